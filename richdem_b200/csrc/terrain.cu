// Seeded synthetic fractal DEM generated directly in HBM (benchmark / test input only).
// Value-noise fractional Brownian motion: octave o interpolates (quintic fade) a hashed integer
// lattice of cell size 2^(k-o); amplitudes fall off as 2^(-0.75 o).  Cell (x, y0+y) only depends
// on its global coordinates and the seed, so row bands generated on different GPUs agree.
#include "common.cuh"

namespace rdb {

namespace {

__device__ __forceinline__ uint32_t hash3(uint32_t x, uint32_t y, uint32_t s) {
  uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u + 0x7F4A7C15u) ^ (s * 0xC2B2AE3Du);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ float lattice(uint32_t ix, uint32_t iy, uint32_t s) {
  return (float)(hash3(ix, iy, s) >> 8) * (1.0f / 16777216.0f);
}

// quintic fade t^3 (t (6 t - 15) + 10), one rounding per operation
__device__ __forceinline__ float fade(float t) {
  const float inner = __fadd_rn(__fmul_rn(t, __fsub_rn(__fmul_rn(t, 6.f), 15.f)), 10.f);
  return __fmul_rn(__fmul_rn(__fmul_rn(t, t), t), inner);
}

__global__ void __launch_bounds__(256) fbm_kernel(float *dem, int W, int H, int y0, uint32_t seed, int octaves,
                                                   int top_log2, float quantum) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  for (int yl = blockIdx.y; yl < H; yl += gridDim.y) {
    const int y = y0 + yl;
    float sum = 0.f, amp = 1.f, norm = 0.f;
    for (int o = 0; o < octaves; o++) {
      const int lg = top_log2 - o;
      if (lg < 1) break;
      const uint32_t cell = 1u << lg;
      const uint32_t ix = (uint32_t)x >> lg, iy = (uint32_t)y >> lg;
      // every operation is an explicitly rounded IEEE single (no FMA contraction), so that the CPU restatement
      // of this generator (oracle/oracle.c: orc_generate_fbm_f32, compiled with -ffp-contract=off) produces the
      // same bits: the reference arm of bench.py runs on exactly the raster the GPU arm runs on
      float tx = __fdiv_rn((float)((uint32_t)x & (cell - 1)), (float)cell);
      float ty = __fdiv_rn((float)((uint32_t)y & (cell - 1)), (float)cell);
      tx = fade(tx);
      ty = fade(ty);
      const uint32_t s = seed * 131u + (uint32_t)o;
      const float v00 = lattice(ix, iy, s), v01 = lattice(ix + 1, iy, s);
      const float v10 = lattice(ix, iy + 1, s), v11 = lattice(ix + 1, iy + 1, s);
      const float top = __fadd_rn(__fmul_rn(v00, __fsub_rn(1.f, tx)), __fmul_rn(v01, tx));
      const float bot = __fadd_rn(__fmul_rn(v10, __fsub_rn(1.f, tx)), __fmul_rn(v11, tx));
      const float v = __fadd_rn(__fmul_rn(top, __fsub_rn(1.f, ty)), __fmul_rn(bot, ty));
      sum = __fadd_rn(sum, __fmul_rn(amp, v));
      norm = __fadd_rn(norm, amp);
      amp = __fmul_rn(amp, 0.5946035575f);  // 2^-0.75
    }
    float z = __fdiv_rn(__fmul_rn(1000.0f, sum), norm);
    if (quantum > 0.f) z = __fmul_rn(rintf(__fdiv_rn(z, quantum)), quantum);
    dem[(size_t)yl * W + x] = z;
  }
}

}  // namespace

void generate_fbm_dev(float *d_dem, int w, int h, int y0, uint32_t seed, int octaves, float quantum) {
  Ctx &c = ctx();
  if (octaves <= 0) octaves = 12;
  const int top_log2 = 12;  // coarsest lattice: 4096 cells
  dim3 blk(256), grd((w + 255) / 256, h < 16384 ? h : 16384);
  fbm_kernel<<<grd, blk, 0, c.stream>>>(d_dem, w, h, y0, seed, octaves, top_log2, quantum);
  RDB_CK(cudaGetLastError());
  count_launch();
}

}  // namespace rdb
