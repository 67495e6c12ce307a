// Per-cell flow-metric device functions shared by the materialising kernels (FM_D8 /
// FM_Tarboton -> 9 floats per cell) and the fused accumulation path (compact code per cell).
#pragma once
#include "common.cuh"

namespace rdb {

constexpr float kNoFlowGen = -1.0f;   // reference common/constants.hpp:83
constexpr float kHasFlowGen = 0.0f;   // :84
constexpr float kNoDataGen = -2.0f;   // :85

// Compact per-cell flow code used by the fused accumulation:
//   0        no receiver (no-flow cell, raster-edge cell)
//   1..8     one receiver in D8 direction n, proportion 1
//   16|n     two receivers: n with proportion p1 and nwrap(n+1) with proportion p2 (D-infinity)
//   255      NoData cell
constexpr uint8_t kCodeNoData = 255;
constexpr uint8_t kCodeTwo = 16;

__device__ __forceinline__ int nwrap(int n) { return n == 9 ? 1 : n; }
__device__ __forceinline__ int d8_inverse(int n) { return n == 0 ? 0 : ((n + 3) & 7) + 1; }  // {0,5,6,7,8,1,2,3,4}

// FM_OCallaghan<D8> for one cell (reference flowmet/OCallaghan1984.hpp:37-75).
// returns kCodeNoData, 0 (no flow) or the receiver direction 1..8.
__device__ __forceinline__ int fm_d8_cell(const float *__restrict__ dem, int x, int y, int W, int H,
                                          float nodata) {
  const size_t i = (size_t)y * W + x;
  const float e = __ldg(dem + i);
  if (e == nodata) return kCodeNoData;                          // :37-40
  if (x == 0 || y == 0 || x == W - 1 || y == H - 1) return 0;  // :42-43
  int lowest_n = 0;
  float lowest = 3.402823466e+38f;  // numeric_limits<float>::max(), :48
#pragma unroll
  for (int n = 1; n <= 8; n++) {
    const float ne = __ldg(dem + (size_t)(y + d8dy(n)) * W + (x + d8dx(n)));
    if (ne == nodata) continue;  // :53-54
    if (ne >= e) continue;       // :58-59
    if (ne < lowest) {           // :61-64  strict: the first minimum wins
      lowest = ne;
      lowest_n = n;
    }
  }
  return lowest_n;
}

// FM_OCallaghan<D4> for one cell (reference flowmet/OCallaghan1984.hpp:37-75 with the D4 offset tables
// common/constants.hpp:53-54: 1 = W, 2 = N, 3 = E, 4 = S).  Returns kCodeNoData, 0 or the D4 index 1..4 --
// the reference stores the proportion in THAT slot of the 9-slot cell (and its generic accumulation then
// reads slots as D8 directions, flow_accumulation_generic.hpp:53-56); both are reproduced as they are.
__device__ __forceinline__ int fm_d4_cell(const float *__restrict__ dem, int x, int y, int W, int H, float nodata) {
  const size_t i = (size_t)y * W + x;
  const float e = __ldg(dem + i);
  if (e == nodata) return kCodeNoData;
  if (x == 0 || y == 0 || x == W - 1 || y == H - 1) return 0;
  int lowest_n = 0;
  float lowest = 3.402823466e+38f;
#pragma unroll
  for (int n = 1; n <= 4; n++) {
    const int dx = n == 1 ? -1 : (n == 3 ? 1 : 0), dy = n == 2 ? -1 : (n == 4 ? 1 : 0);
    const float ne = __ldg(dem + (size_t)(y + dy) * W + (x + dx));
    if (ne == nodata) continue;
    if (ne >= e) continue;
    if (ne < lowest) {
      lowest = ne;
      lowest_n = n;
    }
  }
  return lowest_n;
}

// FM_Holmgren (HOLMGREN = true; reference flowmet/Holmgren1994.hpp:33-80; FM_Quinn is xparam = 1,
// Quinn1991.hpp:15) and FM_Freeman (HOLMGREN = false; flowmet/Freeman1991.hpp:28-77) for one cell.
// p[0..8] arrives filled with NO_FLOW_GEN.  The reference's mixed precision is kept step by step:
// rise is a float difference widened to double, the gradient and the power are double; Holmgren rounds
// each power to float BEFORE summing (it sums props(x,y,n)), Freeman sums the unrounded doubles;
// normalisation multiplies the stored float by the double 1/C and rounds once.  pow(g, 1.0) is g in
// glibc, so the exponent-1 case takes the value itself instead of the device pow (<= 2 ulp).
template <bool HOLMGREN>
__device__ __forceinline__ void fm_mfd_cell(const float *__restrict__ dem, int x, int y, int W, int H, float nodata,
                                            double xparam, float (&p)[9]) {
  const size_t i = (size_t)y * W + x;
  const float e = __ldg(dem + i);
  if (e == nodata) {
    p[0] = kNoDataGen;
    return;
  }
  if (x == 0 || y == 0 || x == W - 1 || y == H - 1) return;
  double C = 0;
#pragma unroll
  for (int n = 1; n <= 8; n++) {
    const float ne = __ldg(dem + (size_t)(y + d8dy(n)) * W + (x + d8dx(n)));
    if (ne == nodata) continue;
    if (ne < e) {
      const double rise = (double)__fsub_rn(e, ne);
      const double run = (n & 1) ? 1.0 : 1.414213562373095048801688724209698078569671875376948;
      double g = __ddiv_rn(rise, run);
      if (HOLMGREN) g = __dmul_rn(g, (n & 1) ? 0.5 : 0.354);
      const double cval = xparam == 1.0 ? g : pow(g, xparam);
      p[n] = (float)cval;
      C = __dadd_rn(C, HOLMGREN ? (double)p[n] : cval);
    }
  }
  if (C > 0) {
    p[0] = kHasFlowGen;
    C = __ddiv_rn(1.0, C);
#pragma unroll
    for (int n = 1; n <= 8; n++) p[n] = p[n] > 0 ? (float)__dmul_rn((double)p[n], C) : 0.0f;
  }
}

// FM_Tarboton for one cell (reference flowmet/Tarboton1997.hpp:62-141).
// returns kCodeNoData, 0 (no flow) or nmax in 1..8 with *rmax_out = rmax after the facet-parity
// flip (:121-126).  All double arithmetic uses explicit round-to-nearest intrinsics so that no
// FMA contraction happens (the CPU reference is compiled without it).
__device__ __forceinline__ int fm_tarboton_cell(const float *__restrict__ dem, int x, int y, int W, int H,
                                                float nodata, float *rmax_out) {
  const size_t i = (size_t)y * W + x;
  const float e0f = __ldg(dem + i);
  if (e0f == nodata) return kCodeNoData;
  if (x == 0 || y == 0 || x == W - 1 || y == H - 1) return 0;
  // facet tables, :48-53 (remapped facets 1..8)
  //  dy_e1 = {0, 0,-1,-1, 0, 0, 1, 1, 0}   dx_e1 = {0,-1, 0, 0, 1, 1, 0, 0,-1}
  //  dy_e2 = {0,-1,-1,-1,-1, 1, 1, 1, 1}   dx_e2 = {0,-1,-1, 1, 1, 1, 1,-1,-1}
  //  af    = {0,-1, 1,-1, 1,-1, 1,-1, 1}
  const float dang = 0.78539818525314331f;  // float(atan2(1,1)), :29
  const double dangd = (double)dang;
  const double e0 = (double)e0f;
  int nmax = -1, bmax = 0;
  double smax = 0, s1max = 0, s2max = 0;
  float rmax = 0;
#pragma unroll
  for (int n = 1; n <= 8; n++) {
    // e1 is the cardinal neighbour of the facet, e2 the diagonal one
    const int dx1 = (n == 1 || n == 8) ? -1 : ((n == 4 || n == 5) ? 1 : 0);
    const int dy1 = (n == 2 || n == 3) ? -1 : ((n == 6 || n == 7) ? 1 : 0);
    const int dx2 = (n == 1 || n == 2 || n == 7 || n == 8) ? -1 : 1;
    const int dy2 = (n <= 4) ? -1 : 1;
    // interior cell: both neighbours are in the grid (:76-83 only filter NoData here)
    const float e1f = __ldg(dem + (size_t)(y + dy1) * W + (x + dx1));
    const float e2f = __ldg(dem + (size_t)(y + dy2) * W + (x + dx2));
    if (e1f == nodata || e2f == nodata) continue;
    const double e1 = (double)e1f, e2 = (double)e2f;
    const double s1 = __dsub_rn(e0, e1);  // (e0-e1)/d1, d1 = 1
    const double s2 = __dsub_rn(e1, e2);
    // r = atan2(s2, s1) only matters through the two threshold tests below and, for the steepest facet, as rmax.
    // The tests are decided from the signs and the ratio s2 / s1 (atan2 is monotone in it) whenever the ratio is
    // clear of tan(1e-7) and tan(dang - 1e-7) by a relative 1e-9 -- ten million ulps, far beyond any atan2's error --
    // and by the function itself otherwise; the one atan2 a cell needs is taken at the end.  (8 double atan2 per cell
    // were 77 of the 575 ms of FA_Dinf at 32768^2.)
    int branch;  // 0: r < 1e-7   1: r > dang - 1e-7   2: in between
    if (s2 < 0.0 || (s2 == 0.0 && s1 >= 0.0)) {
      branch = 0;  // angles in (-pi, 0], and atan2(0, 0) = 0
    } else if (s1 <= 0.0) {
      branch = 1;  // s2 > 0 (or s2 == 0 with s1 < 0): angles in [pi/2, pi]
    } else {       // first quadrant, both positive
      const double tlo = 1.0000000000000033e-07, thi = 0.9999998437114023;  // tan(1e-7), tan(dang - 1e-7)
      const double eps = 1e-9;
      if (s2 < s1 * (tlo * (1.0 - eps))) branch = 0;
      else if (s2 > s1 * (thi * (1.0 + eps))) branch = 1;
      else if (s2 > s1 * (tlo * (1.0 + eps)) && s2 < s1 * (thi * (1.0 - eps))) branch = 2;
      else {
        const double ra = atan2(s2, s1);
        branch = ra < 1e-7 ? 0 : (ra > __dsub_rn(dangd, 1e-7) ? 1 : 2);
      }
    }
    double s;
    if (branch == 0) {  // :99-101
      s = s1;
    } else if (branch == 1) {  // :102-104
      s = __ddiv_rn(__dsub_rn(e0, e2), 1.4142135623730951);  // sqrt(d1*d1+d2*d2) = sqrt(2.0)
    } else {
      s = __dsqrt_rn(__dadd_rn(__dmul_rn(s1, s1), __dmul_rn(s2, s2)));  // :106
    }
    if (s > smax) {  // :109-113
      smax = s;
      nmax = n;
      bmax = branch;
      s1max = s1;
      s2max = s2;
    }
  }
  if (nmax == -1) return 0;
  rmax = bmax == 0 ? 0.0f : (bmax == 1 ? (float)dangd : (float)atan2(s2max, s1max));
  const bool af_pos = (nmax & 1) == 0;  // af[n] == +1 for even n
  if (af_pos && rmax == 0.0f) rmax = dang;
  else if (af_pos && rmax == dang) rmax = 0.0f;
  else if (af_pos) rmax = (float)__dsub_rn(0.78539816339744830962, (double)rmax);  // M_PI/4 - rmax
  *rmax_out = rmax;
  return nmax;
}

// proportions written by FM_Tarboton, :134-141
__device__ __forceinline__ void tarboton_props(float rmax, float *p1, float *p2) {
  const double q = __ddiv_rn((double)rmax, 0.78539816339744830962);
  *p1 = (float)q;
  *p2 = (float)__dsub_rn(1.0, q);
}
constexpr float kDang = 0.78539818525314331f;

}  // namespace rdb
