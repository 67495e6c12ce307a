// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  A minimal single-threaded CUDA execution model for the CPU,
// just large enough to run richdem_b200/csrc/*.cu functionally (tests/emu/build_emu.py rewrites the
// sources against this header and links them into tests/_bin/librdb200_emu_test.so).
//
// It exists so that the LOGIC of the shipped kernels (tile worklists, dirty-block lists, union-find,
// dependency counters, band protocols ...) can be checked against the oracle on machines without a
// GPU.  It is not a backend: richdem_b200/_lib.py refuses to load a library that exports
// `rdb200_emulated`, nothing in the package, bench.py or smoke() references it, and it is far too
// slow for anything but toy rasters.  It says nothing about performance, memory ordering or races:
// every CUDA thread is a fiber on ONE OS thread, switched only at barriers / warp collectives /
// grid syncs, so atomics are trivially atomic and interleavings are deterministic.
//
// Model: a launch runs its blocks one after another; the threads of a block are fibers scheduled
// round-robin.  __syncthreads / __syncthreads_count wait for all live threads of the block, warp
// collectives (full masks only) for all live lanes of the warp, cooperative grid.sync() for all
// live threads of the launch (cooperative launches run all their blocks concurrently and must use
// their own shared memory; the emulated device reports RDB_EMU_SMS SMs, default 1).
// TMA / mbarrier inline PTX is not emulated: the sources run with fill_use_tma = 0.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define RDB_EMU 1

// ---- qualifiers ----
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ----
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- driver / runtime types ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 999 };
typedef struct rdb_emu_stream *cudaStream_t;
typedef struct rdb_emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount, cooperativeLaunch;
  size_t totalGlobalMem;
};
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaEnableDefault = 0 };
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
struct alignas(64) CUtensorMap { unsigned long long opaque[16]; };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

cudaError_t cudaMalloc(void **p, size_t n);
template <class T>
static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t n);
template <class T>
static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMallocHost((void **)p, n); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
enum : unsigned { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }  // execution is synchronous
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventDestroy(cudaEvent_t e);
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, int flags, cudaDriverEntryPointQueryResult *q);
template <class F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) {
  *n = 1;
  return cudaSuccess;
}

// ---- execution model ----
namespace rdb_emu {
void launch(dim3 grid, dim3 block, const std::function<void()> &thread_body, bool cooperative = false);
void sync_block();
int sync_block_count(int pred);
void sync_grid();
// warp collectives over the live lanes of the calling thread's warp; `val` travels as 64 bits
uint64_t warp_exchange(uint64_t val, int src_lane_or_neg, unsigned *ballot_out, int pred);
int lane_id();
bool in_kernel();
void yield_now();
void maybe_yield();  // RDB_EMU_CHAOS=<seed>: atomics yield at random so that blocks interleave differently
void asm_stub(const char *text);
// block-local "shared memory": one object per __shared__ declaration and concurrently running block
struct SharedSlot {
  std::vector<void *> per_block;
  size_t size = 0;
};
void *shared_mem(SharedSlot *slot, size_t size, size_t align);

template <class... A, size_t... I>
void call_unpacked(void (*f)(A...), void **args, std::index_sequence<I...>) {
  f(*reinterpret_cast<typename std::remove_reference<A>::type *>(args[I])...);
}
template <class... A>
cudaError_t launch_coop(void (*f)(A...), dim3 grid, dim3 block, void **args, size_t, cudaStream_t) {
  launch(grid, block, [&]() { call_unpacked(f, args, std::index_sequence_for<A...>{}); }, true);
  return cudaSuccess;
}
}  // namespace rdb_emu

static inline void __syncthreads() { rdb_emu::sync_block(); }
static inline int __syncthreads_count(int pred) { return rdb_emu::sync_block_count(pred); }
static inline int __syncthreads_or(int pred) { return rdb_emu::sync_block_count(pred) != 0; }
void rdb_emu_nanosleep();
static inline void __nanosleep(unsigned) { rdb_emu_nanosleep(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {
  unsigned b;
  rdb_emu::warp_exchange(0, -1, &b, 0);
}

static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned b;
  rdb_emu::warp_exchange(0, -1, &b, pred);
  return b;
}
template <class T>
static inline T rdb_emu_shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  unsigned b;
  const uint64_t r = rdb_emu::warp_exchange(bits, src, &b, 0);
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return rdb_emu_shfl(v, src & 31); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return rdb_emu_shfl(v, (rdb_emu::lane_id() ^ m) & 31); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  const int s = rdb_emu::lane_id() + (int)d;
  return rdb_emu_shfl(v, s < 32 ? s : rdb_emu::lane_id());
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  const int s = rdb_emu::lane_id() - (int)d;
  return rdb_emu_shfl(v, s >= 0 ? s : rdb_emu::lane_id());
}

// ---- memory access intrinsics ----
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline T __ldca(const T *p) { return *p; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
template <class T> static inline void __stwt(T *p, T v) { *p = v; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }

// ---- bit casts / integer intrinsics ----
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
// round-to-nearest arithmetic that the compiler may not contract (build with -ffp-contract=off)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }

// ---- atomics (single OS thread: plain read-modify-write) ----
template <class T> struct rdb_emu_id { typedef T type; };
#define RDB_EMU_ARG(T) typename rdb_emu_id<T>::type
template <class T> static inline T atomicAdd(T *p, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T *p, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicMin(T *p, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicExch(T *p, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, RDB_EMU_ARG(T) cmp, RDB_EMU_ARG(T) v) { rdb_emu::maybe_yield(); const T o = *p; if (o == cmp) *p = v; return o; }

// ---- cooperative groups (the subset the sources use) ----
namespace cooperative_groups {
struct grid_group {
  void sync() const { rdb_emu::sync_grid(); }
};
static inline grid_group this_grid() { return grid_group(); }
// a "coalesced group" is whatever subset of a warp happens to be converged; a single lane is a
// legal outcome on hardware too, and it is what a fiber-per-thread model gives
struct coalesced_group {
  unsigned thread_rank() const { return 0; }
  unsigned size() const { return 1; }
  template <class T> T shfl(T v, int) const { return v; }
};
static inline coalesced_group coalesced_threads() { return coalesced_group(); }
}  // namespace cooperative_groups
