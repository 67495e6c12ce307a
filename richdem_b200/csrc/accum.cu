// Flow accumulation as a topological wavefront over the flow graph:
//
//     A(c) = w(c) + sum over donors u -> c of p(u,c) * A(u)
//
// (reference methods/flow_accumulation_generic.hpp:33-100 and methods/d8_methods.hpp:47-139).
// Every cell carries a remaining-dependency counter; cells with no donors are the sources.  One
// thread starts at each source and pushes its accumulated value downstream with atomicAdd, then
// decrements the receiver's counter with an atomic; the thread that brings a counter to zero is
// the last donor to arrive, so it owns the receiver and keeps walking.  Receivers that become
// ready while the walker already has a successor (two-receiver D-infinity / multi-receiver
// proportions) are appended to a compacted frontier array that seeds the next launch.
// D8 with unit weights only ever adds integers < 2^53 in double, so the result is exact and
// independent of the order in which atomics land (bit-identical to the serial reference).
#include "flowmet.cuh"

namespace rdb {

namespace {

constexpr uint32_t kDepsMask = 0xFFu;
constexpr uint32_t kSrcFlag = 0x80000000u;

// ---- K1: dem -> compact flow code (+ rmax for D-infinity), weights/NoData initialisation ------
template <bool DINF>
__global__ void __launch_bounds__(256) flow_code_kernel(const float *__restrict__ dem, uint8_t *__restrict__ code,
                                                         float *__restrict__ rmaxArr, double *__restrict__ accum,
                                                         int W, int H, float nodata, int ones) {
  const size_t n = (size_t)W * H;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
  int cd;
  if (DINF) {
    float rmax = 0;
    const int nm = fm_tarboton_cell(dem, x, y, W, H, nodata, &rmax);
    if (nm == kCodeNoData || nm == 0) {
      cd = nm;
    } else if (rmax == 0.0f) {
      cd = nm;  // single receiver, proportion 1 (Tarboton1997.hpp:134-135)
    } else if (rmax == kDang) {
      cd = nwrap(nm + 1);  // :136-137
    } else {
      cd = kCodeTwo | nm;  // :138-141
      rmaxArr[i] = rmax;
    }
  } else {
    cd = fm_d8_cell(dem, x, y, W, H, nodata);
  }
  code[i] = (uint8_t)cd;
  if (cd == kCodeNoData) accum[i] = -1.0;  // flow_accumulation_generic.hpp:95-97 (never touched again)
  else if (ones) accum[i] = 1.0;
}

// ---- K2: dependency counters by gathering over the 8 neighbours' codes; marks sources ----------
__global__ void __launch_bounds__(256) deps_gather_kernel(const uint8_t *__restrict__ code, uint32_t *__restrict__ st,
                                                           int W, int H) {
  const size_t n = (size_t)W * H;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
  if (code[i] == kCodeNoData) {
    st[i] = 0;
    return;
  }
  uint32_t deps = 0;
#pragma unroll
  for (int k = 1; k <= 8; k++) {
    const int nx = x + d8dx(k), ny = y + d8dy(k);
    if (nx < 0 || ny < 0 || nx >= W || ny >= H) continue;
    const int cn = code[(size_t)ny * W + nx];
    if (cn == 0 || cn == kCodeNoData) continue;
    const int inv = d8_inverse(k);  // direction from that neighbour to me
    const int first = cn & 15;
    if (first == inv) deps++;
    else if ((cn & kCodeTwo) && nwrap(first + 1) == inv) deps++;
  }
  st[i] = deps | (deps == 0 ? kSrcFlag : 0u);
}

// ---- proportions path: scatter dependency counts, then mark sources ----------------------------
__global__ void __launch_bounds__(256) deps_scatter_props_kernel(const float *__restrict__ props, uint32_t *st, int W,
                                                                  int H) {
  const size_t n = (size_t)W * H;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
  if (x == 0 || y == 0 || x == W - 1 || y == H - 1) return;  // generic.hpp:47-48 interior only
  const float *p = props + 9 * i;
  if (p[0] == kNoDataGen) return;
#pragma unroll
  for (int k = 1; k <= 8; k++)
    if (p[k] > 0) atomicAdd(&st[i + (ptrdiff_t)d8dy(k) * W + d8dx(k)], 1u);
}

__global__ void __launch_bounds__(256) mark_sources_props_kernel(const float *__restrict__ props, uint32_t *st,
                                                                  double *accum, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (props[9 * i] == kNoDataGen) {
    st[i] = 0;
    accum[i] = -1.0;  // generic.hpp:95-97
    return;
  }
  const uint32_t d = st[i];
  if ((d & kDepsMask) == 0) st[i] = kSrcFlag;
}

// ---- the walk ----------------------------------------------------------------------------------
// MODE 0: compact code, at most one receiver (D8, direction grids)
// MODE 1: compact code, up to two receivers (D-infinity)
// MODE 2: 9-float proportions, up to eight receivers
template <class A>
struct WalkArgs {
  const uint8_t *code;
  const float *rmaxArr;
  const float *props;
  A *accum;
  uint32_t *st;
  const int *frontier;  // nullptr: round 0, every cell whose st has kSrcFlag
  int nfrontier;        // round 0: number of cells
  int *next_frontier;
  int *next_count;
  int W, H;
};

template <class A>
__device__ __forceinline__ A ld_acc(const A *p) {
  return __ldcg(p);
}

template <int MODE, bool CHECK, class A>
__global__ void __launch_bounds__(256) accum_walk_kernel(const WalkArgs<A> a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.nfrontier) return;
  int c;
  if (a.frontier) {
    c = a.frontier[t];
  } else {
    c = t;
    if (!(a.st[c] & kSrcFlag)) return;
  }
  const int W = a.W;
  A acc = ld_acc(a.accum + c);
  for (;;) {
    int next = -1;
    if (MODE == 0) {
      const int cd = a.code[c];
      if (cd == 0 || cd == kCodeNoData) break;
      const int dx = d8dx(cd), dy = d8dy(cd);
      if (CHECK) {  // direction grids may point off the raster (d8_methods.hpp:121-122)
        const int y = c / W, x = c - y * W;
        const int nx = x + dx, ny = y + dy;
        if (nx < 0 || ny < 0 || nx >= W || ny >= a.H) break;
      }
      const int r = c + dy * W + dx;
      if (a.code[r] == kCodeNoData) break;  // flow into NoData is dropped
      atomicAdd(a.accum + r, acc);
      __threadfence();
      const uint32_t old = atomicSub(a.st + r, 1u);
      if ((old & kDepsMask) == 1u) next = r;
    } else if (MODE == 1) {
      const int cd = a.code[c];
      if (cd == 0 || cd == kCodeNoData) break;
      const int n1 = cd & 15;
      const int r1 = c + d8dy(n1) * W + d8dx(n1);
      int r2 = -1;
      if (cd & kCodeTwo) {
        const int n2 = nwrap(n1 + 1);
        r2 = c + d8dy(n2) * W + d8dx(n2);
        float p1, p2;
        tarboton_props(a.rmaxArr[c], &p1, &p2);
        // generic.hpp:87  accum(ni) += props(ci,n)*c_accum  (float * double)
        if (p1 > 0) atomicAdd(a.accum + r1, (A)((double)p1 * (double)acc));
        if (p2 > 0) atomicAdd(a.accum + r2, (A)((double)p2 * (double)acc));
        __threadfence();
        if (p1 > 0) {
          const uint32_t o1 = atomicSub(a.st + r1, 1u);
          if ((o1 & kDepsMask) == 1u) next = r1;
        }
        if (p2 > 0) {
          const uint32_t o2 = atomicSub(a.st + r2, 1u);
          if ((o2 & kDepsMask) == 1u) {
            if (next < 0) next = r2;
            else a.next_frontier[atomicAdd(a.next_count, 1)] = r2;
          }
        }
      } else {
        atomicAdd(a.accum + r1, acc);
        __threadfence();
        const uint32_t o1 = atomicSub(a.st + r1, 1u);
        if ((o1 & kDepsMask) == 1u) next = r1;
      }
    } else {
      const int y = c / W, x = c - y * W;
      if (x == 0 || y == 0 || x == W - 1 || y == a.H - 1) break;  // edge cells carry no flow
      const float *p = a.props + (size_t)9 * c;
      uint32_t sent = 0;
#pragma unroll
      for (int k = 1; k <= 8; k++) {
        const float pk = p[k];
        if (pk <= 0) continue;  // generic.hpp:82-83
        const int r = c + d8dy(k) * W + d8dx(k);
        if (a.props[(size_t)9 * r] == kNoDataGen) continue;  // :85-86
        atomicAdd(a.accum + r, (A)((double)pk * (double)acc));
        sent |= 1u << k;
      }
      if (!sent) break;
      __threadfence();
#pragma unroll
      for (int k = 1; k <= 8; k++) {
        if (!(sent & (1u << k))) continue;
        const int r = c + d8dy(k) * W + d8dx(k);
        const uint32_t o = atomicSub(a.st + r, 1u);
        if ((o & kDepsMask) == 1u) {
          if (next < 0) next = r;
          else a.next_frontier[atomicAdd(a.next_count, 1)] = r;
        }
      }
    }
    if (next < 0) break;
    __threadfence();
    c = next;
    acc = ld_acc(a.accum + c);
  }
}

template <int MODE, bool CHECK, class A>
void run_walk(WalkArgs<A> a, size_t ncells) {
  Ctx &c = ctx();
  DevBuf<int> fr0, fr1;
  DevBuf<int> cnt(2);
  int *hcnt = (int *)c.pinned;
  if (MODE != 0) {
    fr0.alloc(ncells);
    fr1.alloc(ncells);
  }
  RDB_CK(cudaMemsetAsync(cnt.p, 0, 2 * sizeof(int), c.stream));
  a.frontier = nullptr;
  a.nfrontier = (int)ncells;
  a.next_frontier = fr0.p;
  a.next_count = cnt.p;
  int rounds = 0;
  KernelTimer kt;
  for (;;) {
    const unsigned blocks = (unsigned)(((size_t)a.nfrontier + 255) / 256);
    accum_walk_kernel<MODE, CHECK, A><<<blocks, 256, 0, c.stream>>>(a);
    RDB_CK(cudaGetLastError());
    count_launch();
    rounds++;
    if (MODE == 0) break;
    RDB_CK(cudaMemcpyAsync(hcnt, a.next_count, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    RDB_CK(cudaStreamSynchronize(c.stream));
    const int nn = *hcnt;
    if (nn == 0) break;
    // swap frontiers
    a.frontier = a.next_frontier;
    a.nfrontier = nn;
    a.next_frontier = (a.frontier == fr0.p) ? fr1.p : fr0.p;
    a.next_count = (a.next_count == cnt.p) ? cnt.p + 1 : cnt.p;
    RDB_CK(cudaMemsetAsync(a.next_count, 0, sizeof(int), c.stream));
  }
  kt.stop_async();
  RDB_CK(cudaStreamSynchronize(c.stream));
  c.stats.ms_main_kernel += kt.ms();
  c.stats.accum_rounds = rounds;
}

__global__ void area_init_kernel(const uint8_t *__restrict__ dirs, int32_t *__restrict__ area, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) area[i] = (dirs[i] == kCodeNoData) ? -1 : 1;  // d8_methods.hpp:71-74, :111
}

__global__ void sanitize_dirs_kernel(const uint8_t *__restrict__ dirs, uint8_t *__restrict__ code, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint8_t d = dirs[i];
    code[i] = (d <= 8 || d == kCodeNoData) ? d : 0;  // anything else: treated as no flow
  }
}

}  // namespace

// FA_D8 / FA_Tarboton fused (reference methods/flow_accumulation.hpp:27,16): no 36 B/cell props
void fa_fused_dev(const float *d_dem, double *d_accum, int w, int h, float nodata, bool ones, bool dinf) {
  Ctx &c = ctx();
  const size_t n = (size_t)w * h;
  DevBuf<uint8_t> code(n);
  DevBuf<uint32_t> st(n);
  DevBuf<float> rmax;
  if (dinf) rmax.alloc(n);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dinf)
    flow_code_kernel<true><<<blocks, 256, 0, c.stream>>>(d_dem, code.p, rmax.p, d_accum, w, h, nodata, ones ? 1 : 0);
  else
    flow_code_kernel<false><<<blocks, 256, 0, c.stream>>>(d_dem, code.p, nullptr, d_accum, w, h, nodata, ones ? 1 : 0);
  RDB_CK(cudaGetLastError());
  deps_gather_kernel<<<blocks, 256, 0, c.stream>>>(code.p, st.p, w, h);
  RDB_CK(cudaGetLastError());
  count_launch(2);
  WalkArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.code = code.p;
  a.rmaxArr = rmax.p;
  a.accum = d_accum;
  a.st = st.p;
  a.W = w;
  a.H = h;
  if (dinf) run_walk<1, false, double>(a, n);
  else run_walk<0, false, double>(a, n);
}

// FlowAccumulation(props, accum) (reference methods/flow_accumulation_generic.hpp:33-100)
void flow_accumulation_props_dev(const float *d_props, double *d_accum, int w, int h) {
  Ctx &c = ctx();
  const size_t n = (size_t)w * h;
  DevBuf<uint32_t> st(n);
  RDB_CK(cudaMemsetAsync(st.p, 0, n * sizeof(uint32_t), c.stream));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  deps_scatter_props_kernel<<<blocks, 256, 0, c.stream>>>(d_props, st.p, w, h);
  RDB_CK(cudaGetLastError());
  mark_sources_props_kernel<<<blocks, 256, 0, c.stream>>>(d_props, st.p, d_accum, n);
  RDB_CK(cudaGetLastError());
  count_launch(2);
  WalkArgs<double> a;
  memset(&a, 0, sizeof(a));
  a.props = d_props;
  a.accum = d_accum;
  a.st = st.p;
  a.W = w;
  a.H = h;
  run_walk<2, false, double>(a, n);
}

// d8_flow_accum(dirs, area) (reference methods/d8_methods.hpp:47-139)
void d8_flow_accum_dev(const uint8_t *d_dirs, int32_t *d_area, int w, int h) {
  Ctx &c = ctx();
  const size_t n = (size_t)w * h;
  DevBuf<uint8_t> code(n);
  DevBuf<uint32_t> st(n);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  sanitize_dirs_kernel<<<blocks, 256, 0, c.stream>>>(d_dirs, code.p, n);
  area_init_kernel<<<blocks, 256, 0, c.stream>>>(d_dirs, d_area, n);
  deps_gather_kernel<<<blocks, 256, 0, c.stream>>>(code.p, st.p, w, h);
  RDB_CK(cudaGetLastError());
  count_launch(3);
  WalkArgs<int32_t> a;
  memset(&a, 0, sizeof(a));
  a.code = code.p;
  a.accum = d_area;
  a.st = st.p;
  a.W = w;
  a.H = h;
  run_walk<0, true, int32_t>(a, n);
}

}  // namespace rdb
