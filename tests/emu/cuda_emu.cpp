// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).  Fiber scheduler + host stand-ins for the
// handful of CUDA runtime calls the sources make.  x86-64 SysV only.
#include "cuda_emu.h"

#include <sys/mman.h>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

extern "C" int rdb200_emulated(void) { return 1; }  // marker: the product loader rejects this library
// scheduler passes over all launches so far: every pass advances each live thread to its next barrier, so
// the count is a crude, hardware-free proxy for the critical path of a schedule (tools only)
static unsigned long long g_total_passes = 0;
extern "C" unsigned long long rdb_emu_passes(void) { return g_total_passes; }

// ---- context switch: callee-saved registers + stack pointer ----
extern "C" void rdb_emu_switch(void **save_sp, void *load_sp);
asm(R"(
    .text
    .globl rdb_emu_switch
    .type rdb_emu_switch, @function
rdb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size rdb_emu_switch, .-rdb_emu_switch
)");

namespace rdb_emu {

namespace {

constexpr size_t kStack = 256 * 1024;

struct Warp {
  int live = 0;       // lanes that have not exited
  int arrived = 0;    // lanes waiting in the current collective
  unsigned gen = 0;   // completed collectives
  uint64_t vals[2][32];
  unsigned ballot[2] = {0, 0};
  unsigned pending_ballot = 0;
};

struct Block {
  int live = 0;
  int arrived = 0;
  unsigned gen = 0;
  int count[2] = {0, 0};
  int pending_count = 0;
  std::vector<Warp> warps;
};

struct Fiber {
  int slot = 0;    // index of its block among the blocks running concurrently
  void *sp = nullptr;
  bool done = false;
  uint3 tid{0, 0, 0}, bid{0, 0, 0};
  int lin = 0;     // linear thread index in its block
  Block *blk = nullptr;
};

struct Launch {
  std::vector<Fiber> fibers;
  std::vector<Block> blocks;
  const std::function<void()> *body = nullptr;
  int grid_live = 0, grid_arrived = 0;
  unsigned grid_gen = 0;
  bool cooperative = false;
};

Launch *g_launch = nullptr;
Fiber *g_cur = nullptr;
void *g_sched_sp = nullptr;
char *g_stacks = nullptr;
size_t g_nstacks = 0;

unsigned long long g_progress = 0;  // barrier releases + fiber exits (deadlock detection)

void yield() { rdb_emu_switch(&g_cur->sp, g_sched_sp); }
}  // namespace
bool in_kernel() { return g_cur != nullptr; }
void yield_now() { yield(); }
namespace {

void release_block_if_complete(Block *b) {
  if (b->live > 0 && b->arrived == b->live) {
    b->count[(b->gen + 1) & 1] = b->pending_count;
    b->pending_count = 0;
    b->arrived = 0;
    b->gen++;
    g_progress++;
  }
}
void release_warp_if_complete(Warp *w) {
  if (w->live > 0 && w->arrived == w->live) {
    w->ballot[(w->gen + 1) & 1] = w->pending_ballot;
    w->pending_ballot = 0;
    w->arrived = 0;
    w->gen++;
    g_progress++;
  }
}
void release_grid_if_complete(Launch *l) {
  if (l->grid_live > 0 && l->grid_arrived == l->grid_live) {
    l->grid_arrived = 0;
    l->grid_gen++;
    g_progress++;
  }
}

extern "C" void rdb_emu_fiber_main() {
  Fiber *f = g_cur;
  (*g_launch->body)();
  f->done = true;
  g_progress++;
  Block *b = f->blk;
  b->live--;
  release_block_if_complete(b);
  Warp *w = &b->warps[f->lin >> 5];
  w->live--;
  release_warp_if_complete(w);
  g_launch->grid_live--;
  release_grid_if_complete(g_launch);
  yield();
  fprintf(stderr, "rdb_emu: resumed a finished fiber\n");
  abort();
}

void prepare(Fiber *f, char *stack_top) {
  // initial frame: six callee-saved registers, then the "return address" = fiber entry.
  // After `ret` the entry sees rsp % 16 == 8, as after a call.
  uintptr_t top = ((uintptr_t)stack_top) & ~(uintptr_t)15;
  void **sp = (void **)(top - 8);
  *--sp = (void *)&rdb_emu_fiber_main;
  for (int i = 0; i < 6; i++) *--sp = nullptr;
  f->sp = sp;
}

void ensure_stacks(size_t n) {
  if (n <= g_nstacks) return;
  if (g_stacks) munmap(g_stacks, g_nstacks * kStack);
  g_stacks = (char *)mmap(nullptr, n * kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (g_stacks == (char *)MAP_FAILED) {
    fprintf(stderr, "rdb_emu: cannot map %zu fiber stacks\n", n);
    abort();
  }
  g_nstacks = n;
}

void run_fibers(Launch &l, size_t first, size_t count) {
  size_t remaining = count;
  size_t idle_passes = 0;
  while (remaining) {
    g_total_passes++;
    const unsigned long long before = g_progress;
    for (size_t i = first; i < first + count; i++) {
      Fiber *f = &l.fibers[i];
      if (f->done) continue;
      g_cur = f;
      threadIdx = f->tid;
      blockIdx = f->bid;
      rdb_emu_switch(&g_sched_sp, f->sp);
      if (f->done) remaining--;
    }
    // every live fiber was resumed once; if none finished and no barrier completed, they are all
    // parked at barriers that can never complete
    // (spin-waits with __nanosleep may legitimately take a few passes without a barrier completing)
    idle_passes = (remaining && g_progress == before) ? idle_passes + 1 : 0;
    if (idle_passes > 20000) {
      fprintf(stderr, "rdb_emu: deadlock -- %zu threads made no progress for 20000 scheduler passes\n", remaining);
      abort();
    }
  }
}

}  // namespace

void launch(dim3 grid, dim3 block, const std::function<void()> &body, bool cooperative) {
  if (g_launch) {
    fprintf(stderr, "rdb_emu: nested launch\n");
    abort();
  }
  const size_t nthreads = (size_t)block.x * block.y * block.z;
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  if (nthreads == 0 || nblocks == 0) return;
  if (nthreads > 1024) {
    fprintf(stderr, "rdb_emu: %zu threads per block\n", nthreads);
    abort();
  }
  static const bool trace = getenv("RDB_EMU_TRACE") != nullptr;
  if (trace) fprintf(stderr, "rdb_emu: launch grid=(%u,%u,%u) block=(%u,%u,%u)%s\n", grid.x, grid.y, grid.z, block.x, block.y,
                     block.z, cooperative ? " cooperative" : "");
  Launch l;
  l.body = &body;
  l.cooperative = cooperative;
  g_launch = &l;
  blockDim = block;
  gridDim = grid;
  // RDB_EMU_CONCURRENT=1: also ordinary launches of up to 2048 blocks run their blocks side by side (for the
  // pass-count comparison of schedules; the default runs them one after another, which is faster)
  static const bool all_concurrent = getenv("RDB_EMU_CONCURRENT") != nullptr;
  const size_t concurrent_blocks = (cooperative || (all_concurrent && nblocks <= 2048)) ? nblocks : 1;
  ensure_stacks(concurrent_blocks * nthreads);
  l.fibers.resize(concurrent_blocks * nthreads);
  l.blocks.resize(concurrent_blocks);
  const size_t nwarps = (nthreads + 31) / 32;
  for (size_t b0 = 0; b0 < nblocks; b0 += concurrent_blocks) {
    l.grid_live = (int)(concurrent_blocks * nthreads);
    l.grid_arrived = 0;
    for (size_t k = 0; k < concurrent_blocks; k++) {
      const size_t b = b0 + k;
      Block &bs = l.blocks[k];
      bs = Block();
      bs.live = (int)nthreads;
      bs.warps.assign(nwarps, Warp());
      for (size_t w = 0; w < nwarps; w++) bs.warps[w].live = (int)std::min<size_t>(32, nthreads - 32 * w);
      uint3 bid;
      bid.x = (unsigned)(b % grid.x);
      bid.y = (unsigned)((b / grid.x) % grid.y);
      bid.z = (unsigned)(b / ((size_t)grid.x * grid.y));
      for (size_t t = 0; t < nthreads; t++) {
        Fiber &f = l.fibers[k * nthreads + t];
        f = Fiber();
        f.bid = bid;
        f.tid.x = (unsigned)(t % block.x);
        f.tid.y = (unsigned)((t / block.x) % block.y);
        f.tid.z = (unsigned)(t / ((size_t)block.x * block.y));
        f.lin = (int)t;
        f.slot = (int)k;
        f.blk = &bs;
        prepare(&f, g_stacks + (k * nthreads + t + 1) * kStack);
      }
    }
    run_fibers(l, 0, concurrent_blocks * nthreads);
  }
  if (trace) fprintf(stderr, "rdb_emu:   ... %llu scheduler passes so far\n", g_total_passes);
  g_launch = nullptr;
  g_cur = nullptr;
}

static uint64_t g_chaos = 0;
static bool g_chaos_init = false;
void maybe_yield() {
  if (!g_chaos_init) {
    g_chaos_init = true;
    const char *s = getenv("RDB_EMU_CHAOS");
    g_chaos = s ? (uint64_t)strtoull(s, nullptr, 10) * 0x9E3779B97F4A7C15ull + 1 : 0;
  }
  if (!g_chaos || !g_cur) return;
  g_chaos ^= g_chaos << 13;
  g_chaos ^= g_chaos >> 7;
  g_chaos ^= g_chaos << 17;
  if ((g_chaos & 7) == 0) yield();
}

void sync_block() {
  Block *b = g_cur->blk;
  const unsigned gen = b->gen;
  b->arrived++;
  release_block_if_complete(b);
  while (b->gen == gen) yield();
}

int sync_block_count(int pred) {
  Block *b = g_cur->blk;
  const unsigned gen = b->gen;
  b->pending_count += pred ? 1 : 0;
  b->arrived++;
  release_block_if_complete(b);
  while (b->gen == gen) yield();
  return b->count[(gen + 1) & 1];
}

void sync_grid() {
  Launch *l = g_launch;
  if (!l->cooperative) {
    fprintf(stderr, "rdb_emu: grid.sync() in a non-cooperative launch\n");
    abort();
  }
  const unsigned gen = l->grid_gen;
  l->grid_arrived++;
  release_grid_if_complete(l);
  while (l->grid_gen == gen) yield();
}

int lane_id() { return g_cur->lin & 31; }

uint64_t warp_exchange(uint64_t val, int src, unsigned *ballot_out, int pred) {
  Warp *w = &g_cur->blk->warps[g_cur->lin >> 5];
  const int lane = g_cur->lin & 31;
  const unsigned gen = w->gen;
  const int slot = (gen + 1) & 1;
  w->vals[slot][lane] = val;
  if (pred) w->pending_ballot |= 1u << lane;
  w->arrived++;
  release_warp_if_complete(w);
  while (w->gen == gen) yield();
  *ballot_out = w->ballot[slot];
  // lanes that exited before the collective contribute nothing: reading them returns the caller's value
  if (src < 0) return val;
  return w->vals[slot][src];
}

void *shared_mem(SharedSlot *slot, size_t size, size_t align) {
  const size_t k = (size_t)g_cur->slot;
  if (slot->size != size) {  // first use (or a different template instantiation sharing the slot: never happens)
    for (void *p : slot->per_block) free(p);
    slot->per_block.clear();
    slot->size = size;
  }
  if (slot->per_block.size() <= k) slot->per_block.resize(k + 1, nullptr);
  if (!slot->per_block[k]) {
    if (align < 16) align = 16;
    slot->per_block[k] = aligned_alloc(align, (size + align - 1) / align * align);
    memset(slot->per_block[k], 0xA5, size);  // shared memory starts as garbage
  }
  return slot->per_block[k];
}

void asm_stub(const char *text) {
  if (strstr(text, "cp.async.bulk") || strstr(text, "try_wait") || strstr(text, "expect_tx")) {
    fprintf(stderr, "rdb_emu: TMA / mbarrier PTX is not emulated (run with fill_use_tma = 0): %s\n", text);
    abort();
  }
}

}  // namespace rdb_emu

void rdb_emu_nanosleep() {
  if (rdb_emu::in_kernel()) rdb_emu::yield_now();
}

// ---- runtime stand-ins ----
struct rdb_emu_stream { int dummy; };
struct rdb_emu_event { std::chrono::steady_clock::time_point t; };

cudaError_t cudaMalloc(void **p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  return *p ? cudaSuccess : cudaErrorEmu;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new rdb_emu_stream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new rdb_emu_event(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorEmu; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof(*p));
  snprintf(p->name, sizeof(p->name), "rdb_emu (CPU fibers, test only)");
  p->major = 10;
  p->minor = 0;
  const char *sms = getenv("RDB_EMU_SMS");  // > 1: cooperative kernels run several blocks side by side
  p->multiProcessorCount = sms && atoi(sms) > 0 ? atoi(sms) : 1;
  p->cooperativeLaunch = 1;
  p->totalGlobalMem = (size_t)8 << 30;
  return cudaSuccess;
}
static CUresult emu_encode_tiled(CUtensorMap *m, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  memset(m, 0, sizeof(*m));
  return CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, int, cudaDriverEntryPointQueryResult *q) {
  if (strcmp(name, "cuTensorMapEncodeTiled") == 0) {
    *fn = (void *)&emu_encode_tiled;
    if (q) *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}
