// fillsim -- CPU model of the tile-level schedule of csrc/fill.cu (developer tool, not product code).
//
// Counts sweep rounds and tile visits of the chaotic tile relaxation on the benchmark terrain for
// different seedings / admission schedules, so that schedule ideas can be evaluated without a GPU.
// One "visit" relaxes a 64x64 tile to its local fixed point given its apron; tiles of one round
// read the state left by the previous round (Jacobi between tiles, the pessimistic model of the
// concurrent CTAs).  Every variant must end at the same surface; the tool checks that.
//
//   gcc -O2 -fopenmp -o fillsim fillsim.c -lm
//   ./fillsim N [seed_mode] [ordered] [Rfactor] [echo_filter] [bands]
//     seed_mode 0: border cells only (what fill.cu does)
//               1: + cells with a strictly descending steepest-descent path to the border (W = Z is exact there)
//               2 / 3 / 4 / 5: interior starts at the lifted fill of the 8x8 / 16x16 / 4x4 / 2x2 max-pooled raster (an upper bound)
//     ordered   0/1: level-ordered admission (quantile schedule with R = Rfactor * tiles across)
//     echo_filter 0/1/2: activate a neighbour always / only if the new edge value is below its adjacent cell / and that cell can go down
//     bands     G > 1: afterwards, flood again with the rows at every band seam preset to their exact values
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define TS 64
static int N, TN;  // raster side, tiles per side
static float *Z, *Wg;

static inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t s) {
  uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u + 0x7F4A7C15u) ^ (s * 0xC2B2AE3Du);
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}
static inline float lattice(uint32_t ix, uint32_t iy, uint32_t s) { return (float)(hash3(ix, iy, s) >> 8) * (1.0f / 16777216.0f); }

static void gen_terrain(uint32_t seed, int octaves, int top_log2) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < N; y++)
    for (int x = 0; x < N; x++) {
      float sum = 0.f, amp = 1.f, norm = 0.f;
      for (int o = 0; o < octaves; o++) {
        const int lg = top_log2 - o;
        if (lg < 1) break;
        const uint32_t cell = 1u << lg;
        const uint32_t ix = (uint32_t)x >> lg, iy = (uint32_t)y >> lg;
        float tx = (float)((uint32_t)x & (cell - 1)) / (float)cell;
        float ty = (float)((uint32_t)y & (cell - 1)) / (float)cell;
        tx = tx * tx * tx * (tx * (tx * 6.f - 15.f) + 10.f);
        ty = ty * ty * ty * (ty * (ty * 6.f - 15.f) + 10.f);
        const uint32_t s = seed * 131u + (uint32_t)o;
        const float v00 = lattice(ix, iy, s), v01 = lattice(ix + 1, iy, s);
        const float v10 = lattice(ix, iy + 1, s), v11 = lattice(ix + 1, iy + 1, s);
        const float v = (v00 * (1.f - tx) + v01 * tx) * (1.f - ty) + (v10 * (1.f - tx) + v11 * tx) * ty;
        sum += amp * v; norm += amp; amp *= 0.5946035575f;
      }
      Z[(size_t)y * N + x] = 1000.0f * sum / norm;
    }
}

// ---- steepest-descent receivers in the total order (Z, index); status by memoised path following ----
enum { ST_UNKNOWN = 0, ST_DRAINED = 1, ST_PIT = 2 };
static const int DX[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, DY[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
static inline int64_t receiver(int x, int y) {
  const size_t i = (size_t)y * N + x;
  float bz = Z[i];
  size_t bi = i;
  for (int k = 0; k < 8; k++) {
    const size_t j = (size_t)(y + DY[k]) * N + (x + DX[k]);
    const float zj = Z[j];
    if (zj < bz || (zj == bz && j < bi)) { bz = zj; bi = j; }
  }
  return bi == i ? -1 : (int64_t)bi;
}
static uint8_t *status;
static void compute_drained(void) {
  status = calloc((size_t)N * N, 1);
  for (int y = 0; y < N; y++)
    for (int x = 0; x < N; x++)
      if (x == 0 || y == 0 || x == N - 1 || y == N - 1) status[(size_t)y * N + x] = ST_DRAINED;
  int64_t *stack = malloc(sizeof(int64_t) * 1 << 24);
  for (int y = 1; y < N - 1; y++)
    for (int x = 1; x < N - 1; x++) {
      int64_t c = (int64_t)y * N + x;
      if (status[c]) continue;
      int sp = 0;
      uint8_t st;
      for (;;) {
        if (status[c]) { st = status[c]; break; }
        const int64_t r = receiver((int)(c % N), (int)(c / N));
        if (r < 0) { st = ST_PIT; status[c] = st; break; }
        stack[sp++] = c;
        if (sp >= (1 << 24)) { fprintf(stderr, "stack overflow\n"); exit(1); }
        c = r;
      }
      while (sp) status[stack[--sp]] = st;
    }
  free(stack);
}

// ---- tile machinery ----
typedef struct { float w[TS * TS]; int tile; int sides; float key; } TileOut;
enum { S_N = 1, S_S = 2, S_W = 4, S_E = 8, S_NW = 16, S_NE = 32, S_SW = 64, S_SE = 128 };

static long long g_passes = 0;
static int g_vk = 0, g_vM = 0, g_vevery = 0;  // V-cycle: pool size, coarse side, fine rounds between coarse corrections
static float *g_Zc = NULL, *g_Wc = NULL;
static int g_echo_filter = 0;
static long long g_chist[8];  // visits by changed cells: 0, 1-16, 17-64, 65-256, 257-1024, 1025-2048, 2049-4095, 4096
static long long g_changed_cells = 0;

// relax tile t to its local fixed point; returns 1 if anything changed
static int relax_tile(int t, TileOut *out) {
  const int ty = t / TN, tx = t % TN, x0 = tx * TS, y0 = ty * TS;
  float w[(TS + 2) * (TS + 2)], z[TS * TS];
  const int P = TS + 2;
  for (int j = -1; j <= TS; j++)
    for (int i = -1; i <= TS; i++) {
      const int x = x0 + i, y = y0 + j;
      w[(j + 1) * P + i + 1] = (x >= 0 && y >= 0 && x < N && y < N) ? Wg[(size_t)y * N + x] : INFINITY;
    }
  for (int j = 0; j < TS; j++) memcpy(&z[j * TS], &Z[(size_t)(y0 + j) * N + x0], TS * 4);
  int any = 0, passes = 0;
  for (;;) {
    int ch = 0;
    for (int j = 1; j <= TS; j++)
      for (int i = 1; i <= TS; i++) {
        float *c = &w[j * P + i];
        float m = fminf(fminf(fminf(c[-P - 1], c[-P]), fminf(c[-P + 1], c[-1])), fminf(fminf(c[1], c[P - 1]), fminf(c[P], c[P + 1])));
        const float nw = fmaxf(z[(j - 1) * TS + i - 1], m);
        if (nw < *c) { *c = nw; ch = 1; }
      }
    for (int j = TS; j >= 1; j--)
      for (int i = TS; i >= 1; i--) {
        float *c = &w[j * P + i];
        float m = fminf(fminf(fminf(c[-P - 1], c[-P]), fminf(c[-P + 1], c[-1])), fminf(fminf(c[1], c[P - 1]), fminf(c[P], c[P + 1])));
        const float nw = fmaxf(z[(j - 1) * TS + i - 1], m);
        if (nw < *c) { *c = nw; ch = 1; }
      }
    passes++;
    if (!ch) break;
    any = 1;
  }
#pragma omp atomic
  g_passes += passes;
  if (!any) {
#pragma omp atomic
    g_chist[0]++;
    return 0;
  }
  out->tile = t;
  out->sides = 0;
  out->key = INFINITY;
  for (int j = 0; j < TS; j++)
    for (int i = 0; i < TS; i++) {
      const float nv = w[(j + 1) * P + i + 1];
      out->w[j * TS + i] = nv;
      const float ov = Wg[(size_t)(y0 + j) * N + x0 + i];
      if (nv < ov) {
        if (!g_echo_filter) {
          if (j == 0) out->sides |= S_N;
          if (j == TS - 1) out->sides |= S_S;
          if (i == 0) out->sides |= S_W;
          if (i == TS - 1) out->sides |= S_E;
          if (j == 0 && i == 0) out->sides |= S_NW;
          if (j == 0 && i == TS - 1) out->sides |= S_NE;
          if (j == TS - 1 && i == 0) out->sides |= S_SW;
          if (j == TS - 1 && i == TS - 1) out->sides |= S_SE;
        } else if (j == 0 || i == 0 || j == TS - 1 || i == TS - 1) {
          // activate a neighbour only if the new value is below one of ITS cells next to this one
          // (apron copy; the live value can only be lower, so this never misses a needed visit)
          for (int dj = -1; dj <= 1; dj++)
            for (int di = -1; di <= 1; di++) {
              const int aj = j + dj, ai = i + di;
              if (aj >= 0 && aj < TS && ai >= 0 && ai < TS) continue;  // inside this tile
              if (!(nv < w[(aj + 1) * P + ai + 1])) continue;
              if (g_echo_filter >= 2) {  // exact: the neighbour's cell must be able to go down (needs its Z)
                const int gx = x0 + ai, gy = y0 + aj;
                if (gx < 0 || gy < 0 || gx >= N || gy >= N) continue;
                if (!(fmaxf(Z[(size_t)gy * N + gx], nv) < w[(aj + 1) * P + ai + 1])) continue;
              }
              const int sy = aj < 0 ? -1 : (aj >= TS ? 1 : 0), sx = ai < 0 ? -1 : (ai >= TS ? 1 : 0);
              if (sy == -1 && sx == 0) out->sides |= S_N;
              if (sy == 1 && sx == 0) out->sides |= S_S;
              if (sy == 0 && sx == -1) out->sides |= S_W;
              if (sy == 0 && sx == 1) out->sides |= S_E;
              if (sy == -1 && sx == -1) out->sides |= S_NW;
              if (sy == -1 && sx == 1) out->sides |= S_NE;
              if (sy == 1 && sx == -1) out->sides |= S_SW;
              if (sy == 1 && sx == 1) out->sides |= S_SE;
            }
        }
        if (j < 4 || i < 4 || j >= TS - 4 || i >= TS - 4) out->key = fminf(out->key, nv);
      }
    }
  {
    int nc = 0;
    for (int j = 0; j < TS; j++)
      for (int i = 0; i < TS; i++)
        if (out->w[j * TS + i] < Wg[(size_t)(y0 + j) * N + x0 + i]) nc++;
    const int b = nc <= 16 ? 1 : nc <= 64 ? 2 : nc <= 256 ? 3 : nc <= 1024 ? 4 : nc <= 2048 ? 5 : nc < 4096 ? 6 : 7;
#pragma omp atomic
    g_chist[b]++;
#pragma omp atomic
    g_changed_cells += nc;
  }
  return 1;
}

int main(int argc, char **argv) {
  N = argc > 1 ? atoi(argv[1]) : 4096;
  const int seed_mode = argc > 2 ? atoi(argv[2]) : 0;
  const int ordered = argc > 3 ? atoi(argv[3]) : 1;
  const double rfac = argc > 4 ? atof(argv[4]) : 0.8;
  g_echo_filter = argc > 5 ? atoi(argv[5]) : 0;
  const int bandsG = argc > 6 ? atoi(argv[6]) : 0;
  g_vevery = argc > 7 ? atoi(argv[7]) : 0;  // > 0 (with seed_mode 2..5): coarse-grid correction every that many fine rounds  // > 1: also measure a flood whose band seam rows start at their exact values
  const char *dump = NULL;
  TN = N / TS;
  if (N % TS) { fprintf(stderr, "N must be a multiple of %d\n", TS); return 1; }
  Z = malloc((size_t)N * N * 4);
  Wg = malloc((size_t)N * N * 4);
  double t0 = omp_get_wtime();
  gen_terrain(42, 12, 12);
  fprintf(stderr, "terrain %.1fs\n", omp_get_wtime() - t0);

  // init
  size_t nseed = 0;
  if (seed_mode == 1) {
    t0 = omp_get_wtime();
    compute_drained();
    fprintf(stderr, "drained %.1fs\n", omp_get_wtime() - t0);
  }
  float *W0 = NULL;
  if (seed_mode >= 2 && seed_mode <= 5) {
    // coarse upper bound: max-pool Z over k x k blocks, fill the coarse raster exactly, lift (every fine cell of a
    // block can reach the border below the block's coarse level: blocks are internally connected and 8-adjacent
    // blocks share a fine 8-adjacency)
    const int k = seed_mode == 2 ? 8 : (seed_mode == 3 ? 16 : (seed_mode == 4 ? 4 : 2)), M = N / k;
    float *Zc = malloc((size_t)M * M * 4), *Wc = malloc((size_t)M * M * 4);
    for (int by = 0; by < M; by++)
      for (int bx = 0; bx < M; bx++) {
        float m = -INFINITY;
        for (int j = 0; j < k; j++)
          for (int i = 0; i < k; i++) m = fmaxf(m, Z[(size_t)(by * k + j) * N + bx * k + i]);
        Zc[(size_t)by * M + bx] = m;
      }
    for (int i = 0; i < M * M; i++) Wc[i] = INFINITY;
    for (int y = 0; y < M; y++)
      for (int x = 0; x < M; x++)
        if (!x || !y || x == M - 1 || y == M - 1) Wc[(size_t)y * M + x] = Zc[(size_t)y * M + x];
    for (;;) {
      int ch = 0;
      for (int y = 1; y < M - 1; y++)
        for (int x = 1; x < M - 1; x++) {
          float *c = &Wc[(size_t)y * M + x];
          float m = fminf(fminf(fminf(c[-M - 1], c[-M]), fminf(c[-M + 1], c[-1])), fminf(fminf(c[1], c[M - 1]), fminf(c[M], c[M + 1])));
          const float nw = fmaxf(Zc[(size_t)y * M + x], m);
          if (nw < *c) { *c = nw; ch = 1; }
        }
      for (int y = M - 2; y >= 1; y--)
        for (int x = M - 2; x >= 1; x--) {
          float *c = &Wc[(size_t)y * M + x];
          float m = fminf(fminf(fminf(c[-M - 1], c[-M]), fminf(c[-M + 1], c[-1])), fminf(fminf(c[1], c[M - 1]), fminf(c[M], c[M + 1])));
          const float nw = fmaxf(Zc[(size_t)y * M + x], m);
          if (nw < *c) { *c = nw; ch = 1; }
        }
      if (!ch) break;
    }
    W0 = malloc((size_t)N * N * 4);
    for (int y = 0; y < N; y++)
      for (int x = 0; x < N; x++) W0[(size_t)y * N + x] = Wc[(size_t)(y / k) * M + x / k];
    g_vk = k; g_vM = M; g_Zc = Zc; g_Wc = Wc;  // kept for the optional V-cycles
  }
  size_t npit = 0;
  for (int y = 0; y < N; y++)
    for (int x = 0; x < N; x++) {
      const size_t i = (size_t)y * N + x;
      const int border = x == 0 || y == 0 || x == N - 1 || y == N - 1;
      int s = border;
      if (seed_mode == 1 && status[i] == ST_DRAINED) s = 1;
      if (seed_mode == 1 && !border && receiver(x, y) < 0) npit++;
      Wg[i] = s ? Z[i] : (W0 ? W0[i] : INFINITY);
      nseed += s;
    }
  printf("N=%d tiles=%d seed_mode=%d ordered=%d seeded_cells=%zu (%.2f%%) pits=%zu (%.3f%%)\n", N, TN * TN, seed_mode, ordered, nseed,
         100.0 * nseed / ((double)N * N), npit, 100.0 * npit / ((double)N * N));

  // level schedule
  int R = (int)(rfac * TN);
  float *levels = NULL;
  if (ordered && R >= 8) {
    const int HB = 1024;
    float zmin = INFINITY, zmax = -INFINITY;
    for (size_t i = 0; i < (size_t)N * N; i++) { zmin = fminf(zmin, Z[i]); zmax = fmaxf(zmax, Z[i]); }
    double *h = calloc(HB, sizeof(double));
    for (int y = 0; y < N; y += 16)
      for (int x = 0; x < N; x++) {
        int b = (int)((Z[(size_t)y * N + x] - zmin) / (zmax - zmin) * HB);
        if (b >= HB) b = HB - 1;
        h[b]++;
      }
    double total = 0;
    for (int k = 0; k < HB; k++) total += h[k];
    levels = malloc(sizeof(float) * R);
    double cum = 0; int bin = 0;
    for (int k = 0; k < R; k++) {
      const double want = total * (k + 1) / R;
      while (bin < HB - 1 && cum + h[bin] < want) cum += h[bin++];
      levels[k] = zmin + (zmax - zmin) * (float)(bin + 1) / HB;
    }
  } else R = 0;

  const int NT = TN * TN;
  int *list = malloc(sizeof(int) * NT), *next = malloc(sizeof(int) * NT);
  int *stamp = calloc(NT, sizeof(int));
  float *key[2] = {malloc(sizeof(float) * NT), malloc(sizeof(float) * NT)};
  for (int t = 0; t < NT; t++) key[0][t] = key[1][t] = INFINITY;
  int n = 0;
  for (int t = 0; t < NT; t++) {
    const int ty = t / TN, tx = t % TN;
    if (seed_mode >= 1 || ty == 0 || tx == 0 || ty == TN - 1 || tx == TN - 1) { list[n++] = t; key[1][t] = -INFINITY; stamp[t] = 1; }
  }
  TileOut *outs = malloc(sizeof(TileOut) * (size_t)NT);
  int *proc = malloc(sizeof(int) * NT);
  long long visits = 0, deferred = 0, vcycles = 0, vlowered = 0;
  int round = 1;
  long long hist_small = 0;  // rounds with fewer tiles than 888 CTAs
  double model_us = 0;
  t0 = omp_get_wtime();
  while (n > 0) {
    const float level = (round - 1 < R) ? levels[round - 1] : INFINITY;
    int nn = 0, np = 0;
    float *kc = key[round & 1], *kn = key[(round + 1) & 1];
    for (int i = 0; i < n; i++) {
      const int t = list[i];
      if (kc[t] <= level) { proc[np++] = t; kc[t] = INFINITY; }
      else {  // postponed
        if (stamp[t] != round + 1) { stamp[t] = round + 1; next[nn++] = t; }
        kn[t] = fminf(kn[t], kc[t]); kc[t] = INFINITY; deferred++;
      }
    }
    int nout = 0;
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < np; i++) {
      TileOut o;
      if (relax_tile(proc[i], &o)) {
        int k;
#pragma omp atomic capture
        k = nout++;
        outs[k] = o;
      }
    }
    visits += np;
    if (np < 888) hist_small++;
    // time model: 888 resident CTAs, ~35 us per visit, ~8 us launch floor (x2 launches with admit)
    { double waves = ceil(np / 888.0); model_us += fmax(8.0, waves * 35.0) + (round - 1 < R ? 4.0 : 0.0); }
    for (int k = 0; k < nout; k++) {
      const TileOut *o = &outs[k];
      const int t = o->tile, ty = t / TN, tx = t % TN;
      for (int j = 0; j < TS; j++) memcpy(&Wg[(size_t)(ty * TS + j) * N + tx * TS], &o->w[j * TS], TS * 4);
      const int nb[8][3] = {{S_N, 0, -1}, {S_S, 0, 1}, {S_W, -1, 0}, {S_E, 1, 0}, {S_NW, -1, -1}, {S_NE, 1, -1}, {S_SW, -1, 1}, {S_SE, 1, 1}};
      for (int q = 0; q < 8; q++)
        if (o->sides & nb[q][0]) {
          const int ux = tx + nb[q][1], uy = ty + nb[q][2];
          if (ux < 0 || uy < 0 || ux >= TN || uy >= TN) continue;
          const int u = uy * TN + ux;
          kn[u] = fminf(kn[u], o->key);
          if (stamp[u] != round + 1) { stamp[u] = round + 1; next[nn++] = u; }
        }
    }
    int *tmp = list; list = next; next = tmp;
    n = nn;
    round++;
    if (g_vevery > 0 && g_Wc && n > 0 && (round - 1) % g_vevery == 0) {
      // coarse-grid correction: restrict (block maximum of the current fine upper bound), relax the coarse raster,
      // prolong (fine = min(fine, lifted)); both steps keep every value an upper bound of the answer
      const int k = g_vk, M = g_vM;
#pragma omp parallel for schedule(static)
      for (int by = 0; by < M; by++)
        for (int bx = 0; bx < M; bx++) {
          float m = -INFINITY;
          for (int j = 0; j < k; j++)
            for (int i = 0; i < k; i++) m = fmaxf(m, Wg[(size_t)(by * k + j) * N + bx * k + i]);
          float *wc = &g_Wc[(size_t)by * M + bx];
          if (m < *wc) *wc = m;
        }
      for (;;) {
        int ch = 0;
        for (int y = 1; y < M - 1; y++)
          for (int x = 1; x < M - 1; x++) {
            float *c = &g_Wc[(size_t)y * M + x];
            float m = fminf(fminf(fminf(c[-M - 1], c[-M]), fminf(c[-M + 1], c[-1])), fminf(fminf(c[1], c[M - 1]), fminf(c[M], c[M + 1])));
            const float nw = fmaxf(g_Zc[(size_t)y * M + x], m);
            if (nw < *c) { *c = nw; ch = 1; }
          }
        for (int y = M - 2; y >= 1; y--)
          for (int x = M - 2; x >= 1; x--) {
            float *c = &g_Wc[(size_t)y * M + x];
            float m = fminf(fminf(fminf(c[-M - 1], c[-M]), fminf(c[-M + 1], c[-1])), fminf(fminf(c[1], c[M - 1]), fminf(c[M], c[M + 1])));
            const float nw = fmaxf(g_Zc[(size_t)y * M + x], m);
            if (nw < *c) { *c = nw; ch = 1; }
          }
        if (!ch) break;
      }
      long long lowered = 0;
      float *kc2 = key[round & 1];
      for (int y = 1; y < N - 1; y++)
        for (int x = 1; x < N - 1; x++) {
          const float l = g_Wc[(size_t)(y / k) * M + x / k];
          float *w = &Wg[(size_t)y * N + x];
          if (l < *w) {
            *w = l;
            lowered++;
            const int t = (y / TS) * TN + x / TS;
            if (stamp[t] != round) { stamp[t] = round; list[n++] = t; }
            kc2[t] = -INFINITY;
          }
        }
      vcycles++;
      vlowered += lowered;
    }
    if (round % 50 == 0) fprintf(stderr, "round %d active %d visits %lld (%.1fs)\n", round, n, visits, omp_get_wtime() - t0);
  }
  if (g_vevery) printf("coarse corrections=%lld, cells lowered by prolongation=%lld\n", vcycles, vlowered);
  printf("rounds=%d visits=%lld (%.2f raster-equivalents) deferred=%lld passes/visit=%.2f small_rounds=%lld model_ms=%.1f\n", round - 1, visits,
         (double)visits / NT, deferred, (double)g_passes / visits, hist_small, model_us / 1000.0);
  printf("visits by changed cells [0 | 1-16 | 17-64 | 65-256 | 257-1024 | 1025-2048 | 2049-4095 | 4096]:");
  for (int k = 0; k < 8; k++) printf(" %lld", g_chist[k]);
  printf("  cell updates=%.2f per cell\n", (double)g_changed_cells / ((double)N * N));
  if (bandsG > 1) {
    // Seam experiment: how many rounds / visits does the flood need when the rows on both sides of every band seam
    // already hold their exact values (what a seam spill-graph solve would provide)?  Every band then floods on its own.
    float *exact = malloc((size_t)N * N * 4);
    memcpy(exact, Wg, (size_t)N * N * 4);
    const int bh = N / bandsG;
    for (int y = 0; y < N; y++)
      for (int x = 0; x < N; x++) {
        const size_t i = (size_t)y * N + x;
        const int border = x == 0 || y == 0 || x == N - 1 || y == N - 1;
        const int seam = (y % bh == 0 || y % bh == bh - 1) && y > 0 && y < N - 1;
        Wg[i] = (border || seam) ? exact[i] : INFINITY;
      }
    for (int t = 0; t < NT; t++) { key[0][t] = key[1][t] = INFINITY; stamp[t] = 0; }
    n = 0;
    for (int t = 0; t < NT; t++) {
      const int ty = t / TN, tx = t % TN;
      const int tb = bh / TS;  // tile rows per band
      if (tx == 0 || tx == TN - 1 || ty % tb == 0 || ty % tb == tb - 1) { list[n++] = t; key[1][t] = -INFINITY; stamp[t] = 1; }
    }
    long long v2 = 0; int r2 = 1;
    while (n > 0) {
      const float level = (r2 - 1 < R) ? levels[r2 - 1] : INFINITY;
      int nn = 0, np = 0;
      float *kc = key[r2 & 1], *kn = key[(r2 + 1) & 1];
      for (int i = 0; i < n; i++) {
        const int t = list[i];
        if (kc[t] <= level) { proc[np++] = t; kc[t] = INFINITY; }
        else { if (stamp[t] != r2 + 1) { stamp[t] = r2 + 1; next[nn++] = t; } kn[t] = fminf(kn[t], kc[t]); kc[t] = INFINITY; }
      }
      int nout = 0;
#pragma omp parallel for schedule(dynamic, 8)
      for (int i = 0; i < np; i++) {
        TileOut o;
        if (relax_tile(proc[i], &o)) { int k;
#pragma omp atomic capture
          k = nout++;
          outs[k] = o; }
      }
      v2 += np;
      for (int k = 0; k < nout; k++) {
        const TileOut *o = &outs[k];
        const int t = o->tile, ty = t / TN, tx = t % TN;
        for (int j = 0; j < TS; j++) memcpy(&Wg[(size_t)(ty * TS + j) * N + tx * TS], &o->w[j * TS], TS * 4);
        const int nb[8][3] = {{S_N, 0, -1}, {S_S, 0, 1}, {S_W, -1, 0}, {S_E, 1, 0}, {S_NW, -1, -1}, {S_NE, 1, -1}, {S_SW, -1, 1}, {S_SE, 1, 1}};
        for (int q = 0; q < 8; q++)
          if (o->sides & nb[q][0]) {
            const int ux = tx + nb[q][1], uy = ty + nb[q][2];
            if (ux < 0 || uy < 0 || ux >= TN || uy >= TN) continue;
            const int u = uy * TN + ux;
            kn[u] = fminf(kn[u], o->key);
            if (stamp[u] != r2 + 1) { stamp[u] = r2 + 1; next[nn++] = u; }
          }
      }
      int *tmp = list; list = next; next = tmp; n = nn; r2++;
    }
    size_t bad2 = 0;
    for (size_t i = 0; i < (size_t)N * N; i++) if (Wg[i] != exact[i]) bad2++;
    printf("bands=%d with exact seam rows: rounds=%d visits=%lld (%.2f raster-equivalents) mismatches=%zu\n", bandsG, r2 - 1, v2, (double)v2 / NT, bad2);
    free(exact);
  }
  // fixed-point check + checksum
  size_t bad = 0; double sum = 0; size_t nfilled = 0;
  for (int y = 1; y < N - 1; y++)
    for (int x = 1; x < N - 1; x++) {
      const float *c = &Wg[(size_t)y * N + x];
      float m = fminf(fminf(fminf(c[-N - 1], c[-N]), fminf(c[-N + 1], c[-1])), fminf(fminf(c[1], c[N - 1]), fminf(c[N], c[N + 1])));
      if (*c != fmaxf(Z[(size_t)y * N + x], m)) bad++;
      sum += *c;
      if (*c > Z[(size_t)y * N + x]) nfilled++;
    }
  printf("fixed-point violations=%zu checksum=%.6f filled_cells=%zu (%.2f%%)\n", bad, sum, nfilled, 100.0 * nfilled / ((double)N * N));
  if (dump) { FILE *f = fopen(dump, "wb"); fwrite(Wg, 4, (size_t)N * N, f); fclose(f); }
  return 0;
}
