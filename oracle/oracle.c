/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the RichDEM CPU algorithms on the
 * fill -> flats -> flow-accumulation hot path.  It exists so the CUDA path can be checked
 * on a machine where /root/reference is absent (the GPU box).  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (richdem_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (a) the reference's own known-answer fixtures (tests/depressions/testdem1.*,
 *       tests/flow_accum/*.d8/.out, re-encoded in tests/golden/), and
 *   (b) outputs of the unmodified reference compiled from /root/reference/include
 *       (oracle/ref_shim.cpp -> oracle/_ref/libref_richdem.so) on seeded random and
 *       real-terrain (Beauford crop) inputs, stored in tests/golden/ and, when the
 *       reference library is present, re-compared live.
 *
 * Layout everywhere: row-major i = y*W + x (reference common/Array2D.hpp:592-595).
 * D8 neighbour numbering (reference common/constants.hpp:44-45,65):
 *      2 3 4
 *      1 0 5
 *      8 7 6
 * Citations are file:line relative to /root/reference/include/richdem/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int D8X[9] = {0, -1, -1, 0, 1, 1, 1, 0, -1};
static const int D8Y[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1};

#define NO_FLOW_GEN (-1.0f) /* common/constants.hpp:83 */
#define HAS_FLOW_GEN (0.0f) /* :84 */
#define NO_DATA_GEN (-2.0f) /* :85 */

/* ------------------------------------------------------------------------------------ */
/* small containers                                                                      */

typedef struct {
  float z;
  int64_t i;
} HeapItem;

typedef struct {
  HeapItem *a;
  size_t n, cap;
} Heap;

static void heap_push(Heap *h, float z, int64_t i) {
  if (h->n == h->cap) {
    h->cap = h->cap ? h->cap * 2 : 1024;
    h->a = (HeapItem *)realloc(h->a, h->cap * sizeof(HeapItem));
  }
  size_t k = h->n++;
  while (k > 0) {
    size_t p = (k - 1) / 2;
    if (h->a[p].z <= z) break;
    h->a[k] = h->a[p];
    k = p;
  }
  h->a[k].z = z;
  h->a[k].i = i;
}

static HeapItem heap_pop(Heap *h) {
  HeapItem top = h->a[0];
  HeapItem last = h->a[--h->n];
  size_t k = 0;
  for (;;) {
    size_t c = 2 * k + 1;
    if (c >= h->n) break;
    if (c + 1 < h->n && h->a[c + 1].z < h->a[c].z) c++;
    if (last.z <= h->a[c].z) break;
    h->a[k] = h->a[c];
    k = c;
  }
  if (h->n) h->a[k] = last;
  return top;
}

typedef struct {
  int64_t *a;
  size_t head, tail, cap; /* simple growing FIFO (never wraps) */
} Fifo;

static void fifo_push(Fifo *q, int64_t v) {
  if (q->tail == q->cap) {
    if (q->head > q->cap / 2) { /* compact */
      memmove(q->a, q->a + q->head, (q->tail - q->head) * sizeof(int64_t));
      q->tail -= q->head;
      q->head = 0;
    } else {
      q->cap = q->cap ? q->cap * 2 : 1024;
      q->a = (int64_t *)realloc(q->a, q->cap * sizeof(int64_t));
    }
  }
  q->a[q->tail++] = v;
}
static int fifo_empty(const Fifo *q) { return q->head == q->tail; }
static int64_t fifo_pop(Fifo *q) { return q->a[q->head++]; }

/* ------------------------------------------------------------------------------------ */
/* a1/a2  Priority-Flood depression filling (D8)                                         */
/*                                                                                       */
/* Restates depressions/Barnes2014.hpp:230-304 (PriorityFlood_Barnes2014: priority queue  */
/* of edge cells + plain FIFO for pit cells).  depressions/Zhou2016.hpp:125-191, which    */
/* FillDepressions<D8> dispatches to (depressions/depressions.hpp:13-21), produces the    */
/* same raster (the reference's tests assert that: tests/tests.cpp:238-271); the output is*/
/* the unique min-over-paths-of-max-elevation surface, so queue order does not matter.    */
/* NoData is NOT special-cased (neither reference function looks at it).                  */
static void orc_fill_topo(float *dem, int w, int h, int d4) {
  const size_t n = (size_t)w * h;
  uint8_t *closed = (uint8_t *)calloc(n, 1);
  Heap open = {0, 0, 0};
  Fifo pit = {0, 0, 0, 0};
  float *pitz = NULL; /* level carried with each pit entry: store in parallel FIFO */
  Fifo pitlev = {0, 0, 0, 0};
  (void)pitz;

  for (int x = 0; x < w; x++) { /* Barnes2014.hpp:257-262 */
    heap_push(&open, dem[x], x);
    closed[x] = 1;
    if (h > 1) {
      size_t i = (size_t)(h - 1) * w + x;
      heap_push(&open, dem[i], (int64_t)i);
      closed[i] = 1;
    }
  }
  for (int y = 1; y < h - 1; y++) { /* :263-268 */
    size_t i = (size_t)y * w;
    heap_push(&open, dem[i], (int64_t)i);
    closed[i] = 1;
    if (w > 1) {
      i = (size_t)y * w + (w - 1);
      heap_push(&open, dem[i], (int64_t)i);
      closed[i] = 1;
    }
  }

  while (open.n > 0 || !fifo_empty(&pit)) { /* :272-299 */
    int64_t ci;
    float cz;
    if (!fifo_empty(&pit)) {
      ci = fifo_pop(&pit);
      int64_t bits = fifo_pop(&pitlev);
      uint32_t u = (uint32_t)bits;
      memcpy(&cz, &u, 4);
    } else {
      HeapItem t = heap_pop(&open);
      ci = t.i;
      cz = t.z;
    }
    const int cx = (int)(ci % w), cy = (int)(ci / w);
    for (int k = 1; k <= 8; k++) {
      if (d4 && !(k & 1)) continue; /* D4 (common/constants.hpp:53-54): the cardinal neighbours are the odd D8 codes */
      const int nx = cx + D8X[k], ny = cy + D8Y[k];
      if (nx < 0 || ny < 0 || nx >= w || ny >= h) continue;
      const size_t ni = (size_t)ny * w + nx;
      if (closed[ni]) continue;
      closed[ni] = 1;
      if (dem[ni] <= cz) { /* :289-294 */
        dem[ni] = cz;
        uint32_t u;
        memcpy(&u, &cz, 4);
        fifo_push(&pit, (int64_t)ni);
        fifo_push(&pitlev, (int64_t)u);
      } else {
        heap_push(&open, dem[ni], (int64_t)ni);
      }
    }
  }
  free(closed);
  free(open.a);
  free(pit.a);
  free(pitlev.a);
}

void orc_fill_depressions_d8_f32(float *dem, int w, int h) { orc_fill_topo(dem, w, h, 0); }
/* FillDepressions<Topology::D4> = PriorityFlood_Barnes2014<D4> (depressions/depressions.hpp:16-17) */
void orc_fill_depressions_d4_f32(float *dem, int w, int h) { orc_fill_topo(dem, w, h, 1); }

/* ------------------------------------------------------------------------------------ */
/* a3  FindFlats  (flats/find_flats.hpp:28-69)                                            */
/* -1 NoData, 0 NOT_A_FLAT (raster-edge cell, or has a lower / NoData D8 neighbour), 1 flat */
void orc_find_flats_f32(const float *dem, int w, int h, float nodata, int8_t *flats) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      if (dem[i] == nodata) { /* :42-45 */
        flats[i] = -1;
        continue;
      }
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) { /* :47-50 */
        flats[i] = 0;
        continue;
      }
      int8_t f = 1;
      for (int k = 1; k <= 8; k++) { /* :55-62 */
        const float ne = dem[(size_t)(y + D8Y[k]) * w + (x + D8X[k])];
        if (ne < dem[i] || ne == nodata) {
          f = 0;
          break;
        }
      }
      flats[i] = f;
    }
}

/* ------------------------------------------------------------------------------------ */
/* a4-a7  GetFlatMask (flats/Barnes2014.hpp:398-467)                                      */
/* Outputs: mask[] (increments per cell) and labels[] (flat id, 0 = none; ids are         */
/* traversal-order dependent, only the partition is meaningful).                          */
void orc_get_flat_mask_f32(const float *dem, int w, int h, float nodata, int32_t *mask,
                           int32_t *labels) {
  const size_t n = (size_t)w * h;
  int8_t *flats = (int8_t *)malloc(n);
  orc_find_flats_f32(dem, w, h, nodata, flats);
  memset(mask, 0, n * sizeof(int32_t));
  memset(labels, 0, n * sizeof(int32_t));

  Fifo low = {0, 0, 0, 0}, high = {0, 0, 0, 0};
  /* FindFlatEdges, flats/Barnes2014.hpp:309-369 */
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      if (flats[i] == -1) continue; /* :330-331 */
      for (int k = 1; k <= 8; k++) {
        const int nx = x + D8X[k], ny = y + D8Y[k];
        if (nx < 0 || ny < 0 || nx >= w || ny >= h) continue;
        const size_t ni = (size_t)ny * w + nx;
        if (flats[i] == 0 && flats[ni] == 1 && dem[ni] == dem[i]) { /* :343-350 low edge */
          fifo_push(&low, (int64_t)i);
          break;
        } else if (flats[i] == 1 && dem[i] < dem[ni]) { /* :354-360 high edge */
          fifo_push(&high, (int64_t)i);
          break;
        }
      }
    }

  if (fifo_empty(&low)) { /* :429-435 */
    free(flats);
    free(low.a);
    free(high.a);
    return;
  }

  /* LabelFlat for every unlabelled low edge, :437-441 and :244-280 */
  int32_t group = 1;
  Fifo q = {0, 0, 0, 0};
  for (size_t e = low.head; e < low.tail; e++) {
    const int64_t s = low.a[e];
    if (labels[s] != 0) continue;
    const float target = dem[s];
    const int32_t label = group++;
    q.head = q.tail = 0;
    fifo_push(&q, s);
    while (!fifo_empty(&q)) {
      const int64_t c = fifo_pop(&q);
      if (dem[c] != target) continue; /* :262-263 */
      if (labels[c] > 0) continue;    /* :266-267 */
      labels[c] = label;
      const int cx = (int)(c % w), cy = (int)(c / w);
      for (int k = 1; k <= 8; k++) {
        const int nx = cx + D8X[k], ny = cy + D8Y[k];
        if (nx < 0 || ny < 0 || nx >= w || ny >= h) continue;
        const int64_t ni = (int64_t)ny * w + nx;
        /* the reference pushes every in-grid neighbour (:273-277); filtering here on the
           two tests applied at pop time is equivalent and keeps the queue small */
        if (dem[ni] == target && labels[ni] == 0) fifo_push(&q, ni);
      }
    }
  }

  /* drop high edges of flats without outlets, :445-454 */
  Fifo high2 = {0, 0, 0, 0};
  for (size_t e = high.head; e < high.tail; e++)
    if (labels[high.a[e]] != 0) fifo_push(&high2, high.a[e]);

  int32_t *flat_height = (int32_t *)calloc((size_t)group, sizeof(int32_t));

  /* BuildAwayGradient, :62-110 : level-synchronous BFS from the high edges */
  {
    Fifo cur = high2, nxt = {0, 0, 0, 0};
    int32_t loops = 1;
    while (!fifo_empty(&cur)) {
      nxt.head = nxt.tail = 0;
      while (!fifo_empty(&cur)) {
        const int64_t c = fifo_pop(&cur);
        if (mask[c] > 0) continue; /* :89-90 */
        mask[c] = loops;           /* :93 */
        flat_height[labels[c]] = loops; /* :94 */
        const int cx = (int)(c % w), cy = (int)(c / w);
        for (int k = 1; k <= 8; k++) {
          const int nx = cx + D8X[k], ny = cy + D8Y[k];
          if (nx < 0 || ny < 0 || nx >= w || ny >= h) continue;
          const int64_t ni = (int64_t)ny * w + nx;
          if (labels[ni] == labels[c] && flats[ni] == 1) fifo_push(&nxt, ni); /* :98-104 */
        }
      }
      Fifo t = cur;
      cur = nxt;
      nxt = t;
      loops++;
    }
    free(cur.a);
    free(nxt.a);
  }

  /* BuildTowardsCombinedGradient, :152-211 */
  for (size_t i = 0; i < n; i++) mask[i] = -mask[i]; /* :169-172 */
  {
    Fifo cur = low, nxt = {0, 0, 0, 0};
    int32_t loops = 1;
    while (!fifo_empty(&cur)) {
      nxt.head = nxt.tail = 0;
      while (!fifo_empty(&cur)) {
        const int64_t c = fifo_pop(&cur);
        if (mask[c] > 0) continue; /* :187-188 */
        if (mask[c] != 0)          /* :191-194 */
          mask[c] = (flat_height[labels[c]] + mask[c]) + 2 * loops;
        else
          mask[c] = 2 * loops;
        const int cx = (int)(c % w), cy = (int)(c / w);
        for (int k = 1; k <= 8; k++) {
          const int nx = cx + D8X[k], ny = cy + D8Y[k];
          if (nx < 0 || ny < 0 || nx >= w || ny >= h) continue;
          const int64_t ni = (int64_t)ny * w + nx;
          if (labels[ni] == labels[c] && flats[ni] == 1) fifo_push(&nxt, ni); /* :196-206 */
        }
      }
      Fifo t = cur;
      cur = nxt;
      nxt = t;
      loops++;
    }
    free(cur.a);
    free(nxt.a);
  }
  /* every IS_A_FLAT cell of a labelled flat is reached by the towards-BFS (each IS_A_FLAT
     region borders an equal-elevation NOT_A_FLAT cell, which is a low edge), so no negated
     value from :169-172 survives. */
  free(flat_height);
  free(flats);
  free(high.a);
  free(q.a);
}

/* a8  ResolveFlatsEpsilon (flats/flats.hpp:21-28 -> flats/Barnes2014.hpp:496-550)        */
void orc_resolve_flats_epsilon_f32(float *dem, int w, int h, float nodata) {
  const size_t n = (size_t)w * h;
  int32_t *mask = (int32_t *)malloc(n * sizeof(int32_t));
  int32_t *labels = (int32_t *)malloc(n * sizeof(int32_t));
  orc_get_flat_mask_f32(dem, w, h, nodata, mask, labels);
  for (int y = 1; y < h - 1; y++)
    for (int x = 1; x < w - 1; x++) {
      const size_t i = (size_t)y * w + x;
      if (labels[i] == 0) continue; /* :515-516 */
      float z = dem[i];
      for (int32_t k = 0; k < mask[i]; k++) z = nextafterf(z, INFINITY); /* :527-528 */
      dem[i] = z;
    }
  free(mask);
  free(labels);
}

/* ------------------------------------------------------------------------------------ */
/* a12  d8_flow_directions (flowmet/d8_flowdirs.hpp:96-123, helper d8_FlowDir :32-74)     */
void orc_d8_flow_directions_f32(const float *dem, int w, int h, float nodata, uint8_t *dirs) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      if (dem[i] == nodata) { /* :117-118 */
        dirs[i] = 255;
        continue;
      }
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) { /* :36-54 */
        uint8_t d;
        if (x == 0 && y == 0) d = 2;
        else if (x == 0 && y == h - 1) d = 8;
        else if (x == w - 1 && y == 0) d = 4;
        else if (x == w - 1 && y == h - 1) d = 6;
        else if (x == 0) d = 1;
        else if (x == w - 1) d = 5;
        else if (y == 0) d = 3;
        else d = 7;
        dirs[i] = d;
        continue;
      }
      float minimum = dem[i];
      int flowdir = 0;
      for (int k = 1; k <= 8; k++) { /* :63-71 */
        const float ne = dem[(size_t)(y + D8Y[k]) * w + (x + D8X[k])];
        if (ne < minimum || (ne == minimum && flowdir > 0 && flowdir % 2 == 0 && k % 2 == 1)) {
          minimum = ne;
          flowdir = k;
        }
      }
      dirs[i] = (uint8_t)flowdir;
    }
}

/* a13  d8_flow_accum (methods/d8_methods.hpp:47-139) ; dirs NoData = `dir_nodata`         */
void orc_d8_flow_accum_i32(const int32_t *dirs, int w, int h, int32_t dir_nodata, int32_t *area) {
  const size_t n = (size_t)w * h;
  int8_t *dep = (int8_t *)calloc(n, 1);
  Fifo src = {0, 0, 0, 0};
  for (size_t i = 0; i < n; i++) area[i] = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) { /* :68-92 */
      const size_t i = (size_t)y * w + x;
      if (dirs[i] == dir_nodata) {
        area[i] = -1;
        continue;
      }
      const int k = dirs[i];
      if (k == 0) continue;
      const int nx = x + D8X[k], ny = y + D8Y[k];
      if (nx < 0 || ny < 0 || nx >= w || ny >= h) continue;
      dep[(size_t)ny * w + nx]++;
    }
  for (size_t i = 0; i < n; i++) /* :96-99 */
    if (dep[i] == 0 && dirs[i] != dir_nodata) fifo_push(&src, (int64_t)i);
  while (!fifo_empty(&src)) { /* :104-131 */
    const int64_t c = fifo_pop(&src);
    area[c]++;
    const int k = dirs[c];
    if (k == 0) continue;
    const int nx = (int)(c % w) + D8X[k], ny = (int)(c / w) + D8Y[k];
    if (nx < 0 || ny < 0 || nx >= w || ny >= h) continue;
    const size_t ni = (size_t)ny * w + nx;
    if (dirs[ni] == dir_nodata) continue;
    area[ni] += area[c];
    if (--dep[ni] == 0) fifo_push(&src, (int64_t)ni);
  }
  free(dep);
  free(src.a);
}

void orc_d8_flow_accum_u8_i32(const uint8_t *dirs, int w, int h, int32_t *area) {
  const size_t n = (size_t)w * h;
  int32_t *d = (int32_t *)malloc(n * sizeof(int32_t));
  for (size_t i = 0; i < n; i++) d[i] = dirs[i];
  orc_d8_flow_accum_i32(d, w, h, 255, area);
  free(d);
}

/* ------------------------------------------------------------------------------------ */
/* a9  FM_D8 = FM_OCallaghan<D8> (flowmet/OCallaghan1984.hpp:13-77,81-84)                 */
/* props is [y][x][9] float (common/Array3D.hpp:203-206)                                   */
void orc_fm_d8_f32(const float *dem, int w, int h, float nodata, float *props) {
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < 9 * n; i++) props[i] = NO_FLOW_GEN; /* :26 */
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      if (dem[i] == nodata) { /* :37-40 */
        props[9 * i] = NO_DATA_GEN;
        continue;
      }
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) continue; /* :42-43 */
      const float e = dem[i];
      int lowest_n = 0;
      float lowest = 3.402823466e+38f; /* numeric_limits<float>::max(), :48 */
      for (int k = 1; k <= 8; k++) {
        const float ne = dem[(size_t)(y + D8Y[k]) * w + (x + D8X[k])];
        if (ne == nodata) continue; /* :53-54 */
        if (ne >= e) continue;      /* :58-59 */
        if (ne < lowest) {          /* :61-64 */
          lowest = ne;
          lowest_n = k;
        }
      }
      if (lowest_n == 0) continue;
      props[9 * i] = HAS_FLOW_GEN;
      props[9 * i + lowest_n] = 1.0f;
    }
}

/* a10  FM_Tarboton (flowmet/Tarboton1997.hpp:14-144).  Compile WITHOUT -ffast-math and     */
/* with -ffp-contract=off so the double arithmetic matches the reference build.            */
void orc_fm_tarboton_f32(const float *dem, int w, int h, float nodata, float *props) {
  static const int dy_e1[9] = {0, 0, -1, -1, 0, 0, 1, 1, 0};
  static const int dx_e1[9] = {0, -1, 0, 0, 1, 1, 0, 0, -1};
  static const int dy_e2[9] = {0, -1, -1, -1, -1, 1, 1, 1, 1};
  static const int dx_e2[9] = {0, -1, -1, 1, 1, 1, 1, -1, -1};
  static const double af[9] = {0, -1., 1., -1., 1., -1., 1., -1., 1.};
  const double d1 = 1, d2 = 1;
  const float dang = (float)atan2(d2, d1); /* :29 */
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < 9 * n; i++) props[i] = NO_FLOW_GEN;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      if (dem[i] == nodata) {
        props[9 * i] = NO_DATA_GEN;
        continue;
      }
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) continue;
      int nmax = -1;
      double smax = 0;
      float rmax = 0;
      for (int k = 1; k <= 8; k++) { /* :75-114 */
        const int x1 = x + dx_e1[k], y1 = y + dy_e1[k];
        const int x2 = x + dx_e2[k], y2 = y + dy_e2[k];
        if (x1 < 0 || y1 < 0 || x1 >= w || y1 >= h) continue;
        if (dem[(size_t)y1 * w + x1] == nodata) continue;
        if (x2 < 0 || y2 < 0 || x2 >= w || y2 >= h) continue;
        if (dem[(size_t)y2 * w + x2] == nodata) continue;
        const double e0 = dem[i];
        const double e1 = dem[(size_t)y1 * w + x1];
        const double e2 = dem[(size_t)y2 * w + x2];
        const double s1 = (e0 - e1) / d1;
        const double s2 = (e1 - e2) / d2;
        double r = atan2(s2, s1);
        double s;
        if (r < 1e-7) {
          r = 0;
          s = s1;
        } else if (r > dang - 1e-7) {
          r = dang;
          s = (e0 - e2) / sqrt(d1 * d1 + d2 * d2);
        } else {
          s = sqrt(s1 * s1 + s2 * s2);
        }
        if (s > smax) {
          smax = s;
          nmax = k;
          rmax = (float)r;
        }
      }
      if (nmax == -1) continue;
      props[9 * i] = HAS_FLOW_GEN;
      if (af[nmax] == 1 && rmax == 0) rmax = dang; /* :121-126 */
      else if (af[nmax] == 1 && rmax == dang) rmax = 0;
      else if (af[nmax] == 1) rmax = (float)(M_PI / 4 - rmax);
      const int nn = (nmax + 1 == 9) ? 1 : nmax + 1;
      if (rmax == 0) { /* :134-141 */
        props[9 * i + nmax] = 1;
      } else if (rmax == dang) {
        props[9 * i + nn] = 1;
      } else {
        props[9 * i + nmax] = (float)(rmax / (M_PI / 4.));
        props[9 * i + nn] = (float)(1 - rmax / (M_PI / 4.));
      }
    }
}

/* a11  FlowAccumulation(props, accum) (methods/flow_accumulation_generic.hpp:33-100)     */
/* accum arrives holding the per-cell weights.                                            */
void orc_flow_accumulation_props_f64(const float *props, int w, int h, double *accum) {
  const size_t n = (size_t)w * h;
  static const int D8X_[9] = {0, -1, -1, 0, 1, 1, 1, 0, -1};
  static const int D8Y_[9] = {0, 0, -1, -1, -1, 0, 1, 1, 1};
  int64_t nshift[9];
  for (int k = 0; k < 9; k++) nshift[k] = (int64_t)D8Y_[k] * w + D8X_[k];
  int8_t *deps = (int8_t *)calloc(n, 1);
  for (int y = 1; y < h - 1; y++) /* :47-58 */
    for (int x = 1; x < w - 1; x++) {
      const size_t ci = (size_t)y * w + x;
      if (props[9 * ci] == NO_DATA_GEN) continue;
      for (int k = 1; k <= 8; k++)
        if (props[9 * ci + k] > 0) deps[ci + nshift[k]]++;
    }
  Fifo q = {0, 0, 0, 0};
  for (size_t i = 0; i < n; i++) /* :61-64 */
    if (deps[i] == 0 && props[9 * i] != NO_DATA_GEN) fifo_push(&q, (int64_t)i);
  while (!fifo_empty(&q)) { /* :71-92 */
    const int64_t ci = fifo_pop(&q);
    const double c_accum = accum[ci];
    const int cx = (int)(ci % w), cy = (int)(ci / w);
    const int interior = cx > 0 && cy > 0 && cx < w - 1 && cy < h - 1;
    for (int k = 1; k <= 8; k++) {
      if (props[9 * ci + k] <= 0) continue;
      if (!interior) continue; /* the reference would index off-grid here; FM_* never emit it */
      const int64_t ni = ci + nshift[k];
      if (props[9 * ni] == NO_DATA_GEN) continue;
      accum[ni] += props[9 * ci + k] * c_accum;
      if (--deps[ni] == 0) fifo_push(&q, ni);
    }
  }
  for (size_t i = 0; i < n; i++) /* :95-97 */
    if (props[9 * i] == NO_DATA_GEN) accum[i] = -1;
  free(deps);
  free(q.a);
}

/* FA_D8 / FA_Tarboton (methods/flow_accumulation.hpp:27,16) */
void orc_fa_d8_f32_f64(const float *dem, int w, int h, float nodata, double *accum) {
  float *props = (float *)malloc(sizeof(float) * 9 * (size_t)w * h);
  orc_fm_d8_f32(dem, w, h, nodata, props);
  orc_flow_accumulation_props_f64(props, w, h, accum);
  free(props);
}

void orc_fa_tarboton_f32_f64(const float *dem, int w, int h, float nodata, double *accum) {
  float *props = (float *)malloc(sizeof(float) * 9 * (size_t)w * h);
  orc_fm_tarboton_f32(dem, w, h, nodata, props);
  orc_flow_accumulation_props_f64(props, w, h, accum);
  free(props);
}

/* ------------------------------------------------------------------------------------ */
/* f1  FM_D4 = FM_OCallaghan<D4> (flowmet/OCallaghan1984.hpp:13-77,89-91): the same scan  */
/* over the 4 cardinal neighbours (common/constants.hpp:53-54: 1=W 2=N 3=E 4=S); the       */
/* proportion lands in slot n of THAT numbering, as in the reference.                      */
static void orc_fm_d4_f32(const float *dem, int w, int h, float nodata, float *props) {
  static const int D4X[5] = {0, -1, 0, 1, 0};
  static const int D4Y[5] = {0, 0, -1, 0, 1};
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < 9 * n; i++) props[i] = NO_FLOW_GEN;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      if (dem[i] == nodata) {
        props[9 * i] = NO_DATA_GEN;
        continue;
      }
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) continue;
      const float e = dem[i];
      int lowest_n = 0;
      float lowest = 3.402823466e+38f;
      for (int k = 1; k <= 4; k++) {
        const float ne = dem[(size_t)(y + D4Y[k]) * w + (x + D4X[k])];
        if (ne == nodata) continue;
        if (ne >= e) continue;
        if (ne < lowest) {
          lowest = ne;
          lowest_n = k;
        }
      }
      if (lowest_n == 0) continue;
      props[9 * i] = HAS_FLOW_GEN;
      props[9 * i + lowest_n] = 1.0f;
    }
}

/* f1  FM_Holmgren (flowmet/Holmgren1994.hpp:13-83; FM_Quinn = exponent 1, Quinn1991.hpp:15) */
/* and FM_Freeman (flowmet/Freeman1991.hpp:13-80).  Holmgren sums the float-rounded powers  */
/* (`C += props(x,y,n)`, :65), Freeman the double ones (`C += cval`, :60).                  */
static void orc_fm_mfd_f32(const float *dem, int w, int h, float nodata, double xparam, int holmgren, float *props) {
  const double SQRT2 = 1.414213562373095048801688724209698078569671875376948; /* constants.hpp:34 */
  const double L1 = 0.5, L2 = 0.354;                                           /* Holmgren1994.hpp:25-26 */
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < 9 * n; i++) props[i] = NO_FLOW_GEN;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t i = (size_t)y * w + x;
      float *p = props + 9 * i;
      if (dem[i] == nodata) {
        p[0] = NO_DATA_GEN;
        continue;
      }
      if (x == 0 || y == 0 || x == w - 1 || y == h - 1) continue;
      const float e = dem[i];
      double C = 0;
      for (int k = 1; k <= 8; k++) {
        const float ne = dem[(size_t)(y + D8Y[k]) * w + (x + D8X[k])];
        if (ne == nodata) continue;
        if (ne < e) {
          const double rise = e - ne; /* float subtraction, then widened */
          const double run = (k & 1) ? 1.0 : SQRT2;
          const double grad = rise / run;
          if (holmgren) {
            p[k] = (float)pow(grad * ((k & 1) ? L1 : L2), xparam);
            C += p[k];
          } else {
            const double cval = pow(grad, xparam);
            p[k] = (float)cval;
            C += cval;
          }
        }
      }
      if (C > 0) {
        p[0] = HAS_FLOW_GEN;
        C = 1 / C;
        for (int k = 1; k <= 8; k++) p[k] = p[k] > 0 ? (float)(p[k] * C) : 0.0f;
      }
    }
}

/* method: 2 FM_D4, 3 FM_Holmgren (FM_Quinn: xparam 1), 4 FM_Freeman -- numbering of the C ABI */
void orc_fm_method_f32(int method, const float *dem, int w, int h, float nodata, double xparam, float *props) {
  if (method == 0) orc_fm_d8_f32(dem, w, h, nodata, props);
  else if (method == 1) orc_fm_tarboton_f32(dem, w, h, nodata, props);
  else if (method == 2) orc_fm_d4_f32(dem, w, h, nodata, props);
  else orc_fm_mfd_f32(dem, w, h, nodata, xparam, method == 3, props);
}

/* FA_D4 / FA_Quinn / FA_Holmgren / FA_Freeman (methods/flow_accumulation.hpp:28,19,18,20) */
void orc_fa_method_f32_f64(int method, const float *dem, int w, int h, float nodata, double xparam, double *accum) {
  float *props = (float *)malloc(sizeof(float) * 9 * (size_t)w * h);
  orc_fm_method_f32(method, dem, w, h, nodata, xparam, props);
  orc_flow_accumulation_props_f64(props, w, h, accum);
  free(props);
}

/* ------------------------------------------------------------------------------------ */
/* Terrain attributes (methods/terrain_attributes.hpp).  attribute: 0 slope rise/run (:239-248), 1 percentage   */
/* (:299-302), 2 degrees (:317-320), 3 radians (:308-311), 4 aspect (:222-236), 5 curvature (:254-258),           */
/* 6 planform (:264-272), 7 profile (:278-286) -- numbering of the C ABI.  TerrainProcessor (:336-354): NoData    */
/* cells get the output's NoData; TerrainSetup (:161-194): neighbours outside the raster or NoData take the       */
/* centre's value, everything scaled by zscale in double.  Built with -ffp-contract=off like the stock reference. */
static double ta_cell(int attribute, const float *dem, int w, int h, int x, int y, float nodata, float zscale, double lx,
                      double ly) {
  double z[9];
  int k = 0;
  for (int dy = -1; dy <= 1; dy++)
    for (int dx = -1; dx <= 1; dx++, k++) {
      const int nx = x + dx, ny = y + dy;
      double v = dem[(size_t)y * w + x];
      if (nx >= 0 && ny >= 0 && nx < w && ny < h && dem[(size_t)ny * w + nx] != nodata) v = dem[(size_t)ny * w + nx];
      z[k] = v * zscale;
    }
  const double a = z[0], b = z[1], c = z[2], d = z[3], e = z[4], f = z[5], g = z[6], hh = z[7], i = z[8];
  if (attribute <= 4) {
    const double dzdx = ((c + 2 * f + i) - (a + 2 * d + g)) / 8 / lx;
    const double dzdy = ((g + 2 * hh + i) - (a + 2 * b + c)) / 8 / ly;
    if (attribute == 4) {
      const double asp = 180.0 / M_PI * atan2(dzdy, -dzdx);
      if (asp < 0) return 90 - asp;
      if (asp > 90.0) return 360.0 - asp + 90.0;
      return 90.0 - asp;
    }
    const double rr = sqrt(dzdx * dzdx + dzdy * dzdy);
    if (attribute == 0) return rr;
    if (attribute == 1) return rr * 100;
    if (attribute == 2) return atan(rr) * 180 / M_PI;
    return atan(rr);
  }
  const double L = lx;
  const double D = ((d + f) / 2 - e) / L / L;
  const double E = ((b + hh) / 2 - e) / L / L;
  const double F = (-a + c + g - i) / 4 / L / L;
  const double G = (-d + f) / 2 / L;
  const double H = (b - hh) / 2 / L;
  if (attribute == 5) return -2 * (D + E) * 100;
  if (G == 0 && H == 0) return 0;
  if (attribute == 6) return -2 * (D * H * H + E * G * G - F * G * H) / (G * G + H * H) * 100;
  return 2 * (D * G * G + E * H * H + F * G * H) / (G * G + H * H) * 100;
}

void orc_terrain_attribute_f32(int attribute, const float *dem, int w, int h, float nodata_in, float nodata_out, float zscale,
                               double cell_x, double cell_y, float *out) {
#pragma omp parallel for
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const size_t ci = (size_t)y * w + x;
      out[ci] = dem[ci] == nodata_in ? nodata_out : (float)ta_cell(attribute, dem, w, h, x, y, nodata_in, zscale, cell_x, cell_y);
    }
}

/* ------------------------------------------------------------------------------------ */
/* Benchmark input: CPU restatement of the device terrain generator (richdem_b200/csrc/   */
/* terrain.cu, fbm_kernel -- this repository's own synthetic DEM, not a reference         */
/* function).  Same integer hash, same single-precision operation order, one rounding per */
/* operation (build with -ffp-contract=off), so rdb200_dev_generate_fbm_f32 and this      */
/* function produce identical bits: bench.py --impl reference times the reference on      */
/* exactly the raster the GPU arm uses without touching the GPU.                           */
static uint32_t fbm_hash3(uint32_t x, uint32_t y, uint32_t s) {
  uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u + 0x7F4A7C15u) ^ (s * 0xC2B2AE3Du);
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}
static float fbm_lattice(uint32_t ix, uint32_t iy, uint32_t s) {
  return (float)(fbm_hash3(ix, iy, s) >> 8) * (1.0f / 16777216.0f);
}
static float fbm_fade(float t) {
  const float inner = (t * (t * 6.f - 15.f)) + 10.f;
  return ((t * t) * t) * inner;
}
void orc_generate_fbm_f32(float *dem, int w, int h, int y0, uint32_t seed, int octaves, float quantum) {
  const int top_log2 = 12;
  if (octaves <= 0) octaves = 12;
#pragma omp parallel for schedule(static)
  for (int yl = 0; yl < h; yl++) {
    const int y = y0 + yl;
    for (int x = 0; x < w; x++) {
      float sum = 0.f, amp = 1.f, norm = 0.f;
      for (int o = 0; o < octaves; o++) {
        const int lg = top_log2 - o;
        if (lg < 1) break;
        const uint32_t cell = 1u << lg;
        const uint32_t ix = (uint32_t)x >> lg, iy = (uint32_t)y >> lg;
        float tx = (float)((uint32_t)x & (cell - 1)) / (float)cell;
        float ty = (float)((uint32_t)y & (cell - 1)) / (float)cell;
        tx = fbm_fade(tx);
        ty = fbm_fade(ty);
        const uint32_t s = seed * 131u + (uint32_t)o;
        const float v00 = fbm_lattice(ix, iy, s), v01 = fbm_lattice(ix + 1, iy, s);
        const float v10 = fbm_lattice(ix, iy + 1, s), v11 = fbm_lattice(ix + 1, iy + 1, s);
        const float top = v00 * (1.f - tx) + v01 * tx;
        const float bot = v10 * (1.f - tx) + v11 * tx;
        const float v = top * (1.f - ty) + bot * ty;
        sum = sum + amp * v;
        norm = norm + amp;
        amp = amp * 0.5946035575f;
      }
      float z = (1000.0f * sum) / norm;
      if (quantum > 0.f) z = rintf(z / quantum) * quantum;
      dem[(size_t)yl * w + x] = z;
    }
  }
}
