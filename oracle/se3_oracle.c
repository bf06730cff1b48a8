/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the integer / index part of the
 * reference hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this; the product (se3conv3d_b200/) never does.
 *
 * Each function follows the cited reference lines (paths relative to /root/reference/point_cloud_lib):
 *   oracle_compute_keys   custom_ops/ball_query/compute_keys.cu:33-72, grid_utils.cuh:56-93
 *   oracle_ball_query     custom_ops/ball_query/ball_query.cu:22-104, find_ranges_grid_ds.cu:40-166
 *                         (candidate cells), count_neighbors.cu:86 + math_helper.cuh:304-320 (predicate)
 *   oracle_knn_query      custom_ops/knn_query/knn_query.cu:18-197 (sweep, tie rule, pruning)
 * Float operations are written with explicit fmaf where nvcc contracts (verified against the SASS
 * of the reference build: FADD, FMUL, FFMA x3, IEEE sqrt, compare), and this file must be compiled
 * with -ffp-contract=off so gcc adds no contraction of its own.
 *
 * Pinning: checked against the reference CUDA ops (oracle/_ref) run on B200, frozen as fixtures in
 * tests/golden/ref_ops_*.npz by tests/golden/gen_ref_ops_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void cell_of(const float* p, const float* mn, const int* nc, const float* inv, int* c) {
  for (int d = 0; d < 3; ++d) {
    float rel = (p[d] - mn[d]) * inv[d];
    c[d] = clampi((int)floorf(rel), 0, nc[d] - 1);
  }
}

static int64_t key_of(const int* c, const int* nc, int b) {
  return (((int64_t)b * nc[0] + c[0]) * nc[1] + c[1]) * nc[2] + c[2];
}

void oracle_compute_keys(const float* pts, const int32_t* batch, int64_t n, const float* aabb_min,
                         const int32_t* num_cells, const float* cell_size, int64_t* keys) {
  float inv[3];
  for (int d = 0; d < 3; ++d) inv[d] = 1.0f / cell_size[d];
  for (int64_t i = 0; i < n; ++i) {
    int c[3];
    cell_of(pts + 3 * i, aabb_min + 3 * batch[i], num_cells, inv, c);
    keys[i] = key_of(c, num_cells, batch[i]);
  }
}

static int hit(const float* s, const float* p, const float* inv) {
  float dx = (s[0] - p[0]) * inv[0];
  float dy = (s[1] - p[1]) * inv[1];
  float dz = (s[2] - p[2]) * inv[2];
  float d2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, 0.0f)));
  return sqrtf(d2) < 1.0f;
}

/* Two passes: counts (neighbors == NULL) then fill.  Output rows are grouped by sample and, inside
 * a row, sorted by source index (the canonical order parity tests compare in).  Returns E. */
int64_t oracle_ball_query(const float* src, const float* dst, const int32_t* bsrc, const int32_t* bdst,
                          int64_t n, int64_t m, const float* min_pt, const int32_t* num_cells,
                          const float* radius, int64_t* neighbors, int32_t* ends) {
  float inv[3];
  for (int d = 0; d < 3; ++d) inv[d] = 1.0f / radius[d];
  int* cells = (int*)malloc(sizeof(int) * 3 * (size_t)(n > 0 ? n : 1));
  for (int64_t j = 0; j < n; ++j) cell_of(src + 3 * j, min_pt + 3 * bsrc[j], num_cells, inv, cells + 3 * j);
  int64_t e = 0;
  for (int64_t i = 0; i < m; ++i) {
    int c[3];
    cell_of(dst + 3 * i, min_pt + 3 * bdst[i], num_cells, inv, c);
    for (int64_t j = 0; j < n; ++j) {
      if (bsrc[j] != bdst[i]) continue;
      const int* cj = cells + 3 * j;
      /* candidate set: the 27 cells around the (clamped) sample cell */
      if (abs(cj[0] - c[0]) > 1 || abs(cj[1] - c[1]) > 1 || abs(cj[2] - c[2]) > 1) continue;
      if (!hit(dst + 3 * i, src + 3 * j, inv)) continue;
      if (neighbors) {
        neighbors[2 * e] = i;
        neighbors[2 * e + 1] = j;
      }
      ++e;
    }
    if (ends) ends[i] = (int32_t)e;
  }
  free(cells);
  return e;
}

/* ---- knn ---------------------------------------------------------------------------------- */
typedef struct { uint64_t key; int32_t idx; } sort_item;
static int cmp_item(const void* a, const void* b) {
  const sort_item* x = (const sort_item*)a;
  const sort_item* y = (const sort_item*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

static void knn_insert(float* bd, int* bi, int k, float d, int idx) {
  for (int e1 = 0; e1 < k; ++e1) {
    if (bd[e1] > d) {
      for (int e2 = k - 1; e2 > e1; --e2) { bd[e2] = bd[e2 - 1]; bi[e2] = bi[e2 - 1]; }
      bd[e1] = d; bi[e1] = idx;
      break;
    }
  }
}

/* out [n,k] int32 (-1 padded).  The ordering key is the exact (batch, coordinate) pair -- the
 * reference sorts a rounded float key (knn_query.cu:151-155); both give the same sweep order except
 * for float ties.  dist_out (optional, [n,k]) receives the squared distances for tie-aware checks. */
void oracle_knn_query(const float* pts, const int32_t* batch, int64_t n, int32_t k, int32_t* out, float* dist_out) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      if (pts[3 * i + d] < lo[d]) lo[d] = pts[3 * i + d];
      if (pts[3 * i + d] > hi[d]) hi[d] = pts[3 * i + d];
    }
  int sd = 0;
  float best = hi[0] - lo[0];
  for (int d = 1; d < 3; ++d)
    if (hi[d] - lo[d] > best) { best = hi[d] - lo[d]; sd = d; }
  sort_item* it = (sort_item*)malloc(sizeof(sort_item) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, pts + 3 * i + sd, 4);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    it[i].key = ((uint64_t)(uint32_t)batch[i] << 32) | u;
    it[i].idx = (int32_t)i;
  }
  qsort(it, (size_t)n, sizeof(sort_item), cmp_item);
  float* bd = (float*)malloc(sizeof(float) * k);
  int* bi = (int*)malloc(sizeof(int) * k);
  for (int64_t p = 0; p < n; ++p) {
    const int self = it[p].idx;
    const float* c = pts + 3 * (int64_t)self;
    for (int e = 0; e < k; ++e) { bd[e] = 1e10f; bi[e] = -1; }
    for (int dir = 0; dir < 2; ++dir) {
      for (int64_t q = dir == 0 ? p : p - 1; q >= 0 && q < n; q += dir == 0 ? 1 : -1) {
        const int j = it[q].idx;
        if (batch[j] != batch[self]) break;
        const float* x = pts + 3 * (int64_t)j;
        float d0 = x[0] - c[0], d1 = x[1] - c[1], d2 = x[2] - c[2];
        float dist = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
        knn_insert(bd, bi, k, dist, j);
        float s = x[sd] - c[sd];
        if (bd[k - 1] < s * s) break;
      }
    }
    for (int e = 0; e < k; ++e) {
      out[(int64_t)self * k + e] = bi[e];
      if (dist_out) dist_out[(int64_t)self * k + e] = bd[e];
    }
  }
  free(bd); free(bi); free(it);
}
