// Fused SE(3) group convolution for sm_100a: geometry -> basis -> aggregation -> projection in ONE persistent,
// warp-specialised kernel whose every matrix product runs on tcgen05 with accumulators in tensor memory.
// The [K x Cin] tile T of a row never reaches HBM (unless the caller asks for a copy for the weight gradient).
//
// Reference semantics: layers/PNEConvLayerRotEquiv.py:160-216 (g -> matmul + bias -> GELU -> FeatBasisProj -> einsum),
// custom_ops/feature_aggregation/feat_basis_proj.cu:55-118 (T[r,c,k] = sum_e x[src(e),c] h[e,k]); transposed mode =
// the data gradient of the same sums over the transposed CSR (feat_basis_proj_grads.cu:129-140 without atomics).
//
// Work decomposition.  A CTA (one per SM) owns a contiguous range of row points with an equal share of the CSR
// entries.  The (edge x gathered frame) entries of a row point are padded to GROUPS of 16 "slots"; eight consecutive
// groups form a BATCH of 128 slots.  Per batch:
//   producers (4 warps, one thread per slot)  gather the 48-byte neighbour record and the bf16 feature row of the
//            slot with cp.async straight into the MN-major SWIZZLE_128B operand tile X, evaluate the 9-vector g for
//            every row frame and write it (tf32, plus a constant 1 that carries the bias) into the K-major tile G
//   MMA1    pre[slot, (a,k)] = G . Wext          tcgen05.mma kind::tf32, M = 128 slots, N = 32 per row frame  -> TMEM
//   activation (4 warps, thread = slot = TMEM lane)  tcgen05.ld pre, h = act(pre), bf16 -> tile H (aliases G),
//            MN-major: one 128-byte line per slot = the 64 (a,k) values, i.e. the line layout MMA2 wants
//   MMA2    T^T[(a,k), c] (+)= sum over the 16 slots of a group H[slot,(a,k)] X[slot,c]   kind::f16, M = 128 (64 used),
//            N = C, one instruction per group, accumulating over the groups of a row point in a TMEM ring
//   transposers (warp a = TMEM lane quadrant a)  tcgen05.ld T^T, bf16, 64-byte pieces into the K-major tile TT
//            [16 rows (point, frame)] x [K3 = 32 * C in (k,c) order]  (+ optional copy to global for dW)
//   MMA3    y^T[o, rows] = W3^T . TT^T           kind::f16, M = 128 (Cout used), N = 16 rows, K3 / 16 steps; W3 (the
//            conv weights in (k,c) order, pre-swizzled image) is resident in shared memory, loaded once per CTA by
//            the TMA engine (cp.async.bulk)
//   epilogue tcgen05.ld y^T, scale, coalesced 128-byte row stores.
// All hand-offs are mbarriers; MMA completion is signalled with tcgen05.commit.  No block-wide barrier inside the loop.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <vector>
#include "conv_fused.cuh"
#include "tc_common.cuh"
#include "umma.cuh"

namespace se3 {

using namespace umma;

namespace {

constexpr int NX = 4;    // X / record ring depth (batches)
constexpr int NG = 3;    // G/H ring depth (= pre buffers in TMEM)
constexpr int NT = 8;    // T^T accumulators in TMEM (NT * CP columns; 4 for CP = 64)
constexpr int NR = 16;   // rows (point, frame) of a projection tile
constexpr int NPS = 2;   // producer warp sets (set s owns batches s, s + NPS, ...)
constexpr int NAS = 2;   // activation warp sets
constexpr int NTS = 2;   // transposer / epilogue warp sets (set s owns row points s, s + NTS, ...)
constexpr int TILE16K = 128 * 128;
constexpr int REC_STAGE = 128 * 48;
constexpr int RR_STAGE = 8 * 2 * 48;   // row-point records of the 8 groups of a batch (up to 2 row frames)
constexpr int PRE_COL = 0, T_COL = NG * 64, D3_COL = 480;
constexpr int NTHREADS = 896;
// warp roles
constexpr int NI2 = 2;   // MMA2 issuer warps (issuer j owns the accumulators tb with tb % NI2 == j)
constexpr int W_ACT = 4 * NPS, W_TE = W_ACT + 4 * NAS, W_I1 = W_TE + 4 * NTS, W_I2 = W_I1 + 1, W_I3 = W_I2 + NI2;
static_assert((W_I3 + 1) * 32 == NTHREADS, "warp roles must cover the block");
static_assert(W_ACT % 4 == 0 && W_TE % 4 == 0, "TMEM lane quadrant = warp % 4");

#ifdef SE3_FUSED_TRACE
// Pipeline trace (debug builds only): lane 0 of every warp of block 0 appends (event, index, clock) triples to its own
// slice of a global buffer (no atomics: a trace point costs a clock read and three stores); the launcher dumps the
// buffer to /tmp/se3_fused_trace.txt.
__device__ long long* g_trace;
constexpr int TRACE_PER_WARP = 2048;
#define TRACE_DECL int trace_cnt = 0;
#define TRACE(ev, idx)                                                              \
  do {                                                                              \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_cnt < TRACE_PER_WARP) { \
      long long* _p = g_trace + 3 * ((size_t)(threadIdx.x >> 5) * TRACE_PER_WARP + trace_cnt++); \
      _p[0] = (ev); _p[1] = (idx); _p[2] = clock64();                               \
    }                                                                               \
  } while (0)
#else
#define TRACE_DECL
#define TRACE(ev, idx) do {} while (0)
#endif

struct Bars {
  uint64_t g_full[NG], pre_full[NG], h_full[NG], gh_free[NG], x_free[NX], t_full[NT], t_empty[NT], tt_full, tt_empty,
      d3_full[2], d3_empty[2], w_full;
};

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// first p in [0, n] with start(p) >= target, start(p) = p ? row_ends[p-1] : 0 (non-decreasing); 33-ary warp search
__device__ int warp_lower_bound(const int* __restrict__ row_ends, int n, int target, int lane) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int span = hi - lo;
    const int step = (span + 32) / 33;
    const int p = lo + lane * step;
    bool ge = true;
    if (p < hi) ge = (p ? __ldg(row_ends + p - 1) : 0) >= target;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    if (m == 0) {
      lo = lo + 31 * step + 1;
    } else {
      const int f = __ffs(m) - 1;
      const int nhi = lo + f * step;
      if (f > 0) lo = lo + (f - 1) * step + 1;
      hi = nhi < hi ? nhi : hi;
    }
  }
  return lo;
}

// Sliding window of 32 row points: lane i <-> point wbase + i.  Maps a group index of the CTA to its row point.
struct Window {
  int wbase, gbase, gtot;  // warp-uniform: first point of the window, first group of the window, groups in the window
  int lo, cnt, goff, gn;   // per lane: CSR start, entries (edges x gathered frames), first group (relative), groups
};
__device__ __forceinline__ void window_load(Window& w, const int* __restrict__ row_ends, int p1, int fg, int lane) {
  const int p = w.wbase + lane;
  const bool valid = p < p1;
  const int hi = valid ? __ldg(row_ends + p) : 0;
  int lo = __shfl_up_sync(0xffffffffu, hi, 1);
  if (lane == 0) lo = (valid && p > 0) ? __ldg(row_ends + p - 1) : 0;
  w.lo = lo;
  w.cnt = valid ? (hi - lo) * fg : 0;
  w.gn = valid ? max(1, (w.cnt + 15) >> 4) : 0;
  int s = w.gn;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += v;
  }
  w.goff = s - w.gn;
  w.gtot = __shfl_sync(0xffffffffu, s, 31);
}
// group G (CTA-relative, warp-uniform, must be < total groups of the CTA): point, group index inside the point, ...
__device__ __forceinline__ void window_locate(Window& w, const int* __restrict__ row_ends, int p1, int fg, int lane, int G,
                                              int& point, int& j, int& lo, int& cnt, int& gn) {
  while (G >= w.gbase + w.gtot) {
    w.gbase += w.gtot;
    w.wbase += 32;
    window_load(w, row_ends, p1, fg, lane);
  }
  const unsigned m = __ballot_sync(0xffffffffu, w.gn > 0 && w.gbase + w.goff <= G);
  const int l = 31 - __clz((int)m);
  point = w.wbase + l;
  j = G - (w.gbase + __shfl_sync(0xffffffffu, w.goff, l));
  lo = __shfl_sync(0xffffffffu, w.lo, l);
  cnt = __shfl_sync(0xffffffffu, w.cnt, l);
  gn = __shfl_sync(0xffffffffu, w.gn, l);
}

template <bool TR>
__device__ __forceinline__ void geometry9f(const float (&Frow)[9], const float (&Fq)[9], float dx, float dy, float dz,
                                           float (&g)[9]) {
  const float(&Ro)[9] = TR ? Fq : Frow;
  const float(&Ri)[9] = TR ? Frow : Fq;
#pragma unroll
  for (int c = 0; c < 3; ++c) g[c] = dx * Ro[c] + dy * Ro[3 + c] + dz * Ro[6 + c];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 3; ++n) g[3 + 3 * m + n] = Ro[m] * Ri[n] + Ro[3 + m] * Ri[3 + n] + Ro[6 + m] * Ri[6 + n];
}

template <int ACT>
__device__ __forceinline__ float act_sel(float x, int act) {
  if (ACT >= 0) return act_fast<ACT>(x);
  switch (act) {
    case 1: return act_fast<1>(x);
    case 2: return act_fast<2>(x);
    case 3: return act_fast<3>(x);
    default: return x;
  }
}

}  // namespace

// CP: gathered channels padded to 16 / 32 / 64 (the N of MMA2, a row of TT holds 32 * CP values);
// FR: row frames (1 or 2); TR: transposed (data gradient) geometry; ACT: activation (-1 = runtime switch)
template <int CP, int FR, bool TR, int ACT>
__global__ void __launch_bounds__(NTHREADS, 1) k_conv_fused(const FusedArgs a) {
  constexpr int NKB = CP / 2;                 // 64-element k-blocks of K3 = 32 * CP
  constexpr int TT_BYTES = NKB * NR * 128;
  constexpr int PT = NR / FR;                 // row points per projection tile
  // feature rows of up to 64 bytes share a 128-byte operand line with the batch of the neighbouring ring slot
  constexpr int XPT = CP <= 32 ? 2 : 1;       // batches per 16 KB X tile
  constexpr int NXT = NX / XPT;               // X tiles
  constexpr int NTC = CP <= 32 ? NT : 4;      // T^T accumulators (TMEM columns T_COL .. T_COL + NTC * CP <= D3_COL)
  static_assert(T_COL + NTC * CP <= D3_COL, "tensor memory budget");
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) Bars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ int range_s[4];                  // p0, p1, groups of this CTA
  __shared__ __align__(16) int ginfo[NX][8];                // per group of a batch in flight: valid | first << 1 | last << 2 | ordinal << 8
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TRACE_DECL
  const uint32_t raw = smem_addr(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gen = smem_dyn + (base - raw);       // generic pointer to the aligned base
  const bool proj = a.w3img != nullptr;
  const int cop = (a.co + 7) & ~7;
  const uint32_t w3_bytes = proj ? (uint32_t)(NKB * cop * 128) : 0u;
  const uint32_t off_gh = (w3_bytes + 1023u) & ~1023u;
  const uint32_t off_x = off_gh + NG * TILE16K;
  const uint32_t off_tt = off_x + NXT * TILE16K;
  const uint32_t off_wx = off_tt + (proj ? TT_BYTES : 0);
  const uint32_t off_rec = off_wx + 4096;
  const uint32_t off_rr = off_rec + NX * REC_STAGE;
  const int fg = a.f_g;

  // ---- range of this CTA: row points with an equal share of the CSR entries (only static geometry is read here,
  //      so this runs ahead of the grid dependency)
  if (warp == 0) {
    const int n = (int)a.n_rows;
    const int64_t e = a.n_edges;
    const int nb = gridDim.x, b = blockIdx.x;
    const int p0 = b == 0 ? 0 : warp_lower_bound(a.row_ends, n, (int)(e * b / nb), lane);
    const int p1 = b == nb - 1 ? n : warp_lower_bound(a.row_ends, n, (int)(e * (b + 1) / nb), lane);
    int g = 0;
    for (int p = p0 + lane; p < p1; p += 32) {
      const int hi = __ldg(a.row_ends + p), lo = p ? __ldg(a.row_ends + p - 1) : 0;
      g += max(1, ((hi - lo) * fg + 15) >> 4);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
    if (lane == 0) {
      range_s[0] = p0; range_s[1] = p1; range_s[2] = g;
    }
  }
  __syncthreads();
  const int p0 = range_s[0], p1 = range_s[1], n_groups = range_s[2];
  const int np = p1 - p0;
  if (np <= 0) return;
  const int n_batches = (n_groups + 7) >> 3;
  const int n_tiles = (np + PT - 1) / PT;

  auto B = [&](const uint64_t& b) { return smem_addr(&b); };
  if (warp == W_I1 && lane == 0) {
    for (int i = 0; i < NG; ++i) {
      mbar_init(B(bars.g_full[i]), 128);
      mbar_init(B(bars.pre_full[i]), 1);
      mbar_init(B(bars.h_full[i]), 128);
      mbar_init(B(bars.gh_free[i]), NI2);
    }
    for (int i = 0; i < NX; ++i) mbar_init(B(bars.x_free[i]), NI2);
    for (int i = 0; i < NT; ++i) {
      mbar_init(B(bars.t_full[i]), 1);
      mbar_init(B(bars.t_empty[i]), FR);
    }
    mbar_init(B(bars.tt_full), FR * NTS);
    mbar_init(B(bars.tt_empty), 1);
    const int n_epi = (a.co + 31) >> 5;
    for (int i = 0; i < 2; ++i) {
      mbar_init(B(bars.d3_full[i]), 1);
      mbar_init(B(bars.d3_empty[i]), n_epi);
    }
    mbar_init(B(bars.w_full), 1);
    mbar_init_fence();
  }
  if (warp == W_I2) tmem_alloc(smem_addr(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();       // everything below reads tensors the previous kernels of the stream wrote
  pdl_trigger();

  // ---- Wext tile: B operand of MMA1, K-major SW128, row k = [scale * proj_axes[0..8][k], scale * bias[k], 0 x 6]
  if (warp == W_I1) {
    const float sc = act_pre_scale(a.act);
    const int k = lane;
    float v[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) v[d] = d < 9 ? sc * __ldg(a.w9 + d * 32 + k) : (d == 9 ? sc * __ldg(a.bias + k) : 0.0f);
    unsigned char* row = gen + off_wx + (k >> 3) * 1024 + (k & 7) * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 q = j < 4 ? make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(row + ((j ^ (k & 7)) << 4)) = q;
    }
    fence_proxy_async();
  }
  // ---- W3 image (resident): one thread drives the TMA engine
  if (proj && warp == W_I3 && lane == 0) {
    mbar_arrive_expect_tx(B(bars.w_full), w3_bytes);
    for (uint32_t o = 0; o < w3_bytes; o += 16384u) {
      const uint32_t n = min(16384u, w3_bytes - o);
      bulk_g2s(base + o, a.w3img + o, n, B(bars.w_full));
    }
  }
  __syncthreads();

  // X operand line of a slot of batch i: tile, and the 64-byte half of the 128-byte line when two batches share a tile
  auto x_line = [&](int i, int slot) -> uint32_t {
    const int xs = i % NX;
    return off_x + (xs / XPT) * TILE16K + (slot >> 3) * 1024 + (slot & 7) * 128;
  };

  if (warp < W_ACT) {
    // =================================================================================== producers
    const int set = warp >> 2, wq = warp & 3;
    const int gi0 = 2 * wq, half = lane >> 4, e16 = lane & 15;
    const int slot = 32 * wq + lane;
    Window w;
    w.wbase = p0; w.gbase = 0;
    window_load(w, a.row_ends, p1, fg, lane);
    // slot info of a batch: gathered row (or -1), row point of the slot's group (or -1: padding group), and the group's
    // descriptor word for the MMA2 issuer
    auto prefetch = [&](int b, int& grow, int& point, int& gword) {
      int pts[2], js[2], los[2], cnts[2], gns[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int G = 8 * b + gi0 + h;
        pts[h] = -1; js[h] = 0; los[h] = 0; cnts[h] = 0; gns[h] = 0;
        if (G < n_groups) window_locate(w, a.row_ends, p1, fg, lane, G, pts[h], js[h], los[h], cnts[h], gns[h]);
      }
      point = half ? pts[1] : pts[0];
      const int j = half ? js[1] : js[0], lo = half ? los[1] : los[0], cnt = half ? cnts[1] : cnts[0];
      const int gn = half ? gns[1] : gns[0];
      // word for the MMA2 issuer: valid | first << 1 | last << 2 | must wait for the accumulator << 3 | parity of that
      // wait << 4 | accumulator index << 8
      {
        const int n = point - p0, u = n / NTC;
        gword = point >= 0 ? (1 | ((j == 0) << 1) | ((j == gn - 1) << 2) | ((j == 0 && u > 0) << 3) | (((u - 1) & 1) << 4) |
                              ((n - u * NTC) << 8))
                           : 0;
      }
      const int ent = 16 * j + e16;
      grow = -1;
      if (point >= 0 && ent < cnt) {
        int e, f;
        if (fg == 1) { e = ent; f = 0; }
        else if (fg == 2) { e = ent >> 1; f = ent & 1; }
        else if (fg == 4) { e = ent >> 2; f = ent & 3; }
        else { e = ent / 3; f = ent - 3 * e; }
        grow = __ldg(a.nbr + lo + e) * fg + f;
      }
    };
    auto issue = [&](int b, int grow, int point, int gword) {
      const int xs = b % NX;
      const bool ok = grow >= 0;
      // neighbour record -> rec[xs][slot][48 B]
      const float* rsrc = a.rec_g + (ok ? (size_t)grow * 12 : 0);
      const uint32_t rdst = base + off_rec + xs * REC_STAGE + slot * 48;
      cp16(rdst, rsrc, ok);
      cp16(rdst + 16, rsrc + 4, ok);
      cp16(rdst + 32, rsrc + 8, ok);
      // feature row -> line `slot` of the MN-major tile X
      const __nv_bfloat16* xsrc = a.feat + (ok ? (size_t)grow * a.cs : 0);
      const uint32_t xdst = base + x_line(b, slot);
      const int ch0 = (XPT == 2) ? 4 * (xs & 1) : 0;
#pragma unroll
      for (int j = 0; j < CP / 8; ++j) {
        const bool cok = ok && (j * 8 < a.cs);
        cp16(xdst + (((ch0 + j) ^ (slot & 7)) << 4), xsrc + (cok ? j * 8 : 0), cok);
      }
      if (CP == 16 && true) {   // the MMA reads 16 channels only; nothing else to fill
      }
      // row-point records of the group (FR x 48 B), three 16-byte pieces per frame, by the first lanes of the half
      if (e16 < FR * 3) {
        const bool pok = point >= 0;
        const float* src = a.rec_row + (pok ? (size_t)point * (FR * 12) : 0) + e16 * 4;
        cp16(base + off_rr + xs * RR_STAGE + (gi0 + half) * (2 * 48) + e16 * 16, src, pok);
      }
      if (e16 == 0) ginfo[xs][gi0 + half] = gword;
    };
    int grow1 = -1, point1 = -1, gword1 = 0, grow2 = -1, point2 = -1, gword2 = 0;
    bool valid = false;   // is this thread's slot of the batch being finished a real entry?
    if (set < n_batches) {
      int g0, q0, gw0;
      prefetch(set, g0, q0, gw0);
      issue(set, g0, q0, gw0);
      cp_async_commit();
      valid = g0 >= 0;
      if (set + NPS < n_batches) prefetch(set + NPS, grow1, point1, gword1);
    }
    for (int i = set; i < n_batches; i += NPS) {
      if (i + 2 * NPS < n_batches) prefetch(i + 2 * NPS, grow2, point2, gword2);
      cp_async_wait<0>();     // batch i of this thread has landed (the only group in flight)
      __syncwarp();           // ... and the row records copied by the other lanes of the half warp
      if (wq == 0) TRACE(2, i);
      {
        const int u = i / NG;
        if (u > 0) mbar_wait(B(bars.gh_free[i % NG]), (uint32_t)((u - 1) & 1));
      }
      if (wq == 0) TRACE(3, i);
      const int xs = i % NX, gs = i % NG;
      const float4* rp = reinterpret_cast<const float4*>(gen + off_rec + xs * REC_STAGE + slot * 48);
      const float4 r0 = rp[0], r1 = rp[1], r2 = rp[2];
      const float4* rr = reinterpret_cast<const float4*>(gen + off_rr + xs * RR_STAGE + (gi0 + half) * (2 * 48));
      const float4 q0 = rr[0], q1 = rr[1], q2 = rr[2];
      const float dx = (TR ? (q0.x - r0.x) : (r0.x - q0.x)) * a.norm;
      const float dy = (TR ? (q0.y - r0.y) : (r0.y - q0.y)) * a.norm;
      const float dz = (TR ? (q0.z - r0.z) : (r0.z - q0.z)) * a.norm;
      const float Fq[9] = {r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
      unsigned char* grow_p = gen + off_gh + gs * TILE16K + (slot >> 3) * 1024 + (slot & 7) * 128;
#pragma unroll
      for (int f = 0; f < FR; ++f) {
        float4 s0 = q0, s1 = q1, s2 = q2;
        if (f == 1) { s0 = rr[3]; s1 = rr[4]; s2 = rr[5]; }
        const float Frow[9] = {s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
        float g[9];
        geometry9f<TR>(Frow, Fq, dx, dy, dz, g);
        if (!valid) {
#pragma unroll
          for (int q = 0; q < 9; ++q) g[q] = 0.0f;
        }
        const float one = valid ? 1.0f : 0.0f;
        const float4 c0 = make_float4(g[0], g[1], g[2], g[3]), c1 = make_float4(g[4], g[5], g[6], g[7]);
        const float4 c2 = make_float4(g[8], one, 0.f, 0.f), c3 = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(grow_p + (((4 * f + 0) ^ (slot & 7)) << 4)) = c0;
        *reinterpret_cast<float4*>(grow_p + (((4 * f + 1) ^ (slot & 7)) << 4)) = c1;
        *reinterpret_cast<float4*>(grow_p + (((4 * f + 2) ^ (slot & 7)) << 4)) = c2;
        *reinterpret_cast<float4*>(grow_p + (((4 * f + 3) ^ (slot & 7)) << 4)) = c3;
      }
      fence_proxy_async();
      mbar_arrive(B(bars.g_full[gs]));
      if (wq == 0) TRACE(4, i);
      // the next batch of this set: its ring slot was last used NX batches earlier
      if (i + NPS < n_batches) {
        const int u = (i + NPS) / NX;
        if (u > 0) mbar_wait(B(bars.x_free[(i + NPS) % NX]), (uint32_t)((u - 1) & 1));
        if (wq == 0) TRACE(1, i + NPS);
        issue(i + NPS, grow1, point1, gword1);
      }
      cp_async_commit();
      valid = grow1 >= 0;
      grow1 = grow2; point1 = point2; gword1 = gword2;
    }
    cp_async_wait<0>();
  } else if (warp < W_TE) {
    // =================================================================================== activation
    const int set = (warp - W_ACT) >> 2, q = warp & 3;
    const int slot = 32 * q + lane;
    for (int i = set; i < n_batches; i += NAS) {
      const int gs = i % NG;
      mbar_wait(B(bars.pre_full[gs]), (uint32_t)((i / NG) & 1));
      tc_fence_after();
      if (q == 0) TRACE(7, i);
      unsigned char* hrow = gen + off_gh + gs * TILE16K + (slot >> 3) * 1024 + (slot & 7) * 128;
#pragma unroll
      for (int f = 0; f < FR; ++f) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(PRE_COL + gs * 64 + f * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 p;
          p.x = pack_bf16(act_sel<ACT>(__uint_as_float(r[8 * j + 0]), a.act), act_sel<ACT>(__uint_as_float(r[8 * j + 1]), a.act));
          p.y = pack_bf16(act_sel<ACT>(__uint_as_float(r[8 * j + 2]), a.act), act_sel<ACT>(__uint_as_float(r[8 * j + 3]), a.act));
          p.z = pack_bf16(act_sel<ACT>(__uint_as_float(r[8 * j + 4]), a.act), act_sel<ACT>(__uint_as_float(r[8 * j + 5]), a.act));
          p.w = pack_bf16(act_sel<ACT>(__uint_as_float(r[8 * j + 6]), a.act), act_sel<ACT>(__uint_as_float(r[8 * j + 7]), a.act));
          *reinterpret_cast<uint4*>(hrow + (((4 * f + j) ^ (slot & 7)) << 4)) = p;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(B(bars.h_full[gs]));
      if (q == 0) TRACE(8, i);
    }
  } else if (warp < W_I1) {
    // =================================================================================== transposers + epilogue
    const int q = warp & 3, tset = (warp - W_TE) >> 2;
    const bool do_t = q < FR;
    const bool do_e = proj && tset == 0 && (32 * q < a.co);
    const int K3 = 32 * CP;
    auto epilogue = [&](int t) {
      const int db = t & 1;
      mbar_wait(B(bars.d3_full[db]), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(D3_COL + db * NR), r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(B(bars.d3_empty[db]));
      const int o = 32 * q + lane;
      const int row0 = (p0 + t * PT) * FR;
      const int rows = min(NR, (p1 - (p0 + t * PT)) * FR);
      if (o < a.co) {
#pragma unroll
        for (int rr = 0; rr < NR; ++rr)
          if (rr < rows) a.out[(size_t)(row0 + rr) * a.co + o] = a.out_scale * __uint_as_float(r[rr]);
      }
    };
    for (int t = 0; t < n_tiles; ++t) {
      if (do_t) {
        if (proj && t > 0) mbar_wait(B(bars.tt_empty), (uint32_t)((t - 1) & 1));
        const int n_end = min(np, (t + 1) * PT);
        for (int n = t * PT + tset; n < n_end; n += NTS) {
          const int tb = n % NTC;
          mbar_wait(B(bars.t_full[tb]), (uint32_t)((n / NTC) & 1));
          tc_fence_after();
          if (q == 0) TRACE(11, n);
          uint32_t v[CP];
          if constexpr (CP >= 32) {
#pragma unroll
            for (int h = 0; h < CP / 32; ++h) {
              uint32_t r[32];
              tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(T_COL + tb * CP + 32 * h), r);
#pragma unroll
              for (int c = 0; c < 32; ++c) v[32 * h + c] = r[c];
            }
          } else {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(T_COL + tb * CP), r);
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = r[c];
          }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(B(bars.t_empty[tb]));
          if (q == 0) TRACE(14, n);
          // bf16 pairs (c, c+1); this thread's values are K3 positions lane * CP .. lane * CP + CP - 1
          uint32_t pk[CP / 2];
#pragma unroll
          for (int c = 0; c < CP / 2; ++c) pk[c] = pack_bf16(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
          const int r = (n - t * PT) * FR + q;     // row of the tile
          const size_t row_g = (size_t)(p0 + n) * FR + q;
#pragma unroll
          for (int j = 0; j < CP / 8; ++j) {
            const uint4 p = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            if (proj) {
              const int kap = lane * CP + 8 * j;
              const int kb = kap >> 6, ch = (kap >> 3) & 7;
              *reinterpret_cast<uint4*>(gen + off_tt + kb * (NR * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4)) = p;
            }
            if (a.t_save) *reinterpret_cast<uint4*>(a.t_save + row_g * K3 + lane * CP + 8 * j) = p;
          }
          if (q == 0) TRACE(15, n);
        }
        if (proj) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(B(bars.tt_full));
        }
      }
      if (do_e && t > 0) epilogue(t - 1);
    }
    if (do_e) epilogue(n_tiles - 1);
  } else if (warp == W_I1) {
    // =================================================================================== MMA1 issuer
    constexpr uint32_t ID1 = idesc(FMT_TF32, FMT_TF32, 128, 32, false, false);
    if (lane == 0) {
      const uint64_t gd0 = desc_kmajor_sw128(base + off_gh), wd0 = desc_kmajor_sw128(base + off_wx);
      int gs = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_batches; ++i) {
        mbar_wait(B(bars.g_full[gs]), ph);
        tc_fence_after();
        TRACE(5, i);
        const uint64_t gd = gd0 + (uint64_t)(gs * (TILE16K >> 4));
#pragma unroll
        for (int f = 0; f < FR; ++f)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)   // + 64 B per row frame, + 32 B per k-step (descriptor address unit = 16 B)
            mma_tf32(tmem + (uint32_t)(PRE_COL + gs * 64 + f * 32), gd + (uint64_t)(4 * f + 2 * ks), wd0 + (uint64_t)(2 * ks), ID1,
                     ks ? 1u : 0u);
        commit(B(bars.pre_full[gs]));
        TRACE(6, i);
        if (++gs == NG) { gs = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp < W_I3) {
    // =================================================================================== MMA2 issuers (one thread each)
    constexpr uint32_t ID2 = idesc(FMT_BF16, FMT_BF16, 128, CP, true, true);
    const int mine = warp - W_I2;
    if (lane == 0) {
      int gs = 0, xs = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_batches; ++i) {
        mbar_wait(B(bars.g_full[gs]), ph);    // group descriptors of the batch are visible
        mbar_wait(B(bars.h_full[gs]), ph);
        tc_fence_after();
        if (mine == 0) TRACE(9, i);
        const int4 w0 = *reinterpret_cast<const int4*>(&ginfo[xs][0]), w1 = *reinterpret_cast<const int4*>(&ginfo[xs][4]);
        const int gw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const uint64_t hd = desc_mnmajor_sw128(base + off_gh + gs * TILE16K, NG * TILE16K);
        const uint64_t xd = desc_mnmajor_sw128(base + off_x + (xs / XPT) * TILE16K + ((XPT == 2) ? 64 * (xs & 1) : 0), 8192);
#pragma unroll
        for (int gi = 0; gi < 8; ++gi) {
          const int g = gw[gi];
          const int tb = (g >> 8) & 0xff;
          if ((g & 1) && (tb % NI2) == mine) {
            if (g & 8) {
              mbar_wait(B(bars.t_empty[tb]), (uint32_t)((g >> 4) & 1));
              tc_fence_after();
            }
            mma_f16(tmem + (uint32_t)(T_COL + tb * CP), hd + (uint64_t)(gi * 128), xd + (uint64_t)(gi * 128), ID2, (g & 2) ? 0u : 1u);
            if (g & 4) commit(B(bars.t_full[tb]));
          }
        }
        commit(B(bars.gh_free[gs]));
        commit(B(bars.x_free[xs]));
        if (mine == 0) TRACE(10, i);
        if (++gs == NG) { gs = 0; ph ^= 1u; }
        if (++xs == NX) xs = 0;
      }
    }
    __syncwarp();
  } else if (warp == W_I3) {
    // =================================================================================== MMA3 issuer
    if (proj && lane == 0) {
      constexpr uint32_t ID3 = idesc(FMT_BF16, FMT_BF16, 128, NR, false, false);
      mbar_wait(B(bars.w_full), 0u);
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(B(bars.tt_full), (uint32_t)(t & 1));
        if (t >= 2) mbar_wait(B(bars.d3_empty[t & 1]), (uint32_t)(((t >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t d3 = tmem + (uint32_t)(D3_COL + (t & 1) * NR);
        uint64_t wd = desc_kmajor_sw128(base), td = desc_kmajor_sw128(base + off_tt);
        const uint64_t wstep = (uint64_t)((cop * 128) >> 4), tstep = (uint64_t)((NR * 128) >> 4);
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_f16(d3, wd + 2 * ks, td + 2 * ks, ID3, (kb | ks) ? 1u : 0u);
          wd += wstep;
          td += tstep;
        }
        commit(B(bars.tt_empty));
        commit(B(bars.d3_full[t & 1]));
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_I2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// W3 image: the conv weights as the A operand of MMA3, [K3 / 64 k-blocks][COP rows][64 values], K3 index = k * CP + c,
// rows = output channel of the projection, every 128-byte row pre-swizzled (chunk ^= row % 8) so that a plain bulk
// copy drops it into shared memory in the K-major SWIZZLE_128B layout.
//   forward:     W3[(k,c), o ] = W[c ][k][o]      (c < Cin,  o < Cout)
//   transposed:  W3[(k,c), o'] = W[o'][k][c]      (c < Cout, o' < Cin)
// plain != 0: the same values as a row-major [co][K3] matrix (operand of the stand-alone projection GEMM).
__global__ void k_w3_image(const float* __restrict__ w, int c_in, int c_out, int tr, int cp, int cop, int plain,
                           __nv_bfloat16* __restrict__ img) {
  pdl_wait();
  const int c_g = tr ? c_out : c_in, co = tr ? c_in : c_out;
  const int64_t total = (int64_t)(cp / 2) * cop * 64;  // (32 * cp / 64) k-blocks * cop rows * 64 values
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i & 63);
    const int row = (int)((i >> 6) % cop);
    const int kb = (int)((i >> 6) / cop);
    const int kap = kb * 64 + col, k = kap / cp, c = kap - k * cp;
    float v = 0.0f;
    if (row < co && c < c_g) v = tr ? w[((int64_t)row * 32 + k) * c_out + c] : w[((int64_t)c * 32 + k) * c_out + row];
    const int chunk = col >> 3;
    if (plain) {
      if (row < co) img[(int64_t)row * (32 * cp) + kap] = __float2bfloat16(v);
    } else {
      img[((int64_t)kb * cop + row) * 64 + ((chunk ^ (row & 7)) << 3) + (col & 7)] = __float2bfloat16(v);
    }
  }
}

size_t fused_w3_bytes(int c_gathered, int co) {
  const int cp = fused_cp(c_gathered), cop = (co + 7) & ~7;
  return cp ? (size_t)(cp / 2) * cop * 128 : 0;
}

int launch_w3_image(const float* w, int c_in, int c_out, bool tr, bool plain, __nv_bfloat16* img, cudaStream_t st) {
  const int c_g = tr ? c_out : c_in, co = tr ? c_in : c_out;
  const int cp = fused_cp(c_g), cop = (co + 7) & ~7;
  const int64_t total = (int64_t)(cp / 2) * cop * 64;
  int blocks = (int)std::min<int64_t>((total + 255) / 256, 4 * num_sms());
  SE3_CUDA(launch_pdl(k_w3_image, dim3(blocks), dim3(256), 0, st, w, c_in, c_out, tr ? 1 : 0, cp, cop, plain ? 1 : 0, img));
  SE3_LAUNCH_CHECK();
  return SE3_OK;
}

static size_t fused_smem_bytes(int cp, int co, bool project) {
  const int cop = (co + 7) & ~7;
  const int nxt = cp <= 32 ? NX / 2 : NX;
  size_t smem = 1024;
  smem += project ? (((size_t)(cp / 2) * cop * 128 + 1023) & ~(size_t)1023) : 0;
  smem += (NG + nxt) * TILE16K + (project ? (cp / 2) * NR * 128 : 0) + 4096 + NX * REC_STAGE + NX * RR_STAGE;
  return smem;
}

// Fused-kernel mode: 0 = off (default: on the measured shapes the stand-alone aggregation + GEMM kernels are faster, see
// profiles/r02_fused_kernel.md), 1 = fused aggregation + stand-alone projection GEMM, 2 = aggregation and projection in the
// fused kernel where the projection weights fit in shared memory.  Initial value from SE3_FUSED, se3_conv_set_fused() after.
static int g_fused_mode = -1;
int fused_mode() {
  if (g_fused_mode < 0) g_fused_mode = getenv("SE3_FUSED") ? atoi(getenv("SE3_FUSED")) : 0;
  return g_fused_mode;
}
void fused_set_mode(int m) { g_fused_mode = m < 0 ? 0 : (m > 2 ? 2 : m); }

bool fused_supported(int c_gathered, int co, int f_row, int f_g, bool project) {
  const int mode = fused_mode();
  if (!mode || (project && mode < 2)) return false;
  if (fused_cp(c_gathered) == 0 || f_row < 1 || f_row > 2 || f_g < 1 || f_g > 4) return false;
  if (project && (co > 128 || fused_smem_bytes(fused_cp(c_gathered), co, true) > 226 * 1024)) return false;
  return true;
}

template <int CP, int FR, bool TR>
static int launch_fused_cfg(const FusedArgs& a, cudaStream_t st) {
  const size_t smem = fused_smem_bytes(CP, a.co, a.w3img != nullptr);
  int64_t blocks = std::min<int64_t>(num_sms(), std::max<int64_t>(1, (a.n_rows + 15) / 16));
#ifdef SE3_FUSED_TRACE
  static long long* trace_dev = nullptr;
  const size_t trace_n = (size_t)(NTHREADS / 32) * TRACE_PER_WARP * 3;
  if (!trace_dev) cudaMalloc(&trace_dev, trace_n * sizeof(long long));
  cudaMemcpyToSymbol(g_trace, &trace_dev, sizeof(trace_dev));
  cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), st);
#endif
  ProfScope prof(TR ? 1 : 0, st);
  if (a.act == 2) {
    auto kern = k_conv_fused<CP, FR, TR, 2>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(NTHREADS), smem, st, a));
  } else {
    auto kern = k_conv_fused<CP, FR, TR, -1>;
    SE3_SMEM_ONCE(kern, smem);
    SE3_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(NTHREADS), smem, st, a));
  }
  SE3_LAUNCH_CHECK();
#ifdef SE3_FUSED_TRACE
  {
    cudaStreamSynchronize(st);
    std::vector<long long> h(trace_n);
    cudaMemcpy(h.data(), trace_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    FILE* f = fopen(TR ? "/tmp/se3_fused_trace_tr.txt" : "/tmp/se3_fused_trace.txt", "w");
    if (f) {
      for (size_t i = 0; i < trace_n / 3; ++i)
        if (h[3 * i]) fprintf(f, "%lld %lld %lld\n", h[3 * i], h[3 * i + 1], h[3 * i + 2]);
      fclose(f);
    }
  }
#endif
  return SE3_OK;
}

int launch_conv_fused(const FusedArgs& a, int f_row, bool tr, cudaStream_t st) {
  if (a.n_rows == 0) return SE3_OK;
  if (a.n_rows >= ((int64_t)1 << 30) || a.n_edges >= ((int64_t)1 << 30)) {
    set_error("launch_conv_fused: problem too large for 32-bit offsets");
    return SE3_EINVAL;
  }
  const int cp = fused_cp(a.c);
#define SE3_FUSED_CASE(CP_)                                                                         \
  if (cp == CP_) {                                                                                  \
    if (f_row == 1) return tr ? launch_fused_cfg<CP_, 1, true>(a, st) : launch_fused_cfg<CP_, 1, false>(a, st); \
    if (f_row == 2) return tr ? launch_fused_cfg<CP_, 2, true>(a, st) : launch_fused_cfg<CP_, 2, false>(a, st); \
  }
  SE3_FUSED_CASE(16)
  SE3_FUSED_CASE(32)
  SE3_FUSED_CASE(64)
#undef SE3_FUSED_CASE
  set_error("launch_conv_fused: unsupported configuration (channels %d, row frames %d)", a.c, f_row);
  return SE3_EINVAL;
}

}  // namespace se3
