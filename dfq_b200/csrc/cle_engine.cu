// Cross-layer equalization engine: ONE persistent cooperative kernel runs every sweep of
// dfq.py:78-117 (cross_layer_equalization) on the device, including the exit rule.
//
// Reference semantics being reproduced (file:line under /root/reference):
//   dfq.py:48-55   per-channel range of row c of W1 and of input column c of W2
//   dfq.py:58-59   s = (1/(r1+eps)) * sqrt(r1*r2+eps), clamped with Python min/max semantics
//   dfq.py:62-73   W1[c] *= s, bn_weight/bn_bias/bias[c] *= s, W2[:, c] *= 1/s
//   dfq.py:84,105-115  convergence: sum over layers of mean|W - W_prev|, exit rule
//   relation.py:20-24  Relation.S accumulates the product of per-sweep s
//
// Data-parallel restructuring (bit-identical):
//   * the reference's per-channel Python loop is independent across channels, so all ranges of a
//     relation are formed first, then all scalings;
//   * relations are visited in forward chain order, so the column scaling of relation A on layer l
//     and the row scaling of relation B on the same layer are applied in ONE pass over l
//     (v -> fl(v*inv_A[col]) -> fl(.*s_B[row])), 8 bytes of HBM traffic per weight per sweep;
//   * column ranges needed by the *next* sweep are not re-read: rounding is monotone, so after a
//     pure column scaling  max(fl(v*a)) = fl(max(v)*a)  for a > 0 (likewise min), and the running
//     column extrema are updated analytically ("derived").  Only general middle layers (both
//     column- and row-scaled with cols > 1) are re-scanned.
//   * the step barrier between chain positions is a grid barrier of the persistent kernel.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dfq {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kScanRows = 32;      // rows per column-scan tile
constexpr int kScanCols = 2048;    // columns kept in shared memory by a scan tile

// Convergence state of one GROUP of chains.  The reference's exit rule (dfq.py:81-115) is evaluated per model: one
// group.  A batch of independent models (the synthetic stack: every block is its own model) is calibrated in one
// launch with one group per model, each stopping on its own.
struct GroupState {
  double acc[3];   // rotating per-sweep accumulators of sum_l mean|dW_l|
  double diff;     // `diff` of dfq.py:81
  int count;       // `count` of dfq.py:82
  int n_sweeps;
  int done;
  int converged;
};
struct CleCtl {
  int active[2];   // groups still iterating, double-buffered by sweep parity
  double diffs[64];  // diff_tmp per sweep of group 0
};

enum RowKind { RK_W4_1 = 0, RK_W4_4, RK_W4_8, RK_C4_8, RK_WS_1, RK_WS_8, RK_GENERIC };

__host__ __device__ inline int row_kind(int64_t w_off, int row_len) {
  const bool v4 = (row_len % 4 == 0) && (w_off % 4 == 0);
  if (v4) {
    if (row_len <= 128) return RK_W4_1;
    if (row_len <= 512) return RK_W4_4;
    if (row_len <= 1024) return RK_W4_8;
    if (row_len <= 8192) return RK_C4_8;
    return RK_GENERIC;
  }
  if (row_len <= 32) return RK_WS_1;
  if (row_len <= 256) return RK_WS_8;
  return RK_GENERIC;
}
__host__ __device__ inline int rows_per_tile(int kind) {
  return (kind == RK_C4_8 || kind == RK_GENERIC) ? 1 : kWarps;
}
__host__ __device__ inline int pass_tiles(const DfqLayer& l) {
  const int rpt = rows_per_tile(row_kind(l.w_off, l.cols * l.kk));
  return (l.rows + rpt - 1) / rpt;
}
__host__ __device__ inline int scan_tiles(int G, int go) {
  return G * ((go + kScanRows - 1) / kScanRows);
}

// Everything a row pass needs to know about its layer; uniform across the CTA.
struct RowCtx {
  float* w;
  int rows, cols, kk, row_len;
  const float* inv_in;  // reciprocal scales of rel_in (applied to columns), or nullptr
  int in_gi, in_go;
  int has_out;
  const float* cmin_rd;  // column extrema of the second layer of rel_out, buffer of this sweep
  const float* cmax_rd;
  float* cmin_wr;        // next sweep's buffer (derived update), or nullptr when it is re-scanned
  float* cmax_wr;
  float *s_step, *inv_out, *s_acc, *bias, *bnw, *bnb;
  float* own_cmin_wr;    // col_mode 1: this layer's own next-sweep extrema get the row factor too
  float* own_cmax_wr;
  double inv_n;
  int first_sweep;
};

// dfq.py:58-59 + :73.  Returns s; *inv is the factor applied to the columns of the second layer.
__device__ __forceinline__ float solve_scale(float r1, float r2, const DfqCleParams& P, float* inv) {
  const float a = __frcp_rn(__fadd_rn(r1, P.eps));
  const float b = __fsqrt_rn(__fadd_rn(__fmul_rn(r1, r2), P.eps));
  const float s = __fmul_rn(a, b);
  // Python: m = min(hi, s) -> s if s < hi else hi   (NaN -> hi);  max(lo, m) -> m if m > lo else lo
  if (!(s < P.s_hi)) {
    if (P.s_hi > P.s_lo) { *inv = P.inv_hi; return P.s_hi; }
    *inv = P.inv_lo; return P.s_lo;
  }
  if (s > P.s_lo) { *inv = __frcp_rn(s); return s; }
  *inv = P.inv_lo;
  return P.s_lo;
}

__device__ __forceinline__ float range_of(float mn, float mx, int signed_mode) {
  return signed_mode ? fmaxf(fabsf(mn), fabsf(mx)) : __fsub_rn(mx, mn);
}

__device__ __forceinline__ int col_of(int e, int kk) {
  return kk == 1 ? e : (kk == 9 ? e / 9 : e / kk);
}

// CTA-wide min/max with one barrier (double-buffered scratch, see parity argument in DESIGN.md).
__device__ __forceinline__ void cta_minmax(float& mn, float& mx, float* red, int& parity) {
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float* r = red + (parity & 1) * 2 * kWarps;
  parity++;
  if (l == 0) { r[w] = mn; r[kWarps + w] = mx; }
  __syncthreads();
  float a = r[l & (kWarps - 1)], b = r[kWarps + (l & (kWarps - 1))];
#pragma unroll
  for (int o = kWarps / 2; o > 0; o >>= 1) {
    a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  }
  mn = a; mx = b;
}

// All threads of the row's group call this with the reduced row extrema; the leader publishes.
__device__ __forceinline__ float solve_and_publish(const RowCtx& c, const DfqCleParams& P, int o,
                                                   float mn, float mx, bool leader) {
  const float r1 = range_of(mn, mx, P.signed_mode);
  const float cmn = __ldcg(c.cmin_rd + o), cmx = __ldcg(c.cmax_rd + o);
  const float r2 = range_of(cmn, cmx, P.signed_mode);
  float inv;
  float s;
  if (P.apply_only) {          // replay a given scale vector (multi-GPU replicas): s = S[o], columns get 1/S[o]
    s = __ldcg(c.s_acc + o);
    inv = __frcp_rn(s);
  } else {
    s = solve_scale(r1, r2, P, &inv);
  }
  if (leader) {
    c.s_step[o] = s;
    __stcg(c.inv_out + o, inv);
    if (!P.apply_only) c.s_acc[o] = c.first_sweep ? s : __fmul_rn(c.s_acc[o], s);
    c.bias[o] = __fmul_rn(c.bias[o], s);
    if (c.bnw) c.bnw[o] = __fmul_rn(c.bnw[o], s);
    if (c.bnb) c.bnb[o] = __fmul_rn(c.bnb[o], s);
    if (c.cmin_wr) {  // derived column extrema of the second layer after its column scaling
      __stcg(c.cmin_wr + o, __fmul_rn(cmn, inv));
      __stcg(c.cmax_wr + o, __fmul_rn(cmx, inv));
    }
    if (c.own_cmin_wr) {  // depthwise middle layer: its single-row column is this row
      __stcg(c.own_cmin_wr + o, __fmul_rn(__ldcg(c.own_cmin_wr + o), s));
      __stcg(c.own_cmax_wr + o, __fmul_rn(__ldcg(c.own_cmax_wr + o), s));
    }
  }
  return s;
}

// One row, held in registers between the range and the rescale: NV 128-bit (or 32-bit) loads per
// thread are all in flight before the first use.  TPR = threads per row (32: warp, 256: CTA).
template <int NV, int TPR, bool VEC>
__device__ __forceinline__ void cle_row(const RowCtx& c, const DfqCleParams& P, int o, int lane,
                                        float* red, int& parity, double& dacc) {
  float* rowp = c.w + (size_t)o * c.row_len;
  const int cbase = c.inv_in ? (o / c.in_go) * c.in_gi : 0;
  const bool leader = (lane == 0);
  float dsum = 0.f;
  float s = 1.f;
  if (VEC) {
    const int n4 = c.row_len >> 2;
    float4 u[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int i4 = lane + i * TPR;
      if (i4 < n4) u[i] = ldg_stream((const float4*)rowp + i4);
    }
    if (c.has_out) {
      float mn = DFQ_INF, mx = -DFQ_INF;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int i4 = lane + i * TPR;
        if (i4 < n4) {
          float4 t = u[i];
          if (c.inv_in) {
            const int e = i4 * 4;
            t.x = __fmul_rn(t.x, __ldcg(c.inv_in + cbase + col_of(e, c.kk)));
            t.y = __fmul_rn(t.y, __ldcg(c.inv_in + cbase + col_of(e + 1, c.kk)));
            t.z = __fmul_rn(t.z, __ldcg(c.inv_in + cbase + col_of(e + 2, c.kk)));
            t.w = __fmul_rn(t.w, __ldcg(c.inv_in + cbase + col_of(e + 3, c.kk)));
          }
          mn = fminf(mn, fminf(fminf(t.x, t.y), fminf(t.z, t.w)));
          mx = fmaxf(mx, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
        }
      }
      if (TPR == 32) { mn = warp_min(mn); mx = warp_max(mx); }
      else cta_minmax(mn, mx, red, parity);
      s = solve_and_publish(c, P, o, mn, mx, leader);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int i4 = lane + i * TPR;
      if (i4 < n4) {
        float4 t = u[i];
        if (c.inv_in) {
          const int e = i4 * 4;
          t.x = __fmul_rn(t.x, __ldcg(c.inv_in + cbase + col_of(e, c.kk)));
          t.y = __fmul_rn(t.y, __ldcg(c.inv_in + cbase + col_of(e + 1, c.kk)));
          t.z = __fmul_rn(t.z, __ldcg(c.inv_in + cbase + col_of(e + 2, c.kk)));
          t.w = __fmul_rn(t.w, __ldcg(c.inv_in + cbase + col_of(e + 3, c.kk)));
        }
        if (c.has_out) {
          t.x = __fmul_rn(t.x, s); t.y = __fmul_rn(t.y, s);
          t.z = __fmul_rn(t.z, s); t.w = __fmul_rn(t.w, s);
        }
        stg_stream((float4*)rowp + i4, t);
        dsum += fabsf(__fsub_rn(t.x, u[i].x)) + fabsf(__fsub_rn(t.y, u[i].y)) +
                fabsf(__fsub_rn(t.z, u[i].z)) + fabsf(__fsub_rn(t.w, u[i].w));
      }
    }
  } else {
    const int n = c.row_len;
    float u[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = lane + i * TPR;
      if (e < n) u[i] = ldg_stream1(rowp + e);
    }
    if (c.has_out) {
      float mn = DFQ_INF, mx = -DFQ_INF;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int e = lane + i * TPR;
        if (e < n) {
          float t = u[i];
          if (c.inv_in) t = __fmul_rn(t, __ldcg(c.inv_in + cbase + col_of(e, c.kk)));
          mn = fminf(mn, t); mx = fmaxf(mx, t);
        }
      }
      if (TPR == 32) { mn = warp_min(mn); mx = warp_max(mx); }
      else cta_minmax(mn, mx, red, parity);
      s = solve_and_publish(c, P, o, mn, mx, leader);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = lane + i * TPR;
      if (e < n) {
        float t = u[i];
        if (c.inv_in) t = __fmul_rn(t, __ldcg(c.inv_in + cbase + col_of(e, c.kk)));
        if (c.has_out) t = __fmul_rn(t, s);
        stg_stream1(rowp + e, t);
        dsum += fabsf(__fsub_rn(t, u[i]));
      }
    }
  }
  dacc += (double)dsum * c.inv_n;
}

// Any row length / alignment: CTA per row, the row is read twice (second read is an L2 hit).
__device__ __forceinline__ void cle_row_generic(const RowCtx& c, const DfqCleParams& P, int o,
                                                float* red, int& parity, double& dacc) {
  float* rowp = c.w + (size_t)o * c.row_len;
  const int cbase = c.inv_in ? (o / c.in_go) * c.in_gi : 0;
  float s = 1.f;
  if (c.has_out) {
    float mn = DFQ_INF, mx = -DFQ_INF;
    for (int e = threadIdx.x; e < c.row_len; e += kThreads) {
      float t = ldg_stream1(rowp + e);
      if (c.inv_in) t = __fmul_rn(t, __ldcg(c.inv_in + cbase + col_of(e, c.kk)));
      mn = fminf(mn, t); mx = fmaxf(mx, t);
    }
    cta_minmax(mn, mx, red, parity);
    s = solve_and_publish(c, P, o, mn, mx, threadIdx.x == 0);
  }
  float dsum = 0.f;
  for (int e = threadIdx.x; e < c.row_len; e += kThreads) {
    const float u = ldg_stream1(rowp + e);
    float t = u;
    if (c.inv_in) t = __fmul_rn(t, __ldcg(c.inv_in + cbase + col_of(e, c.kk)));
    if (c.has_out) t = __fmul_rn(t, s);
    stg_stream1(rowp + e, t);
    dsum += fabsf(__fsub_rn(t, u));
  }
  dacc += (double)dsum * c.inv_n;
}

__device__ __forceinline__ void make_ctx(RowCtx& c, float* arena, const DfqLayer* L, const DfqRelation* R,
                                         int li, int sweep) {
  const DfqLayer l = L[li];
  c.w = arena + l.w_off;
  c.rows = l.rows; c.cols = l.cols; c.kk = l.kk; c.row_len = l.cols * l.kk;
  c.inv_n = 1.0 / ((double)l.rows * (double)c.row_len);
  c.first_sweep = (sweep == 0);
  const int rd = sweep & 1, wr = rd ^ 1;
  c.inv_in = nullptr; c.in_gi = 1; c.in_go = 1;
  c.own_cmin_wr = c.own_cmax_wr = nullptr;
  if (l.rel_in >= 0) {
    const DfqRelation r = R[l.rel_in];
    c.inv_in = arena + r.inv_off;
    c.in_gi = r.gi; c.in_go = r.go;
    if (l.col_mode == 1 && l.rel_out >= 0) {
      c.own_cmin_wr = arena + l.cmin_off + (size_t)wr * r.channels;
      c.own_cmax_wr = arena + l.cmax_off + (size_t)wr * r.channels;
    }
  }
  c.has_out = (l.rel_out >= 0);
  c.cmin_rd = c.cmax_rd = nullptr; c.cmin_wr = c.cmax_wr = nullptr;
  c.s_step = c.inv_out = c.s_acc = c.bnw = c.bnb = nullptr;
  c.bias = arena + l.bias_off;
  if (c.has_out) {
    const DfqRelation r = R[l.rel_out];
    const DfqLayer l2 = L[r.second];
    c.cmin_rd = arena + l2.cmin_off + (size_t)rd * r.channels;
    c.cmax_rd = arena + l2.cmax_off + (size_t)rd * r.channels;
    if (l2.col_mode != 2) {
      c.cmin_wr = arena + l2.cmin_off + (size_t)wr * r.channels;
      c.cmax_wr = arena + l2.cmax_off + (size_t)wr * r.channels;
    }
    c.s_step = arena + r.s_step_off;
    c.inv_out = arena + r.inv_off;
    c.s_acc = arena + r.s_acc_off;
    c.bnw = r.bn_w_off >= 0 ? arena + r.bn_w_off : nullptr;
    c.bnb = r.bn_b_off >= 0 ? arena + r.bn_b_off : nullptr;
  }
}

// Column extrema of rows [r0, r1) of group g of layer l, folded into dst (global float atomics).
__device__ void scan_cols_tile(const float* w, int J, int kk, int g, int gi, int r0, int r1,
                               float* dmin, float* dmax, float* smin, float* smax) {
  const int row_len = J * kk;
  const bool use_smem = (J <= kScanCols);
  if (use_smem) {
    for (int j = threadIdx.x; j < J; j += kThreads) { smin[j] = DFQ_INF; smax[j] = -DFQ_INF; }
    __syncthreads();
  }
  for (int p = threadIdx.x; p < row_len; p += kThreads) {
    float mn = DFQ_INF, mx = -DFQ_INF;
    const float* q = w + (size_t)r0 * row_len + p;
#pragma unroll 8
    for (int r = r0; r < r1; ++r, q += row_len) {
      const float v = ldg_stream1(q);
      mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    const int j = col_of(p, kk);
    if (use_smem) { atomic_min_f(smin + j, mn); atomic_max_f(smax + j, mx); }
    else { atomic_min_f(dmin + g * gi + j, mn); atomic_max_f(dmax + g * gi + j, mx); }
  }
  if (use_smem) {
    __syncthreads();
    for (int j = threadIdx.x; j < J; j += kThreads) {
      atomic_min_f(dmin + g * gi + j, smin[j]);
      atomic_max_f(dmax + g * gi + j, smax[j]);
    }
    __syncthreads();
  }
}

// scan tiles [t0, t1) of layer li (tile = 32 rows of one group)
__device__ __forceinline__ void scan_layer(float* arena, const DfqLayer* L, const DfqRelation* R, int li,
                                           int buf, long long t0, long long t1, float* smin, float* smax) {
  const DfqLayer l = L[li];
  const DfqRelation r = R[l.rel_in];
  const int nb = (r.go + kScanRows - 1) / kScanRows;
  for (long long t = t0; t < t1; ++t) {
    const int g = (int)(t / nb), b = (int)(t - (long long)g * nb);
    const int r0 = g * r.go + b * kScanRows;
    const int r1 = min(r0 + kScanRows, (g + 1) * r.go);
    scan_cols_tile(arena + l.w_off, l.cols, l.kk, g, r.gi, r0, r1,
                   arena + l.cmin_off + (size_t)buf * r.channels,
                   arena + l.cmax_off + (size_t)buf * r.channels, smin, smax);
  }
}

__device__ __forceinline__ void reset_cols(float* arena, const DfqLayer& l, const DfqRelation& r, int buf) {
  float* a = arena + l.cmin_off + (size_t)buf * r.channels;
  float* b = arena + l.cmax_off + (size_t)buf * r.channels;
  for (int j = threadIdx.x; j < r.channels; j += kThreads) { __stcg(a + j, DFQ_INF); __stcg(b + j, -DFQ_INF); }
}

__global__ void __launch_bounds__(kThreads, 2)
k_cle_engine(float* arena, const DfqLayer* __restrict__ L, int nL, const DfqRelation* __restrict__ R, int nR,
             const int* __restrict__ step_ptr, const int* __restrict__ step_layers, int n_steps,
             const int* __restrict__ step_rescan, const long long* __restrict__ pass_ptr,
             const long long* __restrict__ scan_ptr, const int* __restrict__ scan_layers, int n_scan,
             DfqCleParams P, CleCtl* ctl, GroupState* G, int nG) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float red[2 * 2 * kWarps];
  __shared__ float smin[kScanCols];
  __shared__ float smax[kScanCols];
  __shared__ double dred[kWarps];
  __shared__ RowCtx sctx;
  int parity = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- phase 0: column extrema of every `second` layer (buffer 0) -----------------------------
  for (int g = blockIdx.x * kThreads + threadIdx.x; g < nG; g += gridDim.x * kThreads) G[g].diff = 10.0;   // dfq.py:81
  for (int li = blockIdx.x; li < nL; li += gridDim.x)
    if (L[li].rel_in >= 0) reset_cols(arena, L[li], R[L[li].rel_in], 0);
  grid.sync();
  {  // scan_ptr[0 .. n_scan]: tile prefix over scan_layers (all `second` layers)
    const TileSpan sp = tile_span(scan_ptr, 0, n_scan);
    for (int q = sp.q; q < n_scan && scan_ptr[q] < sp.hi; ++q) {
      const long long base = scan_ptr[q];
      scan_layer(arena, L, R, scan_layers[q], 0, max(sp.lo, base) - base, min(sp.hi, scan_ptr[q + 1]) - base, smin, smax);
    }
  }
  grid.sync();

  for (int sweep = 0;; ++sweep) {
    const int slot = sweep % 3;
    for (int p = 0; p < n_steps; ++p) {
      double dacc = 0.0;
      int cur_g = -1;
      // sum the CTA's partial of group cur_g into that group's accumulator (one atomic)
      auto flush = [&]() {
        if (cur_g < 0) return;
        dacc = warp_sum(dacc);
        __syncthreads();
        if (lane == 0) dred[warp] = dacc;
        __syncthreads();
        if (threadIdx.x == 0) {
          double t = 0.0;
#pragma unroll
          for (int i = 0; i < kWarps; ++i) t += dred[i];
          if (t != 0.0) atomicAdd(&G[cur_g].acc[slot], t);
        }
        dacc = 0.0;
      };
      const TileSpan sp = tile_span(pass_ptr, step_ptr[p], step_ptr[p + 1]);
      for (int q = sp.q; q < step_ptr[p + 1] && pass_ptr[q] < sp.hi; ++q) {
        const int li = step_layers[q];
        const long long base = pass_ptr[q];
        const int t0 = (int)(max(sp.lo, base) - base), t1 = (int)(min(sp.hi, pass_ptr[q + 1]) - base);
        if (t1 <= t0) continue;
        const int g = L[li].group;
        if (*((volatile int*)&G[g].done)) continue;          // this model has converged: its weights are final
        if (g != cur_g) { flush(); cur_g = g; }
        __syncthreads();                       // previous layer's tiles are done with sctx
        if (threadIdx.x == 0) make_ctx(sctx, arena, L, R, li, sweep);
        __syncthreads();
        const RowCtx& c = sctx;
        const int kind = row_kind(L[li].w_off, c.row_len);
        const int rpt = rows_per_tile(kind);
        if (L[li].col_mode == 2 && L[li].rel_in >= 0 && t0 == 0)
          reset_cols(arena, L[li], R[L[li].rel_in], (sweep & 1) ^ 1);
        for (int t = t0; t < t1; ++t) {
          if (rpt == 1) {
            if (kind == RK_C4_8) cle_row<8, kThreads, true>(c, P, t, threadIdx.x, red, parity, dacc);
            else cle_row_generic(c, P, t, red, parity, dacc);
          } else {
            const int o = t * kWarps + warp;
            if (o < c.rows) {
              switch (kind) {
                case RK_W4_1: cle_row<1, 32, true>(c, P, o, lane, red, parity, dacc); break;
                case RK_W4_4: cle_row<4, 32, true>(c, P, o, lane, red, parity, dacc); break;
                case RK_W4_8: cle_row<8, 32, true>(c, P, o, lane, red, parity, dacc); break;
                case RK_WS_1: cle_row<1, 32, false>(c, P, o, lane, red, parity, dacc); break;
                default: cle_row<8, 32, false>(c, P, o, lane, red, parity, dacc); break;
              }
            }
          }
        }
      }
      flush();
      grid.sync();
      if (step_rescan[p]) {   // general middle layers of this step: round-robin over their scan tiles
        long long sbase = 0;
        for (int q = step_ptr[p]; q < step_ptr[p + 1]; ++q) {
          const int li = step_layers[q];
          if (L[li].col_mode == 2 && L[li].rel_in >= 0 && !*((volatile int*)&G[L[li].group].done)) {
            const DfqRelation r = R[L[li].rel_in];
            const long long nt = scan_tiles(r.groups, r.go);
            long long first = ((long long)blockIdx.x - sbase) % (long long)gridDim.x;
            if (first < 0) first += gridDim.x;
            for (long long t = first; t < nt; t += gridDim.x)
              scan_layer(arena, L, R, li, (sweep & 1) ^ 1, t, t + 1, smin, smax);
            sbase += nt;
          }
        }
        grid.sync();
      }
    }
    // ---- exit rule of dfq.py:105-115, one thread per group ------------------------------------------------
    const int n = sweep + 1;
    for (int g = blockIdx.x * kThreads + threadIdx.x; g < nG; g += gridDim.x * kThreads) {
      GroupState& st = G[g];
      if (st.done) continue;
      const double diff_tmp = st.acc[slot];
      st.acc[(slot + 2) % 3] = 0.0;   // last read before this sweep's final barrier, next used in sweep+2
      if (fabs(st.diff - diff_tmp) > 1e-9) { st.count = 0; st.diff = diff_tmp; }
      else st.count++;
      if (g == 0 && sweep < 64) ctl->diffs[sweep] = diff_tmp;
      const bool cont = (st.diff > P.converge_thres) && (st.count < P.converge_count);
      // safety net: the reference's loop has no bound; 4096 sweeps is ~80x what any of its models needs
      const int cap = P.max_sweeps > 0 ? P.max_sweeps : 4096;
      if (!cont || n >= cap) { st.n_sweeps = n; st.converged = !cont; st.done = 1; }
      else atomicAdd(&ctl->active[n & 1], 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->active[sweep & 1] = 0;   // read at the end of the previous sweep
    __threadfence();
    grid.sync();
    if (*((volatile int*)&ctl->active[n & 1]) == 0) break;
  }
}

}  // namespace dfq

using namespace dfq;

extern "C" int dfq_cle_run(float* arena, int64_t arena_floats, const DfqLayer* layers, int32_t n_layers,
                           const DfqRelation* rels, int32_t n_rels, const int32_t* step_ptr,
                           const int32_t* step_layers, int32_t n_steps, const DfqCleParams* params,
                           DfqCleResult* result, int32_t n_groups, int32_t* group_sweeps, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DFQ_REQUIRE(arena && layers && rels && step_ptr && step_layers && params && result, "null argument");
  DFQ_REQUIRE(n_layers > 0 && n_rels > 0 && n_steps > 0 && n_groups > 0, "empty problem");
  memset(result, 0, sizeof(*result));
  // Python `while diff > thres and count < converge_count` with diff = 10, count = 0 (dfq.py:81-83)
  if (!(10.0 > params->converge_thres) || !(0 < params->converge_count)) { result->converged = 1; result->last_diff = 10.0; return 0; }

  // ---- validate descriptors, find the widest phase -----------------------------------------------
  int64_t max_tiles = 1;
  std::vector<int32_t> rescan(n_steps, 0);
  for (int i = 0; i < n_rels; ++i) {
    const DfqRelation& r = rels[i];
    DFQ_REQUIRE(r.first >= 0 && r.first < n_layers && r.second >= 0 && r.second < n_layers, "relation layer index");
    const DfqLayer& a = layers[r.first];
    const DfqLayer& b = layers[r.second];
    DFQ_REQUIRE(r.channels == a.rows, "relation.channels != rows(first)");
    DFQ_REQUIRE(r.groups >= 1 && r.groups * r.gi == r.channels && r.groups * r.go == b.rows, "relation grouping");
    DFQ_REQUIRE(r.gi == b.cols, "first.rows / groups must equal second.cols (dfq.py:29-35)");
    DFQ_REQUIRE(a.rel_out == i && b.rel_in == i, "layer/relation cross links");
    DFQ_REQUIRE(a.group == b.group, "both layers of a relation must belong to the same convergence group");
    DFQ_REQUIRE(r.s_acc_off >= 0 && r.s_step_off >= 0 && r.inv_off >= 0, "relation scratch offsets");
    DFQ_REQUIRE(b.cmin_off >= 0 && b.cmax_off >= 0, "second layer needs column range scratch");
    DFQ_REQUIRE(b.col_mode != 1 || (b.cols == 1 && r.go == 1), "col_mode 1 requires cols==1 and one row per group");
    DFQ_REQUIRE(b.col_mode != 0 || b.rel_out < 0, "col_mode 0 is for chain ends");
  }
  for (int i = 0; i < n_layers; ++i) {
    const DfqLayer& l = layers[i];
    DFQ_REQUIRE(l.rows > 0 && l.cols > 0 && l.kk > 0, "layer shape");
    DFQ_REQUIRE(l.group >= 0 && l.group < n_groups, "layer group index");
    DFQ_REQUIRE(l.w_off >= 0 && l.w_off + (int64_t)l.rows * l.cols * l.kk <= arena_floats, "weight outside arena");
    DFQ_REQUIRE(l.bias_off >= 0 && l.bias_off + l.rows <= arena_floats, "bias outside arena");
    if (l.rel_in >= 0 && l.rel_out >= 0) DFQ_REQUIRE(l.rel_in < l.rel_out, "relations must be in forward chain order");
  }
  int64_t scan_total = 0;
  for (int i = 0; i < n_layers; ++i)
    if (layers[i].rel_in >= 0) scan_total += scan_tiles(rels[layers[i].rel_in].groups, rels[layers[i].rel_in].go);
  max_tiles = std::max(max_tiles, scan_total);
  for (int p = 0; p < n_steps; ++p) {
    int64_t t = 0;
    for (int q = step_ptr[p]; q < step_ptr[p + 1]; ++q) {
      const int li = step_layers[q];
      DFQ_REQUIRE(li >= 0 && li < n_layers, "step layer index");
      const DfqLayer& l = layers[li];
      DFQ_REQUIRE(l.rel_in >= 0 || l.rel_out >= 0, "step layer without relation");
      t += pass_tiles(l);
      if (l.rel_in >= 0 && l.col_mode == 2) rescan[p] = 1;
    }
    max_tiles = std::max(max_tiles, t);
  }

  int dev = 0, sms = 0, per_sm = 0, coop = 0;
  DFQ_CUDA(cudaGetDevice(&dev));
  DFQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DFQ_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) { set_error("device does not support cooperative launch"); return DFQ_E_NOT_COOPERATIVE; }
  DFQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cle_engine, kThreads, 0));
  if (per_sm < 1) { set_error("persistent kernel does not fit on an SM"); return DFQ_E_NOT_COOPERATIVE; }
  const int grid = (int)std::min<int64_t>((int64_t)sms * per_sm, max_tiles);

  const int n_entries = step_ptr[n_steps];
  std::vector<long long> pass_ptr(n_entries + 1, 0);
  for (int q = 0; q < n_entries; ++q) pass_ptr[q + 1] = pass_ptr[q] + pass_tiles(layers[step_layers[q]]);
  std::vector<int32_t> scan_layers;
  std::vector<long long> scan_ptr(1, 0);
  for (int i = 0; i < n_layers; ++i)
    if (layers[i].rel_in >= 0) {
      scan_layers.push_back(i);
      scan_ptr.push_back(scan_ptr.back() + scan_tiles(rels[layers[i].rel_in].groups, rels[layers[i].rel_in].go));
    }
  int n_scan = (int)scan_layers.size();
  long long *dPP = nullptr, *dSCP = nullptr; int32_t* dSCL = nullptr;
  DfqLayer* dL = nullptr; DfqRelation* dR = nullptr; int32_t *dSP = nullptr, *dSL = nullptr, *dRS = nullptr;
  CleCtl* dctl = nullptr;
  GroupState* dG = nullptr;
  int rc;
  if ((rc = upload(layers, n_layers, &dL, st))) return rc;
  if ((rc = upload(rels, n_rels, &dR, st))) return rc;
  if ((rc = upload(step_ptr, n_steps + 1, &dSP, st))) return rc;
  if ((rc = upload(step_layers, step_ptr[n_steps], &dSL, st))) return rc;
  if ((rc = upload(rescan.data(), n_steps, &dRS, st))) return rc;
  if ((rc = upload(pass_ptr.data(), n_entries + 1, &dPP, st))) return rc;
  if ((rc = upload(scan_ptr.data(), n_scan + 1, &dSCP, st))) return rc;
  if ((rc = upload(scan_layers.data(), n_scan, &dSCL, st))) return rc;
  DFQ_CUDA(cudaMallocAsync((void**)&dctl, sizeof(CleCtl), st));
  DFQ_CUDA(cudaMemsetAsync(dctl, 0, sizeof(CleCtl), st));
  DFQ_CUDA(cudaMallocAsync((void**)&dG, sizeof(GroupState) * n_groups, st));
  DFQ_CUDA(cudaMemsetAsync(dG, 0, sizeof(GroupState) * n_groups, st));

  DfqCleParams P = *params;
  void* args[] = {&arena, &dL, (void*)&n_layers, &dR, (void*)&n_rels, &dSP, &dSL, (void*)&n_steps, &dRS,
                  &dPP, &dSCP, &dSCL, (void*)&n_scan, &P, &dctl, &dG, (void*)&n_groups};
  DFQ_CUDA(cudaLaunchCooperativeKernel((void*)k_cle_engine, dim3(grid), dim3(kThreads), args, 0, st));
  CleCtl h;
  std::vector<GroupState> hg(n_groups);
  DFQ_CUDA(cudaMemcpyAsync(&h, dctl, sizeof(CleCtl), cudaMemcpyDeviceToHost, st));
  DFQ_CUDA(cudaMemcpyAsync(hg.data(), dG, sizeof(GroupState) * n_groups, cudaMemcpyDeviceToHost, st));
  free_async(dL, st); free_async(dR, st); free_async(dSP, st); free_async(dSL, st); free_async(dRS, st);
  free_async(dPP, st); free_async(dSCP, st); free_async(dSCL, st);
  free_async(dctl, st); free_async(dG, st);
  DFQ_CUDA(cudaStreamSynchronize(st));
  result->n_sweeps = 0;
  result->converged = 1;
  for (int g = 0; g < n_groups; ++g) {
    result->n_sweeps = std::max(result->n_sweeps, hg[g].n_sweeps);
    result->converged &= hg[g].converged;
    if (group_sweeps) group_sweeps[g] = hg[g].n_sweeps;
  }
  result->last_diff = hg[0].diff;
  memcpy(result->diffs, h.diffs, sizeof(h.diffs));
  return 0;
}
