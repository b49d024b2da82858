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

struct CleCtl {
  double acc[3];   // rotating per-sweep accumulators of sum_l mean|dW_l|
  double diff;     // `diff` of dfq.py:81
  int count;       // `count` of dfq.py:82
  int n_sweeps;
  int done;
  int converged;
  double diffs[64];
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
  const float s = solve_scale(r1, r2, P, &inv);
  if (leader) {
    c.s_step[o] = s;
    __stcg(c.inv_out + o, inv);
    c.s_acc[o] = c.first_sweep ? s : __fmul_rn(c.s_acc[o], s);
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

__device__ __forceinline__ void scan_layer(float* arena, const DfqLayer* L, const DfqRelation* R, int li,
                                           int buf, int& tile_base, float* smin, float* smax) {
  const DfqLayer l = L[li];
  const DfqRelation r = R[l.rel_in];
  const int nb = (r.go + kScanRows - 1) / kScanRows;
  const int nt = r.groups * nb;
  int first = (int)(((long long)blockIdx.x - tile_base) % (long long)gridDim.x);
  if (first < 0) first += gridDim.x;
  for (int t = first; t < nt; t += gridDim.x) {
    const int g = t / nb, b = t - g * nb;
    const int r0 = g * r.go + b * kScanRows;
    const int r1 = min(r0 + kScanRows, (g + 1) * r.go);
    scan_cols_tile(arena + l.w_off, l.cols, l.kk, g, r.gi, r0, r1,
                   arena + l.cmin_off + (size_t)buf * r.channels,
                   arena + l.cmax_off + (size_t)buf * r.channels, smin, smax);
  }
  tile_base += nt;
}

__device__ __forceinline__ void reset_cols(float* arena, const DfqLayer& l, const DfqRelation& r, int buf) {
  float* a = arena + l.cmin_off + (size_t)buf * r.channels;
  float* b = arena + l.cmax_off + (size_t)buf * r.channels;
  for (int j = threadIdx.x; j < r.channels; j += kThreads) { __stcg(a + j, DFQ_INF); __stcg(b + j, -DFQ_INF); }
}

__global__ void __launch_bounds__(kThreads, 2)
k_cle_engine(float* arena, const DfqLayer* __restrict__ L, int nL, const DfqRelation* __restrict__ R, int nR,
             const int* __restrict__ step_ptr, const int* __restrict__ step_layers, int n_steps,
             const int* __restrict__ step_rescan, DfqCleParams P, CleCtl* ctl) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float red[2 * 2 * kWarps];
  __shared__ float smin[kScanCols];
  __shared__ float smax[kScanCols];
  __shared__ double dred[kWarps];
  __shared__ RowCtx sctx;
  int parity = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- phase 0: column extrema of every `second` layer (buffer 0) -----------------------------
  for (int li = blockIdx.x; li < nL; li += gridDim.x)
    if (L[li].rel_in >= 0) reset_cols(arena, L[li], R[L[li].rel_in], 0);
  grid.sync();
  {
    int base = 0;
    for (int li = 0; li < nL; ++li)
      if (L[li].rel_in >= 0) scan_layer(arena, L, R, li, 0, base, smin, smax);
  }
  grid.sync();

  // exit-rule state, replicated in every CTA's thread 0 (all see the same accumulators)
  double diff = 10.0;
  int count = 0;

  for (int sweep = 0;; ++sweep) {
    const int slot = sweep % 3;
    for (int p = 0; p < n_steps; ++p) {
      double dacc = 0.0;
      int base = 0;
      for (int q = step_ptr[p]; q < step_ptr[p + 1]; ++q) {
        const int li = step_layers[q];
        __syncthreads();                       // previous layer's tiles are done with sctx
        if (threadIdx.x == 0) make_ctx(sctx, arena, L, R, li, sweep);
        __syncthreads();
        const RowCtx& c = sctx;
        const int kind = row_kind(L[li].w_off, c.row_len);
        const int rpt = rows_per_tile(kind);
        const int nt = (c.rows + rpt - 1) / rpt;
        int first = (int)(((long long)blockIdx.x - base) % (long long)gridDim.x);
        if (first < 0) first += gridDim.x;
        if (L[li].col_mode == 2 && L[li].rel_in >= 0 && first == 0)
          reset_cols(arena, L[li], R[L[li].rel_in], (sweep & 1) ^ 1);
        for (int t = first; t < nt; t += gridDim.x) {
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
        base += nt;
      }
      // one atomic per CTA per step
      dacc = warp_sum(dacc);
      if (lane == 0) dred[warp] = dacc;
      __syncthreads();
      if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) t += dred[i];
        if (t != 0.0) atomicAdd(&ctl->acc[slot], t);
      }
      grid.sync();
      if (step_rescan[p]) {
        int sbase = 0;
        for (int q = step_ptr[p]; q < step_ptr[p + 1]; ++q) {
          const int li = step_layers[q];
          if (L[li].col_mode == 2 && L[li].rel_in >= 0)
            scan_layer(arena, L, R, li, (sweep & 1) ^ 1, sbase, smin, smax);
        }
        grid.sync();
      }
    }
    // ---- exit rule of dfq.py:105-115, evaluated identically by every CTA -----------------------
    const double diff_tmp = *((volatile double*)&ctl->acc[slot]);
    if (fabs(diff - diff_tmp) > 1e-9) { count = 0; diff = diff_tmp; }
    else count++;
    const int n = sweep + 1;
    const bool cont = (diff > P.converge_thres) && (count < P.converge_count);
    const bool stop = !cont || (P.max_sweeps > 0 && n >= P.max_sweeps);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      ctl->acc[(slot + 2) % 3] = 0.0;  // last read before this sweep's final barrier, next used in sweep+2
      if (sweep < 64) ctl->diffs[sweep] = diff_tmp;
      if (stop) { ctl->n_sweeps = n; ctl->diff = diff; ctl->count = count; ctl->converged = !cont; ctl->done = 1; }
      __threadfence();
    }
    if (stop) break;
  }
}

}  // namespace dfq

using namespace dfq;

extern "C" int dfq_cle_run(float* arena, int64_t arena_floats, const DfqLayer* layers, int32_t n_layers,
                           const DfqRelation* rels, int32_t n_rels, const int32_t* step_ptr,
                           const int32_t* step_layers, int32_t n_steps, const DfqCleParams* params,
                           DfqCleResult* result, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DFQ_REQUIRE(arena && layers && rels && step_ptr && step_layers && params && result, "null argument");
  DFQ_REQUIRE(n_layers > 0 && n_rels > 0 && n_steps > 0, "empty problem");
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
    DFQ_REQUIRE(r.s_acc_off >= 0 && r.s_step_off >= 0 && r.inv_off >= 0, "relation scratch offsets");
    DFQ_REQUIRE(b.cmin_off >= 0 && b.cmax_off >= 0, "second layer needs column range scratch");
    DFQ_REQUIRE(b.col_mode != 1 || (b.cols == 1 && r.go == 1), "col_mode 1 requires cols==1 and one row per group");
    DFQ_REQUIRE(b.col_mode != 0 || b.rel_out < 0, "col_mode 0 is for chain ends");
  }
  for (int i = 0; i < n_layers; ++i) {
    const DfqLayer& l = layers[i];
    DFQ_REQUIRE(l.rows > 0 && l.cols > 0 && l.kk > 0, "layer shape");
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

  DfqLayer* dL = nullptr; DfqRelation* dR = nullptr; int32_t *dSP = nullptr, *dSL = nullptr, *dRS = nullptr;
  CleCtl* dctl = nullptr;
  int rc;
  if ((rc = upload(layers, n_layers, &dL, st))) return rc;
  if ((rc = upload(rels, n_rels, &dR, st))) return rc;
  if ((rc = upload(step_ptr, n_steps + 1, &dSP, st))) return rc;
  if ((rc = upload(step_layers, step_ptr[n_steps], &dSL, st))) return rc;
  if ((rc = upload(rescan.data(), n_steps, &dRS, st))) return rc;
  DFQ_CUDA(cudaMallocAsync((void**)&dctl, sizeof(CleCtl), st));
  DFQ_CUDA(cudaMemsetAsync(dctl, 0, sizeof(CleCtl), st));

  DfqCleParams P = *params;
  void* args[] = {&arena, &dL, (void*)&n_layers, &dR, (void*)&n_rels, &dSP, &dSL, (void*)&n_steps, &dRS, &P, &dctl};
  DFQ_CUDA(cudaLaunchCooperativeKernel((void*)k_cle_engine, dim3(grid), dim3(kThreads), args, 0, st));
  CleCtl h;
  DFQ_CUDA(cudaMemcpyAsync(&h, dctl, sizeof(CleCtl), cudaMemcpyDeviceToHost, st));
  free_async(dL, st); free_async(dR, st); free_async(dSP, st); free_async(dSL, st); free_async(dRS, st);
  free_async(dctl, st);
  DFQ_CUDA(cudaStreamSynchronize(st));
  result->n_sweeps = h.n_sweeps;
  result->converged = h.converged;
  result->last_diff = h.diff;
  memcpy(result->diffs, h.diffs, sizeof(h.diffs));
  return 0;
}
