// Batched arena passes other than equalization: BN fold, bias correction, weight/bias fake-quant.
//
//   dfq_bn_fold           utils/layer_transform.py:231-276  (merge_batchnorm)
//   dfq_bias_correct      dfq.py:173-293                    (bias_correction)
//   dfq_quantize_tensors  utils/layer_transform.py:279-296  (quantize_targ_layer)
#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dfq {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = kThreads * 16;   // floats per flat tile

// round-robin tile ownership shared by all batched kernels: tiles are numbered consecutively over the
// task list; CTA b owns global tile indices congruent to b modulo gridDim.x
__device__ __forceinline__ int first_tile(long long base) {
  int f = (int)(((long long)blockIdx.x - base) % (long long)gridDim.x);
  return f < 0 ? f + gridDim.x : f;
}

// ------------------------------------------------------------------------------------------------
// BN fold
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_bn_fold(float* arena, const DfqLayer* __restrict__ L, const DfqFold* __restrict__ F, int nF,
          const long long* __restrict__ tptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TileSpan sp = tile_span(tptr, 0, nF);
  for (int fi = sp.q; fi < nF && tptr[fi] < sp.hi; ++fi) {
    const long long base = tptr[fi];
    const DfqFold f = F[fi];
    const DfqLayer l = L[f.layer];
    const int row_len = l.cols * l.kk;
    const bool vec = (row_len % 4 == 0) && (l.w_off % 4 == 0);
    const bool cta_row = row_len > 2048;
    const int rpt = cta_row ? 1 : kWarps;
    const int t0 = (int)(max(sp.lo, base) - base), t1 = (int)(min(sp.hi, tptr[fi + 1]) - base);
    for (int t = t0; t < t1; ++t) {
      const int o = cta_row ? t : t * kWarps + warp;
      if (o >= l.rows) continue;
      const int tid = cta_row ? (int)threadIdx.x : lane;
      const int tpr = cta_row ? kThreads : 32;
      // layer_transform.py:251: gamma / sqrt(var + eps) formed first, then multiplied in
      const float gamma = arena[f.gamma_off + o], var = arena[f.var_off + o];
      const float den = __fsqrt_rn(__fadd_rn(var, f.bn_eps));
      const float fac = __fdiv_rn(gamma, den);
      float* rowp = arena + l.w_off + (size_t)o * row_len;
      if (vec) {
        float4* r4 = (float4*)rowp;
        for (int i = tid; i < (row_len >> 2); i += tpr) {
          float4 v = ldg_stream(r4 + i);
          v.x = __fmul_rn(v.x, fac); v.y = __fmul_rn(v.y, fac);
          v.z = __fmul_rn(v.z, fac); v.w = __fmul_rn(v.w, fac);
          stg_stream(r4 + i, v);
        }
      } else {
        for (int i = tid; i < row_len; i += tpr) stg_stream1(rowp + i, __fmul_rn(ldg_stream1(rowp + i), fac));
      }
      if (tid == 0) {
        // layer_transform.py:260-261: b*f + (beta - (gamma*mean)/sqrt(var+eps))
        const float beta = arena[f.beta_off + o], mean = arena[f.mean_off + o];
        const float b = arena[l.bias_off + o];
        const float shift = __fsub_rn(beta, __fdiv_rn(__fmul_rn(gamma, mean), den));
        arena[l.bias_off + o] = __fadd_rn(__fmul_rn(b, fac), shift);
        arena[f.fake_w_off + o] = fabsf(gamma);   // :264
        arena[f.fake_b_off + o] = beta;           // :265
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-tensor min/max over a task list (flat tiles), then in-place fake quantization
// ------------------------------------------------------------------------------------------------
struct FlatTask { int64_t off; int64_t n; int64_t minmax_off; int32_t num_bits; int32_t symmetric; };

__device__ __forceinline__ void cta_minmax_atomic(float mn, float mx, float* dst2, float* red) {
  mn = warp_min(mn); mx = warp_max(mx);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { red[w] = mn; red[kWarps + w] = mx; }
  __syncthreads();
  if (w == 0) {
    float a = red[l & (kWarps - 1)], b = red[kWarps + (l & (kWarps - 1))];
#pragma unroll
    for (int o = kWarps / 2; o > 0; o >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (l == 0) { atomic_min_f(dst2, a); atomic_max_f(dst2 + 1, b); }
  }
}

__device__ __forceinline__ void tile_minmax(const float* x, int64_t n, int64_t t, float& mn, float& mx) {
  const int64_t lo = t * kChunk;
  const int64_t hi = min(lo + (int64_t)kChunk, n);
  if ((((uintptr_t)x) & 15) == 0) {
    const float4* x4 = (const float4*)x;
    const int64_t lo4 = lo >> 2, hi4 = hi >> 2;
    for (int64_t i = lo4 + threadIdx.x; i < hi4; i += kThreads) {
      const float4 v = ldg_stream(x4 + i);
      mn = fminf(mn, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
      mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
    for (int64_t i = (hi4 << 2) + threadIdx.x; i < hi; i += kThreads) {
      const float v = ldg_stream1(x + i); mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
  } else {
    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
      const float v = ldg_stream1(x + i); mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
  }
}

__global__ void k_minmax_init(float* arena, const FlatTask* __restrict__ T, int nT) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nT; i += gridDim.x * blockDim.x) {
    arena[T[i].minmax_off] = DFQ_INF;
    arena[T[i].minmax_off + 1] = -DFQ_INF;
  }
}

__global__ void __launch_bounds__(kThreads)
k_minmax_tasks(float* arena, const FlatTask* __restrict__ T, int nT, const long long* __restrict__ tptr) {
  __shared__ float red[2 * kWarps];
  const TileSpan sp = tile_span(tptr, 0, nT);
  for (int ti = sp.q; ti < nT && tptr[ti] < sp.hi; ++ti) {
    const FlatTask t = T[ti];
    const long long base = tptr[ti];
    const long long k0 = max(sp.lo, base) - base, k1 = min(sp.hi, tptr[ti + 1]) - base;
    if (k1 <= k0) continue;
    float mn = DFQ_INF, mx = -DFQ_INF;
    for (long long k = k0; k < k1; ++k) tile_minmax(arena + t.off, t.n, k, mn, mx);
    cta_minmax_atomic(mn, mx, arena + t.minmax_off, red);
  }
}

template <bool RECIP>
__global__ void __launch_bounds__(kThreads)
k_quant_tasks(float* arena, const FlatTask* __restrict__ T, int nT, const long long* __restrict__ tptr) {
  const TileSpan sp = tile_span(tptr, 0, nT);
  for (int ti = sp.q; ti < nT && tptr[ti] < sp.hi; ++ti) {
    const FlatTask t = T[ti];
    const long long base = tptr[ti];
    const long long k0 = max(sp.lo, base) - base, k1 = min(sp.hi, tptr[ti + 1]) - base;
    if (k1 > k0) {
      // float(param.min()), float(param.max()) -> Python doubles (layer_transform.py:289,294)
      const QuantScalars q = quant_scalars((double)__ldcg(arena + t.minmax_off), (double)__ldcg(arena + t.minmax_off + 1),
                                           t.num_bits, t.symmetric);
      float* x = arena + t.off;
      for (long long k = k0; k < k1; ++k) {
        const int64_t lo = k * kChunk, hi = min(lo + (int64_t)kChunk, t.n);
        if ((t.off & 3) == 0) {
          float4* x4 = (float4*)x;
          const int64_t lo4 = lo >> 2, hi4 = hi >> 2;
          for (int64_t i = lo4 + threadIdx.x; i < hi4; i += kThreads) {
            float4 v = ldg_stream(x4 + i);
            v.x = fake_quant<RECIP>(v.x, q); v.y = fake_quant<RECIP>(v.y, q);
            v.z = fake_quant<RECIP>(v.z, q); v.w = fake_quant<RECIP>(v.w, q);
            stg_stream(x4 + i, v);
          }
          for (int64_t i = (hi4 << 2) + threadIdx.x; i < hi; i += kThreads) stg_stream1(x + i, fake_quant<RECIP>(ldg_stream1(x + i), q));
        } else {
          for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) stg_stream1(x + i, fake_quant<RECIP>(ldg_stream1(x + i), q));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bias correction engine (persistent cooperative kernel)
// ------------------------------------------------------------------------------------------------
// scipy.stats.norm.pdf / cdf in float64 on an fp32 argument, rounded to fp32 (dfq.py:182-183)
__device__ __forceinline__ float std_pdf(float x32) {
  const double x = (double)x32;
  return (float)(exp(-x * x / 2.0) / 2.5066282746310002);   // sqrt(2*pi)
}
__device__ __forceinline__ float std_cdf(float x32) {
  // scipy.special.ndtr (cephes)
  const double x = (double)x32 * 0.70710678118654752440;
  const double z = fabs(x);
  double y;
  if (z < 0.70710678118654752440) y = 0.5 + 0.5 * erf(x);
  else { y = 0.5 * erfc(z); if (x > 0) y = 1.0 - y; }
  return (float)y;
}
// dfq.py:184 calculate_mean + :240 clamp; every op an individually rounded fp32 op as in eager PyTorch
__device__ __forceinline__ float relu_gauss_mean(float g, float b) {
  const float q = __fdiv_rn(-b, g);
  const float t1 = __fmul_rn(g, std_pdf(q));
  const float t2 = __fmul_rn(b, __fsub_rn(1.0f, std_cdf(q)));
  const float e = __fadd_rn(t1, t2);
  return e < 0.f ? 0.f : e;      // NaN stays NaN (expect[expect < 0] = 0)
}

__global__ void __launch_bounds__(kThreads, 2)
k_bc_engine(float* arena, const DfqLayer* __restrict__ L, const DfqBcLayer* __restrict__ B, int nB,
            const DfqExpectTerm* __restrict__ T, const int* __restrict__ level_ptr, int n_levels, int num_bits,
            const long long* __restrict__ mm_ptr, const long long* __restrict__ row_ptr) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float red[2 * kWarps];
  __shared__ double dred[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- per-tensor min/max of every corrected weight (dfq.py:14 via :218) ------------------------
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < nB; i += gridDim.x * kThreads) {
    __stcg(arena + B[i].minmax_off, DFQ_INF);
    __stcg(arena + B[i].minmax_off + 1, -DFQ_INF);
  }
  grid.sync();
  {
    const TileSpan sp = tile_span(mm_ptr, 0, nB);
    for (int bi = sp.q; bi < nB && mm_ptr[bi] < sp.hi; ++bi) {
      const long long base = mm_ptr[bi];
      const long long k0 = max(sp.lo, base) - base, k1 = min(sp.hi, mm_ptr[bi + 1]) - base;
      if (k1 <= k0) continue;
      const DfqLayer l = L[B[bi].layer];
      const int64_t n = (int64_t)l.rows * l.cols * l.kk;
      float mn = DFQ_INF, mx = -DFQ_INF;
      for (long long k = k0; k < k1; ++k) tile_minmax(arena + l.w_off, n, k, mn, mx);
      cta_minmax_atomic(mn, mx, arena + B[bi].minmax_off, red);
    }
  }
  grid.sync();

  for (int lev = 0; lev < n_levels; ++lev) {
    // ---- E[x] of every layer of this level (dfq.py:228-278); one CTA per layer, terms in order ----
    for (int bi = level_ptr[lev] + blockIdx.x; bi < level_ptr[lev + 1]; bi += gridDim.x) {
      const DfqBcLayer b = B[bi];
      float* ex = arena + b.expect_off;
      for (int ti = b.term_begin; ti < b.term_end; ++ti) {
        const DfqExpectTerm t = T[ti];
        for (int ch = threadIdx.x; ch < t.n; ch += kThreads) {
          const float fb = __ldcg(arena + t.bn_b_off + ch);
          const float v = t.relu ? relu_gauss_mean(__ldcg(arena + t.bn_w_off + ch), fb) : fb;
          float* d = ex + t.dst_off + ch;
          __stcg(d, t.accumulate ? __fadd_rn(__ldcg(d), v) : v);
        }
        __syncthreads();
      }
    }
    grid.sync();
    // ---- eps . E[x] per output row (dfq.py:216-219,281-293) ------------------------------------------
    const TileSpan sp = tile_span(row_ptr, level_ptr[lev], level_ptr[lev + 1]);
    for (int bi = sp.q; bi < level_ptr[lev + 1] && row_ptr[bi] < sp.hi; ++bi) {
      const long long base = row_ptr[bi];
      const int t0 = (int)(max(sp.lo, base) - base), t1 = (int)(min(sp.hi, row_ptr[bi + 1]) - base);
      const DfqBcLayer b = B[bi];
      const DfqLayer l = L[b.layer];
      const int row_len = l.cols * l.kk;
      const bool cta_row = row_len > 2048;
      if (t1 > t0) {
        const QuantScalars q = quant_scalars((double)__ldcg(arena + b.minmax_off), (double)__ldcg(arena + b.minmax_off + 1),
                                             num_bits, b.signed_mode);
        const int G = b.expect_len / l.cols;
        const int so = l.rows / G;
        for (int t = t0; t < t1; ++t) {
          const int o = cta_row ? t : t * kWarps + warp;
          const bool live = o < l.rows;
          double acc = 0.0;
          if (live) {
            const float* rowp = arena + l.w_off + (size_t)o * row_len;
            const float* ex = arena + b.expect_off + (size_t)(o / so) * l.cols;
            const int tid = cta_row ? (int)threadIdx.x : lane;
            const int tpr = cta_row ? kThreads : 32;
            for (int j = tid; j < l.cols; j += tpr) {
              float E = 0.f;
              const float* p = rowp + (size_t)j * l.kk;
              if (b.flags & 1) {          // raw weight sum: bias absorption, dfq.py:150-153
                for (int k = 0; k < l.kk; ++k) E = __fadd_rn(E, p[k]);
              } else {
                for (int k = 0; k < l.kk; ++k) {
                  const float w = p[k];
                  E = __fadd_rn(E, __fsub_rn(fake_quant<false>(w, q), w));
                }
              }
              acc += (double)E * (double)__ldcg(ex + j);
            }
          }
          acc = warp_sum(acc);
          if (cta_row) {
            __syncthreads();
            if (lane == 0) dred[warp] = acc;
            __syncthreads();
            acc = 0.0;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) acc += dred[i];
          }
          const bool leader = live && (cta_row ? threadIdx.x == 0 : lane == 0);
          if (leader) {
            const float d = (float)acc;
            __stcg(arena + b.delta_off + o, d);
            __stcg(arena + l.bias_off + o, __fadd_rn(__ldcg(arena + l.bias_off + o), (b.flags & 2) ? d : -d));  // dfq.py:292 / :164
            if (b.next_bn_b_off >= 0)                                                                // dfq.py:204-206,293
              __stcg(arena + b.next_bn_b_off + o, __fadd_rn(__ldcg(arena + b.next_bn_b_off + o), -d));
          }
        }
      }
    }
    grid.sync();
  }
}

}  // namespace dfq

using namespace dfq;

static int pick_grid(const void* kernel, int64_t max_tiles, int* grid) {
  int dev = 0, sms = 0, per_sm = 0;
  DFQ_CUDA(cudaGetDevice(&dev));
  DFQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DFQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
  if (per_sm < 1) { set_error("kernel does not fit on an SM"); return DFQ_E_NOT_COOPERATIVE; }
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * per_sm, max_tiles));
  return 0;
}

extern "C" int dfq_bn_fold(float* arena, int64_t arena_floats, const DfqLayer* layers, int32_t n_layers,
                           const DfqFold* folds, int32_t n_folds, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DFQ_REQUIRE(arena && layers && folds, "null argument");
  if (n_folds <= 0) return 0;
  std::vector<long long> tptr(n_folds + 1, 0);
  for (int i = 0; i < n_folds; ++i) {
    DFQ_REQUIRE(folds[i].layer >= 0 && folds[i].layer < n_layers, "fold layer index");
    const DfqLayer& l = layers[folds[i].layer];
    DFQ_REQUIRE(l.w_off >= 0 && l.w_off + (int64_t)l.rows * l.cols * l.kk <= arena_floats, "weight outside arena");
    tptr[i + 1] = tptr[i] + ((l.cols * l.kk > 2048) ? l.rows : (l.rows + kWarps - 1) / kWarps);
  }
  int grid, rc;
  if ((rc = pick_grid((const void*)k_bn_fold, tptr[n_folds], &grid))) return rc;
  DfqLayer* dL; DfqFold* dF; long long* dP;
  if ((rc = upload(layers, n_layers, &dL, st))) return rc;
  if ((rc = upload(folds, n_folds, &dF, st))) return rc;
  if ((rc = upload(tptr.data(), n_folds + 1, &dP, st))) return rc;
  k_bn_fold<<<grid, kThreads, 0, st>>>(arena, dL, dF, n_folds, dP);
  DFQ_CUDA(cudaGetLastError());
  free_async(dL, st); free_async(dF, st); free_async(dP, st);
  return 0;
}

extern "C" int dfq_quantize_tensors(float* arena, int64_t arena_floats, const DfqQuantTask* tasks, int32_t n_tasks,
                                    int div_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DFQ_REQUIRE(arena && tasks, "null argument");
  if (n_tasks <= 0) return 0;
  std::vector<FlatTask> ft(n_tasks);
  std::vector<long long> tptr(n_tasks + 1, 0);
  for (int i = 0; i < n_tasks; ++i) {
    DFQ_REQUIRE(tasks[i].off >= 0 && tasks[i].n > 0 && tasks[i].off + tasks[i].n <= arena_floats, "tensor outside arena");
    DFQ_REQUIRE(tasks[i].minmax_off >= 0 && tasks[i].minmax_off + 2 <= arena_floats, "minmax scratch outside arena");
    DFQ_REQUIRE(tasks[i].num_bits >= 1 && tasks[i].num_bits <= 32, "num_bits");
    ft[i] = {tasks[i].off, tasks[i].n, tasks[i].minmax_off, tasks[i].num_bits, tasks[i].symmetric};
    tptr[i + 1] = tptr[i] + (tasks[i].n + kChunk - 1) / kChunk;
  }
  int grid, rc;
  if ((rc = pick_grid((const void*)k_minmax_tasks, tptr[n_tasks], &grid))) return rc;
  FlatTask* dT; long long* dP;
  if ((rc = upload(ft.data(), n_tasks, &dT, st))) return rc;
  if ((rc = upload(tptr.data(), n_tasks + 1, &dP, st))) return rc;
  k_minmax_init<<<std::min(148, (n_tasks + 255) / 256), 256, 0, st>>>(arena, dT, n_tasks);
  k_minmax_tasks<<<grid, kThreads, 0, st>>>(arena, dT, n_tasks, dP);
  if (div_mode) k_quant_tasks<true><<<grid, kThreads, 0, st>>>(arena, dT, n_tasks, dP);
  else          k_quant_tasks<false><<<grid, kThreads, 0, st>>>(arena, dT, n_tasks, dP);
  DFQ_CUDA(cudaGetLastError());
  free_async(dT, st); free_async(dP, st);
  return 0;
}

extern "C" int dfq_bias_correct(float* arena, int64_t arena_floats, const DfqLayer* layers, int32_t n_layers,
                                const DfqBcLayer* bc, int32_t n_bc, const DfqExpectTerm* terms, int32_t n_terms,
                                const int32_t* level_ptr, int32_t n_levels, int32_t num_bits, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DFQ_REQUIRE(arena && layers && bc && terms && level_ptr, "null argument");
  if (n_bc <= 0 || n_levels <= 0) return 0;
  DFQ_REQUIRE(level_ptr[0] == 0 && level_ptr[n_levels] == n_bc, "levels must partition the layer list");
  int64_t max_tiles = 1, mm_tiles = 0;
  std::vector<long long> mm_ptr(n_bc + 1, 0), row_ptr(n_bc + 1, 0);
  for (int i = 0; i < n_bc; ++i) {
    const DfqBcLayer& b = bc[i];
    DFQ_REQUIRE(b.layer >= 0 && b.layer < n_layers, "bc layer index");
    const DfqLayer& l = layers[b.layer];
    DFQ_REQUIRE(b.expect_len > 0 && b.expect_len % l.cols == 0, "expectation length must be groups*cols");
    DFQ_REQUIRE(l.rows % (b.expect_len / l.cols) == 0, "rows not divisible by groups");
    DFQ_REQUIRE(b.term_begin >= 0 && b.term_end <= n_terms && b.term_begin <= b.term_end, "term range");
    for (int t = b.term_begin; t < b.term_end; ++t)
      DFQ_REQUIRE(terms[t].dst_off >= 0 && terms[t].dst_off + terms[t].n <= b.expect_len, "term outside expectation vector");
    DFQ_REQUIRE(b.expect_off >= 0 && b.expect_off + b.expect_len <= arena_floats, "expect scratch outside arena");
    mm_tiles += ((int64_t)l.rows * l.cols * l.kk + kChunk - 1) / kChunk;
    mm_ptr[i + 1] = mm_tiles;
    row_ptr[i + 1] = row_ptr[i] + ((l.cols * l.kk > 2048) ? l.rows : (l.rows + kWarps - 1) / kWarps);
  }
  max_tiles = std::max(max_tiles, mm_tiles);
  for (int lev = 0; lev < n_levels; ++lev) {
    int64_t t = 0;
    for (int i = level_ptr[lev]; i < level_ptr[lev + 1]; ++i) {
      const DfqLayer& l = layers[bc[i].layer];
      t += (l.cols * l.kk > 2048) ? l.rows : (l.rows + kWarps - 1) / kWarps;
    }
    max_tiles = std::max(max_tiles, t);
  }
  int grid, rc;
  if ((rc = pick_grid((const void*)k_bc_engine, max_tiles, &grid))) return rc;
  DfqLayer* dL; DfqBcLayer* dB; DfqExpectTerm* dT; int32_t* dLP; long long *dMP, *dRP;
  if ((rc = upload(layers, n_layers, &dL, st))) return rc;
  if ((rc = upload(bc, n_bc, &dB, st))) return rc;
  if ((rc = upload(terms, n_terms, &dT, st))) return rc;
  if ((rc = upload(level_ptr, n_levels + 1, &dLP, st))) return rc;
  if ((rc = upload(mm_ptr.data(), n_bc + 1, &dMP, st))) return rc;
  if ((rc = upload(row_ptr.data(), n_bc + 1, &dRP, st))) return rc;
  void* args[] = {&arena, &dL, &dB, (void*)&n_bc, &dT, &dLP, (void*)&n_levels, (void*)&num_bits, &dMP, &dRP};
  DFQ_CUDA(cudaLaunchCooperativeKernel((void*)k_bc_engine, dim3(grid), dim3(kThreads), args, 0, st));
  free_async(dL, st); free_async(dB, st); free_async(dT, st); free_async(dLP, st); free_async(dMP, st); free_async(dRP, st);
  return 0;
}
