// Column extrema of row tiles: the running [C] min/max arrays of a `second` layer that the equalization solves s from
// (dfq.py:54-55 `weight_second.view(...).max/min`).  Shared by the equalization engine (initial scan through its ring) and
// the BN fold (which has every folded tile in shared memory anyway and fills the arrays for free).
#pragma once
#include "common.cuh"

namespace dfq {

// Rows [row0, row0 + nrows) of a layer with J columns of kk taps per row, `go` rows and `gi` columns per group.
// One item = the kk taps of one (row, column).  `buf` = the tile (shared memory, or global memory with GLOBAL = true).
// The partial extrema go to the CTA's shared-memory arrays smin/smax indexed by column (`own`: the calling thread is the only
// one that ever touches its columns -> plain read-modify-write; otherwise shared-memory atomics), or straight to the global
// arrays dmin/dmax when the layer has more columns than the scratch holds (use_smem = false).
// NT threads cooperate; tid = 0 .. NT-1.
template <int NT, bool GLOBAL>
__device__ __forceinline__ void colscan_tile(const float* __restrict__ buf, int tid, int row0, int nrows, int J, int kk, int go,
                                             int gi, bool single_group, bool own, bool use_smem, float* smin, float* smax,
                                             float* dmin, float* dmax) {
  if (!GLOBAL && kk == 1 && single_group && use_smem) {
    // 1x1 rows of an ungrouped layer: a thread per COLUMN walks down the tile's rows (conflict-free, no division, no atomics).
    // Column j belongs to thread j % NT in every tile of the layer, and the scratch is private to the NT threads between two
    // flushes: plain read-modify-write.
    for (int j = tid; j < J; j += NT) {
      float mn = buf[j], mx = mn;
      for (int r = 1; r < nrows; ++r) { const float v = buf[(size_t)r * J + j]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
      smin[j] = fminf(smin[j], mn); smax[j] = fmaxf(smax[j], mx);
    }
    return;
  }
  const int items = nrows * J;
  for (int idx = tid; idx < items; idx += NT) {
    const float* p = buf + (size_t)idx * kk;
    float mn = GLOBAL ? ldg_stream1(p) : p[0], mx = mn;
    if (!GLOBAL && kk == 9) {
#pragma unroll
      for (int k = 1; k < 9; ++k) { const float v = p[k]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    } else {
      for (int k = 1; k < kk; ++k) { const float v = GLOBAL ? ldg_stream1(p + k) : p[k]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    }
    int row = 0, j = idx;
    if (nrows > 1) { row = idx / J; j = idx - row * J; }
    const int col = single_group ? j : ((row0 + row) / go) * gi + j;
    if (use_smem) {
      if (own) { smin[col] = fminf(smin[col], mn); smax[col] = fmaxf(smax[col], mx); }
      else { atomic_min_f(smin + col, mn); atomic_max_f(smax + col, mx); }
    } else {
      atomic_min_f(dmin + col, mn); atomic_max_f(dmax + col, mx);
    }
  }
}

// Fold the CTA's partial extrema into the global arrays and leave the scratch reset.  Caller: barrier before and after.
template <int NT>
__device__ __forceinline__ void colscan_flush(int tid, int nch, float* smin, float* smax, float* dmin, float* dmax) {
  for (int j = tid; j < nch; j += NT) {
    const float mn = smin[j], mx = smax[j];
    if (mn != DFQ_INF || mx != -DFQ_INF) {
      atomic_min_f(dmin + j, mn); atomic_max_f(dmax + j, mx);
      smin[j] = DFQ_INF; smax[j] = -DFQ_INF;
    }
  }
}

}  // namespace dfq
