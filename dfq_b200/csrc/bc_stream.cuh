// k_bc_stream: bias correction (dfq.py:173-293) for LARGE phases - the bandwidth-bound variant of k_bc_engine.
//
// k_bc_engine processes one tile per CTA at a time: every 18 KB row pays two CTA-wide barriers, a serial 8-way fp64
// sum and three dependent read-modify-writes by thread 0, and every element an IEEE division (MUFU.RCP + FCHK + 5 FFMA)
// and an FRND - 2.15 TB/s on the synthetic stack (round 1: barrier 4.0 and wait 2.3 of the warp stall samples, issue
// slots 52 % busy with the XU pipe saturated by MUFU.RCP + FRND: 2 x 4608 quarter-rate ops per row).
//
// Here every consumer WARP owns whole tiles:
//   * one CTA per SM: 8 consumer warps + 1 producer warp, a ring of kBcStages 18 KB stages (all of the SM's shared
//     memory); per stage a `full` mbarrier (TMA transaction bytes) and an `empty` mbarrier (the consuming warp's lane 0):
//     NO CTA-wide barrier anywhere in the streaming loop;
//   * the producer's lane 0 walks the CTA's tile list (same block-cyclic order as RowPipe), waits for the stage to be
//     empty, writes the descriptor and issues the bulk copy; consumer w takes sequence numbers n = w (mod 8).  A phase
//     ends with padding SKIP items up to a multiple of 8 and then 8 END items, so every consumer sees exactly one END;
//   * a tile's rows are reduced by the warp alone (lanes over input columns, 9 taps each, warp shuffle at the end);
//     lane r requests row r's read-modify-write operands before the tile and retires them after it;
//   * the expectation vector of the current (layer, group) stays in REGISTERS (cols <= 512: 16 per lane) across tiles;
//   * the quantization error is computed with no XU-pipe instruction:
//       - the quotient t / scale as Markstein's correction of the product with the correctly rounded reciprocal
//         (q0 = t*y, r = fma(-q0, s, t), q = fma(r, y, q0); y = __frcp_rn(s)): equal to the IEEE quotient when no
//         intermediate underflows and the mantissa of s is not all ones - both checked once per tensor (BcFastQuant::ok),
//         otherwise the plain __fdiv_rn chain runs for that tensor.  t lies in [0, max-min], q in [qmin, qmax];
//       - rint() for |t| <= 2^22 as (t + 1.5*2^23) - 1.5*2^23 (two FADDs, round-half-even like FRND).
//     tests/test_gpu_engine.py::test_bc_fast_quotient_equals_ieee_division checks codes AND quotients against __fdiv_rn
//     over thousands of scales x dense numerators including every half-integer neighbourhood.
#pragma once
#include "common.cuh"
#include "rowpipe.cuh"

namespace dfq {

constexpr int kBcConsumers = 8;
constexpr int kBcThreads = (kBcConsumers + 1) * 32;
#ifndef DFQ_BC_STAGES
#define DFQ_BC_STAGES 11
#endif
constexpr int kBcStages = DFQ_BC_STAGES;
constexpr int kBcExRegs = 16;                 // expectation values per lane kept in registers (cols <= 512)

enum { TK_SKIP = 3, TK_END = 4 };

struct BcFastQuant {
  float neg_min, min_v, scale, rcp, qmin, qmax;
  int ok;        // 1: the three-instruction quotient is exact for every numerator this tensor can produce
};

__device__ __forceinline__ BcFastQuant bc_fast_quant(const QuantScalars& q, int num_bits) {
  BcFastQuant f;
  f.neg_min = q.neg_min; f.min_v = q.min_v; f.scale = q.scale; f.qmin = q.qmin; f.qmax = q.qmax;
  f.rcp = __frcp_rn(q.scale);
  const unsigned bits = __float_as_uint(q.scale);
  const int e = (int)((bits >> 23) & 0xff);
  // exponent window: the residual r = t - q0*s is ~2^-24 of t and must stay normal for every t whose quotient can reach
  // 0.5 (smaller t give code 0 or qmin's neighbourhood either way): t >= s/4 -> r >= s * 2^-27.  The reciprocal must be
  // normal as well.  Mantissa all ones: the one divisor for which RN(1/s) is not close enough (Markstein).
  f.ok = (num_bits <= 16) && e >= 40 && e <= 200 && ((bits & 0x7fffffu) != 0x7fffffu) && !(q.scale != q.scale);
  return f;
}

// quantize.py:70-74 for one element, error form: Q(w) - w
template <bool FAST>
__device__ __forceinline__ float bc_qerr(float w, const BcFastQuant& f) {
  float t = __fadd_rn(w, f.neg_min);
  if (FAST) {
    const float q0 = __fmul_rn(t, f.rcp);
    const float r = __fmaf_rn(-q0, f.scale, t);
    t = __fmaf_rn(r, f.rcp, q0);
    t = fminf(fmaxf(t, f.qmin), f.qmax);
    t = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
  } else {
    t = __fdiv_rn(t, f.scale);
    t = fminf(fmaxf(t, f.qmin), f.qmax);
    t = rintf(t);
  }
  t = __fmul_rn(t, f.scale);
  return __fsub_rn(__fadd_rn(t, f.min_v), w);
}

struct BcRing {
  unsigned char* base;
  __device__ __forceinline__ float* stage(int s) const { return (float*)(base + (size_t)s * kStageBytes); }
  uint64_t* full;      // [kBcStages]
  uint64_t* empty;     // [kBcStages]
  TileDesc* desc;      // [kBcStages]
  __device__ void init(unsigned char* smem) {
    base = smem;
    unsigned char* p = smem + (size_t)kBcStages * kStageBytes;
    full = (uint64_t*)p;
    empty = (uint64_t*)(p + 8 * kBcStages);
    desc = (TileDesc*)(p + 16 * kBcStages + 16);
    if (threadIdx.x == 0) {
      for (int i = 0; i < kBcStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
      mbar_fence_init();
    }
    __syncthreads();
  }
  static constexpr size_t smem_bytes() { return (size_t)kBcStages * kStageBytes + 16 * kBcStages + 16 + kBcStages * sizeof(TileDesc) + 64; }
};

// Producer lane: put item `d` at sequence number n.
__device__ __forceinline__ void bc_produce(BcRing& ring, unsigned long long n, const TileDesc& d) {
  const int s = (int)(n % kBcStages);
  if (n >= kBcStages) mbar_wait(ring.empty + s, (uint32_t)(((n / kBcStages) - 1) & 1));
  ring.desc[s] = d;
  if (d.kind == TK_BULK) {
    mbar_arrive_expect_tx(ring.full + s, (uint32_t)d.floats * 4u);
    bulk_g2s(ring.stage(s), d.gptr, (uint32_t)d.floats * 4u, ring.full + s);
  } else {
    mbar_arrive(ring.full + s);
  }
}

// One output row (shared or global memory) against an expectation vector held in registers / global memory.
template <bool FAST, bool RAW, bool EXREG>
__device__ __forceinline__ double bc_stream_row(const float* __restrict__ row, int cols, int kk, const float* __restrict__ ex,
                                                const float (&exr)[kBcExRegs], const BcFastQuant& f, int lane) {
  double acc = 0.0;
  if (EXREG) {
#pragma unroll
    for (int c = 0; c < kBcExRegs; ++c) {
      const int j = lane + 32 * c;
      if (j < cols) {
        const float* p = row + (size_t)j * kk;
        float E = 0.f;
        if (kk == 9) {
#pragma unroll
          for (int k = 0; k < 9; ++k) E = __fadd_rn(E, RAW ? p[k] : bc_qerr<FAST>(p[k], f));
        } else {
          for (int k = 0; k < kk; ++k) E = __fadd_rn(E, RAW ? p[k] : bc_qerr<FAST>(p[k], f));
        }
        acc += (double)E * (double)exr[c];
      }
    }
  } else {
    for (int j = lane; j < cols; j += 32) {
      const float* p = row + (size_t)j * kk;
      float E = 0.f;
      for (int k = 0; k < kk; ++k) E = __fadd_rn(E, RAW ? p[k] : bc_qerr<FAST>(p[k], f));
      acc += (double)E * (double)__ldcg(ex + j);
    }
  }
  return warp_sum(acc);
}

}  // namespace dfq
