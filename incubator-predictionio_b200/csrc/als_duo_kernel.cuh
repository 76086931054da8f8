// als_duo_kernel.cuh -- rank 33..64 half-step, third generation: a WORKER is two warps that accumulate one destination
// row's Gramian together -- warp 0 owns the accumulator tiles of m-tiles {0, 3}, warp 1 those of m-tiles {1, 2} (ten
// 16x8 tiles each) -- over four rows in a round, and then each warp solves two of the four normal equations with the
// lockstep Cholesky (als_lockstep.cuh).
//
// Why split a row over two warps (ncu on the pair kernel, profiles/r02_pair_kernel_summary.md): one warp holding all
// twenty accumulator tiles (80 registers) has room for only 2-3 tiles in flight, and every tile is a chain of three
// dependent HMMAs plus an FADD -- `wait` was 36-53 % of the samples of the accumulate loop; and because all four
// 16-feature blocks served as both A and B fragments, ~90 MOVs per chunk rebuilt operand pairs.  With ten tiles per warp
// there are registers for all ten chains at once, and most blocks now serve one role only (warp 0: blocks 1, 2 are B
// fragments only; warp 1: block 0), so few pairs need MOVs.  The staging ring is filled by both warps (two cp.async per
// lane and chunk) behind one 64-thread named barrier per chunk; the right-hand side is accumulated by warp 1 alone (it
// carries fewer splits).
//
// Same arithmetic per row as als_pair_kernel.cuh (same chunk order, same per-tile products, round-to-nearest sum over
// chunks; b summed over the same lanes), so the two kernels produce identical factors; same parts mode and finish kernel.
// Replaces NormalEquation.add + CholeskySolver.solve of Spark 2.4 ml.recommendation.ALS,
// reached from examples/scala-parallel-recommendation/.../ALSAlgorithm.scala:76-86.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "als_pair_kernel.cuh"

namespace pio {
namespace duo {

using pr::CH;
using pr::KP;
using pr::LL;
using pr::NSTAGE;
using pr::PART_FLOATS;
using pr::RSTR;
using pr::SLOT_STRIDE;
using pr::STAGE;
using pr::VSTR;

constexpr int ROWS_PER_ROUND = 4;
constexpr int WORKERS = 2;                       // per CTA (four warps)
// per-worker shared memory (floats): four slots (the staging ring aliases the last one), b vectors, pivot lines, ratings
constexpr int D_BVEC = ROWS_PER_ROUND * SLOT_STRIDE;
constexpr int D_COL = D_BVEC + ROWS_PER_ROUND * VSTR;
constexpr int D_MVAL = D_COL + ROWS_PER_ROUND * VSTR;
constexpr int D_FLOATS = D_MVAL + NSTAGE * CH + 8;
constexpr size_t SMEM_BYTES = sizeof(float) * (size_t)D_FLOATS * WORKERS;

__device__ __forceinline__ void worker_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// One warp's share of a row: tiles of m-tiles MT0 and MT1 (HALF 0: {0, 3}; HALF 1: {1, 2}).
template <bool IMPLICIT, int HALF>
__device__ __forceinline__ void accumulate_half(const SolveParams& p, long long beg, long long end, float* ring, float* mval,
                                                float* slot, float* bv, int bar_id) {
  constexpr int MT0 = HALF == 0 ? 0 : 1, MT1 = HALF == 0 ? 3 : 2;
  constexpr int NBLK = MT1 + 1;                   // 16-feature blocks this warp needs as B fragments
  constexpr int NT0 = 2 * MT0 + 2, NT = NT0 + 2 * MT1 + 2;   // tiles of MT0, all tiles (= 10)
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nchunks = (int)((end - beg + CH - 1) / CH);
  // staging: the worker's 64 lanes copy the 128 16-byte pieces of a chunk; lane's pieces: staged rows prow, prow + 4
  const int wl = HALF * 32 + lane;                // lane inside the worker
  const int prow = wl >> 4, psl = wl & 15;

  int nidx[2];
  float nval;
  auto prefetch_meta = [&](int c) {
    const long long e0 = beg + (long long)c * CH;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long e = e0 + prow + 4 * j;
      nidx[j] = (c < nchunks && e < end) ? __ldg(p.idx + e) : -1;
    }
    nval = 0.f;
    if (HALF == 0 && lane < CH && c < nchunks && e0 + lane < end) nval = __ldg(p.val + e0 + lane);
  };
  auto issue = [&](int c) {
    if (c < nchunks) {
      float* sbuf = ring + (c % NSTAGE) * STAGE;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float4* d4 = reinterpret_cast<float4*>(sbuf + (prow + 4 * j) * RSTR + psl * 4);
        if (nidx[j] >= 0) cp_async16(d4, p.src + (size_t)nidx[j] * KP + psl * 4);
        else *d4 = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (HALF == 0 && lane < CH) mval[(c % NSTAGE) * CH + lane] = nval;
    }
    cp_async_commit();
  };

  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  float pb[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) pb[i][0] = pb[i][1] = 0.f;

  prefetch_meta(0);
  issue(0);
  prefetch_meta(1);
  issue(1);
  prefetch_meta(2);

#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<1>();
    worker_barrier(bar_id);        // chunk c has landed for both warps; stage (c + 2) % 3 was consumed by both
    issue(c + 2);
    prefetch_meta(c + 3);
    const float* X = ring + (c % NSTAGE) * STAGE;
    const float* mv = mval + (c % NSTAGE) * CH;
    const float r0 = mv[t], r1 = mv[t + 4];
    float sc0 = 1.f, sc1 = 1.f, wb0 = r0, wb1 = r1;
    if (IMPLICIT) {
      const float c0 = p.alpha * fabsf(r0), c1 = p.alpha * fabsf(r1);
      sc0 = sqrtf(c0);
      sc1 = sqrtf(c1);
      wb0 = r0 > 0.f ? 1.f + c0 : 0.f;
      wb1 = r1 > 0.f ? 1.f + c1 : 0.f;
    }
    uint32_t hi[NBLK][4], lo[NBLK][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      // warp 0 splits every block; warp 1 reads block 3 for the right-hand side only
      float v[4];
      v[0] = X[t * RSTR + 16 * b + g];
      v[1] = X[t * RSTR + 16 * b + 8 + g];
      v[2] = X[(t + 4) * RSTR + 16 * b + g];
      v[3] = X[(t + 4) * RSTR + 16 * b + 8 + g];
      if (HALF == 1) {   // the right-hand side rides on warp 1 (same lanes, same order as the pair kernel)
        pb[b][0] = fmaf(wb0, v[0], pb[b][0]);
        pb[b][1] = fmaf(wb0, v[1], pb[b][1]);
        pb[b][0] = fmaf(wb1, v[2], pb[b][0]);
        pb[b][1] = fmaf(wb1, v[3], pb[b][1]);
      }
      if (b >= NBLK) continue;
      if (IMPLICIT) {
        v[0] *= sc0; v[1] *= sc0; v[2] *= sc1; v[3] *= sc1;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float cc = __fmul_rn(v[e], 8193.f);       // Veltkamp split (see als_pair_kernel.cuh)
        const float h = __fsub_rn(cc, __fsub_rn(cc, v[e]));
        hi[b][e] = __float_as_uint(h);
        lo[b][e] = __float_as_uint(__fsub_rn(v[e], h));
      }
    }
#pragma unroll
    for (int j = 0; j < 2 * MT1 + 2; ++j) {
      const int bi = j >> 1, be = j & 1;
      const uint32_t bh0 = hi[bi][be], bh1 = hi[bi][be + 2], bl0 = lo[bi][be], bl1 = lo[bi][be + 2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int mt = s == 0 ? MT0 : MT1;
        if (j > 2 * mt + 1) continue;
        const int tile = (s == 0 ? 0 : NT0) + j;
        float d[4];
        pr::mma_tf32_z(d, lo[mt], bh0, bh1);
        pr::mma_tf32(d, hi[mt], bl0, bl1);
        pr::mma_tf32(d, hi[mt], bh0, bh1);
        acc[tile][0] += d[0];
        acc[tile][1] += d[1];
        acc[tile][2] += d[2];
        acc[tile][3] += d[3];
      }
    }
  }
  cp_async_wait<0>();
  worker_barrier(bar_id);   // the ring is dead for both warps: the last slot may be written

  if (HALF == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v = pb[i][e];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (t == 0) bv[16 * i + 8 * e + g] = v;
      }
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int mt = s == 0 ? MT0 : MT1;
#pragma unroll
    for (int j = 0; j <= 2 * mt + 1; ++j) {
      const int tile = (s == 0 ? 0 : NT0) + j;
      const int cb = j >> 1;
      const int cc = 8 * (j & 1) + 2 * t;
      if (cb < mt) {
        *reinterpret_cast<float2*>(slot + LL::offd(mt, cb, g, cc)) = make_float2(acc[tile][0], acc[tile][1]);
        *reinterpret_cast<float2*>(slot + LL::offd(mt, cb, g + 8, cc)) = make_float2(acc[tile][2], acc[tile][3]);
      } else {
        if (cc <= g) slot[LL::diag(mt, g, cc)] = acc[tile][0];
        if (cc + 1 <= g) slot[LL::diag(mt, g, cc + 1)] = acc[tile][1];
        if (cc <= g + 8) slot[LL::diag(mt, g + 8, cc)] = acc[tile][2];
        if (cc + 1 <= g + 8) slot[LL::diag(mt, g + 8, cc + 1)] = acc[tile][3];
      }
    }
  }
  worker_barrier(bar_id);   // slot and b are complete
}

template <bool IMPLICIT>
__global__ void __launch_bounds__(64 * WORKERS, 3) als_solve_duo_kernel(const SolveParams p, int n_items) {
  extern __shared__ __align__(16) float smem_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wk = warp >> 1, half = warp & 1;
  float* smem = smem_all + wk * D_FLOATS;
  float* ring = smem + (ROWS_PER_ROUND - 1) * SLOT_STRIDE;   // aliases the last slot
  float* bvec = smem + D_BVEC;
  float* colbuf = smem + D_COL;
  float* mval = smem + D_MVAL;
  const int grp = lane >> 4;
  const int nquads = (n_items + ROWS_PER_ROUND - 1) / ROWS_PER_ROUND;
  const int bar_id = 1 + wk;

#pragma unroll 1
  for (int base = blockIdx.x * WORKERS; base < nquads; base += gridDim.x * WORKERS) {
    const int quad = base + wk;
    int row0 = -1, row1 = -1, row2 = -1, row3 = -1;
    if (quad < nquads) {
#pragma unroll 1
      for (int h = 0; h < ROWS_PER_ROUND; ++h) {
        const int item = ROWS_PER_ROUND * quad + h;
        float* slot = smem + h * SLOT_STRIDE;
        float* bv = bvec + h * VSTR;
        if (item >= n_items) {
          if (half == 0) pr::fill_identity(slot, bv);
          worker_barrier(bar_id);
          continue;
        }
        long long beg, end;
        if (p.partial) {
          beg = p.wl_beg[item];
          end = p.wl_end[item];
        } else {
          const int r = p.row_begin + item;
          beg = p.ptr[r];
          end = p.ptr[r + 1];
          if (h == 0) row0 = r;
          else if (h == 1) row1 = r;
          else if (h == 2) row2 = r;
          else row3 = r;
        }
        if (half == 0) accumulate_half<IMPLICIT, 0>(p, beg, end, ring, mval, slot, bv, bar_id);
        else accumulate_half<IMPLICIT, 1>(p, beg, end, ring, mval, slot, bv, bar_id);
        if (p.partial) {
          float* out = p.partial + (size_t)item * PART_FLOATS;
          const int wl = half * 32 + lane;
          for (int o = wl; o < LL::SIZE / 4; o += 64)
            reinterpret_cast<float4*>(out)[o] = reinterpret_cast<const float4*>(slot)[o];
          for (int o = wl; o < KP; o += 64) out[LL::SIZE + o] = bv[o];
          worker_barrier(bar_id);
        }
      }
    }
    if (p.partial) continue;
    __syncthreads();   // the CTA's warps enter the (large, unrolled) solver together: instruction-cache locality
    if (quad < nquads) {
      // warp `half` solves rows 2 * half and 2 * half + 1 of the quad, one per 16-lane group
      const int h = 2 * half + grp;
      const int myrow = h == 0 ? row0 : h == 1 ? row1 : h == 2 ? row2 : row3;
      const int rr = myrow < 0 ? p.row_begin : myrow;
      chol_lockstep<KP, IMPLICIT>(smem + h * SLOT_STRIDE, bvec + h * VSTR, p.yty, p.lambda * p.nreg[rr], p.k,
                                  colbuf + h * VSTR, p.dst + (size_t)(p.dst_row_offset + rr) * KP, myrow >= 0, p.fail);
    }
    __syncthreads();   // slots are reused by the next round
  }
}

}  // namespace duo
}  // namespace pio
