// topk.cuh -- top-k scoring over the item factor matrix (serving side of the hot path) and the
// multinomial NaiveBayes reductions of the classification template.
//
//  score_dot_topk_kernel    : recommendProducts(WithFilter) -- <x_u, y_i> over all candidate items
//      (examples/scala-parallel-recommendation/blacklist-items/src/main/scala/ALSModel.scala:44-60)
//  score_cos_topk_kernel    : similarproduct predict -- sum_q cosine(y_q, y_i), score > 0 only
//      (examples/scala-parallel-similarproduct/multi-events-multi-algos/src/main/scala/ALSAlgorithm.scala:160-187,220-234)
//  topk_merge_kernel        : merges the per-tile winners (getTopN, :200-217)
//
// Scores are accumulated in fp64 over the fp32 factors in index order, exactly like the reference's
// blas.ddot / cosine loops over Array[Double], so scores and rankings are bit-identical to the oracle.
// These are HBM-bound scans: every item row is read once per query batch tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pio {

constexpr int TK_THREADS = 256;
constexpr int TK_ITEMS = 4;                       // items per thread
constexpr int TK_TILE = TK_THREADS * TK_ITEMS;    // items per CTA
constexpr int TK_MAXK = 128;                      // max supported topk

struct ScoreIdx {
  double s;
  int i;
};
__device__ __forceinline__ bool better(double s1, int i1, double s2, int i2) {
  return (s1 > s2) || (s1 == s2 && i1 < i2);
}

// Block-wide selection: every thread holds TK_ITEMS candidates (score, index; index -1 = none).
// Extracts the best `topk` in order and writes them to out[0..topk).
__device__ __forceinline__ void block_select_topk(double (&sc)[TK_ITEMS], int (&ix)[TK_ITEMS], int topk,
                                                  ScoreIdx* out) {
  __shared__ double ws[TK_THREADS / 32];
  __shared__ int wi[TK_THREADS / 32];
  __shared__ int wowner[TK_THREADS / 32];
  __shared__ int s_owner;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t = 0; t < topk; ++t) {
    // local best
    double bs = 0.0;
    int bi = -1, bslot = -1;
#pragma unroll
    for (int j = 0; j < TK_ITEMS; ++j)
      if (ix[j] >= 0 && (bi < 0 || better(sc[j], ix[j], bs, bi))) { bs = sc[j]; bi = ix[j]; bslot = j; }
    double rs = bs;
    int ri = bi, ro = threadIdx.x;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double os = __shfl_xor_sync(0xffffffffu, rs, d);
      const int oi = __shfl_xor_sync(0xffffffffu, ri, d);
      const int oo = __shfl_xor_sync(0xffffffffu, ro, d);
      if (oi >= 0 && (ri < 0 || better(os, oi, rs, ri))) { rs = os; ri = oi; ro = oo; }
    }
    if (lane == 0) { ws[w] = rs; wi[w] = ri; wowner[w] = ro; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double fs = ws[0];
      int fi = wi[0], fo = wowner[0];
      for (int q = 1; q < TK_THREADS / 32; ++q)
        if (wi[q] >= 0 && (fi < 0 || better(ws[q], wi[q], fs, fi))) { fs = ws[q]; fi = wi[q]; fo = wowner[q]; }
      out[t].s = fs;
      out[t].i = fi;
      s_owner = fi >= 0 ? fo : -1;
    }
    __syncthreads();
    if (s_owner == (int)threadIdx.x && bslot >= 0) ix[bslot] = -1;
    __syncthreads();
  }
}

// grid: (n_tiles, n_queries). xq: query vectors [n_queries][KP] (device, zero padded);
// qvalid[q] == 0 -> no candidates. cand: [n_queries][n_tiles][topk].
__global__ void __launch_bounds__(TK_THREADS)
score_dot_topk_kernel(const float* __restrict__ Y, int n_items, int kp, int k,
                      const float* __restrict__ xq, const uint8_t* __restrict__ qvalid,
                      const int* __restrict__ cand_ext, const uint8_t* __restrict__ mask,
                      int topk, ScoreIdx* __restrict__ cand) {
  extern __shared__ float sx[];
  const int q = blockIdx.y;
  for (int t = threadIdx.x; t < k; t += TK_THREADS) sx[t] = xq[(size_t)q * kp + t];
  __syncthreads();
  double sc[TK_ITEMS];
  int ix[TK_ITEMS];
  const bool qok = qvalid[q] != 0;
#pragma unroll
  for (int j = 0; j < TK_ITEMS; ++j) {
    const int i = blockIdx.x * TK_TILE + j * TK_THREADS + threadIdx.x;
    sc[j] = 0.0;
    ix[j] = -1;
    const int ext = (qok && i < n_items) ? cand_ext[i] : -1;  // external id, -1 = owns no factor
    if (ext >= 0 && !(mask && mask[ext])) {
      const float* y = Y + (size_t)i * kp;
      double s = 0.0;
      for (int t = 0; t < k; ++t) s += (double)sx[t] * (double)y[t];
      sc[j] = s;
      ix[j] = ext;
    }
  }
  block_select_topk(sc, ix, topk, cand + ((size_t)q * gridDim.x + blockIdx.x) * topk);
}

// one query = a set of item vectors. qf: [nqv][KP] vectors of the query items that own a factor
// (query order kept); qid: all nq_all query item ids (external) -- every one of them is excluded
// from the candidates (ALSAlgorithm.scala:243-245 `!queryList.contains(i)`).
__global__ void __launch_bounds__(TK_THREADS)
score_cos_topk_kernel(const float* __restrict__ Y, int n_items, int kp, int k,
                      const float* __restrict__ qf, const int* __restrict__ qid, int nq_all, int nqv,
                      const int* __restrict__ cand_ext, const uint8_t* __restrict__ mask,
                      int topk, ScoreIdx* __restrict__ cand) {
  double sc[TK_ITEMS];
  int ix[TK_ITEMS];
#pragma unroll
  for (int j = 0; j < TK_ITEMS; ++j) {
    const int i = blockIdx.x * TK_TILE + j * TK_THREADS + threadIdx.x;
    sc[j] = 0.0;
    ix[j] = -1;
    const int ext = i < n_items ? cand_ext[i] : -1;
    if (ext >= 0 && !(mask && mask[ext])) {
      bool isq = false;
      for (int t = 0; t < nq_all; ++t) isq |= (qid[t] == ext);
      if (!isq) {
        const float* f = Y + (size_t)i * kp;
        double score = 0.0;
        for (int t = 0; t < nqv; ++t) {
          const float* v1 = qf + (size_t)t * kp;
          double n1 = 0.0, n2 = 0.0, d = 0.0;
          for (int c = 0; c < k; ++c) {
            const double a = (double)v1[c], b = (double)f[c];
            n1 += a * a;
            n2 += b * b;
            d += a * b;
          }
          const double n1n2 = sqrt(n1) * sqrt(n2);
          score += (n1n2 == 0.0) ? 0.0 : d / n1n2;
        }
        if (score > 0.0) { sc[j] = score; ix[j] = ext; }
      }
    }
  }
  block_select_topk(sc, ix, topk, cand + (size_t)blockIdx.x * topk);
}

// grid: n_queries. Merges n_tiles*topk candidates per query -> final topk.
__global__ void __launch_bounds__(TK_THREADS)
topk_merge_kernel(const ScoreIdx* __restrict__ cand, int n_cand, int topk, int* __restrict__ out_items,
                  float* __restrict__ out_scores, int* __restrict__ out_count) {
  __shared__ ScoreIdx best[TK_MAXK];
  const ScoreIdx* c = cand + (size_t)blockIdx.x * n_cand;
  // candidates may exceed TK_TILE: fold them through repeated selection rounds
  double sc[TK_ITEMS];
  int ix[TK_ITEMS];
  __shared__ ScoreIdx carry[TK_MAXK];
  int ncarry = 0;
  for (int base = 0; base < n_cand || base == 0; base += TK_TILE - TK_MAXK) {
    // slots [0, ncarry) of this round come from carry, the rest from cand[base...]
#pragma unroll
    for (int j = 0; j < TK_ITEMS; ++j) {
      const int slot = j * TK_THREADS + threadIdx.x;
      sc[j] = 0.0;
      ix[j] = -1;
      if (slot < ncarry) {
        sc[j] = carry[slot].s;
        ix[j] = carry[slot].i;
      } else {
        const int o = base + slot - ncarry;
        if (slot - ncarry < TK_TILE - TK_MAXK && o < n_cand) { sc[j] = c[o].s; ix[j] = c[o].i; }
      }
    }
    __syncthreads();
    block_select_topk(sc, ix, topk, best);
    __syncthreads();
    for (int t = threadIdx.x; t < topk; t += TK_THREADS) carry[t] = best[t];
    ncarry = topk;
    __syncthreads();
    if (base + (TK_TILE - TK_MAXK) >= n_cand) break;
  }
  int cnt = 0;
  for (int t = threadIdx.x; t < topk; t += TK_THREADS) {
    const ScoreIdx b = best[t];
    out_items[(size_t)blockIdx.x * topk + t] = b.i;
    out_scores[(size_t)blockIdx.x * topk + t] = b.i >= 0 ? (float)b.s : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int t = 0; t < topk; ++t) cnt += best[t].i >= 0;
    if (out_count) out_count[blockIdx.x] = cnt;
  }
}

// ------------------------------------------------------------------------------------------
// NaiveBayes: per-class counts and feature sums (fp64), deterministic two-stage reduction.
// partial: [gridDim.x][n_class * (n_feat + 1)]  (slot n_feat = count)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nb_partial_kernel(const int* __restrict__ label, const float* __restrict__ x, long long n, int n_feat,
                  int n_class, double* __restrict__ partial) {
  extern __shared__ double acc[];  // [warps][n_class*(n_feat+1)]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = n_class * (n_feat + 1);
  double* my = acc + w * width;
  for (int o = lane; o < width; o += 32) my[o] = 0.0;
  __syncwarp();
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per;
  const long long r1 = r0 + per < n ? r0 + per : n;
  // each warp walks its strided rows; lanes serialise their updates in lane order so the
  // summation order is fixed (values are typically small integers, sums exact in fp64)
  for (long long base = r0 + (long long)w * 32; base < r1; base += 8 * 32) {
    const long long r = base + lane;
    int c = -1;
    if (r < r1) c = label[r];
    for (int src = 0; src < 32; ++src) {
      const int cc = __shfl_sync(0xffffffffu, c, src);
      if (cc < 0) continue;
      const long long rr = base + src;
      if (lane <= n_feat) {
        const double v = lane < n_feat ? (double)x[rr * n_feat + lane] : 1.0;
        my[cc * (n_feat + 1) + lane] += v;
      }
      for (int f = lane + 32; f < n_feat; f += 32) my[cc * (n_feat + 1) + f] += (double)x[rr * n_feat + f];
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < width; o += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += acc[q * width + o];
    partial[(size_t)blockIdx.x * width + o] = s;
  }
}

__global__ void nb_reduce_kernel(const double* __restrict__ partial, int nparts, int width,
                                 double* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= width) return;
  double s = 0.0;
  for (int q = 0; q < nparts; ++q) s += partial[(size_t)q * width + o];
  out[o] = s;
}

__global__ void __launch_bounds__(256)
nb_predict_kernel(const float* __restrict__ x, long long n, int n_feat, int n_class,
                  const double* __restrict__ pi, const double* __restrict__ theta, int* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int bestc = 0;
  double bests = -INFINITY;
  for (int c = 0; c < n_class; ++c) {
    double s = pi[c];
    // separate multiply and add (no FMA contraction) so the result is bit-identical to the
    // reference-order host arithmetic
    for (int j = 0; j < n_feat; ++j) s = __dadd_rn(s, __dmul_rn(theta[c * n_feat + j], (double)x[r * n_feat + j]));
    if (s > bests) { bests = s; bestc = c; }
  }
  out[r] = bestc;
}

}  // namespace pio
