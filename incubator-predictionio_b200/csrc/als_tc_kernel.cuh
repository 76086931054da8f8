// als_tc_kernel.cuh -- rank-64 ALS half-step with the Gramian on the 5th-gen tensor cores.
//
// Same mathematics as als_solve_kernel (als_kernels.cuh) for KP = 64, but  A = sum (s y)(s y)^T  is a
// split-precision SYRK on tcgen05:
//   every gathered, sqrt(c1)-scaled source row x (64 floats) is staged as  [hi | lo]  (128 floats),
//   hi = tf32_round(x), lo = x - hi;  8 ratings form one K-block in the MN-major
//   SWIZZLE_128B_BASE32B UMMA shared-memory layout (per 32 columns a 1 KB atom: one 128-byte row per
//   rating, its four 32-byte chunks XOR-swizzled by rating%4) -- the only MN-major layout kind::tf32
//   accepts (verified with tools/umma_probe.cu; the un-swizzled MN-major forms silently produce 0);
//   one  tcgen05.mma.cta_group::1.kind::tf32  with M = N = 128, K = 8 and the SAME tile as A and B
//   operand accumulates  D += [hi|lo]^T [hi|lo]  in TMEM; the four 64x64 quadrants of D sum to
//   (hi+lo)^T (hi+lo) -- fp32-class products (error ~2^-21 relative per term) at tensor-core rate.
// D is symmetric, so operand-order / transpose conventions cannot change the result.
//
// Warp roles (512 threads, 1 CTA per SM, persistent; rows are claimed from a global counter in
// degree-descending order):
//   warp 0      scheduler + MMA issuer (one elected lane): builds batch descriptors, issues the MMAs,
//               commits to mbarriers
//   warps 1-3   producers: gather source rows (LDG.128), scale, split hi/lo, STS into the stage ring,
//               accumulate the right-hand side b
//   warps 4-15  three teams of four warps (one warp per TMEM lane quarter): drain the four
//               accumulators of their batch (tcgen05.ld -> quadrant sum -> packed A slot in smem),
//               then each warp Cholesky-solves one of the four rows (chol_solve_warp<PACKED_IN>).
// A batch = up to four (row, rating-segment) pairs, one per TMEM accumulator (4 x 128 columns);
// rows longer than SEG ratings are accumulated segment by segment with fp32 adds in shared memory
// between segments (bounds the length of any single tensor-core accumulation chain).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "als_kernels.cuh"

namespace pio {
namespace tc {

constexpr int KP = 64;
constexpr int NCOL = 128;                        // [hi | lo]
constexpr int KB_RATINGS = 8;                    // K of one tf32 MMA
constexpr int KB_BYTES = KB_RATINGS * NCOL * 4;  // 4096
constexpr int STAGE_KB = 3;                      // one K-block per producer warp
constexpr int STAGE_RATINGS = STAGE_KB * KB_RATINGS;  // 24
constexpr int STAGE_BYTES = STAGE_KB * KB_BYTES;      // 12288
constexpr int NSTAGE = 4;
constexpr int SEG = 21 * STAGE_RATINGS;          // 504 ratings per accumulation segment
constexpr int NTEAM = 3;
constexpr int NSLOT = 4;
constexpr int NTHREADS = 512;
constexpr int H = 32;
constexpr int L21S = H + 4;
constexpr int OFF21 = H * (H + 1) / 2;           // 528
constexpr int OFF22 = OFF21 + H * L21S;          // 1680
constexpr int ASLOT = OFF22 + H * (H + 1) / 2;   // 2208 floats

struct BatchDesc {
  long long beg[NSLOT];
  int row[NSLOT];    // local row, -1 = empty slot
  int len[NSLOT];
  int first[NSLOT];
  int last[NSLOT];
  int exit;
  int pad[3];
};

struct Smem {
  alignas(1024) unsigned char stage[NSTAGE][STAGE_BYTES];
  alignas(16) float aslot[NTEAM][NSLOT][ASLOT];
  alignas(16) float bslot[NTEAM][NSLOT][KP];
  alignas(16) float bstage[NTEAM][NSLOT][KP];
  alignas(16) float bpart[STAGE_KB][KP];
  alignas(16) float colbuf[NTEAM * NSLOT][2 * KP];
  alignas(16) float dinv[NTEAM * NSLOT][KP];
  BatchDesc desc[NTEAM];
  alignas(8) unsigned long long full[NSTAGE];
  unsigned long long empty[NSTAGE];
  unsigned long long descfull[NTEAM];
  unsigned long long accfull[NTEAM];
  unsigned long long teamdone[NTEAM];
  unsigned long long bfull[NTEAM];
  unsigned long long tmemfree;
  unsigned int tmem_base;
};

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(s32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned int* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): MN-major,
// layout_type SWIZZLE_128B_BASE32B = 1 at [61,64); start address >> 4 at [0,14); LBO = stride between
// 32-column atoms (1024 B) >> 4 at [16,30); SBO = stride between groups of 4 ratings (512 B) >> 4 at [32,46);
// version 1 at [46,48).  Element (rating k, column mn) of a K-block sits at
//   (mn/32)*1024 + k*128 + ((((mn%32)/8) ^ (k%4)) * 32) + (mn%8)*4      (tools/umma_probe.cu: exact).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format F32 = 1 at [4,6),
// a/b format TF32 = 2 at [7,10)/[10,13), a/b major MN = 1 at bits 15/16, N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct TcParams {
  SolveParams sp;
  int* counter;  // next row to claim (starts at sp.row_begin)
  float* dbg;    // debug only (PIO_ALS_TC_DEBUG=1): per local row ASLOT + KP floats (A as drained, b), else null
};

template <bool IMPLICIT>
__global__ void __launch_bounds__(NTHREADS, 1) als_solve_tc_kernel(const TcParams tp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const SolveParams& p = tp.sp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&sm.full[i], STAGE_KB);
      mbar_init(&sm.empty[i], 1);
    }
    for (int t = 0; t < NTEAM; ++t) {
      mbar_init(&sm.descfull[t], 1);
      mbar_init(&sm.accfull[t], 1);
      mbar_init(&sm.teamdone[t], NSLOT);
      mbar_init(&sm.bfull[t], 1);
    }
    mbar_init(&sm.tmemfree, NSLOT);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ================= scheduler + MMA issuer =================
    if (lane == 0) {
      int l_row[NTEAM * NSLOT];
      long long l_pos[NTEAM * NSLOT], l_end[NTEAM * NSLOT];
      for (int i = 0; i < NTEAM * NSLOT; ++i) { l_row[i] = -1; l_pos[i] = 0; l_end[i] = 0; }
      bool exited[NTEAM] = {false, false, false};
      int nbatch_team[NTEAM] = {0, 0, 0};
      bool out_of_rows = false;
      uint32_t it = 0;       // global stage counter
      uint32_t nmma_batches = 0;
      // returns false when team t has nothing left
      auto build = [&](int t) -> bool {
        mbar_wait(&sm.teamdone[t], (nbatch_team[t] & 1) ^ 1);  // previous batch of this team fully consumed
        BatchDesc& d = sm.desc[t];
        bool any = false;
        for (int s = 0; s < NSLOT; ++s) {
          const int li = t * NSLOT + s;
          if (l_row[li] >= 0 && l_pos[li] >= l_end[li]) l_row[li] = -1;
          if (l_row[li] < 0 && !out_of_rows) {
            const int r = atomicAdd(tp.counter, 1);
            if (r < p.row_end) {
              l_row[li] = r;
              l_pos[li] = p.ptr[r];
              l_end[li] = p.ptr[r + 1];
            } else {
              out_of_rows = true;
            }
          }
          if (l_row[li] >= 0) {
            long long e = l_pos[li] + SEG;
            if (e > l_end[li]) e = l_end[li];
            d.row[s] = l_row[li];
            d.beg[s] = l_pos[li];
            d.len[s] = (int)(e - l_pos[li]);
            d.first[s] = l_pos[li] == p.ptr[l_row[li]];
            d.last[s] = e == l_end[li];
            l_pos[li] = e;
            any = true;
          } else {
            d.row[s] = -1;
            d.beg[s] = 0;
            d.len[s] = 0;
            d.first[s] = 0;
            d.last[s] = 0;
          }
        }
        d.exit = any ? 0 : 1;
        mbar_arrive(&sm.descfull[t]);
        ++nbatch_team[t];
        if (!any) exited[t] = true;
        return any;
      };
      // software pipeline: the descriptor of batch b+1 is published before the MMAs of batch b are issued,
      // unless b+1 belongs to the same team as b (then its teamdone can only complete after b's MMAs).
      int b = 0;
      bool have[2] = {false, false};
      int team_of[2] = {0, 0};
      int cursor = 0;
      auto next_team = [&]() -> int {
        for (int k = 0; k < NTEAM; ++k) {
          const int t = (cursor + k) % NTEAM;
          if (!exited[t]) return t;
        }
        return -1;
      };
      // returns true when publishing for slot `nxt` is settled (a batch was found or every team exited)
      auto try_publish = [&](int cur, int nxt, bool allow_cur) -> bool {
        while (true) {
          const int t = next_team();
          if (t < 0) return true;
          if (!allow_cur && have[cur] && t == team_of[cur]) return false;
          cursor = (t + 1) % NTEAM;
          team_of[nxt] = t;
          if (build(t)) { have[nxt] = true; return true; }
        }
      };
      try_publish(1, 0, true);
      while (true) {
        const int cur = b & 1, nxt = cur ^ 1;
        have[nxt] = false;
        const bool settled = try_publish(cur, nxt, false);
        if (have[cur]) {
          const int t = team_of[cur];
          const BatchDesc& d = sm.desc[t];
          mbar_wait(&sm.tmemfree, (nmma_batches & 1) ^ 1);  // accumulators of the previous batch drained
          tc_fence_after();
          for (int s = 0; s < NSLOT; ++s) {
            const int len = d.len[s];
            const int nst = (len + STAGE_RATINGS - 1) / STAGE_RATINGS;
            for (int q = 0; q < nst; ++q, ++it) {
              const int st = it % NSTAGE;
              mbar_wait(&sm.full[st], (it / NSTAGE) & 1);
              tc_fence_after();
              const int valid = len - q * STAGE_RATINGS;
              const int nkb = valid >= STAGE_RATINGS ? STAGE_KB : (valid + KB_RATINGS - 1) / KB_RATINGS;
              const uint32_t base = s32(&sm.stage[st][0]);
              for (int kb = 0; kb < nkb; ++kb) {
                const uint64_t dsc = make_desc(base + kb * KB_BYTES);
                umma_tf32(tmem + s * NCOL, dsc, dsc, IDESC, (q > 0 || kb > 0) ? 1u : 0u);
              }
              umma_commit(&sm.empty[st]);
            }
          }
          umma_commit(&sm.accfull[t]);
          ++nmma_batches;
        }
        if (!settled) try_publish(cur, nxt, true);
        if (!have[nxt]) break;
        ++b;
      }
    }
    __syncwarp();
  } else if (warp <= STAGE_KB) {
    // ================= producers =================
    const int pw = warp - 1;            // K-block inside the stage
    // lane -> (rating kr inside the K-block, source float4 groups sg_j): a quarter-warp holds 4 ratings
    // (kr%4 = 0..3) x 2 halves of one 32-byte chunk -> its 8 STS.128 hit 8 distinct 16-byte bank groups
    const int kr = ((lane >> 3) & 1) * 4 + (lane & 3);
    const int sgb = ((lane >> 2) & 1) + 2 * (lane >> 4);   // sg_j = sgb + 4 j
    bool exited[NTEAM] = {false, false, false};
    int nbatch_team[NTEAM] = {0, 0, 0};
    uint32_t it = 0;
    int cursor = 0;
    while (true) {
      int t = -1;
      for (int k = 0; k < NTEAM; ++k) {
        const int c = (cursor + k) % NTEAM;
        if (!exited[c]) { t = c; break; }
      }
      if (t < 0) break;
      cursor = (t + 1) % NTEAM;
      mbar_wait(&sm.descfull[t], nbatch_team[t] & 1);
      ++nbatch_team[t];
      const BatchDesc& d = sm.desc[t];
      if (d.exit) { exited[t] = true; continue; }
      for (int s = 0; s < NSLOT; ++s) {
        const int len = d.len[s];
        if (d.row[s] < 0) continue;
        const long long beg = d.beg[s];
        float bacc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) bacc[j][c] = 0.f;
        const int nst = (len + STAGE_RATINGS - 1) / STAGE_RATINGS;
        // prefetch of the first stage's index / rating
        int k0 = pw * KB_RATINGS + kr;
        int nidx = -1;
        float nval = 0.f;
        if (k0 < len) { nidx = __ldg(p.idx + beg + k0); nval = __ldg(p.val + beg + k0); }
        for (int q = 0; q < nst; ++q, ++it) {
          const int st = it % NSTAGE;
          const int cidx = nidx;
          const float cval = nval;
          // prefetch next stage's metadata
          const int kn = (q + 1) * STAGE_RATINGS + pw * KB_RATINGS + kr;
          nidx = -1;
          nval = 0.f;
          if (kn < len) { nidx = __ldg(p.idx + beg + kn); nval = __ldg(p.val + beg + kn); }
          float4 y[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cidx >= 0) y[j] = __ldg(reinterpret_cast<const float4*>(p.src + (size_t)cidx * KP) + sgb + 4 * j);
          }
          float wb, sc;
          if (IMPLICIT) {
            const float c1 = p.alpha * fabsf(cval);
            wb = cval > 0.f ? 1.f + c1 : 0.f;
            sc = sqrtf(c1);
          } else {
            wb = cval;
            sc = 1.f;
          }
          const int valid = len - q * STAGE_RATINGS;
          const bool kb_used = pw * KB_RATINGS < valid;
          mbar_wait(&sm.empty[st], ((it / NSTAGE) & 1) ^ 1);
          if (kb_used) {
            unsigned char* kbp = &sm.stage[st][0] + pw * KB_BYTES + kr * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int sg = sgb + 4 * j;
              const int off = (sg >> 3) * 1024 + ((((sg & 7) >> 1) ^ (kr & 3)) * 32) + (sg & 1) * 16;
              bacc[j][0] = fmaf(wb, y[j].x, bacc[j][0]);
              bacc[j][1] = fmaf(wb, y[j].y, bacc[j][1]);
              bacc[j][2] = fmaf(wb, y[j].z, bacc[j][2]);
              bacc[j][3] = fmaf(wb, y[j].w, bacc[j][3]);
              float4 x = y[j];
              if (IMPLICIT) { x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
              float4 hi, lo;
              hi.x = tf32_round(x.x); hi.y = tf32_round(x.y); hi.z = tf32_round(x.z); hi.w = tf32_round(x.w);
              lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
              *reinterpret_cast<float4*>(kbp + off) = hi;          // columns 4sg..4sg+3 of the hi half
              *reinterpret_cast<float4*>(kbp + 2048 + off) = lo;   // same columns of the lo half (mn + 64)
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.full[st]);
        }
        // b of this segment: reduce over the 8 ratings of the warp (lanes with equal sgb), then over warps
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float v = bacc[j][c];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            bacc[j][c] = v;
          }
        named_bar_sync(1, STAGE_KB * 32);  // bpart free (previous segment's sum was read)
        if (kr == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(&sm.bpart[pw][(sgb + 4 * j) * 4]) = make_float4(bacc[j][0], bacc[j][1], bacc[j][2], bacc[j][3]);
        }
        named_bar_sync(1, STAGE_KB * 32);
        if (pw == 0) {
          for (int c = lane; c < KP; c += 32) sm.bstage[t][s][c] = (sm.bpart[0][c] + sm.bpart[1][c]) + sm.bpart[2][c];
        }
      }
      if (pw == 0) {  // all right-hand sides of this batch are in bstage[t]
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bfull[t]);
      }
    }
  } else {
    // ================= teams: drain + solve =================
    const int t = (warp - 4) / 4;
    const int slot = (warp - 4) % 4;     // the row this warp solves
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int barid = 2 + t;
    int nb = 0;
    while (true) {
      mbar_wait(&sm.descfull[t], nb & 1);
      const BatchDesc& d = sm.desc[t];
      if (d.exit) break;
      mbar_wait(&sm.accfull[t], nb & 1);
      tc_fence_after();
      // ---- drain: row i of D lives in TMEM lane i; A = D[0:64,0:64] + D[0:64,64:128] + D[64:128,0:64] + D[64:128,64:128]
      const int drow = (q & 1) * 32 + lane;  // destination row 0..63 handled by this thread
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const bool mine = pass == 0 ? (q >= 2) : (q < 2);  // lo rows first (they initialise), then hi rows add
        if (mine) {
#pragma unroll 1
          for (int s = 0; s < NSLOT; ++s) {
            if (d.row[s] < 0) continue;
            float* as = &sm.aslot[t][s][0];
            const bool overwrite = (pass == 0) && d.first[s];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * NCOL);
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
              float a[16], b2[16];
              tmem_ld16(taddr + cb * 16, a);
              tmem_ld16(taddr + 64 + cb * 16, b2);
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const int col = cb * 16 + c;
                if (col <= drow) {
                  const float v = a[c] + b2[c];
                  int off;
                  if (drow < H) off = drow * (drow + 1) / 2 + col;
                  else if (col < H) off = OFF21 + (drow - H) * L21S + col;
                  else off = OFF22 + (drow - H) * (drow - H + 1) / 2 + (col - H);
                  as[off] = overwrite ? v : as[off] + v;
                }
              }
            }
          }
        }
        if (pass == 0) named_bar_sync(barid, 128);
      }
      // right-hand side of this segment
      mbar_wait(&sm.bfull[t], nb & 1);
      {
        const int tt = (warp - 4) % 4 * 32 + lane;  // 0..127 within the team
        for (int o = tt; o < NSLOT * KP; o += 128) {
          const int s = o / KP, c = o % KP;
          if (d.row[s] >= 0) sm.bslot[t][s][c] = d.first[s] ? sm.bstage[t][s][c] : sm.bslot[t][s][c] + sm.bstage[t][s][c];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmemfree);
      named_bar_sync(barid, 128);
      // ---- solve
      const int r = d.row[slot];
      if (r >= 0 && d.last[slot] && tp.dbg) {
        float* o = tp.dbg + (size_t)r * (ASLOT + KP);
        for (int c = lane; c < ASLOT; c += 32) o[c] = sm.aslot[t][slot][c];
        for (int c = lane; c < KP; c += 32) o[ASLOT + c] = sm.bslot[t][slot][c];
        __syncwarp();
      }
      if (r >= 0 && d.last[slot]) {
        const float ridge = p.lambda * p.nreg[r];
        chol_solve_warp<KP, 8, 72, IMPLICIT, true>(&sm.aslot[t][slot][0], &sm.bslot[t][slot][0], p.yty, ridge, p.k,
                                                   &sm.colbuf[t * NSLOT + slot][0], &sm.dinv[t * NSLOT + slot][0],
                                                   p.dst + (size_t)(p.dst_row_offset + r) * KP, p.fail);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.teamdone[t]);
      ++nb;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace pio
