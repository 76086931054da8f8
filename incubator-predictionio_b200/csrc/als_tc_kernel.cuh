// als_tc_kernel.cuh -- rank 33..64 ALS half-step with the Gramian on the 5th-gen tensor cores (tcgen05 + TMEM).
//
// Same mathematics as als_solve_kernel (als_kernels.cuh) for KP = 64, but  A = sum (s y)(s y)^T  is a
// split-precision SYRK on tcgen05:
//   every gathered, sqrt(c1)-scaled source row x (64 floats) is staged as  [hi | lo]  (128 floats),
//   hi = the top 10 mantissa bits of x, lo = x - hi (exact);  8 ratings form one K-block in the MN-major
//   SWIZZLE_128B_BASE32B UMMA shared-memory layout (per 32 columns a 1 KB atom: one 128-byte row per
//   rating, its four 32-byte chunks XOR-swizzled by rating%4) -- the only MN-major layout kind::tf32
//   accepts (verified with tools/umma_probe.cu; the un-swizzled MN-major forms silently produce 0);
//   per K-block ONE  tcgen05.mma.cta_group::1.kind::tf32  with A = all 128 staged columns (M = 128) and
//   B = the 64 hi columns (N = 64, K = 8) accumulates  D = [hi^T hi ; lo^T hi]  into a 128-lane x 64-column
//   TMEM accumulator (row r in lane r); the drain forms  A = HH + LH + LH^T  (the dropped lo^T lo term is
//   ~2^-20 relative).  TMEM (512 columns) holds two batches of four rows: the MMAs of batch b+1 overlap the
//   drain of batch b.
//
// Warp roles (512 threads, 1 CTA per SM, persistent; degree-sorted rows are dealt statically: groups of
// NTEAM*NSLOT consecutive rows go to the CTAs in snake order):
//   warp 0              scheduler + MMA issuer, warp-parallel: lane t*4+s keeps row lane (team t, slot s) in
//                       registers, builds the flat stage list of a batch with shuffles, publishes it, issues the
//                       batch's MMAs from one lane and commits to mbarriers
//   warps 1..NGATHER    gather warps: gw owns every NGATHER-th stage; source indices two stages ahead in
//                       registers, 12 LDGSTS per 24-rating stage into the raw ring (rows padded to 288 bytes),
//                       completion by cp.async.mbarrier.arrive.noinc
//   next NCONV warps    converters: cw owns the stages whose number inside their segment is cw mod NCONV: LDS.128
//                       of the raw rows, scale, hi/lo split, 24 conflict-free STS.128 into the UMMA stage ring,
//                       fence.proxy.async, arrive on full[stage]; right-hand side b reduced at segment ends
//   last 4*NTEAM warps  teams of four warps (one per TMEM lane quarter): drain the four accumulators of their
//                       batch (tcgen05.ld: two warps store HH, two add LH and then its transpose), then each
//                       warp Cholesky-solves one of the four rows (chol_solve_warp<PACKED_IN>), or - split mode -
//                       stores the normal equations for als_solve_packed_kernel.
// A batch = up to four (row, rating-segment) pairs, one per TMEM accumulator (64 columns each); rows longer
// than SEG ratings are accumulated segment by segment with fp32 adds in shared memory between segments (bounds
// the length of any single tensor-core accumulation chain).  Every sum has a fixed order that depends on the
// row alone, so results do not depend on the grid or on how rows are sharded over GPUs.
// Every blocking mbarrier wait has a watchdog (mbar_wait): a protocol error traps with a diagnostic.
// This header is included once per role partition (pio_als.cu): TC_NS names the namespace, TC_NTEAM / TC_NCONV /
// TC_NGATHER / TC_NSTAGE / TC_NRAW fix how the sixteen warps and the shared memory of the CTA are divided.
#ifndef TC_NS
#error "define TC_NS, TC_NTEAM, TC_NCONV, TC_NGATHER, TC_NSTAGE, TC_NRAW before including als_tc_kernel.cuh"
#endif
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "als_kernels.cuh"

#ifndef PIO_TC_TIMING
#define PIO_TC_TIMING 0
#endif

namespace pio {
namespace TC_NS {

constexpr int KP = 64;
constexpr int NCOL = 128;                        // staged columns per rating: [hi | lo]
constexpr int ACOL = 64;                         // TMEM columns of one accumulator
constexpr int KB_RATINGS = 8;                    // K of one tf32 MMA
constexpr int KB_BYTES = KB_RATINGS * NCOL * 4;  // 4096
constexpr int STAGE_KB = 3;                      // one K-block per producer warp
constexpr int STAGE_RATINGS = STAGE_KB * KB_RATINGS;  // 24
constexpr int STAGE_BYTES = STAGE_KB * KB_BYTES;      // 12288
constexpr int NSTAGE = TC_NSTAGE;                 // UMMA stage ring
constexpr int NRAW = TC_NRAW;                          // raw ring: stages of gathered rows in flight
constexpr int RAW_ROW = 288;                     // bytes per gathered row in the raw ring (256 + 32 pad: conflict-free LDS.128)
constexpr int NGATHER = TC_NGATHER;                       // gather warps (each owns every NGATHER-th stage)
constexpr int NCONV = TC_NCONV;                         // converter warps (each owns every NCONV-th stage)
constexpr int CONV0 = 1 + NGATHER;               // first converter warp
constexpr int TEAM0 = CONV0 + NCONV;             // first drain+solve warp (a multiple of 4: warp%4 = TMEM lane quarter = slot)
static_assert(TEAM0 % 4 == 0, "team warps must start at a multiple of four");
constexpr int SEG = 21 * STAGE_RATINGS;          // 504 ratings per accumulation segment
constexpr int NTEAM = TC_NTEAM;                   // drain + solve teams of four warps
constexpr int NSLOT = 4;
constexpr int NTHREADS = 512;
static_assert((TEAM0 + 4 * NTEAM) * 32 == NTHREADS, "the role partition must fill the sixteen warps of the CTA");
constexpr int H = 32;
constexpr int L21S = H + 4;
constexpr int OFF21 = H * (H + 1) / 2;           // 528
constexpr int OFF22 = OFF21 + H * L21S;          // 1680
constexpr int ASLOT = OFF22 + H * (H + 1) / 2;   // 2208 floats

// one UMMA stage (<= 24 ratings of one segment), pre-expanded by the scheduler so producers need no bookkeeping
struct StageEnt {
  long long beg;   // offset of the stage's first rating in idx/val
  int valid;       // ratings in this stage (1..24)
  int info;        // bits 0-1 slot, bit 2 last stage of its segment, bit 3 last stage of the batch, bit 4 first stage of its segment
};
constexpr int MAX_STAGES = NSLOT * (SEG / STAGE_RATINGS);   // 84

struct BatchDesc {
  int row[NSLOT];    // local row, -1 = empty slot
  int first[NSLOT];  // segment starts its row (overwrite the A slot instead of accumulating)
  int last[NSLOT];   // segment ends its row (solve after draining)
  int exit;
  int seq;           // global batch number (selects the TMEM accumulator set)
  int nstages;
  int pad;
  StageEnt st[MAX_STAGES];
};

struct Smem {
  alignas(1024) unsigned char stage[NSTAGE][STAGE_BYTES];
  alignas(16) float aslot[NTEAM][NSLOT][ASLOT];
  alignas(16) float bslot[NTEAM][NSLOT][KP];
  alignas(16) float bstage[NTEAM][2][NSLOT][KP];
  alignas(16) float bpart[NCONV][KP];
  alignas(16) unsigned char raw[NRAW][STAGE_RATINGS * RAW_ROW];   // gathered source rows (cp.async.bulk destinations)
  float rawval[NRAW][32];                                          // their ratings
  alignas(16) float colbuf[NTEAM * NSLOT][2 * KP];
  alignas(16) float dinv[NTEAM * NSLOT][KP];
  BatchDesc desc[NTEAM][2];   // double-buffered per team: batch n+1 is produced while the team still solves batch n
  alignas(8) unsigned long long full[NSTAGE];
  unsigned long long empty[NSTAGE];
  unsigned long long descfull[NTEAM][2];
  unsigned long long accfull[NTEAM][2];
  unsigned long long teamdone[NTEAM][2];
  unsigned long long bfull[NTEAM][2];
  unsigned long long tmemfree[2];
  unsigned long long rawfull[NRAW];    // gather warp -> converters (transaction-count barrier)
  unsigned long long rawempty[NRAW];   // converters -> gather warp
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
__device__ __forceinline__ bool mbar_try(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(s32(bar)), "r"(parity), "r"(200000u)   // suspend-time hint (ns): the waiting warp sleeps instead of polling
      : "memory");
  return ok != 0;
}
// Blocking wait with a watchdog: a protocol error must end the kernel with a diagnostic, never hang the GPU.
__device__ __noinline__ void mbar_timeout(unsigned long long* bar, uint32_t parity) {
  extern __shared__ __align__(1024) unsigned char tc_smem_base_[];
  if ((threadIdx.x & 31) == 0) printf("pio tc kernel: mbarrier wait timed out: cta %d warp %d lane %d barrier word %d (full4 empty4 descfull6 accfull6 teamdone6 bfull6 tmemfree2 rawfull6 rawempty6) parity %u\n", (int)blockIdx.x,
         (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31), (int)(s32(bar) - s32(tc_smem_base_) - (uint32_t)offsetof(Smem, full)) / 8, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 16000000000ll) mbar_timeout(bar, parity);   // ~8 s (a half-step takes tens of ms; tools like compute-sanitizer slow the kernel ~100x)
  }
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"   // test_wait never suspends (try_wait may)
      "selp.u32 %0, 1, 0, P1;\n"
      "}"
      : "=r"(ok)
      : "r"(s32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
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
// the mbarrier receives one arrival from this thread once all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
// one contiguous global -> shared copy (UBLKCP); completion is signalled on the mbarrier's transaction count
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s32(smem_dst)), "l"(gsrc));
}
// allow at most `pend` of the most recent commit groups to stay in flight (exact up to 8, else conservative)
__device__ __forceinline__ void cp_async_wait_dyn(uint32_t pend) {
  switch (pend) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    case 3: cp_async_wait<3>(); break;
    case 4: cp_async_wait<4>(); break;
    case 5: cp_async_wait<5>(); break;
    case 6: cp_async_wait<6>(); break;
    case 7: cp_async_wait<7>(); break;
    default: cp_async_wait<8>(); break;
  }
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// hi part of the split: the top 10 mantissa bits by truncation (one LOP3 on the full-rate integer pipe; cvt.rna.tf32
// goes through the quarter-rate conversion pipe, 48 per lane and stage).  x - hi is exact and has at most 13 significant
// bits, so the split loses nothing; the tensor core's own truncation of lo costs 2^-21 relative, the dropped lo*lo
// term 2^-20.
__device__ __forceinline__ float tf32_round(float x) {
  return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
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
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // M = 128 ([hi|lo] columns), N = 64 (hi columns)

static_assert(sizeof(Smem) + 1024 <= 232448, "shared memory budget (227 KB) exceeded");

struct TcParams {
  SolveParams sp;
  long long* timing;  // debug only (PIO_ALS_TC_TIMING=1): [grid][16 warps][8] cycle counters, else null
  float* dbg;    // debug only (PIO_ALS_TC_DEBUG=1): per local row ASLOT + KP floats (A as drained, b), else null
  float* out;    // split mode: per local row ASLOT + KP floats (packed lower triangle of the Gramian, then b); the rows
                 // are solved afterwards by als_solve_packed_kernel with every warp of the SM; null = solve in this kernel
  int out_row0;  // local row stored at out[0] (the buffer covers one tile of rows)
};

template <bool IMPLICIT>
__global__ void __launch_bounds__(NTHREADS, 1) als_solve_tc_kernel(const TcParams tp) {
  // direct cast (no integer round-trip) so the compiler keeps the shared address space: LDS/STS, not generic LD/ST
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  if (threadIdx.x == 0 && (s32(smem_raw) & 1023u) != 0u) __trap();   // UMMA swizzle atoms need 1 KB alignment
  const SolveParams& p = tp.sp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    for (int i = 0; i < NRAW; ++i) {
      mbar_init(&sm.rawfull[i], 32);   // every lane of the gather warp: cp.async.mbarrier.arrive.noinc
      mbar_init(&sm.rawempty[i], 1);
    }
    for (int t = 0; t < NTEAM; ++t)
      for (int bf = 0; bf < 2; ++bf) {
        mbar_init(&sm.descfull[t][bf], 1);
        mbar_init(&sm.accfull[t][bf], 1);
        mbar_init(&sm.teamdone[t][bf], NSLOT);
        mbar_init(&sm.bfull[t][bf], 1);
      }
    mbar_init(&sm.tmemfree[0], NSLOT);
    mbar_init(&sm.tmemfree[1], NSLOT);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // per-role cycle counters are compiled in only with -DPIO_TC_TIMING=1 (tools/tc_timing.py); the default build carries no
  // instrumentation in the hot loops
  const bool timing = PIO_TC_TIMING && tp.timing != nullptr;
  long long t0_ = 0;
#define T_BEGIN() do { if (timing) t0_ = clock64(); } while (0)
#define T_END(i) do { if (timing) tacc[i] += clock64() - t0_; } while (0)
  const long long tstart_ = timing ? clock64() : 0;

  // register rebalancing: scheduler/producer warpgroup gives registers to the three drain+solve warpgroups
  // (setmaxnreg rebalancing was tried and removed: shrinking the producer warpgroup forced its state into local memory)

  if (warp == 0) {
    // ================= scheduler + MMA issuer =================
    // Warp-parallel: lane li = t*4+s (< 12) keeps the state of row lane (team t, slot s) in registers; every control
    // decision is warp-uniform.  Static row assignment (no atomics): rows are degree-sorted; groups of 12 consecutive
    // rows are dealt to the CTAs in snake order, lane li of this CTA takes row  g*12 + li  of its k-th group g.  The row
    // pointers of a lane's NEXT row are prefetched while the current one is processed.
    const int per_cta = NTEAM * NSLOT;
    const int nrows = p.row_end - p.row_begin;
    const int ngroups = (nrows + per_cta - 1) / per_cta;
    const int G = (int)gridDim.x;
    auto row_of = [&](int k) -> int {
      const long long g = (long long)k * G + ((k & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x);
      if (lane >= per_cta || g >= ngroups) return -1;
      const long long r = (long long)p.row_begin + g * per_cta + lane;
      return r < p.row_end ? (int)r : -1;
    };
    int my_row = -1, my_k = 0;
    long long my_pos = 0, my_end = 0;
    bool my_first = false;
    int nx_row = row_of(0);
    long long nx_beg = 0, nx_end = 0;
    if (nx_row >= 0) { nx_beg = p.ptr[nx_row]; nx_end = p.ptr[nx_row + 1]; }
    uint32_t exited = 0;
    uint32_t nbp = 0;                // batches published per team, 8 bits each (only the low two bits are ever used)
    uint32_t it = 0;                 // global stage counter
    uint32_t nmma_batches = 0;
    uint32_t nbuilt = 0;             // batches with work, in publish order == MMA order
    auto nb_of = [&](int t) -> int { return (int)((nbp >> (8 * t)) & 3u); };
    // batch n of team t lives in buffer n&1; the buffer is free once batch n-2 (its previous user) is done
    auto team_free = [&](int t) -> bool {
      const int n = nb_of(t);
      return mbar_test(&sm.teamdone[t][n & 1], ((n >> 1) & 1) ^ 1);
    };
    // publish the next batch of team t; returns false when the team has nothing left (an exit descriptor is published)
    auto build = [&](int t) -> bool {
      const int nbt = nb_of(t);
      const int bf = nbt & 1;
      T_BEGIN();
      mbar_wait(&sm.teamdone[t][bf], ((nbt >> 1) & 1) ^ 1);
      T_END(0);
      T_BEGIN();
      BatchDesc& d = sm.desc[t][bf];
      const bool mine = lane < per_cta && (lane >> 2) == t;
      int len = 0;
      long long beg = 0;
      if (mine) {
        if (my_row >= 0 && my_pos >= my_end) my_row = -1;
        if (my_row < 0 && nx_row >= 0) {
          my_row = nx_row;
          my_pos = nx_beg;
          my_end = nx_end;
          my_first = true;
          ++my_k;
          nx_row = row_of(my_k);   // prefetch the pointers of the row after this one
          if (nx_row >= 0) { nx_beg = __ldg(p.ptr + nx_row); nx_end = __ldg(p.ptr + nx_row + 1); }
        }
        const int s = lane & 3;
        if (my_row >= 0) {
          long long e = my_pos + SEG;
          if (e > my_end) e = my_end;
          len = (int)(e - my_pos);
          beg = my_pos;
          d.row[s] = my_row;
          d.first[s] = my_first ? 1 : 0;
          d.last[s] = e == my_end;
          my_first = false;
          my_pos = e;
        } else {
          d.row[s] = -1;
          d.first[s] = 0;
          d.last[s] = 0;
        }
      }
      const int nst = (len + STAGE_RATINGS - 1) / STAGE_RATINGS;
      int n_[NSLOT], l_[NSLOT];
      long long b_[NSLOT];
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        n_[s] = __shfl_sync(0xffffffffu, nst, t * NSLOT + s);
        l_[s] = __shfl_sync(0xffffffffu, len, t * NSLOT + s);
        b_[s] = __shfl_sync(0xffffffffu, beg, t * NSLOT + s);
      }
      const int ns = n_[0] + n_[1] + n_[2] + n_[3];
      for (int e = lane; e < ns; e += 32) {   // the flat stage list, one entry per lane and pass
        int s = 0, q = e;
#pragma unroll
        for (int k = 0; k < NSLOT - 1; ++k)
          if (s == k && q >= n_[k]) { q -= n_[k]; s = k + 1; }
        const int ls = s == 0 ? l_[0] : (s == 1 ? l_[1] : (s == 2 ? l_[2] : l_[3]));
        const int nn = s == 0 ? n_[0] : (s == 1 ? n_[1] : (s == 2 ? n_[2] : n_[3]));
        const long long bs = s == 0 ? b_[0] : (s == 1 ? b_[1] : (s == 2 ? b_[2] : b_[3]));
        StageEnt se;
        se.beg = bs + (long long)q * STAGE_RATINGS;
        const int v = ls - q * STAGE_RATINGS;
        se.valid = v < STAGE_RATINGS ? v : STAGE_RATINGS;
        se.info = s | (q == nn - 1 ? 4 : 0) | (q == 0 ? 16 : 0) | (e == ns - 1 ? 8 : 0);
        d.st[e] = se;
      }
      const bool any = ns > 0;
      if (lane == 0) {
        d.nstages = ns;
        d.exit = any ? 0 : 1;
        d.seq = any ? (int)nbuilt : 0;
      }
      if (any) ++nbuilt;
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.descfull[t][bf]);
      nbp = (nbp & ~(0xffu << (8 * t))) | ((((nbp >> (8 * t)) + 1u) & 3u) << (8 * t));
      if (!any) exited |= 1u << t;
      T_END(3);
      return any;
    };
    // software pipeline: the descriptor of batch b+1 is published before the MMAs of batch b are issued,
    // unless b+1 belongs to the same team as b (then its teamdone can only complete after b's MMAs).
    int b = 0;
    bool have[2] = {false, false};
    int team_of[2] = {0, 0};
    int buf_of[2] = {0, 0};
    int cursor = 0;
    auto next_team = [&]() -> int {
      for (int k = 0; k < NTEAM; ++k) {
        const int t = (cursor + k) % NTEAM;
        if (!((exited >> t) & 1u)) return t;
      }
      return -1;
    };
    // returns true when publishing for slot `nxt` is settled (a batch was found or every team exited)
    auto try_publish = [&](int cur, int nxt, bool allow_cur) -> bool {
      while (true) {
        const int t = next_team();
        if (t < 0) return true;
        if (!allow_cur && have[cur] && !team_free(t)) return false;  // do not stall the MMAs of `cur`
        cursor = (t + 1) % NTEAM;
        team_of[nxt] = t;
        buf_of[nxt] = nb_of(t) & 1;
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
        const BatchDesc& d = sm.desc[t][buf_of[cur]];
        const uint32_t set = nmma_batches & 1;                         // TMEM accumulator set of this batch
        T_BEGIN();
        mbar_wait(&sm.tmemfree[set], ((nmma_batches >> 1) & 1) ^ 1);   // batch b-2 (same set) has been drained
        T_END(1);
        tc_fence_after();
        const int nstg = d.nstages;
        for (int i = 0; i < nstg; ++i, ++it) {
          const int st = it % NSTAGE;
          const int valid = d.st[i].valid;
          const int info = d.st[i].info;
          T_BEGIN();
          mbar_wait(&sm.full[st], (it / NSTAGE) & 1);
          T_END(2);
          tc_fence_after();
          T_BEGIN();
          if (lane == 0) {
            const int nkb = (valid + KB_RATINGS - 1) / KB_RATINGS;
            const uint32_t base = s32(&sm.stage[st][0]);
            const uint32_t acc = tmem + set * (NSLOT * ACOL) + (info & 3) * ACOL;
            for (int kb = 0; kb < nkb; ++kb) {
              // A = all 128 staged columns [hi | lo] (M = 128), B = the 64 hi columns (N = 64):
              // D rows 0..63 = hi^T hi, rows 64..127 = lo^T hi; the drain adds lo^T hi and its transpose to hi^T hi
              const uint64_t dh = make_desc(base + kb * KB_BYTES);
              umma_tf32(acc, dh, dh, IDESC, ((info & 16) && kb == 0) ? 0u : 1u);
            }
            umma_commit(&sm.empty[st]);
          }
          __syncwarp();
          T_END(4);
        }
        if (lane == 0) umma_commit(&sm.accfull[t][buf_of[cur]]);
        __syncwarp();
        ++nmma_batches;
      }
      if (!settled) try_publish(cur, nxt, true);
      if (!have[nxt]) break;
      ++b;
    }
    __syncwarp();
  } else if (warp < TEAM0) {
    // ================= producers =================
    // warp 1 = gather warp: walks the scheduler's stage lists and, per stage, starts one 256-byte cp.async.bulk
    //          copy per rating into the raw ring (completion counted on rawfull[]); source indices are loaded
    //          two stages ahead into a register shift queue.
    // converters: converter cw owns the stages whose number inside their segment is cw mod NCONV: it waits for the
    //          raw rows, scales, splits hi/lo and writes all three K-blocks of the UMMA tile (three independent
    //          pieces of work per dependency chain), then hands the stage to the MMA issuer.
    // cursor over the published stage lists (scalars only)
    uint32_t c_ex = 0, c_par = 0;
    int c_cur = 0, c_t = 0, c_b = 0, c_i = -1, c_n = 0;
    bool c_done = false;
    auto next_batch = [&](bool blocking) -> bool {   // -> first stage of the next batch; false if finished / not published yet
      while (true) {
        int nt = -1;
#pragma unroll
        for (int k = 0; k < NTEAM; ++k) {
          const int cand = (c_cur + k) % NTEAM;
          if (nt < 0 && !((c_ex >> cand) & 1u)) nt = cand;
        }
        if (nt < 0) { c_done = true; return false; }
        const uint32_t cnt = (c_par >> (2 * nt)) & 3u;   // batches of team nt consumed so far, mod 4
        const int nbf = (int)(cnt & 1u);
        const uint32_t pb = (cnt >> 1) & 1u;
        if (blocking) mbar_wait(&sm.descfull[nt][nbf], pb);
        else if (!mbar_test(&sm.descfull[nt][nbf], pb)) return false;
        c_par = (c_par & ~(3u << (2 * nt))) | (((cnt + 1u) & 3u) << (2 * nt));
        c_cur = (nt + 1) % NTEAM;
        if (sm.desc[nt][nbf].exit) { c_ex |= 1u << nt; continue; }
        c_t = nt;
        c_b = nbf;
        c_n = sm.desc[nt][nbf].nstages;
        c_i = 0;
        return true;
      }
    };
    auto advance = [&](bool blocking) -> bool {
      if (c_done) return false;
      if (c_i >= 0 && c_i + 1 < c_n) { ++c_i; return true; }
      return next_batch(blocking);
    };
    if (warp < CONV0) {
      // ---------------- gather warps: warp gw issues every stage with (stage number % NGATHER == gw) ----------------
      // queue of this warp's next two stages (global stage number, beg, valid, source index) in registers; q0 is issued next
      const int gw = warp - 1;
      long long qbeg0 = 0, qbeg1 = 0;
      int qval0 = -1, qval1 = -1;      // valid count; -1 = empty queue slot
      int qidx0 = -1, qidx1 = -1;
      uint32_t qn0 = 0, qn1 = 0;
      uint32_t n_seen = 0;             // stages walked so far (all warps' stages)
      bool more = true;
      // pull this warp's next stage and start its index load; the look-ahead never blocks on a descriptor (the scheduler
      // may hold the next batch back until the MMAs of the current one - which need the stage still queued here - are issued)
      auto fetch = [&](uint32_t& qn, long long& qb, int& qv, int& qi, bool blocking) {
        qv = -1;
        while (more) {
          T_BEGIN();
          const bool ok = advance(blocking);
          T_END(0);
          if (!ok) {
            if (c_done) more = false;
            return;
          }
          const uint32_t n = n_seen++;
          if ((n % NGATHER) != (uint32_t)gw) continue;
          const StageEnt se = sm.desc[c_t][c_b].st[c_i];
          qn = n;
          qb = se.beg;
          qv = se.valid;
          qi = -1;
          if (lane < se.valid) qi = __ldg(p.idx + se.beg + lane);
          return;
        }
      };
      while (true) {
        if (qval0 < 0) {
          if (qval1 >= 0) { qn0 = qn1; qbeg0 = qbeg1; qval0 = qval1; qidx0 = qidx1; qval1 = -1; }
          else fetch(qn0, qbeg0, qval0, qidx0, true);    // nothing pending in this warp: safe to block
          if (qval0 < 0) break;                          // finished
        }
        if (qval1 < 0) fetch(qn1, qbeg1, qval1, qidx1, false);
        const int rs = qn0 % NRAW;
        T_BEGIN();
        mbar_wait(&sm.rawempty[rs], ((qn0 / NRAW) & 1) ^ 1);   // the converter is done with this raw slot
        T_END(1);
        T_BEGIN();
        // lane = (row parity, 16-byte chunk): one LDGSTS moves two whole 256-byte rows per warp instruction
#pragma unroll
        for (int i = 0; i < STAGE_RATINGS / 2; ++i) {
          const int r = 2 * i + (lane >> 4);
          const int si = __shfl_sync(0xffffffffu, qidx0, r);
          if (2 * i < qval0 && si >= 0)
            cp_async16(&sm.raw[rs][r * RAW_ROW + (lane & 15) * 16], p.src + (size_t)si * KP + (lane & 15) * 4);
        }
        if (lane < qval0) cp_async4(&sm.rawval[rs][lane], p.val + qbeg0 + lane);
        cp_async_mbar_arrive_noinc(&sm.rawfull[rs]);
        T_END(2);
        qval0 = -1;
      }
    } else {
      // ---------------- converters ----------------
      const int cw = warp - CONV0;
      // lane -> (rating kr inside a K-block, source float4 groups sg_j = sgb + 4 j): a quarter-warp holds 4 ratings
      // (kr%4 = 0..3) x 2 halves of one 32-byte chunk -> its 8 STS.128 hit 8 distinct 16-byte bank groups
      const int kr = ((lane >> 3) & 1) * 4 + (lane & 3);
      const int sgb = ((lane >> 2) & 1) + 2 * (lane >> 4);
      uint32_t it = 0;   // global stage counter (UMMA ring + raw ring)
      uint32_t qs = 0;   // stage number inside the current segment: decides the owner, so that the grouping of a row's
                         // right-hand-side partial sums depends on the row alone (not on what else this CTA processes)
      float bacc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) bacc[j][c] = 0.f;
      while (true) {
        T_BEGIN();
        const bool ok = advance(true);
        T_END(0);
        if (!ok) break;
        const StageEnt se = sm.desc[c_t][c_b].st[c_i];
        if (se.info & 16) qs = 0;
        const bool mine = (qs % NCONV) == (uint32_t)cw;
        ++qs;
        if (mine) {
          const int rs = it % NRAW;
          const int st = it % NSTAGE;
          T_BEGIN();
          mbar_wait(&sm.rawfull[rs], (it / NRAW) & 1);           // gathered rows have landed
          T_END(1);
          T_BEGIN();
          // all shared-memory loads of the stage first (12 independent LDS.128 + 3 ratings), then arithmetic + stores
          float4 y[STAGE_KB][4];
          float cv[STAGE_KB];
#pragma unroll
          for (int kb = 0; kb < STAGE_KB; ++kb) {
            const int r = kb * KB_RATINGS + kr;                  // rating inside the stage
            const bool live = r < se.valid;
            cv[kb] = live ? sm.rawval[rs][r] : 0.f;
            const unsigned char* rrow = &sm.raw[rs][r * RAW_ROW];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              y[kb][j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (live) y[kb][j] = *reinterpret_cast<const float4*>(rrow + (sgb + 4 * j) * 16);
            }
          }
          // the raw slot is in registers now: hand it back to the gather warps before converting
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.rawempty[rs]);
          T_END(3);
          T_BEGIN();
          mbar_wait(&sm.empty[st], ((it / NSTAGE) & 1) ^ 1);     // UMMA stage is free
          T_END(2);
          T_BEGIN();
#pragma unroll
          for (int kb = 0; kb < STAGE_KB; ++kb) {
            if (kb * KB_RATINGS < se.valid) {
              const float cval = cv[kb];
              float wb, sc;
              if (IMPLICIT) {
                const float c1 = p.alpha * fabsf(cval);
                wb = cval > 0.f ? 1.f + c1 : 0.f;
                sc = sqrtf(c1);
              } else {
                wb = cval;
                sc = 1.f;
              }
              unsigned char* kbp = &sm.stage[st][0] + kb * KB_BYTES + kr * 128;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int sg = sgb + 4 * j;
                const int off = (sg >> 3) * 1024 + ((((sg & 7) >> 1) ^ (kr & 3)) * 32) + (sg & 1) * 16;
                const float4 yy = y[kb][j];
                bacc[j][0] = fmaf(wb, yy.x, bacc[j][0]);
                bacc[j][1] = fmaf(wb, yy.y, bacc[j][1]);
                bacc[j][2] = fmaf(wb, yy.z, bacc[j][2]);
                bacc[j][3] = fmaf(wb, yy.w, bacc[j][3]);
                float4 x = yy;
                if (IMPLICIT) { x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
                float4 hi, lo;
                hi.x = tf32_round(x.x); hi.y = tf32_round(x.y); hi.z = tf32_round(x.z); hi.w = tf32_round(x.w);
                lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
                *reinterpret_cast<float4*>(kbp + off) = hi;          // columns 4sg..4sg+3 of the hi half
                *reinterpret_cast<float4*>(kbp + 2048 + off) = lo;   // same columns of the lo half (mn + 64)
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.full[st]);       // -> MMA issuer
          T_END(4);
        }
        ++it;
        if (se.info & 4) {
          // end of a segment: every converter reduces its share of b over the 8 rating lanes, then the shares are added
          const int sslot = se.info & 3;
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
          named_bar_sync(1, NCONV * 32);  // bpart free (previous segment's sum was read)
          if (kr == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(&sm.bpart[cw][(sgb + 4 * j) * 4]) = make_float4(bacc[j][0], bacc[j][1], bacc[j][2], bacc[j][3]);
          }
          named_bar_sync(1, NCONV * 32);
          if (cw == 0) {
            for (int c = lane; c < KP; c += 32) {
              float acc = sm.bpart[0][c];
#pragma unroll
              for (int w = 1; w < NCONV; ++w) acc += sm.bpart[w][c];   // fixed order
              sm.bstage[c_t][c_b][sslot][c] = acc;
            }
            if (se.info & 8) {   // every right-hand side of this batch is in bstage[t][buf]
              __syncwarp();
              if (lane == 0) mbar_arrive(&sm.bfull[c_t][c_b]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) bacc[j][c] = 0.f;
        }
      }
    }
  } else {
    // ================= teams: drain + solve =================
    const int t = (warp - TEAM0) / 4;
    const int slot = (warp - TEAM0) % 4;     // the row this warp solves
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int barid = 2 + t;
    int nb = 0;
    while (true) {
      const int bf = nb & 1;
      const uint32_t ph = (nb >> 1) & 1;
      T_BEGIN();
      mbar_wait(&sm.descfull[t][bf], ph);
      T_END(0);
      const BatchDesc& d = sm.desc[t][bf];
      if (d.exit) break;
      T_BEGIN();
      mbar_wait(&sm.accfull[t][bf], ph);
      T_END(1);
      tc_fence_after();
      // every warp of the team must have finished solving the previous batch before any A slot is overwritten
      named_bar_sync(barid, 128);
      T_BEGIN();
      // ---- drain: the accumulator is 128 lanes x 64 columns: lane r < 64 = row r of hi^T hi, lane 64 + r = row r of
      // lo^T hi (LH).  A = HH + LH + LH^T: warps q = 0,1 store/accumulate their HH rows, then warps q = 2,3 add the lower
      // part of their LH rows, then (after a team barrier) its transpose - a fixed order, so the sums are deterministic.
      const uint32_t set = (uint32_t)d.seq & 1u;
      const int drow = (q & 1) * 32 + lane;
      auto load_slot = [&](int s, float* v) {
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + set * (NSLOT * ACOL) + (uint32_t)(s * ACOL);
        tmem_ld16_nowait(taddr, &v[0]);
        tmem_ld16_nowait(taddr + 16, &v[16]);
        tmem_ld16_nowait(taddr + 32, &v[32]);
        tmem_ld16_nowait(taddr + 48, &v[48]);
        tmem_ld_wait();
      };
      // row `drow` of a 64x64 block into the lower-triangle slot (columns <= drow), overwriting or accumulating
      auto put_lower = [&](float* as, const float* v, bool overwrite) {
        if (drow < H) {
          float* rowp = as + drow * (drow + 1) / 2;
#pragma unroll
          for (int c = 0; c < H; ++c)
            if (c <= drow) rowp[c] = overwrite ? v[c] : rowp[c] + v[c];
        } else {
          float4* r4 = reinterpret_cast<float4*>(as + OFF21 + (drow - H) * L21S);
#pragma unroll
          for (int c = 0; c < H; c += 4) {
            float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            if (!overwrite) { const float4 e = r4[c / 4]; o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
            r4[c / 4] = o;
          }
          float* rowp = as + OFF22 + (drow - H) * (drow - H + 1) / 2;
#pragma unroll
          for (int c = 0; c < H; ++c)
            if (c <= drow - H) rowp[c] = overwrite ? v[H + c] : rowp[c] + v[H + c];
        }
      };
      if (q < 2) {
#pragma unroll 1
        for (int s = 0; s < NSLOT; ++s) {
          if (d.row[s] < 0) continue;
          float v[64];
          load_slot(s, v);
          put_lower(&sm.aslot[t][s][0], v, d.first[s] != 0);
        }
      }
      named_bar_sync(barid, 128);
      if (q >= 2) {
#pragma unroll 1
        for (int s = 0; s < NSLOT; ++s) {
          if (d.row[s] < 0) continue;
          float v[64];
          load_slot(s, v);
          put_lower(&sm.aslot[t][s][0], v, false);
        }
      }
      named_bar_sync(barid, 128);
      if (q >= 2) {
#pragma unroll 1
        for (int s = 0; s < NSLOT; ++s) {
          if (d.row[s] < 0) continue;
          float v[64];
          load_slot(s, v);
          float* as = &sm.aslot[t][s][0];
          // transpose: LH[drow][c] (c >= drow) goes to A[c][drow]; consecutive lanes -> consecutive addresses
          if (drow < H) {
#pragma unroll
            for (int c = 0; c < H; ++c)
              if (c >= drow) as[c * (c + 1) / 2 + drow] += v[c];
#pragma unroll
            for (int c = H; c < KP; ++c) as[OFF21 + (c - H) * L21S + drow] += v[c];
          } else {
#pragma unroll
            for (int c = H; c < KP; ++c)
              if (c >= drow) as[OFF22 + (c - H) * (c - H + 1) / 2 + (drow - H)] += v[c];
          }
        }
      }
      __syncwarp();
      T_END(2);
      // right-hand side of this segment
      T_BEGIN();
      mbar_wait(&sm.bfull[t][bf], ph);
      T_END(3);
      {
        const int tt = (warp - TEAM0) % 4 * 32 + lane;  // 0..127 within the team
        for (int o = tt; o < NSLOT * KP; o += 128) {
          const int s = o / KP, c = o % KP;
          if (d.row[s] >= 0) sm.bslot[t][s][c] = d.first[s] ? sm.bstage[t][bf][s][c] : sm.bslot[t][s][c] + sm.bstage[t][bf][s][c];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tmemfree[set]);
      T_BEGIN();
      named_bar_sync(barid, 128);
      T_END(4);
      T_BEGIN();
      // ---- solve
      const int r = d.row[slot];
      if (r >= 0 && d.last[slot] && tp.dbg) {
        float* o = tp.dbg + (size_t)r * (ASLOT + KP);
        for (int c = lane; c < ASLOT; c += 32) o[c] = sm.aslot[t][slot][c];
        for (int c = lane; c < KP; c += 32) o[ASLOT + c] = sm.bslot[t][slot][c];
        __syncwarp();
      }
      if (r >= 0 && d.last[slot] && tp.out) {
        // split mode: hand the finished normal equations to the solver kernel (coalesced 16-byte stores)
        float4* o4 = reinterpret_cast<float4*>(tp.out + (size_t)(r - tp.out_row0) * (ASLOT + KP));
        const float4* a4 = reinterpret_cast<const float4*>(&sm.aslot[t][slot][0]);
        for (int c = lane; c < ASLOT / 4; c += 32) o4[c] = a4[c];
        const float4* b4 = reinterpret_cast<const float4*>(&sm.bslot[t][slot][0]);
        if (lane < KP / 4) o4[ASLOT / 4 + lane] = b4[lane];
      } else if (r >= 0 && d.last[slot]) {
        const float ridge = p.lambda * p.nreg[r];
        chol_solve_warp<KP, 8, 72, IMPLICIT, true>(&sm.aslot[t][slot][0], &sm.bslot[t][slot][0], p.yty, ridge, p.k,
                                                   &sm.colbuf[t * NSLOT + slot][0], &sm.dinv[t * NSLOT + slot][0],
                                                   p.dst + (size_t)(p.dst_row_offset + r) * KP, p.fail);
      }
      __syncwarp();
      T_END(5);
      if (lane == 0) mbar_arrive(&sm.teamdone[t][bf]);
      ++nb;
    }
  }
  if (timing && lane == 0) {
    tacc[7] = clock64() - tstart_;
    for (int i = 0; i < 8; ++i) tp.timing[((size_t)blockIdx.x * 16 + warp) * 8 + i] = tacc[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- split mode, second half: one warp per row solves the normal equations the tensor-core kernel left in `in`
// (ASLOT + KP floats per local row). 128-thread CTAs, four per SM by registers: all 16 resident warps factorise.
constexpr int SOLVE_WARPS = 4;
constexpr int SOLVE_SMEM_PER_WARP = ASLOT + KP + 2 * KP + KP;   // slot, b, column double buffer, 1/diag
template <bool IMPLICIT>
__global__ void __launch_bounds__(SOLVE_WARPS * 32, 3) als_solve_packed_kernel(const SolveParams p, const float* __restrict__ in, int in_row0) {
  extern __shared__ __align__(16) float solve_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* slot = solve_smem + warp * SOLVE_SMEM_PER_WARP;
  float* bv = slot + ASLOT;
  float* colbuf = bv + KP;
  float* dinv = colbuf + 2 * KP;
  const int nwarps = gridDim.x * SOLVE_WARPS;
  for (int r = p.row_begin + blockIdx.x * SOLVE_WARPS + warp; r < p.row_end; r += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(in + (size_t)(r - in_row0) * (ASLOT + KP));
    float4* dst4 = reinterpret_cast<float4*>(slot);
    for (int c = lane; c < (ASLOT + KP) / 4; c += 32) cp_async16(dst4 + c, src + c);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    chol_solve_warp<KP, 8, 72, IMPLICIT, true>(slot, bv, p.yty, p.lambda * p.nreg[r], p.k, colbuf, dinv,
                                               p.dst + (size_t)(p.dst_row_offset + r) * KP, p.fail);
    __syncwarp();
  }
}

// what the host launcher needs to know about this role partition
struct Api {
  using Params = TcParams;
  static constexpr int kPerCta = NTEAM * NSLOT;       // rows in flight per CTA (one group of the static assignment)
  static constexpr int kThreads = NTHREADS;
  static constexpr size_t kSmem = sizeof(Smem) + 1024;
  static constexpr int kRowFloats = ASLOT + KP;       // split mode / debug dump: floats per row
  static constexpr int kSolveWarps = SOLVE_WARPS;
  static constexpr size_t kSolveSmem = sizeof(float) * SOLVE_WARPS * SOLVE_SMEM_PER_WARP;
  static void (*kernel(bool imp))(const TcParams) { return imp ? als_solve_tc_kernel<true> : als_solve_tc_kernel<false>; }
  static void (*solver(bool imp))(const SolveParams, const float*, int) {
    return imp ? als_solve_packed_kernel<true> : als_solve_packed_kernel<false>;
  }
};

}  // namespace TC_NS
}  // namespace pio
