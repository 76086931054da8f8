// Standalone probe: one tcgen05.mma (kind::tf32, M=N=128, K=8, A = B = same smem tile) with candidate
// shared-memory layouts / descriptor encodings; prints the max error of D against X^T X.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const unsigned char* img, int img_bytes, uint64_t desc_hi_bits, uint32_t idesc, float* out, uint32_t* info, uint32_t start_off) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ unsigned long long bar;
  __shared__ unsigned int tbase;
  for (int i = threadIdx.x; i < img_bytes; i += blockDim.x) sm[i] = img[i];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  if (threadIdx.x == 0) {
    info[0] = tmem;
    uint64_t d = desc_hi_bits | (uint64_t)(((s32(sm) + start_off) >> 4) & 0x3FFF);
    info[1] = (uint32_t)d; info[2] = (uint32_t)(d >> 32);
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(d), "l"(d), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(&bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb = 0; cb < 8; ++cb) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(w * 32) << 16) + cb * 16;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(w * 32 + lane) * 128 + cb * 16 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int K = 32, MN = 128;
  static float X[K][MN];
  srand(1);
  for (int k = 0; k < K; ++k) for (int m = 0; m < MN; ++m) X[k][m] = tf32((float)(rand() % 2001 - 1000) / 256.0f);
  static double E[MN][MN];
  unsigned char* dimg; float* dout; uint32_t* dinfo;
  const int IMG = 16384;
  cudaMalloc(&dimg, IMG); cudaMalloc(&dout, MN * MN * 4); cudaMalloc(&dinfo, 64);
  static unsigned char img[IMG]; static float out[MN * MN]; uint32_t info[4];
  struct Cand { const char* name; int mn_major; int layout; uint32_t lbo, sbo; int layout_type; int kb; uint32_t start_off; };
  // layout 0: MN-major interleave: (k,mn) at (mn/4)*128 + (k%8)*16 + (mn%4)*4
  // layout 1: K-major interleave: (mn,k) at (mn/8)*256 + (k/4)*128 + (mn%8)*16 + (k%4)*4
  // layout 2: MN-major SWIZZLE_128B atoms: (k, mn) at (mn/32)*1024 + k*128 + (((mn%32)/4) ^ k)*16 + (mn%4)*4
  // layout 1: K-major interleave (8 ratings): (mn,k) at (mn/8)*256 + (k/4)*128 + (mn%8)*16 + (k%4)*4
  // layout 3: K-major SWIZZLE_128B, 32 rating slots per row: (mn,k) at (mn/8)*1024 + (mn%8)*128 + (((k/4) ^ (mn%8))*16) + (k%4)*4
  Cand cands[] = {
      {"K-major none  LBO=128 SBO=256 kb0", 0, 1, 128, 256, 0, 0, 0},
      {"K-major sw128 LBO=16 SBO=1024 kb0", 0, 3, 16, 1024, 2, 0, 0},
      {"K-major sw128 LBO=16 SBO=1024 kb1 (+32B)", 0, 3, 16, 1024, 2, 1, 32},
      {"K-major sw128 LBO=16 SBO=1024 kb2 (+64B)", 0, 3, 16, 1024, 2, 2, 64},
      {"K-major sw128 LBO=0  SBO=1024 kb3 (+96B)", 0, 3, 0, 1024, 2, 3, 96},
      {"MN-major sw128b32 LBO=1024 SBO=512", 1, 4, 1024, 512, 1, 0, 0},
  };
  for (auto& c : cands) {
    memset(img, 0, IMG);
    for (int a = 0; a < MN; ++a) for (int b = 0; b < MN; ++b) { double s2 = 0; for (int k = c.kb * 8; k < c.kb * 8 + 8; ++k) s2 += (double)X[k][a] * X[k][b]; E[a][b] = s2; }
    for (int k = 0; k < K; ++k) for (int m = 0; m < MN; ++m) {
      size_t off;
      if (c.layout == 1) { if (k >= 8) continue; off = (size_t)(m / 8) * 256 + (k / 4) * 128 + (m % 8) * 16 + (k % 4) * 4; }
      else if (c.layout == 3) off = (size_t)(m / 8) * 1024 + (m % 8) * 128 + ((((k / 4) ^ (m % 8))) * 16) + (k % 4) * 4;
      else { if (k >= 8) continue; /* MN-major 128B_BASE32B guess: 32B units */ off = (size_t)(m / 32) * 1024 + (k % 4) * 128 * 0 + (size_t)k * 128 + ((((m % 32) / 8) ^ (k % 4)) * 32) + (m % 8) * 4; }
      memcpy(img + off, &X[k][m], 4);
    }
    uint64_t hi = 0;
    hi |= (uint64_t)((c.lbo >> 4) & 0x3FFF) << 16;
    hi |= (uint64_t)((c.sbo >> 4) & 0x3FFF) << 32;
    hi |= (uint64_t)1 << 46;
    hi |= (uint64_t)c.layout_type << 61;
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.mn_major << 15) | ((uint32_t)c.mn_major << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    cudaMemcpy(dimg, img, IMG, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xFF, MN * MN * 4);
    probe<<<1, 128, IMG + 1024>>>(dimg, IMG, hi, idesc, dout, dinfo, c.start_off);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out, dout, MN * MN * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(info, dinfo, 16, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxabs = 0; int nz = 0;
    for (int a = 0; a < MN; ++a) for (int b = 0; b < MN; ++b) { double d = fabs(out[a * MN + b] - E[a][b]); if (d > maxerr) maxerr = d; if (fabs(out[a*MN+b]) > maxabs) maxabs = fabs(out[a*MN+b]); if (out[a*MN+b] != 0) nz++; }
    printf("%-44s err=%s tmem=%08x desc=%08x%08x idesc=%08x maxerr=%.4g maxabs=%.4g nonzero=%d  D[0][0..3]=%.3f %.3f %.3f %.3f exp %.3f %.3f %.3f %.3f\n", c.name,
           cudaGetErrorString(e), info[0], info[2], info[1], idesc, maxerr, maxabs, nz, out[0], out[1], out[2], out[3], E[0][0], E[0][1], E[0][2], E[0][3]);
  }
  return 0;
}
