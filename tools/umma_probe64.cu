// Probe 2: M=64,N=64,K=8 tf32, MN-major SWIZZLE_128B_BASE32B operands; three MMAs hi*hi + lo*hi + hi*lo into one
// accumulator. Dumps all 128 TMEM lanes x 64 columns to find which lanes hold D rows 0..63.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const unsigned char* img, int img_bytes, uint32_t idesc, float* out) {
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tbase;
  if (threadIdx.x == 0) {
    uint64_t dh = (uint64_t)(((s32(sm)) >> 4) & 0x3FFF);
    dh |= (uint64_t)(1024 >> 4) << 16; dh |= (uint64_t)(512 >> 4) << 32; dh |= (uint64_t)1 << 46; dh |= (uint64_t)1 << 61;
    uint64_t dl = (uint64_t)(((s32(sm) + 2048) >> 4) & 0x3FFF);
    dl |= (uint64_t)(1024 >> 4) << 16; dl |= (uint64_t)(512 >> 4) << 32; dl |= (uint64_t)1 << 46; dl |= (uint64_t)1 << 61;
    uint64_t aa[3] = {dh, dl, dh}, bb[3] = {dh, dh, dl};
    for (int i = 0; i < 3; ++i)
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(aa[i]), "l"(bb[i]), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(s32(&bar)), "r"(0u) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(w * 32) << 16) + cb * 16;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(w * 32 + lane) * 64 + cb * 16 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}
static float tf32r(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
int main() {
  const int K = 8, F = 64;
  static float X[K][F], HI[K][F], LO[K][F];
  srand(2);
  for (int k = 0; k < K; ++k) for (int m = 0; m < F; ++m) { X[k][m] = (float)(rand() % 200001 - 100000) / 77777.0f; HI[k][m] = tf32r(X[k][m]); LO[k][m] = X[k][m] - HI[k][m]; }
  static double E[F][F];
  for (int a = 0; a < F; ++a) for (int b = 0; b < F; ++b) { double s = 0; for (int k = 0; k < K; ++k) s += (double)X[k][a] * X[k][b]; E[a][b] = s; }
  const int IMG = 4096;
  static unsigned char img[IMG];
  for (int k = 0; k < K; ++k) for (int mn = 0; mn < 128; ++mn) {
    float v = mn < 64 ? HI[k][mn] : LO[k][mn - 64];
    size_t off = (size_t)(mn / 32) * 1024 + k * 128 + ((((mn % 32) / 8) ^ (k % 4)) * 32) + (mn % 8) * 4;
    memcpy(img + off, &v, 4);
  }
  unsigned char* dimg; float* dout;
  cudaMalloc(&dimg, IMG); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dimg, img, IMG, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 128 * 64 * 4);
  uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
  probe<<<1, 128, IMG + 1024>>>(dimg, IMG, idesc, dout);
  printf("err=%s idesc=%08x\n", cudaGetErrorString(cudaDeviceSynchronize()), idesc);
  static float out[128 * 64];
  cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost);
  for (int lane = 0; lane < 128; ++lane) {
    int best = -1; double beste = 1e30; double nz = 0;
    for (int c = 0; c < 64; ++c) nz += fabs(out[lane * 64 + c]);
    for (int r = 0; r < 64; ++r) { double e = 0; for (int c = 0; c < 64; ++c) e = fmax(e, fabs(out[lane * 64 + c] - E[r][c])); if (e < beste) { beste = e; best = r; } }
    printf("lane %3d: |row|1=%9.4f best D row %2d maxerr %.3g\n", lane, nz, best, beste);
  }
  return 0;
}
