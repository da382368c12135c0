// Probe tcgen05 tf32 MMA with MN-major SW128 operands: which (LBO,SBO) convention is right?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt = 2) {
  uint64_t d = 0; d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)lt << 61; return d;
}
// A logical [M=128][K=16], B logical [N=64][K=16]; smem MN-major: block j (32 elems of MN), k row, chunk swizzle
__global__ void probe(const float* A, const float* Bm, float* D, int variant, int KP) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (s32(raw) + 1023) & ~1023u; uint8_t* gen = raw + (base - s32(raw));
  const int REGION = KP * 128;
  uint8_t* sA = gen; uint8_t* sB = gen + 4 * REGION;
  __shared__ uint32_t tslot; __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 30 * 1024 / 4; i += blockDim.x) ((float*)gen)[i] = 0.f;
  __syncthreads();
  if (variant == 2) sB = gen + 16384;
  for (int idx = threadIdx.x; idx < 128 * KP; idx += blockDim.x) {
    int m = idx % 128, k = idx / 128; int blk = m >> 5, e = m & 31; int chunk = e >> 2;
    uint32_t off = blk * REGION + (k >> 3) * 1024 + (k & 7) * 128 + ((chunk ^ (k & 7)) << 4) + (e & 3) * 4;
    if (variant == 2) off = (m >> 3) * 1024 + (m & 7) * 128 + (((k >> 2) ^ (m & 7)) << 4) + (k & 3) * 4;
    if (variant >= 3) off = blk * REGION + k * 128 + ((((e >> 3) ^ (k & 3))) << 5) + (e & 7) * 4;
    *(float*)(sA + off) = A[m * KP + k];
  }
  for (int idx = threadIdx.x; idx < 64 * KP; idx += blockDim.x) {
    int n = idx % 64, k = idx / 64; int blk = n >> 5, e = n & 31; int chunk = e >> 2;
    uint32_t off = blk * REGION + (k >> 3) * 1024 + (k & 7) * 128 + ((chunk ^ (k & 7)) << 4) + (e & 3) * 4;
    if (variant == 2) off = (n >> 3) * 1024 + (n & 7) * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
    if (variant >= 3) off = blk * REGION + k * 128 + ((((e >> 3) ^ (k & 3))) << 5) + (e & 7) * 4;
    *(float*)(sB + off) = Bm[n * KP + k];
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(64u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int j = 0; j < KP / 8; ++j) {
      uint32_t lbo = REGION, sbo = 1024;
      if (variant == 1) { lbo = 1024; sbo = REGION; }
      uint64_t ad = mkdesc(base + j * 1024, lbo, sbo), bd = mkdesc(base + 4 * REGION + j * 1024, lbo, sbo);
      if (variant == 3) { ad = mkdesc(base + j * 1024, REGION, 512, 1); bd = mkdesc(base + 4 * REGION + j * 1024, REGION, 512, 1); }
      if (variant == 4) { ad = mkdesc(base + j * 1024, 512, REGION, 1); bd = mkdesc(base + 4 * REGION + j * 1024, 512, REGION, 1); }
      if (variant == 2) { idesc &= ~((1u << 15) | (1u << 16)); ad = mkdesc(base + j * 32, 16, 1024); bd = mkdesc(base + 16384 + j * 32, 16, 1024); }
      uint32_t acc = j > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra Dn;\n\tbra W;\n\tDn:\n\t}" ::"r"(s32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 128) {
    int w = threadIdx.x >> 5;
    for (int n0 = 0; n0 < 64; n0 += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(tm + ((uint32_t)(w * 32) << 16) + n0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[threadIdx.x * 64 + n0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u));
}
int main() {
  for (int KP : {8, 16}) for (int variant = 2; variant < 5; ++variant) {
    std::vector<float> A(128 * KP), B(64 * KP), D(128 * 64), R(128 * 64, 0.f);
    for (auto& v : A) v = (float)((rand() % 17) - 8);
    for (auto& v : B) v = (float)((rand() % 13) - 6);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < KP; ++k) s += A[m * KP + k] * B[n * KP + k]; R[m * 64 + n] = s; }
    float *dA, *dB, *dD; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    probe<<<1, 128, 40 * 1024>>>(dA, dB, dD, variant, KP);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; float maxe = 0; for (int i = 0; i < 128 * 64; ++i) { float d = fabsf(D[i] - R[i]); if (d > 1e-3f) ++bad; maxe = fmaxf(maxe, d); }
    printf("KP %d variant %d: %s bad %d / %d maxerr %g  D[0..3]= %g %g %g %g  R= %g %g %g %g | D[33*64]= %g R %g\n", KP, variant, cudaGetErrorString(e), bad, 128 * 64, maxe, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3], D[33*64], R[33*64]);
  }
  return 0;
}
