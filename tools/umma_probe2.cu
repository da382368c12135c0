// Probe tcgen05 kind::tf32 shared-memory descriptor semantics needed by the window-staged convolutions:
//   T_A  K-major, no swizzle, 16-byte chunk planes [chunk][position][16 B]; which of LBO/SBO is which, and
//        does a start address shifted by j positions (j*16 B) address rows j..j+127?
//   T_B  K-major SWIZZLE_128B rows of 128 B, start shifted by j rows (j*128 B), base_offset 0 vs (j&7)
//   T_C  MN-major SWIZZLE_128B_BASE32B, start shifted by j k-rows (j*128 B), base_offset variants
// Generic kernel: the host builds the raw smem image and a list of MMA ops (descriptor high bits + byte offsets).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe2 tools/umma_probe2.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <vector>

struct MmaOp { uint32_t a_off, b_off; uint64_t a_hi, b_hi; uint32_t acc, pad; };

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const uint8_t* img, int bytes, const MmaOp* ops, int nops, uint32_t idesc, float* D, int N) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (s32(raw) + 1023) & ~1023u;
  uint8_t* gen = raw + (base - s32(raw));
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) ((uint32_t*)gen)[i] = ((const uint32_t*)img)[i];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nops; ++i) {
      MmaOp o = ops[i];
      uint64_t ad = o.a_hi | (uint64_t)(((base + o.a_off) >> 4) & 0x3FFF);
      uint64_t bd = o.b_hi | (uint64_t)(((base + o.b_off) >> 4) & 0x3FFF);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
                   "l"(ad), "l"(bd), "r"(idesc), "r"(o.acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra Dn;\n\tbra W;\n\tDn:\n\t}" ::"r"(s32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 128) {
    int w = threadIdx.x >> 5;
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                     "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tm + ((uint32_t)(w * 32) << 16) + n0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[threadIdx.x * N + n0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u));
}

static uint64_t desc_hi(uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}

static int run(const char* name, const std::vector<uint8_t>& img, const std::vector<MmaOp>& ops, uint32_t idesc, int N,
               const std::vector<float>& R) {
  uint8_t* dimg; MmaOp* dops; float* dD;
  cudaMalloc(&dimg, img.size()); cudaMalloc(&dops, ops.size() * sizeof(MmaOp)); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dops, ops.data(), ops.size() * sizeof(MmaOp), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  probe<<<1, 128, img.size() + 2048>>>(dimg, (int)img.size(), dops, (int)ops.size(), idesc, dD, N);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> D(128 * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; float maxe = 0;
  for (int i = 0; i < 128 * N; ++i) { float d = fabsf(D[i] - R[i]); if (d > 1e-3f) ++bad; maxe = fmaxf(maxe, d); }
  printf("%-58s %s bad %5d / %d maxerr %g\n", name, cudaGetErrorString(e), bad, 128 * N, maxe);
  fflush(stdout);
  cudaFree(dimg); cudaFree(dops); cudaFree(dD);
  return bad;
}

static float rv(int m) { return (float)((rand() % (2 * m + 1)) - m); }

int main() {
  const int N = 64;
  // instruction descriptors: D fp32 (1<<4), A tf32 (2<<7), B tf32 (2<<10), N>>3 at 17, M>>4 at 24; bits 15/16 = A/B MN-major
  const uint32_t idesc_k = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_mn = idesc_k | (1u << 15) | (1u << 16);

  // ---------------- T_A: K-major, no swizzle, chunk planes, shifted start ----------------
  {
    const int P = 176, K = 16;                 // positions in the window; K = 16 -> 4 chunks of 16 B, two MMAs
    const int PLA = P * 16 + 16, PLB = 64 * 16;  // plane strides (bytes)
    const int offB = 4 * PLA + 1024 - ((4 * PLA) % 1024);
    std::vector<float> W(P * K), Bm(N * K);
    for (auto& v : W) v = rv(8);
    for (auto& v : Bm) v = rv(6);
    std::vector<uint8_t> img(offB + 4 * PLB, 0);
    for (int p = 0; p < P; ++p) for (int k = 0; k < K; ++k) memcpy(&img[(k >> 2) * PLA + p * 16 + (k & 3) * 4], &W[p * K + k], 4);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) memcpy(&img[offB + (k >> 2) * PLB + n * 16 + (k & 3) * 4], &Bm[n * K + k], 4);
    for (int conv = 0; conv < 1; ++conv)
      for (int j : {0, 1, 3, 8, 35}) {
        std::vector<float> R(128 * N, 0.f);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += W[(m + j) * K + k] * Bm[n * K + k]; R[m * N + n] = s; }
        std::vector<MmaOp> ops;
        for (int ks = 0; ks < K / 8; ++ks) {
          MmaOp o; memset(&o, 0, sizeof(o));
          o.a_off = 2 * ks * PLA + j * 16; o.b_off = offB + 2 * ks * PLB;
          o.a_hi = conv == 0 ? desc_hi(PLA, 128, 0, 0) : desc_hi(128, PLA, 0, 0);
          o.b_hi = conv == 0 ? desc_hi(PLB, 128, 0, 0) : desc_hi(128, PLB, 0, 0);
          o.acc = ks > 0;
          ops.push_back(o);
        }
        char nm[128]; snprintf(nm, sizeof(nm), "T_A K-major noswz planes %s shift %d", conv == 0 ? "LBO=plane SBO=128" : "LBO=128 SBO=plane", j);
        run(nm, img, ops, idesc_k, N, R);
      }
  }
  // ---------------- T_B: K-major SW128 rows, start shifted by j rows ----------------
  {
    const int P = 176, K = 32;
    const int offB = P * 128 + 1024 - ((P * 128) % 1024);
    std::vector<float> W(P * K), Bm(N * K);
    for (auto& v : W) v = rv(8);
    for (auto& v : Bm) v = rv(6);
    std::vector<uint8_t> img(offB + N * 128, 0);
    for (int p = 0; p < P; ++p) for (int k = 0; k < K; ++k) memcpy(&img[p * 128 + (((k >> 2) ^ (p & 7)) << 4) + (k & 3) * 4], &W[p * K + k], 4);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) memcpy(&img[offB + n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4], &Bm[n * K + k], 4);
    for (int bo = 0; bo < 2; ++bo)
      for (int j : {0, 1, 3, 8, 35}) {
        std::vector<float> R(128 * N, 0.f);
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += W[(m + j) * K + k] * Bm[n * K + k]; R[m * N + n] = s; }
        std::vector<MmaOp> ops;
        for (int ks = 0; ks < K / 8; ++ks) {
          MmaOp o; memset(&o, 0, sizeof(o));
          o.a_off = j * 128 + ks * 32; o.b_off = offB + ks * 32;
          o.a_hi = desc_hi(16, 1024, 2, bo ? (j & 7) : 0);
          o.b_hi = desc_hi(16, 1024, 2, 0);
          o.acc = ks > 0;
          ops.push_back(o);
        }
        char nm[128]; snprintf(nm, sizeof(nm), "T_B K-major SW128 row shift %d base_offset %d", j, bo ? (j & 7) : 0);
        run(nm, img, ops, idesc_k, N, R);
      }
  }
  // ---------------- T_C: MN-major SW128_BASE32B, start shifted by j k-rows ----------------
  {
    const int P = 64, M = 128;            // P positions (k), A logical [k][m], B logical [k][n]
    const int REG = P * 128;              // bytes per 32-channel MN block
    const int offB = 4 * REG;
    std::vector<float> A(P * M), Bm(P * N);
    for (auto& v : A) v = rv(8);
    for (auto& v : Bm) v = rv(6);
    std::vector<uint8_t> img(offB + 2 * REG, 0);
    auto put = [&](int off0, int k, int e, float v) {  // e = channel inside the operand
      int blk = e >> 5, c = e & 31;
      int off = off0 + blk * REG + k * 128 + ((((c >> 3) ^ (k & 3))) << 5) + (c & 7) * 4;
      memcpy(&img[off], &v, 4);
    };
    for (int k = 0; k < P; ++k) { for (int m = 0; m < M; ++m) put(0, k, m, A[k * M + m]); for (int n = 0; n < N; ++n) put(offB, k, n, Bm[k * N + n]); }
    const int KT = 16;                    // positions reduced: 2 MMAs of K=8
    for (int mode = 0; mode < 3; ++mode)
      for (int j : {0, 1, 2, 4, 5}) {
        // A rows shifted by j, B rows unshifted:  R[m][n] = sum_k A[k + j][m] * B[k][n]
        std::vector<float> R(128 * N, 0.f);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < KT; ++k) s += A[(k + j) * M + m] * Bm[k * N + n]; R[m * N + n] = s; }
        std::vector<MmaOp> ops;
        for (int ks = 0; ks < KT / 8; ++ks) {
          MmaOp o; memset(&o, 0, sizeof(o));
          o.a_off = j * 128 + ks * 1024; o.b_off = offB + ks * 1024;
          const int bo = mode == 0 ? 0 : mode == 1 ? (j & 3) : (j & 7);
          o.a_hi = desc_hi(REG, 512, 1, bo);
          o.b_hi = desc_hi(REG, 512, 1, 0);
          o.acc = ks > 0;
          ops.push_back(o);
        }
        char nm[128]; snprintf(nm, sizeof(nm), "T_C MN-major BASE32B k-row shift %d base_offset mode %d", j, mode);
        run(nm, img, ops, idesc_mn, N, R);
      }
  }
  // ---------------- T_D: M = 64 (cta_group::1): which TMEM lanes hold accumulator rows 0..63? ----------------
  {
    const int P = 64, K = 32;
    const int offB = 64 * 128;
    std::vector<float> W(P * K), Bm(N * K);
    for (auto& v : W) v = rv(8);
    for (auto& v : Bm) v = rv(6);
    std::vector<uint8_t> img(offB + N * 128, 0);
    for (int p = 0; p < P; ++p) for (int k = 0; k < K; ++k) memcpy(&img[p * 128 + (((k >> 2) ^ (p & 7)) << 4) + (k & 3) * 4], &W[p * K + k], 4);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) memcpy(&img[offB + n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4], &Bm[n * K + k], 4);
    std::vector<float> R(64 * N, 0.f);
    for (int m = 0; m < 64; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += W[m * K + k] * Bm[n * K + k]; R[m * N + n] = s; }
    std::vector<MmaOp> ops;
    for (int ks = 0; ks < K / 8; ++ks) {
      MmaOp o; memset(&o, 0, sizeof(o));
      o.a_off = ks * 32; o.b_off = offB + ks * 32;
      o.a_hi = desc_hi(16, 1024, 2, 0); o.b_hi = desc_hi(16, 1024, 2, 0);
      o.acc = ks > 0;
      ops.push_back(o);
    }
    const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((64u >> 4) << 24);
    uint8_t* dimg; MmaOp* dops; float* dD;
    cudaMalloc(&dimg, img.size()); cudaMalloc(&dops, ops.size() * sizeof(MmaOp)); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dops, ops.data(), ops.size() * sizeof(MmaOp), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, 128 * N * 4);
    probe<<<1, 128, img.size() + 2048>>>(dimg, (int)img.size(), dops, (int)ops.size(), idesc64, dD, N);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    printf("T_D M=64: %s ; accumulator row -> TMEM lane:", cudaGetErrorString(e));
    for (int m = 0; m < 64; ++m) {
      int found = -1;
      for (int l = 0; l < 128 && found < 0; ++l) {
        bool ok = true;
        for (int n = 0; n < N && ok; ++n) ok = fabsf(D[l * N + n] - R[m * N + n]) < 1e-3f;
        if (ok) found = l;
      }
      if (m % 16 == 0) printf("\n   rows %2d..: ", m);
      printf("%d ", found);
    }
    printf("\n");
  }
  return 0;
}
