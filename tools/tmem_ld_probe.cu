// Probe: register layout of tcgen05.ld.16x256b.x2 (used by the conv_win epilogue).  TMEM is filled through
// tcgen05.st.32x32b.x16 (lane = row, register c = column c: the known layout), then read back in the 16x256b form
// at lane offsets 0 and 16 of each warp's quarter; prints (row, column) held by every register of lanes 0-7 of warp 1
// and checks the documented mapping  reg[4k + 2h + e] of lane L = (base + L/4 + 8h, 8k + 2(L%4) + e)  for all warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_ld_probe tools/tmem_ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(int* out, int* bad) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = (uint32_t)((warp * 32 + lane) * 100 + c);
  const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
               "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int h16 = 0; h16 < 2; ++h16) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tm + ((uint32_t)(warp * 32 + h16 * 16) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) {
      out[((warp * 2 + h16) * 32 + lane) * 8 + i] = (int)r[i];
      const int k = i >> 2, h = (i >> 1) & 1, e = i & 1;
      const int want = (warp * 32 + h16 * 16 + lane / 4 + 8 * h) * 100 + 8 * k + 2 * (lane % 4) + e;
      if ((int)r[i] != want) atomicAdd(bad, 1);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}

int main() {
  int *out, *bad;
  cudaMallocManaged(&out, 4 * 2 * 32 * 8 * sizeof(int));
  cudaMallocManaged(&bad, sizeof(int));
  *bad = 0;
  probe<<<1, 128>>>(out, bad);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe: %s, mismatches vs documented mapping: %d\n", cudaGetErrorString(e), *bad);
  for (int h16 = 0; h16 < 2; ++h16)
    for (int lane = 0; lane < 8; ++lane) {
      printf("warp 1 half %d lane %d:", h16, lane);
      for (int i = 0; i < 8; ++i) {
        const int v = out[((1 * 2 + h16) * 32 + lane) * 8 + i];
        printf(" r%d=(%d,%d)", i, v / 100, v % 100);
      }
      printf("\n");
    }
  return *bad != 0;
}
