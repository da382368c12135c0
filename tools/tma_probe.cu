// Probe cp.async.bulk.tensor (TMA tiled mode, SWIZZLE_128B) semantics needed by a TMA-fed window:
//   * is the 128B swizzle a function of the ABSOLUTE shared-memory address (like the tcgen05 descriptors), i.e.
//     may the destination be 128-byte aligned but not 1024-byte aligned (rows of a 33-position pitch)?
//   * out-of-bounds coordinates (x = W, y = H, negative start) -> zero fill?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_probe tools/tma_probe.cu   (no -lcuda: the
// driver entry point is fetched through cudaGetDriverEntryPoint)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int off, int c0, int x0, int y0, int n0, int bytes, float* out) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (s32(raw) + 1023) & ~1023u;
  uint8_t* gen = raw + (base - s32(raw));
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) ((float*)gen)[i] = -1.0f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(base + off),
        "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(x0), "r"(y0), "r"(n0)
        : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra Dn;\n\tbra W;\n\tDn:\n\t}" ::"r"(s32(&bar)) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) out[i] = ((float*)gen)[i];
}

int main() {
  const int N = 2, H = 4, W = 8, C = 64;
  std::vector<float> t(N * H * W * C);
  for (int i = 0; i < (int)t.size(); ++i) t[i] = (float)i;  // value = flat index
  float *dT, *dOut;
  cudaMalloc(&dT, t.size() * 4); cudaMalloc(&dOut, 8192);
  cudaMemcpy(dT, t.data(), t.size() * 4, cudaMemcpyHostToDevice);
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres);
  printf("entry point: %s qres %d ptr %p\n", cudaGetErrorString(e), (int)qres, (void*)enc);
  if (!enc) return 1;
  CUtensorMap tm;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, W + 1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dT, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  struct Case { int off, c0, x0, y0, n0; } cases[] = {{0, 32, 0, 1, 1}, {128, 32, 0, 1, 1}, {384, 0, 0, 2, 0}, {1152, 32, 0, 3, 1},
                                                        {128, 0, -1, 1, 1}, {256, 0, 0, 4, 1}, {256, 0, 0, 1, 2}};
  for (auto cs : cases) {
    probe<<<1, 128, 16 * 1024>>>(tm, cs.off, cs.c0, cs.x0, cs.y0, cs.n0, (W + 1) * 128, dOut);
    cudaError_t ee = cudaDeviceSynchronize();
    std::vector<float> o(2048);
    cudaMemcpy(o.data(), dOut, 8192, cudaMemcpyDeviceToHost);
    printf("case off %d c0 %d x0 %d y %d n %d: %s\n", cs.off, cs.c0, cs.x0, cs.y0, cs.n0, cudaGetErrorString(ee));
    if (ee != cudaSuccess) return 1;
    int abs_ok = 0, rel_ok = 0, zero_rows = 0, tot = 0;
    for (int bx = 0; bx <= W; ++bx) {
      const int x = cs.x0 + bx;
      const bool inb = x >= 0 && x < W && cs.y0 >= 0 && cs.y0 < H && cs.n0 >= 0 && cs.n0 < N;
      const int row_byte = cs.off + bx * 128;
      for (int c = 0; c < 8; ++c) {
        const float want = inb ? (float)(((cs.n0 * H + cs.y0) * W + x) * C + cs.c0 + c * 4) : 0.f;
        const int slot_abs = c ^ ((row_byte >> 7) & 7), slot_rel = c ^ (bx & 7);
        const float got_abs = o[(row_byte + slot_abs * 16) / 4], got_rel = o[(row_byte + slot_rel * 16) / 4];
        ++tot;
        if (got_abs == want) ++abs_ok;
        if (got_rel == want) ++rel_ok;
      }
      if (!inb) ++zero_rows;
    }
    // untouched bytes before / after the box keep the -1 sentinel
    const bool before = cs.off == 0 || o[(cs.off - 16) / 4] == -1.0f, after = o[(cs.off + (W + 1) * 128) / 4] == -1.0f;
    printf("   absolute-address swizzle matches %d/%d, box-relative %d/%d, oob rows %d, sentinel before %d after %d\n", abs_ok, tot,
           rel_ok, tot, zero_rows, (int)before, (int)after);
  }
  // ---------------- part 2: SWIZZLE_128B_ATOM_32B (the MN-major BASE32B operand layout) and elementStrides = 2 ----------------
  {
    CUtensorMap tm2;
    cuuint32_t box2[4] = {32, 2 * 5, 1, 1};   // 5 pixels with traversal stride 2 -> boxDim = 10
    cuuint32_t es2[4] = {1, 2, 1, 1};
    CUresult r2 = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dT, dims, strides, box2, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode ATOM_32B + stride 2: %d\n", (int)r2);
    if (r2 == CUDA_SUCCESS) {
      struct Case2 { int off, c0, x0, y0, n0; } cases2[] = {{0, 0, 0, 1, 1}, {128, 32, 1, 2, 0}, {384, 0, 1, 3, 1}, {640, 0, 0, 1, 1}};
      for (auto cs : cases2) {
        probe<<<1, 128, 16 * 1024>>>(tm2, cs.off, cs.c0, cs.x0, cs.y0, cs.n0, 5 * 128, dOut);
        cudaError_t ee = cudaDeviceSynchronize();
        std::vector<float> o(2048);
        cudaMemcpy(o.data(), dOut, 8192, cudaMemcpyDeviceToHost);
        printf("case2 off %d c0 %d x0 %d y %d n %d: %s\n", cs.off, cs.c0, cs.x0, cs.y0, cs.n0, cudaGetErrorString(ee));
        if (ee != cudaSuccess) return 1;
        int abs_ok = 0, rel_ok = 0, tot = 0;
        for (int bx = 0; bx < 5; ++bx) {
          const int x = cs.x0 + 2 * bx;   // strided traversal
          const bool inb = x < W;
          const int row_byte = cs.off + bx * 128;
          for (int c = 0; c < 8; ++c) {   // 16-byte chunk c: granule c >> 1, half c & 1
            const float want = inb ? (float)(((cs.n0 * H + cs.y0) * W + x) * C + cs.c0 + c * 4) : 0.f;
            const int slot_abs = (((c >> 1) ^ ((row_byte >> 7) & 3)) << 1) | (c & 1);
            const int slot_rel = (((c >> 1) ^ (bx & 3)) << 1) | (c & 1);
            ++tot;
            if (o[(row_byte + slot_abs * 16) / 4] == want) ++abs_ok;
            if (o[(row_byte + slot_rel * 16) / 4] == want) ++rel_ok;
          }
        }
        printf("   ATOM_32B, stride 2: absolute-address granule swizzle matches %d/%d, box-relative %d/%d, sentinel after %d\n", abs_ok, tot, rel_ok,
               tot, (int)(o[(cs.off + 5 * 128) / 4] == -1.0f));
      }
    }
  }
  return 0;
}
