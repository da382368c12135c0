// Shared device helpers for libbsde.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/bsde.h"

namespace bsde {

void set_error(const char* fmt, ...);

// One-time per-DEVICE initialisation (function attributes, symbol addresses are per device / context: a process that drives
// several GPUs must repeat them on each).  `first()` is true the first time it is called while device d is current.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done[4];
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long m = 1ull << (d & 63);
    return (done[(d >> 6) & 3].fetch_or(m) & m) == 0;
  }
};
void count_launch(int n = 1);  // bookkeeping for bsde_launch_count()
// SMs a persistent launch may size its grid for (default 148).  The two-stream reverse pass (api.cu) halves it so that a data-gradient
// launch and the weight-gradient launch that reads the same dY are co-resident (every kernel here holds one CTA per SM).
int sm_budget();
void set_sm_budget(int n);

#define BSDE_CHECK_ARG(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      bsde::set_error(__VA_ARGS__);      \
      return BSDE_ERR_INVALID_ARG;       \
    }                                    \
  } while (0)

#define BSDE_CUDA_OK(expr)                                                          \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      bsde::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return BSDE_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

// ---- activations (brax/_impl/arch.py:31-32; torch.nn.Softplus threshold 20 is also what
// jax.nn.softplus = logaddexp(x,0) evaluates to in fp32 for x > 20) -------------------------------
// rbf (exp(-x^2), arch.py:31) as the conv epilogues store it: the sign of the pre-activation rides in the last mantissa bit
__device__ __forceinline__ float rbf_fwd_tagged(float x) {
  const uint32_t b = (__float_as_uint(expf(-x * x)) & ~1u) | (x < 0.f ? 1u : 0u);
  return __uint_as_float(b);
}
// d/dx exp(-x^2) = -2 x a from the tagged output a
__device__ __forceinline__ float rbf_grad_from_tagged(float a) {
  // a == 0 (|x| > ~10.2: expf underflows) would give -log(0) = inf and inf * 0 = NaN; the true derivative is 0 there
  if (!(fabsf(a) >= 1.17549435e-38f)) return 0.f;
  const float ax = sqrtf(fmaxf(-logf(a), 0.f));
  return (__float_as_uint(a) & 1u) ? 2.f * ax * a : -2.f * ax * a;
}
__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case BSDE_ACT_RBF: return expf(-x * x);
    case BSDE_ACT_SOFTPLUS: return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    case BSDE_ACT_TANH: return tanhf(x);
    case BSDE_ACT_ELU: return x > 0.f ? x : expm1f(x);
    default: return x;
  }
}
// derivative expressed through the activation OUTPUT a = act(x)
__device__ __forceinline__ float act_grad_from_out(int act, float a) {
  switch (act) {
    case BSDE_ACT_SOFTPLUS: return -expm1f(-a);  // sigmoid(x) = 1 - exp(-softplus(x))
    case BSDE_ACT_TANH: return 1.f - a * a;
    case BSDE_ACT_ELU: return a > 0.f ? 1.f : a + 1.f;
    case BSDE_ACT_RBF: return rbf_grad_from_tagged(a);
    default: return 1.f;
  }
}
// derivative expressed through the pre-activation x
__device__ __forceinline__ float act_grad(int act, float x) {
  switch (act) {
    case BSDE_ACT_SOFTPLUS: return 1.f / (1.f + expf(-x));
    case BSDE_ACT_TANH: { float a = tanhf(x); return 1.f - a * a; }
    case BSDE_ACT_ELU: return x > 0.f ? 1.f : expf(x);
    case BSDE_ACT_RBF: return -2.f * x * expf(-x * x);
    default: return 1.f;
  }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based ----------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  c[0] = hi1 ^ c[1] ^ k0;
  c[1] = lo1;
  c[2] = hi0 ^ c[3] ^ k1;
  c[3] = lo0;
}
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// one standard normal for (element p, iteration k, sample s, stream id)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t p, uint32_t k, uint32_t s,
                                               uint32_t stream_id) {
  uint32_t c[4] = {p, k, s, stream_id};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  // Box-Muller on two 32-bit uniforms in (0,1]
  const float u1 = ((float)c[0] + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float u2 = (float)c[1] * 2.3283064365386963e-10f;           // [0,1)
  const float r = sqrtf(-2.0f * logf(fmaxf(u1, 1e-37f)));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  (void)sn;
  return r * cs;
}

// four standard normals of the increments stream from ONE Philox call: counter (element, iteration, sample / 4, stream);
// sample s takes component s & 3 (two Box-Muller pairs).  Shared by the K1 kernels and bsde_philox_increments, which must agree
// bit for bit, so both use this one function.  lg2 / sin / cos are the MUFU approximations (absolute error ~1e-6 on the normal,
// far below the fp32 rounding of sigma * dW in the update; the law tests bound mean / variance / kurtosis): the K1 pass over P
// is bound by this arithmetic once its state is resident.
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t p, uint32_t k, uint32_t sgroup, uint32_t stream_id,
                                               float (&z)[4]) {
  uint32_t c[4] = {p, k, sgroup, stream_id};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float u1 = ((float)c[2 * h] + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
    const float u2 = (float)c[2 * h + 1] * 2.3283064365386963e-10f;       // [0,1)
    const float r = sqrtf(-1.3862943611198906f * __log2f(fmaxf(u1, 1e-37f)));   // sqrt(-2 ln u1)
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    z[2 * h] = r * cs;
    z[2 * h + 1] = r * sn;
  }
}

// arguments of the fixed-order split reduction of weight-gradient partial tiles (tconv_simt.cu)
struct WreduceArgs {
  const float* partial;
  const float* bias_partial;  // swapped form: [S][splits][cout] column sums, else NULL
  float* gw;
  long long gw_sample_stride;
  int S, splits, rows, cols, Kc, Ktot, has_time, time_row, real_base, swapped, cin, cout;
  long long w_off, b_off, ts, ks, ns;
  float scale;
  int nbias;  // bias partial rows Ktot .. Ktot+nbias-1 are summed (0 means 1)
};
int launch_wgrad_reduce(const WreduceArgs& a, int blocks, cudaStream_t st);

// Launch with programmatic dependent launch allowed: the kernel may be scheduled while its stream predecessor drains (after the
// predecessor's griddepcontrol.launch_dependents) and MUST execute griddepcontrol.wait before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const int pdl = getenv("BSDE_PDL") ? atoi(getenv("BSDE_PDL")) : 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, fn, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace bsde
