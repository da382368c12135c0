// Config 2 (torch toy, row a24): the Euler-Maruyama posterior/prior solve of examples/torch/sdebnn_toy1d.py:19-32
// (em_solver) with the drift net of construct_odenet([1, h1, h2, 1]) (:35-41):
//     ConcatLinear(1,h1) -> TimeDependentSwish -> ConcatLinear(h1,h2) -> TimeDependentSwish -> ConcatLinear(h2,1)
// (torchbnn/_impl/diffeq_layers.py:178-188: time channel FIRST; torchbnn/_impl/basic.py:298-312: x*sigmoid(x*beta(t))).
// beta(t) depends on t only, so the host evaluates it for the nt-1 grid points (torch autograd owns its little
// MLP); the kernels take beta1[k][h1], beta2[k][h2] and return dL/dbeta.
//
// One warp integrates one sample over all steps (the state is a scalar); lanes own hidden units.  Per step
// (quirk Q6: the prior drift is evaluated at the NEW state):
//     f = f(t_k, z_k);  z_{k+1} = z_k + f dt + sigma dW_k;  u = (f + theta z_{k+1}) / sigma;  kl += dt u^2 / 2
// The reverse pass recomputes the hidden layers from the stored z_k and emits the per-(step, sample) layer
// gradients; the parameter gradients are their sums over (step, sample), taken by the caller.
#include "common.cuh"

namespace bsde {

constexpr int TOY_MAXH = 128;
constexpr int TOY_UPL = TOY_MAXH / 32;  // hidden units per lane

struct ToyNet {
  const float* W1; const float* b1;   // [h1][2], [h1]
  const float* W2; const float* b2;   // [h2][h1+1], [h2]
  const float* W3; const float* b3;   // [1][h2+1], [1]
  const float* beta1; const float* beta2;  // [nt-1][h1], [nt-1][h2]
};

__device__ __forceinline__ float toy_sig(float x) { return 1.f / (1.f + expf(-x)); }

// forward through the drift net for one sample held by one warp; returns f (all lanes), fills a/h registers
__device__ __forceinline__ float toy_drift(const bsde_toy_desc& d, const ToyNet& p, int k, float t, float z, int lane,
                                           float (&a1)[TOY_UPL], float (&h1)[TOY_UPL], float (&a2)[TOY_UPL], float (&h2)[TOY_UPL]) {
#pragma unroll
  for (int u = 0; u < TOY_UPL; ++u) {
    const int i = lane + 32 * u;
    a1[u] = 0.f; h1[u] = 0.f;
    if (i < d.h1) {
      a1[u] = fmaf(__ldg(p.W1 + 2 * i), t, fmaf(__ldg(p.W1 + 2 * i + 1), z, __ldg(p.b1 + i)));
      h1[u] = a1[u] * toy_sig(a1[u] * __ldg(p.beta1 + (long long)k * d.h1 + i));
    }
  }
#pragma unroll
  for (int u = 0; u < TOY_UPL; ++u) {
    const int j = lane + 32 * u;
    a2[u] = 0.f; h2[u] = 0.f;
    if (j < d.h2) a2[u] = fmaf(__ldg(p.W2 + (long long)j * (d.h1 + 1)), t, __ldg(p.b2 + j));
  }
  for (int i = 0; i < d.h1; ++i) {
    const float hi = __shfl_sync(0xffffffffu, h1[i >> 5], i & 31);
#pragma unroll
    for (int u = 0; u < TOY_UPL; ++u) {
      const int j = lane + 32 * u;
      if (j < d.h2) a2[u] = fmaf(__ldg(p.W2 + (long long)j * (d.h1 + 1) + 1 + i), hi, a2[u]);
    }
  }
  float part = 0.f;
#pragma unroll
  for (int u = 0; u < TOY_UPL; ++u) {
    const int j = lane + 32 * u;
    if (j < d.h2) {
      h2[u] = a2[u] * toy_sig(a2[u] * __ldg(p.beta2 + (long long)k * d.h2 + j));
      part = fmaf(__ldg(p.W3 + 1 + j), h2[u], part);
    }
  }
  return warp_sum(part) + fmaf(__ldg(p.W3), t, __ldg(p.b3));
}

__global__ void toy_em_fwd_kernel(bsde_toy_desc d, ToyNet p, const float* __restrict__ ts, const float* __restrict__ z0,
                                  const float* __restrict__ dW, unsigned long long seed, float* __restrict__ zt,
                                  float* __restrict__ kl) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= d.n) return;
  float z = z0[s], acc = 0.f;
  if (lane == 0) zt[s] = z;
  float a1[TOY_UPL], h1[TOY_UPL], a2[TOY_UPL], h2[TOY_UPL];
  for (int k = 0; k + 1 < d.nt; ++k) {
    const float t = ts[k], dt = ts[k + 1] - ts[k];
    const float f = toy_drift(d, p, k, t, z, lane, a1, h1, a2, h2);
    const float w = dW ? dW[(long long)k * d.n + s] : sqrtf(fmaxf(dt, 0.f)) * philox_normal(seed, (uint32_t)s, (uint32_t)k, 0u, 0x70790001u);
    z = z + f * dt + d.sigma * w;
    const float u = (f + d.theta * z) / d.sigma;
    acc += dt * fabsf(u * u) * 0.5f;
    if (lane == 0) zt[(long long)(k + 1) * d.n + s] = z;
  }
  if (lane == 0) kl[s] = acc;
}

__global__ void toy_em_bwd_kernel(bsde_toy_desc d, ToyNet p, const float* __restrict__ ts, const float* __restrict__ zt,
                                  const float* __restrict__ gzt, const float* __restrict__ gkl, float* __restrict__ gz0,
                                  float* __restrict__ DA1, float* __restrict__ DA2, float* __restrict__ DF,
                                  float* __restrict__ H1, float* __restrict__ H2, float* __restrict__ GB1, float* __restrict__ GB2) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= d.n) return;
  const int K = d.nt - 1;
  float lam = gzt ? gzt[(long long)K * d.n + s] : 0.f;
  const float gk = gkl ? gkl[s] : 0.f;
  float a1[TOY_UPL], h1[TOY_UPL], a2[TOY_UPL], h2[TOY_UPL];
  for (int k = K - 1; k >= 0; --k) {
    const float t = ts[k], dt = ts[k + 1] - ts[k];
    const float z = zt[(long long)k * d.n + s], zn = zt[(long long)(k + 1) * d.n + s];
    const float f = toy_drift(d, p, k, t, z, lane, a1, h1, a2, h2);
    const float u = (f + d.theta * zn) / d.sigma;
    const float du = gk * dt * u;
    const float dzn = lam + du * d.theta / d.sigma;
    const float df = du / d.sigma + dzn * dt;
    const long long ks = (long long)k * d.n + s;
    if (lane == 0) DF[ks] = df;
    // layer 3 -> h2 -> a2
    float da2[TOY_UPL];
#pragma unroll
    for (int uu = 0; uu < TOY_UPL; ++uu) {
      const int j = lane + 32 * uu;
      da2[uu] = 0.f;
      if (j < d.h2) {
        const float b = __ldg(p.beta2 + (long long)k * d.h2 + j);
        const float sg = toy_sig(a2[uu] * b);
        const float dh = df * __ldg(p.W3 + 1 + j);
        da2[uu] = dh * (sg + a2[uu] * b * sg * (1.f - sg));
        GB2[ks * d.h2 + j] = dh * a2[uu] * a2[uu] * sg * (1.f - sg);
        DA2[ks * d.h2 + j] = da2[uu];
        H2[ks * d.h2 + j] = h2[uu];
      }
    }
    // dh1[i] = sum_j da2[j] W2[j][1+i]
    float dzp = 0.f;
#pragma unroll
    for (int uu = 0; uu < TOY_UPL; ++uu) {
      const int i0 = 32 * uu;
      if (i0 >= d.h1) break;
      float dh1 = 0.f;
      // every lane needs the sum over ALL j for its own i: loop j over the warp's units by shuffle
      for (int j = 0; j < d.h2; ++j) {
        const float dj = __shfl_sync(0xffffffffu, da2[j >> 5], j & 31);
        const int i = lane + i0;
        if (i < d.h1) dh1 = fmaf(dj, __ldg(p.W2 + (long long)j * (d.h1 + 1) + 1 + i), dh1);
      }
      const int i = lane + i0;
      if (i < d.h1) {
        const float b = __ldg(p.beta1 + (long long)k * d.h1 + i);
        const float sg = toy_sig(a1[uu] * b);
        const float da1 = dh1 * (sg + a1[uu] * b * sg * (1.f - sg));
        GB1[ks * d.h1 + i] = dh1 * a1[uu] * a1[uu] * sg * (1.f - sg);
        DA1[ks * d.h1 + i] = da1;
        H1[ks * d.h1 + i] = h1[uu];
        dzp = fmaf(da1, __ldg(p.W1 + 2 * i + 1), dzp);
      }
    }
    lam = dzn + warp_sum(dzp);
    if (gzt) lam += gzt[(long long)k * d.n + s];
  }
  if (lane == 0) gz0[s] = lam;
}

static int toy_check(const bsde_toy_desc* d) {
  BSDE_CHECK_ARG(d != nullptr, "NULL descriptor");
  BSDE_CHECK_ARG(d->n >= 1 && d->nt >= 2, "n >= 1 and nt >= 2 required");
  BSDE_CHECK_ARG(d->h1 >= 1 && d->h1 <= TOY_MAXH && d->h2 >= 1 && d->h2 <= TOY_MAXH, "hidden widths must be in [1, %d]", TOY_MAXH);
  BSDE_CHECK_ARG(d->sigma != 0.f, "sigma must be non-zero");
  return 0;
}

}  // namespace bsde

using namespace bsde;

extern "C" int bsde_toy_em_fwd(const bsde_toy_desc* d, const float* d_ts, const float* d_z0, const float* d_dW, uint64_t seed,
                               const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                               const float* b3, const float* beta1, const float* beta2, float* d_zt, float* d_kl,
                               void* stream) {
  if (int e = toy_check(d)) return e;
  BSDE_CHECK_ARG(d_ts && d_z0 && W1 && b1 && W2 && b2 && W3 && b3 && beta1 && beta2 && d_zt && d_kl, "NULL pointer");
  ToyNet p{W1, b1, W2, b2, W3, b3, beta1, beta2};
  const int threads = 128, blocks = (d->n * 32 + threads - 1) / threads;
  toy_em_fwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*d, p, d_ts, d_z0, d_dW, seed, d_zt, d_kl);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int bsde_toy_em_bwd(const bsde_toy_desc* d, const float* d_ts, const float* d_zt, const float* d_gzt,
                               const float* d_gkl, const float* W1, const float* b1, const float* W2, const float* b2,
                               const float* W3, const float* b3, const float* beta1, const float* beta2, float* d_gz0,
                               float* d_DA1, float* d_DA2, float* d_DF, float* d_H1, float* d_H2, float* d_GB1,
                               float* d_GB2, void* stream) {
  if (int e = toy_check(d)) return e;
  BSDE_CHECK_ARG(d_ts && d_zt && W1 && b1 && W2 && b2 && W3 && b3 && beta1 && beta2 && d_gz0 && d_DA1 && d_DA2 && d_DF && d_H1 &&
                     d_H2 && d_GB1 && d_GB2,
                 "NULL pointer");
  ToyNet p{W1, b1, W2, b2, W3, b3, beta1, beta2};
  const int threads = 128, blocks = (d->n * 32 + threads - 1) / threads;
  toy_em_bwd_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*d, p, d_ts, d_zt, d_gzt, d_gkl, d_gz0, d_DA1, d_DA2, d_DF, d_H1,
                                                                  d_H2, d_GB1, d_GB2);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}
