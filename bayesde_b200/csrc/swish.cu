// Trainable swish of the conv stack f_x (brax/_impl/arch.py:14-29,172-174: a = z * sigmoid(z * softplus(beta)), beta per channel,
// beta LIVES IN THE FLAT WEIGHT VECTOR w(t) right after the conv it follows, initial value 0.5).  swish is not monotone, so its
// derivative cannot be rebuilt from the activation output the way the other activations' epilogues do: for fx_actfn = "swish" the
// convolution launches write the PRE-activation z (epilogue BIAS) and these elementwise kernels apply the activation (forward) and
// its derivative plus the per-channel beta gradient (reverse).  Two extra passes over the 64-channel tensors per layer: the price of
// an optional, non-default activation; the default path is untouched.
#include <algorithm>

#include "common.cuh"

namespace bsde {

__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// a[s][pix][c] = z * sigmoid(z * softplus(beta[s][c])),  beta = w + s * w_sample_stride + beta_off
__global__ void swish_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w, long long wss, long long beta_off,
                                 long long per_sample, int C, float* __restrict__ a) {
  const int s = blockIdx.y;
  const float* __restrict__ beta = w + (long long)s * wss + beta_off;
  const float* __restrict__ zs = z + (long long)s * per_sample;
  float* __restrict__ as = a + (long long)s * per_sample;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float g = softplusf_(__ldg(beta + c));
    const float zz = zs[i];
    as[i] = zz * sigmoidf_(zz * g);
  }
}

// dz <- scale * q * da/dz  in place (q = the data gradient the conv launch left in dz), and the partial sums of
// dL/dbeta[c] = sum_pix scale * q * z^2 * sig (1 - sig) * sigmoid(beta[c]) per (sample, CTA): partial[s][blockIdx.x][c].
// Thread t owns channel t % C for all its positions (blockDim.x is a multiple of C), so the per-thread sums are per channel.
__global__ void swish_bwd_kernel(float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ w, long long wss,
                                 long long beta_off, long long per_sample, int C, float scale, float* __restrict__ partial) {
  extern __shared__ float sh[];
  const int s = blockIdx.y;
  const float* __restrict__ beta = w + (long long)s * wss + beta_off;
  float* __restrict__ d = dz + (long long)s * per_sample;
  const float* __restrict__ zs = z + (long long)s * per_sample;
  const int c = threadIdx.x % C;
  const float b = __ldg(beta + c);
  const float g = softplusf_(b), sb = sigmoidf_(b);
  float acc = 0.f;
  // stride gridDim.x * blockDim.x is a multiple of C, so i % C == c for every element this thread visits
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
    const float zz = zs[i], q = scale * d[i];
    const float sg = sigmoidf_(zz * g);
    const float t1 = sg * (1.f - sg);
    d[i] = q * (sg + zz * g * t1);
    acc = fmaf(q, zz * zz * t1 * sb, acc);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < C) {
    float r = 0.f;
    for (int j = threadIdx.x; j < blockDim.x; j += C) r += sh[j];   // fixed order
    partial[((long long)s * gridDim.x + blockIdx.x) * C + threadIdx.x] = r;
  }
}

// gw[s][beta_off + c] = sum over CTAs (fixed order) of the partial sums
__global__ void swish_beta_reduce_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ gw, long long gss,
                                         long long beta_off) {
  const int s = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float r = 0.f;
    for (int j = 0; j < nblk; ++j) r += partial[((long long)s * nblk + j) * C + c];
    gw[(long long)s * gss + beta_off + c] = r;
  }
}

static int swish_grid(long long per_sample, int C, int& threads) {
  threads = (256 / C) * C;            // a multiple of C (C <= 256)
  long long blocks = (per_sample + threads * 8LL - 1) / (threads * 8LL);
  return (int)std::max(1LL, std::min(blocks, 148LL * 4));
}

size_t swish_partial_floats(int S, long long per_sample, int C) {
  int threads;
  return (size_t)S * swish_grid(per_sample, C, threads) * C;
}

int swish_forward(const float* z, const float* w, long long wss, long long beta_off, int S, long long per_sample, int C, float* a,
                  cudaStream_t st) {
  BSDE_CHECK_ARG(C >= 1 && C <= 256, "swish: %d channels unsupported", C);
  int threads;
  const int blocks = swish_grid(per_sample, C, threads);
  swish_fwd_kernel<<<dim3(blocks, S), 256, 0, st>>>(z, w, wss, beta_off, per_sample, C, a);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int swish_backward(float* dz, const float* z, const float* w, long long wss, long long beta_off, int S, long long per_sample, int C,
                   float scale, float* partial, float* gw, long long gss, cudaStream_t st) {
  BSDE_CHECK_ARG(C >= 1 && C <= 256, "swish: %d channels unsupported", C);
  int threads;
  const int blocks = swish_grid(per_sample, C, threads);
  swish_bwd_kernel<<<dim3(blocks, S), threads, threads * sizeof(float), st>>>(dz, z, w, wss, beta_off, per_sample, C, scale, partial);
  BSDE_CUDA_OK(cudaGetLastError());
  swish_beta_reduce_kernel<<<S, 64, 0, st>>>(partial, blocks, C, gw, gss, beta_off);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return 0;
}

}  // namespace bsde
