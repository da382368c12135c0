// K1: the weight-SDE rows [w, kl] of one SDEBNN block, all scan iterations in ONE cooperative
// persistent launch (forward) and one for the reverse pass.
//
// Reference arithmetic (per scan iteration k, brax/_impl/brax.py:49-78 + jaxsde/sdeint.py:48-50):
//     dw  = f_w(w_k, t_k)            MLP of ConcatSquashLinear layers P->d1->..->dL->P
//     u   = (dw + w_k) / sigma        hard-coded OU prior drift -w
//     kl += dt_k * u^2  (+ u_stale(t_k) * dW_k  when stl)
//     w_{k+1} = w_k + dt_k * dw + sigma * dW_k
// The only cross-element coupling is the first layer's P->d1 dot product.  Each pass over P that
// produces w_{k+1} also accumulates w_{k+1} . W1, so one grid-wide barrier per iteration suffices;
// the d1->..->dL middle of the MLP (a few hundred MACs) is recomputed redundantly by every CTA.
// Reductions are two-level with a fixed order (no float atomics): results are run-to-run
// deterministic and identical on every CTA.
//
// HBM traffic: the w trajectory is written once per iteration (the x rows need w_k anyway) and
// the 32 B/element of f_w parameters stay L2-resident across iterations, so the kernel moves far
// fewer DRAM bytes than the 48*P algorithmic figure of SURVEY.md 8.4.
#include <cooperative_groups.h>

#include <mutex>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace bsde {

constexpr int SC = 8;      // samples per launch
constexpr int ME = BSDE_FW_MAX_EDGE;
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;

struct TinyDims {
  int nhid;
  int dims[BSDE_FW_MAX_MID + 2];
  int off[BSDE_FW_MAX_MID + 2];  // offset of layer i's activations in the per-sample smem row
  int wsum;
};

struct WsdeArgs {
  bsde_wsde_desc d;
  bsde_fw_params fw;
  bsde_fw_params gfw;  // bwd only
  TinyDims td;
  const float* w0;
  const float* tk;
  const float* dtk;
  const float* dW;
  const float* wstale;
  float* wtraj;   // [S][nit+1][P]
  float* z1;      // [nit][S][d1]
  float* kl;      // [S]
  // bwd
  const float* gw_traj;
  const float* gkl;
  const float* gwT;
  float* gw0;
  int accumulate;
  // workspace
  float* partials;  // [2][G][SC*ME]
  float* klpart;    // [G][SC]
  float* stpart;    // [G][ME]
  float* glam;      // [S][P] (bwd)
  float* ghold;     // [S][P] (bwd, stage tables with back = 1)
  float ou_coef;
  int s_base;  // global index of local sample 0 (Philox counter)
};

// ---- block-level reduction of NV per-thread values; result (sum over the CTA) to dst[v] -------
template <int NV>
__device__ __forceinline__ void block_reduce_store(float (&v)[NV], float* sh_red, float* dst, int nv_used) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float r = warp_sum(v[i]);
    if (lane == 0) sh_red[warp * NV + i] = r;
  }
  __syncthreads();
  if (threadIdx.x < nv_used) {
    float r = 0.f;
#pragma unroll
    for (int wv = 0; wv < NWARPS; ++wv) r += sh_red[wv * NV + threadIdx.x];
    dst[threadIdx.x] = r;
  }
  __syncthreads();
}

// sum partials over CTAs in a fixed order: out[v] = sum_g part[g*stride + v], v < nv
__device__ __forceinline__ void gather_partials(const float* part, int G, int stride, int nv, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int v = warp; v < nv; v += NWARPS) {
    float r = 0.f;
    for (int g = lane; g < G; g += 32) r += __ldcg(part + (long long)g * stride + v);
    r = warp_sum(r);
    if (lane == 0) out[v] = r;
  }
  __syncthreads();
}

// ---- the tiny middle of f_w, per sample row.  Stores a (pre-gate), g (gate), y (pre-activation)
// and h (activation) for every hidden layer so the reverse pass can reuse them. ----------------
struct TinySmem {
  float* a;  // [SC][wsum]
  float* g;  // [wsum]   (gate is sample-independent)
  float* y;  // [SC][wsum]
  float* h;  // [SC][wsum]
};

__device__ __forceinline__ void tiny_forward(const WsdeArgs& A, const TinySmem& T, const float* z1raw /*[S][d1] smem*/,
                                             int S, float t) {
  const TinyDims& td = A.td;
  const int wsum = td.wsum;
  // layer 0 (first hidden): a = z1 + b1
  {
    const int d1 = td.dims[0];
    for (int idx = threadIdx.x; idx < S * d1; idx += NTHREADS) {
      const int s = idx / d1, j = idx - s * d1;
      const float a = z1raw[s * ME + j] + __ldg(A.fw.b1 + j);
      const float g = A.fw.plain ? 1.f : sigmoidf_(__ldg(A.fw.wt1 + j) * t + __ldg(A.fw.wtb1 + j));
      const float y = A.fw.plain ? a : a * g + __ldg(A.fw.bt1 + j) * t;
      T.a[s * wsum + j] = a;
      if (s == 0) T.g[j] = g;
      T.y[s * wsum + j] = y;
      T.h[s * wsum + j] = act_fwd(A.fw.act, y);
    }
    __syncthreads();
  }
  for (int l = 0; l + 1 < td.nhid; ++l) {
    const int din = td.dims[l], dout = td.dims[l + 1];
    const int oi = td.off[l], oo = td.off[l + 1];
    const float* __restrict__ W = A.fw.Wm[l];
    for (int idx = threadIdx.x; idx < S * dout; idx += NTHREADS) {
      const int s = idx / dout, j = idx - s * dout;
      float a = __ldg(A.fw.bm[l] + j);
      for (int q = 0; q < din; ++q) a = fmaf(T.h[s * wsum + oi + q], __ldg(W + q * dout + j), a);
      const float g = A.fw.plain ? 1.f : sigmoidf_(__ldg(A.fw.wtm[l] + j) * t + __ldg(A.fw.wtbm[l] + j));
      const float y = A.fw.plain ? a : a * g + __ldg(A.fw.btm[l] + j) * t;
      T.a[s * wsum + oo + j] = a;
      if (s == 0) T.g[oo + j] = g;
      T.y[s * wsum + oo + j] = y;
      T.h[s * wsum + oo + j] = act_fwd(A.fw.act, y);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float inv_or_zero(float s) { return s != 0.f ? 1.f / s : 0.f; }

// =============================================================================================
// forward
// =============================================================================================
__global__ void __launch_bounds__(NTHREADS) wsde_fwd_kernel(WsdeArgs A) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float smem[];
  const int S = A.d.S, nit = A.d.nit, P = A.fw.P;
  const int d1 = A.td.dims[0], dL = A.td.dims[A.td.nhid - 1], offL = A.td.off[A.td.nhid - 1];
  const int wsum = A.td.wsum;
  const int G = gridDim.x;
  const long long NT = (long long)G * NTHREADS;
  const long long gtid = (long long)blockIdx.x * NTHREADS + threadIdx.x;

  TinySmem T;
  T.a = smem;
  T.g = T.a + (SC + 1) * wsum;
  T.y = T.g + wsum;
  T.h = T.y + (SC + 1) * wsum;
  float* sh_z1 = T.h + (SC + 1) * wsum;  // [(SC+1)][ME]  (row SC = stale vector)
  float* sh_red = sh_z1 + (SC + 1) * ME;  // [NWARPS][SC*ME]
  float* sh_hst = sh_red + NWARPS * SC * ME;  // [ME] hL of the stale vector

  const float sigma = A.d.sigma, inv_sigma = inv_or_zero(sigma);
  const long long wstride = (long long)(nit + 1) * P;

  // ---- phase 0: w_0 -> trajectory slot 0, partial dots of w_0 (and of w_stale)
  {
    float acc[SC * ME];
#pragma unroll
    for (int i = 0; i < SC * ME; ++i) acc[i] = 0.f;
    float accst[ME];
#pragma unroll
    for (int j = 0; j < ME; ++j) accst[j] = 0.f;
    for (long long p = gtid; p < P; p += NT) {
      float w1[ME];
#pragma unroll
      for (int j = 0; j < ME; ++j) w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f;
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s < S) {
          const float w = __ldg(A.w0 + (A.d.w0_per_sample ? (long long)s * P : 0) + p);
          A.wtraj[s * wstride + p] = w;
#pragma unroll
          for (int j = 0; j < ME; ++j) acc[s * ME + j] = fmaf(w, w1[j], acc[s * ME + j]);
        }
      }
      if (A.d.stl) {
        const float ws = __ldg(A.wstale + p);
#pragma unroll
        for (int j = 0; j < ME; ++j) accst[j] = fmaf(ws, w1[j], accst[j]);
      }
    }
    block_reduce_store<SC * ME>(acc, sh_red, A.partials + (long long)blockIdx.x * SC * ME, SC * ME);
    if (A.d.stl) block_reduce_store<ME>(accst, sh_red, A.stpart + (long long)blockIdx.x * ME, ME);
  }
  grid.sync();
  if (A.d.stl) gather_partials(A.stpart, G, ME, ME, sh_z1 + SC * ME);

  float klacc[SC];
#pragma unroll
  for (int s = 0; s < SC; ++s) klacc[s] = 0.f;

  for (int k = 0; k < nit; ++k) {
    const float t = __ldg(A.tk + k), dt = __ldg(A.dtk + k);
    // stage tables (bsde.h): KL step, noise scale / Philox step index, base slot of the update
    const float dtkl = A.d.d_dtkl ? __ldg(A.d.d_dtkl + k) : dt;
    const float sq = A.d.d_nscale ? __ldg(A.d.d_nscale + k) : sqrtf(fmaxf(dt, 0.f));
    const uint32_t nidx = A.d.d_nidx ? (uint32_t)__ldg(A.d.d_nidx + k) : (uint32_t)k;
    const int back = A.d.d_back ? __ldg(A.d.d_back + k) : 0;
    const float* pin = A.partials + (long long)(k & 1) * G * SC * ME;
    float* pout = A.partials + (long long)((k + 1) & 1) * G * SC * ME;
    gather_partials(pin, G, SC * ME, SC * ME, sh_z1);
    if (blockIdx.x == 0) {
      for (int idx = threadIdx.x; idx < S * d1; idx += NTHREADS) {
        const int s = idx / d1, j = idx - s * d1;
        A.z1[((long long)k * S + s) * d1 + j] = sh_z1[s * ME + j];
      }
    }
    // tiny MLP for the S samples (+ the stale vector as row SC when stl)
    tiny_forward(A, T, sh_z1, A.d.stl ? SC + 1 : S, t);
    if (A.d.stl && threadIdx.x < ME) sh_hst[threadIdx.x] = threadIdx.x < dL ? T.h[SC * wsum + offL + threadIdx.x] : 0.f;
    __syncthreads();

    float acc[SC * ME];
#pragma unroll
    for (int i = 0; i < SC * ME; ++i) acc[i] = 0.f;
    for (long long p = gtid; p < P; p += NT) {
      float w1[ME], w4[ME];
#pragma unroll
      for (int j = 0; j < ME; ++j) {
        w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f;
        w4[j] = j < dL ? __ldg(A.fw.W4 + (long long)j * P + p) : 0.f;
      }
      const float b4 = __ldg(A.fw.b4 + p), bt4 = A.fw.plain ? 0.f : __ldg(A.fw.bt4 + p);
      const float s4 = A.fw.plain ? 1.f : sigmoidf_(__ldg(A.fw.wt4 + p) * t + __ldg(A.fw.wtb4 + p));
      float ust = 0.f;
      if (A.d.stl) {
        float pre = b4;
#pragma unroll
        for (int j = 0; j < ME; ++j) pre = fmaf(sh_hst[j], w4[j], pre);
        const float ws = __ldg(A.wstale + p);
        ust = (pre * s4 + bt4 * t + (A.ou_coef + 1.f) * ws) * inv_sigma;
      }
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s < S) {
          const float w = A.wtraj[s * wstride + (long long)k * P + p];
          float pre = b4;
#pragma unroll
          for (int j = 0; j < ME; ++j) pre = fmaf(T.h[s * wsum + offL + (j < dL ? j : 0)], w4[j], pre);
          const float mlp = pre * s4 + bt4 * t;
          const float dw = mlp + A.ou_coef * w;
          const float u = (mlp + (A.ou_coef + 1.f) * w) * inv_sigma;
          const float dWv = A.dW ? __ldg(A.dW + ((long long)s * nit + k) * P + p)
                                 : sq * philox_normal(A.d.seed, (uint32_t)p, nidx, (uint32_t)(s + A.s_base), A.d.stream_id);
          const float klinc = dtkl * A.d.kl_weight * u * u + ust * dWv;
          klacc[s] += klinc;
          if (A.d.d_kltraj) {   // optional: the kl rows of ys (sdeint.py:138), read-modify-write by the element's owner thread
            float* kt = A.d.d_kltraj + s * wstride + (long long)k * P + p;
            kt[P] = (k == 0 ? 0.f : kt[0]) + klinc;
            if (k == 0) kt[0] = 0.f;
          }
          // base of the update: w_k, w_{k-1} (back = 1) or their average (back = 2: the corrector of a trapezoidal step)
          const float wb = back == 1 ? A.wtraj[s * wstride + (long long)(k - 1) * P + p]
                         : back == 2 ? 0.5f * (A.wtraj[s * wstride + (long long)(k - 1) * P + p] + w) : w;
          const float wn = wb + dt * dw + sigma * dWv;
          A.wtraj[s * wstride + (long long)(k + 1) * P + p] = wn;
#pragma unroll
          for (int j = 0; j < ME; ++j) acc[s * ME + j] = fmaf(wn, w1[j], acc[s * ME + j]);
        }
      }
    }
    if (k + 1 < nit) {
      block_reduce_store<SC * ME>(acc, sh_red, pout + (long long)blockIdx.x * SC * ME, SC * ME);
      grid.sync();
    }
  }
  // ---- kl: block partials, then CTA 0 sums in order
  block_reduce_store<SC>(klacc, sh_red, A.klpart + (long long)blockIdx.x * SC, SC);
  grid.sync();
  if (blockIdx.x == 0) {
    gather_partials(A.klpart, G, SC, SC, sh_z1);
    if (threadIdx.x < S) A.kl[threadIdx.x] = sh_z1[threadIdx.x];
  }
}

// =============================================================================================
// reverse pass
// =============================================================================================
__global__ void __launch_bounds__(NTHREADS) wsde_bwd_kernel(WsdeArgs A) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float smem[];
  const int S = A.d.S, nit = A.d.nit, P = A.fw.P;
  const TinyDims& td = A.td;
  const int d1 = td.dims[0], dL = td.dims[td.nhid - 1], offL = td.off[td.nhid - 1];
  const int wsum = td.wsum;
  const int G = gridDim.x;
  const long long NT = (long long)G * NTHREADS;
  const long long gtid = (long long)blockIdx.x * NTHREADS + threadIdx.x;

  TinySmem T;
  T.a = smem;
  T.g = T.a + (SC + 1) * wsum;
  T.y = T.g + wsum;
  T.h = T.y + (SC + 1) * wsum;
  float* sh_z1 = T.h + (SC + 1) * wsum;       // [(SC+1)][ME]
  float* sh_red = sh_z1 + (SC + 1) * ME;      // [NWARPS][SC*ME]
  float* sh_gh = sh_red + NWARPS * SC * ME;   // [SC][wsum]  dL/dh (then reused as dL/da)
  float* sh_dz1 = sh_gh + SC * wsum;          // [SC][ME]    pending dL/dz1 of iteration k+1
  float* sh_ghL = sh_dz1 + SC * ME;           // [SC][ME]

  const float sigma = A.d.sigma, inv_sigma = inv_or_zero(sigma);
  const long long wstride = (long long)(nit + 1) * P;
  const bool acc_in = A.accumulate != 0;

  // ---- prologue: zero the per-element gradient accumulators, seed lambda
  for (long long p = gtid; p < P; p += NT) {
    if (!acc_in) {
      for (int j = 0; j < d1; ++j) A.gfw.W1[p * d1 + j] = 0.f;
      for (int j = 0; j < dL; ++j) A.gfw.W4[(long long)j * P + p] = 0.f;
      A.gfw.b4[p] = 0.f;
      if (!A.fw.plain) { A.gfw.wt4[p] = 0.f; A.gfw.wtb4[p] = 0.f; A.gfw.bt4[p] = 0.f; }
    }
    for (int s = 0; s < S; ++s) A.glam[(long long)s * P + p] = A.gwT ? __ldg(A.gwT + (long long)s * P + p) : 0.f;
  }
  if (blockIdx.x == 0 && !acc_in) {
    for (int j = threadIdx.x; j < d1; j += NTHREADS) {
      A.gfw.b1[j] = 0.f;
      if (!A.fw.plain) { A.gfw.wt1[j] = 0.f; A.gfw.wtb1[j] = 0.f; A.gfw.bt1[j] = 0.f; }
    }
    for (int l = 0; l + 1 < td.nhid; ++l) {
      const int din = td.dims[l], dout = td.dims[l + 1];
      for (int i = threadIdx.x; i < din * dout; i += NTHREADS) A.gfw.Wm[l][i] = 0.f;
      for (int j = threadIdx.x; j < dout; j += NTHREADS) {
        A.gfw.bm[l][j] = 0.f;
        if (!A.fw.plain) { A.gfw.wtm[l][j] = 0.f; A.gfw.wtbm[l][j] = 0.f; A.gfw.btm[l][j] = 0.f; }
      }
    }
  }
  for (int i = threadIdx.x; i < SC * ME; i += NTHREADS) sh_dz1[i] = 0.f;
  __syncthreads();

  bool pending = false;
  for (int k = nit - 1; k >= 0; --k) {
    const float t = __ldg(A.tk + k), dt = __ldg(A.dtk + k);
    const float dtkl = A.d.d_dtkl ? __ldg(A.d.d_dtkl + k) : dt;
    const int back = A.d.d_back ? __ldg(A.d.d_back + k) : 0;                        // w_{k+1} starts from w_{k-back}
    const int hold_in = (A.d.d_back && k + 1 < nit) ? __ldg(A.d.d_back + k + 1) : 0;  // iteration k+1 started from w_k
    // tiny forward of iteration k from the stored first-layer dots
    for (int idx = threadIdx.x; idx < SC * ME; idx += NTHREADS) {
      const int s = idx / ME, j = idx - s * ME;
      sh_z1[idx] = (s < S && j < d1) ? __ldg(A.z1 + ((long long)k * S + s) * d1 + j) : 0.f;
    }
    __syncthreads();
    tiny_forward(A, T, sh_z1, S, t);

    float part[SC * ME];
#pragma unroll
    for (int i = 0; i < SC * ME; ++i) part[i] = 0.f;
    for (long long p = gtid; p < P; p += NT) {
      float w1[ME], w4[ME], gw1[ME], gw4[ME];
#pragma unroll
      for (int j = 0; j < ME; ++j) {
        w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f;
        w4[j] = j < dL ? __ldg(A.fw.W4 + (long long)j * P + p) : 0.f;
        gw1[j] = 0.f; gw4[j] = 0.f;
      }
      const float b4 = __ldg(A.fw.b4 + p), bt4 = A.fw.plain ? 0.f : __ldg(A.fw.bt4 + p);
      const float s4 = A.fw.plain ? 1.f : sigmoidf_(__ldg(A.fw.wt4 + p) * t + __ldg(A.fw.wtb4 + p));
      float gb4 = 0.f, gwt4 = 0.f, gwtb4 = 0.f, gbt4 = 0.f;
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s < S) {
          float lam = A.glam[(long long)s * P + p];
          if (pending) {
            const float wn = __ldg(A.wtraj + s * wstride + (long long)(k + 1) * P + p);
#pragma unroll
            for (int j = 0; j < ME; ++j) {
              lam = fmaf(w1[j], sh_dz1[s * ME + j], lam);
              gw1[j] = fmaf(wn, sh_dz1[s * ME + j], gw1[j]);
            }
          }
          const float w = __ldg(A.wtraj + s * wstride + (long long)k * P + p);
          float pre = b4;
#pragma unroll
          for (int j = 0; j < ME; ++j) pre = fmaf(T.h[s * wsum + offL + (j < dL ? j : 0)], w4[j], pre);
          const float mlp = pre * s4 + bt4 * t;
          const float u = (mlp + (A.ou_coef + 1.f) * w) * inv_sigma;
          const float gu = __ldg(A.gkl + s) * dtkl * A.d.kl_weight * 2.f * u * inv_sigma;  // dL/d(mlp) via kl
          const float q = dt * lam + gu;  // dL/d mlp
          const float gwk = A.gw_traj ? __ldg(A.gw_traj + ((long long)s * nit + k) * P + p) : 0.f;
          // dL/dw_k without the first-layer term (added next iteration / epilogue)
          float gl = dt * A.ou_coef * lam + (A.ou_coef + 1.f) * gu + gwk;
          if (back == 1) A.ghold[(long long)s * P + p] = lam;   // the identity path of this update ends at w_{k-1}
          else if (back == 2) { A.ghold[(long long)s * P + p] = 0.5f * lam; gl += 0.5f * lam; }  // ... at (w_{k-1} + w_k)/2
          else gl += lam;
          if (hold_in) gl += A.ghold[(long long)s * P + p];
          A.glam[(long long)s * P + p] = gl;
          const float qs = q * s4;
#pragma unroll
          for (int j = 0; j < ME; ++j) {
            part[s * ME + j] = fmaf(w4[j], qs, part[s * ME + j]);
            gw4[j] = fmaf(T.h[s * wsum + offL + (j < dL ? j : 0)], qs, gw4[j]);
          }
          gb4 += qs;
          const float e = q * pre * s4 * (1.f - s4);
          gwt4 = fmaf(e, t, gwt4);
          gwtb4 += e;
          gbt4 = fmaf(q, t, gbt4);
        }
      }
      for (int j = 0; j < d1; ++j) if (pending) A.gfw.W1[p * d1 + j] += gw1[j];
      for (int j = 0; j < dL; ++j) A.gfw.W4[(long long)j * P + p] += gw4[j];
      A.gfw.b4[p] += gb4;
      if (!A.fw.plain) { A.gfw.wt4[p] += gwt4; A.gfw.wtb4[p] += gwtb4; A.gfw.bt4[p] += gbt4; }
    }
    float* pout = A.partials + (long long)(k & 1) * G * SC * ME;
    block_reduce_store<SC * ME>(part, sh_red, pout + (long long)blockIdx.x * SC * ME, SC * ME);
    grid.sync();
    gather_partials(pout, G, SC * ME, SC * ME, sh_ghL);

    // ---- tiny reverse: dL/dhL -> ... -> dL/dz1 ; CTA 0 accumulates the tiny parameter grads
    for (int idx = threadIdx.x; idx < SC * wsum; idx += NTHREADS) sh_gh[idx] = 0.f;
    __syncthreads();
    for (int idx = threadIdx.x; idx < S * dL; idx += NTHREADS) {
      const int s = idx / dL, j = idx - s * dL;
      sh_gh[s * wsum + offL + j] = sh_ghL[s * ME + j];
    }
    __syncthreads();
    for (int l = td.nhid - 1; l >= 0; --l) {
      const int dout = td.dims[l], oo = td.off[l];
      // gh -> ga (in place), and the gate / shift parameter grads
      for (int idx = threadIdx.x; idx < S * dout; idx += NTHREADS) {
        const int s = idx / dout, j = idx - s * dout;
        const float gy = sh_gh[s * wsum + oo + j] * act_grad(A.fw.act, T.y[s * wsum + oo + j]);
        sh_gh[s * wsum + oo + j] = gy;  // keep gy for the parameter grads below
      }
      __syncthreads();
      if (blockIdx.x == 0) {
        float* gb = l == 0 ? A.gfw.b1 : A.gfw.bm[l - 1];
        float* gwt = l == 0 ? A.gfw.wt1 : A.gfw.wtm[l - 1];
        float* gwtb = l == 0 ? A.gfw.wtb1 : A.gfw.wtbm[l - 1];
        float* gbt = l == 0 ? A.gfw.bt1 : A.gfw.btm[l - 1];
        for (int j = threadIdx.x; j < dout; j += NTHREADS) {
          const float g = T.g[oo + j];
          float sb = 0.f, se = 0.f, sy = 0.f;
          for (int s = 0; s < S; ++s) {
            const float gy = sh_gh[s * wsum + oo + j];
            sb += gy * g;
            se += gy * T.a[s * wsum + oo + j] * g * (1.f - g);
            sy += gy;
          }
          gb[j] += sb;
          if (!A.fw.plain) { gwt[j] += se * t; gwtb[j] += se; gbt[j] += sy * t; }
        }
        if (l > 0) {
          const int din = td.dims[l - 1], oi = td.off[l - 1];
          float* gW = A.gfw.Wm[l - 1];
          for (int idx = threadIdx.x; idx < din * dout; idx += NTHREADS) {
            const int qi = idx / dout, j = idx - qi * dout;
            float sw = 0.f;
            for (int s = 0; s < S; ++s) sw += T.h[s * wsum + oi + qi] * sh_gh[s * wsum + oo + j] * T.g[oo + j];
            gW[idx] += sw;
          }
        }
      }
      if (l > 0) {
        const int din = td.dims[l - 1], oi = td.off[l - 1];
        const float* __restrict__ W = A.fw.Wm[l - 1];
        for (int idx = threadIdx.x; idx < S * din; idx += NTHREADS) {
          const int s = idx / din, qi = idx - s * din;
          float r = 0.f;
          for (int j = 0; j < dout; ++j) r = fmaf(__ldg(W + qi * dout + j), sh_gh[s * wsum + oo + j] * T.g[oo + j], r);
          sh_gh[s * wsum + oi + qi] = r;
        }
      } else {
        for (int idx = threadIdx.x; idx < SC * ME; idx += NTHREADS) {
          const int s = idx / ME, j = idx - s * ME;
          sh_dz1[idx] = (s < S && j < d1) ? sh_gh[s * wsum + j] * T.g[j] : 0.f;
        }
      }
      __syncthreads();
    }
    pending = true;
  }

  // ---- epilogue: close lambda_0 with the first-layer term of iteration 0
  for (long long p = gtid; p < P; p += NT) {
    float w1[ME], gw1[ME];
#pragma unroll
    for (int j = 0; j < ME; ++j) { w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f; gw1[j] = 0.f; }
    float tot = 0.f;
#pragma unroll
    for (int s = 0; s < SC; ++s) {
      if (s < S) {
        float lam = A.glam[(long long)s * P + p];
        const float w = __ldg(A.wtraj + s * wstride + p);
#pragma unroll
        for (int j = 0; j < ME; ++j) {
          lam = fmaf(w1[j], sh_dz1[s * ME + j], lam);
          gw1[j] = fmaf(w, sh_dz1[s * ME + j], gw1[j]);
        }
        if (A.d.w0_per_sample) A.gw0[(long long)s * P + p] = lam;
        tot += lam;
      }
    }
    for (int j = 0; j < d1; ++j) A.gfw.W1[p * d1 + j] += gw1[j];
    if (!A.d.w0_per_sample) A.gw0[p] = acc_in ? A.gw0[p] + tot : tot;
  }
}

// =============================================================================================
// host side
// =============================================================================================
static std::mutex g_mu;
static int g_num_sms[64] = {0};  // per device

static int fill_tiny(const bsde_fw_params* fw, TinyDims& td) {
  BSDE_CHECK_ARG(fw->nhid >= 1 && fw->nhid <= BSDE_FW_MAX_MID + 1, "len(fw_dims)=%d unsupported", fw->nhid);
  td.nhid = fw->nhid;
  int off = 0;
  for (int i = 0; i < fw->nhid; ++i) {
    td.dims[i] = fw->dims[i];
    td.off[i] = off;
    off += fw->dims[i];
    BSDE_CHECK_ARG(fw->dims[i] >= 1 && fw->dims[i] <= BSDE_FW_MAX_HID, "fw_dims[%d]=%d out of range", i, fw->dims[i]);
  }
  td.wsum = off;
  BSDE_CHECK_ARG(fw->dims[0] <= ME && fw->dims[fw->nhid - 1] <= ME,
                 "first/last hidden widths must be <= %d (got %d, %d)", ME, fw->dims[0], fw->dims[fw->nhid - 1]);
  BSDE_CHECK_ARG(fw->act >= BSDE_ACT_SOFTPLUS && fw->act <= BSDE_ACT_RBF, "bad fw activation %d", fw->act);
  return 0;
}

static size_t smem_bytes(const TinyDims& td, bool bwd) {
  size_t f = (size_t)(3 * (SC + 1) + 1) * td.wsum + (SC + 1) * ME + NWARPS * SC * ME + ME;
  if (bwd) f += (size_t)SC * td.wsum + 2 * SC * ME;
  return f * sizeof(float);
}

static int grid_for(int P, bool bwd, size_t smem, int& G) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  BSDE_CUDA_OK(cudaGetDevice(&dev));
  int& num_sms = g_num_sms[dev & 63];
  if (num_sms == 0) BSDE_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  int occ = 0;
  if (bwd) BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wsde_bwd_kernel, NTHREADS, smem));
  else BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wsde_fwd_kernel, NTHREADS, smem));
  if (occ < 1) { set_error("wsde kernel does not fit on an SM (smem %zu)", smem); return BSDE_ERR_UNSUPPORTED; }
  if (occ > 2) occ = 2;
  int want = (P + NTHREADS - 1) / NTHREADS;
  G = num_sms * occ;
  if (G > want) G = want;
  if (G < 1) G = 1;
  return 0;
}

constexpr int MAXG = 148 * 2 * 2;  // workspace is sized for at most this many CTAs

static size_t ws_floats(int S, int P, bool bwd) {
  size_t f = (size_t)2 * MAXG * SC * ME + (size_t)MAXG * SC + (size_t)MAXG * ME;
  if (bwd) f += (size_t)2 * (S < SC ? S : SC) * P;  // lambda + the held identity path of back = 1 stages
  return f;
}

}  // namespace bsde

using namespace bsde;

extern "C" size_t bsde_wsde_workspace_bytes(const bsde_wsde_desc* d, const bsde_fw_params* fw) {
  if (!d || !fw) return 0;
  return ws_floats(d->S, fw->P, true) * sizeof(float);
}

static int common_checks(const bsde_wsde_desc* d, const bsde_fw_params* fw) {
  BSDE_CHECK_ARG(d && fw, "NULL descriptor");
  BSDE_CHECK_ARG(d->S >= 1, "S must be >= 1");
  BSDE_CHECK_ARG(d->nit >= 1, "nit must be >= 1");
  BSDE_CHECK_ARG(fw->P >= 1, "P must be >= 1");
  return 0;
}

extern "C" int bsde_wsde_fwd(const bsde_wsde_desc* d, const bsde_fw_params* fw, const float* d_w0,
                             const float* d_tk, const float* d_dtk, const float* d_dW, const float* d_w_stale,
                             float* d_wtraj, float* d_z1, float* d_kl, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (int e = common_checks(d, fw)) return e;
  BSDE_CHECK_ARG(d_w0 && d_tk && d_dtk && d_wtraj && d_z1 && d_kl, "NULL tensor pointer");
  BSDE_CHECK_ARG(!d->stl || d_w_stale, "stl=1 needs d_w_stale");
  WsdeArgs A;
  memset(&A, 0, sizeof(A));
  if (int e = fill_tiny(fw, A.td)) return e;
  const size_t need = ws_floats(d->S, fw->P, false) * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    set_error("wsde_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  const size_t smem = smem_bytes(A.td, false);
  int G = 0;
  if (int e = grid_for(fw->P, false, smem, G)) return e;
  if (G > MAXG) G = MAXG;
  const int S = d->S, P = fw->P, nit = d->nit, d1 = fw->dims[0];
  for (int s0 = 0; s0 < S; s0 += SC) {
    const int sc = S - s0 < SC ? S - s0 : SC;
    A.d = *d; A.d.S = sc; A.fw = *fw; A.ou_coef = d->ou_coef; A.s_base = s0 + (int)d->sample_offset;
    A.w0 = d_w0 + (d->w0_per_sample ? (size_t)s0 * P : 0);
    A.tk = d_tk; A.dtk = d_dtk;
    A.dW = d_dW ? d_dW + (size_t)s0 * nit * P : nullptr;
    A.wstale = d_w_stale;
    A.wtraj = d_wtraj + (size_t)s0 * (nit + 1) * P;
    if (d->d_kltraj) A.d.d_kltraj = d->d_kltraj + (size_t)s0 * (nit + 1) * P;
    A.kl = d_kl + s0;
    float* ws = (float*)workspace;
    A.partials = ws; ws += (size_t)2 * MAXG * SC * ME;
    A.klpart = ws; ws += (size_t)MAXG * SC;
    A.stpart = ws;
    // z1 layout for chunked launches: chunk c occupies [nit][sc][d1] at offset s0*nit*d1
    A.z1 = d_z1 + (size_t)s0 * nit * d1;
    void* args[] = {&A};
    BSDE_CUDA_OK(cudaLaunchCooperativeKernel((void*)wsde_fwd_kernel, dim3(G), dim3(NTHREADS), args, smem,
                                             (cudaStream_t)stream));
    count_launch();
  }
  return 0;
}


extern "C" int bsde_wsde_bwd(const bsde_wsde_desc* d, const bsde_fw_params* fw, const float* d_tk,
                             const float* d_dtk, const float* d_wtraj, const float* d_z1,
                             const float* d_gw_traj, const float* d_gkl, const float* d_gwT, float* d_gw0,
                             const bsde_fw_params* gfw, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = common_checks(d, fw)) return e;
  BSDE_CHECK_ARG(gfw && d_tk && d_dtk && d_wtraj && d_z1 && d_gkl && d_gw0, "NULL tensor pointer");
  WsdeArgs A;
  memset(&A, 0, sizeof(A));
  if (int e = fill_tiny(fw, A.td)) return e;
  const size_t need = ws_floats(d->S, fw->P, true) * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    set_error("wsde_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  const size_t smem = smem_bytes(A.td, true);
  int G = 0;
  if (int e = grid_for(fw->P, true, smem, G)) return e;
  if (G > MAXG) G = MAXG;
  const int S = d->S, P = fw->P, nit = d->nit, d1 = fw->dims[0];
  for (int s0 = 0; s0 < S; s0 += SC) {
    const int sc = S - s0 < SC ? S - s0 : SC;
    A.d = *d; A.d.S = sc; A.fw = *fw; A.gfw = *gfw; A.ou_coef = d->ou_coef; A.s_base = s0 + (int)d->sample_offset;
    A.tk = d_tk; A.dtk = d_dtk;
    A.wtraj = const_cast<float*>(d_wtraj) + (size_t)s0 * (nit + 1) * P;
    A.z1 = const_cast<float*>(d_z1) + (size_t)s0 * nit * d1;
    A.gw_traj = d_gw_traj ? d_gw_traj + (size_t)s0 * nit * P : nullptr;
    A.gkl = d_gkl + s0;
    A.gwT = d_gwT ? d_gwT + (size_t)s0 * P : nullptr;
    A.gw0 = d_gw0 + (d->w0_per_sample ? (size_t)s0 * P : 0);
    A.accumulate = s0 > 0;
    float* ws = (float*)workspace;
    A.partials = ws; ws += (size_t)2 * MAXG * SC * ME;
    A.klpart = ws; ws += (size_t)MAXG * SC;
    A.stpart = ws; ws += (size_t)MAXG * ME;
    A.glam = ws; ws += (size_t)sc * P;
    A.ghold = ws;
    void* args[] = {&A};
    BSDE_CUDA_OK(cudaLaunchCooperativeKernel((void*)wsde_bwd_kernel, dim3(G), dim3(NTHREADS), args, smem,
                                             (cudaStream_t)stream));
    count_launch();
  }
  return 0;
}

