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

#include <algorithm>
#include <cstdlib>
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
  int ept;     // RES kernels: resident elements per thread (ceil(P / (G * NTHREADS)))
  int stamp;   // tools only
  int tp_floats;  // > 0: the tiny MLP's parameters are staged in shared memory (that many floats)
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

// sum partials over CTAs in a fixed order: out[v] = sum_g part[g*stride + v], v < nv <= 32.  Thread (v = tid & 31, c = tid >> 5)
// adds the partials of CTAs c, c + NWARPS, ... (independent, coalesced loads: one L2 round trip), then NWARPS chunk sums are added
// in order -- identical on every CTA.  `scratch` = NWARPS * 32 floats of shared memory.
__device__ __forceinline__ void gather_partials(const float* part, int G, int stride, int nv, float* out, float* scratch) {
  const int v = threadIdx.x & 31, c = threadIdx.x >> 5;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
  if (v < nv) {
    const float* q = part + v;
    int g = c;
    for (; g + 3 * NWARPS < G; g += 4 * NWARPS) {
      const float a0 = __ldcg(q + (long long)g * stride), a1 = __ldcg(q + (long long)(g + NWARPS) * stride);
      const float a2 = __ldcg(q + (long long)(g + 2 * NWARPS) * stride), a3 = __ldcg(q + (long long)(g + 3 * NWARPS) * stride);
      r0 += a0; r1 += a1; r2 += a2; r3 += a3;
    }
    for (; g < G; g += NWARPS) r0 += __ldcg(q + (long long)g * stride);
  }
  scratch[c * 32 + v] = (r0 + r1) + (r2 + r3);
  __syncthreads();
  if (threadIdx.x < nv) {
    float t = 0.f;
#pragma unroll
    for (int c2 = 0; c2 < NWARPS; ++c2) t += scratch[c2 * 32 + threadIdx.x];
    out[threadIdx.x] = t;
  }
  __syncthreads();
}

// ---- the tiny middle of f_w, per sample row.  Stores a (pre-gate), g (gate), y (pre-activation)
// and h (activation) for every hidden layer so the reverse pass can reuse them. ----------------
struct TinySmem {
  float* a;    // [SC+1][wsum]
  float* g;    // [wsum]   gate sigmoid(w_t t + w_tb) (sample-independent)
  float* y;    // [SC+1][wsum]
  float* h;    // [SC+1][wsum]
  float* btt;  // [wsum]   shift b_t * t
};

constexpr int TL = BSDE_FW_MAX_MID + 1;   // hidden layers of the tiny MLP
constexpr int TP_MAX = 6144;              // floats of tiny parameters staged in shared memory once per launch (else: global)
// b, w_t, w_tb, b_t of hidden layer l (dims[l] each) and, for l >= 1, W [dims[l-1]][dims[l]]: shared-memory copies or the originals
struct TinyPar {
  const float* b[TL]; const float* wt[TL]; const float* wtb[TL]; const float* bt[TL]; const float* W[TL];
  const float* beta[TL];   // trainable swish (arch.py:14-29): beta of the activation after hidden layer l, else unused
};
struct TinyGrad {
  float* b[TL]; float* wt[TL]; float* wtb[TL]; float* bt[TL]; float* W[TL]; float* beta[TL];
};
// five vectors per hidden layer (b, w_t, w_tb, b_t, beta) + the middle matrices
__host__ __device__ inline int tiny_par_floats(const TinyDims& td) {
  int f = 0;
  for (int l = 0; l < td.nhid; ++l) f += 5 * td.dims[l] + (l > 0 ? td.dims[l - 1] * td.dims[l] : 0);
  return f;
}
// activation of a hidden unit and its derivative in the pre-activation (swish: per-unit beta)
__device__ __forceinline__ float tiny_act(int act, float y, float beta) {
  if (act == BSDE_ACT_SWISH) return y * sigmoidf_(y * (fmaxf(beta, 0.f) + log1pf(expf(-fabsf(beta)))));
  return act_fwd(act, y);
}
__device__ __forceinline__ float tiny_act_grad(int act, float y, float beta) {
  if (act == BSDE_ACT_SWISH) {
    const float g = fmaxf(beta, 0.f) + log1pf(expf(-fabsf(beta)));
    const float sg = sigmoidf_(y * g);
    return sg + y * g * sg * (1.f - sg);
  }
  return act_grad(act, y);
}
// d act / d beta of the trainable swish: y^2 sig (1 - sig) sigmoid(beta)
__device__ __forceinline__ float tiny_act_dbeta(float y, float beta) {
  const float g = fmaxf(beta, 0.f) + log1pf(expf(-fabsf(beta)));
  const float sg = sigmoidf_(y * g);
  return y * y * sg * (1.f - sg) * sigmoidf_(beta);
}
__device__ __forceinline__ void tiny_par_init(const bsde_fw_params& fw, const TinyDims& td, float* stage, TinyPar& TP) {
  for (int l = 0; l < td.nhid; ++l) {
    TP.b[l] = l == 0 ? fw.b1 : fw.bm[l - 1];
    TP.wt[l] = l == 0 ? fw.wt1 : fw.wtm[l - 1];
    TP.wtb[l] = l == 0 ? fw.wtb1 : fw.wtbm[l - 1];
    TP.bt[l] = l == 0 ? fw.bt1 : fw.btm[l - 1];
    TP.W[l] = l == 0 ? nullptr : fw.Wm[l - 1];
    TP.beta[l] = l == 0 ? fw.beta1 : fw.betam[l - 1];
  }
  if (!stage) return;
  float* q = stage;
  for (int l = 0; l < td.nhid; ++l) {
    const int d = td.dims[l];
    const float* src[5] = {TP.b[l], TP.wt[l], TP.wtb[l], TP.bt[l], TP.beta[l]};
    for (int f = 0; f < 5; ++f) {
      const bool used = f == 0 || (f < 4 && !fw.plain) || (f == 4 && fw.act == BSDE_ACT_SWISH);
      if (used)
        for (int j = threadIdx.x; j < d; j += NTHREADS) q[j] = __ldg(src[f] + j);
      if (f == 0) TP.b[l] = q; else if (f == 1) TP.wt[l] = q; else if (f == 2) TP.wtb[l] = q; else if (f == 3) TP.bt[l] = q; else TP.beta[l] = q;
      q += d;
    }
    if (l > 0) {
      const int n = td.dims[l - 1] * d;
      for (int j = threadIdx.x; j < n; j += NTHREADS) q[j] = __ldg(TP.W[l] + j);
      TP.W[l] = q;
      q += n;
    }
  }
  __syncthreads();
}

// gates and shifts of every hidden unit at time t (sample-independent; no dependence on the state): callers run it BEFORE
// the partials gather so that it overlaps that round trip; a later __syncthreads publishes it
__device__ __forceinline__ void tiny_gates(const bsde_fw_params& fw, const TinyDims& td, const TinyPar& TP, const TinySmem& T, float t) {
  for (int idx = threadIdx.x; idx < td.wsum; idx += NTHREADS) {
    int l = 0;
    while (l + 1 < td.nhid && idx >= td.off[l + 1]) ++l;
    const int j = idx - td.off[l];
    T.g[idx] = fw.plain ? 1.f : sigmoidf_(TP.wt[l][j] * t + TP.wtb[l][j]);
    T.btt[idx] = fw.plain ? 0.f : TP.bt[l][j] * t;
  }
}

__device__ __forceinline__ void tiny_forward(const WsdeArgs& A, const TinyPar& TP, const TinySmem& T, const float* z1raw /*[S][d1] smem*/,
                                             int S) {
  const TinyDims& td = A.td;
  const int wsum = td.wsum;
  // layer 0 (first hidden): a = z1 + b1
  {
    const int d1 = td.dims[0];
    for (int idx = threadIdx.x; idx < S * d1; idx += NTHREADS) {
      const int s = idx / d1, j = idx - s * d1;
      const float a = z1raw[s * ME + j] + TP.b[0][j];
      const float y = a * T.g[j] + T.btt[j];
      T.a[s * wsum + j] = a;
      T.y[s * wsum + j] = y;
      T.h[s * wsum + j] = tiny_act(A.fw.act, y, A.fw.act == BSDE_ACT_SWISH ? TP.beta[0][j] : 0.f);
    }
    __syncthreads();
  }
  for (int l = 0; l + 1 < td.nhid; ++l) {
    const int din = td.dims[l], dout = td.dims[l + 1];
    const int oi = td.off[l], oo = td.off[l + 1];
    const float* W = TP.W[l + 1];
    for (int idx = threadIdx.x; idx < S * dout; idx += NTHREADS) {
      const int s = idx / dout, j = idx - s * dout;
      float a0 = TP.b[l + 1][j], a1 = 0.f;
      int q = 0;
      for (; q + 1 < din; q += 2) {
        a0 = fmaf(T.h[s * wsum + oi + q], W[q * dout + j], a0);
        a1 = fmaf(T.h[s * wsum + oi + q + 1], W[(q + 1) * dout + j], a1);
      }
      if (q < din) a0 = fmaf(T.h[s * wsum + oi + q], W[q * dout + j], a0);
      const float a = a0 + a1;
      const float y = a * T.g[oo + j] + T.btt[oo + j];
      T.a[s * wsum + oo + j] = a;
      T.y[s * wsum + oo + j] = y;
      T.h[s * wsum + oo + j] = tiny_act(A.fw.act, y, A.fw.act == BSDE_ACT_SWISH ? TP.beta[l + 1][j] : 0.f);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float inv_or_zero(float s) { return s != 0.f ? 1.f / s : 0.f; }

// =============================================================================================
// forward
// =============================================================================================
// Per-thread resident storage in shared memory (RES kernels): slot (field f, element e) of thread tid lives at
// base[(f * ept + e) * NTHREADS + tid] -- consecutive threads hit consecutive banks.  A thread owns the elements
// p = gtid + e * NT, e < ept, for the whole launch, so its w state, the f_w parameter vectors of its elements and (reverse
// pass) its adjoint and gradient accumulators stay on chip across all scan iterations (north star (a)).
// tools-only phase stamps (BSDE_K1_STAMP=1; read back with bsde_debug_k1_stamps): %globaltimer of CTA 0 / thread 0 at the
// phase boundaries of every forward / reverse iteration
constexpr int STAMP_PH = 8, STAMP_IT = 64;
__device__ unsigned long long g_k1_stamps[2][STAMP_IT][STAMP_PH];
__device__ __forceinline__ void k1_stamp(int on, int which, int k, int ph) {
  if (on && blockIdx.x == 0 && threadIdx.x == 0 && k < STAMP_IT) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_k1_stamps[which][k][ph] = t;
  }
}

struct Slots {
  float* base;  // already offset by threadIdx.x
  int ept;
  __device__ __forceinline__ float& operator()(int f, int e) const { return base[(f * ept + e) * NTHREADS]; }
};
// field indices
constexpr int F_W1 = 0, F_W4 = ME, F_B4 = 2 * ME, F_WT4 = 2 * ME + 1, F_WTB4 = 2 * ME + 2, F_BT4 = 2 * ME + 3;
constexpr int F_NPAR = 2 * ME + 4;
constexpr int FF_STALE = F_NPAR, FF_W = F_NPAR + 1, FF_N = FF_W + SC;                              // forward
constexpr int FB_GW1 = F_NPAR, FB_GW4 = FB_GW1 + ME, FB_GB4 = FB_GW4 + ME, FB_LAM = FB_GB4 + 4, FB_N = FB_LAM + SC;  // reverse
constexpr int MAXE = 6;   // resident elements per thread; larger P runs the streaming instantiation

template <bool RES>
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
  T.btt = T.h + (SC + 1) * wsum;
  float* sh_z1 = T.btt + wsum;            // [(SC+1)][ME]  (row SC = stale vector)
  float* sh_red = sh_z1 + (SC + 1) * ME;  // [NWARPS][SC*ME]
  float* sh_hst = sh_red + NWARPS * SC * ME;  // [ME] hL of the stale vector
  float* sh_stage = sh_hst + ME;          // [tp_floats] staged tiny parameters
  Slots R;
  R.base = sh_stage + A.tp_floats + threadIdx.x;
  R.ept = A.ept;
  TinyPar TP;
  tiny_par_init(A.fw, A.td, A.tp_floats ? sh_stage : nullptr, TP);

  const float sigma = A.d.sigma, inv_sigma = inv_or_zero(sigma);
  const long long wstride = (long long)(nit + 1) * P;
  const uint32_t sg0 = (uint32_t)A.s_base >> 2, sph = (uint32_t)A.s_base & 3u;  // sample group / phase of local sample 0

  // ---- phase 0: w_0 -> trajectory slot 0, partial dots of w_0 (and of w_stale); RES: parameters and state -> slots
  {
    float acc[SC * ME];
#pragma unroll
    for (int i = 0; i < SC * ME; ++i) acc[i] = 0.f;
    float accst[ME];
#pragma unroll
    for (int j = 0; j < ME; ++j) accst[j] = 0.f;
    int e = 0;
    for (long long p = gtid; p < P; p += NT, ++e) {
      float w1[ME];
#pragma unroll
      for (int j = 0; j < ME; ++j) w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f;
      if (RES) {
#pragma unroll
        for (int j = 0; j < ME; ++j) {
          R(F_W1 + j, e) = w1[j];
          R(F_W4 + j, e) = j < dL ? __ldg(A.fw.W4 + (long long)j * P + p) : 0.f;
        }
        R(F_B4, e) = __ldg(A.fw.b4 + p);
        R(F_WT4, e) = A.fw.plain ? 0.f : __ldg(A.fw.wt4 + p);
        R(F_WTB4, e) = A.fw.plain ? 0.f : __ldg(A.fw.wtb4 + p);
        R(F_BT4, e) = A.fw.plain ? 0.f : __ldg(A.fw.bt4 + p);
        R(FF_STALE, e) = A.d.stl ? __ldg(A.wstale + p) : 0.f;
      }
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s < S) {
          const float w = __ldg(A.w0 + (A.d.w0_per_sample ? (long long)s * P : 0) + p);
          A.wtraj[s * wstride + p] = w;
          if (RES) R(FF_W + s, e) = w;
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
  if (A.d.stl) gather_partials(A.stpart, G, ME, ME, sh_z1 + SC * ME, sh_red);

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
    k1_stamp(A.stamp, 0, k, 0);
    tiny_gates(A.fw, A.td, TP, T, t);   // independent of the partials: overlaps their round trip
    gather_partials(pin, G, SC * ME, SC * ME, sh_z1, sh_red);
    k1_stamp(A.stamp, 0, k, 1);
    if (blockIdx.x == 0) {
      for (int idx = threadIdx.x; idx < S * d1; idx += NTHREADS) {
        const int s = idx / d1, j = idx - s * d1;
        A.z1[((long long)k * S + s) * d1 + j] = sh_z1[s * ME + j];
      }
    }
    // tiny MLP for the S samples (+ the stale vector as row SC when stl)
    tiny_forward(A, TP, T, sh_z1, A.d.stl ? SC + 1 : S);
    if (A.d.stl && threadIdx.x < ME) sh_hst[threadIdx.x] = threadIdx.x < dL ? T.h[SC * wsum + offL + threadIdx.x] : 0.f;
    __syncthreads();
    k1_stamp(A.stamp, 0, k, 2);

    float acc[SC * ME];
#pragma unroll
    for (int i = 0; i < SC * ME; ++i) acc[i] = 0.f;
    // last hidden activations of the samples: registers for the whole pass over P
    float hL[SC][ME];
#pragma unroll
    for (int s = 0; s < SC; ++s)
#pragma unroll
      for (int j = 0; j < ME; ++j) hL[s][j] = (s < S && j < dL) ? T.h[s * wsum + offL + j] : 0.f;
    int e = 0;
    for (long long p = gtid; p < P; p += NT, ++e) {
      float w1[ME], w4[ME], b4, bt4, s4, wsv = 0.f;
      if (RES) {
#pragma unroll
        for (int j = 0; j < ME; ++j) { w1[j] = R(F_W1 + j, e); w4[j] = R(F_W4 + j, e); }
        b4 = R(F_B4, e); bt4 = R(F_BT4, e);
        s4 = A.fw.plain ? 1.f : sigmoidf_(R(F_WT4, e) * t + R(F_WTB4, e));
        if (A.d.stl) wsv = R(FF_STALE, e);
      } else {
#pragma unroll
        for (int j = 0; j < ME; ++j) {
          w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f;
          w4[j] = j < dL ? __ldg(A.fw.W4 + (long long)j * P + p) : 0.f;
        }
        b4 = __ldg(A.fw.b4 + p); bt4 = A.fw.plain ? 0.f : __ldg(A.fw.bt4 + p);
        s4 = A.fw.plain ? 1.f : sigmoidf_(__ldg(A.fw.wt4 + p) * t + __ldg(A.fw.wtb4 + p));
        if (A.d.stl) wsv = __ldg(A.wstale + p);
      }
      float ust = 0.f;
      if (A.d.stl) {
        float pre = b4;
#pragma unroll
        for (int j = 0; j < ME; ++j) pre = fmaf(sh_hst[j], w4[j], pre);
        ust = (pre * s4 + bt4 * t + (A.ou_coef + 1.f) * wsv) * inv_sigma;
      }
      float z4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s < S) {
          const float w = RES ? R(FF_W + s, e) : A.wtraj[s * wstride + (long long)k * P + p];
          float pre = b4;
#pragma unroll
          for (int j = 0; j < ME; ++j) pre = fmaf(hL[s][j], w4[j], pre);
          const float mlp = pre * s4 + bt4 * t;
          const float dw = mlp + A.ou_coef * w;
          const float u = (mlp + (A.ou_coef + 1.f) * w) * inv_sigma;
          float dWv;
          if (A.dW) {
            dWv = __ldg(A.dW + ((long long)s * nit + k) * P + p);
          } else {
            const uint32_t sl = (uint32_t)s + sph;   // position inside the global groups of four samples
            if (s == 0 || (sl & 3u) == 0u) philox_normal4(A.d.seed, (uint32_t)p, nidx, sg0 + (sl >> 2), A.d.stream_id, z4);
            const uint32_t zc = sl & 3u;
            dWv = sq * (zc == 0u ? z4[0] : zc == 1u ? z4[1] : zc == 2u ? z4[2] : z4[3]);
          }
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
          if (RES) R(FF_W + s, e) = wn;
#pragma unroll
          for (int j = 0; j < ME; ++j) acc[s * ME + j] = fmaf(wn, w1[j], acc[s * ME + j]);
        }
      }
    }
    k1_stamp(A.stamp, 0, k, 3);
    if (k + 1 < nit) {
      block_reduce_store<SC * ME>(acc, sh_red, pout + (long long)blockIdx.x * SC * ME, SC * ME);
      k1_stamp(A.stamp, 0, k, 4);
      grid.sync();
      k1_stamp(A.stamp, 0, k, 5);
    }
  }
  // ---- kl: block partials, then CTA 0 sums in order
  block_reduce_store<SC>(klacc, sh_red, A.klpart + (long long)blockIdx.x * SC, SC);
  grid.sync();
  if (blockIdx.x == 0) {
    gather_partials(A.klpart, G, SC, SC, sh_z1, sh_red);
    if (threadIdx.x < S) A.kl[threadIdx.x] = sh_z1[threadIdx.x];
  }
}

// =============================================================================================
// reverse pass
// =============================================================================================
template <bool RES>
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
  T.btt = T.h + (SC + 1) * wsum;
  float* sh_z1 = T.btt + wsum;                // [(SC+1)][ME]
  float* sh_red = sh_z1 + (SC + 1) * ME;      // [NWARPS][SC*ME]
  float* sh_gh = sh_red + NWARPS * SC * ME;   // [SC][wsum]  dL/dh (then reused as dL/da)
  float* sh_dz1 = sh_gh + SC * wsum;          // [SC][ME]    pending dL/dz1 of iteration k+1
  float* sh_ghL = sh_dz1 + SC * ME;           // [SC][ME]
  float* sh_stage = sh_ghL + SC * ME;         // [2 * tp_floats] staged tiny parameters, then CTA 0's tiny gradient accumulators
  Slots R;
  R.base = sh_stage + 2 * A.tp_floats + threadIdx.x;
  R.ept = A.ept;
  TinyPar TP;
  tiny_par_init(A.fw, A.td, A.tp_floats ? sh_stage : nullptr, TP);
  // tiny parameter gradients: only CTA 0 accumulates them (every CTA computes the same tiny reverse pass); staged: in shared
  // memory across all iterations, written once in the epilogue -- else read-modify-write in global memory every iteration
  TinyGrad TG;
  {
    float* q = sh_stage + A.tp_floats;
    for (int l = 0; l < td.nhid; ++l) {
      const int dd = td.dims[l];
      float* gsrc[6] = {l == 0 ? A.gfw.b1 : A.gfw.bm[l - 1], l == 0 ? A.gfw.wt1 : A.gfw.wtm[l - 1], l == 0 ? A.gfw.wtb1 : A.gfw.wtbm[l - 1],
                        l == 0 ? A.gfw.bt1 : A.gfw.btm[l - 1], l == 0 ? A.gfw.beta1 : A.gfw.betam[l - 1], l == 0 ? nullptr : A.gfw.Wm[l - 1]};
      const int nW = l > 0 ? td.dims[l - 1] * dd : 0;
      const bool sw = A.fw.act == BSDE_ACT_SWISH;
      if (A.tp_floats) {
        const bool acc0 = A.accumulate != 0;
        for (int f = 0; f < 5; ++f) {
          const bool used = f == 0 || (f < 4 && !A.fw.plain) || (f == 4 && sw);
          for (int j = threadIdx.x; j < dd; j += NTHREADS) q[f * dd + j] = (acc0 && used) ? gsrc[f][j] : 0.f;
        }
        for (int j = threadIdx.x; j < nW; j += NTHREADS) q[5 * dd + j] = acc0 ? gsrc[5][j] : 0.f;
        TG.b[l] = q; TG.wt[l] = q + dd; TG.wtb[l] = q + 2 * dd; TG.bt[l] = q + 3 * dd; TG.beta[l] = q + 4 * dd; TG.W[l] = q + 5 * dd;
        q += 5 * dd + nW;
      } else {
        TG.b[l] = gsrc[0]; TG.wt[l] = gsrc[1]; TG.wtb[l] = gsrc[2]; TG.bt[l] = gsrc[3]; TG.beta[l] = gsrc[4]; TG.W[l] = gsrc[5];
      }
    }
    __syncthreads();
  }

  const float sigma = A.d.sigma, inv_sigma = inv_or_zero(sigma);
  const long long wstride = (long long)(nit + 1) * P;
  const bool acc_in = A.accumulate != 0;

  // ---- prologue: zero the per-element gradient accumulators, seed lambda; RES: parameters / accumulators / lambda -> slots
  {
    int e = 0;
    for (long long p = gtid; p < P; p += NT, ++e) {
      if (RES) {
#pragma unroll
        for (int j = 0; j < ME; ++j) {
          R(F_W1 + j, e) = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f;
          R(F_W4 + j, e) = j < dL ? __ldg(A.fw.W4 + (long long)j * P + p) : 0.f;
          R(FB_GW1 + j, e) = (acc_in && j < d1) ? A.gfw.W1[p * d1 + j] : 0.f;
          R(FB_GW4 + j, e) = (acc_in && j < dL) ? A.gfw.W4[(long long)j * P + p] : 0.f;
        }
        R(F_B4, e) = __ldg(A.fw.b4 + p);
        R(F_WT4, e) = A.fw.plain ? 0.f : __ldg(A.fw.wt4 + p);
        R(F_WTB4, e) = A.fw.plain ? 0.f : __ldg(A.fw.wtb4 + p);
        R(F_BT4, e) = A.fw.plain ? 0.f : __ldg(A.fw.bt4 + p);
        R(FB_GB4 + 0, e) = acc_in ? A.gfw.b4[p] : 0.f;
        R(FB_GB4 + 1, e) = (acc_in && !A.fw.plain) ? A.gfw.wt4[p] : 0.f;
        R(FB_GB4 + 2, e) = (acc_in && !A.fw.plain) ? A.gfw.wtb4[p] : 0.f;
        R(FB_GB4 + 3, e) = (acc_in && !A.fw.plain) ? A.gfw.bt4[p] : 0.f;
        for (int s = 0; s < S; ++s) R(FB_LAM + s, e) = A.gwT ? __ldg(A.gwT + (long long)s * P + p) : 0.f;
      } else {
        if (!acc_in) {
          for (int j = 0; j < d1; ++j) A.gfw.W1[p * d1 + j] = 0.f;
          for (int j = 0; j < dL; ++j) A.gfw.W4[(long long)j * P + p] = 0.f;
          A.gfw.b4[p] = 0.f;
          if (!A.fw.plain) { A.gfw.wt4[p] = 0.f; A.gfw.wtb4[p] = 0.f; A.gfw.bt4[p] = 0.f; }
        }
        for (int s = 0; s < S; ++s) A.glam[(long long)s * P + p] = A.gwT ? __ldg(A.gwT + (long long)s * P + p) : 0.f;
      }
    }
  }
  if (blockIdx.x == 0 && !acc_in && !A.tp_floats) {
    for (int j = threadIdx.x; j < d1; j += NTHREADS) {
      A.gfw.b1[j] = 0.f;
      if (!A.fw.plain) { A.gfw.wt1[j] = 0.f; A.gfw.wtb1[j] = 0.f; A.gfw.bt1[j] = 0.f; }
      if (A.fw.act == BSDE_ACT_SWISH) A.gfw.beta1[j] = 0.f;
    }
    for (int l = 0; l + 1 < td.nhid; ++l) {
      const int din = td.dims[l], dout = td.dims[l + 1];
      for (int i = threadIdx.x; i < din * dout; i += NTHREADS) A.gfw.Wm[l][i] = 0.f;
      for (int j = threadIdx.x; j < dout; j += NTHREADS) {
        A.gfw.bm[l][j] = 0.f;
        if (!A.fw.plain) { A.gfw.wtm[l][j] = 0.f; A.gfw.wtbm[l][j] = 0.f; A.gfw.btm[l][j] = 0.f; }
        if (A.fw.act == BSDE_ACT_SWISH) A.gfw.betam[l][j] = 0.f;
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
    k1_stamp(A.stamp, 1, k, 0);
    // the next iteration's streamed operands (w_{k-1}, the x rows' dL/dw_{k-1}) -> L2 while this iteration computes: the pass over
    // P is otherwise bound by the DRAM latency of these loads (8 warps per SM cannot hide it)
    if (k > 0 && ((threadIdx.x & 31) == 0 || (threadIdx.x & 31) == 31)) {   // a warp's 32 consecutive floats span <= 2 lines
      for (long long p = gtid; p < P; p += NT) {
        for (int s2 = 0; s2 < S; ++s2) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(A.wtraj + s2 * wstride + (long long)(k - 1) * P + p));
          if (A.gw_traj) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.gw_traj + ((long long)s2 * nit + (k - 1)) * P + p));
        }
      }
    }
    // tiny forward of iteration k from the stored first-layer dots
    for (int idx = threadIdx.x; idx < SC * ME; idx += NTHREADS) {
      const int s = idx / ME, j = idx - s * ME;
      sh_z1[idx] = (s < S && j < d1) ? __ldg(A.z1 + ((long long)k * S + s) * d1 + j) : 0.f;
    }
    tiny_gates(A.fw, A.td, TP, T, t);
    __syncthreads();
    tiny_forward(A, TP, T, sh_z1, S);
    k1_stamp(A.stamp, 1, k, 1);

    float part[SC * ME];
#pragma unroll
    for (int i = 0; i < SC * ME; ++i) part[i] = 0.f;
    float hL[SC][ME];
#pragma unroll
    for (int s = 0; s < SC; ++s)
#pragma unroll
      for (int j = 0; j < ME; ++j) hL[s][j] = (s < S && j < dL) ? T.h[s * wsum + offL + j] : 0.f;
    int e = 0;
    for (long long p = gtid; p < P; p += NT, ++e) {
      float w1[ME], w4[ME], gw1[ME], gw4[ME];
      float b4, bt4, s4;
      if (RES) {
#pragma unroll
        for (int j = 0; j < ME; ++j) { w1[j] = R(F_W1 + j, e); w4[j] = R(F_W4 + j, e); gw1[j] = 0.f; gw4[j] = 0.f; }
        b4 = R(F_B4, e); bt4 = R(F_BT4, e);
        s4 = A.fw.plain ? 1.f : sigmoidf_(R(F_WT4, e) * t + R(F_WTB4, e));
      } else {
#pragma unroll
        for (int j = 0; j < ME; ++j) {
          w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f;
          w4[j] = j < dL ? __ldg(A.fw.W4 + (long long)j * P + p) : 0.f;
          gw1[j] = 0.f; gw4[j] = 0.f;
        }
        b4 = __ldg(A.fw.b4 + p); bt4 = A.fw.plain ? 0.f : __ldg(A.fw.bt4 + p);
        s4 = A.fw.plain ? 1.f : sigmoidf_(__ldg(A.fw.wt4 + p) * t + __ldg(A.fw.wtb4 + p));
      }
      float gb4 = 0.f, gwt4 = 0.f, gwtb4 = 0.f, gbt4 = 0.f;
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s < S) {
          float lam = RES ? R(FB_LAM + s, e) : A.glam[(long long)s * P + p];
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
          for (int j = 0; j < ME; ++j) pre = fmaf(hL[s][j], w4[j], pre);
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
          if (RES) R(FB_LAM + s, e) = gl; else A.glam[(long long)s * P + p] = gl;
          const float qs = q * s4;
#pragma unroll
          for (int j = 0; j < ME; ++j) {
            part[s * ME + j] = fmaf(w4[j], qs, part[s * ME + j]);
            gw4[j] = fmaf(hL[s][j], qs, gw4[j]);
          }
          gb4 += qs;
          const float ee = q * pre * s4 * (1.f - s4);
          gwt4 = fmaf(ee, t, gwt4);
          gwtb4 += ee;
          gbt4 = fmaf(q, t, gbt4);
        }
      }
      if (RES) {
#pragma unroll
        for (int j = 0; j < ME; ++j) { if (pending) R(FB_GW1 + j, e) += gw1[j]; R(FB_GW4 + j, e) += gw4[j]; }
        R(FB_GB4 + 0, e) += gb4; R(FB_GB4 + 1, e) += gwt4; R(FB_GB4 + 2, e) += gwtb4; R(FB_GB4 + 3, e) += gbt4;
      } else {
        for (int j = 0; j < d1; ++j) if (pending) A.gfw.W1[p * d1 + j] += gw1[j];
        for (int j = 0; j < dL; ++j) A.gfw.W4[(long long)j * P + p] += gw4[j];
        A.gfw.b4[p] += gb4;
        if (!A.fw.plain) { A.gfw.wt4[p] += gwt4; A.gfw.wtb4[p] += gwtb4; A.gfw.bt4[p] += gbt4; }
      }
    }
    k1_stamp(A.stamp, 1, k, 2);
    float* pout = A.partials + (long long)(k & 1) * G * SC * ME;
    block_reduce_store<SC * ME>(part, sh_red, pout + (long long)blockIdx.x * SC * ME, SC * ME);
    k1_stamp(A.stamp, 1, k, 3);
    grid.sync();
    k1_stamp(A.stamp, 1, k, 4);
    gather_partials(pout, G, SC * ME, SC * ME, sh_ghL, sh_red);
    k1_stamp(A.stamp, 1, k, 5);

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
        const float bj = A.fw.act == BSDE_ACT_SWISH ? TP.beta[l][j] : 0.f;
        const float gh = sh_gh[s * wsum + oo + j];
        const float gy = gh * tiny_act_grad(A.fw.act, T.y[s * wsum + oo + j], bj);
        // trainable swish: dL/dbeta needs dL/dh, which gy overwrites -> stash it in the (now free) activation slot
        if (A.fw.act == BSDE_ACT_SWISH) T.h[s * wsum + oo + j] = gh * tiny_act_dbeta(T.y[s * wsum + oo + j], bj);
        sh_gh[s * wsum + oo + j] = gy;  // keep gy for the parameter grads below
      }
      __syncthreads();
      if (blockIdx.x == 0) {
        float* gb = TG.b[l];
        float* gwt = TG.wt[l];
        float* gwtb = TG.wtb[l];
        float* gbt = TG.bt[l];
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
          if (A.fw.act == BSDE_ACT_SWISH) {
            float sbeta = 0.f;
            for (int s = 0; s < S; ++s) sbeta += T.h[s * wsum + oo + j];
            TG.beta[l][j] += sbeta;
          }
        }
        if (l > 0) {
          const int din = td.dims[l - 1], oi = td.off[l - 1];
          float* gW = TG.W[l];
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
        const float* W = TP.W[l];
        for (int idx = threadIdx.x; idx < S * din; idx += NTHREADS) {
          const int s = idx / din, qi = idx - s * din;
          float r0 = 0.f, r1 = 0.f;
          int j = 0;
          for (; j + 1 < dout; j += 2) {
            r0 = fmaf(W[qi * dout + j], sh_gh[s * wsum + oo + j] * T.g[oo + j], r0);
            r1 = fmaf(W[qi * dout + j + 1], sh_gh[s * wsum + oo + j + 1] * T.g[oo + j + 1], r1);
          }
          if (j < dout) r0 = fmaf(W[qi * dout + j], sh_gh[s * wsum + oo + j] * T.g[oo + j], r0);
          sh_gh[s * wsum + oi + qi] = r0 + r1;
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
    k1_stamp(A.stamp, 1, k, 6);
  }

  // ---- epilogue: close lambda_0 with the first-layer term of iteration 0; RES: accumulators -> global, once
  {
    int e = 0;
    for (long long p = gtid; p < P; p += NT, ++e) {
      float w1[ME], gw1[ME];
#pragma unroll
      for (int j = 0; j < ME; ++j) { w1[j] = j < d1 ? __ldg(A.fw.W1 + p * d1 + j) : 0.f; gw1[j] = 0.f; }
      float tot = 0.f;
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        if (s < S) {
          float lam = RES ? R(FB_LAM + s, e) : A.glam[(long long)s * P + p];
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
      if (RES) {
        for (int j = 0; j < d1; ++j) A.gfw.W1[p * d1 + j] = R(FB_GW1 + j, e) + gw1[j];
        for (int j = 0; j < dL; ++j) A.gfw.W4[(long long)j * P + p] = R(FB_GW4 + j, e);
        A.gfw.b4[p] = R(FB_GB4 + 0, e);
        if (!A.fw.plain) { A.gfw.wt4[p] = R(FB_GB4 + 1, e); A.gfw.wtb4[p] = R(FB_GB4 + 2, e); A.gfw.bt4[p] = R(FB_GB4 + 3, e); }
      } else {
        for (int j = 0; j < d1; ++j) A.gfw.W1[p * d1 + j] += gw1[j];
      }
      if (!A.d.w0_per_sample) A.gw0[p] = acc_in ? A.gw0[p] + tot : tot;
    }
  }
  if (A.tp_floats && blockIdx.x == 0) {   // staged tiny gradient accumulators -> global, once
    __syncthreads();
    for (int l = 0; l < td.nhid; ++l) {
      const int dd = td.dims[l];
      float* gdst[5] = {l == 0 ? A.gfw.b1 : A.gfw.bm[l - 1], l == 0 ? A.gfw.wt1 : A.gfw.wtm[l - 1], l == 0 ? A.gfw.wtb1 : A.gfw.wtbm[l - 1],
                        l == 0 ? A.gfw.bt1 : A.gfw.btm[l - 1], l == 0 ? nullptr : A.gfw.Wm[l - 1]};
      float* gbeta = l == 0 ? A.gfw.beta1 : A.gfw.betam[l - 1];
      for (int j = threadIdx.x; j < dd; j += NTHREADS) {
        gdst[0][j] = TG.b[l][j];
        if (!A.fw.plain) { gdst[1][j] = TG.wt[l][j]; gdst[2][j] = TG.wtb[l][j]; gdst[3][j] = TG.bt[l][j]; }
        if (A.fw.act == BSDE_ACT_SWISH) gbeta[j] = TG.beta[l][j];
      }
      if (l > 0)
        for (int j = threadIdx.x; j < td.dims[l - 1] * dd; j += NTHREADS) gdst[4][j] = TG.W[l][j];
    }
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
  BSDE_CHECK_ARG(fw->act >= BSDE_ACT_SOFTPLUS && fw->act <= BSDE_ACT_SWISH && fw->act != BSDE_ACT_NONE, "bad fw activation %d", fw->act);
  if (fw->act == BSDE_ACT_SWISH) {
    BSDE_CHECK_ARG(fw->beta1 != nullptr, "swish f_w needs beta1");
    for (int i = 0; i + 1 < fw->nhid; ++i) BSDE_CHECK_ARG(fw->betam[i] != nullptr, "swish f_w needs betam[%d]", i);
  }
  return 0;
}

static int staged_floats(const TinyDims& td) {
  const int f = tiny_par_floats(td);
  return f <= TP_MAX ? f : 0;
}

static size_t smem_bytes(const TinyDims& td, bool bwd, int ept = 0) {
  size_t f = (size_t)(3 * (SC + 1) + 2) * td.wsum + (SC + 1) * ME + NWARPS * SC * ME + ME;
  if (bwd) f += (size_t)SC * td.wsum + 2 * SC * ME;
  f += (size_t)(bwd ? 2 : 1) * staged_floats(td);
  f += (size_t)(bwd ? FB_N : FF_N) * ept * NTHREADS;   // resident slots (RES instantiations)
  return f * sizeof(float);
}

// Grid and resident-slot plan: the RES instantiation (one CTA per SM, ept = ceil(P / (G * NTHREADS)) <= MAXE elements per
// thread held in shared memory for the whole launch) unless P is too large for that, then the streaming instantiation.
static int grid_for(int P, bool bwd, const TinyDims& td, int& G, int& ept, size_t& smem) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  BSDE_CUDA_OK(cudaGetDevice(&dev));
  int& num_sms = g_num_sms[dev & 63];
  if (num_sms == 0) BSDE_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  static const int res_on = getenv("BSDE_K1_RES") ? atoi(getenv("BSDE_K1_RES")) : 1;
  const int want = (P + NTHREADS - 1) / NTHREADS;
  {
    const int g = std::min(num_sms, want);
    const int e = (int)(((long long)P + (long long)g * NTHREADS - 1) / ((long long)g * NTHREADS));
    const size_t sm = smem_bytes(td, bwd, e);
    if (res_on && e <= MAXE && sm <= 227 * 1024) {
      static PerDeviceOnce once_f, once_b;
      if (bwd ? once_b.first() : once_f.first()) {
        if (bwd) BSDE_CUDA_OK(cudaFuncSetAttribute(wsde_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else BSDE_CUDA_OK(cudaFuncSetAttribute(wsde_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      }
      int occ = 0;
      if (bwd) BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wsde_bwd_kernel<true>, NTHREADS, sm));
      else BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wsde_fwd_kernel<true>, NTHREADS, sm));
      if (occ >= 1) { G = g; ept = e; smem = sm; return 0; }
    }
  }
  ept = 0;
  smem = smem_bytes(td, bwd, 0);
  int occ = 0;
  if (bwd) BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wsde_bwd_kernel<false>, NTHREADS, smem));
  else BSDE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wsde_fwd_kernel<false>, NTHREADS, smem));
  if (occ < 1) { set_error("wsde kernel does not fit on an SM (smem %zu)", smem); return BSDE_ERR_UNSUPPORTED; }
  if (occ > 2) occ = 2;
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
  size_t smem = 0;
  int G = 0, ept = 0;
  if (int e = grid_for(fw->P, false, A.td, G, ept, smem)) return e;
  if (G > MAXG) G = MAXG;
  A.ept = ept;
  A.tp_floats = staged_floats(A.td);
  { static const int st_on = getenv("BSDE_K1_STAMP") ? atoi(getenv("BSDE_K1_STAMP")) : 0; A.stamp = st_on; }
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
    BSDE_CUDA_OK(cudaLaunchCooperativeKernel(ept ? (void*)wsde_fwd_kernel<true> : (void*)wsde_fwd_kernel<false>, dim3(G), dim3(NTHREADS),
                                             args, smem, (cudaStream_t)stream));
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
  size_t smem = 0;
  int G = 0, ept = 0;
  if (int e = grid_for(fw->P, true, A.td, G, ept, smem)) return e;
  if (G > MAXG) G = MAXG;
  A.ept = ept;
  A.tp_floats = staged_floats(A.td);
  { static const int st_on = getenv("BSDE_K1_STAMP") ? atoi(getenv("BSDE_K1_STAMP")) : 0; A.stamp = st_on; }
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
    BSDE_CUDA_OK(cudaLaunchCooperativeKernel(ept ? (void*)wsde_bwd_kernel<true> : (void*)wsde_bwd_kernel<false>, dim3(G), dim3(NTHREADS),
                                             args, smem, (cudaStream_t)stream));
    count_launch();
  }
  return 0;
}

// tools only: copy the phase stamps of the last stamped K1 launches to the host ([2][64][8] uint64, ns)
extern "C" int bsde_debug_k1_stamps(unsigned long long* h_out) {
  BSDE_CUDA_OK(cudaMemcpyFromSymbol(h_out, g_k1_stamps, sizeof(unsigned long long) * 2 * STAMP_IT * STAMP_PH));
  return 0;
}
