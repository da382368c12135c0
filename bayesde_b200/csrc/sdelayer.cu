// Config 1 (JAX toy, row a20): the whole `SDELayer` solve in ONE launch per direction.
//
//   brax/_impl/layers.py:54-118    augmented_post_drift    state [w (P), x (B, D), kl ()]  (ravel order: w first)
//   brax/_impl/layers.py:120-143   augmented_prior_drift
//   brax/_impl/layers.py:145-157   augmented_diffusion     g = [sigma 1_P, 0, 0]
//   brax/_impl/layers.py:164-248   SDELayer.apply_fun      _sdeint(post / prior, ..., linspace(0, 1, 20), rng)
//   brax/utils/sdeint.py:8-60      _sdeint / sdeint        state += h f(t0, state, bm[:P]) + sqrt(h) bm g(t0, state)
//   examples/jax/sdebnn_toy1d.py:32-95  mlp(): driftx = CSL(H) act CSL(H) act CSL(D) with its weights IN THE STATE (w),
//                                   driftw = [Additive(prior,] CSL(H) act CSL(H) act CSL(P) [)], prior = -w ("ou") or 0, g = diff_const
//   CSL = ConcatSquashLineartx (layers.py:19-42): (x W + b) sigmoid(w_t t + w_tb) + b_t t
//
// The reference traces ~40 small XLA ops per step (launch bound).  Here one CTA integrates one MC sample over all
// steps: w, x and the hidden activations of both networks stay in shared memory; the only per-step global traffic is the two big
// driftw matrices W1 [P][H], W3 [H][P] (L2 resident, shared by all samples) and the state row written to the trajectory.
// The reverse pass (hand written) walks the trajectory backwards, recomputing the hidden layers from the stored state:
//   a_k = a_{k+1} + h (df/ds)^T a_{k+1},   dL/dtheta += h (df/dtheta)^T a_{k+1}
// with the sticking-the-landing term sum(u(stop_gradient(theta), w) eps) (layers.py:81-112) feeding w but not theta, so the driftw
// net is back-propagated with two cotangents (one for theta, one for w) when `stop_grad` is set.
// Every reduction has a single owner thread per output and a fixed order: run-to-run deterministic, no float atomics.
#include <algorithm>

#include "common.cuh"

namespace bsde {

constexpr int SL_NT = 1024;

struct SLOff {  // offsets of the driftx leaves inside w (ravel_pytree order)
  int W[3], b[3], wt[3], wtb[3], bt[3], beta[3];
};

struct SLArgs {
  bsde_sdelayer_desc d;
  bsde_fw_params fw, gfw;
  SLOff xo;
  const float* ts; const float* w0; const float* x0; const float* eps;
  float* traj;             // [S][nt][Dtot]
  const float* gtraj;      // [S][nt][Dtot] cotangents of every state, or NULL
  const float* gT;         // [S][Dtot] cotangent of the final state, or NULL
  float* g0;               // [S][Dtot] cotangent of the initial state
  int Bc;                  // rows of x per chunk (shared-memory budget)
};

__device__ __forceinline__ float sl_softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sl_act(int act, float y, float beta) {
  if (act == BSDE_ACT_SWISH) return y * sigmoidf_(y * sl_softplus(beta));
  return act_fwd(act, y);
}
__device__ __forceinline__ float sl_act_grad(int act, float y, float beta) {
  if (act == BSDE_ACT_SWISH) {
    const float g = sl_softplus(beta), sg = sigmoidf_(y * g);
    return sg + y * g * sg * (1.f - sg);
  }
  return act_grad(act, y);
}
__device__ __forceinline__ float sl_act_dbeta(int act, float y, float beta) {
  if (act != BSDE_ACT_SWISH) return 0.f;
  const float g = sl_softplus(beta), sg = sigmoidf_(y * g);
  return y * y * sg * (1.f - sg) * sigmoidf_(beta);
}

// sum_i vec[i] * col[i * stride] with the global loads of 16 elements in flight before the first FMA (vec in shared memory)
__device__ __forceinline__ float sl_dot_col(const float* __restrict__ col, long long stride, const float* vec, int n) {
  float acc = 0.f;
  int i = 0;
  for (; i + 16 <= n; i += 16) {
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __ldg(col + (long long)(i + u) * stride);
#pragma unroll
    for (int u = 0; u < 16; ++u) acc = fmaf(vec[i + u], v[u], acc);
  }
  for (; i < n; ++i) acc = fmaf(vec[i], __ldg(col + (long long)i * stride), acc);
  return acc;
}

__host__ __device__ __forceinline__ int sl_al4(int x) { return (x + 3) & ~3; }

// Small shared-memory product with a 1 x 4 register tile: for r < R, 4-column group j0 < N,
//   acc[0..3] = sum_k A[r * a_rs + k * a_ks] * Bm[k * ldb + j0 .. j0 + 3]          (Bm rows 16-byte aligned, N % 4 == 0)
// one scalar (warp-broadcast) and one 128-bit load per four FMAs instead of two loads per FMA; `store(r, j0, acc)` finishes.
template <class Store>
__device__ __forceinline__ void sl_mm4(const float* A, int a_rs, int a_ks, const float* Bm, int ldb, int R, int K, int N, Store&& store) {
  const int n4 = N >> 2;
  for (int idx = threadIdx.x; idx < R * n4; idx += SL_NT) {
    const int r = idx / n4, j0 = (idx - r * n4) << 2;
    const float* a = A + r * a_rs;
    const float4* b = reinterpret_cast<const float4*>(Bm + j0);
    const int ldb4 = ldb >> 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float av = a[k * a_ks];
      const float4 bv = b[k * ldb4];
      acc.x = fmaf(av, bv.x, acc.x); acc.y = fmaf(av, bv.y, acc.y); acc.z = fmaf(av, bv.z, acc.z); acc.w = fmaf(av, bv.w, acc.w);
    }
    store(r, j0, acc);
  }
}

// fixed-order block sum; `scr` holds SL_NT / 32 floats; every thread gets the result
__device__ __forceinline__ float sl_block_sum(float v, float* scr) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scr[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < SL_NT / 32; ++i) r += scr[i];
  return r;
}

__device__ __forceinline__ float sl_noise(const SLArgs& A, int s, int k, int p) {
  if (A.eps) return A.eps[((long long)s * (A.d.nt - 1) + k) * A.d.P + p];
  float z[4];
  const unsigned sg = (unsigned)s + A.d.sample_offset;
  philox_normal4(A.d.seed, (uint32_t)p, (uint32_t)k, sg >> 2, A.d.stream_id, z);
  return z[sg & 3u];
}

// hidden layers of the driftw net at w (in shared memory): z1 -> a1 -> h1 -> z2 -> a2 -> h2, gates gA / gB.  `red` holds SL_NT floats.
struct SLHid { float *z1, *a1, *h1, *z2, *a2, *h2, *gA, *gB; };
__device__ __forceinline__ void sl_driftw_hidden(const SLArgs& A, const float* sw, float t, const SLHid& hd, float* red) {
  const int H = A.d.H, P = (int)A.d.P, tid = threadIdx.x;
  const int slices = SL_NT / H;
  {
    const int j = tid % H, sl = tid / H;
    float acc = 0.f;
    if (sl < slices) {
      // the loads of a group are issued before the first FMA that needs one (explicit arrays: the compiler otherwise serialises
      // load -> FMA pairs on one register and the loop runs at one L2 latency per element)
      const float* __restrict__ col = A.fw.W1 + j;
      int p = sl;
      for (; p + 15 * slices < P; p += 16 * slices) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldg(col + (long long)(p + u * slices) * H);
#pragma unroll
        for (int u = 0; u < 16; ++u) acc = fmaf(sw[p + u * slices], v[u], acc);
      }
      for (; p < P; p += slices) acc = fmaf(sw[p], __ldg(col + (long long)p * H), acc);
    }
    red[tid] = acc;
  }
  __syncthreads();
  if (tid < H) {
    float z = __ldg(A.fw.b1 + tid);
    for (int sl = 0; sl < slices; ++sl) z += red[sl * H + tid];
    const float g = sigmoidf_(__ldg(A.fw.wt1 + tid) * t + __ldg(A.fw.wtb1 + tid));
    const float a = z * g + __ldg(A.fw.bt1 + tid) * t;
    hd.z1[tid] = z; hd.gA[tid] = g; hd.a1[tid] = a;
    hd.h1[tid] = sl_act(A.d.act, a, A.d.act == BSDE_ACT_SWISH ? __ldg(A.fw.beta1 + tid) : 0.f);
  }
  __syncthreads();
  if (tid < H) {
    float z = __ldg(A.fw.bm[0] + tid);
    for (int i = 0; i < H; ++i) z = fmaf(hd.h1[i], __ldg(A.fw.Wm[0] + i * H + tid), z);
    const float g = sigmoidf_(__ldg(A.fw.wtm[0] + tid) * t + __ldg(A.fw.wtbm[0] + tid));
    const float a = z * g + __ldg(A.fw.btm[0] + tid) * t;
    hd.z2[tid] = z; hd.gB[tid] = g; hd.a2[tid] = a;
    hd.h2[tid] = sl_act(A.d.act, a, A.d.act == BSDE_ACT_SWISH ? __ldg(A.fw.betam[0] + tid) : 0.f);
  }
  __syncthreads();
}

// gates of the three driftx layers at time t: gx[0..H), gx[H..2H), gx[2H..2H+D)
__device__ __forceinline__ void sl_x_gates(const SLArgs& A, const float* sw, float t, float* gx) {
  const int H = A.d.H, D = A.d.D;
  for (int i = threadIdx.x; i < 2 * H + D; i += SL_NT) {
    const int l = i < H ? 0 : (i < 2 * H ? 1 : 2), j = i - l * H;
    gx[i] = sigmoidf_(sw[A.xo.wt[l] + j] * t + sw[A.xo.wtb[l] + j]);
  }
}

// driftx on rows [r0, r0 + nr): Z1 / Z2 pre-gate values (KEEP), H1 / H2 activations; returns nothing, layer 3 is done by the caller
template <bool KEEP>
__device__ __forceinline__ void sl_x_hidden(const SLArgs& A, const float* sw, const float* sx, const float* gx, float t, int r0, int nr,
                                            float* Z1, float* Z2, float* H1, float* H2) {
  const int H = A.d.H, D = A.d.D, act = A.d.act;
  const bool swi = act == BSDE_ACT_SWISH;
  for (int idx = threadIdx.x; idx < nr * H; idx += SL_NT) {
    const int r = idx / H, j = idx - r * H;
    float z = sw[A.xo.b[0] + j];
    for (int dd = 0; dd < D; ++dd) z = fmaf(sx[(r0 + r) * D + dd], sw[A.xo.W[0] + dd * H + j], z);
    if (KEEP) Z1[idx] = z;
    H1[idx] = sl_act(act, z * gx[j] + sw[A.xo.bt[0] + j] * t, swi ? sw[A.xo.beta[0] + j] : 0.f);
  }
  __syncthreads();
  sl_mm4(H1, H, 1, sw + A.xo.W[1], H, nr, H, H, [&](int r, int j0, const float4& acc) {
    const float zz[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      const float z = zz[u] + sw[A.xo.b[1] + j];
      if (KEEP) Z2[r * H + j] = z;
      H2[r * H + j] = sl_act(act, z * gx[H + j] + sw[A.xo.bt[1] + j] * t, swi ? sw[A.xo.beta[1] + j] : 0.f);
    }
  });
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------------------
// forward: one CTA per sample, all steps
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SL_NT) sdelayer_fwd_kernel(SLArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int s = blockIdx.x, tid = threadIdx.x;
  const int P = (int)A.d.P, B = A.d.B, D = A.d.D, H = A.d.H, BD = B * D;
  const long long Dtot = (long long)P + BD + 1;
  float* sw = sm;                      // every segment starts on a 16-byte boundary (P, H are multiples of 4)
  float* sx = sw + P;
  float* gx = sx + sl_al4(BD);
  float* red = gx + sl_al4(2 * H + D);
  float* scr = red + SL_NT;
  SLHid hd;
  hd.z1 = scr + 32; hd.a1 = hd.z1 + H; hd.h1 = hd.a1 + H; hd.z2 = hd.h1 + H; hd.a2 = hd.z2 + H; hd.h2 = hd.a2 + H; hd.gA = hd.h2 + H; hd.gB = hd.gA + H;
  float* H1 = hd.gB + H;
  float* H2 = H1 + A.Bc * H;
  float* tr = A.traj + (long long)s * A.d.nt * Dtot;
  const float* w0 = A.w0 + (A.d.w0_per_sample ? (long long)s * P : 0);
  for (int p = tid; p < P; p += SL_NT) { sw[p] = w0[p]; tr[p] = w0[p]; }
  for (int i = tid; i < BD; i += SL_NT) { sx[i] = A.x0[i]; tr[P + i] = A.x0[i]; }
  if (tid == 0) tr[P + BD] = 0.f;
  float kl = 0.f;
  const bool mlp = A.d.driftw && !A.d.prior_solve;
  const float ou = A.d.prior_ou ? 1.f : 0.f;
  const float sig = A.d.sigma, isig = A.d.diffw ? 1.f / A.d.sigma : 0.f;
  __syncthreads();
  for (int k = 0; k + 1 < A.d.nt; ++k) {
    const float t = A.ts[k], h = A.ts[k + 1] - A.ts[k], sq = sqrtf(h);
    float* nx = tr + (long long)(k + 1) * Dtot;
    if (mlp) sl_driftw_hidden(A, sw, t, hd, red);
    sl_x_gates(A, sw, t, gx);
    __syncthreads();
    // x rows at the OLD w
    for (int r0 = 0; r0 < B; r0 += A.Bc) {
      const int nr = min(A.Bc, B - r0);
      sl_x_hidden<false>(A, sw, sx, gx, t, r0, nr, nullptr, nullptr, H1, H2);
      for (int idx = tid; idx < nr * D; idx += SL_NT) {
        const int r = idx / D, dd = idx - r * D;
        float z = sw[A.xo.b[2] + dd];
        for (int i = 0; i < H; ++i) z = fmaf(H2[r * H + i], sw[A.xo.W[2] + i * D + dd], z);
        const float dx = z * gx[2 * H + dd] + sw[A.xo.bt[2] + dd] * t;
        const float xn = sx[(r0 + r) * D + dd] + h * dx;   // rows are independent; layer 1 of this chunk is complete
        sx[(r0 + r) * D + dd] = xn;
        nx[P + (r0 + r) * D + dd] = xn;
      }
      __syncthreads();
    }
    // w rows and the kl integrand
    float klacc = 0.f;
    for (int p = tid; p < P; p += SL_NT) {
      const float w = sw[p];
      float m = 0.f;
      if (mlp) {
        float z = sl_dot_col(A.fw.W4 + p, P, hd.h2, H) + __ldg(A.fw.b4 + p);
        m = z * sigmoidf_(__ldg(A.fw.wt4 + p) * t + __ldg(A.fw.wtb4 + p)) + __ldg(A.fw.bt4 + p) * t;
      }
      const float prior = -ou * w;
      float dw, u;
      if (A.d.prior_solve) { dw = prior; u = dw * isig; }
      else { dw = A.d.driftw ? (A.d.diff_drift ? m + prior : m) : 0.f; u = (dw - prior) * isig; }
      const float e = sl_noise(A, s, k, p);
      klacc += 0.5f * u * u + u * e;
      const float wn = w + h * dw + sq * e * (A.d.diffw ? sig : 0.f);
      sw[p] = wn;
      nx[p] = wn;
    }
    kl += h * sl_block_sum(klacc, scr);
    if (tid == 0) nx[P + BD] = kl;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// reverse pass
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SL_NT) sdelayer_bwd_kernel(SLArgs A) {
  extern __shared__ __align__(16) float sm[];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = (int)A.d.P, B = A.d.B, D = A.d.D, H = A.d.H, BD = B * D, act = A.d.act;
  const bool swi = act == BSDE_ACT_SWISH;
  const long long Dtot = (long long)P + BD + 1;
  float* sw = sm;
  float* aw = sw + P;
  float* daw = aw + P;
  float* gzt = daw + P;
  float* gzw = A.d.stop_grad ? gzt + P : gzt;
  float* sx = gzt + (A.d.stop_grad ? 2 : 1) * P;
  float* ax = sx + sl_al4(BD);
  float* dax = ax + sl_al4(BD);
  float* gx = dax + sl_al4(BD);
  float* red = gx + sl_al4(2 * H + D);
  float* scr = red + SL_NT;
  SLHid hd;
  hd.z1 = scr + 32; hd.a1 = hd.z1 + H; hd.h1 = hd.a1 + H; hd.z2 = hd.h1 + H; hd.a2 = hd.z2 + H; hd.h2 = hd.a2 + H; hd.gA = hd.h2 + H; hd.gB = hd.gA + H;
  float* ght = hd.gB + H;     // cotangent chains of the driftw hidden layers: theta and w versions
  float* ghw = ght + H;
  float* gz2t = ghw + H;
  float* gz2w = gz2t + H;
  float* g3 = gz2w + H;       // per (row, d) scratch of driftx layer 3: GZ3, Q3, GDX  [Bc * D each]
  float* Z1 = g3 + sl_al4(3 * A.Bc * D);
  float* Z2 = Z1 + A.Bc * H;
  float* H1 = Z2 + A.Bc * H;
  float* H2 = H1 + A.Bc * H;
  float* G = H2 + A.Bc * H;
  float* WT = G + A.Bc * H;   // transposed copy of driftx's middle matrix at the current state: WT[j][i] = W[i][j]
  const float* tr = A.traj + (long long)s * A.d.nt * Dtot;
  const bool mlp = A.d.driftw && !A.d.prior_solve;
  const float ou = A.d.prior_ou ? 1.f : 0.f;
  const float isig = A.d.diffw ? 1.f / A.d.sigma : 0.f;
  // per-sample gradient accumulators of the driftw parameters
  const long long PH = (long long)P * H;
  float* gW1 = mlp ? A.gfw.W1 + s * PH : nullptr;
  float* gW4 = mlp ? A.gfw.W4 + s * PH : nullptr;
  float* gWm = mlp ? A.gfw.Wm[0] + (long long)s * H * H : nullptr;
  float akl = 0.f;
  for (int p = tid; p < P; p += SL_NT) aw[p] = 0.f;
  for (int i = tid; i < BD; i += SL_NT) ax[i] = 0.f;
  __syncthreads();

  for (int k = A.d.nt - 2; k >= 0; --k) {
    const float t = A.ts[k], h = A.ts[k + 1] - A.ts[k];
    // cotangent injected at state k+1
    {
      const float* gin = A.gtraj ? A.gtraj + ((long long)s * A.d.nt + k + 1) * Dtot : (k == A.d.nt - 2 ? A.gT + (long long)s * Dtot : nullptr);
      if (gin) {
        for (int p = tid; p < P; p += SL_NT) aw[p] += gin[p];
        for (int i = tid; i < BD; i += SL_NT) ax[i] += gin[P + i];
        akl += gin[P + BD];
      }
    }
    const float* st = tr + (long long)k * Dtot;
    for (int p = tid; p < P; p += SL_NT) { sw[p] = st[p]; daw[p] = 0.f; }
    for (int i = tid; i < BD; i += SL_NT) { sx[i] = st[P + i]; dax[i] = 0.f; }
    __syncthreads();
    if (mlp) sl_driftw_hidden(A, sw, t, hd, red);
    sl_x_gates(A, sw, t, gx);
    for (int e2 = tid; e2 < H * H; e2 += SL_NT) WT[(e2 % H) * H + e2 / H] = sw[A.xo.W[1] + e2];
    __syncthreads();

    // ---- w rows: phase A, one thread per p -----------------------------------------------------------------------------
    for (int p = tid; p < P; p += SL_NT) {
      const float w = sw[p];
      float z3 = 0.f, gt3 = 0.f, m = 0.f;
      if (mlp) {
        z3 = sl_dot_col(A.fw.W4 + p, P, hd.h2, H) + __ldg(A.fw.b4 + p);
        gt3 = sigmoidf_(__ldg(A.fw.wt4 + p) * t + __ldg(A.fw.wtb4 + p));
        m = z3 * gt3 + __ldg(A.fw.bt4 + p) * t;
      }
      const float prior = -ou * w;
      float u, ddw_dm, du_dm, ddw_dw, du_dw;
      if (A.d.prior_solve) {
        u = prior * isig; ddw_dm = 0.f; du_dm = 0.f; ddw_dw = -ou; du_dw = -ou * isig;
      } else {
        const float dw = A.d.driftw ? (A.d.diff_drift ? m + prior : m) : 0.f;
        u = (dw - prior) * isig;
        ddw_dm = A.d.driftw ? 1.f : 0.f;
        du_dm = ddw_dm * isig;
        ddw_dw = (A.d.driftw && A.d.diff_drift) ? -ou : 0.f;
        du_dw = (ddw_dw + ou) * isig;
      }
      const float e = sl_noise(A, s, k, p);
      const float abar = aw[p];
      const float ct = h * (abar * ddw_dm + akl * (u + (A.d.stop_grad ? 0.f : e)) * du_dm);
      const float cw = ct + h * akl * (A.d.stop_grad ? e : 0.f) * du_dm;
      daw[p] += h * (abar * ddw_dw + akl * (u + e) * du_dw);
      if (mlp) {
        const float gzT = ct * gt3;
        gzt[p] = gzT;
        if (A.d.stop_grad) gzw[p] = cw * gt3;
        const float q = ct * z3 * gt3 * (1.f - gt3);
        A.gfw.bt4[(long long)s * P + p] += ct * t;
        A.gfw.wt4[(long long)s * P + p] += q * t;
        A.gfw.wtb4[(long long)s * P + p] += q;
        A.gfw.b4[(long long)s * P + p] += gzT;
        {   // read-modify-write of this sample's accumulator column: loads of a group first, then the stores (no aliasing between rows)
          float* __restrict__ col = gW4 + p;
          int i = 0;
          for (; i + 8 <= H; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = col[(long long)(i + u) * P];
#pragma unroll
            for (int u = 0; u < 8; ++u) col[(long long)(i + u) * P] = fmaf(hd.h2[i + u], gzT, v[u]);
          }
          for (; i < H; ++i) col[(long long)i * P] += hd.h2[i] * gzT;
        }
      }
    }
    __syncthreads();
    if (mlp) {
      // ---- phase B: gh2[i] = sum_p W4[i][p] gz[p], one warp per hidden unit, lanes stride p (coalesced rows) ----------------
      for (int i = warp; i < H; i += SL_NT / 32) {
        float at = 0.f, awv = 0.f;
        const float* row = A.fw.W4 + (long long)i * P;
        int p = lane;
        for (; p + 15 * 32 < P; p += 16 * 32) {
          float v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = __ldg(row + p + u * 32);
#pragma unroll
          for (int u = 0; u < 16; ++u) { at = fmaf(v[u], gzt[p + u * 32], at); awv = fmaf(v[u], gzw[p + u * 32], awv); }
        }
        for (; p < P; p += 32) {
          const float wv = __ldg(row + p);
          at = fmaf(wv, gzt[p], at);
          awv = fmaf(wv, gzw[p], awv);
        }
        at = warp_sum(at); awv = warp_sum(awv);
        if (lane == 0) { ght[i] = at; ghw[i] = awv; }
      }
      __syncthreads();
      // ---- phase C: hidden layer 2 -----------------------------------------------------------------------------------------
      if (tid < H) {
        const float b2v = swi ? __ldg(A.fw.betam[0] + tid) : 0.f;
        const float dact = sl_act_grad(act, hd.a2[tid], b2v);
        const float gat = ght[tid] * dact, gaw = ghw[tid] * dact;
        const float gB = hd.gB[tid];
        if (swi) A.gfw.betam[0][(long long)s * H + tid] += ght[tid] * sl_act_dbeta(act, hd.a2[tid], b2v);
        const float q = gat * hd.z2[tid] * gB * (1.f - gB);
        A.gfw.btm[0][(long long)s * H + tid] += gat * t;
        A.gfw.wtm[0][(long long)s * H + tid] += q * t;
        A.gfw.wtbm[0][(long long)s * H + tid] += q;
        A.gfw.bm[0][(long long)s * H + tid] += gat * gB;
        gz2t[tid] = gat * gB; gz2w[tid] = gaw * gB;
      }
      __syncthreads();
      for (int e2 = tid; e2 < H * H; e2 += SL_NT) gWm[e2] += hd.h1[e2 / H] * gz2t[e2 % H];
      if (tid < H) {   // gh1[i] = sum_j Wm[i][j] gz2[j]
        float at = 0.f, awv = 0.f;
        for (int j = 0; j < H; ++j) {
          const float wv = __ldg(A.fw.Wm[0] + tid * H + j);
          at = fmaf(wv, gz2t[j], at);
          awv = fmaf(wv, gz2w[j], awv);
        }
        ght[tid] = at; ghw[tid] = awv;   // ght / ghw of layer 2 were consumed above (same thread)
      }
      __syncthreads();
      if (tid < H) {   // hidden layer 1
        const float b1v = swi ? __ldg(A.fw.beta1 + tid) : 0.f;
        const float dact = sl_act_grad(act, hd.a1[tid], b1v);
        const float gat = ght[tid] * dact, gaw = ghw[tid] * dact;
        const float gA = hd.gA[tid];
        if (swi) A.gfw.beta1[(long long)s * H + tid] += ght[tid] * sl_act_dbeta(act, hd.a1[tid], b1v);
        const float q = gat * hd.z1[tid] * gA * (1.f - gA);
        A.gfw.bt1[(long long)s * H + tid] += gat * t;
        A.gfw.wt1[(long long)s * H + tid] += q * t;
        A.gfw.wtb1[(long long)s * H + tid] += q;
        A.gfw.b1[(long long)s * H + tid] += gat * gA;
        gz2t[tid] = gat * gA; gz2w[tid] = gaw * gA;   // now the layer-1 pre-gate cotangents gz1
      }
      __syncthreads();
      // ---- phase D: first layer ----------------------------------------------------------------------------------------------
      {
        int e1 = tid;
        for (; e1 + 7 * SL_NT < (int)PH; e1 += 8 * SL_NT) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = gW1[e1 + u * SL_NT];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const int e = e1 + u * SL_NT; gW1[e] = fmaf(sw[e / H], gz2t[e % H], v[u]); }
        }
        for (; e1 < (int)PH; e1 += SL_NT) gW1[e1] += sw[e1 / H] * gz2t[e1 % H];
      }
      for (int p = tid; p < P; p += SL_NT) {
        float acc = 0.f;
        acc = sl_dot_col(A.fw.W1 + (long long)p * H, 1, gz2w, H);
        daw[p] += acc;
      }
      __syncthreads();
    }

    // ---- x rows: driftx(w, x) back-propagated to x (dax) and to its weights, which are state entries (daw) ----------------------
    for (int r0 = 0; r0 < B; r0 += A.Bc) {
      const int nr = min(A.Bc, B - r0);
      sl_x_hidden<true>(A, sw, sx, gx, t, r0, nr, Z1, Z2, H1, H2);
      float* GZ3 = g3; float* Q3 = g3 + A.Bc * D; float* GDX = Q3 + A.Bc * D;
      for (int idx = tid; idx < nr * D; idx += SL_NT) {
        const int r = idx / D, dd = idx - r * D;
        float z = sw[A.xo.b[2] + dd];
        for (int i = 0; i < H; ++i) z = fmaf(H2[r * H + i], sw[A.xo.W[2] + i * D + dd], z);
        const float gdx = h * ax[(r0 + r) * D + dd], g = gx[2 * H + dd];
        GDX[idx] = gdx; GZ3[idx] = gdx * g; Q3[idx] = gdx * z * g * (1.f - g);
      }
      __syncthreads();
      // layer-3 leaves: W [H][D], b, w_t, w_tb, b_t [D]
      for (int o = tid; o < H * D + 4 * D; o += SL_NT) {
        float acc = 0.f;
        if (o < H * D) {
          const int i = o / D, dd = o - i * D;
          for (int r = 0; r < nr; ++r) acc = fmaf(H2[r * H + i], GZ3[r * D + dd], acc);
          daw[A.xo.W[2] + o] += acc;
        } else {
          const int v = (o - H * D) / D, dd = (o - H * D) % D;
          const float* src = v == 0 ? GZ3 : (v == 3 ? GDX : Q3);
          for (int r = 0; r < nr; ++r) acc += src[r * D + dd];
          if (v == 0) daw[A.xo.b[2] + dd] += acc;
          else if (v == 1) daw[A.xo.wt[2] + dd] += acc * t;
          else if (v == 2) daw[A.xo.wtb[2] + dd] += acc;
          else daw[A.xo.bt[2] + dd] += acc * t;
        }
      }
      // GH2[r][i] = sum_d GZ3[r][d] W3x[i][d]
      for (int idx = tid; idx < nr * H; idx += SL_NT) {
        const int r = idx / H, i = idx - r * H;
        float acc = 0.f;
        for (int dd = 0; dd < D; ++dd) acc = fmaf(GZ3[r * D + dd], sw[A.xo.W[2] + i * D + dd], acc);
        G[idx] = acc;
      }
      __syncthreads();
      // two hidden layers, l = 1 then 0: G holds dL/d(activation output) on entry
      for (int l = 1; l >= 0; --l) {
        float* Z = l == 1 ? Z2 : Z1;
        const float* Hin = l == 1 ? H1 : nullptr;   // layer input (l == 0: rows of x)
        const int din = l == 1 ? H : D;
        // pass A: one thread per unit, rows in order: vector leaves + G <- pre-gate cotangent
        if (tid < H) {
          const int j = tid;
          const float g = gx[l * H + j], btv = sw[A.xo.bt[l] + j], bv = swi ? sw[A.xo.beta[l] + j] : 0.f;
          float sb = 0.f, sq_ = 0.f, sa = 0.f, sbeta = 0.f;
          for (int r = 0; r < nr; ++r) {
            const float z = Z[r * H + j], a = z * g + btv * t, gh = G[r * H + j];
            const float ga = gh * sl_act_grad(act, a, bv);
            sbeta += gh * sl_act_dbeta(act, a, bv);
            sa += ga; sq_ += ga * z * g * (1.f - g);
            const float gz = ga * g;
            sb += gz;
            G[r * H + j] = gz;
          }
          daw[A.xo.b[l] + j] += sb;
          daw[A.xo.wt[l] + j] += sq_ * t;
          daw[A.xo.wtb[l] + j] += sq_;
          daw[A.xo.bt[l] + j] += sa * t;
          if (swi) daw[A.xo.beta[l] + j] += sbeta;
        }
        __syncthreads();
        // pass B: W leaf [din][H]
        {   // dW[i][j] = sum_r in[r][i] G[r][j]: the layer input read transposed (row stride 1, reduction stride = its pitch)
          float* dW = daw + A.xo.W[l];
          sl_mm4(l == 1 ? Hin : sx + r0 * D, 1, l == 1 ? H : D, G, H, din, nr, H, [&](int i, int j0, const float4& acc) {
            float* q = dW + i * H + j0;
            q[0] += acc.x; q[1] += acc.y; q[2] += acc.z; q[3] += acc.w;
          });
        }
        // pass C: cotangent of the layer input
        if (l == 1) {
          // GH1[r][i] = sum_j G[r][j] W[i][j] = sum_j G[r][j] WT[j][i] (WT: this step's transposed copy), into Z2 (consumed by pass A)
          sl_mm4(G, H, 1, WT, H, nr, H, H, [&](int r, int i0, const float4& acc) {
            *reinterpret_cast<float4*>(Z2 + r * H + i0) = acc;
          });
          __syncthreads();
          for (int idx = tid; idx < nr * H; idx += SL_NT) G[idx] = Z2[idx];
          __syncthreads();
        } else {
          for (int idx = tid; idx < nr * D; idx += SL_NT) {
            const int r = idx / D, dd = idx - r * D;
            float acc = 0.f;
            for (int j = 0; j < H; ++j) acc = fmaf(G[r * H + j], sw[A.xo.W[0] + dd * H + j], acc);
            dax[(r0 + r) * D + dd] += acc;
          }
          __syncthreads();
        }
      }
    }
    // a_k = a_{k+1} + contributions
    for (int p = tid; p < P; p += SL_NT) aw[p] += daw[p];
    for (int i = tid; i < BD; i += SL_NT) ax[i] += dax[i];
    __syncthreads();
  }
  float* g0 = A.g0 + (long long)s * Dtot;
  // cotangent injected at the initial state itself (row 0 of gtraj)
  const float* gin0 = A.gtraj ? A.gtraj + (long long)s * A.d.nt * Dtot : nullptr;
  for (int p = tid; p < P; p += SL_NT) g0[p] = aw[p] + (gin0 ? gin0[p] : 0.f);
  for (int i = tid; i < BD; i += SL_NT) g0[P + i] = ax[i] + (gin0 ? gin0[P + i] : 0.f);
  if (tid == 0) g0[P + BD] = akl + (gin0 ? gin0[P + BD] : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------------
static int sl_check(const bsde_sdelayer_desc* d, const bsde_fw_params* fw, SLOff& xo) {
  BSDE_CHECK_ARG(d != nullptr, "NULL descriptor");
  BSDE_CHECK_ARG(d->S >= 1 && d->B >= 1 && d->D >= 1 && d->nt >= 2, "S, B, D must be positive and nt >= 2");
  BSDE_CHECK_ARG(d->H >= 4 && d->H <= 128 && d->H % 4 == 0, "hidden width %d unsupported (a multiple of 4 in 4..128)", d->H);
  BSDE_CHECK_ARG(d->D <= 16, "x width %d unsupported (<= 16)", d->D);
  BSDE_CHECK_ARG((d->act >= BSDE_ACT_SOFTPLUS && d->act <= BSDE_ACT_ELU) || d->act == BSDE_ACT_RBF || d->act == BSDE_ACT_SWISH,
                 "activation %d unsupported", d->act);
  BSDE_CHECK_ARG(!d->diffw || d->sigma > 0.f, "diffw needs sigma > 0");
  const int H = d->H, D = d->D;
  const int dims[4] = {D, H, H, D};
  int off = 0;
  for (int l = 0; l < 3; ++l) {
    const int din = dims[l], dout = dims[l + 1];
    xo.W[l] = off; xo.b[l] = off + din * dout; xo.wt[l] = xo.b[l] + dout; xo.wtb[l] = xo.wt[l] + dout; xo.bt[l] = xo.wtb[l] + dout;
    off = xo.bt[l] + dout;
    xo.beta[l] = -1;
    if (l < 2 && d->act == BSDE_ACT_SWISH) { xo.beta[l] = off; off += dout; }
  }
  BSDE_CHECK_ARG((long long)off * H < (1LL << 30), "P * H too large");
  BSDE_CHECK_ARG(d->P == off, "P = %lld does not match the driftx parameter count %d for D = %d, H = %d", (long long)d->P, off, D, H);
  if (d->driftw && !d->prior_solve) {
    BSDE_CHECK_ARG(fw != nullptr && fw->nhid == 2 && fw->dims[0] == H && fw->dims[1] == H && fw->P == d->P,
                   "driftw net must be P -> H -> H -> P");
    BSDE_CHECK_ARG(fw->W1 && fw->b1 && fw->wt1 && fw->wtb1 && fw->bt1 && fw->Wm[0] && fw->bm[0] && fw->wtm[0] && fw->wtbm[0] && fw->btm[0] &&
                       fw->W4 && fw->b4 && fw->wt4 && fw->wtb4 && fw->bt4, "NULL driftw parameter");
    BSDE_CHECK_ARG(d->act != BSDE_ACT_SWISH || (fw->beta1 && fw->betam[0]), "swish driftw needs beta1 / betam[0]");
  }
  return 0;
}

static size_t sl_fixed_floats(const bsde_sdelayer_desc* d, int backward) {
  const size_t P = (size_t)d->P, BD = (size_t)sl_al4(d->B * d->D), H = d->H;
  const size_t common = sl_al4(2 * d->H + d->D) + SL_NT + 32 + 8 * H;
  if (!backward) return P + BD + common;
  return (d->stop_grad ? 5 : 4) * P + 3 * BD + common + 4 * H + H * H + 4;
}
static size_t sl_row_floats(const bsde_sdelayer_desc* d, int backward) { return backward ? 5 * (size_t)d->H + 3 * d->D : 2 * (size_t)d->H; }

static int sl_chunk(const bsde_sdelayer_desc* d, int backward, size_t& smem) {
  const size_t budget = 200 * 1024 / sizeof(float);
  const size_t fixed = sl_fixed_floats(d, backward), row = sl_row_floats(d, backward);
  if (fixed + row > budget) return 0;
  const int Bc = (int)std::min<size_t>((budget - fixed) / row, (size_t)d->B);
  smem = (fixed + (size_t)Bc * row) * sizeof(float);
  return Bc;
}

}  // namespace bsde

using namespace bsde;

extern "C" int bsde_sdelayer_fwd(const bsde_sdelayer_desc* d, const bsde_fw_params* fw, const float* d_ts, const float* d_w0,
                                 const float* d_x0, const float* d_eps, float* d_traj, void* stream) {
  SLArgs A;
  memset(&A, 0, sizeof(A));
  if (int e = sl_check(d, fw, A.xo)) return e;
  BSDE_CHECK_ARG(d_ts && d_w0 && d_x0 && d_traj, "NULL pointer");
  size_t smem = 0;
  A.Bc = sl_chunk(d, 0, smem);
  BSDE_CHECK_ARG(A.Bc >= 1, "state does not fit in shared memory (P = %lld, B*D = %d)", (long long)d->P, d->B * d->D);
  A.d = *d;
  if (fw) A.fw = *fw;
  A.ts = d_ts; A.w0 = d_w0; A.x0 = d_x0; A.eps = d_eps; A.traj = d_traj;
  BSDE_CUDA_OK(cudaFuncSetAttribute(sdelayer_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sdelayer_fwd_kernel<<<d->S, SL_NT, smem, (cudaStream_t)stream>>>(A);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int bsde_sdelayer_bwd(const bsde_sdelayer_desc* d, const bsde_fw_params* fw, const float* d_ts, const float* d_eps,
                                 const float* d_traj, const float* d_gtraj, const float* d_gT, float* d_g0, const bsde_fw_params* gfw,
                                 void* stream) {
  SLArgs A;
  memset(&A, 0, sizeof(A));
  if (int e = sl_check(d, fw, A.xo)) return e;
  BSDE_CHECK_ARG(d_ts && d_traj && d_g0 && (d_gtraj || d_gT), "NULL pointer");
  if (d->driftw && !d->prior_solve) {
    BSDE_CHECK_ARG(gfw && gfw->W1 && gfw->b1 && gfw->wt1 && gfw->wtb1 && gfw->bt1 && gfw->Wm[0] && gfw->bm[0] && gfw->wtm[0] && gfw->wtbm[0] &&
                       gfw->btm[0] && gfw->W4 && gfw->b4 && gfw->wt4 && gfw->wtb4 && gfw->bt4, "NULL driftw gradient accumulator");
    BSDE_CHECK_ARG(d->act != BSDE_ACT_SWISH || (gfw->beta1 && gfw->betam[0]), "swish driftw needs beta gradient accumulators");
    A.gfw = *gfw;
  }
  size_t smem = 0;
  A.Bc = sl_chunk(d, 1, smem);
  BSDE_CHECK_ARG(A.Bc >= 1, "state does not fit in shared memory (P = %lld, B*D = %d)", (long long)d->P, d->B * d->D);
  A.d = *d;
  if (fw) A.fw = *fw;
  A.ts = d_ts; A.eps = d_eps; A.traj = const_cast<float*>(d_traj); A.gtraj = d_gtraj; A.gT = d_gT; A.g0 = d_g0;
  BSDE_CUDA_OK(cudaFuncSetAttribute(sdelayer_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sdelayer_bwd_kernel<<<d->S, SL_NT, smem, (cudaStream_t)stream>>>(A);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
