// K2 (fp32 mode): time-dependent convolution as an implicit GEMM on CUDA cores.
//
// Replaces the reference's  concat(x, t) -> lax.conv_general_dilated / F.conv2d  call
// (brax/_impl/diffeq_layers.py:143-148, torchbnn/_impl/diffeq_layers.py:100-109,148-156) with the
// kernel taken from the flat SDE state w(t) of the image's Monte-Carlo sample.  The time channel is
// never materialised: the gather substitutes `t` for the extra reduction index wherever the tap is
// inside the image (zero padding applies to it too, as in the reference).
//
// The same gather kernel evaluates forward convs (stride 1/2, SAME, lhs-dilated "transposed") and
// every data gradient, by the per-axis map   num = o*sn + k*kn + off,  i = num/sd  (see bsde.h).
// fp32 FMA accumulation: this is the bit-faithful path; BSDE_MATH_TC is the tensor-core path.
#include "common.cuh"

namespace bsde {

struct ConvArgs {
  const float* in;
  const float* w;
  const float* aux;
  float* out;
  int S, B;
  long long w_sample_stride;
  int cin, cout, ksize, hin, win, hout, wout, sn, kn, off, sd;
  int has_time, time_row, real_base;
  long long w_off, b_off, ts, ks, ns;
  float t, scale;
  int act, epi;
  int Kc, Ktot, Mper;
};

__device__ __forceinline__ bool map_axis(int o, int k, int sn, int kn, int off, int sd, int size, int& i) {
  int num = o * sn + k * kn + off;
  if (num < 0) return false;
  if (sd == 2) {
    if (num & 1) return false;
    num >>= 1;
  }
  i = num;
  return num < size;
}

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) conv_gather_kernel(ConvArgs a) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  constexpr int PPT = BM / 16;  // pixels gathered per thread per k-tile
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int s = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const float* __restrict__ in = a.in + (long long)s * a.B * a.hin * a.win * a.cin;
  const float* __restrict__ w = a.w + (long long)s * a.w_sample_stride;

  // pixel decode for the gather role
  const int kl = tid % BK;
  const int ml0 = tid / BK;
  int p_oy[PPT], p_ox[PPT], p_img[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    int m = m0 + ml0 + i * 16;
    if (m < a.Mper) {
      int hw = a.hout * a.wout;
      int img = m / hw;
      int r = m - img * hw;
      p_img[i] = img;
      p_oy[i] = r / a.wout;
      p_ox[i] = r - p_oy[i] * a.wout;
    } else {
      p_img[i] = -1; p_oy[i] = 0; p_ox[i] = 0;
    }
  }

  const int ty = tid / (BN / TN);
  const int tx = tid % (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.Ktot; k0 += BK) {
    // ---- gather A tile
    {
      const int k = k0 + kl;
      const bool kvalid = k < a.Ktot;
      int tap = 0, c = 0;
      if (kvalid) { tap = k / a.Kc; c = k - tap * a.Kc; }
      const int kh = tap / a.ksize, kw = tap - kh * a.ksize;
      const bool is_time = a.has_time && (c == a.Kc - 1);
#pragma unroll
      for (int i = 0; i < PPT; ++i) {
        float v = 0.f;
        if (kvalid && p_img[i] >= 0) {
          int iy, ix;
          if (map_axis(p_oy[i], kh, a.sn, a.kn, a.off, a.sd, a.hin, iy) &&
              map_axis(p_ox[i], kw, a.sn, a.kn, a.off, a.sd, a.win, ix)) {
            v = is_time ? a.t
                        : __ldg(in + (((long long)p_img[i] * a.hin + iy) * a.win + ix) * a.cin + c);
          }
        }
        As[kl][ml0 + i * 16] = v;
      }
    }
    // ---- weight tile
    for (int idx = tid; idx < BK * BN; idx += 256) {
      const int kk = idx / BN, nn = idx - kk * BN;
      const int k = k0 + kk, n = n0 + nn;
      float v = 0.f;
      if (k < a.Ktot && n < a.cout) {
        const int tap = k / a.Kc, c = k - tap * a.Kc;
        const int row = (a.has_time && c == a.Kc - 1) ? a.time_row : c + a.real_base;
        v = __ldg(w + a.w_off + tap * a.ts + row * a.ks + n * a.ns);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 q = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        av[i] = q.x; av[i + 1] = q.y; av[i + 2] = q.z; av[i + 3] = q.w;
      }
      if (TN >= 4) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          float4 q = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
          bv[j] = q.x; bv[j + 1] = q.y; bv[j + 2] = q.z; bv[j + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  float* __restrict__ out = a.out + (long long)s * a.Mper * a.cout;
  const float* __restrict__ aux = a.aux ? a.aux + (long long)s * a.Mper * a.cout : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= a.Mper) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= a.cout) continue;
      const long long o = (long long)m * a.cout + n;
      float v = acc[i][j];
      const float b = (a.b_off >= 0 && a.epi <= BSDE_EPI_EULER) ? __ldg(w + a.b_off + n) : 0.f;
      switch (a.epi) {
        case BSDE_EPI_BIAS: v = v + b; break;
        case BSDE_EPI_BIAS_ACT: v = act_fwd(a.act, v + b); break;
        case BSDE_EPI_EULER: v = aux[o] + a.scale * (v + b); break;
        case BSDE_EPI_DACT: v = a.scale * v * act_grad_from_out(a.act, aux[o]); break;
        case BSDE_EPI_ACCUM: v = aux[o] + a.scale * v; break;
      }
      out[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient:  G[i][j] = sum_m A(m, i) * dy[m][j],  i over (tap, channel) rows plus one
// trailing "ones" row that yields the bias gradient.  Pixels are split across CTAs; partial
// tiles go to the workspace and a second kernel reduces them in a fixed order (deterministic).
// ---------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* in;
  const float* dy;
  float* partial;  // [S][splits][rows][cout]
  int S, B, splits, chunk;
  int cin, cout, ksize, hin, win, hout, wout, sn, kn, off, sd;
  int has_time;
  float t;
  int Kc, Ktot, rows, Mper;
};

template <int BI, int BJ, int TM, int TN>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(WgradArgs a) {
  static_assert((BI / TM) * (BJ / TN) == 256, "256 threads");
  __shared__ __align__(16) float As[BK][BI + 4];
  __shared__ __align__(16) float Bs[BK][BJ + 4];
  const int tid = threadIdx.x;
  const int s = blockIdx.z / a.splits;
  const int split = blockIdx.z - s * a.splits;
  const int i0 = blockIdx.x * BI;
  const int j0 = blockIdx.y * BJ;
  const float* __restrict__ in = a.in + (long long)s * a.B * a.hin * a.win * a.cin;
  const float* __restrict__ dy = a.dy + (long long)s * a.Mper * a.cout;
  const int mbeg = split * a.chunk;
  const int mend = min(a.Mper, mbeg + a.chunk);
  const int hw = a.hout * a.wout;

  // gather role: a fixed row i per (tid % BI), BK*BI/256 pixels per k-tile
  constexpr int RPT = BK * BI / 256;  // loads per thread
  constexpr int MSTEP = 256 / BI;     // pixel stride between a thread's loads
  const int il = tid % BI;
  const int i_row = i0 + il;
  int r_kind = 0;  // 0 invalid, 1 data, 2 time, 3 ones (bias)
  int r_kh = 0, r_kw = 0, r_c = 0;
  if (i_row < a.Ktot) {
    const int tap = i_row / a.Kc;
    r_c = i_row - tap * a.Kc;
    r_kh = tap / a.ksize;
    r_kw = tap - r_kh * a.ksize;
    r_kind = (a.has_time && r_c == a.Kc - 1) ? 2 : 1;
  } else if (i_row == a.Ktot) {
    r_kind = 3;
  }

  const int ty = tid / (BJ / TN);
  const int tx = tid % (BJ / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int mb = mbeg; mb < mend; mb += BK) {
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int mm = tid / BI + r * MSTEP;
      const int m = mb + mm;
      float v = 0.f;
      if (m < mend && r_kind) {
        if (r_kind == 3) {
          v = 1.f;
        } else {
          const int img = m / hw;
          const int rem = m - img * hw;
          const int oy = rem / a.wout, ox = rem - oy * a.wout;
          int iy, ix;
          if (map_axis(oy, r_kh, a.sn, a.kn, a.off, a.sd, a.hin, iy) &&
              map_axis(ox, r_kw, a.sn, a.kn, a.off, a.sd, a.win, ix)) {
            v = (r_kind == 2) ? a.t
                              : __ldg(in + (((long long)img * a.hin + iy) * a.win + ix) * a.cin + r_c);
          }
        }
      }
      As[mm][il] = v;
    }
    for (int idx = tid; idx < BK * BJ; idx += 256) {
      const int mm = idx / BJ, jj = idx - mm * BJ;
      const int m = mb + mm, j = j0 + jj;
      Bs[mm][jj] = (m < mend && j < a.cout) ? __ldg(dy + (long long)m * a.cout + j) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 q = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        av[i] = q.x; av[i + 1] = q.y; av[i + 2] = q.z; av[i + 3] = q.w;
      }
      if (TN >= 4) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          float4 q = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
          bv[j] = q.x; bv[j + 1] = q.y; bv[j + 2] = q.z; bv[j + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* __restrict__ part = a.partial + ((long long)(s * a.splits + split) * a.rows) * a.cout;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = i0 + ty * TM + i;
    if (row >= a.rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = j0 + tx * TN + j;
      if (col < a.cout) part[(long long)row * a.cout + col] = acc[i][j];
    }
  }
}

struct WreduceArgs {
  const float* partial;
  float* gw;
  long long gw_sample_stride;
  int S, splits, rows, cout, Kc, Ktot, has_time, time_row, real_base;
  long long w_off, b_off, ts, ks, ns;
  float scale;
};

__global__ void wgrad_reduce_kernel(WreduceArgs a) {
  const long long per = (long long)a.rows * a.cout;
  const long long total = per * a.S;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(idx / per);
    const long long e = idx - s * per;
    const int row = (int)(e / a.cout), col = (int)(e - (long long)row * a.cout);
    const float* p = a.partial + (long long)s * a.splits * per + e;
    float v = 0.f;
    for (int q = 0; q < a.splits; ++q) v += p[q * per];
    v *= a.scale;
    float* g = a.gw + s * a.gw_sample_stride;
    if (row == a.Ktot) {
      if (a.b_off >= 0) g[a.b_off + col] = v;
    } else {
      const int tap = row / a.Kc, c = row - tap * a.Kc;
      const int wr = (a.has_time && c == a.Kc - 1) ? a.time_row : c + a.real_base;
      g[a.w_off + tap * a.ts + wr * a.ks + col * a.ns] = v;
    }
  }
}

static int check_layer(const bsde_conv_layer* L) {
  BSDE_CHECK_ARG(L != nullptr, "conv layer is NULL");
  BSDE_CHECK_ARG(L->ksize == 1 || L->ksize == 3, "ksize must be 1 or 3, got %d", L->ksize);
  BSDE_CHECK_ARG(L->sd == 1 || L->sd == 2, "sd must be 1 or 2, got %d", L->sd);
  BSDE_CHECK_ARG(L->cin > 0 && L->cout > 0 && L->hin > 0 && L->win > 0 && L->hout > 0 && L->wout > 0,
                 "conv layer has a non-positive extent");
  return 0;
}

int tconv_fwd_simt(const bsde_conv_layer* L, int S, int B, const float* in, const float* w,
                   long long w_sample_stride, float t, int act, int epi, float scale, const float* aux,
                   float* out, cudaStream_t st) {
  if (int e = check_layer(L)) return e;
  BSDE_CHECK_ARG(S > 0 && B > 0, "S and B must be positive");
  BSDE_CHECK_ARG(epi >= BSDE_EPI_BIAS && epi <= BSDE_EPI_ACCUM, "bad epilogue %d", epi);
  BSDE_CHECK_ARG(!(epi >= BSDE_EPI_EULER) || aux != nullptr, "epilogue %d needs aux", epi);
  ConvArgs a;
  a.in = in; a.w = w; a.aux = aux; a.out = out; a.S = S; a.B = B; a.w_sample_stride = w_sample_stride;
  a.cin = L->cin; a.cout = L->cout; a.ksize = L->ksize; a.hin = L->hin; a.win = L->win;
  a.hout = L->hout; a.wout = L->wout; a.sn = L->sn; a.kn = L->kn; a.off = L->off; a.sd = L->sd;
  a.has_time = L->has_time; a.time_row = L->time_first ? 0 : L->cin;
  a.real_base = (L->has_time && L->time_first) ? 1 : 0;
  a.w_off = L->w_off; a.b_off = L->b_off; a.ts = L->w_tap_stride; a.ks = L->w_k_stride; a.ns = L->w_n_stride;
  a.t = t; a.scale = scale; a.act = act; a.epi = epi;
  a.Kc = L->cin + (L->has_time ? 1 : 0);
  a.Ktot = L->ksize * L->ksize * a.Kc;
  a.Mper = B * L->hout * L->wout;
  if (L->cout > 16) {
    dim3 g((a.Mper + 127) / 128, (L->cout + 63) / 64, S);
    conv_gather_kernel<128, 64, 8, 4><<<g, 256, 0, st>>>(a);
  } else if (L->cout > 8) {
    dim3 g((a.Mper + 255) / 256, 1, S);
    conv_gather_kernel<256, 16, 4, 4><<<g, 256, 0, st>>>(a);
  } else {
    dim3 g((a.Mper + 255) / 256, 1, S);
    conv_gather_kernel<256, 8, 4, 2><<<g, 256, 0, st>>>(a);
  }
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

static void wgrad_plan(const bsde_conv_layer* L, int S, int B, int& rows, int& splits, int& chunk,
                       int& ti, int& tj, int& bi, int& bj) {
  const int Kc = L->cin + (L->has_time ? 1 : 0);
  rows = L->ksize * L->ksize * Kc + 1;
  if (L->cout > 16) { bi = 128; bj = 64; }
  else if (L->cout > 8) { bi = 256; bj = 16; }
  else { bi = 256; bj = 8; }
  ti = (rows + bi - 1) / bi;
  tj = (L->cout + bj - 1) / bj;
  const int Mper = B * L->hout * L->wout;
  int want = (148 * 4 + ti * tj * S - 1) / (ti * tj * S);
  int maxs = (Mper + 511) / 512;
  splits = want < 1 ? 1 : (want > maxs ? maxs : want);
  if (splits < 1) splits = 1;
  chunk = (Mper + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  splits = (Mper + chunk - 1) / chunk;
}

size_t tconv_wgrad_ws_simt(const bsde_conv_layer* L, int S, int B) {
  int rows, splits, chunk, ti, tj, bi, bj;
  wgrad_plan(L, S, B, rows, splits, chunk, ti, tj, bi, bj);
  return (size_t)S * splits * rows * L->cout * sizeof(float);
}

int tconv_wgrad_simt(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t,
                     float scale, float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  if (int e = check_layer(L)) return e;
  int rows, splits, chunk, ti, tj, bi, bj;
  wgrad_plan(L, S, B, rows, splits, chunk, ti, tj, bi, bj);
  const size_t need = (size_t)S * splits * rows * L->cout * sizeof(float);
  if (ws == nullptr || ws_bytes < need) {
    set_error("tconv_wgrad: workspace %zu < %zu bytes", ws_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  WgradArgs a;
  a.in = in; a.dy = dy; a.partial = (float*)ws; a.S = S; a.B = B; a.splits = splits; a.chunk = chunk;
  a.cin = L->cin; a.cout = L->cout; a.ksize = L->ksize; a.hin = L->hin; a.win = L->win;
  a.hout = L->hout; a.wout = L->wout; a.sn = L->sn; a.kn = L->kn; a.off = L->off; a.sd = L->sd;
  a.has_time = L->has_time; a.t = t;
  a.Kc = L->cin + (L->has_time ? 1 : 0);
  a.Ktot = L->ksize * L->ksize * a.Kc;
  a.rows = rows;
  a.Mper = B * L->hout * L->wout;
  dim3 g(ti, tj, S * splits);
  if (bi == 128) conv_wgrad_kernel<128, 64, 8, 4><<<g, 256, 0, st>>>(a);
  else if (bj == 16) conv_wgrad_kernel<256, 16, 4, 4><<<g, 256, 0, st>>>(a);
  else conv_wgrad_kernel<256, 8, 4, 2><<<g, 256, 0, st>>>(a);
  BSDE_CUDA_OK(cudaGetLastError());
  WreduceArgs r;
  r.partial = (const float*)ws; r.gw = gw; r.gw_sample_stride = gw_sample_stride; r.S = S; r.splits = splits;
  r.rows = rows; r.cout = L->cout; r.Kc = a.Kc; r.Ktot = a.Ktot; r.has_time = L->has_time;
  r.time_row = L->time_first ? 0 : L->cin; r.real_base = (L->has_time && L->time_first) ? 1 : 0;
  r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
  r.scale = scale;
  const long long total = (long long)S * rows * L->cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(r);
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace bsde
