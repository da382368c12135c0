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
#include <cstring>

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

// valid taps along one axis for output parity `par` (sd == 2), or all taps (sd == 1)
__device__ __forceinline__ int valid_taps(int ksize, int sn, int kn, int off, int sd, int par, int* kv) {
  int n = 0;
  for (int k = 0; k < ksize; ++k)
    if (sd == 1 || (((par * sn + k * kn + off) & 1) == 0)) kv[n++] = k;
  return n;
}

// When sd == 2 (lhs-dilated "transposed" conv, or the data gradient of a stride-2 conv) only the taps
// whose parity matches the output pixel's parity touch real data.  Output pixels are therefore
// processed per parity class (blockIdx.y selects the class): each class is a dense conv over its
// 1, 2 or 4 valid taps, so no multiply-by-structural-zero work is issued.
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) conv_gather_kernel(ConvArgs a) {
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  constexpr int PPT = BM / 16;  // pixels gathered per thread per k-tile
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int s = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int ncls = a.sd == 2 ? 4 : 1;
  const int cls = blockIdx.y % ncls;
  const int n0 = (blockIdx.y / ncls) * BN;
  const int py = cls >> 1, px = cls & 1;
  const int step = a.sd == 2 ? 2 : 1;
  const int hq = (a.hout - py + step - 1) / step, wq = (a.wout - px + step - 1) / step;
  const int Mcls = a.B * hq * wq;
  if (m0 >= Mcls) return;
  int khv[3], kwv[3];
  const int nkh = valid_taps(a.ksize, a.sn, a.kn, a.off, a.sd, py, khv);
  const int nkw = valid_taps(a.ksize, a.sn, a.kn, a.off, a.sd, px, kwv);
  const int Ktot = nkh * nkw * a.Kc;
  const float* __restrict__ in = a.in + (long long)s * a.B * a.hin * a.win * a.cin;
  const float* __restrict__ w = a.w + (long long)s * a.w_sample_stride;

  // pixel decode for the gather role
  const int kl = tid % BK;
  const int ml0 = tid / BK;
  int p_oy[PPT], p_ox[PPT], p_img[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    int m = m0 + ml0 + i * 16;
    if (m < Mcls) {
      int hw = hq * wq;
      int img = m / hw;
      int r = m - img * hw;
      int qy = r / wq;
      p_img[i] = img;
      p_oy[i] = qy * step + py;
      p_ox[i] = (r - qy * wq) * step + px;
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

  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    // ---- gather A tile
    {
      const int k = k0 + kl;
      const bool kvalid = k < Ktot;
      int tl = 0, c = 0;
      if (kvalid) { tl = k / a.Kc; c = k - tl * a.Kc; }
      const int kh = khv[tl / nkw], kw = kwv[tl % nkw];
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
      if (k < Ktot && n < a.cout) {
        const int tl = k / a.Kc, c = k - tl * a.Kc;
        const int tap = khv[tl / nkw] * a.ksize + kwv[tl % nkw];
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
    if (m >= Mcls) continue;
    long long pix = m;
    if (a.sd == 2) {
      const int hw = hq * wq;
      const int img = m / hw;
      const int r = m - img * hw;
      const int qy = r / wq;
      pix = ((long long)img * a.hout + (qy * 2 + py)) * a.wout + ((r - qy * wq) * 2 + px);
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= a.cout) continue;
      const long long o = pix * a.cout + n;
      float v = acc[i][j];
      const float b = (a.b_off >= 0 && a.epi <= BSDE_EPI_EULER) ? __ldg(w + a.b_off + n) : 0.f;
      switch (a.epi) {
        case BSDE_EPI_BIAS: v = v + b; break;
        case BSDE_EPI_BIAS_ACT: v = a.act == BSDE_ACT_RBF ? rbf_fwd_tagged(v + b) : act_fwd(a.act, v + b); break;
        case BSDE_EPI_EULER: v = aux[o] + a.scale * (v + b); break;
        case BSDE_EPI_DACT: v = a.scale * v * act_grad_from_out(a.act, aux[o]); break;
        case BSDE_EPI_ACCUM: v = aux[o] + a.scale * v; break;
      }
      out[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient:  G[i][j] = sum_m A(m, i) * Bm[m][j]
//   direct form  (forward layers with sd == 1): m runs over OUTPUT pixels, A = gathered input
//       (rows i = (tap, channel incl. time) plus one trailing "ones" row = bias gradient), Bm = dy.
//   swapped form (forward layers with sd == 2, i.e. lhs-dilated convs): m runs over INPUT pixels,
//       A = dy gathered through the data-gradient map (rows i = (tap, cout)), Bm = [x, t]
//       (columns j = input channel incl. time); G is stored transposed.  This keeps the pixel loop
//       dense: the direct form would multiply 3/4 structural zeros.  The bias gradient is a
//       separate column sum.
// Pixels are split across CTAs; partial tiles go to the workspace and a second kernel reduces
// them in a fixed order (deterministic, no float atomics).
// ---------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* in;   // gathered operand
  const float* bm;   // plain operand [pixels][bcols]
  float* partial;    // [S][splits][rows][cols]
  int S, B, splits, chunk;
  int cin, ksize, hin, win, hout, wout, sn, kn, off, sd;  // gather geometry; (hout,wout) = pixel loop space
  int has_time;      // gathered side carries the time channel (direct form)
  int bias_row;      // trailing ones row (direct form)
  int bcols, b_time; // plain operand: real columns, +1 constant-t column (swapped form)
  float t;
  int Kc, Ktot, rows, cols, Mper;
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
  const float* __restrict__ bm = a.bm + (long long)s * a.Mper * a.bcols;
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
  } else if (i_row == a.Ktot && a.bias_row) {
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
      float v = 0.f;
      if (m < mend) {
        if (j < a.bcols) v = __ldg(bm + (long long)m * a.bcols + j);
        else if (j == a.bcols && a.b_time) v = a.t;
      }
      Bs[mm][jj] = v;
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
  float* __restrict__ part = a.partial + ((long long)(s * a.splits + split) * a.rows) * a.cols;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = i0 + ty * TM + i;
    if (row >= a.rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = j0 + tx * TN + j;
      if (col < a.cols) part[(long long)row * a.cols + col] = acc[i][j];
    }
  }
}

// column sums of dy per (sample, pixel chunk): partial[s][split][c]   (bias gradient, swapped form)
__global__ void colsum_partial_kernel(const float* __restrict__ dy, float* __restrict__ partial, int Mper, int cout,
                                      int splits, int chunk) {
  const int s = blockIdx.y, split = blockIdx.x;
  const int mbeg = split * chunk, mend = min(Mper, mbeg + chunk);
  const float* __restrict__ d = dy + (long long)s * Mper * cout;
  __shared__ float red[256];
  // threads: column c = tid % cpad, pixel lane = tid / cpad
  for (int c0 = 0; c0 < cout; c0 += 64) {
    const int cw = min(64, cout - c0);
    const int c = threadIdx.x % 64, lane = threadIdx.x / 64;  // 4 pixel lanes
    float v = 0.f;
    if (c < cw)
      for (int m = mbeg + lane; m < mend; m += 4) v += __ldg(d + (long long)m * cout + c0 + c);
    red[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x < 64 && c < cw)
      partial[((long long)s * splits + split) * cout + c0 + c] = red[c] + red[64 + c] + red[128 + c] + red[192 + c];
    __syncthreads();
  }
}


__global__ void wgrad_reduce_kernel(WreduceArgs a) {
  // launched with programmatic dependent launch from the tensor-core weight-gradient paths: wait for the partial tiles
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long per = (long long)a.rows * a.cols;
  const long long total = per * a.S;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(idx / per);
    const long long e = idx - s * per;
    const int row = (int)(e / a.cols), col = (int)(e - (long long)row * a.cols);
    const float* p = a.partial + (long long)s * a.splits * per + e;
    float v = 0.f;
    for (int q = 0; q < a.splits; ++q) v += p[q * per];
    v *= a.scale;
    float* g = a.gw + s * a.gw_sample_stride;
    if (!a.swapped) {
      if (row > a.Ktot) continue;  // extra bias rows are folded into row Ktot below
      if (row == a.Ktot) {
        for (int b = 1; b < a.nbias; ++b) {
          float vb = 0.f;
          for (int q = 0; q < a.splits; ++q) vb += p[q * per + (long long)b * a.cols];
          v += vb * a.scale;
        }
        if (a.b_off >= 0) g[a.b_off + col] = v;
      } else {
        const int tap = row / a.Kc, c = row - tap * a.Kc;
        const int wr = (a.has_time && c == a.Kc - 1) ? a.time_row : c + a.real_base;
        g[a.w_off + tap * a.ts + wr * a.ks + col * a.ns] = v;
      }
    } else {  // row = (tap, cout), col = input channel (last = time)
      const int tap = row / a.cout, co = row - tap * a.cout;
      const int wr = (a.has_time && col == a.cin) ? a.time_row : col + a.real_base;
      g[a.w_off + tap * a.ts + wr * a.ks + co * a.ns] = v;
    }
  }
  if (a.swapped && a.bias_partial && a.b_off >= 0) {
    const long long tb = (long long)a.S * a.cout;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tb;
         idx += (long long)gridDim.x * blockDim.x) {
      const int s = (int)(idx / a.cout), c = (int)(idx - (long long)s * a.cout);
      float v = 0.f;
      for (int q = 0; q < a.splits; ++q) v += a.bias_partial[((long long)s * a.splits + q) * a.cout + c];
      a.gw[s * a.gw_sample_stride + a.b_off + c] = v * a.scale;
    }
  }
}

int launch_wgrad_reduce(const WreduceArgs& a, int blocks, cudaStream_t st) {
  BSDE_CUDA_OK(launch_pdl(wgrad_reduce_kernel, dim3(blocks), dim3(256), 0, st, a));
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
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
  BSDE_CHECK_ARG(L->sd == 1 || (L->sn & 1), "sd=2 needs an odd sn");
  const int ncls = L->sd == 2 ? 4 : 1;
  const int Mcls = L->sd == 2 ? B * ((L->hout + 1) / 2) * ((L->wout + 1) / 2) : a.Mper;  // largest class
  if (L->cout > 16) {
    dim3 g((Mcls + 127) / 128, ((L->cout + 63) / 64) * ncls, S);
    conv_gather_kernel<128, 64, 8, 4><<<g, 256, 0, st>>>(a);
  } else if (L->cout > 8) {
    dim3 g((Mcls + 255) / 256, ncls, S);
    conv_gather_kernel<256, 16, 4, 4><<<g, 256, 0, st>>>(a);
  } else {
    dim3 g((Mcls + 255) / 256, ncls, S);
    conv_gather_kernel<256, 8, 4, 2><<<g, 256, 0, st>>>(a);
  }
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

struct WgradPlan {
  int swapped, rows, cols, splits, chunk, ti, tj, bi, bj, Mper;
  size_t main_floats, bias_floats;
};

static WgradPlan wgrad_plan(const bsde_conv_layer* L, int S, int B) {
  WgradPlan p;
  const int taps = L->ksize * L->ksize;
  p.swapped = L->sd == 2;
  if (!p.swapped) {
    p.rows = taps * (L->cin + (L->has_time ? 1 : 0)) + 1;
    p.cols = L->cout;
    p.Mper = B * L->hout * L->wout;
  } else {
    p.rows = taps * L->cout;
    p.cols = L->cin + (L->has_time ? 1 : 0);
    p.Mper = B * L->hin * L->win;
  }
  if (p.cols > 16) { p.bi = 128; p.bj = 64; }
  else if (p.cols > 8) { p.bi = 256; p.bj = 16; }
  else { p.bi = 256; p.bj = 8; }
  p.ti = (p.rows + p.bi - 1) / p.bi;
  p.tj = (p.cols + p.bj - 1) / p.bj;
  int want = (148 * 4 + p.ti * p.tj * S - 1) / (p.ti * p.tj * S);
  int maxs = (p.Mper + 511) / 512;
  p.splits = want < 1 ? 1 : (want > maxs ? maxs : want);
  if (p.splits < 1) p.splits = 1;
  p.chunk = (p.Mper + p.splits - 1) / p.splits;
  p.chunk = (p.chunk + BK - 1) / BK * BK;
  p.splits = (p.Mper + p.chunk - 1) / p.chunk;
  p.main_floats = (size_t)S * p.splits * p.rows * p.cols;
  p.bias_floats = p.swapped ? (size_t)S * 64 * L->cout : 0;
  return p;
}

size_t tconv_wgrad_ws_simt(const bsde_conv_layer* L, int S, int B) {
  WgradPlan p = wgrad_plan(L, S, B);
  return (p.main_floats + p.bias_floats) * sizeof(float);
}

int tconv_wgrad_simt(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t,
                     float scale, float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  if (int e = check_layer(L)) return e;
  WgradPlan p = wgrad_plan(L, S, B);
  const size_t need = (p.main_floats + p.bias_floats) * sizeof(float);
  if (ws == nullptr || ws_bytes < need) {
    set_error("tconv_wgrad: workspace %zu < %zu bytes", ws_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  WgradArgs a;
  a.partial = (float*)ws; a.S = S; a.B = B; a.splits = p.splits; a.chunk = p.chunk;
  a.ksize = L->ksize; a.t = t; a.rows = p.rows; a.cols = p.cols; a.Mper = p.Mper;
  if (!p.swapped) {
    a.in = in; a.bm = dy;
    a.cin = L->cin; a.hin = L->hin; a.win = L->win; a.hout = L->hout; a.wout = L->wout;
    a.sn = L->sn; a.kn = L->kn; a.off = L->off; a.sd = L->sd;
    a.has_time = L->has_time; a.bias_row = 1; a.bcols = L->cout; a.b_time = 0;
    a.Kc = L->cin + (L->has_time ? 1 : 0);
  } else {
    // gather dy through the data-gradient map, looping over the forward layer's input pixels
    a.in = dy; a.bm = in;
    a.cin = L->cout; a.hin = L->hout; a.win = L->wout; a.hout = L->hin; a.wout = L->win;
    a.sn = L->sd; a.kn = -L->kn; a.off = -L->off; a.sd = L->sn;
    BSDE_CHECK_ARG(a.sd == 1, "wgrad: forward layer with sd=2 needs sn=1");
    a.has_time = 0; a.bias_row = 0; a.bcols = L->cin; a.b_time = L->has_time;
    a.Kc = L->cout;
  }
  a.Ktot = L->ksize * L->ksize * a.Kc;
  dim3 g(p.ti, p.tj, S * p.splits);
  if (p.bi == 128) conv_wgrad_kernel<128, 64, 8, 4><<<g, 256, 0, st>>>(a);
  else if (p.bj == 16) conv_wgrad_kernel<256, 16, 4, 4><<<g, 256, 0, st>>>(a);
  else conv_wgrad_kernel<256, 8, 4, 2><<<g, 256, 0, st>>>(a);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  float* bias_partial = nullptr;
  int bsplits = 0;
  if (p.swapped && L->b_off >= 0) {
    bias_partial = (float*)ws + p.main_floats;
    const int Mout = B * L->hout * L->wout;
    bsplits = 64;
    int bchunk = (Mout + bsplits - 1) / bsplits;
    if (bchunk < 1) bchunk = 1;
    bsplits = (Mout + bchunk - 1) / bchunk;
    colsum_partial_kernel<<<dim3(bsplits, S), 256, 0, st>>>(dy, bias_partial, Mout, L->cout, bsplits, bchunk);
    BSDE_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  WreduceArgs r;
  memset(&r, 0, sizeof(r));
  r.partial = (const float*)ws; r.bias_partial = bias_partial; r.gw = gw; r.gw_sample_stride = gw_sample_stride;
  r.S = S; r.splits = p.splits; r.rows = p.rows; r.cols = p.cols; r.Kc = a.Kc; r.Ktot = a.Ktot;
  r.has_time = L->has_time; r.time_row = L->time_first ? 0 : L->cin;
  r.real_base = (L->has_time && L->time_first) ? 1 : 0; r.swapped = p.swapped; r.cin = L->cin; r.cout = L->cout;
  r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
  r.scale = scale;
  const long long total = (long long)S * p.rows * p.cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (bias_partial) {
    // bias partials use their own split count: reduce them with a dedicated tiny launch
    WreduceArgs rb = r;
    rb.partial = nullptr; rb.rows = 0; rb.cols = 1; rb.splits = bsplits;
    wgrad_reduce_kernel<<<(S * L->cout + 255) / 256, 256, 0, st>>>(rb);
    BSDE_CUDA_OK(cudaGetLastError());
    count_launch();
    r.bias_partial = nullptr;
  }
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(r);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bsde
