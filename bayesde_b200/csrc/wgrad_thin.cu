// K3 weight gradient of the THIN 3x3 stride-1 layers of an SDEBNN block (5 -> 64 and 64 -> 5 at 32x32): one side of the
// layer has 64 channels (the "big" operand, 268 MB per launch at S*B = 1024 images), the other at most 7 (the "small" one).
//
//   G[tap][ci][co] = sum over positions q of  in(q + s_tap)[ci] * dY(q)[co]          s_tap = dy * pitch + dx
//
// wgrad_win_kernel serves these layers with three N = 96 MMAs per k-step in which 15 of 96 columns are real.  Here the nine
// taps go to the N side of ONE MMA per k-step, as in tconv_tapn.cu: the big operand is streamed UN-shifted through TMA
// (MN-major SWIZZLE_128B_BASE32B rows, one position = 128 bytes per 32-channel block: exactly how NHWC memory holds it),
// and the small operand is expanded on the fly into the shifted, tap-blocked matrix
//   case A (big = input,  64 -> cs):  Bm[q][tap * cs + co]       = dY(q - s_tap)[co]
//   case B (big = dY,     cs -> 64):  Bm[q][tap * (cs + 1) + ci] = in(q + s_tap)[ci],   column ci = cs: 1 at valid pixels
// built by eight warps from a shared-memory window of the small tensor into the K-major (no-swizzle) operand layout, so that
//   D[m][n] += sum_q big(q)[m] * Bm[q][n]           tcgen05.mma kind::tf32, M = 64, N = 64, K = 8 positions
// The ones column of case B yields the time-channel row (times t) and, for the centre tap, the bias row; in case A those
// rows only involve the small tensor and are accumulated on the CUDA cores by the threads that build Bm.
// CTAs split the positions of a sample; partial tiles go to the workspace in the layout of wgrad_reduce_kernel
// (tconv_simt.cu), which finishes in a fixed order (deterministic).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

namespace bsde {

using namespace um;

constexpr int WT_BUILD = 256;           // warps 0-7 build Bm (warps 0-3 also run the epilogue)
constexpr int WT_PROD_WARP = 8;         // TMA producer
constexpr int WT_MMA_WARP = 9;
constexpr int WT_THREADS = 10 * 32;
constexpr int WT_KS = 128;              // positions per stage
constexpr int WT_NPAD = 64;             // columns of Bm
constexpr int WT_ASTAGES = 3;
constexpr int WT_CW = 8;                // floats per window row: up to 7 channels + the valid flag

struct ThinArgs {
  const float* small;
  float* partial;
  int S, B, Hq, Wq, pitch, IP, Q, halo, NW;
  int SPL, nst_total;
  int cs, cpt, ncols, big_is_input;     // small channels; columns per tap (cs or cs + 1); real columns
  int NR, box_bytes, region, astage_bytes;
  int cin, cout, Kc, Ktot, rows, cols, has_time;
  float t;
  int delta[9];                          // shift of the small operand for tap j: -s_tap (case A) or +s_tap (case B)
};

__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_thin_kernel(const __grid_constant__ ThinArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / a.SPL, split = blockIdx.x - s * a.SPL;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  int soff = 0;
  const uint32_t s_a = base;                                       soff += WT_ASTAGES * a.astage_bytes;
  const uint32_t s_b = base + soff; float* g_b = reinterpret_cast<float*>(gen + soff);   soff += 2 * WT_KS * WT_NPAD * 4;
  float* s_win = reinterpret_cast<float*>(gen + soff);            soff += 2 * a.NW * WT_CW * 4;
  const uint32_t bars = base + soff;                              soff += 8 * (2 * WT_ASTAGES + 4 + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + soff);
  auto afull = [&](int st) { return bars + 8u * st; };
  auto aempty = [&](int st) { return bars + 8u * (WT_ASTAGES + st); };
  auto bfull = [&](int b) { return bars + 8u * (2 * WT_ASTAGES + b); };
  auto bempty = [&](int b) { return bars + 8u * (2 * WT_ASTAGES + 2 + b); };
  const uint32_t done_bar = bars + 8u * (2 * WT_ASTAGES + 4);

  // this CTA's contiguous range of stages (WT_KS positions each)
  const int per = a.nst_total / a.SPL, rem = a.nst_total - per * a.SPL;
  const int st0 = split * per + min(split, rem), nst = per + (split < rem ? 1 : 0);

  if (warp == WT_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < WT_ASTAGES; ++i) { mbar_init(afull(i), 1); mbar_init(aempty(i), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(bfull(b), WT_BUILD); mbar_init(bempty(b), 1); }
      mbar_init(done_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 64u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == WT_PROD_WARP) {
    // =============================== TMA producer: the big operand ===============================
    const int rows_img = a.Hq + 1;
#pragma unroll 1
    for (int it = 0; it < nst; ++it) {
      const int q0 = (st0 + it) * WT_KS;
      const int g0 = q0 / a.pitch + lane;
      const int img = g0 / rows_img;
      const int y = g0 - img * rows_img;
      const int st = it % WT_ASTAGES;
      if (it >= WT_ASTAGES) mbar_wait(aempty(st), ((it / WT_ASTAGES) & 1) ^ 1);
      const uint32_t sb = s_a + st * a.astage_bytes;
      if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(afull(st)), "r"(2 * a.NR * a.box_bytes) : "memory");
      __syncwarp();
      if (lane < a.NR) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const uint32_t dst = sb + (uint32_t)b * (uint32_t)a.region + (uint32_t)(lane * a.box_bytes);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
              "l"(&tmap), "r"(afull(st)), "r"(b * 32), "r"(0), "r"(y), "r"(s * a.B + img)
              : "memory");
        }
      }
    }
  } else if (warp == WT_MMA_WARP) {
    // =============================== MMA issuer ===============================
    // A: MN-major SWIZZLE_128B_BASE32B (transpose bit 15), two 32-channel blocks `region` apart, 8 positions = 1024 B per k-step;
    // B: K-major no-swizzle chunk planes [k >> 2][n][4 floats] (transpose bit 16 clear).  M = 64, N = 64, K = 8.
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(WT_NPAD >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    const uint64_t a_hi = desc_hi((uint32_t)a.region, 512u, 1);
    const uint64_t b_hi = desc_hi((uint32_t)WT_NPAD * 16u, 128u, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
#pragma unroll 1
    for (int it = 0; it < nst; ++it) {
      const int st = it % WT_ASTAGES, bb = it & 1;
      mbar_wait(afull(st), (it / WT_ASTAGES) & 1);
      mbar_wait(bfull(bb), (it >> 1) & 1);
      fence_proxy_async();
      tc_fence_after();
      const int q0 = (st0 + it) * WT_KS;
      const uint32_t oU = (uint32_t)(q0 % a.pitch) * 8u;
      const uint32_t sa16 = ((s_a + st * a.astage_bytes) >> 4) + oU;
      const uint32_t sb16 = (s_b + bb * (WT_KS * WT_NPAD * 4)) >> 4;
#pragma unroll 1
      for (int ks = 0; ks < WT_KS / 8; ++ks)
        umma_tf32_elect(tm, a_hi | (uint64_t)((sa16 + (uint32_t)ks * 64u) & 0x3FFFu),
                        b_hi | (uint64_t)((sb16 + (uint32_t)ks * (WT_NPAD * 2)) & 0x3FFFu), idesc, (it > 0 || ks > 0) ? 1u : 0u);
      umma_commit_elect(aempty(st));
      umma_commit_elect(bempty(bb));
    }
    umma_commit_elect(done_bar);
    __syncwarp();
  } else {
    // =============================== builders of Bm (warps 0-7) ===============================
    const float* __restrict__ small = a.small + (long long)s * a.B * a.Hq * a.Wq * a.cs;
    const float rIP = 1.0f / (float)a.IP, rpitch = 1.0f / (float)a.pitch;
    // window row `tid` (< NW) of a stage: small position q0 - halo + tid -> [c0 .. c_{cs-1}, 0.., valid flag at WT_CW - 1]
    auto load_row = [&](int q0, float (&v)[WT_CW]) {
#pragma unroll
      for (int c = 0; c < WT_CW; ++c) v[c] = 0.f;
      const int q = q0 - a.halo + tid;
      if (tid < a.NW && q >= 0 && q < a.Q) {
        int img = (int)((float)q * rIP);
        int r = q - img * a.IP;
        if (r < 0) { --img; r += a.IP; } else if (r >= a.IP) { ++img; r -= a.IP; }
        int y = (int)((float)r * rpitch);
        int x = r - y * a.pitch;
        if (x < 0) { --y; x += a.pitch; } else if (x >= a.pitch) { ++y; x -= a.pitch; }
        if (y < a.Hq && x < a.Wq) {
          const float* p = small + ((long long)(img * a.Hq + y) * a.Wq + x) * a.cs;
#pragma unroll
          for (int c = 0; c < WT_CW - 1; ++c) if (c < a.cs) v[c] = __ldg(p + c);
          v[WT_CW - 1] = 1.f;
        }
      }
    };
    // this thread's fixed part of Bm: column n (a warp covers 32 consecutive columns: conflict-free 16-byte stores) and the
    // k-quads kq, kq + 4, ... of every stage: four window reads feed ONE 16-byte store of the K-major chunk [n][k .. k + 3]
    // (the scalar form -- one 4-byte store per element -- made the builders the critical path: 62 % of the launch's
    // instructions on the two lines of the read and the store, ncu source page)
    const int n = (warp & 1) * 32 + lane, kq = warp >> 1;
    const int tap = n / a.cpt, c = n - tap * a.cpt;
    const bool real = n < a.ncols;
    // case B ones column: the valid flag of the shifted position
    const int wc = real ? ((!a.big_is_input && c == a.cs) ? WT_CW - 1 : c) : 0;
    const int wofs = real ? (a.halo + a.delta[tap]) : 0;
    float tsum = 0.f;   // case A: sum over valid big positions of Bm[.][n] -> time row (x t) / bias row
    float row[WT_CW];
    if (nst > 0) {
      load_row(st0 * WT_KS, row);
      if (tid < a.NW) {
#pragma unroll
        for (int cc = 0; cc < WT_CW; ++cc) s_win[tid * WT_CW + cc] = row[cc];
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(WT_BUILD) : "memory");
#pragma unroll 1
    for (int it = 0; it < nst; ++it) {
      const int bb = it & 1;
      const int q0 = (st0 + it) * WT_KS;
      if (it + 1 < nst) load_row(q0 + WT_KS, row);          // next stage's window row: in flight while this stage is built
      if (it >= 2) mbar_wait(bempty(bb), ((it >> 1) & 1) ^ 1);
      const float* win = s_win + bb * a.NW * WT_CW;
      float* bt = g_b + bb * (WT_KS * WT_NPAD);
      const float* wsrc = win + wofs * WT_CW + wc;
      const float* wflag = win + a.halo * WT_CW + WT_CW - 1;
      const int kreal = real ? a.Q - q0 : 0;               // positions k < kreal of this stage exist
#pragma unroll 4
      for (int i = 0; i < WT_KS / 16; ++i) {
        const int k = (kq + 4 * i) * 4;
        float4 v;
        v.x = k + 0 < kreal ? wsrc[(k + 0) * WT_CW] : 0.f;
        v.y = k + 1 < kreal ? wsrc[(k + 1) * WT_CW] : 0.f;
        v.z = k + 2 < kreal ? wsrc[(k + 2) * WT_CW] : 0.f;
        v.w = k + 3 < kreal ? wsrc[(k + 3) * WT_CW] : 0.f;
        // tile of k-step k >> 3: [chunk (k >> 2) & 1][n][4 floats]
        *reinterpret_cast<float4*>(bt + (k >> 3) * (WT_NPAD * 8) + (((k >> 2) & 1) * WT_NPAD + n) * 4) = v;
        if (a.big_is_input) {
          tsum = fmaf(wflag[(k + 0) * WT_CW], v.x, tsum); tsum = fmaf(wflag[(k + 1) * WT_CW], v.y, tsum);
          tsum = fmaf(wflag[(k + 2) * WT_CW], v.z, tsum); tsum = fmaf(wflag[(k + 3) * WT_CW], v.w, tsum);
        }
      }
      fence_proxy_async();
      mbar_arrive(bfull(bb));
      if (it + 1 < nst && tid < a.NW) {
        float* wn = s_win + (bb ^ 1) * a.NW * WT_CW;
#pragma unroll
        for (int cc = 0; cc < WT_CW; ++cc) wn[tid * WT_CW + cc] = row[cc];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(WT_BUILD) : "memory");   // the next window is complete; this one may be overwritten
    }

    // =============================== epilogue ===============================
    float* __restrict__ part = a.partial + ((long long)(s * a.SPL + split) * a.rows) * a.cols;
    if (a.big_is_input) {
      // the four k-quad partial sums of column n live in four warps: fixed-order sum through the (now free) window buffer
      s_win[kq * WT_NPAD + n] = tsum;
      asm volatile("bar.sync 1, %0;" ::"n"(WT_BUILD) : "memory");
      if (kq == 0 && real) {
        tsum = (s_win[n] + s_win[WT_NPAD + n]) + (s_win[2 * WT_NPAD + n] + s_win[3 * WT_NPAD + n]);
        if (a.has_time) part[(long long)(tap * a.Kc + a.cin) * a.cols + c] = a.t * tsum;
        if (tap == 4) part[(long long)a.Ktot * a.cols + c] = tsum;
      }
    }
    if (warp < 4) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int m = warp * 16 + lane;   // M = 64: accumulator rows 16 w .. 16 w + 15 live in lanes 0-15 of quarter w
#pragma unroll 1
      for (int n0 = 0; n0 < WT_NPAD; n0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
        if (lane < 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int nn = n0 + j;
            if (nn >= a.ncols) continue;
            const int tp = nn / a.cpt, cc = nn - tp * a.cpt;
            const float val = nst > 0 ? __uint_as_float(v[j]) : 0.f;
            if (a.big_is_input) {
              part[(long long)(tp * a.Kc + m) * a.cols + cc] = val;                       // m = input channel, cc = output channel
            } else if (cc < a.cs) {
              part[(long long)(tp * a.Kc + cc) * a.cols + m] = val;                       // m = output channel, cc = input channel
            } else {
              if (a.has_time) part[(long long)(tp * a.Kc + a.cin) * a.cols + m] = a.t * val;
              if (tp == 4) part[(long long)a.Ktot * a.cols + m] = val;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WT_MMA_WARP) tmem_dealloc(tmem_base, 64u);
}

// ---- host side ----------------------------------------------------------------------------------------------
struct ThinPlan {
  bool ok;
  int big_is_input, cs, cpt, ncols, Hq, Wq, pitch, IP, halo, NW, NR, box_bytes, region, astage_bytes, Kc, Ktot, rows, cols;
  size_t smem;
};

static ThinPlan thin_plan(const bsde_conv_layer* L) {
  ThinPlan p;
  memset(&p, 0, sizeof(p));
  if (!L || L->ksize != 3 || L->sd != 1 || L->sn != 1 || L->kn != 1 || L->off != -1) return p;   // forward stride-1 SAME 3x3
  if (L->hin != L->hout || L->win != L->wout) return p;
  if (L->has_time && L->time_first) return p;                          // JAX front end layout (time row last)
  if (L->cin == 64 && L->cout >= 1 && 9 * L->cout <= WT_NPAD && L->cout <= WT_CW - 1) {
    p.big_is_input = 1; p.cs = L->cout; p.cpt = L->cout;
  } else if (L->cout == 64 && L->cin >= 1 && 9 * (L->cin + 1) <= WT_NPAD && L->cin <= WT_CW - 1 && L->has_time) {
    p.big_is_input = 0; p.cs = L->cin; p.cpt = L->cin + 1;
  } else {
    return p;
  }
  p.ncols = 9 * p.cpt;
  p.Hq = L->hout; p.Wq = L->wout; p.pitch = p.Wq + 1; p.IP = (p.Hq + 1) * p.pitch;
  p.halo = p.pitch + 1;
  p.NW = WT_KS + 2 * p.halo;
  if (p.NW > WT_BUILD || p.pitch > 256) return p;
  p.NR = (WT_KS + p.pitch - 2) / p.pitch + 1;
  if (p.NR > 32) return p;
  p.box_bytes = p.pitch * 128;
  p.region = (p.NR * p.box_bytes + 1023) & ~1023;
  p.astage_bytes = 2 * p.region;
  p.Kc = L->cin + (L->has_time ? 1 : 0);
  p.Ktot = 9 * p.Kc;
  p.rows = p.Ktot + 1;
  p.cols = L->cout;
  p.smem = 1024 + (size_t)WT_ASTAGES * p.astage_bytes + 2 * WT_KS * WT_NPAD * 4 + (size_t)2 * p.NW * WT_CW * 4 +
           8 * (2 * WT_ASTAGES + 5) + 16;
  if (p.smem > 227 * 1024) return p;
  p.ok = true;
  return p;
}

static int thin_splits(const ThinPlan& p, int S, int B) {
  const int nst_total = (B * p.IP + WT_KS - 1) / WT_KS;
  return std::max(1, std::min(sm_budget() / S, nst_total));
}

bool tconv_wgrad_thin_supported(const bsde_conv_layer* L, int S, int B) {
  static const int on = [] { const char* e = getenv("BSDE_WTHIN"); return e ? atoi(e) : 1; }();
  if (!on || !L || S < 1 || S > 148 || B < 1) return false;
  ThinPlan p = thin_plan(L);
  if (!p.ok) return false;
  if ((long long)B * p.IP + 1024 > (1LL << 24)) return false;
  if ((long long)B * L->hin * L->win * 64 > 0x7FFFFFFFLL) return false;
  return true;
}

size_t tconv_wgrad_ws_thin(const bsde_conv_layer* L, int S, int B) {
  ThinPlan p = thin_plan(L);
  if (!p.ok) return 0;
  return (size_t)S * thin_splits(p, S, B) * p.rows * p.cols * sizeof(float);
}

int tconv_wgrad_thin(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t, float scale,
                     float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes, cudaStream_t st) {
  ThinPlan p = thin_plan(L);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the thin weight-gradient path");
  const size_t need = tconv_wgrad_ws_thin(L, S, B);
  if (!ws || ws_bytes < need) {
    set_error("tconv_wgrad(thin): workspace %zu < %zu bytes", ws_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  ThinArgs a;
  memset(&a, 0, sizeof(a));
  const float* big = p.big_is_input ? in : dy;
  a.small = p.big_is_input ? dy : in;
  a.partial = (float*)ws;
  a.S = S; a.B = B; a.Hq = p.Hq; a.Wq = p.Wq; a.pitch = p.pitch; a.IP = p.IP; a.Q = B * p.IP; a.halo = p.halo; a.NW = p.NW;
  a.nst_total = (a.Q + WT_KS - 1) / WT_KS;
  a.SPL = thin_splits(p, S, B);
  a.cs = p.cs; a.cpt = p.cpt; a.ncols = p.ncols; a.big_is_input = p.big_is_input;
  a.NR = p.NR; a.box_bytes = p.box_bytes; a.region = p.region; a.astage_bytes = p.astage_bytes;
  a.cin = L->cin; a.cout = L->cout; a.Kc = p.Kc; a.Ktot = p.Ktot; a.rows = p.rows; a.cols = p.cols; a.has_time = L->has_time;
  a.t = t;
  for (int tap = 0; tap < 9; ++tap) {
    const int sh = (tap / 3 - 1) * p.pitch + (tap % 3 - 1);
    a.delta[tap] = p.big_is_input ? -sh : sh;
  }
  static PerDeviceOnce once;
  if (once.first()) BSDE_CUDA_OK(cudaFuncSetAttribute(wgrad_thin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult qres;
    BSDE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return BSDE_ERR_CUDA; }
  }
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  const cuuint64_t dims[4] = {64, (cuuint64_t)p.Wq, (cuuint64_t)p.Hq, (cuuint64_t)S * B};
  const cuuint64_t strides[3] = {(cuuint64_t)64 * 4, (cuuint64_t)p.Wq * 64 * 4, (cuuint64_t)p.Hq * p.Wq * 64 * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)p.pitch, 1, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(big), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return BSDE_ERR_CUDA; }
  wgrad_thin_kernel<<<dim3(S * a.SPL), dim3(WT_THREADS), p.smem, st>>>(a, tmap);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();

  WreduceArgs rd;
  memset(&rd, 0, sizeof(rd));
  rd.partial = a.partial; rd.bias_partial = nullptr; rd.gw = gw; rd.gw_sample_stride = gw_sample_stride;
  rd.S = S; rd.splits = a.SPL; rd.rows = a.rows; rd.cols = a.cols; rd.Kc = a.Kc; rd.Ktot = a.Ktot;
  rd.has_time = L->has_time; rd.time_row = L->time_first ? 0 : L->cin;
  rd.real_base = (L->has_time && L->time_first) ? 1 : 0; rd.swapped = 0; rd.cin = L->cin; rd.cout = L->cout;
  rd.w_off = L->w_off; rd.b_off = L->b_off; rd.ts = L->w_tap_stride; rd.ks = L->w_k_stride; rd.ns = L->w_n_stride;
  rd.scale = scale; rd.nbias = 1;
  const long long total = (long long)S * a.rows * a.cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  return launch_wgrad_reduce(rd, blocks, st);
}

}  // namespace bsde
