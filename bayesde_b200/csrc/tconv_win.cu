// K2 (tensor-core mode, window-staged): the time-dependent convolution as a persistent implicit GEMM
// whose A operand is never gathered per tap.
//
// Positions.  The output grid of one sample (B images of Hq x Wq) is flattened with ONE zero pad column per
// row and ONE zero pad row per image:  q = img*IP + y*(Wq+1) + x,  IP = (Hq+1)*(Wq+1).  A tile is any 128
// consecutive q (one accumulator row each); a tap displaced by (sy, sx) reads position q + sy*(Wq+1) + sx,
// and the pad column / pad row supply the zero padding of every border.  So the A operand of every tap is
// the SAME staged window, addressed through a shared-memory descriptor whose start address is shifted by
// `shift` rows (tools/umma_probe2.cu: the 128B swizzle is a function of absolute smem address bits, so
// row-shifted starts are exact with base_offset = 0).
//
// Shared memory:
//   * window: per 32-channel group ("half") one K-major SWIZZLE_128B region [NPOSw rows][128 B]; the four
//     producer warps fill it with 16-byte cp.async (zero-fill at pads), 8 lanes per row, so that BOTH sides of
//     an LDGSTS are whole 128-byte lines (4 global lines + 4 smem rows per warp instruction).  Every input
//     element is read ONCE per tile.  Slabs (all halves of a tile, or one half) cycle through a ring.
//   * weights of the whole layer, repacked once per (sample, iteration) by win_repack_kernel into K-major
//     no-swizzle chunk planes, stay resident for all tiles of the CTA (<= 184 KB);
//   * stride-2 convolutions stage the four input parity planes (taps pick plane + shift); lhs-dilated
//     (transposed) convolutions keep four accumulators, one per output parity class, over one window.
// One warp issues tcgen05.mma.kind::tf32 (M=128, N=npad, K=8) per (k-step, tap) from descriptor tables;
// accumulators are double buffered in TMEM so the eight epilogue warps overlap the next tile's MMAs:
// tcgen05.ld -> bias / time term / activation in registers (one accumulator row per lane) -> warp-private
// swizzled smem transpose -> coalesced Euler / data-gradient combine and store.
// The time channel never enters the GEMM (spatially constant): t * (sum of the time rows of the taps that
// fall inside the image), looked up from a 16-entry border-class table per accumulator.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

namespace bsde {

using namespace um;

constexpr int WN_EPI_WARPS = 8;   // warps 0-7 epilogue (TMEM lane quarter = warp & 3)
constexpr int WN_PROD = 256;      // warps 8-15 producers
constexpr int WN_MMA_WARP = 16;   // warp 16 MMA issuer + TMEM owner
constexpr int WN_THREADS = 17 * 32;
constexpr int WN_MAXENT = 16;
constexpr int WN_MAXRING = 6;
constexpr int WN_MAXBUF = 4;       // TMEM accumulator buffers

struct WinEntry {
  int acc, plane, shift, kh, kw;
};

struct WinArgs {
  const float* in;
  const float* aux;
  float* out;
  const float* wimg;
  long long wimg_sample_stride;  // floats
  int S, B, G, ntiles;
  int cin, cout, npad, nks, nhalf;
  int opitch;  // floats per output pixel in `out` / `aux` (> cout when this launch produces a channel slice of the layer)
  int hin, win, hout, wout, sn, kn, off, sd, has_time, mode;
  int Hq, Wq, pitch, IP, Q, lead, NPOSw, NP, NA, nent;
  int ring, slab_bytes, w_bytes, tb_floats, nbuf, tmem_cols;
  int hslab, nslab, region;  // halves per slab, slabs per tile, bytes of one (plane, half) region
  int gather;                // stride-2 layers: slabs are per-(tap, half) gathered tiles of 128 rows (no window, no pads)
  int tma, NR, box_bytes;    // TMA-fed window: NR row boxes {32 ch, Wq+1 px} per (tile, half), each box_bytes
  float t, scale;
  uint32_t aoff[WN_MAXENT];  // A start offset of each entry in 16-byte units (plane region + shifted row)
  uint32_t dcol[WN_MAXENT];  // accumulator column offset | (first tap of its accumulator) << 31
  int rev;    // walk the tiles from the last to the first: a consumer then starts on what its producer wrote last (still in L2)
  int pf;     // epilogue: L2 prefetch of the next tile's combine operand (BSDE_WIN_PF, default on)
  int wsafe;  // the weight image was written before the previous kernel in the stream started: load it before pdl_wait()
  int dbg;  // tools/ only (BSDE_WIN_DBG): 1 skip window copies, 2 skip MMAs, 4 skip epilogue stores
  WinEntry ent[WN_MAXENT];
};

__device__ __host__ __forceinline__ bool wn_map_axis(int o, int k, int sn, int kn, int off, int sd, int size) {
  int num = o * sn + k * kn + off;
  if (num < 0) return false;
  if (sd == 2) {
    if (num & 1) return false;
    num >>= 1;
  }
  return num < size;
}

// ---- weight repack: flat w(t_k) -> the CTA's resident smem image --------------------------------------
//   tiles  [(entry e, k-step ks)][chunk c in 0..1][n in 0..npad)[4 floats]   (K-major chunk planes)
//   tables [0] bias; with a time channel: [1 + acc*16 + cls] = sum of the time rows of the taps of accumulator
//          `acc` that are inside the image for border class cls = fy*4 + fx,
//          fy = (y == 0) | (y == Hq-1) << 1 in position space (fx alike)
struct WinRepackArgs {
  const float* w;
  float* img;
  long long w_sample_stride, img_sample_stride, w_iter_stride, img_iter_stride;  // blockIdx.z = scan iteration
  int nent, nks, npad, cin, cout, ksize, NA, real_base, has_time, time_row, use_bias, w_floats, tb_rows;
  int mode, Hq, Wq, hin, win, sn, kn, off, sd;
  long long w_off, b_off, ts, ks, ns;
  int ent_tap[WN_MAXENT], ent_acc[WN_MAXENT], ent_kh[WN_MAXENT], ent_kw[WN_MAXENT];
};

__device__ __forceinline__ int wn_class_rep(int f, int size) { return f == 0 ? 1 : (f == 2 ? size - 1 : 0); }

__global__ void win_repack_kernel(WinRepackArgs a) {
  const int s = blockIdx.y;
  const float* __restrict__ w = a.w + (long long)s * a.w_sample_stride + (long long)blockIdx.z * a.w_iter_stride;
  float* __restrict__ img = a.img + (long long)s * a.img_sample_stride + (long long)blockIdx.z * a.img_iter_stride;
  if ((int)blockIdx.x == a.nent * a.nks) {
    float* tb = img + a.w_floats;
    for (int idx = threadIdx.x; idx < a.tb_rows * a.npad; idx += blockDim.x) {
      const int r = idx / a.npad, n = idx - r * a.npad;
      float v = 0.f;
      if (n < a.cout) {
        if (r == 0) {
          v = (a.b_off >= 0 && a.use_bias) ? __ldg(w + a.b_off + n) : 0.f;
        } else {
          const int ac = (r - 1) >> 4, cls = (r - 1) & 15;
          const int y = wn_class_rep(cls >> 2, a.Hq), x = wn_class_rep(cls & 3, a.Wq);
          const int oy = a.mode == 3 ? 2 * y + (ac >> 1) : y, ox = a.mode == 3 ? 2 * x + (ac & 1) : x;
          for (int e = 0; e < a.nent; ++e) {
            if (a.ent_acc[e] != ac) continue;
            if (!wn_map_axis(oy, a.ent_kh[e], a.sn, a.kn, a.off, a.sd, a.hin) ||
                !wn_map_axis(ox, a.ent_kw[e], a.sn, a.kn, a.off, a.sd, a.win))
              continue;
            v += __ldg(w + a.w_off + a.ent_tap[e] * a.ts + (long long)a.time_row * a.ks + (long long)n * a.ns);
          }
        }
      }
      tb[idx] = v;
    }
    return;
  }
  const int e = blockIdx.x / a.nks, ks = blockIdx.x - e * a.nks;
  float* __restrict__ tile = img + (long long)blockIdx.x * a.npad * 8;
  const int tap = a.ent_tap[e];
  for (int idx = threadIdx.x; idx < a.npad * 8; idx += blockDim.x) {
    const int n = idx % a.npad, kk = idx / a.npad;
    const int c = ks * 8 + kk;
    float v = 0.f;
    if (c < a.cin && n < a.cout)
      v = __ldg(w + a.w_off + tap * a.ts + (long long)(c + a.real_base) * a.ks + (long long)n * a.ns);
    tile[((kk >> 2) * a.npad + n) * 4 + (kk & 3)] = v;
  }
}

// ---- main kernel -----------------------------------------------------------------------------------------
__device__ __forceinline__ float wn_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float wn_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float wn_act_fwd(float x) {
  if (ACT == BSDE_ACT_SOFTPLUS) {
    // softplus(x) = max(x, 0) + log1p(e), e = exp(-|x|) in (0, 1]: one MUFU (ex2); log1p(e) = e * p(e) with a
    // degree-4 p (max abs error 1.2e-5 on [0, 1], below the TF32 rounding of the consumer) on the FMA pipe -- the
    // epilogue is instruction-issue / MUFU bound (ncu: issue active 58 %), lg2 + ex2 cost 2 MUFU per element
    const float e = wn_ex2(-1.44269504f * fabsf(x));
    float p = 0.03137759119272232f;
    p = fmaf(p, e, -0.134135439991951f);
    p = fmaf(p, e, 0.2878262996673584f);
    p = fmaf(p, e, -0.49134793877601624f);
    p = fmaf(p, e, 0.9994350075721741f);
    return fmaf(p, e, fmaxf(x, 0.f));
  }
  if (ACT == BSDE_ACT_TANH) return tanhf(x);
  if (ACT == BSDE_ACT_ELU) return x > 0.f ? x : wn_ex2(1.44269504f * x) - 1.f;
  if (ACT == BSDE_ACT_RBF) return rbf_fwd_tagged(x);
  return x;
}
template <int ACT>
__device__ __forceinline__ float wn_act_grad_out(float a) {
  if (ACT == BSDE_ACT_SOFTPLUS) return 1.f - wn_ex2(-1.44269504f * a);
  if (ACT == BSDE_ACT_TANH) return 1.f - a * a;
  if (ACT == BSDE_ACT_ELU) return a > 0.f ? 1.f : a + 1.f;
  if (ACT == BSDE_ACT_RBF) return rbf_grad_from_tagged(a);
  return 1.f;
}
// second half of the epilogue (needs the value already stored at the output position)
template <int EPI, int ACT>
__device__ __forceinline__ float wn_combine(float q, float x, float scale) {
  if (EPI == BSDE_EPI_DACT) return scale * q * wn_act_grad_out<ACT>(x);
  if (EPI == BSDE_EPI_EULER || EPI == BSDE_EPI_ACCUM) return fmaf(scale, q, x);
  return q;
}

// ---- packed pairs: Blackwell's FFMA2 / FADD2 / FMUL2 carry two fp32 operations per issue slot (same IEEE results as the scalar
// forms: the epilogues are issue bound with two warps per scheduler, not FP32-pipe bound)
__device__ __forceinline__ float2 wn_f2(float a) { return make_float2(a, a); }
template <int ACT>
__device__ __forceinline__ float2 wn_act_fwd2(float2 x) {
  if (ACT == BSDE_ACT_SOFTPLUS) {
    const float2 t = __fmul2_rn(x, wn_f2(-1.44269504f));
    float2 e;
    e.x = wn_ex2(-fabsf(t.x)); e.y = wn_ex2(-fabsf(t.y));
    float2 p = wn_f2(0.03137759119272232f);
    p = __ffma2_rn(p, e, wn_f2(-0.134135439991951f));
    p = __ffma2_rn(p, e, wn_f2(0.2878262996673584f));
    p = __ffma2_rn(p, e, wn_f2(-0.49134793877601624f));
    p = __ffma2_rn(p, e, wn_f2(0.9994350075721741f));
    return __ffma2_rn(p, e, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
  }
  return make_float2(wn_act_fwd<ACT>(x.x), wn_act_fwd<ACT>(x.y));
}
template <int EPI, int ACT>
__device__ __forceinline__ float2 wn_combine2(float2 q, float2 x, float scale) {
  if (EPI == BSDE_EPI_DACT) {
    float2 g;
    if (ACT == BSDE_ACT_SOFTPLUS) {
      const float2 t = __fmul2_rn(x, wn_f2(-1.44269504f));
      g = __ffma2_rn(make_float2(wn_ex2(t.x), wn_ex2(t.y)), wn_f2(-1.f), wn_f2(1.f));
    } else {
      g = make_float2(wn_act_grad_out<ACT>(x.x), wn_act_grad_out<ACT>(x.y));
    }
    return __fmul2_rn(__fmul2_rn(wn_f2(scale), q), g);
  }
  if (EPI == BSDE_EPI_EULER || EPI == BSDE_EPI_ACCUM) return __ffma2_rn(wn_f2(scale), q, x);
  return q;
}

template <bool VEC, int EPI, int ACT>
__global__ void __launch_bounds__(WN_THREADS, 1) conv_win_kernel(const __grid_constant__ WinArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / a.G, g = blockIdx.x - s * a.G;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  // carve: ring first (1024-byte aligned swizzle regions), then weights, tables, epilogue scratch
  int soff = 0;
  const uint32_t s_ring = base;  uint8_t* g_ring = gen;         soff += a.ring * a.slab_bytes;
  const uint32_t s_w = base + soff;                             soff += a.w_bytes;
  float* s_tb = reinterpret_cast<float*>(gen + soff);           soff += a.tb_floats * 4;
  float* s_stage = reinterpret_cast<float*>(gen + soff);        soff += WN_EPI_WARPS * 32 * 16 * 4;
  int* s_tab = reinterpret_cast<int*>(gen + soff);              soff += (a.gather ? 2 : a.NP) * a.NPOSw * 4;
  int* s_pix = reinterpret_cast<int*>(gen + soff);              soff += WN_EPI_WARPS * 32 * 4;
  const uint32_t bars = base + soff;                            soff += 8 * (2 * WN_MAXRING + 2 * WN_MAXBUF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + soff);
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto empty_bar = [&](int st) { return bars + 8u * (WN_MAXRING + st); };
  auto accf_bar = [&](int b) { return bars + 8u * (2 * WN_MAXRING + b); };
  auto acce_bar = [&](int b) { return bars + 8u * (2 * WN_MAXRING + WN_MAXBUF + b); };

  // resident weights + tables: one cooperative copy per CTA
  if (!a.wsafe) pdl_wait();
  {
    const float* __restrict__ wsrc = a.wimg + (long long)s * a.wimg_sample_stride;
    const int n16 = (a.w_bytes + a.tb_floats * 4) >> 4;
    for (int i = tid; i < n16; i += WN_THREADS) cp_async16(s_w + i * 16, wsrc + i * 4, true);
    cp_async_commit();
    // zero the ring once: bytes of padded channels are never written by the scalar producer path
    for (int i = tid; i < (a.ring * a.slab_bytes) >> 4; i += WN_THREADS)
      reinterpret_cast<uint4*>(g_ring)[i] = make_uint4(0, 0, 0, 0);
    cp_async_wait_all();
  }
  if (warp == WN_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < a.ring; ++i) { mbar_init(full_bar(i), a.tma ? 1 : WN_PROD); mbar_init(empty_bar(i), 1); }
      for (int b = 0; b < WN_MAXBUF; ++b) { mbar_init(accf_bar(b), 1); mbar_init(acce_bar(b), WN_EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (a.wsafe) pdl_wait();      // everything above overlapped the predecessor's tail
  pdl_launch_dependents();

  if (warp >= WN_EPI_WARPS && warp < WN_MMA_WARP) {
    // =============================== producers ===============================
    const float* __restrict__ in = a.in + (long long)s * a.B * a.hin * a.win * a.cin;
    const int pt = tid - WN_EPI_WARPS * 32;
    const int c = pt & 7;  // 16-byte chunk inside the 128-byte row: 8 lanes cover one row / one global line
    const float rIP = 1.0f / (float)a.IP, rpitch = 1.0f / (float)a.pitch;
    uint32_t cnt = 0;
    if (a.tma) {
      // ---- TMA-fed window: one elected thread issues NR row boxes {32 channels, Wq+1 pixels} per half.  The pixel
      // past the row end, the row past the image end and images outside [0, S*B) are out of bounds -> zero filled:
      // exactly the pad column / pad row of the flattened position grid.  Destinations are 128-byte aligned rows of
      // the SWIZZLE_128B region (tools/tma_probe.cu: the TMA swizzle is a function of the absolute smem address).
      if (pt < 32) {
        // lane r issues the box of window row r (r < NR); lane 0 arms the barrier first
        const int rows_img = a.Hq + 1;
#pragma unroll 1
        for (int tile = g; tile < a.ntiles; tile += a.G) {
          const int tl = a.rev ? a.ntiles - 1 - tile : tile;   // physical tile
          const int n0 = tl * 128 - a.lead;
          const int g0 = (n0 + 2 * a.pitch) / a.pitch - 2 + pt;  // floor(n0 / pitch) + r, n0 >= -(pitch + 1)
          const int img = (g0 + rows_img) / rows_img - 1;         // floor(g0 / rows_img), g0 >= -2
          const int y = g0 - img * rows_img;
#pragma unroll 1
          for (int sl = 0; sl < a.nslab; ++sl, ++cnt) {
            const int slot = cnt % a.ring;
            if (cnt >= (uint32_t)a.ring) mbar_wait(empty_bar(slot), ((cnt / a.ring) & 1) ^ 1);
            const uint32_t sb = s_ring + slot * a.slab_bytes;
            const int h0 = sl * a.hslab;
            const int hend = min(a.hslab, a.nhalf - h0);
            if (!(a.dbg & 1)) {
              if (pt == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(slot)), "r"(hend * a.NR * a.box_bytes) : "memory");
              __syncwarp();
              if (pt < a.NR) {
                for (int hs = 0; hs < hend; ++hs) {
                  const uint32_t dst = sb + (uint32_t)hs * (uint32_t)a.region + (uint32_t)(pt * a.box_bytes);
                  asm volatile(
                      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                      "l"(&tmap), "r"(full_bar(slot)), "r"((h0 + hs) * 32), "r"(0), "r"(y), "r"(s * a.B + img)
                      : "memory");
                }
              }
            } else if (pt == 0) {
              mbar_arrive(full_bar(slot));
            }
          }
        }
      }
    } else
    if (a.gather) {
      // ---- stride-2 layers: per-(tap, half) gathered slabs.  Per tile, one table entry per output position: the offset of
      // input pixel (2y, 2x) and a bit mask of the taps whose source pixel is inside the image.  Every thread then keeps the
      // source pointers and masks of its four rows in registers for the 18 slabs of the tile (one test, one 64-bit add and
      // one select per 16-byte copy; the per-copy coordinate arithmetic made the producers the critical path: ncu source
      // page, 60 % of the launch's instructions).
      const int r0 = pt >> 3;
      const uint32_t swz = (uint32_t)((c ^ (r0 & 7)) << 4);
      constexpr int RPT = 128 / (WN_PROD / 8);   // rows per thread
#pragma unroll 1
      for (int tile = g; tile < a.ntiles; tile += a.G) {
        const int tl = a.rev ? a.ntiles - 1 - tile : tile;   // physical tile
        asm volatile("bar.sync 1, %0;" ::"n"(WN_PROD) : "memory");
        if (pt < 128) {
          const int q = tl * 128 + pt;
          int base = 0, mask = 0;
          if (q < a.Q) {
            int img = (int)((float)q * rIP);
            int r = q - img * a.IP;
            if (r < 0) { --img; r += a.IP; } else if (r >= a.IP) { ++img; r -= a.IP; }
            int y = (int)((float)r * rpitch);
            int x = r - y * a.pitch;
            if (x < 0) { --y; x += a.pitch; } else if (x >= a.pitch) { ++y; x -= a.pitch; }
            base = ((img * a.hin + 2 * y) * a.win + 2 * x) * a.cin;
            for (int e = 0; e < a.nent; ++e) {
              const int iy = 2 * y + a.ent[e].kh * a.kn + a.off, ix = 2 * x + a.ent[e].kw * a.kn + a.off;
              if ((unsigned)iy < (unsigned)a.hin && (unsigned)ix < (unsigned)a.win) mask |= 1 << e;
            }
          }
          s_tab[2 * pt] = base;
          s_tab[2 * pt + 1] = mask;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WN_PROD) : "memory");
        const float* rp[RPT];
        int rm[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
          const int r = r0 + k * (WN_PROD / 8);
          rp[k] = in + s_tab[2 * r] + c * 4;
          rm[k] = s_tab[2 * r + 1];
        }
#pragma unroll 1
        for (int e = 0; e < a.nent; ++e) {
          const int th = a.ent[e].kh * a.kn + a.off, tw = a.ent[e].kw * a.kn + a.off;
          const int toff = (th * a.win + tw) * a.cin;
#pragma unroll 1
          for (int h = 0; h < a.nhalf; ++h, ++cnt) {
            const int slot = cnt % a.ring;
            if (cnt >= (uint32_t)a.ring) mbar_wait(empty_bar(slot), ((cnt / a.ring) & 1) ^ 1);
            const uint32_t sb = s_ring + slot * a.slab_bytes + swz;
            const bool cok = h * 32 + c * 4 < a.cin;
            const int off = toff + h * 32;
            if (!(a.dbg & 1)) {
#pragma unroll
              for (int k = 0; k < RPT; ++k) {
                const bool ok = cok && ((rm[k] >> e) & 1);
                cp_async16(sb + (r0 + k * (WN_PROD / 8)) * 128, ok ? rp[k] + off : in, ok);
              }
            }
            cp_async_arrive_noinc(full_bar(slot));
          }
        }
      }
    } else
#pragma unroll 1
    for (int tile = g; tile < a.ntiles; tile += a.G) {
      const int tl = a.rev ? a.ntiles - 1 - tile : tile;   // physical tile
      const int q0 = tl * 128 - a.lead;
      // decode the window positions once per tile into the shared source table (read by other producer threads)
      asm volatile("bar.sync 1, %0;" ::"n"(WN_PROD) : "memory");  // last tile's readers are done
#pragma unroll 1
      for (int i = pt; i < a.NPOSw; i += WN_PROD) {
        const int q = q0 + i;
        int img = 0, y = 0, x = 0;
        bool pv = q >= 0 && q < a.Q;
        if (pv) {
          // q / IP and r / pitch through fp32 reciprocals with a +-1 fix-up (q < 2^24 is not required: the
          // fix-up loop makes the result exact for any estimate within one of the quotient)
          img = (int)((float)q * rIP);
          int r = q - img * a.IP;
          if (r < 0) { --img; r += a.IP; } else if (r >= a.IP) { ++img; r -= a.IP; }
          y = (int)((float)r * rpitch);
          x = r - y * a.pitch;
          if (x < 0) { --y; x += a.pitch; } else if (x >= a.pitch) { ++y; x -= a.pitch; }
          pv = y < a.Hq && x < a.Wq;
        }
        for (int plane = 0; plane < a.NP; ++plane) {
          int src = -1;
          if (pv) {
            int iy = y, ix = x;
            if (a.mode == 2) { iy = 2 * y + (plane >> 1); ix = 2 * x + (plane & 1); }
            if (iy < a.hin && ix < a.win) src = ((img * a.hin + iy) * a.win + ix) * a.cin;
          }
          s_tab[plane * a.NPOSw + i] = src;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(WN_PROD) : "memory");
#pragma unroll 1
      for (int sl = 0; sl < a.nslab; ++sl, ++cnt) {
        const int slot = cnt % a.ring;
        if (cnt >= (uint32_t)a.ring) mbar_wait(empty_bar(slot), ((cnt / a.ring) & 1) ^ 1);
        const uint32_t sb = s_ring + slot * a.slab_bytes;
        const int h0 = sl * a.hslab;
        const int hend = min(a.hslab, a.nhalf - h0);
        if (a.dbg & 1) {
        } else if (VEC) {
          // thread = (chunk c, rows i0 + 16 k): i & 7 is constant per thread, so the swizzled chunk slot is too
          const int i0 = pt >> 3;
          const uint32_t swz = (uint32_t)((c ^ (i0 & 7)) << 4);
          const float* __restrict__ inc = in + h0 * 32 + c * 4;
          const int cfirst = h0 * 32 + c * 4;
#pragma unroll 1
          for (int plane = 0; plane < a.NP; ++plane) {
            const uint32_t reg = sb + (uint32_t)(plane * a.hslab) * (uint32_t)a.region + swz;
            const int* tabp = s_tab + plane * a.NPOSw;
#pragma unroll 4
            for (int i = i0; i < a.NPOSw; i += WN_PROD / 8) {
              const int src = tabp[i];
              const float* gp = inc + (src >= 0 ? src : 0);
              uint32_t d = reg + i * 128;
              for (int hs = 0; hs < hend; ++hs) {
                const bool ok = src >= 0 && cfirst + hs * 32 < a.cin;
                cp_async16(d, ok ? gp + hs * 32 : in, ok);
                d += a.region;
              }
            }
          }
        } else {
#pragma unroll 1
          for (int plane = 0; plane < a.NP; ++plane) {
            const uint32_t reg = sb + (uint32_t)(plane * a.hslab) * (uint32_t)a.region;
            const int cbase = h0 * 32;
#pragma unroll 1
            for (int i = pt; i < a.NPOSw; i += WN_PROD) {
              const int src = s_tab[plane * a.NPOSw + i];
              const bool ok = src >= 0;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch)
                if (cbase + ch < a.cin)
                  cp_async4(reg + i * 128 + (((ch >> 2) ^ (i & 7)) << 4) + (ch & 3) * 4, ok ? in + src + cbase + ch : in, ok);
            }
          }
        }
        cp_async_arrive_noinc(full_bar(slot));
      }
    }
  } else if (warp == WN_MMA_WARP) {
    // =============================== MMA issuer ===============================
    // The whole warp runs the loop; every operand is built from kernel parameters and loop counters only, so
    // the descriptors live in uniform registers and one elected lane issues each tcgen05.mma.
    {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_hi = desc_hi(16u, 1024u, 2);  // K-major SWIZZLE_128B, 8-row groups 1024 B apart
      const uint64_t b_hi = desc_hi((uint32_t)a.npad * 16u, 128u, 0);
      const uint32_t w16 = s_w >> 4, bstep = (uint32_t)(a.npad * 32) >> 4, region16 = (uint32_t)a.region >> 4;
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      // per-tap constants of the unrolled 3x3 issue path, loaded once (uniform registers)
      uint32_t e_ao[9], e_bo[9], e_dc[9], e_fl[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        e_ao[e] = a.aoff[e]; e_bo[e] = (uint32_t)(e * a.nks) * bstep;
        e_dc[e] = a.dcol[e] & 0xFFFFu; e_fl[e] = (a.dcol[e] >> 31) ^ 1u;
      }
      uint32_t cnt = 0;
      int it = 0;
#pragma unroll 1
      for (int tile = g; tile < a.ntiles; tile += a.G, ++it) {
        const int tl = a.rev ? a.ntiles - 1 - tile : tile;   // physical tile
        const int buf = it % a.nbuf;
        if (it >= a.nbuf) {
          mbar_wait(acce_bar(buf), ((it / a.nbuf) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t dbase = tm + (uint32_t)(buf * a.NA * a.npad);
        uint32_t o0 = 0;  // TMA-fed window: positions of the first grid row before position q0 - lead, in 16-byte units
        if (a.tma) {
          const int n0 = tl * 128 - a.lead;
          const int g0 = (n0 + 2 * a.pitch) / a.pitch - 2;
          o0 = (uint32_t)(n0 - g0 * a.pitch) * 8u;
        }
        if (a.gather) {
          // slabs arrive in (tap, half) order; each holds four k-steps of one tap: unshifted K-major SW128 tile
#pragma unroll 1
          for (int e = 0; e < a.nent; ++e) {
#pragma unroll 1
            for (int h = 0; h < a.nhalf; ++h, ++cnt) {
              const int slot = cnt % a.ring;
              mbar_wait(full_bar(slot), (cnt / a.ring) & 1);
              fence_proxy_async();
              tc_fence_after();
              const uint32_t sb16 = (s_ring + slot * a.slab_bytes) >> 4;
              const int kend = min(4, a.nks - h * 4);
              if (!(a.dbg & 2)) {
#pragma unroll 1
                for (int kk = 0; kk < kend; ++kk) {
                  const int ks = h * 4 + kk;
                  const uint64_t ad = a_hi | (uint64_t)((sb16 + (uint32_t)kk * 2u) & 0x3FFFu);
                  const uint64_t bd = b_hi | (uint64_t)((w16 + (uint32_t)(e * a.nks + ks) * bstep) & 0x3FFFu);
                  umma_tf32_elect(dbase, ad, bd, idesc, (e > 0 || ks > 0) ? 1u : 0u);
                }
              }
              umma_commit_elect(empty_bar(slot));
            }
          }
          umma_commit_elect(accf_bar(buf));
          continue;
        }
#pragma unroll 1
        for (int sl = 0; sl < a.nslab; ++sl, ++cnt) {
          const int slot = cnt % a.ring;
          mbar_wait(full_bar(slot), (cnt / a.ring) & 1);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t sb16 = (s_ring + slot * a.slab_bytes) >> 4;
          const int ks0 = sl * a.hslab * 4;
          const int kend = min(a.hslab * 4, a.nks - ks0);
          if (!(a.dbg & 2)) {
#pragma unroll 1
            for (int kk = 0; kk < kend; ++kk) {
              // k-step kk of the slab: half kk >> 2 (region stride), 32-byte step kk & 3 inside the 128-byte row
              const uint32_t ka = sb16 + o0 + (uint32_t)(kk >> 2) * region16 + (uint32_t)(kk & 3) * 2u;
              const uint32_t kb = w16 + (uint32_t)(ks0 + kk) * bstep;
              const uint32_t later = (ks0 + kk) > 0 ? 1u : 0u;
#define WN_ISSUE(e)                                                                                         \
  {                                                                                                         \
    const uint32_t dc = a.dcol[e];                                                                          \
    const uint64_t ad = a_hi | (uint64_t)((ka + a.aoff[e]) & 0x3FFFu);                                      \
    const uint64_t bd = b_hi | (uint64_t)((kb + (uint32_t)((e) * a.nks) * bstep) & 0x3FFFu);                \
    umma_tf32_elect(dbase + (dc & 0xFFFFu), ad, bd, idesc, later | ((dc >> 31) ^ 1u));                      \
  }
              if (a.nent == 9) {
#pragma unroll
                for (int e = 0; e < 9; ++e)
                  umma_tf32_elect(dbase + e_dc[e], a_hi | (uint64_t)((ka + e_ao[e]) & 0x3FFFu), b_hi | (uint64_t)((kb + e_bo[e]) & 0x3FFFu),
                                  idesc, later | e_fl[e]);
              } else {
#pragma unroll 1
                for (int e = 0; e < a.nent; ++e) WN_ISSUE(e)
              }
#undef WN_ISSUE
            }
          }
          umma_commit_elect(empty_bar(slot));
        }
        umma_commit_elect(accf_bar(buf));
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    // warp w: TMEM lane quarter w & 3 (rows), and every second (accumulator, 16-column chunk) work item
    const int quarter = warp & 3, half = warp >> 2;
    float* stage = s_stage + warp * 32 * 16;
    int* w_pix = s_pix + warp * 32;
    float* __restrict__ out = a.out + (long long)s * a.B * a.hout * a.wout * a.opitch;
    const float* __restrict__ aux = a.aux ? a.aux + (long long)s * a.B * a.hout * a.wout * a.opitch : nullptr;
    const bool vec_out = ((a.cout | a.opitch) & 3) == 0;
    const int nchunk = (a.cout + 15) >> 4;
    const int c4 = lane & 3, r8 = lane >> 2;
    const float e_rIP = 1.0f / (float)a.IP, e_rpitch = 1.0f / (float)a.pitch;
    // fold the time term into the table once per CTA: row 1 + ac * 16 + cls becomes bias + t * (time rows of that border class),
    // so that an item costs one table read and one add per element instead of two reads, an add and a multiply-add
    if (a.has_time) {
      for (int r = warp; r < 16 * a.NA; r += WN_EPI_WARPS)
        for (int n = lane; n < a.npad; n += 32) s_tb[(1 + r) * a.npad + n] = fmaf(a.t, s_tb[(1 + r) * a.npad + n], s_tb[n]);
      asm volatile("bar.sync 2, %0;" ::"n"(WN_EPI_WARPS * 32) : "memory");
    }
    int it = 0;
#pragma unroll 1
    for (int tile = g; tile < a.ntiles; tile += a.G, ++it) {
      const int tl = a.rev ? a.ntiles - 1 - tile : tile;   // physical tile
      const int buf = it % a.nbuf;
      const int q = tl * 128 + quarter * 32 + lane;
      bool valid = q < a.Q;
      int img = 0, y = 0, x = 0;
      if (valid) {
        img = (int)((float)q * e_rIP);
        int r = q - img * a.IP;
        if (r < 0) { --img; r += a.IP; } else if (r >= a.IP) { ++img; r -= a.IP; }
        y = (int)((float)r * e_rpitch);
        x = r - y * a.pitch;
        if (x < 0) { --y; x += a.pitch; } else if (x >= a.pitch) { ++y; x -= a.pitch; }
        valid = y < a.Hq && x < a.Wq;
      }
      // border class of this position (selects the time-term row)
      const int cls = (((y == 0) | ((y == a.Hq - 1) << 1)) << 2) | ((x == 0) | ((x == a.Wq - 1) << 1));
      if (EPI >= BSDE_EPI_EULER && a.pf) {
        // combine operand of this CTA's NEXT tile -> L2 now (one tile of MMAs and epilogue ahead): the loads at the head of
        // every item otherwise wait out a DRAM round trip with two warps per scheduler to cover it (data-gradient launches:
        // 84 -> 188 us with the combine operand against without).  Warp `half` takes every second 128-byte line of a pixel.
        const int tn = tile + a.G;
        const int qn = (a.rev ? a.ntiles - 1 - tn : tn) * 128 + quarter * 32 + lane;
        if (tn < a.ntiles && qn < a.Q) {
          int im = (int)((float)qn * e_rIP);
          int r = qn - im * a.IP;
          if (r < 0) { --im; r += a.IP; } else if (r >= a.IP) { ++im; r -= a.IP; }
          int yn = (int)((float)r * e_rpitch);
          int xn = r - yn * a.pitch;
          if (xn < 0) { --yn; xn += a.pitch; } else if (xn >= a.pitch) { ++yn; xn -= a.pitch; }
          if (yn < a.Hq && xn < a.Wq) {
            const int nline = (a.cout * 4 + 127) >> 7;
            for (int acn = 0; acn < a.NA; ++acn) {
              int oy = yn, ox = xn;
              if (a.mode == 3) { oy = 2 * yn + (acn >> 1); ox = 2 * xn + (acn & 1); }
              const float* pa = aux + ((im * a.hout + oy) * a.wout + ox) * a.opitch;
              for (int l = half; l < nline; l += 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + l * 32));
            }
          }
        }
      }
      // element offsets (< 2^31: tconv_win_supported) of this tile's output pixels, accumulator 0: row `lane` publishes its own,
      // every lane then keeps those of its four store rows r8 + 8 j for all items of the tile; accumulator ac of a
      // transposed layer writes the pixel (ac >> 1) rows down and (ac & 1) columns right of it
      int pp[4];
      {
        int oy = y, ox = x;
        if (a.mode == 3) { oy = 2 * y; ox = 2 * x; }
        __syncwarp();   // the previous tile's readers are done
        w_pix[lane] = valid ? ((img * a.hout + oy) * a.wout + ox) * a.opitch : -1;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) pp[j] = w_pix[r8 + 8 * j];
      }
      int ac = 0, ch_i = half;  // item = ac * nchunk + ch_i, stepping by 2 without a division
      while (ch_i >= nchunk) { ch_i -= nchunk; ++ac; }
      // combine epilogues (Euler / data-gradient / accumulate): the values already stored at the output positions are loaded ONE
      // ITEM AHEAD (the first item's before the accumulator is even complete), so that their L2 latency runs under a whole item
      // of TMEM load, row arithmetic, staging exchange and stores -- two epilogue warps per scheduler cannot hide it otherwise
      float4 xn[4];
#define WN_LOAD_AUX(ac_, ch_)                                                                                              \
  {                                                                                                                        \
    const int n0_ = (ch_) * 16;                                                                                            \
    if (vec_out && a.cout - n0_ >= 16) {                                                                                   \
      const int o_ = (a.mode == 3 ? (((ac_) >> 1) * a.wout + ((ac_) & 1)) * a.opitch : 0) + n0_ + c4 * 4;                  \
      xn[0] = pp[0] >= 0 ? *reinterpret_cast<const float4*>(aux + (pp[0] + o_)) : make_float4(0.f, 0.f, 0.f, 0.f);         \
      xn[1] = pp[1] >= 0 ? *reinterpret_cast<const float4*>(aux + (pp[1] + o_)) : make_float4(0.f, 0.f, 0.f, 0.f);         \
      xn[2] = pp[2] >= 0 ? *reinterpret_cast<const float4*>(aux + (pp[2] + o_)) : make_float4(0.f, 0.f, 0.f, 0.f);         \
      xn[3] = pp[3] >= 0 ? *reinterpret_cast<const float4*>(aux + (pp[3] + o_)) : make_float4(0.f, 0.f, 0.f, 0.f);         \
    }                                                                                                                      \
  }
      const int na_run = (a.dbg & 8) ? 0 : a.NA;
      if (EPI >= BSDE_EPI_EULER && ac < na_run) WN_LOAD_AUX(ac, ch_i)
      mbar_wait(accf_bar(buf), (it / a.nbuf) & 1);
      tc_fence_after();
#pragma unroll 1
      for (; ac < na_run;) {
        const int n0 = ch_i * 16;
        const int dac = a.mode == 3 ? ((ac >> 1) * a.wout + (ac & 1)) * a.opitch : 0;
        const int ncol = min(16, a.cout - n0);
        const bool fast = vec_out && ncol == 16;
        int ac2 = ac, ch2 = ch_i + 2;
        while (ch2 >= nchunk) { ch2 -= nchunk; ++ac2; }
        float4 xv[4];
        if (EPI >= BSDE_EPI_EULER) {
#pragma unroll
          for (int j = 0; j < 4; ++j) xv[j] = xn[j];
          if (ac2 < na_run) WN_LOAD_AUX(ac2, ch2)
        }
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((buf * a.NA + ac) * a.npad + n0), v);
        // row form: bias, time term, activation on this lane's 16 accumulator columns
        float2 f[8];
        {
          // one table row per (accumulator, border class): bias + t * (time rows of the taps inside the image), folded once per CTA
          const float4* bt = reinterpret_cast<const float4*>(s_tb + (a.has_time ? (1 + ac * 16 + cls) * a.npad : 0) + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = bt[j];
            f[2 * j] = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), make_float2(b4.x, b4.y));
            f[2 * j + 1] = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b4.z, b4.w));
          }
          if (EPI == BSDE_EPI_BIAS_ACT) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = wn_act_fwd2<ACT>(f[j]);
          }
        }
        __syncwarp();  // readers of the previous chunk are done with the staging rows
        {
          // row `lane`: 16-byte chunk c stored at position c ^ ((lane >> 1) & 3)  (conflict-free both ways)
          float4* dst = reinterpret_cast<float4*>(stage + lane * 16);
          const int fz = (lane >> 1) & 3;
          dst[0 ^ fz] = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
          dst[1 ^ fz] = make_float4(f[2].x, f[2].y, f[3].x, f[3].y);
          dst[2 ^ fz] = make_float4(f[4].x, f[4].y, f[5].x, f[5].y);
          dst[3 ^ fz] = make_float4(f[6].x, f[6].y, f[7].x, f[7].y);
        }
        __syncwarp();
        if (fast) {
          // lane -> fixed column group c4, rows r8 + 8 j: 64 contiguous bytes per row
          const int n = dac + n0 + c4 * 4;
          float4 qv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = r8 + 8 * j;
            qv[j] = *reinterpret_cast<const float4*>(stage + rr * 16 + ((c4 ^ ((rr >> 1) & 3)) << 2));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (pp[j] < 0 || (a.dbg & 4)) continue;
            float4 qq = qv[j];
            if (EPI >= BSDE_EPI_EULER) {
              const float2 lo = wn_combine2<EPI, ACT>(make_float2(qq.x, qq.y), make_float2(xv[j].x, xv[j].y), a.scale);
              const float2 hi = wn_combine2<EPI, ACT>(make_float2(qq.z, qq.w), make_float2(xv[j].z, xv[j].w), a.scale);
              qq = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
            *reinterpret_cast<float4*>(out + (pp[j] + n)) = qq;
          }
        } else {
          // partial chunk / cout not a multiple of 4: rows r8 + 8 j, columns c4 + 4 i
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int rr = r8 + 8 * j;
            const int pq = w_pix[rr];   // (pp[] stays in registers: no dynamic index)
            if (pq < 0) continue;
#pragma unroll 1
            for (int cc = c4; cc < ncol; cc += 4) {
              const float qq = stage[rr * 16 + ((((cc >> 2) ^ ((rr >> 1) & 3)) << 2) | (cc & 3))];
              const int o = pq + dac + n0 + cc;
              const float xx = EPI >= BSDE_EPI_EULER ? aux[o] : 0.f;
              out[o] = wn_combine<EPI, ACT>(qq, xx, a.scale);
            }
          }
        }
        ac = ac2; ch_i = ch2;
      }
#undef WN_LOAD_AUX
      // this warp has drained its share of accumulator buffer `buf`
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WN_MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ---- host side ---------------------------------------------------------------------------------------------
struct WinPlan {
  bool ok;
  int mode, NA, NP, nent, Hq, Wq, pitch, IP, lead, trail, NPOSw, nks, npad, ring, slab_bytes, w_bytes, tb_floats;
  int tab_units, nbuf, tmem_cols, vec, nhalf, hslab, nslab, region, tb_rows, gather, tma, NR, box_bytes;
  size_t smem, img_floats;
  WinEntry ent[WN_MAXENT];
  int ent_tap[WN_MAXENT];
};

static int wn_pow2_cols(int c) { int p = 32; while (p < c) p <<= 1; return p; }

static WinPlan win_plan(const bsde_conv_layer* L) {
  WinPlan p;
  memset(&p, 0, sizeof(p));
  if (!L || (L->ksize != 1 && L->ksize != 3)) return p;
  if (L->cout < 1 || L->cout > 128 || L->cin < 1) return p;
  p.vec = (L->cin % 4) == 0;
  if (!p.vec && L->cin > 8) return p;  // scalar producer path: one k-step only
  if (L->sd == 1 && L->sn == 1) p.mode = 1;
  else if (L->sd == 1 && L->sn == 2) p.mode = 2;
  else if (L->sd == 2 && L->sn == 1) p.mode = 3;
  else return p;
  if (p.mode == 1) {
    if (L->hin != L->hout || L->win != L->wout) return p;
    p.Hq = L->hout; p.Wq = L->wout; p.NA = 1; p.NP = 1;
  } else if (p.mode == 2) {
    if ((L->hin + 1) / 2 != L->hout || (L->win + 1) / 2 != L->wout) return p;
    p.Hq = L->hout; p.Wq = L->wout; p.NA = 1; p.NP = 4;
  } else {
    if (L->hout != 2 * L->hin || L->wout != 2 * L->win) return p;
    p.Hq = L->hin; p.Wq = L->win; p.NA = 4; p.NP = 1;
  }
  p.pitch = p.Wq + 1;
  p.IP = (p.Hq + 1) * p.pitch;
  int mins = 0, maxs = 0;
  for (int kh = 0; kh < L->ksize; ++kh)
    for (int kw = 0; kw < L->ksize; ++kw) {
      const int th = kh * L->kn + L->off, tw = kw * L->kn + L->off;
      if (p.mode == 3) {
        // output parity class (poh, pow): taps with (po + t) even, displacement (po + t) / 2
        for (int cls = 0; cls < 4; ++cls) {
          const int poh = cls >> 1, pow_ = cls & 1;
          if (((poh + th) & 1) || ((pow_ + tw) & 1)) continue;
          const int sy = (poh + th) / 2, sx = (pow_ + tw) / 2;
          if (abs(sy) > 1 || abs(sx) > 1 || p.nent >= WN_MAXENT) return p;
          WinEntry& e = p.ent[p.nent];
          e.acc = cls; e.plane = 0; e.shift = sy * p.pitch + sx; e.kh = kh; e.kw = kw;
          p.ent_tap[p.nent++] = kh * L->ksize + kw;
        }
      } else {
        int sy = th, sx = tw, plane = 0;
        if (p.mode == 2) {
          const int py = th & 1, px = tw & 1;
          sy = (th - py) / 2; sx = (tw - px) / 2; plane = py * 2 + px;
        }
        if (abs(sy) > 1 || abs(sx) > 1 || p.nent >= WN_MAXENT) return p;
        WinEntry& e = p.ent[p.nent];
        e.acc = 0; e.plane = plane; e.shift = sy * p.pitch + sx; e.kh = kh; e.kw = kw;
        p.ent_tap[p.nent++] = kh * L->ksize + kw;
      }
    }
  for (int e = 0; e < p.nent; ++e) { mins = std::min(mins, p.ent[e].shift); maxs = std::max(maxs, p.ent[e].shift); }
  if (p.nent < 1) return p;
  // every accumulator must receive at least one tap (else its TMEM columns would be read uninitialised)
  for (int ac = 0; ac < p.NA; ++ac) {
    bool any = false;
    for (int e = 0; e < p.nent; ++e) any |= p.ent[e].acc == ac;
    if (!any) return p;
  }
  p.lead = -mins; p.trail = maxs;
  p.NPOSw = (p.lead + 128 + p.trail + 7) & ~7;  // plane stride = NPOSw*16 B (multiple of 128 B)
  p.nks = (L->cin + 7) / 8;
  p.npad = (L->cout + 15) / 16 * 16;
  p.nhalf = (L->cin + 31) / 32;
  p.region = p.NPOSw * 128;  // one (plane, 32-channel half) SWIZZLE_128B region; NPOSw % 8 == 0 keeps it 1024-byte aligned
  // TMA-fed window (input channels a multiple of 32, single input plane): the window buffer holds NR whole grid rows
  static const int tma_on = getenv("BSDE_TMA") ? atoi(getenv("BSDE_TMA")) : 1;
  if (tma_on && p.vec && (L->cin % 32) == 0 && p.NP == 1 && p.Wq + 1 <= 256) {
    const int nr = (p.lead + 128 + p.trail + p.pitch - 2) / p.pitch + 1;   // grid rows any 128-position window (+ halo) can touch
    // one lane of the producer warp issues the box of one window row: grids narrower than 4 pixels (more than 32 rows per
    // window) keep the cp.async-fed window (the barrier would otherwise wait for boxes that no lane issues)
    if (nr <= 32) {
      p.tma = 1;
      p.NR = nr;
      p.box_bytes = p.pitch * 128;
      p.region = (p.NR * p.box_bytes + 128 + 1023) & ~1023;          // + one slack row for the last shifted read
    }
  }
  p.w_bytes = p.nent * p.nks * p.npad * 32;
  p.tb_rows = L->has_time ? 1 + 16 * p.NA : 1;
  p.tb_floats = p.tb_rows * p.npad;
  p.tab_units = p.NP * p.NPOSw;
  p.nbuf = std::max(1, std::min(WN_MAXBUF, 512 / (p.NA * p.npad)));
  if (p.nbuf == 3) p.nbuf = 2;
  if (p.NA * p.npad > 512) return p;
  p.tmem_cols = wn_pow2_cols(p.nbuf * p.NA * p.npad);
  const size_t fixed = 1024 + (size_t)p.w_bytes + (size_t)p.tb_floats * 4 + WN_EPI_WARPS * 32 * 16 * 4 + (size_t)p.tab_units * 4 +
                       WN_EPI_WARPS * 32 * 4 + 8 * (2 * WN_MAXRING + 2 * WN_MAXBUF) + 16;
  const size_t cap = 227 * 1024;
  // slab = all 32-channel halves of a tile's window (ring >= 2), else one half per slab (ring >= 2)
  p.hslab = 0;
  const int cands[2] = {p.nhalf, 1};
  for (int ci = 0; ci < 2 && !p.hslab; ++ci) {
    const int k = cands[ci];
    const size_t slab = (size_t)p.NP * k * p.region;
    if (fixed + (p.tma && k > 1 ? 3 : 2) * slab <= cap) { p.hslab = k; p.slab_bytes = (int)slab; }
  }
  if (!p.hslab) {
    // Stride-2 layer whose four parity-plane windows do not fit beside the resident weights: keep the persistent
    // structure (resident weights, TMEM double buffering, uniform MMA issue) but fill the ring with per-(tap, half)
    // gathered tiles of 128 rows x 128 B; positions are the plain output pixels (no pad column / row needed).
    if (p.mode != 2 || !p.vec) return p;
    p.gather = 1;
    p.pitch = p.Wq; p.IP = p.Hq * p.Wq; p.lead = 0; p.trail = 0; p.NPOSw = 128; p.NP = 1;
    p.region = 128 * 128; p.hslab = 1; p.slab_bytes = p.region;
    const size_t old_tab = (size_t)p.tab_units * 4;
    p.tab_units = 2 * 128;
    const size_t fixed_g = fixed - old_tab + (size_t)p.tab_units * 4;
    if (fixed_g + 3 * (size_t)p.slab_bytes > cap) return p;
    p.nslab = p.nent * p.nhalf;
    p.ring = (int)std::min<size_t>(WN_MAXRING, (cap - fixed_g) / p.slab_bytes);
    p.smem = fixed_g + (size_t)p.ring * p.slab_bytes;
    p.img_floats = ((size_t)p.w_bytes / 4 + p.tb_floats + 3) & ~(size_t)3;
    p.ok = true;
    return p;
  }
  p.nslab = (p.nhalf + p.hslab - 1) / p.hslab;
  p.ring = (int)std::min<size_t>(WN_MAXRING, (cap - fixed) / p.slab_bytes);
  p.ring = std::min(p.ring, std::max(2, 3 * p.nslab));
  p.smem = fixed + (size_t)p.ring * p.slab_bytes;
  p.img_floats = ((size_t)p.w_bytes / 4 + p.tb_floats + 3) & ~(size_t)3;
  p.ok = true;
  return p;
}

static bool win_enabled() {
  static const int on = [] { const char* e = getenv("BSDE_WIN"); return e ? atoi(e) : 1; }();
  return on != 0;
}

// A layer whose resident weight image does not fit beside the window ring (e.g. 80 -> 64 and 64 -> 80 channels: 184 KB) is
// served as TWO launches over output-channel slices [0, c0) and [c0, cout): each slice is an ordinary layer whose kernel
// and bias offsets are shifted by c0 output channels and whose rows are written with the full layer's pixel pitch.
struct WinSplit {
  int n;                     // 0: unsupported, 1: whole layer, 2: two channel slices
  bsde_conv_layer sub[2];
  int n0[2];                 // first output channel of each slice
  size_t img_off[2];         // float offset of each slice's image inside a sample's image
  size_t img_floats;         // floats of one sample's image (all slices)
};

static WinSplit win_split(const bsde_conv_layer* L) {
  WinSplit sp;
  memset(&sp, 0, sizeof(sp));
  if (!L) return sp;
  WinPlan p = win_plan(L);
  if (p.ok) {
    sp.n = 1; sp.sub[0] = *L; sp.img_floats = p.img_floats;
    return sp;
  }
  static const int on = getenv("BSDE_WIN_NSPLIT") ? atoi(getenv("BSDE_WIN_NSPLIT")) : 1;
  if (!on || L->cout < 32 || (L->cout & 3)) return sp;
  const int c0 = ((L->cout / 2 + 15) / 16) * 16;   // 64 -> 32 + 32, 80 -> 48 + 32
  if (c0 >= L->cout) return sp;
  size_t off = 0;
  for (int i = 0; i < 2; ++i) {
    bsde_conv_layer& q = sp.sub[i];
    q = *L;
    sp.n0[i] = i == 0 ? 0 : c0;
    q.cout = i == 0 ? c0 : L->cout - c0;
    q.w_off = L->w_off + (long long)sp.n0[i] * L->w_n_stride;
    if (L->b_off >= 0) q.b_off = L->b_off + sp.n0[i];
    WinPlan ps = win_plan(&q);
    if (!ps.ok) return sp;
    sp.img_off[i] = off;
    off += ps.img_floats;
  }
  sp.n = 2; sp.img_floats = off;
  return sp;
}

bool tconv_win_supported(const bsde_conv_layer* L, int S, int B) {
  if (!win_enabled() || !L || S < 1 || S > 148 || B < 1) return false;
  WinSplit sp = win_split(L);
  if (!sp.n) return false;
  WinPlan p = win_plan(&sp.sub[0]);
  // 32-bit position / offset arithmetic inside the kernel
  if ((long long)B * p.IP + 1024 > (1LL << 24)) return false;  // fp32-reciprocal position decode is exact below 2^24
  if ((long long)B * L->hin * L->win * L->cin > 0x7FFFFFFFLL) return false;
  if ((long long)B * L->hout * L->wout * L->cout > 0x7FFFFFFFLL) return false;
  return true;
}

size_t tconv_win_ws_bytes(const bsde_conv_layer* L, int S) {
  WinSplit sp = win_split(L);
  return sp.n ? sp.img_floats * sizeof(float) * S : 0;
}

// floats of one sample's resident image (weight tiles + tables; all channel slices)
size_t tconv_win_img_floats(const bsde_conv_layer* L) {
  WinSplit sp = win_split(L);
  return sp.n ? sp.img_floats : 0;
}

// Repack the weights of `nit` scan iterations in one launch: iteration z reads w + z*w_iter_stride and writes
// img + z*S*img_floats ([iteration][sample][img_floats]).  `epi` only selects whether the bias row is kept.
static int win_repack_one(const bsde_conv_layer* L, int S, int nit, const float* w, long long w_sample_stride, long long w_iter_stride,
                          int epi, float* img, long long img_sample_stride, cudaStream_t st) {
  WinPlan p = win_plan(L);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the window-staged tensor-core path");
  WinRepackArgs r;
  memset(&r, 0, sizeof(r));
  r.w = w; r.img = img; r.w_sample_stride = w_sample_stride; r.img_sample_stride = img_sample_stride;
  r.w_iter_stride = w_iter_stride; r.img_iter_stride = img_sample_stride * S;
  r.nent = p.nent; r.nks = p.nks; r.npad = p.npad; r.cin = L->cin; r.cout = L->cout; r.ksize = L->ksize; r.NA = p.NA;
  r.real_base = (L->has_time && L->time_first) ? 1 : 0; r.has_time = L->has_time;
  r.time_row = L->time_first ? 0 : L->cin; r.use_bias = epi <= BSDE_EPI_EULER; r.w_floats = p.w_bytes / 4;
  r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
  r.tb_rows = p.tb_rows; r.mode = p.mode; r.Hq = p.Hq; r.Wq = p.Wq; r.hin = L->hin; r.win = L->win;
  r.sn = L->sn; r.kn = L->kn; r.off = L->off; r.sd = L->sd;
  for (int e = 0; e < p.nent; ++e) {
    r.ent_tap[e] = p.ent_tap[e]; r.ent_acc[e] = p.ent[e].acc; r.ent_kh[e] = p.ent[e].kh; r.ent_kw[e] = p.ent[e].kw;
  }
  win_repack_kernel<<<dim3(p.nent * p.nks + 1, S, nit), 256, 0, st>>>(r);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int tconv_win_repack(const bsde_conv_layer* L, int S, int nit, const float* w, long long w_sample_stride, long long w_iter_stride,
                     int epi, float* img, cudaStream_t st) {
  WinSplit sp = win_split(L);
  BSDE_CHECK_ARG(sp.n > 0, "layer not supported by the window-staged tensor-core path");
  for (int i = 0; i < sp.n; ++i)
    if (int e = win_repack_one(&sp.sub[i], S, nit, w, w_sample_stride, w_iter_stride, epi, img + sp.img_off[i], (long long)sp.img_floats, st))
      return e;
  return 0;
}

static thread_local int g_win_rev = 0;   // tile direction of the launches of the current tconv_fwd_win call
static int win_launch_one(const bsde_conv_layer* L, int S, int B, const float* in, float t, int act, int epi, float scale,
                          const float* aux, float* out, int opitch, const float* wimg, long long wimg_sample_stride, int wsafe,
                          cudaStream_t st);

int tconv_fwd_win(const bsde_conv_layer* L, int S, int B, const float* in, const float* w, long long w_sample_stride,
                  float t, int act, int epi, float scale, const float* aux, float* out, void* ws, size_t ws_bytes,
                  cudaStream_t st, const float* prepacked, int prepacked_safe) {
  WinSplit sp = win_split(L);
  BSDE_CHECK_ARG(sp.n > 0, "layer not supported by the window-staged tensor-core path");
  BSDE_CHECK_ARG(!(epi >= BSDE_EPI_EULER) || aux != nullptr, "epilogue %d needs aux", epi);
  if (!prepacked) {
    const size_t need = sp.img_floats * sizeof(float) * S;
    if (!ws || ws_bytes < need) {
      set_error("tconv_fwd(win): workspace %zu < %zu bytes", ws_bytes, need);
      return BSDE_ERR_WORKSPACE;
    }
    if (int e = tconv_win_repack(L, S, 1, w, w_sample_stride, 0, epi, (float*)ws, st)) return e;
  }
  const float* wimg = prepacked ? prepacked : (const float*)ws;
  // consecutive launches of a chain are producer -> consumer over tensors larger than L2: they alternate the tile direction
  static const int rev_on = getenv("BSDE_WIN_REV") ? atoi(getenv("BSDE_WIN_REV")) : 1;
  static thread_local int rev_toggle = 0;
  g_win_rev = rev_on ? rev_toggle : 0;
  rev_toggle ^= 1;
  for (int i = 0; i < sp.n; ++i) {
    // slice i > 0 follows slice i-1 in the stream: its image was written before that kernel started
    const int wsafe = prepacked ? (i > 0 ? 1 : prepacked_safe) : 0;
    if (int e = win_launch_one(&sp.sub[i], S, B, in, t, act, epi, scale, aux ? aux + sp.n0[i] : nullptr, out + sp.n0[i], L->cout,
                               wimg + sp.img_off[i], (long long)sp.img_floats, wsafe, st))
      return e;
  }
  return 0;
}

static int win_launch_one(const bsde_conv_layer* L, int S, int B, const float* in, float t, int act, int epi, float scale,
                          const float* aux, float* out, int opitch, const float* wimg, long long wimg_sample_stride, int wsafe,
                          cudaStream_t st) {
  WinPlan p = win_plan(L);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the window-staged tensor-core path");
  WinArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.aux = aux; a.out = out; a.wimg = wimg; a.wimg_sample_stride = wimg_sample_stride; a.opitch = opitch;
  a.S = S; a.B = B;
  a.G = std::max(1, sm_budget() / S);
  a.Q = B * p.IP;
  a.ntiles = (a.Q + 127) / 128;
  a.G = std::min(a.G, a.ntiles);
  a.cin = L->cin; a.cout = L->cout; a.npad = p.npad; a.nks = p.nks; a.nhalf = p.nhalf;
  a.hin = L->hin; a.win = L->win; a.hout = L->hout; a.wout = L->wout;
  a.sn = L->sn; a.kn = L->kn; a.off = L->off; a.sd = L->sd; a.has_time = L->has_time; a.mode = p.mode;
  a.Hq = p.Hq; a.Wq = p.Wq; a.pitch = p.pitch; a.IP = p.IP; a.lead = p.lead; a.NPOSw = p.NPOSw; a.NP = p.NP; a.NA = p.NA;
  a.nent = p.nent; a.ring = p.ring; a.slab_bytes = p.slab_bytes; a.w_bytes = p.w_bytes; a.tb_floats = p.tb_floats;
  a.nbuf = p.nbuf; a.tmem_cols = p.tmem_cols;
  a.hslab = p.hslab; a.nslab = p.nslab; a.region = p.region; a.gather = p.gather;
  a.tma = p.gather ? 0 : p.tma; a.NR = p.NR; a.box_bytes = p.box_bytes;
  a.t = t; a.scale = scale;
  a.rev = g_win_rev;
  { static const int pf = getenv("BSDE_WIN_PF") ? atoi(getenv("BSDE_WIN_PF")) : 1; a.pf = pf; }
  { static const int dbg = getenv("BSDE_WIN_DBG") ? atoi(getenv("BSDE_WIN_DBG")) : 0; a.dbg = dbg; }
  for (int e = 0; e < p.nent; ++e) {
    a.ent[e] = p.ent[e];
    a.aoff[e] = ((uint32_t)(p.ent[e].plane * p.hslab) * (uint32_t)p.region + (uint32_t)((p.lead + p.ent[e].shift) * 128)) >> 4;
    bool first = true;
    for (int f = 0; f < e; ++f) first &= p.ent[f].acc != p.ent[e].acc;
    a.dcol[e] = (uint32_t)(p.ent[e].acc * p.npad) | (first ? 0x80000000u : 0u);
  }
  // one instantiation per (producer path, epilogue, activation); activation only matters for BIAS_ACT / DACT
  using KernelFn = void (*)(const WinArgs, const CUtensorMap);
  KernelFn fn = nullptr;
  const bool uses_act = epi == BSDE_EPI_BIAS_ACT || epi == BSDE_EPI_DACT;
  const int actk = uses_act ? act : BSDE_ACT_NONE;
#define WN_PICK(E, A)                                                                        \
  if (epi == E && actk == A) fn = p.vec ? (KernelFn)conv_win_kernel<true, E, A> : (KernelFn)conv_win_kernel<false, E, A>;
  WN_PICK(BSDE_EPI_BIAS, BSDE_ACT_NONE)
  WN_PICK(BSDE_EPI_EULER, BSDE_ACT_NONE)
  WN_PICK(BSDE_EPI_ACCUM, BSDE_ACT_NONE)
  WN_PICK(BSDE_EPI_BIAS_ACT, BSDE_ACT_SOFTPLUS)
  WN_PICK(BSDE_EPI_BIAS_ACT, BSDE_ACT_TANH)
  WN_PICK(BSDE_EPI_BIAS_ACT, BSDE_ACT_ELU)
  WN_PICK(BSDE_EPI_BIAS_ACT, BSDE_ACT_RBF)
  WN_PICK(BSDE_EPI_BIAS_ACT, BSDE_ACT_NONE)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_SOFTPLUS)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_TANH)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_ELU)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_RBF)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_NONE)
#undef WN_PICK
  BSDE_CHECK_ARG(fn != nullptr, "epilogue %d / activation %d not instantiated", epi, act);
  BSDE_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  static const int pdl = getenv("BSDE_PDL") ? atoi(getenv("BSDE_PDL")) : 1;
  a.wsafe = pdl ? wsafe : 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(S * a.G); cfg.blockDim = dim3(WN_THREADS); cfg.dynamicSmemBytes = p.smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (a.tma) {
    // input tensor [S*B][hin][win][cin] fp32; box = one grid row of one 32-channel half (+ the out-of-bounds pad pixel)
    using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled enc = nullptr;
    if (!enc) {
      cudaDriverEntryPointQueryResult qres;
      BSDE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
      if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return BSDE_ERR_CUDA; }
    }
    const cuuint64_t dims[4] = {(cuuint64_t)L->cin, (cuuint64_t)L->win, (cuuint64_t)L->hin, (cuuint64_t)S * B};
    const cuuint64_t strides[3] = {(cuuint64_t)L->cin * 4, (cuuint64_t)L->win * L->cin * 4, (cuuint64_t)L->hin * L->win * L->cin * 4};
    const cuuint32_t box[4] = {32, (cuuint32_t)p.pitch, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return BSDE_ERR_CUDA; }
  }
  BSDE_CUDA_OK(cudaLaunchKernelEx(&cfg, fn, a, tmap));
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bsde
