// K2 (tensor-core mode, window-staged): the time-dependent convolution as a persistent implicit GEMM
// whose A operand is never gathered per tap.
//
// Positions.  The output grid of one sample (B images of Hq x Wq) is flattened with ONE zero pad column per
// row and ONE zero pad row per image:  q = img*IP + y*(Wq+1) + x,  IP = (Hq+1)*(Wq+1).  A tile is any 128
// consecutive q (one accumulator row each); a tap displaced by (sy, sx) reads position q + sy*(Wq+1) + sx,
// and the pad column / pad row supply the zero padding of every border.  So the A operand of every tap is
// the SAME staged window, addressed through a shared-memory descriptor whose start address is shifted by
// `shift` positions (tools/umma_probe2.cu: shifted starts are exact for the no-swizzle K-major layout).
//
// Shared memory (all K-major, no swizzle, "chunk planes": [16-byte channel chunk][row][16 B], SBO = 128 B,
// LBO = plane stride):
//   * weights of the whole layer, repacked once per (sample, iteration) by win_repack_kernel, stay resident
//     for all tiles of the CTA (<= 184 KB);
//   * the window streams through a ring of K-slabs (8 input channels = one MMA k-step each), filled by four
//     producer warps with 16-byte cp.async (zero-fill at pads): every input element is read ONCE per tile;
//   * stride-2 convolutions stage the four input parity planes (taps pick plane + shift); lhs-dilated
//     (transposed) convolutions keep four accumulators, one per output parity class, over one window.
// One thread issues tcgen05.mma.kind::tf32 (M=128, N=npad, K=8) per (k-slab, tap); accumulators are double
// buffered in TMEM so the four epilogue warps (tcgen05.ld -> warp-private smem transpose -> coalesced
// bias / time-channel / activation / Euler / data-gradient epilogue) overlap the next tile's MMAs.
// The time channel never enters the GEMM (spatially constant): t * sum of the time rows of the taps that
// fall inside the image, from a per-row tap mask.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

namespace bsde {

using namespace um;

constexpr int WN_EPI_WARPS = 8;   // warps 0-7 epilogue (TMEM lane quarter = warp & 3)
constexpr int WN_PROD = 128;      // warps 8-11 producers
constexpr int WN_MMA_WARP = 12;   // warp 12 MMA issuer + TMEM owner
constexpr int WN_THREADS = 13 * 32;
constexpr int WN_MAXENT = 16;
constexpr int WN_MAXRING = 6;

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
  int cin, cout, npad, nks;
  int hin, win, hout, wout, sn, kn, off, sd, has_time, mode;
  int Hq, Wq, pitch, IP, Q, lead, NPOSw, NP, NA, nent;
  int ring, slab_bytes, w_bytes, tb_floats, tab_units, nbuf, tmem_cols;
  int kslab, nslab, pstride;  // k-steps per slab, slabs per tile, bytes between consecutive 16-byte chunk planes
  float t, scale;
  int act, epi;
  WinEntry ent[WN_MAXENT];
};

__device__ __forceinline__ bool wn_map_axis(int o, int k, int sn, int kn, int off, int sd, int size) {
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
//   tables [0] bias, [1+e] time row of entry e, [1+nent+a] sum of the time rows of accumulator a
struct WinRepackArgs {
  const float* w;
  float* img;
  long long w_sample_stride, img_sample_stride;
  int nent, nks, npad, cin, cout, ksize, NA, real_base, has_time, time_row, use_bias, w_floats;
  long long w_off, b_off, ts, ks, ns;
  int ent_tap[WN_MAXENT], ent_acc[WN_MAXENT];
};

__global__ void win_repack_kernel(WinRepackArgs a) {
  const int s = blockIdx.y;
  const float* __restrict__ w = a.w + (long long)s * a.w_sample_stride;
  float* __restrict__ img = a.img + (long long)s * a.img_sample_stride;
  if ((int)blockIdx.x == a.nent * a.nks) {
    float* tb = img + a.w_floats;
    const int rows = 1 + a.nent + a.NA;
    for (int idx = threadIdx.x; idx < rows * a.npad; idx += blockDim.x) {
      const int r = idx / a.npad, n = idx - r * a.npad;
      float v = 0.f;
      if (n < a.cout) {
        if (r == 0) {
          v = (a.b_off >= 0 && a.use_bias) ? __ldg(w + a.b_off + n) : 0.f;
        } else if (a.has_time) {
          for (int e = 0; e < a.nent; ++e) {
            const bool take = r <= a.nent ? (e == r - 1) : (a.ent_acc[e] == r - 1 - a.nent);
            if (take) v += __ldg(w + a.w_off + a.ent_tap[e] * a.ts + (long long)a.time_row * a.ks + (long long)n * a.ns);
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
    // idx enumerates (n, kk) with n fastest for coalesced-ish reads along the output channel
    const int n = idx % a.npad, kk = idx / a.npad;
    const int c = ks * 8 + kk;
    float v = 0.f;
    if (c < a.cin && n < a.cout)
      v = __ldg(w + a.w_off + tap * a.ts + (long long)(c + a.real_base) * a.ks + (long long)n * a.ns);
    tile[((kk >> 2) * a.npad + n) * 4 + (kk & 3)] = v;
  }
}

// ---- main kernel -----------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float wn_act_fwd(float x) {
  if (ACT == BSDE_ACT_SOFTPLUS) {
    const float r = __logf(1.f + __expf(fminf(x, 15.f)));
    return x > 15.f ? x : r;
  }
  if (ACT == BSDE_ACT_TANH) return tanhf(x);
  if (ACT == BSDE_ACT_ELU) return x > 0.f ? x : __expf(x) - 1.f;
  return x;
}
template <int ACT>
__device__ __forceinline__ float wn_act_grad_out(float a) {
  if (ACT == BSDE_ACT_SOFTPLUS) return 1.f - __expf(-a);
  if (ACT == BSDE_ACT_TANH) return 1.f - a * a;
  if (ACT == BSDE_ACT_ELU) return a > 0.f ? 1.f : a + 1.f;
  return 1.f;
}
template <int EPI, int ACT>
__device__ __forceinline__ float wn_epi(float q, float x, float scale) {
  if (EPI == BSDE_EPI_BIAS_ACT) return wn_act_fwd<ACT>(q);
  if (EPI == BSDE_EPI_DACT) return scale * q * wn_act_grad_out<ACT>(x);
  if (EPI == BSDE_EPI_EULER || EPI == BSDE_EPI_ACCUM) return fmaf(scale, q, x);
  return q;
}

template <bool VEC, int EPI, int ACT>
__global__ void __launch_bounds__(WN_THREADS, 1) conv_win_kernel(const __grid_constant__ WinArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / a.G, g = blockIdx.x - s * a.G;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - raw);
  // carve
  int soff = 0;
  const uint32_t s_w = base;                                    soff += a.w_bytes;
  float* s_tb = reinterpret_cast<float*>(gen + soff);           soff += a.tb_floats * 4;
  soff = (soff + 127) & ~127;
  const uint32_t s_ring = base + soff;  uint8_t* g_ring = gen + soff;  soff += a.ring * a.slab_bytes;
  float* s_stage = reinterpret_cast<float*>(gen + soff);        soff += WN_EPI_WARPS * 32 * 16 * 4;
  int* s_tab = reinterpret_cast<int*>(gen + soff);              soff += a.NP * a.NPOSw * 4;
  int* s_pix = reinterpret_cast<int*>(gen + soff);              soff += WN_EPI_WARPS * 32 * 4;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(gen + soff);   soff += WN_EPI_WARPS * 32 * 4;
  uint64_t* s_bdesc = reinterpret_cast<uint64_t*>(gen + soff);  soff += a.nks * a.nent * 8;   // [ks][e] B descriptors
  uint32_t* s_aoff = reinterpret_cast<uint32_t*>(gen + soff);   soff += WN_MAXENT * 4;        // A start offset (16 B units)
  uint32_t* s_dcol = reinterpret_cast<uint32_t*>(gen + soff);   soff += WN_MAXENT * 4;        // acc column | first-tap flag << 31
  const uint32_t bars = base + soff;                            soff += 8 * (2 * WN_MAXRING + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + soff);
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto empty_bar = [&](int st) { return bars + 8u * (WN_MAXRING + st); };
  auto accf_bar = [&](int b) { return bars + 8u * (2 * WN_MAXRING + b); };
  auto acce_bar = [&](int b) { return bars + 8u * (2 * WN_MAXRING + 2 + b); };

  // resident weights + tables: one cooperative copy per CTA
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
      for (int i = 0; i < a.ring; ++i) { mbar_init(full_bar(i), WN_PROD); mbar_init(empty_bar(i), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(accf_bar(b), 1); mbar_init(acce_bar(b), WN_EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
    // descriptor tables of the issue loop
    const uint64_t b_hi = desc_hi((uint32_t)a.npad * 16u, 128u, 0);
    for (int idx = lane; idx < a.nks * a.nent; idx += 32) {
      const int ks = idx / a.nent, e = idx - ks * a.nent;
      s_bdesc[idx] = b_hi | desc_addr(s_w + (uint32_t)((e * a.nks + ks) * a.npad * 32));
    }
    if (lane < a.nent) {
      const WinEntry en = a.ent[lane];
      s_aoff[lane] = ((uint32_t)(en.plane * a.kslab * 2) * (uint32_t)a.pstride + (uint32_t)((a.lead + en.shift) * 16)) >> 4;
      bool first = true;
      for (int e = 0; e < lane; ++e) first &= a.ent[e].acc != en.acc;
      s_dcol[lane] = (uint32_t)(en.acc * a.npad) | (first ? 0x80000000u : 0u);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= WN_EPI_WARPS && warp < WN_MMA_WARP) {
    // =============================== producers ===============================
    const float* __restrict__ in = a.in + (long long)s * a.B * a.hin * a.win * a.cin;
    const int pt = tid - WN_EPI_WARPS * 32;
    uint32_t cnt = 0;
#pragma unroll 1
    for (int tile = g; tile < a.ntiles; tile += a.G) {
      const int q0 = tile * 128 - a.lead;
      // decode the window positions once per tile into the shared source table (read by other producer threads)
      asm volatile("bar.sync 1, %0;" ::"n"(WN_PROD) : "memory");  // last tile's readers are done
#pragma unroll 1
      for (int i = pt; i < a.NPOSw; i += WN_PROD) {
        const int q = q0 + i;
        int img = 0, y = 0, x = 0;
        bool pv = q >= 0 && q < a.Q;
        if (pv) {
          img = q / a.IP;
          const int r = q - img * a.IP;
          y = r / a.pitch;
          x = r - y * a.pitch;
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
        const int cbase = sl * a.kslab * 8;                       // first channel of this slab
        if (VEC) {
          // lanes run over the 16-byte chunks of a position first: a warp reads whole 128-byte lines
          const int nch = min(a.kslab * 2, (a.cin - cbase + 3) >> 2);   // real chunks in this slab
          const int nchp = a.kslab * 2;                                   // chunk planes per plane set
          int cpw = 1;                                                    // power of two >= nchp
          while (cpw < nchp) cpw <<= 1;
          const int c = pt & (cpw - 1), pstep = WN_PROD / cpw;
#pragma unroll 1
          for (int plane = 0; plane < a.NP; ++plane) {
            const uint32_t d0 = sb + (uint32_t)(plane * nchp + c) * (uint32_t)a.pstride;
            if (c < nchp) {
#pragma unroll 2
              for (int i = pt / cpw; i < a.NPOSw; i += pstep) {
                const int src = s_tab[plane * a.NPOSw + i];
                const bool ok = src >= 0 && c < nch;
                cp_async16(d0 + i * 16, in + (ok ? src + cbase + c * 4 : 0), ok);
              }
            }
          }
        } else {
#pragma unroll 1
          for (int plane = 0; plane < a.NP; ++plane) {
            const uint32_t d0 = sb + (uint32_t)(plane * 2) * (uint32_t)a.pstride;
#pragma unroll 1
            for (int i = pt; i < a.NPOSw; i += WN_PROD) {
              const int src = s_tab[plane * a.NPOSw + i];
              const bool ok = src >= 0;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch)
                if (cbase + ch < a.cin)
                  cp_async4(d0 + (ch >> 2) * a.pstride + i * 16 + (ch & 3) * 4, in + (ok ? src + cbase + ch : 0), ok);
            }
          }
        }
        cp_async_arrive_noinc(full_bar(slot));
      }
    }
  } else if (warp == WN_MMA_WARP) {
    // =============================== MMA issuer ===============================
    // the whole warp runs the loop (uniform address arithmetic); lane 0 issues
    {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_hi = desc_hi((uint32_t)a.pstride, 128u, 0);
      const uint32_t kstep16 = (uint32_t)(2 * a.pstride) >> 4;
      uint32_t cnt = 0;
      int it = 0;
#pragma unroll 1
      for (int tile = g; tile < a.ntiles; tile += a.G, ++it) {
        const int buf = it % a.nbuf;
        if (it >= a.nbuf) {
          mbar_wait(acce_bar(buf), ((it / a.nbuf) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t dbase = tmem_base + (uint32_t)(buf * a.NA * a.npad);
#pragma unroll 1
        for (int sl = 0; sl < a.nslab; ++sl, ++cnt) {
          const int slot = cnt % a.ring;
          mbar_wait(full_bar(slot), (cnt / a.ring) & 1);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t sb16 = (s_ring + slot * a.slab_bytes) >> 4;
          const int kend = min(a.kslab, a.nks - sl * a.kslab);
#pragma unroll 1
          for (int kk = 0; kk < kend; ++kk) {
            const int ks = sl * a.kslab + kk;
            const uint32_t ka = sb16 + kk * kstep16;
            const uint64_t* bd = s_bdesc + ks * a.nent;
            const uint32_t later = ks > 0 ? 1u : 0u;
#pragma unroll 3
            for (int e = 0; e < a.nent; ++e) {
              const uint32_t dc = s_dcol[e];
              const uint64_t ad = a_hi | (uint64_t)((ka + s_aoff[e]) & 0x3FFFu);
              if (lane == 0) umma_tf32(dbase + (dc & 0xFFFFu), ad, bd[e], idesc, later | ((dc >> 31) ^ 1u));
            }
          }
          __syncwarp();
          if (lane == 0) umma_commit(empty_bar(slot));
        }
        if (lane == 0) umma_commit(accf_bar(buf));
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    // warp w: TMEM lane quarter w & 3 (rows), and every second (accumulator, 16-column chunk) work item
    const int quarter = warp & 3, half = warp >> 2;
    float* stage = s_stage + warp * 32 * 16;
    int* w_pix = s_pix + warp * 32;
    uint32_t* w_mask = s_mask + warp * 32;
    float* __restrict__ out = a.out + (long long)s * a.B * a.hout * a.wout * a.cout;
    const float* __restrict__ aux = a.aux ? a.aux + (long long)s * a.B * a.hout * a.wout * a.cout : nullptr;
    const bool vec_out = (a.cout & 3) == 0;
    const int nchunk = (a.cout + 15) >> 4;
    const int c4 = lane & 3, r8 = lane >> 2;
    int it = 0;
#pragma unroll 1
    for (int tile = g; tile < a.ntiles; tile += a.G, ++it) {
      const int buf = it % a.nbuf;
      const int q = tile * 128 + quarter * 32 + lane;
      bool valid = q < a.Q;
      int img = 0, y = 0, x = 0;
      if (valid) {
        img = q / a.IP;
        const int r = q - img * a.IP;
        y = r / a.pitch;
        x = r - y * a.pitch;
        valid = y < a.Hq && x < a.Wq;
      }
      // every tap of an interior position falls inside the image
      const bool interior = y >= 1 && x >= 1 && y < a.Hq - 2 && x < a.Wq - 2;
      mbar_wait(accf_bar(buf), (it / a.nbuf) & 1);
      tc_fence_after();
      int last_ac = -1;
#pragma unroll 1
      for (int item = half; item < a.NA * nchunk; item += 2) {
        const int ac = item / nchunk, n0 = (item - ac * nchunk) * 16;
        if (ac != last_ac) {
          last_ac = ac;
          int oy = y, ox = x;
          if (a.mode == 3) { oy = 2 * y + (ac >> 1); ox = 2 * x + (ac & 1); }
          const bool oval = valid && oy < a.hout && ox < a.wout;
          uint32_t mask = 0xFFFFFFFFu;
          if (a.has_time && oval && !interior) {
            mask = 0;
            bool all = true;
            for (int e = 0; e < a.nent; ++e) {
              if (a.ent[e].acc != ac) continue;
              const bool ok = wn_map_axis(oy, a.ent[e].kh, a.sn, a.kn, a.off, a.sd, a.hin) &&
                              wn_map_axis(ox, a.ent[e].kw, a.sn, a.kn, a.off, a.sd, a.win);
              if (ok) mask |= 1u << e; else all = false;
            }
            if (all) mask = 0xFFFFFFFFu;
          }
          __syncwarp();
          w_pix[lane] = oval ? (img * a.hout + oy) * a.wout + ox : -1;
          w_mask[lane] = mask;
        }
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((buf * a.NA + ac) * a.npad);
        const float* tsum = s_tb + (1 + a.nent + ac) * a.npad;
        uint32_t v[16];
        tmem_ld16(trow + n0, v);
        __syncwarp();  // readers of the previous chunk are done with the staging rows
        {
          // row `lane`: 16-byte chunk c stored at position c ^ ((lane >> 1) & 3)  (conflict-free both ways)
          uint4* dst = reinterpret_cast<uint4*>(stage + lane * 16);
          const int f = (lane >> 1) & 3;
          dst[0 ^ f] = make_uint4(v[0], v[1], v[2], v[3]);
          dst[1 ^ f] = make_uint4(v[4], v[5], v[6], v[7]);
          dst[2 ^ f] = make_uint4(v[8], v[9], v[10], v[11]);
          dst[3 ^ f] = make_uint4(v[12], v[13], v[14], v[15]);
        }
        __syncwarp();
        const int ncol = min(16, a.cout - n0);
        if (vec_out && ncol == 16) {
          // lane -> fixed column group c4, rows r8 + 8 j
          const int n = n0 + c4 * 4;
          const float4 b4 = *reinterpret_cast<const float4*>(s_tb + n);
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.has_time) {
            t4 = *reinterpret_cast<const float4*>(tsum + n);
            t4.x *= a.t; t4.y *= a.t; t4.z *= a.t; t4.w *= a.t;
          }
          int pp[4];
          uint32_t mk[4];
          float4 qv[4], xv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = r8 + 8 * j;
            pp[j] = w_pix[rr];
            mk[j] = w_mask[rr];
            qv[j] = *reinterpret_cast<const float4*>(stage + rr * 16 + ((c4 ^ ((rr >> 1) & 3)) << 2));
          }
          if (EPI >= BSDE_EPI_EULER) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              xv[j] = pp[j] >= 0 ? *reinterpret_cast<const float4*>(aux + (long long)pp[j] * a.cout + n) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (pp[j] < 0) continue;
            float4 qq = qv[j];
            qq.x += b4.x; qq.y += b4.y; qq.z += b4.z; qq.w += b4.w;
            if (a.has_time) {
              if (mk[j] == 0xFFFFFFFFu) {
                qq.x += t4.x; qq.y += t4.y; qq.z += t4.z; qq.w += t4.w;
              } else {
                for (int e = 0; e < a.nent; ++e)
                  if ((mk[j] >> e) & 1) {
                    const float4 tt = *reinterpret_cast<const float4*>(s_tb + (1 + e) * a.npad + n);
                    qq.x = fmaf(a.t, tt.x, qq.x); qq.y = fmaf(a.t, tt.y, qq.y); qq.z = fmaf(a.t, tt.z, qq.z); qq.w = fmaf(a.t, tt.w, qq.w);
                  }
              }
            }
            qq.x = wn_epi<EPI, ACT>(qq.x, xv[j].x, a.scale); qq.y = wn_epi<EPI, ACT>(qq.y, xv[j].y, a.scale);
            qq.z = wn_epi<EPI, ACT>(qq.z, xv[j].z, a.scale); qq.w = wn_epi<EPI, ACT>(qq.w, xv[j].w, a.scale);
            *reinterpret_cast<float4*>(out + (long long)pp[j] * a.cout + n) = qq;
          }
        } else {
          // generic tail / scalar path (cout not a multiple of 4, or a partial chunk): rows r8 + 8 j, columns c4 + 4 i
#pragma unroll 1
          for (int j = 0; j < 4; ++j) {
            const int rr = r8 + 8 * j;
            const int pq = w_pix[rr];
            if (pq < 0) continue;
            const uint32_t mkk = w_mask[rr];
#pragma unroll 1
            for (int cc = c4; cc < ncol; cc += 4) {
              const int n = n0 + cc;
              float qq = stage[rr * 16 + ((((cc >> 2) ^ ((rr >> 1) & 3)) << 2) | (cc & 3))] + s_tb[n];
              if (a.has_time) {
                if (mkk == 0xFFFFFFFFu) {
                  qq = fmaf(a.t, tsum[n], qq);
                } else {
                  for (int e = 0; e < a.nent; ++e)
                    if ((mkk >> e) & 1) qq = fmaf(a.t, s_tb[(1 + e) * a.npad + n], qq);
                }
              }
              const long long o = (long long)pq * a.cout + n;
              const float xx = EPI >= BSDE_EPI_EULER ? aux[o] : 0.f;
              out[o] = wn_epi<EPI, ACT>(qq, xx, a.scale);
            }
          }
        }
      }
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
  int tab_units, nbuf, tmem_cols, vec, kslab, nslab, pstride;
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
  p.pstride = p.NPOSw * 16 + 16;  // +16 B: consecutive chunk planes start 4 banks apart (conflict-free producer writes)
  p.w_bytes = p.nent * p.nks * p.npad * 32;
  p.tb_floats = (1 + p.nent + p.NA) * p.npad;
  p.tab_units = p.NP * p.NPOSw;
  p.nbuf = (2 * p.NA * p.npad <= 512) ? 2 : 1;
  if (p.NA * p.npad > 512) return p;
  p.tmem_cols = wn_pow2_cols(p.nbuf * p.NA * p.npad);
  const size_t fixed = 128 + (size_t)p.w_bytes + (size_t)p.tb_floats * 4 + 128 + WN_EPI_WARPS * 32 * 16 * 4 + (size_t)p.tab_units * 4 +
                       2 * WN_EPI_WARPS * 32 * 4 + 8 * (2 * WN_MAXRING + 4) + 16 + (size_t)p.nks * p.nent * 8 + 2 * WN_MAXENT * 4;
  const size_t cap = 227 * 1024;
  // k-steps per slab: as many as fit with a ring of >= 2 slabs (>= 3 when a tile needs several slabs), so that a
  // warp-level cp.async covers whole 128-byte lines (4 k-steps = 32 channels) instead of 32-byte sectors
  p.kslab = 0;
  const int cands[5] = {p.nks, 8, 4, 2, 1};
  for (int ci = 0; ci < 5 && !p.kslab; ++ci) {
    const int k = cands[ci];
    if (k > p.nks || k < 1) continue;
    if (!p.vec && k != 1) continue;
    const size_t slab = (size_t)p.NP * k * 2 * p.pstride;
    const int need = (k >= p.nks) ? 2 : (k >= 4 ? 2 : 3);
    if (fixed + need * slab <= cap) { p.kslab = k; p.slab_bytes = (int)slab; }
  }
  if (!p.kslab) return p;
  // 8-channel slabs make every lane of a cp.async touch its own 128-byte line (measured ~3 B/cycle/SM): leave such
  // shapes (stride-2 layers with 64x64 weights resident) to the per-tap gather kernel of tconv_tc.cu
  if (p.vec && p.kslab < 4 && p.nks >= 4) return p;
  p.nslab = (p.nks + p.kslab - 1) / p.kslab;
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

bool tconv_win_supported(const bsde_conv_layer* L, int S, int B) {
  if (!win_enabled() || !L || S < 1 || S > 148 || B < 1) return false;
  WinPlan p = win_plan(L);
  if (!p.ok) return false;
  // 32-bit position / offset arithmetic inside the kernel
  if ((long long)B * p.IP + 1024 > 0x7FFFFFFFLL) return false;
  if ((long long)B * L->hin * L->win * L->cin > 0x7FFFFFFFLL) return false;
  if ((long long)B * L->hout * L->wout * L->cout > 0x7FFFFFFFLL) return false;
  return true;
}

size_t tconv_win_ws_bytes(const bsde_conv_layer* L, int S) {
  WinPlan p = win_plan(L);
  return p.ok ? p.img_floats * sizeof(float) * S : 0;
}

int tconv_fwd_win(const bsde_conv_layer* L, int S, int B, const float* in, const float* w, long long w_sample_stride,
                  float t, int act, int epi, float scale, const float* aux, float* out, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  WinPlan p = win_plan(L);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the window-staged tensor-core path");
  BSDE_CHECK_ARG(!(epi >= BSDE_EPI_EULER) || aux != nullptr, "epilogue %d needs aux", epi);
  const size_t need = p.img_floats * sizeof(float) * S;
  if (!ws || ws_bytes < need) {
    set_error("tconv_fwd(win): workspace %zu < %zu bytes", ws_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  {
    WinRepackArgs r;
    memset(&r, 0, sizeof(r));
    r.w = w; r.img = (float*)ws; r.w_sample_stride = w_sample_stride; r.img_sample_stride = (long long)p.img_floats;
    r.nent = p.nent; r.nks = p.nks; r.npad = p.npad; r.cin = L->cin; r.cout = L->cout; r.ksize = L->ksize; r.NA = p.NA;
    r.real_base = (L->has_time && L->time_first) ? 1 : 0; r.has_time = L->has_time;
    r.time_row = L->time_first ? 0 : L->cin; r.use_bias = epi <= BSDE_EPI_EULER; r.w_floats = p.w_bytes / 4;
    r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
    for (int e = 0; e < p.nent; ++e) { r.ent_tap[e] = p.ent_tap[e]; r.ent_acc[e] = p.ent[e].acc; }
    win_repack_kernel<<<dim3(p.nent * p.nks + 1, S), 256, 0, st>>>(r);
    BSDE_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  WinArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.aux = aux; a.out = out; a.wimg = (const float*)ws; a.wimg_sample_stride = (long long)p.img_floats;
  a.S = S; a.B = B;
  a.G = std::max(1, 148 / S);
  a.Q = B * p.IP;
  a.ntiles = (a.Q + 127) / 128;
  a.G = std::min(a.G, a.ntiles);
  a.cin = L->cin; a.cout = L->cout; a.npad = p.npad; a.nks = p.nks;
  a.hin = L->hin; a.win = L->win; a.hout = L->hout; a.wout = L->wout;
  a.sn = L->sn; a.kn = L->kn; a.off = L->off; a.sd = L->sd; a.has_time = L->has_time; a.mode = p.mode;
  a.Hq = p.Hq; a.Wq = p.Wq; a.pitch = p.pitch; a.IP = p.IP; a.lead = p.lead; a.NPOSw = p.NPOSw; a.NP = p.NP; a.NA = p.NA;
  a.nent = p.nent; a.ring = p.ring; a.slab_bytes = p.slab_bytes; a.w_bytes = p.w_bytes; a.tb_floats = p.tb_floats;
  a.tab_units = p.tab_units; a.nbuf = p.nbuf; a.tmem_cols = p.tmem_cols;
  a.kslab = p.kslab; a.nslab = p.nslab; a.pstride = p.pstride;
  a.t = t; a.scale = scale; a.act = act; a.epi = epi;
  for (int e = 0; e < p.nent; ++e) a.ent[e] = p.ent[e];
  // one instantiation per (producer path, epilogue, activation); activation only matters for BIAS_ACT / DACT
  using KernelFn = void (*)(const WinArgs);
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
  WN_PICK(BSDE_EPI_BIAS_ACT, BSDE_ACT_NONE)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_SOFTPLUS)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_TANH)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_ELU)
  WN_PICK(BSDE_EPI_DACT, BSDE_ACT_NONE)
#undef WN_PICK
  BSDE_CHECK_ARG(fn != nullptr, "epilogue %d / activation %d not instantiated", epi, act);
  BSDE_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  fn<<<S * a.G, WN_THREADS, p.smem, st>>>(a);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bsde
