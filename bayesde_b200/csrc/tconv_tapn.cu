// K2, small-cout layers (64 -> 5 at 32x32: the last conv of an SDEBNN block and the data gradient of its first conv):
// "taps on the N side".
//
// conv_win_kernel issues nine MMAs per k-step for a 3x3 layer, one per tap, each reading the same 128-row A operand
// through a row-shifted descriptor; with cout = 5 (N padded to 16) the tensor-core shared-memory port carries the A tile
// nine times (ncu, round 1: 22.6 M wavefronts = 81 of the launch's 168-173 us).  Here ONE product per k-step covers all
// nine taps,
//        T[p][tap * cout + co] = sum_ci in[p][ci] * w[tap][ci][co]           (M = 128 INPUT positions, N = 9 cout -> 48)
// and the tap shift moves into the epilogue:
//        out[q][co] = sum_tap T[q + s_tap][tap * cout + co],   s_tap = dy * pitch + dx
// on the same flattened padded position grid as conv_win (q = img * IP + y * pitch + x, one zero pad column per row and
// one zero pad row per image, supplied by the TMA out-of-bounds fill), so the A operand is read from shared memory once
// per k-step and needs no halo.  The index math was checked against the direct convolution in tools/proto_taps_on_n.py.
//
// A persistent CTA owns a contiguous range of output tiles and walks the T tiles of that range plus one halo tile per side
// (|s_tap| <= pitch + 1 <= 128).  Warp roles: one TMA producer warp (row boxes {32 channels, pitch pixels} into a
// SWIZZLE_128B K-major ring), one MMA warp (tcgen05.mma kind::tf32, M = 128, N = npad, K = 8; accumulators double
// buffered in TMEM), eight epilogue warps: stage 1 copies the finished T tile from TMEM into a three-slot shared-memory
// ring (row stride npad + 1 floats: conflict free), stage 2 forms the outputs of the PREVIOUS tile -- nine gathered adds,
// bias, time term (border-class table, as conv_win), Euler / accumulate combine -- with channel-fastest thread order so
// that the 20-byte output rows of a warp are one contiguous span.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

namespace bsde {

using namespace um;

constexpr int TN_EPI_WARPS = 8;
constexpr int TN_EPI_THREADS = TN_EPI_WARPS * 32;
constexpr int TN_PROD_WARP = 8;
constexpr int TN_MMA_WARP = 9;
constexpr int TN_THREADS = 10 * 32;
constexpr int TN_RING = 3;     // A-operand stages
constexpr int TN_PSLOTS = 3;   // T tiles kept in shared memory
constexpr int TN_MAXN = 64;    // npad <= 64

struct TapnArgs {
  const float* in;
  const float* aux;
  float* out;
  const float* wimg;
  long long wimg_sample_stride;  // floats
  int S, B, G, ntiles;
  int cin, cout, npad, nks, nhalf, opitch;
  int hout, wout, has_time, epi;
  int Hq, Wq, pitch, IP, Q, NR, box_bytes, region, slab_bytes;
  int w_bytes, tb_floats, prow;
  float t, scale;
  int dbg;   // tools only (BSDE_TAPN_DBG): 1 skip TMA loads, 2 skip MMAs, 4 skip stage 2, 8 skip stage 1 copies
  int shift[9];
};

// ---- weight repack: flat w(t_k) -> [ks][chunk c in 0..1][n in 0..npad)[4 floats] (K-major chunk planes, n = tap * cout + co)
// followed by the tables [0] bias, [1 + cls] time term of border class cls (sum of the time rows of the taps inside the image)
struct TapnRepackArgs {
  const float* w;
  float* img;
  long long w_sample_stride, img_sample_stride, w_iter_stride, img_iter_stride;
  int nks, npad, cin, cout, real_base, has_time, time_row, use_bias, w_floats, tb_rows, Hq, Wq, off, kn;
  long long w_off, b_off, ts, ks, ns;
};

__device__ __forceinline__ int tn_class_rep(int f, int size) { return f == 0 ? 1 : (f == 2 ? size - 1 : 0); }

__global__ void tapn_repack_kernel(TapnRepackArgs a) {
  const int s = blockIdx.y;
  const float* __restrict__ w = a.w + (long long)s * a.w_sample_stride + (long long)blockIdx.z * a.w_iter_stride;
  float* __restrict__ img = a.img + (long long)s * a.img_sample_stride + (long long)blockIdx.z * a.img_iter_stride;
  if ((int)blockIdx.x == a.nks) {
    float* tb = img + a.w_floats;
    for (int idx = threadIdx.x; idx < a.tb_rows * 8; idx += blockDim.x) {
      const int r = idx >> 3, n = idx & 7;
      float v = 0.f;
      if (n < a.cout) {
        if (r == 0) {
          v = (a.b_off >= 0 && a.use_bias) ? __ldg(w + a.b_off + n) : 0.f;
        } else {
          const int cls = r - 1;
          const int y = tn_class_rep(cls >> 2, a.Hq), x = tn_class_rep(cls & 3, a.Wq);
          for (int tap = 0; tap < 9; ++tap) {
            const int iy = y + (tap / 3) * a.kn + a.off, ix = x + (tap % 3) * a.kn + a.off;
            if (iy < 0 || iy >= a.Hq || ix < 0 || ix >= a.Wq) continue;
            v += __ldg(w + a.w_off + tap * a.ts + (long long)a.time_row * a.ks + (long long)n * a.ns);
          }
        }
      }
      tb[idx] = v;
    }
    return;
  }
  const int ks = blockIdx.x;
  float* __restrict__ tile = img + (long long)ks * a.npad * 8;
  for (int idx = threadIdx.x; idx < a.npad * 8; idx += blockDim.x) {
    const int n = idx % a.npad, kk = idx / a.npad;
    const int c = ks * 8 + kk;
    float v = 0.f;
    if (c < a.cin && n < 9 * a.cout) {
      const int tap = n / a.cout, co = n - tap * a.cout;
      v = __ldg(w + a.w_off + tap * a.ts + (long long)(c + a.real_base) * a.ks + (long long)co * a.ns);
    }
    tile[((kk >> 2) * a.npad + n) * 4 + (kk & 3)] = v;
  }
}

// ---- main kernel -----------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(TN_THREADS, 1) conv_tapn_kernel(const __grid_constant__ TapnArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / a.G, g = blockIdx.x - s * a.G;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  int soff = 0;
  const uint32_t s_ring = base;                                   soff += TN_RING * a.slab_bytes;
  const uint32_t s_w = base + soff;                               soff += a.w_bytes;
  float* s_tb = reinterpret_cast<float*>(gen + soff);             soff += a.tb_floats * 4;
  float* s_p = reinterpret_cast<float*>(gen + soff);              soff += TN_PSLOTS * 128 * a.prow * 4;
  const uint32_t bars = base + soff;                              soff += 8 * (2 * TN_RING + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + soff);
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto empty_bar = [&](int st) { return bars + 8u * (TN_RING + st); };
  auto accf_bar = [&](int b) { return bars + 8u * (2 * TN_RING + b); };
  auto acce_bar = [&](int b) { return bars + 8u * (2 * TN_RING + 2 + b); };

  // this CTA's contiguous range of output tiles and the T tiles it needs (one halo tile per side)
  const int per = a.ntiles / a.G, rem = a.ntiles - per * a.G;
  const int t0 = g * per + min(g, rem), t1 = t0 + per + (g < rem ? 1 : 0);
  const int jlo = max(t0 - 1, 0), jhi = min(t1, a.ntiles - 1);

  // resident weights + tables
  {
    const float* __restrict__ wsrc = a.wimg + (long long)s * a.wimg_sample_stride;
    const int n16 = (a.w_bytes + a.tb_floats * 4) >> 4;
    for (int i = tid; i < n16; i += TN_THREADS) cp_async16(s_w + i * 16, wsrc + i * 4, true);
    cp_async_commit();
    cp_async_wait_all();
  }
  if (warp == TN_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < TN_RING; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(accf_bar(b), 1); mbar_init(acce_bar(b), TN_EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 128u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (t1 > t0) {
    if (warp == TN_PROD_WARP) {
      // =============================== TMA producer ===============================
      const int rows_img = a.Hq + 1;
      uint32_t cnt = 0;
#pragma unroll 1
      for (int j = jlo; j <= jhi; ++j, ++cnt) {
        const int n0 = j * 128;
        const int g0 = n0 / a.pitch + lane;          // grid row of window row `lane`
        const int img = g0 / rows_img;
        const int y = g0 - img * rows_img;
        const int slot = cnt % TN_RING;
        if (cnt >= (uint32_t)TN_RING) mbar_wait(empty_bar(slot), ((cnt / TN_RING) & 1) ^ 1);
        const uint32_t sb = s_ring + slot * a.slab_bytes;
        if (a.dbg & 1) { if (lane == 0) mbar_arrive(full_bar(slot)); continue; }
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(slot)), "r"(a.nhalf * a.NR * a.box_bytes) : "memory");
        __syncwarp();
        if (lane < a.NR) {
          for (int h = 0; h < a.nhalf; ++h) {
            const uint32_t dst = sb + (uint32_t)h * (uint32_t)a.region + (uint32_t)(lane * a.box_bytes);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                "l"(&tmap), "r"(full_bar(slot)), "r"(h * 32), "r"(0), "r"(y), "r"(s * a.B + img)
                : "memory");
          }
        }
      }
    } else if (warp == TN_MMA_WARP) {
      // =============================== MMA issuer ===============================
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_hi = desc_hi(16u, 1024u, 2);  // K-major SWIZZLE_128B, 8-row groups 1024 B apart
      const uint64_t b_hi = desc_hi((uint32_t)a.npad * 16u, 128u, 0);
      const uint32_t w16 = s_w >> 4, bstep = (uint32_t)(a.npad * 32) >> 4, region16 = (uint32_t)a.region >> 4;
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t cnt = 0;
#pragma unroll 1
      for (int j = jlo; j <= jhi; ++j, ++cnt) {
        const int buf = cnt & 1;
        if (cnt >= 2u) {
          mbar_wait(acce_bar(buf), ((cnt >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const uint32_t dbase = tm + (uint32_t)(buf * a.npad);
        const int n0 = j * 128;
        const uint32_t o0 = (uint32_t)(n0 - (n0 / a.pitch) * a.pitch) * 8u;   // positions of the first window row before n0, 16-byte units
        const int slot = cnt % TN_RING;
        mbar_wait(full_bar(slot), (cnt / TN_RING) & 1);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t sb16 = (s_ring + slot * a.slab_bytes) >> 4;
#pragma unroll 1
        for (int ks = 0; ks < ((a.dbg & 2) ? 0 : a.nks); ++ks) {
          const uint32_t ka = sb16 + o0 + (uint32_t)(ks >> 2) * region16 + (uint32_t)(ks & 3) * 2u;
          const uint32_t kb = w16 + (uint32_t)ks * bstep;
          umma_tf32_elect(dbase, a_hi | (uint64_t)(ka & 0x3FFFu), b_hi | (uint64_t)(kb & 0x3FFFu), idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit_elect(empty_bar(slot));
        umma_commit_elect(accf_bar(buf));
      }
      __syncwarp();
    } else {
      // =============================== epilogue ===============================
      const int quarter = warp & 3, half = warp >> 2;
      const int prow = a.prow;
      float* __restrict__ out = a.out + (long long)s * a.B * a.hout * a.wout * a.opitch;
      const float* __restrict__ aux = a.aux ? a.aux + (long long)s * a.B * a.hout * a.wout * a.opitch : nullptr;
      const float e_rIP = 1.0f / (float)a.IP, e_rpitch = 1.0f / (float)a.pitch;
      const int nch = a.npad >> 4;                 // 16-column chunks of a T row
      const int c_lo = half == 0 ? 0 : (nch + 1) / 2, c_hi = half == 0 ? (nch + 1) / 2 : nch;
      constexpr int PR = TN_PSLOTS * 128;   // ring rows
      const int wrap = PR * prow;
      // stage 2 work split: the two lanes of a pair share output row r = tid / 2; lane `part` adds the taps 2 i + part
      // (i = 0..4; tap 9 does not exist), the partner's sums arrive by one shuffle per channel, the even lane finishes the row.
      // Per-tap displacement inside the ring: row shift * prow + column block of the tap (uniform values, selected per lane).
      const int r2 = tid >> 1, part = tid & 1;
      int dsp[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int tap = min(2 * i + part, 8);
        dsp[i] = a.shift[tap] * prow + tap * COUT;
      }
      uint32_t cnt = 0;
      // T tiles that do not exist (before the first / after the last tile of the sample) read as zero: their ring slots are
      // zero filled, so stage 2 needs no per-tap range test
      if (jlo == 0 && t0 == 0) {
        float* z = s_p + ((TN_PSLOTS - 1) * 128) * prow;   // slot of tile -1
        for (int i = tid; i < 128 * prow; i += TN_EPI_THREADS) z[i] = 0.f;
      }
#pragma unroll 1
      for (int j = jlo; j <= t1; ++j) {
        // ---- outputs of tile jo = j - 1: decode this pair's row and issue its aux loads NOW, so that the (L2 / DRAM)
        // latency runs under stage 1 and the barrier
        const int jo = j - 1;
        const bool do_out = jo >= t0 && jo < t1;
        int obase = -1, cls = 0;   // float index of channel 0 of the output pixel inside the sample (< 2^31, checked on the host)
        float xv[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) xv[c] = 0.f;
        if (do_out) {
          const int q = jo * 128 + r2;
          if (q < a.Q) {
            int img = (int)((float)q * e_rIP);
            int rr = q - img * a.IP;
            if (rr < 0) { --img; rr += a.IP; } else if (rr >= a.IP) { ++img; rr -= a.IP; }
            int y = (int)((float)rr * e_rpitch);
            int x = rr - y * a.pitch;
            if (x < 0) { --y; x += a.pitch; } else if (x >= a.pitch) { ++y; x -= a.pitch; }
            if (y < a.Hq && x < a.Wq) {
              cls = (((y == 0) | ((y == a.Hq - 1) << 1)) << 2) | ((x == 0) | ((x == a.Wq - 1) << 1));
              obase = ((img * a.hout + y) * a.wout + x) * a.opitch;
              if (part == 0 && a.epi >= BSDE_EPI_EULER) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) xv[c] = aux[obase + c];
              }
            }
          }
        }
        if (j <= jhi) {
          // ---- stage 1: T tile j, TMEM -> shared ring slot j % 3 (this warp: rows of its lane quarter, its chunk range)
          const int buf = cnt & 1;
          mbar_wait(accf_bar(buf), (cnt >> 1) & 1);
          tc_fence_after();
          float* prowp = s_p + ((j % TN_PSLOTS) * 128 + quarter * 32 + lane) * prow;
#pragma unroll 1
          for (int c = c_lo; c < ((a.dbg & 8) ? c_lo : c_hi); ++c) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * a.npad + c * 16), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) prowp[c * 16 + i] = __uint_as_float(v[i]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acce_bar(buf));
          ++cnt;
        } else {
          float* z = s_p + ((j % TN_PSLOTS) * 128) * prow;   // tile j == ntiles does not exist
          for (int i = tid; i < 128 * prow; i += TN_EPI_THREADS) z[i] = 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TN_EPI_THREADS) : "memory");   // slot j % 3 is complete
        // ---- stage 2: outputs of tile jo from the T tiles jo - 1, jo, jo + 1 (ring rows, wrapped)
        if (do_out && !(a.dbg & 4)) {
          const int pb = ((jo % TN_PSLOTS) * 128 + r2) * prow;
          float acc[COUT];
#pragma unroll
          for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            int k = pb + dsp[i];
            k += k < 0 ? wrap : 0;          // row below 0 (the column part stays below prow: no carry)
            k -= k >= wrap ? wrap : 0;      // row above the ring
            if (i < 4 || part == 0) {
#pragma unroll
              for (int c = 0; c < COUT; ++c) acc[c] += s_p[k + c];
            }
          }
#pragma unroll
          for (int c = 0; c < COUT; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
          if (part == 0 && obase >= 0) {
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
              float v = acc[c] + s_tb[c];
              if (a.has_time) v = fmaf(a.t, s_tb[(1 + cls) * 8 + c], v);
              out[obase + c] = a.epi >= BSDE_EPI_EULER ? fmaf(a.scale, v, xv[c]) : v;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TN_EPI_THREADS) : "memory");   // readers of slot (j - 2) % 3 are done
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TN_MMA_WARP) tmem_dealloc(tmem_base, 128u);
}

// ---- host side ---------------------------------------------------------------------------------------------
struct TapnPlan {
  bool ok;
  int Hq, Wq, pitch, IP, nks, npad, nhalf, NR, box_bytes, region, slab_bytes, w_bytes, tb_rows, tb_floats, prow;
  size_t smem, img_floats;
};

static TapnPlan tapn_plan(const bsde_conv_layer* L) {
  TapnPlan p;
  memset(&p, 0, sizeof(p));
  if (!L || L->ksize != 3) return p;
  if (L->sd != 1 || L->sn != 1) return p;
  if (!((L->kn == 1 && L->off == -1) || (L->kn == -1 && L->off == 1))) return p;   // stride-1 SAME 3x3: forward form or data-gradient form
  if (L->hin != L->hout || L->win != L->wout) return p;
  if (L->cin % 32 != 0 || L->cin < 32 || L->cin > 128) return p;               // TMA-fed 32-channel halves
  if (L->cout < 1 || L->cout > 7) return p;                                    // 9 cout <= 64, table rows of 8 floats
  p.Hq = L->hout; p.Wq = L->wout;
  p.pitch = p.Wq + 1;
  if (p.pitch + 1 > 128 || p.pitch > 256) return p;                            // one halo tile per side; TMA box limit
  p.IP = (p.Hq + 1) * p.pitch;
  p.nks = L->cin / 8;
  p.npad = (9 * L->cout + 15) / 16 * 16;
  if (p.npad > TN_MAXN) return p;
  p.nhalf = L->cin / 32;
  p.NR = (128 + p.pitch - 2) / p.pitch + 1;                                     // grid rows any 128 consecutive positions can touch
  if (p.NR > 32) return p;
  p.box_bytes = p.pitch * 128;
  p.region = (p.NR * p.box_bytes + 1023) & ~1023;
  p.slab_bytes = p.nhalf * p.region;
  p.w_bytes = p.nks * p.npad * 32;
  p.tb_rows = L->has_time ? 17 : 1;
  p.tb_floats = p.tb_rows * 8;
  p.prow = p.npad + 1;
  p.smem = 1024 + (size_t)TN_RING * p.slab_bytes + p.w_bytes + (size_t)p.tb_floats * 4 + (size_t)TN_PSLOTS * 128 * p.prow * 4 +
           8 * (2 * TN_RING + 4) + 16;
  if (p.smem > 227 * 1024) return p;
  p.img_floats = ((size_t)p.w_bytes / 4 + p.tb_floats + 3) & ~(size_t)3;
  p.ok = true;
  return p;
}

bool tconv_tapn_supported(const bsde_conv_layer* L, int S, int B, int epi) {
  static const int on = [] { const char* e = getenv("BSDE_TAPN"); return e ? atoi(e) : 1; }();
  if (!on || !L || S < 1 || S > 148 || B < 1) return false;
  if (!(epi == BSDE_EPI_BIAS || epi == BSDE_EPI_EULER || epi == BSDE_EPI_ACCUM)) return false;
  TapnPlan p = tapn_plan(L);
  if (!p.ok) return false;
  if ((long long)B * p.IP + 1024 > (1LL << 24)) return false;                  // fp32-reciprocal position decode
  if ((long long)B * L->hin * L->win * L->cin > 0x7FFFFFFFLL) return false;
  if ((long long)B * L->hout * L->wout * L->cout > 0x7FFFFFFFLL) return false;
  return true;
}

size_t tconv_tapn_img_floats(const bsde_conv_layer* L) {
  TapnPlan p = tapn_plan(L);
  return p.ok ? p.img_floats : 0;
}

int tconv_tapn_repack(const bsde_conv_layer* L, int S, int nit, const float* w, long long w_sample_stride, long long w_iter_stride,
                      int epi, float* img, cudaStream_t st) {
  TapnPlan p = tapn_plan(L);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the taps-on-N path");
  TapnRepackArgs r;
  memset(&r, 0, sizeof(r));
  r.w = w; r.img = img; r.w_sample_stride = w_sample_stride; r.img_sample_stride = (long long)p.img_floats;
  r.w_iter_stride = w_iter_stride; r.img_iter_stride = (long long)p.img_floats * S;
  r.nks = p.nks; r.npad = p.npad; r.cin = L->cin; r.cout = L->cout;
  r.real_base = (L->has_time && L->time_first) ? 1 : 0; r.has_time = L->has_time;
  r.time_row = L->time_first ? 0 : L->cin; r.use_bias = epi <= BSDE_EPI_EULER; r.w_floats = p.w_bytes / 4;
  r.tb_rows = p.tb_rows; r.Hq = p.Hq; r.Wq = p.Wq; r.off = L->off; r.kn = L->kn;
  r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
  tapn_repack_kernel<<<dim3(p.nks + 1, S, nit), 256, 0, st>>>(r);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int tconv_fwd_tapn(const bsde_conv_layer* L, int S, int B, const float* in, const float* w, long long w_sample_stride,
                   float t, int act, int epi, float scale, const float* aux, float* out, void* ws, size_t ws_bytes,
                   cudaStream_t st, const float* prepacked) {
  (void)act;
  TapnPlan p = tapn_plan(L);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the taps-on-N path");
  BSDE_CHECK_ARG(!(epi >= BSDE_EPI_EULER) || aux != nullptr, "epilogue %d needs aux", epi);
  if (!prepacked) {
    const size_t need = p.img_floats * sizeof(float) * S;
    if (!ws || ws_bytes < need) {
      set_error("tconv_fwd(tapn): workspace %zu < %zu bytes", ws_bytes, need);
      return BSDE_ERR_WORKSPACE;
    }
    if (int e = tconv_tapn_repack(L, S, 1, w, w_sample_stride, 0, epi, (float*)ws, st)) return e;
  }
  TapnArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.aux = aux; a.out = out; a.wimg = prepacked ? prepacked : (const float*)ws; a.wimg_sample_stride = (long long)p.img_floats;
  a.S = S; a.B = B;
  a.Q = B * p.IP;
  a.ntiles = (a.Q + 127) / 128;
  a.G = std::max(1, std::min(sm_budget() / S, a.ntiles));
  a.cin = L->cin; a.cout = L->cout; a.npad = p.npad; a.nks = p.nks; a.nhalf = p.nhalf; a.opitch = L->cout;
  a.hout = L->hout; a.wout = L->wout; a.has_time = L->has_time; a.epi = epi;
  a.Hq = p.Hq; a.Wq = p.Wq; a.pitch = p.pitch; a.IP = p.IP; a.NR = p.NR; a.box_bytes = p.box_bytes; a.region = p.region;
  a.slab_bytes = p.slab_bytes; a.w_bytes = p.w_bytes; a.tb_floats = p.tb_floats; a.prow = p.prow;
  a.t = t; a.scale = scale;
  { static const int dbg = getenv("BSDE_TAPN_DBG") ? atoi(getenv("BSDE_TAPN_DBG")) : 0; a.dbg = dbg; }
  for (int tap = 0; tap < 9; ++tap) a.shift[tap] = ((tap / 3) * L->kn + L->off) * p.pitch + ((tap % 3) * L->kn + L->off);
  using KernelFn = void (*)(const TapnArgs, const CUtensorMap);
  static const KernelFn fns[8] = {nullptr, conv_tapn_kernel<1>, conv_tapn_kernel<2>, conv_tapn_kernel<3>, conv_tapn_kernel<4>,
                                  conv_tapn_kernel<5>, conv_tapn_kernel<6>, conv_tapn_kernel<7>};
  const KernelFn fn = fns[L->cout];
  BSDE_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
  const cuuint64_t dims[4] = {(cuuint64_t)L->cin, (cuuint64_t)L->win, (cuuint64_t)L->hin, (cuuint64_t)S * B};
  const cuuint64_t strides[3] = {(cuuint64_t)L->cin * 4, (cuuint64_t)L->win * L->cin * 4, (cuuint64_t)L->hin * L->win * L->cin * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)p.pitch, 1, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return BSDE_ERR_CUDA; }
  fn<<<dim3(S * a.G), dim3(TN_THREADS), p.smem, st>>>(a, tmap);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bsde
