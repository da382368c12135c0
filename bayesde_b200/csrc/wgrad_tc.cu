// K3 weight gradient on tcgen05 tensor cores (TF32 operands, fp32 accumulation in TMEM).
//
//   G[(tap, ka)][nb] = sum over pixels m of  A(m; tap, ka) * Bm(m; nb)
//
// The reduction index is the PIXEL, so both operands are consumed "MN-major" exactly as they lie in
// NHWC memory (channels contiguous per pixel): a k-block is KP=16 pixels, every 128-byte (32-channel)
// slice of a pixel row is one row of a SWIZZLE_128B_BASE32B MN block.  Per k-block the producers gather the nine
// shifted copies of the A rows (zero-filled outside the image) plus the plain Bm rows with cp.async;
// one thread issues M=128 MMAs, each covering four consecutive (tap, 32-channel) blocks, all
// accumulating in TMEM for the whole pixel chunk of the CTA.  Chunks are split across CTAs; partial
// tiles go to the workspace and the fixed-order reduction kernel of tconv_simt.cu finishes (no atomics).
//   direct form  (forward layers with sd == 1): A = layer input gathered by the forward map, Bm = dy.
//   swapped form (lhs-dilated layers, sd == 2): A = dy gathered by the data-gradient map, Bm = input.
// The time-channel rows (t * sum of dy over the pixels where the tap is inside the image) and the bias
// row (sum of dy) are tap-masked column sums of dy; a small CUDA-core kernel adds them to the partials.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace bsde {

__device__ long long* g_wg_dbg = nullptr;  // tools/: accumulated clock64 phases of CTA 0

constexpr int WG_KP = 16;        // pixels per k-block
constexpr int WG_PROD = 256;
constexpr int WG_THREADS = WG_PROD + 32;
constexpr int WG_REGION = WG_KP * 128;  // bytes of one MN block of one k-block

struct WgTcArgs {
  const float* ga;   // gathered operand  [S][B*hA*wA][CA]
  const float* bm;   // plain operand     [S][Mper][CB]
  float* partial;    // [S][splits][rows][cols]
  int S, B, splits, chunk;
  int CA, CB, ksize, hA, wA, hM, wM, sn, kn, off, sd;  // gather geometry; (hM,wM) = pixel loop space
  int rows, cols, Mper;
  int row_stride_tap;   // partial row = tap*row_stride_tap + channel
  int bt, nblk, nmma, npad, nbB, stages, tmem_cols;
  int dbg;  // tools/ only: 2 skip MMA, 4 skip A gathers
};

__device__ __forceinline__ uint32_t wg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wg_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void wg_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void wg_cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void wg_cp4(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ bool wg_map_axis(int o, int k, int sn, int kn, int off, int sd, int size, int& i) {
  int num = o * sn + k * kn + off;
  if (num < 0) return false;
  if (sd == 2) {
    if (num & 1) return false;
    num >>= 1;
  }
  i = num;
  return num < size;
}
// MN-major TF32 operands must use the SWIZZLE_128B_BASE32B layout (verified on hardware with
// tools/umma_probe.cu): rows of 128 B (32 channels of one pixel), the four 32-byte granules of a row
// XOR-permuted with (k & 3), k-atoms of 4 rows = 512 B (SBO), MN blocks `lbo` bytes apart (LBO).
__device__ __forceinline__ uint64_t wg_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

template <bool VA, bool VB>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(WgTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / a.splits, split = blockIdx.x - s * a.splits;
  const int mbeg = split * a.chunk;
  const int mend = min(a.Mper, mbeg + a.chunk);
  const int NKB = (mend - mbeg + WG_KP - 1) / WG_KP;
  const int ntap = a.ksize * a.ksize;

  const uint32_t raw = wg_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int a_bytes = a.nblk * WG_REGION, b_bytes = a.nbB * WG_REGION;
  const int stage_bytes = a_bytes + b_bytes;
  const uint32_t bars = base + a.stages * stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + a.stages * stage_bytes + 8 * (2 * 4 + 1));
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto empty_bar = [&](int st) { return bars + 8u * (4 + st); };
  const uint32_t tmem_full_bar = bars + 8u * 8;

  const float* __restrict__ ga = a.ga + (long long)s * a.B * a.hA * a.wA * a.CA;
  const float* __restrict__ bm = a.bm + (long long)s * a.Mper * a.CB;

  // zero all stages once: padded channels / padded blocks are never written by the gathers
  for (int i = tid; i < a.stages * stage_bytes / 16; i += WG_THREADS)
    reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  if (warp == WG_PROD / 32) {
    if (lane == 0) {
      for (int i = 0; i < a.stages; ++i) { wg_mbar_init(full_bar(i), WG_PROD); wg_mbar_init(empty_bar(i), 1); }
      wg_mbar_init(tmem_full_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wg_smem_u32(tmem_slot)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic zero-fill -> visible to the MMA's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: the prologue above overlapped the predecessor's tail
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < WG_PROD / 32) {
    // =============================== producers ===============================
    // The gather map of this kernel always has sd == 1 (direct form: forward layers with sd == 1;
    // swapped form: the data-gradient map of an lhs-dilated layer), so i = o*sn + (k*kn + off).
    const int ncA = VA ? a.CA / 4 : a.CA;   // copy units per pixel row of A (16 B or 4 B)
    const int ncB = VB ? a.CB / 4 : a.CB;
    constexpr int MAXU = 3;
    // fixed (pixel-in-block, channel) units of this thread, and its running pixel coordinates
    int uA_n = 0, uA_ch[MAXU], uA_img[MAXU], uA_oy[MAXU], uA_ox[MAXU];
    uint32_t uA_dst[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int idx = tid + u * WG_PROD;
      uA_ch[u] = 0; uA_img[u] = 0; uA_oy[u] = 0; uA_ox[u] = 0; uA_dst[u] = 0;
      if (idx < WG_KP * ncA) {
        uA_n = u + 1;
        const int p = idx / ncA, c = idx - p * ncA;
        const int ch = VA ? c * 4 : c;
        const int byte_in_row = (ch & 31) * 4;
        uA_ch[u] = ch;
        uA_dst[u] = (uint32_t)(ch >> 5) * WG_REGION + (uint32_t)p * 128u +
                    (((uint32_t)((byte_in_row >> 5) ^ (p & 3)) << 5) | (uint32_t)(byte_in_row & 31));
        const int m = mbeg + p;
        const int hw = a.hM * a.wM;
        uA_img[u] = m / hw;
        const int r = m - uA_img[u] * hw;
        uA_oy[u] = r / a.wM;
        uA_ox[u] = r - uA_oy[u] * a.wM;
      }
    }
    int uB_n = 0, uB_ch[MAXU], uB_p[MAXU];
    uint32_t uB_dst[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int idx = tid + u * WG_PROD;
      uB_ch[u] = 0; uB_p[u] = 0; uB_dst[u] = 0;
      if (idx < WG_KP * ncB) {
        uB_n = u + 1;
        const int p = idx / ncB, c = idx - p * ncB;
        const int ch = VB ? c * 4 : c;
        const int byte_in_row = (ch & 31) * 4;
        uB_ch[u] = ch; uB_p[u] = p;
        uB_dst[u] = (uint32_t)(ch >> 5) * WG_REGION + (uint32_t)p * 128u +
                    (((uint32_t)((byte_in_row >> 5) ^ (p & 3)) << 5) | (uint32_t)(byte_in_row & 31));
      }
    }
    // per-tap displacements, register resident (ksize is 1 or 3)
    int r_ty[9], r_tx[9], r_toff[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int kh = a.ksize == 3 ? tap / 3 : 0, kw = a.ksize == 3 ? tap % 3 : 0;
      r_ty[tap] = kh * a.kn + a.off;
      r_tx[tap] = kw * a.kn + a.off;
      r_toff[tap] = (r_ty[tap] * a.wA + r_tx[tap]) * a.CA;
    }
    long long t_wait = 0, t_issue = 0;
    const bool dbgt = g_wg_dbg && blockIdx.x == 0 && tid == 0;
#pragma unroll 1
    for (int kb = 0; kb < NKB; ++kb) {
      const int st = kb % a.stages;
      long long c0 = dbgt ? clock64() : 0;
      if (kb >= a.stages) wg_mbar_wait(empty_bar(st), ((kb / a.stages) & 1) ^ 1);
      long long c1 = dbgt ? clock64() : 0;
      t_wait += c1 - c0;
      const uint32_t sA = base + st * stage_bytes;
      const uint32_t sB = sA + a_bytes;
      const int mb = mbeg + kb * WG_KP;
#pragma unroll
      for (int u = 0; u < MAXU; ++u) {
        if (u < uA_n) {
          const bool mval = (uA_img[u] < a.B) && (mb + (int)((uA_dst[u] >> 7) & (WG_KP - 1)) < mend);
          const int by = uA_oy[u] * a.sn, bx = uA_ox[u] * a.sn;
          const int boff = ((uA_img[u] * a.hA + by) * a.wA + bx) * a.CA + uA_ch[u];
          const uint32_t tap_stride = (uint32_t)a.bt * WG_REGION;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (tap < ntap) {
              const bool ok = mval && (unsigned)(by + r_ty[tap]) < (unsigned)a.hA && (unsigned)(bx + r_tx[tap]) < (unsigned)a.wA;
              const float* src = ok ? ga + (boff + r_toff[tap]) : ga;
              const uint32_t dst = sA + uA_dst[u] + (uint32_t)tap * tap_stride;
              if (VA) wg_cp16(dst, src, ok); else wg_cp4(dst, src, ok);
            }
          }
          // advance this unit's pixel by KP for the next k-block
          uA_ox[u] += WG_KP;
          while (uA_ox[u] >= a.wM) {
            uA_ox[u] -= a.wM;
            if (++uA_oy[u] == a.hM) { uA_oy[u] = 0; ++uA_img[u]; }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < MAXU; ++u) {
        if (u < uB_n) {
          const int m = mb + uB_p[u];
          const bool ok = m < mend;
          const float* src = ok ? bm + (long long)m * a.CB + uB_ch[u] : bm;
          if (VB) wg_cp16(sB + uB_dst[u], src, ok); else wg_cp4(sB + uB_dst[u], src, ok);
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(st)) : "memory");
      if (dbgt) t_issue += clock64() - c1;
    }
    if (dbgt) { g_wg_dbg[0] = t_wait; g_wg_dbg[1] = t_issue; g_wg_dbg[2] = NKB; }

    // =============================== epilogue ===============================
    wg_mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* __restrict__ part = a.partial + ((long long)(s * a.splits + split) * a.rows) * a.cols;
    const int q = warp & 3;            // TMEM lane quarter
    const int r = q * 32 + lane;       // row inside an M=128 accumulator
#pragma unroll 1
    for (int i = 0; i < a.nmma; ++i) {
      const int blk = 4 * i + (r >> 5);
      const int tap = blk / a.bt;
      const int ch = (blk - tap * a.bt) * 32 + (r & 31);
      const bool rval = tap < ntap && ch < a.CA;
      const long long prow = (long long)(tap * a.row_stride_tap + ch) * a.cols;
#pragma unroll 1
      for (int n0 = (warp >> 2) * 16; n0 < a.npad; n0 += 32) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(i * a.npad + n0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (rval) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + j < a.CB) part[prow + n0 + j] = __uint_as_float(v[j]);
        }
      }
    }
  } else {
    // =============================== MMA issuer ===============================
    // whole warp, uniform operands (kernel parameters + loop counters), one elected lane issues
    {
      // TF32 x TF32 -> F32, both operands MN-major, M=128, N=npad
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(a.npad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t hi = um::desc_hi(WG_REGION, 512u, 1);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
#pragma unroll 1
      for (int kb = 0; kb < NKB; ++kb) {
        const int st = kb % a.stages;
        wg_mbar_wait(full_bar(st), (kb / a.stages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sA16 = (base + st * stage_bytes) >> 4;
        const uint32_t sB16 = sA16 + ((uint32_t)a_bytes >> 4);
#pragma unroll 1
        for (int i = 0; i < ((a.dbg & 2) ? 0 : a.nmma); ++i) {
#pragma unroll
          for (int j = 0; j < WG_KP / 8; ++j) {
            const uint64_t ad = hi | (uint64_t)((sA16 + (uint32_t)(4 * i) * (WG_REGION >> 4) + (uint32_t)j * 64u) & 0x3FFFu);
            const uint64_t bd = hi | (uint64_t)((sB16 + (uint32_t)j * 64u) & 0x3FFFu);
            um::umma_tf32_elect(tm + (uint32_t)(i * a.npad), ad, bd, idesc, (kb > 0 || j > 0) ? 1u : 0u);
          }
        }
        um::umma_commit_elect(empty_bar(st));
      }
      um::umma_commit_elect(tmem_full_bar);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == WG_PROD / 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
}

// ---- tap-masked column sums of dy over OUTPUT pixels (forward validity): time rows + bias row ----
struct TapSumArgs {
  const float* dy;
  float* partial;       // main partial array
  float* bias_partial;  // swapped form only: [S][splits][cout]
  int S, B, splits, chunk, cout, ksize, hin, win, hout, wout, sn, kn, off, sd;
  int rows, cols, swapped, has_time, Kc, Ktot, cin;
  float t;
};

template <int VW>
__global__ void __launch_bounds__(1024) tap_colsum_kernel(TapSumArgs a) {
  // threads = (pixel lane, column group of VW channels).  Interior pixels (every tap inside the image) add
  // into one accumulator; only border pixels pay the per-tap predicated adds.
  const int split = blockIdx.x, s = blockIdx.y;
  const int Mout = a.B * a.hout * a.wout;
  const int mbeg = split * a.chunk, mend = min(Mout, mbeg + a.chunk);
  const float* __restrict__ dy = a.dy + (long long)s * Mout * a.cout;
  float* __restrict__ part = a.partial + ((long long)(s * a.splits + split) * a.rows) * a.cols;
  __shared__ float red[10][1024];
  __shared__ unsigned s_rowmask[64], s_colmask[64];  // per output row / column: valid kh / kw bits
  for (int i = threadIdx.x; i < a.hout + a.wout; i += blockDim.x) {
    const bool isrow = i < a.hout;
    const int o = isrow ? i : i - a.hout;
    unsigned mk = 0;
    for (int k = 0; k < a.ksize; ++k) {
      int ii;
      if (wg_map_axis(o, k, a.sn, a.kn, a.off, a.sd, isrow ? a.hin : a.win, ii)) mk |= 1u << k;
    }
    if (isrow) s_rowmask[o] = mk; else s_colmask[o] = mk;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched with programmatic dependent launch
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();
  const unsigned fullmask = (1u << a.ksize) - 1u;
  const int ngrp = (a.cout + VW - 1) / VW;      // column groups
  int cp = 1;
  while (cp < ngrp && cp < 64) cp <<= 1;
  const int lanes = 1024 / cp;
  const int c = threadIdx.x & (cp - 1), pl = threadIdx.x / cp;
  const int hw = a.hout * a.wout;
  for (int g0 = 0; g0 < ngrp; g0 += 64) {
    const bool cval = g0 + c < ngrp;
    const int ch = (g0 + c) * VW;
    float acc_all[VW], acc_bt[VW], acc[9][VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      acc_all[e] = 0.f; acc_bt[e] = 0.f;
#pragma unroll
      for (int i = 0; i < 9; ++i) acc[i][e] = 0.f;
    }
    int m = mbeg + pl;
    int img = m / hw;
    int r = m - img * hw;
    int oy = r / a.wout, ox = r - oy * a.wout;
#pragma unroll 2
    for (; m < mend; m += lanes) {
      float v[VW];
      if (VW == 4) {
        const float4 q = cval ? __ldg(reinterpret_cast<const float4*>(dy + (long long)m * a.cout + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[0] = q.x; v[1 % VW] = q.y; v[2 % VW] = q.z; v[3 % VW] = q.w;
      } else {
        v[0] = cval ? __ldg(dy + (long long)m * a.cout + ch) : 0.f;
      }
      const unsigned rm = s_rowmask[oy], cm = s_colmask[ox];
      if (rm == fullmask && cm == fullmask) {
#pragma unroll
        for (int e = 0; e < VW; ++e) acc_all[e] += v[e];
      } else {
#pragma unroll
        for (int e = 0; e < VW; ++e) acc_bt[e] += v[e];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            if (((rm >> kh) & 1) && ((cm >> kw) & 1)) {
#pragma unroll
              for (int e = 0; e < VW; ++e) acc[kh * 3 + kw][e] += v[e];
            }
      }
      ox += lanes;
      while (ox >= a.wout) {
        ox -= a.wout;
        if (++oy == a.hout) oy = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 9; ++i) red[i][threadIdx.x] = acc[i][e] + acc_all[e];
      red[9][threadIdx.x] = acc_all[e] + acc_bt[e];
      __syncthreads();
      for (int idx = threadIdx.x; idx < 10 * cp; idx += blockDim.x) {
        const int i = idx / cp, cc = idx - i * cp;
        const int co = (g0 + cc) * VW + e;
        if (g0 + cc >= ngrp || co >= a.cout) continue;
        float vs = 0.f;
        for (int l = 0; l < lanes; ++l) vs += red[i][l * cp + cc];
        if (i == 9) {
          if (!a.swapped) part[(long long)a.Ktot * a.cols + co] = vs;
          else if (a.bias_partial) a.bias_partial[((long long)s * a.splits + split) * a.cout + co] = vs;
        } else if (a.has_time) {
          // accumulators are indexed kh*3+kw; for ksize 1 only (0,0) exists
          const int kh = i / 3, kw = i - kh * 3;
          if (kh < a.ksize && kw < a.ksize) {
            const int tap = kh * a.ksize + kw;
            if (!a.swapped) part[(long long)(tap * a.Kc + a.Kc - 1) * a.cols + co] = a.t * vs;
            else part[(long long)(tap * a.cout + co) * a.cols + a.cin] = a.t * vs;
          }
        }
      }
    }
    __syncthreads();
  }
}

struct WgTcPlan {
  int swapped, rows, cols, splits, chunk, Mper, CA, CB, bt, nblk, nmma, npad, nbB, stages, tmem_cols, Mout, chunk_o;
  size_t main_floats, bias_floats, smem;
  bool ok;
};

static WgTcPlan wg_plan(const bsde_conv_layer* L, int S, int B) {
  WgTcPlan p;
  memset(&p, 0, sizeof(p));
  const int taps = L->ksize * L->ksize;
  p.swapped = L->sd == 2;
  p.Mout = B * L->hout * L->wout;
  if (!p.swapped) {
    p.CA = L->cin; p.CB = L->cout;
    p.rows = taps * (L->cin + (L->has_time ? 1 : 0)) + 1;
    p.cols = L->cout;
    p.Mper = p.Mout;
  } else {
    p.CA = L->cout; p.CB = L->cin;
    p.rows = taps * L->cout;
    p.cols = L->cin + (L->has_time ? 1 : 0);
    p.Mper = B * L->hin * L->win;
  }
  p.bt = (p.CA + 31) / 32;
  p.nmma = (taps * p.bt + 3) / 4;
  p.nblk = p.nmma * 4;
  p.npad = (p.CB + 15) / 16 * 16;
  p.nbB = (p.CB + 31) / 32;
  const int cols_needed = p.nmma * p.npad;
  p.tmem_cols = cols_needed <= 32 ? 32 : cols_needed <= 64 ? 64 : cols_needed <= 128 ? 128 : cols_needed <= 256 ? 256 : 512;
  const int stage_bytes = (p.nblk + p.nbB) * WG_REGION;
  p.stages = std::min(4, (200 * 1024) / stage_bytes);
  p.smem = 1024 + (size_t)p.stages * stage_bytes + 8 * 9 + 16;
  p.ok = cols_needed <= 512 && p.stages >= 2 && p.npad <= 256 && !(p.swapped && L->sn != 1);
  int want = sm_budget() / S;  // one CTA per SM (TMEM / smem allow a single resident CTA): a single wave
  if (want < 1) want = 1;
  int maxs = (p.Mper + 1023) / 1024;
  p.splits = std::max(1, std::min(want, maxs));
  p.chunk = (p.Mper + p.splits - 1) / p.splits;
  p.chunk = (p.chunk + WG_KP - 1) / WG_KP * WG_KP;
  p.splits = (p.Mper + p.chunk - 1) / p.chunk;
  p.chunk_o = (p.Mout + p.splits - 1) / p.splits;
  p.main_floats = (size_t)S * p.splits * p.rows * p.cols;
  p.bias_floats = p.swapped ? (size_t)S * p.splits * L->cout : 0;
  return p;
}

bool tconv_wgrad_tc_supported(const bsde_conv_layer* L, int S, int B) {
  if (!L || (L->ksize != 1 && L->ksize != 3)) return false;
  return wg_plan(L, S, B).ok;
}

size_t tconv_wgrad_ws_tc(const bsde_conv_layer* L, int S, int B) {
  WgTcPlan p = wg_plan(L, S, B);
  return (p.main_floats + p.bias_floats) * sizeof(float);
}

int tconv_wgrad_tc(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t, float scale,
                   float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes, cudaStream_t st) {
  WgTcPlan p = wg_plan(L, S, B);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the tensor-core wgrad");
  const size_t need = (p.main_floats + p.bias_floats) * sizeof(float);
  if (!ws || ws_bytes < need) {
    set_error("tconv_wgrad(tc): workspace %zu < %zu bytes", ws_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  float* partial = (float*)ws;
  float* bias_partial = p.swapped ? partial + p.main_floats : nullptr;
  const int Kc = L->cin + (L->has_time ? 1 : 0);
  const int Ktot = L->ksize * L->ksize * Kc;

  WgTcArgs a;
  memset(&a, 0, sizeof(a));
  a.partial = partial; a.S = S; a.B = B; a.splits = p.splits; a.chunk = p.chunk;
  a.CA = p.CA; a.CB = p.CB; a.ksize = L->ksize; a.rows = p.rows; a.cols = p.cols; a.Mper = p.Mper;
  a.bt = p.bt; a.nblk = p.nblk; a.nmma = p.nmma; a.npad = p.npad; a.nbB = p.nbB; a.stages = p.stages; a.tmem_cols = p.tmem_cols;
  if (!p.swapped) {
    a.ga = in; a.bm = dy;
    a.hA = L->hin; a.wA = L->win; a.hM = L->hout; a.wM = L->wout;
    a.sn = L->sn; a.kn = L->kn; a.off = L->off; a.sd = L->sd;
    a.row_stride_tap = Kc;
  } else {
    a.ga = dy; a.bm = in;
    a.hA = L->hout; a.wA = L->wout; a.hM = L->hin; a.wM = L->win;
    a.sn = L->sd; a.kn = -L->kn; a.off = -L->off; a.sd = L->sn;
    a.row_stride_tap = L->cout;
  }
  static const int wg_dbg = getenv("BSDE_WG_DBG") ? atoi(getenv("BSDE_WG_DBG")) : 0;
  a.dbg = wg_dbg;
  const bool va = p.CA % 4 == 0, vb = p.CB % 4 == 0;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    BSDE_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    BSDE_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    BSDE_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    BSDE_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  }
#define WG_LAUNCH(VA_, VB_) BSDE_CUDA_OK(launch_pdl(wgrad_tc_kernel<VA_, VB_>, dim3(S * p.splits), dim3(WG_THREADS), p.smem, st, a))
  if (va && vb) WG_LAUNCH(true, true);
  else if (va) WG_LAUNCH(true, false);
  else if (vb) WG_LAUNCH(false, true);
  else WG_LAUNCH(false, false);
#undef WG_LAUNCH
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();

  TapSumArgs ta;
  memset(&ta, 0, sizeof(ta));
  ta.dy = dy; ta.partial = partial; ta.bias_partial = (L->b_off >= 0) ? bias_partial : nullptr;
  ta.S = S; ta.B = B; ta.splits = p.splits; ta.chunk = p.chunk_o; ta.cout = L->cout; ta.ksize = L->ksize;
  ta.hin = L->hin; ta.win = L->win; ta.hout = L->hout; ta.wout = L->wout;
  ta.sn = L->sn; ta.kn = L->kn; ta.off = L->off; ta.sd = L->sd;
  ta.rows = p.rows; ta.cols = p.cols; ta.swapped = p.swapped; ta.has_time = L->has_time; ta.Kc = Kc; ta.Ktot = Ktot;
  ta.cin = L->cin; ta.t = t;
  if (!(wg_dbg & 1))
  if (L->cout % 4 == 0) BSDE_CUDA_OK(launch_pdl(tap_colsum_kernel<4>, dim3(p.splits, S), dim3(1024), 0, st, ta));
  else BSDE_CUDA_OK(launch_pdl(tap_colsum_kernel<1>, dim3(p.splits, S), dim3(1024), 0, st, ta));
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();

  WreduceArgs r;
  memset(&r, 0, sizeof(r));
  r.partial = partial; r.bias_partial = nullptr; r.gw = gw; r.gw_sample_stride = gw_sample_stride;
  r.S = S; r.splits = p.splits; r.rows = p.rows; r.cols = p.cols; r.Kc = Kc; r.Ktot = Ktot;
  r.has_time = L->has_time; r.time_row = L->time_first ? 0 : L->cin;
  r.real_base = (L->has_time && L->time_first) ? 1 : 0; r.swapped = p.swapped; r.cin = L->cin; r.cout = L->cout;
  r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
  r.scale = scale;
  if (p.swapped && L->b_off >= 0) {
    WreduceArgs rb = r;
    rb.partial = nullptr; rb.bias_partial = bias_partial; rb.rows = 0; rb.cols = 1;
    if (int e = launch_wgrad_reduce(rb, (S * L->cout + 255) / 256, st)) return e;
  }
  const long long total = (long long)S * p.rows * p.cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  return launch_wgrad_reduce(r, blocks, st);
}

}  // namespace bsde

extern "C" int bsde_debug_set_wg_stamps(long long* d_buf) {
  return cudaMemcpyToSymbol(bsde::g_wg_dbg, &d_buf, sizeof(d_buf)) == cudaSuccess ? 0 : -3;
}
