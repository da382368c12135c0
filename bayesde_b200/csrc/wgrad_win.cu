// K3 weight gradient, window-staged (tcgen05 kind::tf32, fp32 accumulation in TMEM).
//
//   G[tap][ci][co] = sum over positions q of  in(q + s_tap)[ci] * dY(q)[co]
//
// on the flattened, zero-padded position grid of tconv_win.cu (one pad column per row, one pad row per
// image: a neighbour that would wrap around lands on a pad, which both operands hold as zeros).  The
// reduction index is the POSITION, so both operands are MN-major exactly as NHWC memory has them: one
// position = one 128-byte row per 32-channel block, SWIZZLE_128B_BASE32B.  Every tap uses the SAME staged
// rows; only the start address of the shifted operand's descriptor moves (tools/umma_probe2.cu, test T_C):
//
//   D_e[m][n] += sum_k  U_{up(e)}(q_k)[m] * V_{vp(e)}(q_k + sh(e))[n]          one MMA (M=128, K=8) per k-step
//
//   U = unshifted operand on the M side (128 accumulator rows), V = shifted operand on the N side.
//     forward stride 1 : U = input (+ ones), V = dY shifted by -s_tap;  when cin + 1 <= 32 instead
//                        U = dY, V = input (+ ones) shifted by +s_tap
//     forward stride 2 : U = the four input parity planes (+ ones), V = dY shifted by -s
//     lhs-dilated fwd  : U = input (+ ones), V = the four dY parity classes shifted by -s
//   "ones" is an extra channel that is 1 at valid positions: its accumulator row/column is the tap-masked
//   column sum of dY, i.e. the time-channel row (times t) and, for the always-valid tap, the bias row.
//   When the shifted operand has <= 32 channels its three horizontally adjacent taps are consecutive rows of the
//   same region, so ONE MMA with N = 96 and LBO = 128 B covers them ("packed": 3 MMAs per k-step, not 9).
// CTAs split the positions of a sample (and, when 9 accumulators exceed the 512 TMEM columns, the taps);
// partial tiles go to the workspace in the layout of wgrad_reduce_kernel (tconv_simt.cu), which finishes
// in a fixed order (no atomics: run-to-run deterministic).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

namespace bsde {

using namespace um;

constexpr int WW_PROD_WARPS = 16;
constexpr int WW_PROD = WW_PROD_WARPS * 32;  // warps 0-7 producers (warps 0-3 also run the epilogue)
constexpr int WW_MMA_WARP = WW_PROD_WARPS;
constexpr int WW_THREADS = (WW_PROD_WARPS + 1) * 32;
constexpr int WW_MAXENT = 9;
constexpr int WW_MAXSTAGES = 4;

__device__ long long* g_ww_dbg = nullptr;  // tools/stamp_wwin.py: accumulated clock64 phases of CTA 0
__device__ float ww_one_dev = 1.0f;  // source of the ones channel (copied with cp.async like every other element)

struct WwArgs {
  const float* U;
  const float* V;
  const float* one;
  float* partial;
  int S, B, Hq, Wq, pitch, IP, Q;
  int cu, cv, u_ones, v_ones;     // real channels; append a ones channel at index cu / cv
  int nbU, nbV, NPU, NPV;         // 32-channel blocks, parity planes of each operand
  int hU, wU, hV, wV;             // spatial extent of the tensors behind U and V
  int KS, NPOSw, lead, regU, regV, stage_bytes, v_off, nstages;  // rows per stage, V window rows, region bytes
  int tmaU, tmaV, NRU, NRV, box_bytes;  // TMA-fed operands: NR grid-row boxes {32 ch, pitch px} per (plane, block), box_bytes each
  int SPL, TG, chunk, vgch;       // position splits, V channel groups, positions per split, V channels per group
  int Nmma;                       // N of every MMA
  uint32_t lboB;                  // LBO of the B descriptor: regV (channel blocks) or 128 (packed taps)
  int tmem_cols;
  // per tap group: entries
  int ncnt[2];
  uint32_t aoff[2][WW_MAXENT];    // A start (16-byte units) relative to the stage base: U plane
  uint32_t boff[2][WW_MAXENT];    // B start relative to the stage base: V plane + (lead + shift) rows
  int tapj[2][WW_MAXENT][3];      // weight tap of N-block j (packed) or of the entry ([0])
  int bias_row[2][WW_MAXENT][3];  // >= 0: this (entry, block)'s ones sum is bias partial row Ktot + bias_row
  int packed, u_is_input;
  int vpar;                       // V is held as parity planes (lhs-dilated forward layers)
  int tapsplit;                   // the two CTA groups of a sample split the TAPS (by parity class of V), not the V channels
  int vpl[2][4];                  // tapsplit: parity class held by V plane p of group g
  int cin, cout, Kc, Ktot, has_time, rows, cols;
  float t;
  int dbg;  // tools/ only (BSDE_WW_DBG): 1 skip copies, 2 skip MMAs, 4 skip the source tables
};

// Per-thread position state of one staged row: every stage advances by KS positions, so (img, y, x) are updated
// incrementally instead of being decoded (the decode cost ~1500 cycles per stage: tools/stamp_wwin.py).
struct WwPos { int q, img, y, x; };
struct WwGeo { int B, Hq, Wq, pitch, IP, adv_img, adv_y, adv_x; };
__device__ __forceinline__ void ww_pos_init(WwPos& s, int q, const WwGeo& g) {
  s.q = q;
  int qq = q, img = 0;
  if (qq < 0) { qq += g.IP; img = -1; }  // rows before the first position: |q| <= lead < IP
  img += qq / g.IP;
  const int r = qq % g.IP;
  s.img = img; s.y = r / g.pitch; s.x = r % g.pitch;
}
__device__ __forceinline__ void ww_pos_advance(WwPos& s, int KS, const WwGeo& g) {
  s.q += KS;
  s.img += g.adv_img;
  s.x += g.adv_x;
  if (s.x >= g.pitch) { s.x -= g.pitch; ++s.y; }
  s.y += g.adv_y;
  if (s.y > g.Hq) { s.y -= g.Hq + 1; ++s.img; }
}
__device__ __forceinline__ bool ww_pos_valid(const WwPos& s, const WwGeo& g) {
  return (unsigned)s.img < (unsigned)g.B && s.y < g.Hq && s.x < g.Wq;
}

// position -> pixel coordinates on the position grid (returns false on pads / out of range)
__device__ __forceinline__ bool ww_decode(int q, int Q, int IP, int pitch, int Hq, int Wq, float rIP, float rpitch, int& img, int& y, int& x) {
  if (q < 0 || q >= Q) return false;
  img = (int)((float)q * rIP);
  int r = q - img * IP;
  if (r < 0) { --img; r += IP; } else if (r >= IP) { ++img; r -= IP; }
  y = (int)((float)r * rpitch);
  x = r - y * pitch;
  if (x < 0) { --y; x += pitch; } else if (x >= pitch) { ++y; x -= pitch; }
  return y < Hq && x < Wq;
}
__device__ __forceinline__ int ww_plane_src(bool ok, int img, int y, int x, bool par, int plane, int h, int w, int c) {
  if (!ok) return -1;
  if (par) { y = 2 * y + (plane >> 1); x = 2 * x + (plane & 1); }
  if (y >= h || x >= w) return -1;
  return ((img * h + y) * w + x) * c;
}

// position -> source pixel offset (floats) inside one sample, or -1
__device__ __forceinline__ int ww_src(int q, int Q, int IP, int pitch, int Hq, int Wq, float rIP, float rpitch, int np, int plane,
                                      int h, int w, int c) {
  if (q < 0 || q >= Q) return -1;
  int img = (int)((float)q * rIP);
  int r = q - img * IP;
  if (r < 0) { --img; r += IP; } else if (r >= IP) { ++img; r -= IP; }
  int y = (int)((float)r * rpitch);
  int x = r - y * pitch;
  if (x < 0) { --y; x += pitch; } else if (x >= pitch) { ++y; x -= pitch; }
  if (y >= Hq || x >= Wq) return -1;
  if (np == 4) { y = 2 * y + (plane >> 1); x = 2 * x + (plane & 1); }
  if (y >= h || x >= w) return -1;
  return ((img * h + y) * w + x) * c;
}

// Copy `rows` rows of one operand (np planes x NB 32-channel blocks) into its SWIZZLE_128B_BASE32B regions, 16 bytes
// per lane, 8 lanes per 128-byte row (whole lines on both sides).  Rows advance by 32 per pass, so the swizzled
// granule slot ((c >> 1) ^ (r & 3)) is a per-thread constant; the table is read once per row for all blocks.
template <int NB>
__device__ __forceinline__ void ww_copy_vec(const float* __restrict__ src, const float* one, uint32_t sb, const int* tab, int np,
                                            int reg, int rows, int cch, int ones, int ch0, int tid) {
  const int c = tid & 7, r0 = tid >> 3;
  const uint32_t swz = (uint32_t)((((c >> 1) ^ (r0 & 3)) << 5) | ((c & 1) << 4));
  int kind[NB];  // 0: beyond the channels, 1: data chunk, 2: the ones channel starts this chunk
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int ch = ch0 + b * 32 + c * 4;
    kind[b] = ch < cch ? 1 : ((ones && cch >= ch && cch < ch + 4) ? 2 : 0);
  }
  const float* __restrict__ sc = src + ch0 + c * 4;
#pragma unroll 1
  for (int p = 0; p < np; ++p) {
    const int* tp = tab + p * rows;
    const uint32_t rg = sb + (uint32_t)(p * NB * reg) + swz;
#pragma unroll 4
    for (int r = r0; r < rows; r += WW_PROD / 8) {
      const int so = tp[r];
      const bool ok = so >= 0;
      const float* g = sc + (ok ? so : 0);
      const uint32_t d = rg + r * 128;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (kind[b] == 1) cp_async16(d + b * reg, ok ? g + b * 32 : src, ok);
        else if (kind[b] == 2) cp_async4(d + b * reg, one, ok);
      }
    }
  }
}

// Channel counts that are not a multiple of 4 (the 5-channel tensors; one block): 4 bytes per lane with
// (row, channel) flattened over ALL lanes.
__device__ __forceinline__ void ww_copy_rows(const float* __restrict__ src, const float* one, uint32_t sb, const int* tab, int np,
                                             int nb, int reg, int rows, int cch, int ones, int ch0, int tid) {
  if ((cch & 3) == 0) {
    switch (nb) {
      case 1: ww_copy_vec<1>(src, one, sb, tab, np, reg, rows, cch, ones, ch0, tid); break;
      case 2: ww_copy_vec<2>(src, one, sb, tab, np, reg, rows, cch, ones, ch0, tid); break;
      case 3: ww_copy_vec<3>(src, one, sb, tab, np, reg, rows, cch, ones, ch0, tid); break;
      default: ww_copy_vec<4>(src, one, sb, tab, np, reg, rows, cch, ones, ch0, tid); break;
    }
  } else {
    const int ne = cch + ones;
    const float rne = 1.0f / (float)ne;
#pragma unroll 1
    for (int p = 0; p < np; ++p) {
      const uint32_t rg = sb + (uint32_t)(p * nb * reg);
#pragma unroll 2
      for (int idx = tid; idx < rows * ne; idx += WW_PROD) {
        int r = (int)((float)idx * rne);
        int e = idx - r * ne;
        if (e < 0) { --r; e += ne; } else if (e >= ne) { ++r; e -= ne; }
        const int so = tab[p * rows + r];
        const uint32_t d = rg + r * 128 + ((((e >> 3) ^ (r & 3)) << 5) | ((e & 7) << 2));
        cp_async4(d, e < cch ? (so >= 0 ? src + so + e : src) : one, so >= 0);
      }
    }
  }
}

// MMA issue loop of one CTA; NE > 0: entry count known at compile time (descriptor offsets stay in uniform registers)
template <int NE>
__device__ __forceinline__ void ww_issue(const WwArgs& a, int tg, int nst, int qbeg, uint32_t base, uint32_t bars, uint32_t tm) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(a.Nmma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t a_hi = desc_hi((uint32_t)a.regU, 512u, 1);
  const uint64_t b_hi = desc_hi(a.lboB, 512u, 1);
  const int ncnt = NE > 0 ? NE : a.ncnt[tg];
  uint32_t ao[NE > 0 ? NE : 1], bo[NE > 0 ? NE : 1];
  if (NE > 0) {
#pragma unroll
    for (int e = 0; e < NE; ++e) { ao[e] = a.aoff[tg][e]; bo[e] = a.boff[tg][e]; }
  }
  const uint32_t nm = (uint32_t)a.Nmma;
  const bool dbgm = g_ww_dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0;
  long long mw = 0, mi = 0;
#pragma unroll 1
  for (int it = 0; it < nst; ++it) {
    long long m0 = dbgm ? clock64() : 0;
    const int st = it % a.nstages;
    mbar_wait(bars + 8u * st, (it / a.nstages) & 1);
    long long m1 = dbgm ? clock64() : 0;
    fence_proxy_async();
    tc_fence_after();
    const uint32_t sb16 = (base + st * a.stage_bytes) >> 4;
    // TMA-fed regions start at a grid-row boundary: the stage's first position sits oU / oV rows into them
    const int q0 = qbeg + it * a.KS;
    const uint32_t oU = a.tmaU ? (uint32_t)(q0 % a.pitch) * 8u : 0u;
    const int nv = q0 - a.lead;
    const uint32_t oV = a.tmaV ? (uint32_t)(nv - ((nv + 2 * a.pitch) / a.pitch - 2) * a.pitch) * 8u : 0u;
#pragma unroll 1
    for (int ks = 0; ks < ((a.dbg & 2) ? 0 : a.KS / 8); ++ks) {
      const uint32_t kofs = sb16 + (uint32_t)ks * 64u;  // 8 rows x 128 B
      const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
      if (NE > 0) {
#pragma unroll
        for (int e = 0; e < NE; ++e)
          umma_tf32_elect(tm + (uint32_t)e * nm, a_hi | (uint64_t)((kofs + oU + ao[e]) & 0x3FFFu), b_hi | (uint64_t)((kofs + oV + bo[e]) & 0x3FFFu), idesc, acc);
      } else {
#pragma unroll 1
        for (int e = 0; e < ncnt; ++e)
          umma_tf32_elect(tm + (uint32_t)e * nm, a_hi | (uint64_t)((kofs + oU + a.aoff[tg][e]) & 0x3FFFu),
                          b_hi | (uint64_t)((kofs + oV + a.boff[tg][e]) & 0x3FFFu), idesc, acc);
      }
    }
    umma_commit_elect(bars + 8u * (4 + st));
    if (dbgm) { mw += m1 - m0; mi += clock64() - m1; }
  }
  if (dbgm) { g_ww_dbg[8] = mw; g_ww_dbg[9] = mi; }
  umma_commit_elect(bars + 8u * 8);
}

__global__ void __launch_bounds__(WW_THREADS, 1) wgrad_win_kernel(const __grid_constant__ WwArgs a, const __grid_constant__ CUtensorMap tmU,
                                                                  const __grid_constant__ CUtensorMap tmV) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / (a.TG * a.SPL);
  const int rem = blockIdx.x - s * a.TG * a.SPL;
  const int vg = rem / a.SPL, split = rem - vg * a.SPL;  // this CTA's group of V channels (accumulator columns)
  const int tg = a.tapsplit ? vg : 0;   // entry table of this CTA
  const int qbeg = split * a.chunk, qend = min(a.Q, qbeg + a.chunk);
  const int nst = qend > qbeg ? (qend - qbeg + a.KS - 1) / a.KS : 0;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int ring_bytes = a.nstages * a.stage_bytes;
  const int tab_stride = a.NPU * a.KS + a.NPV * a.NPOSw;
  int* s_tabU = reinterpret_cast<int*>(gen + ring_bytes);                      // [2][NPU][KS]
  int* s_tabV = s_tabU + a.NPU * a.KS;                                          // [2][NPV][NPOSw] (interleaved with U per buffer)
  const uint32_t bars = base + ring_bytes + 2 * (a.NPU * a.KS + a.NPV * a.NPOSw) * 4;  // full[4], empty[4], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + ring_bytes + 2 * (a.NPU * a.KS + a.NPV * a.NPOSw) * 4 + 8 * 9);
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto empty_bar = [&](int st) { return bars + 8u * (4 + st); };
  const uint32_t done_bar = bars + 8u * 8;

  // zero both stages once: padded channels / unused blocks are never written by the producers
  for (int i = tid; i < ring_bytes >> 4; i += WW_THREADS) reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  if (warp == WW_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < a.nstages; ++i) { mbar_init(full_bar(i), WW_PROD + ((a.tmaU | a.tmaV) ? 1 : 0)); mbar_init(empty_bar(i), 1); }
      mbar_init(done_bar, 1);
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
  // programmatic dependent launch: everything above (smem zeroing, barriers, TMEM) overlapped the predecessor's tail
  pdl_wait();
  pdl_launch_dependents();

  if (warp < WW_PROD_WARPS) {
    // =============================== producers ===============================
    const float* __restrict__ U = a.U + (long long)s * a.B * a.hU * a.wU * a.cu;
    const float* __restrict__ V = a.V + (long long)s * a.B * a.hV * a.wV * a.cv;
    const float rIP = 1.0f / (float)a.IP, rpitch = 1.0f / (float)a.pitch;
    // thread t owns U row t (t < KS) and V window row t (t < NPOSw) of every stage
    WwGeo geo;
    geo.B = a.B; geo.Hq = a.Hq; geo.Wq = a.Wq; geo.pitch = a.pitch; geo.IP = a.IP;
    geo.adv_img = a.KS / a.IP;
    geo.adv_y = (a.KS % a.IP) / a.pitch;
    geo.adv_x = (a.KS % a.IP) % a.pitch;
    WwPos pu, pv;
    ww_pos_init(pu, qbeg + tid, geo);
    ww_pos_init(pv, qbeg - a.lead + tid, geo);
    const bool dbgp = g_ww_dbg && blockIdx.x == 0 && tid == 0;
    long long ph[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int it = 0; it < nst; ++it) {
      long long c0 = dbgp ? clock64() : 0;
      const int st = it % a.nstages;
      // source tables of this stage (shared by the producer threads; double buffered -> one barrier per stage)
      int* tU = s_tabU + (it & 1) * tab_stride;
      int* tV = s_tabV + (it & 1) * tab_stride;
      if (!(a.dbg & 4)) {
        if (!a.tmaU && tid < a.KS) {
          const bool ok = pu.q < qend && ww_pos_valid(pu, geo);
          for (int p = 0; p < a.NPU; ++p) tU[p * a.KS + tid] = ww_plane_src(ok, pu.img, pu.y, pu.x, a.NPU == 4, p, a.hU, a.wU, a.cu);
        }
        if (!a.tmaV && tid < a.NPOSw) {
          const bool ok = ww_pos_valid(pv, geo);
          for (int p = 0; p < a.NPV; ++p)
            tV[p * a.NPOSw + tid] = ww_plane_src(ok, pv.img, pv.y, pv.x, a.vpar != 0, a.tapsplit ? a.vpl[tg][p] : p, a.hV, a.wV, a.cv);
        }
        ww_pos_advance(pu, a.KS, geo);
        ww_pos_advance(pv, a.KS, geo);
      }
      long long c1 = dbgp ? clock64() : 0;
      if (!(a.tmaU && a.tmaV)) asm volatile("bar.sync 1, %0;" ::"n"(WW_PROD) : "memory");
      long long c2 = dbgp ? clock64() : 0;
      if (it >= a.nstages) mbar_wait(empty_bar(st), ((it / a.nstages) & 1) ^ 1);
      long long c3 = dbgp ? clock64() : 0;
      const uint32_t sb = base + st * a.stage_bytes;
      // ---- U rows: [plane][block][KS rows][128 B], granule g of row r at g ^ (r & 3); then the V window rows
      if (!(a.dbg & 1)) {
        if (a.tmaU | a.tmaV) {
          // ---- TMA-fed operands: one box {32 channels, pitch pixels} per (plane, block, grid row); the pixel past the
          // row end / the row past the image end / images past the sample end are out of bounds -> zero fill = the pads.
          // (tools/tma_probe.cu: SWIZZLE_128B_ATOM_32B == the BASE32B operand layout, absolute-address swizzle, and
          // elementStrides = 2 yields the compact parity-plane rows.)
          const int q0 = qbeg + it * a.KS;
          const int rows_img = a.Hq + 1;
          const int gU = q0 / a.pitch;
          const int gV = (q0 - a.lead + 2 * a.pitch) / a.pitch - 2;   // floor((q0 - lead) / pitch), q0 - lead >= -(pitch + 1)
          const int nbUr = a.cu >> 5;                                   // real 32-channel blocks of U (a ones block is not TMA-fed)
          const int nbxU = a.tmaU ? a.NPU * nbUr * a.NRU : 0;
          const int nbxV = a.tmaV ? a.NPV * a.nbV * a.NRV : 0;
          if (warp == 0) {
            if (lane == 0)
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_bar(st)), "r"((nbxU + nbxV) * a.box_bytes) : "memory");
            __syncwarp();
#pragma unroll 1
            for (int idx = lane; idx < nbxU + nbxV; idx += 32) {
              const bool isV = idx >= nbxU;
              const int j = isV ? idx - nbxU : idx;
              const int NR = isV ? a.NRV : a.NRU, nb = isV ? a.nbV : nbUr, np = isV ? a.NPV : a.NPU;
              const int r = j % NR, tt = j / NR;
              const int b = tt % nb, pl = tt / nb;
              const int gg = (isV ? gV : gU) + r;
              const int img = (gg + 2 * rows_img) / rows_img - 2;        // floor(gg / rows_img), gg >= -2
              const int y = gg - img * rows_img;
              const int yy = np == 4 ? 2 * y + (pl >> 1) : y, x0 = np == 4 ? (pl & 1) : 0;
              const int c0 = (isV ? vg * a.vgch : 0) + b * 32;
              const uint32_t dst = sb + (isV ? (uint32_t)a.v_off + (uint32_t)((pl * a.nbV + b) * a.regV) : (uint32_t)((pl * a.nbU + b) * a.regU)) +
                                   (uint32_t)(r * a.box_bytes);
              const CUtensorMap* tm = isV ? &tmV : &tmU;
              asm volatile(
                  "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                  "l"(tm), "r"(full_bar(st)), "r"(c0), "r"(x0), "r"(yy), "r"(img), "r"(s)
                  : "memory");
            }
          }
          if (a.tmaU && a.u_ones && tid < a.NRU * a.pitch) {
            // the ones channel of a TMA-fed U: one element per region row (region rows are whole grid rows)
            const int ry = tid / a.pitch, x = tid - ry * a.pitch;
            const int gg = gU + ry;
            const int img = gg / rows_img, y = gg - img * rows_img;
            const bool pv = img < a.B && y < a.Hq && x < a.Wq;
            const int e = a.cu & 31;
            for (int pl = 0; pl < a.NPU; ++pl) {
              int yy = y, xx = x;
              if (a.NPU == 4) { yy = 2 * y + (pl >> 1); xx = 2 * x + (pl & 1); }
              const bool ok = pv && yy < a.hU && xx < a.wU;
              const uint32_t d = sb + (uint32_t)((pl * a.nbU + (a.cu >> 5)) * a.regU) + tid * 128 + ((((e >> 3) ^ (tid & 3)) << 5) | ((e & 7) << 2));
              cp_async4(d, a.one, ok);
            }
          }
        }
        if (!a.tmaU) ww_copy_rows(U, a.one, sb, tU, a.NPU, a.nbU, a.regU, a.KS, a.cu, a.u_ones, 0, tid);
        if (!a.tmaV) ww_copy_rows(V, a.one, sb + (uint32_t)a.v_off, tV, a.NPV, a.nbV, a.regV, a.NPOSw, a.cv, a.v_ones, vg * a.vgch, tid);
      }
      if ((a.dbg & 1) && (a.tmaU | a.tmaV) && tid == 0) mbar_arrive(full_bar(st));  // stands in for the expect_tx arrival
      long long c4 = dbgp ? clock64() : 0;
      cp_async_arrive_noinc(full_bar(st));
      if (dbgp) { ph[0] += c1 - c0; ph[1] += c2 - c1; ph[2] += c3 - c2; ph[3] += c4 - c3; ph[4] += clock64() - c4; ph[5] += 1; }
    }
    if (dbgp) { for (int i = 0; i < 6; ++i) g_ww_dbg[i] = ph[i]; }

    // =============================== epilogue (warps 0-3) ===============================
    if (warp < 4) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      float* __restrict__ part = a.partial + ((long long)(s * a.SPL + split) * a.rows) * a.cols;
      const int m = warp * 32 + lane;  // accumulator row = U channel
#pragma unroll 1
      for (int e = 0; e < a.ncnt[tg]; ++e) {
#pragma unroll 1
        for (int n0 = 0; n0 < a.Nmma; n0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(e * a.Nmma + n0), v);
          if (nst == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;  // this split had no positions: its tile is all zeros
          }
          const int blk = a.packed ? n0 >> 5 : 0;
          const int tap = a.tapj[tg][e][blk];
          const int brow = a.bias_row[tg][e][blk];
          if (tap < 0) continue;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = a.packed ? ((n0 & 31) + j) : (vg * a.vgch + n0 + j);  // V channel
            const float val = __uint_as_float(v[j]);
            if (a.u_is_input) {
              // m = input channel (cin = ones), n = output channel
              if (n >= a.cout) continue;
              if (m < a.cin) {
                part[(long long)(tap * a.Kc + m) * a.cols + n] = val;
              } else if (m == a.cin) {
                if (a.has_time) part[(long long)(tap * a.Kc + a.cin) * a.cols + n] = a.t * val;
                if (brow >= 0) part[(long long)(a.Ktot + brow) * a.cols + n] = val;
              }
            } else {
              // m = output channel, n = input channel (cin = ones)
              if (m >= a.cout) continue;
              if (n < a.cin) {
                part[(long long)(tap * a.Kc + n) * a.cols + m] = val;
              } else if (n == a.cin) {
                if (a.has_time) part[(long long)(tap * a.Kc + a.cin) * a.cols + m] = a.t * val;
                if (brow >= 0) part[(long long)(a.Ktot + brow) * a.cols + m] = val;
              }
            }
          }
        }
      }
    }
  } else {
    // =============================== MMA issuer ===============================
    // whole warp, uniform operands, one elected lane issues
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int ncnt = a.ncnt[tg];
    if (ncnt == 9) ww_issue<9>(a, tg, nst, qbeg, base, bars, tm);
    else if (ncnt == 5) ww_issue<5>(a, tg, nst, qbeg, base, bars, tm);
    else if (ncnt == 4) ww_issue<4>(a, tg, nst, qbeg, base, bars, tm);
    else if (ncnt == 3) ww_issue<3>(a, tg, nst, qbeg, base, bars, tm);
    else ww_issue<0>(a, tg, nst, qbeg, base, bars, tm);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WW_MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
}

// ---- host side ----------------------------------------------------------------------------------------------
struct WwPlan {
  bool ok;
  WwArgs a;
  size_t smem, part_floats;
  int nbias;
};

static int ww_pow2_cols(int c) { int p = 32; while (p < c) p <<= 1; return p; }

static bool ww_enabled() {
  static const int on = [] { const char* e = getenv("BSDE_WGWIN"); return e ? atoi(e) : 1; }();
  return on != 0;
}

static WwPlan ww_plan(const bsde_conv_layer* L, int S, int B) {
  WwPlan p;
  memset(&p, 0, sizeof(p));
  WwArgs& a = p.a;
  if (!ww_enabled() || !L || L->ksize != 3 || S < 1 || B < 1 || S > 74) return p;
  int mode;
  if (L->sd == 1 && L->sn == 1) mode = 1;
  else if (L->sd == 1 && L->sn == 2) mode = 2;
  else if (L->sd == 2 && L->sn == 1) mode = 3;
  else return p;
  if (mode == 1 && (L->hin != L->hout || L->win != L->wout)) return p;
  if (mode == 2 && ((L->hin + 1) / 2 != L->hout || (L->win + 1) / 2 != L->wout)) return p;
  if (mode == 3 && (L->hout != 2 * L->hin || L->wout != 2 * L->win)) return p;
  if (mode == 2 && L->hout > 4) return p;  // measured slower than wgrad_tc.cu: the big operand (4 parity planes x 3 blocks)
                                           // sits on the U side, which every V-channel group reloads
  { static const int m3 = getenv("BSDE_WW_M3") ? atoi(getenv("BSDE_WW_M3")) : 1; if (mode == 3 && L->hin > 8 && !m3) return p; }  // 16x16 grids: 234 us vs 256 us (wgrad_tc.cu)
  a.Hq = mode == 3 ? L->hin : L->hout;
  a.Wq = mode == 3 ? L->win : L->wout;
  a.pitch = a.Wq + 1;
  a.IP = (a.Hq + 1) * a.pitch;
  a.S = S; a.B = B;
  a.Q = B * a.IP;
  if ((long long)B * a.IP + 4096 > (1LL << 24)) return p;
  if ((long long)B * L->hin * L->win * L->cin > 0x7FFFFFFFLL || (long long)B * L->hout * L->wout * L->cout > 0x7FFFFFFFLL) return p;
  a.cin = L->cin; a.cout = L->cout; a.has_time = L->has_time;
  a.Kc = L->cin + (L->has_time ? 1 : 0);
  a.Ktot = 9 * a.Kc;
  a.cols = L->cout;

  // taps: displacement in position space, operand planes
  struct Tap { int tap, kh, kw, sy, sx, plane, cls; } taps[9];
  int nt = 0;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int th = kh * L->kn + L->off, tw = kw * L->kn + L->off;
      Tap& t = taps[nt];
      t.tap = kh * 3 + kw; t.kh = kh; t.kw = kw; t.plane = 0; t.cls = 0;
      if (mode == 1) { t.sy = th; t.sx = tw; }
      else if (mode == 2) { const int py = th & 1, px = tw & 1; t.sy = (th - py) / 2; t.sx = (tw - px) / 2; t.plane = py * 2 + px; }
      else {
        // the (unique) output parity class of this tap: (po + t) even
        const int poh = th & 1, pow_ = tw & 1;
        t.cls = poh * 2 + pow_; t.sy = (poh + th) / 2; t.sx = (pow_ + tw) / 2;
      }
      if (abs(t.sy) > 1 || abs(t.sx) > 1) return p;
      ++nt;
    }

  // operand roles
  const bool v_is_input = mode == 1 && L->cin + 1 <= 32;  // packed input taps on the N side
  a.u_is_input = !v_is_input;
  if (a.u_is_input) {
    a.cu = L->cin; a.u_ones = 1; a.cv = L->cout; a.v_ones = 0;
    a.hU = L->hin; a.wU = L->win; a.hV = L->hout; a.wV = L->wout;
    a.NPU = mode == 2 ? 4 : 1; a.NPV = mode == 3 ? 4 : 1;
  } else {
    a.cu = L->cout; a.u_ones = 0; a.cv = L->cin; a.v_ones = 1;
    a.hU = L->hout; a.wU = L->wout; a.hV = L->hin; a.wV = L->win;
    a.NPU = 1; a.NPV = 1;
  }
  if (a.cu + a.u_ones > 128) return p;
  // the 4-byte copy path handles one 32-channel block only
  if (((a.cu & 3) && a.cu + a.u_ones > 32) || ((a.cv & 3) && a.cv + a.v_ones > 32)) return p;
  a.nbU = (a.cu + a.u_ones + 31) / 32;
  a.nbV = (a.cv + a.v_ones + 31) / 32;
  a.packed = (mode == 1 && a.cv + a.v_ones <= 32) ? 1 : 0;
  a.Nmma = a.packed ? 96 : (a.cv + a.v_ones + 15) / 16 * 16;
  if (a.Nmma > 256) return p;

  // entries: (U plane, V plane, shift of V)
  struct Ent { int up, vp, sh, tapj[3], brow[3]; } ents[9];
  int ne = 0, mins = 0, maxs = 0;
  p.nbias = mode == 3 ? 4 : 1;
  bool have_bias[4] = {false, false, false, false};
  if (a.packed) {
    for (int kh = 0; kh < 3; ++kh) {
      // the three taps of this kernel row: consecutive positions; the MMA's N-block j reads shift base + j
      Ent& e = ents[ne];
      e.up = 0; e.vp = 0;
      int base_sh = 1 << 30;
      for (int i = 0; i < 9; ++i)
        if (taps[i].kh == kh) {
          const int sh = (taps[i].sy * a.pitch + taps[i].sx) * (a.u_is_input ? -1 : 1);
          base_sh = std::min(base_sh, sh);
        }
      e.sh = base_sh;
      for (int j = 0; j < 3; ++j) { e.tapj[j] = -1; e.brow[j] = -1; }
      for (int i = 0; i < 9; ++i)
        if (taps[i].kh == kh) {
          const int sh = (taps[i].sy * a.pitch + taps[i].sx) * (a.u_is_input ? -1 : 1);
          const int j = sh - base_sh;
          if (j < 0 || j > 2) return p;
          e.tapj[j] = taps[i].tap;
          if (taps[i].sy == 0 && taps[i].sx == 0) { e.brow[j] = 0; have_bias[0] = true; }
        }
      mins = std::min(mins, e.sh); maxs = std::max(maxs, e.sh + 2);
      ++ne;
    }
  } else {
    for (int i = 0; i < 9; ++i) {
      Ent& e = ents[ne];
      e.up = mode == 2 ? taps[i].plane : 0;
      e.vp = mode == 3 ? taps[i].cls : 0;
      e.sh = -(taps[i].sy * a.pitch + taps[i].sx);
      e.tapj[0] = taps[i].tap; e.tapj[1] = e.tapj[2] = -1;
      e.brow[0] = e.brow[1] = e.brow[2] = -1;
      // bias = sum of dY over all positions: the zero-displacement tap of each class reads an always-valid source
      const bool always = taps[i].sy == 0 && taps[i].sx == 0 && (mode != 2 || taps[i].plane == 0);
      const int bc = mode == 3 ? taps[i].cls : 0;
      if (always && !have_bias[bc]) { e.brow[0] = bc; have_bias[bc] = true; }
      mins = std::min(mins, e.sh); maxs = std::max(maxs, e.sh);
      ++ne;
    }
  }
  for (int b = 0; b < p.nbias; ++b)
    if (!have_bias[b]) return p;
  a.rows = a.Ktot + p.nbias;
  a.lead = -mins;
  const int trail = maxs;

  // the nine accumulators of a CTA must fit the 512 TMEM columns: otherwise split the V channels into groups of one
  // 32-channel block (each CTA then loads all of U but only its block of V)
  a.TG = 1;
  a.vgch = 32 * a.nbV;
  a.vpar = a.NPV == 4;
  int egrp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // CTA group of each entry
  int per = ne;
  if (ne * a.Nmma > 512) {
    if (a.packed) return p;
    // lhs-dilated layers (V = four parity classes of dY): the two CTA groups split the TAPS by class -- {(1,1), (0,0)} = 4 + 1
    // taps and {(0,1), (1,0)} = 2 + 2 taps -- and keep the full N: the same bytes per CTA as a split of the V channels (two
    // of the four class planes, all channels), but 5 / 4 MMAs with N = 64 per k-step instead of 9 with N = 32 (the A operand
    // is read once per MMA whatever N is: 240 / 192 instead of 360 shared-memory cycles per k-step).  BSDE_WW_TAPSPLIT=0: channels.
    static const int tsp = getenv("BSDE_WW_TAPSPLIT") ? atoi(getenv("BSDE_WW_TAPSPLIT")) : 1;
    if (tsp && mode == 3 && a.u_is_input && 5 * a.Nmma <= 512) {
      a.tapsplit = 1;
      a.TG = 2;
      a.NPV = 2;
      a.vgch = 0;
      a.vpl[0][0] = 3; a.vpl[0][1] = 0; a.vpl[1][0] = 1; a.vpl[1][1] = 2;
      per = 0;
      int cnt[2] = {0, 0};
      for (int i = 0; i < ne; ++i) {
        const int cls = ents[i].vp;
        egrp[i] = (cls == 3 || cls == 0) ? 0 : 1;
        ents[i].vp = (cls == 3 || cls == 1) ? 0 : 1;   // local plane of the class inside its group
        per = std::max(per, ++cnt[egrp[i]]);
      }
    } else if (tsp && mode == 1 && ne == 9 && 5 * a.Nmma <= 512) {
      // stride-1 layers with a wide dY (64 / 80 channels): taps 0-4 (with the centre tap and its bias row) and taps 5-8, full N.
      // Both groups stage the same rows; a split of the V channels into 32-channel groups would stage U once per group as well,
      // issue 9 MMAs with N = 32 per k-step, and need THREE groups for 80 channels (216 CTAs for 72 pairs: one and a half waves)
      a.tapsplit = 1;
      a.TG = 2;
      a.vgch = 0;
      a.vpl[0][0] = a.vpl[1][0] = 0;
      for (int i = 0; i < ne; ++i) egrp[i] = i >= 5 ? 1 : 0;
      per = 5;
    } else {
      a.TG = a.nbV;
      a.nbV = 1;
      a.vgch = 32;
      a.Nmma = 32;
    }
  }
  a.tmem_cols = ww_pow2_cols(per * a.Nmma);
  if (a.tmem_cols > 512) return p;

  // stage size and ring depth: prefer >= 3 stages in flight (latency hiding), then the largest KS
  const size_t cap = 227 * 1024 - 2048;
  a.KS = 0;
  // TMA-fed operands are exact (GPU parity green) but measured 3-35 % slower than the LDGSTS producers on every layer
  // (profiles/r01_summary.md): opt-in
  static const int tma_on = getenv("BSDE_WW_TMA") ? atoi(getenv("BSDE_WW_TMA")) : 0;
  a.tmaU = tma_on && (a.cu % 32) == 0 && 2 * a.pitch <= 256;
  a.tmaV = tma_on && (a.cv % 32) == 0 && !a.v_ones && 2 * a.pitch <= 256 && !a.tapsplit;
  a.box_bytes = a.pitch * 128;
  auto rup = [](size_t x) { return (x + 1023) & ~(size_t)1023; };
  // larger stages first (the per-stage fixed cost is ~2-3 k cycles), then ring depth
  for (int ks : {128, 64, 32, 16})
    for (int want = 3; want >= 2 && !a.KS; --want) {
      const int nposw = (a.lead + ks + trail + 7) & ~7;
      const int nru = (ks + a.pitch - 2) / a.pitch + 1, nrv = (a.lead + ks + trail + a.pitch - 2) / a.pitch + 1;
      const size_t regU = a.tmaU ? rup((size_t)nru * a.pitch * 128) : (size_t)ks * 128;
      const size_t regV = a.tmaV ? rup((size_t)nrv * a.pitch * 128 + 256) : (size_t)nposw * 128;
      if ((a.tmaU && nru * a.pitch > WW_PROD) || regU > 0x3FFFFu || regV > 0x3FFFFu) continue;
      // an M=128 MMA always reads four 32-channel blocks `regU` apart: when U has fewer, the surplus blocks must
      // still fall inside the stage (their accumulator rows are ignored), so the stage is padded if V is short
      size_t st = (size_t)a.NPU * a.nbU * regU + (size_t)a.NPV * a.nbV * regV;
      st = std::max(st, ((size_t)(a.NPU - 1) * a.nbU + 4) * regU + 2048);
      const size_t tabs = 2 * ((size_t)a.NPU * ks + (size_t)a.NPV * nposw) * 4;
      if (want * st + tabs + 128 <= cap) {
        a.KS = ks; a.NPOSw = nposw; a.stage_bytes = (int)st; a.NRU = nru; a.NRV = nrv; a.regU = (int)regU; a.regV = (int)regV;
        a.nstages = (int)std::min<size_t>(WW_MAXSTAGES, (cap - tabs - 128) / st);
        break;
      }
    }
  if (!a.KS) return p;
  a.v_off = a.NPU * a.nbU * a.regU;
  a.lboB = a.packed ? 128u : (uint32_t)a.regV;
  p.smem = 1024 + (size_t)a.nstages * a.stage_bytes + 2 * ((size_t)a.NPU * a.KS + (size_t)a.NPV * a.NPOSw) * 4 + 128;

  a.SPL = std::max(1, sm_budget() / (S * a.TG));
  {   // tools only (BSDE_WW_MINST): at least this many stages per CTA -- fewer, longer CTAs and fewer partial tiles for small grids
    static const int minst = getenv("BSDE_WW_MINST") ? atoi(getenv("BSDE_WW_MINST")) : 0;
    if (minst > 0) a.SPL = std::max(1, std::min(a.SPL, (a.Q + minst * a.KS - 1) / (minst * a.KS)));
  }
  a.chunk = (a.Q + a.SPL - 1) / a.SPL;
  a.chunk = (a.chunk + a.KS - 1) / a.KS * a.KS;
  a.SPL = (a.Q + a.chunk - 1) / a.chunk;

  for (int g = 0; g < 2; ++g) a.ncnt[g] = 0;
  for (int i = 0; i < ne; ++i) {
    const int g = egrp[i], k = a.ncnt[g]++;
    a.aoff[g][k] = (uint32_t)(ents[i].up * a.nbU * a.regU) >> 4;
    a.boff[g][k] = (uint32_t)(a.v_off + ents[i].vp * a.nbV * a.regV + (a.lead + ents[i].sh) * 128) >> 4;
    for (int j = 0; j < 3; ++j) { a.tapj[g][k][j] = ents[i].tapj[j]; a.bias_row[g][k][j] = ents[i].brow[j]; }
  }
  p.part_floats = (size_t)S * a.SPL * a.rows * a.cols;
  p.ok = true;
  return p;
}

bool tconv_wgrad_win_supported(const bsde_conv_layer* L, int S, int B) { return ww_plan(L, S, B).ok; }

size_t tconv_wgrad_ws_win(const bsde_conv_layer* L, int S, int B) {
  WwPlan p = ww_plan(L, S, B);
  return p.ok ? p.part_floats * sizeof(float) : 0;
}

int tconv_wgrad_win(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t, float scale,
                    float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes, cudaStream_t st) {
  WwPlan p = ww_plan(L, S, B);
  BSDE_CHECK_ARG(p.ok, "layer not supported by the window-staged weight gradient");
  const size_t need = p.part_floats * sizeof(float);
  if (!ws || ws_bytes < need) {
    set_error("tconv_wgrad(win): workspace %zu < %zu bytes", ws_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  WwArgs& a = p.a;
  a.U = a.u_is_input ? in : dy;
  a.V = a.u_is_input ? dy : in;
  a.partial = (float*)ws;
  a.t = t;
  { static const int dbg = getenv("BSDE_WW_DBG") ? atoi(getenv("BSDE_WW_DBG")) : 0; a.dbg = dbg; }
  {
    // the __device__ symbol has one address per device
    static float* one_ptr[64] = {nullptr};
    int dev = 0;
    BSDE_CUDA_OK(cudaGetDevice(&dev));
    float*& op = one_ptr[dev & 63];
    if (!op) BSDE_CUDA_OK(cudaGetSymbolAddress((void**)&op, ww_one_dev));
    a.one = op;
  }
  static PerDeviceOnce attr_once;
  if (attr_once.first())
    BSDE_CUDA_OK(cudaFuncSetAttribute(wgrad_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CUtensorMap tmU, tmV;
  memset(&tmU, 0, sizeof(tmU));
  memset(&tmV, 0, sizeof(tmV));
  if (a.tmaU || a.tmaV) {
    using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled enc = nullptr;
    if (!enc) {
      cudaDriverEntryPointQueryResult qres;
      BSDE_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
      if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return BSDE_ERR_CUDA; }
    }
    // [S][B][h][w][c] fp32, rank 5 so that images past the end of a sample are out of bounds (zero filled)
    auto make = [&](CUtensorMap* tm, const float* ptr, int c, int w, int h, int np) -> int {
      const cuuint64_t dims[5] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B, (cuuint64_t)S};
      const cuuint64_t strides[4] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4, (cuuint64_t)B * h * w * c * 4};
      const cuuint32_t es1 = np == 4 ? 2u : 1u;
      const cuuint32_t box[5] = {32, (cuuint32_t)a.pitch * es1, 1, 1, 1};
      const cuuint32_t es[5] = {1, es1, 1, 1, 1};
      const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(ptr), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return BSDE_ERR_CUDA; }
      return 0;
    };
    if (a.tmaU) { if (int e = make(&tmU, a.U, a.cu, a.wU, a.hU, a.NPU)) return e; }
    if (a.tmaV) { if (int e = make(&tmV, a.V, a.cv, a.wV, a.hV, a.vpar ? 4 : 1)) return e; }
  }
  BSDE_CUDA_OK(launch_pdl(wgrad_win_kernel, dim3(S * a.TG * a.SPL), dim3(WW_THREADS), p.smem, st, a, tmU, tmV));
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();

  WreduceArgs r;
  memset(&r, 0, sizeof(r));
  r.partial = a.partial; r.bias_partial = nullptr; r.gw = gw; r.gw_sample_stride = gw_sample_stride;
  r.S = S; r.splits = a.SPL; r.rows = a.rows; r.cols = a.cols; r.Kc = a.Kc; r.Ktot = a.Ktot;
  r.has_time = L->has_time; r.time_row = L->time_first ? 0 : L->cin;
  r.real_base = (L->has_time && L->time_first) ? 1 : 0; r.swapped = 0; r.cin = L->cin; r.cout = L->cout;
  r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
  r.scale = scale; r.nbias = p.nbias;
  const long long total = (long long)S * a.rows * a.cols;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  return launch_wgrad_reduce(r, blocks, st);
}

}  // namespace bsde

extern "C" int bsde_debug_set_ww_stamps(long long* d_buf) {
  return cudaMemcpyToSymbol(bsde::g_ww_dbg, &d_buf, sizeof(d_buf)) == cudaSuccess ? 0 : -3;
}
