// K2 (tensor-core mode): the time-dependent convolution as an implicit GEMM on tcgen05 tensor cores.
//
//   D[128 pixels x Npad] (fp32, TMEM)  +=  A[128 x 32] (smem, TF32)  x  B[Npad x 32]^T (smem, TF32)
//
// per k-block of 32 reduction indices, where the reduction index runs over (valid tap, input channel).
//   * A is gathered on the fly (im2col never materialised): four producer warps issue 16-byte (or
//     4-byte, for channel counts that are not a multiple of 4) cp.async copies with zero-fill for
//     padding / out-of-range taps, straight into the 128B-swizzled K-major UMMA layout.
//   * B comes from a per-(sample, iteration) repack of the flat SDE state w(t_k): a small kernel writes
//     ready-to-copy swizzled K-major tile images (zero padded to Npad x 32), so the main loop moves them
//     with straight 16-byte copies.
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=Npad, K=8) x4 per k-block,
//     tcgen05.commit releases the smem stage / signals the epilogue through mbarriers.
//   * the epilogue reads the accumulator with tcgen05.ld (32x32b), adds the bias and the time-channel
//     term  t * sum_{taps inside the image} W[tap, t_ch, :]  (the time channel is spatially constant, so
//     it never enters the GEMM), applies the activation / Euler / data-gradient epilogues and stores.
// The same kernel evaluates all forward layers and all data gradients (see the spatial map in bsde.h);
// lhs-dilated layers are split into four output-parity classes so that no structural zeros are issued.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace bsde {

__device__ long long* g_tc_dbg = nullptr;  // optional phase timestamps (tools/), NULL in production
#define TC_STAMP(slot) do { if (g_tc_dbg && threadIdx.x == (slot == 3 ? 256 : 0) && blockIdx.x < 64 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_dbg[blockIdx.x * 8 + slot] = clock64(); } while (0)

constexpr int TC_BM = 128;
constexpr int TC_KB = 32;  // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_LAG = 2;  // cp.async groups kept in flight per producer thread
constexpr int TC_PROD = 256;
constexpr int TC_THREADS = TC_PROD + 32;
constexpr int TC_A_BYTES = TC_BM * 128;

struct TcArgs {
  const float* in;
  const float* w;
  const float* aux;
  float* out;
  const float* bimg;
  long long w_sample_stride, bimg_sample_stride;
  int S, B;
  int cin, cout, ksize, hin, win, hout, wout, sn, kn, off, sd;
  int has_time, time_row;
  long long w_off, b_off, ts, ks, ns;
  float t, scale;
  int act, epi;
  int Kc, Mper, npad, tmem_cols;
  int cls_off[4];  // float offset of each parity class' image inside a sample's image block
  int cls_tb_off[4];   // float offset of the class' bias / time-row table  [(ntap+2)][npad]
  int cls_tap_off[4];  // float offset of the class' gather table (TcTap entries)
  int dbg;         // tools/ only: 1 skip A gather, 2 skip B copy, 4 skip MMA, 8 skip epilogue stores
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ bool tc_map_axis(int o, int k, int sn, int kn, int off, int sd, int size, int& i) {
  int num = o * sn + k * kn + off;
  if (num < 0) return false;
  if (sd == 2) {
    if (num & 1) return false;
    num >>= 1;
  }
  i = num;
  return num < size;
}
__device__ __host__ __forceinline__ int tc_valid_taps(int ksize, int sn, int kn, int off, int sd, int par, int* kv) {
  int n = 0;
  for (int k = 0; k < ksize; ++k)
    if (sd == 1 || (((par * sn + k * kn + off) & 1) == 0)) kv[n++] = k;
  return n;
}

// one gather-table entry per 16-byte chunk (VEC) or per element (scalar mode) of every k-block:
// input offset relative to the row's base pixel, and the tap displacement for the bounds test.
struct __align__(8) TcTap {
  int toff;       // ((dy*win + dx)*cin + c) in floats; INT_MIN/2 marks "beyond K"
  short dy, dx;
};

// ---- weight repack: flat w(t_k) -> swizzled K-major [kb][npad][32] tile images, plus (last block of
// the grid) the class' epilogue table {bias, sum of time rows, time row per tap} and gather table ----
struct RepackClass {
  int cls_off, tb_off, tap_off, Kcls, nkh, nkw, nkb, py, px, blk0;
  int khv[3], kwv[3];
};
struct RepackArgs {
  const float* w;
  float* img;
  long long w_sample_stride, img_sample_stride;
  int ncls, Kc, ksize, cout, npad, real_base, vec;
  int sn, kn, off, sd, win, cin, has_time, time_row, use_bias;
  long long w_off, b_off, ts, ks, ns;
  RepackClass c[4];
};
// per-class view with the field names the body uses
struct RepackView {
  const float* w;
  float* img;
  long long w_sample_stride, img_sample_stride;
  int cls_off, tb_off, tap_off, Kc, Kcls, ksize, nkh, nkw, cout, npad, real_base, nkb, vec;
  int py, px, sn, kn, off, sd, win, cin, has_time, time_row, use_bias;
  int khv[3], kwv[3];
  long long w_off, b_off, ts, ks, ns;
};

__global__ void tc_repack_kernel(RepackArgs A) {
  // blockIdx.x enumerates (class, k-block) pairs; the last block of every class writes its tables
  int ci = 0;
  while (ci + 1 < A.ncls && (int)blockIdx.x >= A.c[ci + 1].blk0) ++ci;
  RepackView a;
  a.w = A.w; a.img = A.img; a.w_sample_stride = A.w_sample_stride; a.img_sample_stride = A.img_sample_stride;
  a.cls_off = A.c[ci].cls_off; a.tb_off = A.c[ci].tb_off; a.tap_off = A.c[ci].tap_off; a.Kc = A.Kc; a.Kcls = A.c[ci].Kcls;
  a.ksize = A.ksize; a.nkh = A.c[ci].nkh; a.nkw = A.c[ci].nkw; a.cout = A.cout; a.npad = A.npad; a.real_base = A.real_base;
  a.nkb = A.c[ci].nkb; a.vec = A.vec; a.py = A.c[ci].py; a.px = A.c[ci].px; a.sn = A.sn; a.kn = A.kn; a.off = A.off; a.sd = A.sd;
  a.win = A.win; a.cin = A.cin; a.has_time = A.has_time; a.time_row = A.time_row; a.use_bias = A.use_bias;
  for (int i = 0; i < 3; ++i) { a.khv[i] = A.c[ci].khv[i]; a.kwv[i] = A.c[ci].kwv[i]; }
  a.w_off = A.w_off; a.b_off = A.b_off; a.ts = A.ts; a.ks = A.ks; a.ns = A.ns;
  const int kb = (int)blockIdx.x - A.c[ci].blk0, s = blockIdx.y;
  const float* __restrict__ w = a.w + (long long)s * a.w_sample_stride;
  if (kb == a.nkb) {
    const int ntap = a.nkh * a.nkw;
    float* tb = a.img + (long long)s * a.img_sample_stride + a.tb_off;
    for (int idx = threadIdx.x; idx < (ntap + 2) * a.npad; idx += blockDim.x) {
      const int r = idx / a.npad, n = idx - r * a.npad;
      float v = 0.f;
      if (n < a.cout) {
        if (r == 0) v = (a.b_off >= 0 && a.use_bias) ? __ldg(w + a.b_off + n) : 0.f;
        else if (a.has_time) {
          for (int tl = (r == 1 ? 0 : r - 2); tl < (r == 1 ? ntap : r - 1); ++tl) {
            const int tap = a.khv[tl / a.nkw] * a.ksize + a.kwv[tl % a.nkw];
            v += __ldg(w + a.w_off + tap * a.ts + (long long)a.time_row * a.ks + (long long)n * a.ns);
          }
        }
      }
      tb[idx] = v;
    }
    TcTap* tp = reinterpret_cast<TcTap*>(a.img + (long long)s * a.img_sample_stride + a.tap_off);
    const int per_kb = a.vec ? 8 : 32;
    for (int idx = threadIdx.x; idx < a.nkb * per_kb; idx += blockDim.x) {
      const int k = a.vec ? idx * 4 : idx;
      TcTap e;
      e.toff = INT_MIN / 2; e.dy = 0; e.dx = 0;
      if (k < a.Kcls) {
        const int tl = k / a.Kc, c = k - tl * a.Kc;
        const int kh = a.khv[tl / a.nkw], kw = a.kwv[tl % a.nkw];
        // per-axis linear map of the parity class: i = q*sn + (par*sn + k*kn + off)/sd  (exact division)
        const int dy = (a.py * a.sn + kh * a.kn + a.off) / a.sd;
        const int dx = (a.px * a.sn + kw * a.kn + a.off) / a.sd;
        e.dy = (short)dy; e.dx = (short)dx;
        e.toff = (dy * a.win + dx) * a.cin + c;
      }
      tp[idx] = e;
    }
    return;
  }
  float* __restrict__ img = a.img + (long long)s * a.img_sample_stride + a.cls_off + (long long)kb * a.npad * TC_KB;
  for (int idx = threadIdx.x; idx < a.npad * TC_KB; idx += blockDim.x) {
    const int n = idx % a.npad, kk = idx / a.npad;
    const int k = kb * TC_KB + kk;
    float v = 0.f;
    if (k < a.Kcls && n < a.cout) {
      const int tl = k / a.Kc, c = k - tl * a.Kc;
      const int tap = a.khv[tl / a.nkw] * a.ksize + a.kwv[tl % a.nkw];
      v = __ldg(w + a.w_off + tap * a.ts + (long long)(c + a.real_base) * a.ks + (long long)n * a.ns);
    }
    img[(n >> 3) * 256 + (n & 7) * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3))] = v;
  }
}

// ---- main kernel -------------------------------------------------------------------------------
// fast activation forms for this path (TF32 operand rounding dominates their ~1e-6 error)
__device__ __forceinline__ float tc_act_fwd(int act, float x) {
  switch (act) {
    case BSDE_ACT_SOFTPLUS: return x > 15.f ? x : __logf(1.f + __expf(x));
    case BSDE_ACT_TANH: return tanhf(x);
    case BSDE_ACT_ELU: return x > 0.f ? x : __expf(x) - 1.f;
    case BSDE_ACT_RBF: return rbf_fwd_tagged(x);
    default: return x;
  }
}
__device__ __forceinline__ float tc_act_grad_out(int act, float a) {
  switch (act) {
    case BSDE_ACT_SOFTPLUS: return 1.f - __expf(-a);
    case BSDE_ACT_TANH: return 1.f - a * a;
    case BSDE_ACT_ELU: return a > 0.f ? 1.f : a + 1.f;
    case BSDE_ACT_RBF: return rbf_grad_from_tagged(a);
    default: return 1.f;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.z, cls = blockIdx.y;
  const int m0 = blockIdx.x * TC_BM;
  const int py = cls >> 1, px = cls & 1;
  const int step = a.sd == 2 ? 2 : 1;
  const int hq = (a.hout - py + step - 1) / step, wq = (a.wout - px + step - 1) / step;
  const int Mcls = a.B * hq * wq;
  if (m0 >= Mcls) return;
  TC_STAMP(0);
  int khv[3], kwv[3];
  const int nkh = tc_valid_taps(a.ksize, a.sn, a.kn, a.off, a.sd, py, khv);
  const int nkw = tc_valid_taps(a.ksize, a.sn, a.kn, a.off, a.sd, px, kwv);
  const int ntap = nkh * nkw;
  const int Kcls = ntap * a.Kc;
  const int NKB = (Kcls + TC_KB - 1) / TC_KB;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int stage_bytes = TC_A_BYTES + a.npad * 128;
  int soff = TC_STAGES * stage_bytes;
  const uint32_t bars = base + soff;  // full[STAGES], empty[STAGES], tmem_full
  soff += (8 * (2 * TC_STAGES + 1) + 15) & ~15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + soff);
  soff += 16;
  float* s_tb = reinterpret_cast<float*>(gen + soff);  // [0]=bias, [1]=sum of all time rows, [2+q]=time row of tap q
  soff += 11 * a.npad * 4;
  TcTap* s_tap = reinterpret_cast<TcTap*>(gen + soff);  // [NKB][8] (VEC) or [NKB][32]
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto empty_bar = [&](int st) { return bars + 8u * (TC_STAGES + st); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * TC_STAGES);

  const float* __restrict__ in = a.in + (long long)s * a.B * a.hin * a.win * a.cin;
  const float* __restrict__ w = a.w + (long long)s * a.w_sample_stride;
  const float* __restrict__ bimg = a.bimg + (long long)s * a.bimg_sample_stride + a.cls_off[cls];

  // per-axis linear map of this parity class:  i = q*sn + (par*sn + k*kn + off)/sd   (exact division)
  if (warp == TC_PROD / 32) {
    if (lane == 0) {
      for (int i = 0; i < TC_STAGES; ++i) { mbar_init(full_bar(i), TC_PROD); mbar_init(empty_bar(i), 1); }
      mbar_init(tmem_full_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
  } else {
    // epilogue table {bias, sum of time rows, time row per tap} and gather table -> smem
    const float* __restrict__ gtb = a.bimg + (long long)s * a.bimg_sample_stride + a.cls_tb_off[cls];
#pragma unroll 1
    for (int idx = tid; idx < (ntap + 2) * a.npad; idx += TC_PROD) s_tb[idx] = __ldg(gtb + idx);
    const int2* __restrict__ gtap = reinterpret_cast<const int2*>(a.bimg + (long long)s * a.bimg_sample_stride + a.cls_tap_off[cls]);
    int2* s_tap2 = reinterpret_cast<int2*>(s_tap);
#pragma unroll 1
    for (int idx = tid; idx < NKB * (VEC ? 8 : 32); idx += TC_PROD) s_tap2[idx] = __ldg(gtap + idx);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  TC_STAMP(1);

  if (warp < TC_PROD / 32) {
    // =============================== producers ===============================
    constexpr int RPT = TC_BM * 8 / TC_PROD;      // rows per thread
    constexpr int RSTEP = TC_PROD / 8;
    const int c = tid & 7;  // 16-byte chunk column inside the 128-byte row
    int p_by[RPT], p_bx[RPT], p_off[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int m = m0 + (tid >> 3) + RSTEP * i;
      if (m < Mcls) {
        const int hw = hq * wq;
        const int img = m / hw;
        const int r = m - img * hw;
        const int qy = r / wq, qx = r - qy * wq;
        p_by[i] = qy * a.sn * (a.sd == 2 ? 1 : 1);
        p_bx[i] = qx * a.sn;
        if (a.sd == 1) { p_by[i] = (qy * step + py) * a.sn; p_bx[i] = (qx * step + px) * a.sn; }
        p_off[i] = ((img * a.hin + p_by[i]) * a.win + p_bx[i]) * a.cin;
      } else {
        p_by[i] = -100000; p_bx[i] = -100000; p_off[i] = 0;
      }
    }
#pragma unroll 1
    for (int kb = 0; kb < NKB; ++kb) {
      const int st = kb % TC_STAGES;
      if (kb >= TC_STAGES) mbar_wait(empty_bar(st), ((kb / TC_STAGES) & 1) ^ 1);
      const uint32_t sA = base + st * stage_bytes;
      const uint32_t sB = sA + TC_A_BYTES;
      if (a.dbg & 1) {
      } else if (VEC) {
        const TcTap e = s_tap[kb * 8 + c];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = (tid >> 3) + RSTEP * i;
          const bool ok = (unsigned)(p_by[i] + e.dy) < (unsigned)a.hin && (unsigned)(p_bx[i] + e.dx) < (unsigned)a.win &&
                          e.toff > INT_MIN / 4;
          const float* src = ok ? in + (p_off[i] + e.toff) : in;
          cp_async16(sA + r * 128 + ((c ^ (r & 7)) << 4), src, ok);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const TcTap e = s_tap[kb * 32 + c * 4 + j];
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = (tid >> 3) + RSTEP * i;
            const bool ok = (unsigned)(p_by[i] + e.dy) < (unsigned)a.hin &&
                            (unsigned)(p_bx[i] + e.dx) < (unsigned)a.win && e.toff > INT_MIN / 4;
            const float* src = ok ? in + (p_off[i] + e.toff) : in;
            cp_async4(sA + r * 128 + ((c ^ (r & 7)) << 4) + 4 * j, src, ok);
          }
        }
      }
      // weight tile image: straight copy
      const float* bsrc = bimg + (long long)kb * a.npad * TC_KB;
      if (!(a.dbg & 2))
  #pragma unroll 1
        for (int idx = tid; idx < a.npad * 8; idx += TC_PROD) cp_async16(sB + idx * 16, bsrc + idx * 4, true);
      // arrive on the stage's barrier when this thread's copies have landed (no blocking here)
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(st)) : "memory");
    }
    TC_STAMP(2);

    // =============================== epilogue ===============================
    // warp w reads TMEM lanes 32*(w%4)..+31 (its rows) and the 16-column chunks j with j % NHALF == w/4
    constexpr int NHALF = TC_PROD / 128;
    const int row = (warp & 3) * 32 + lane;
    const int m = m0 + row;
    const bool mval = m < Mcls;
    long long pix = 0;
    uint32_t tapmask = 0;
    bool all_taps = true;
    if (mval) {
      const int hw = hq * wq;
      const int img = m / hw;
      const int r = m - img * hw;
      const int qy = r / wq;
      const int oy = qy * step + py, ox = (r - qy * wq) * step + px;
      pix = ((long long)img * a.hout + oy) * a.wout + ox;
      if (a.has_time) {
#pragma unroll 1
        for (int ih = 0; ih < nkh; ++ih)
#pragma unroll 1
          for (int iw = 0; iw < nkw; ++iw) {
            int iy, ix;
            const bool ok = tc_map_axis(oy, khv[ih], a.sn, a.kn, a.off, a.sd, a.hin, iy) &&
                            tc_map_axis(ox, kwv[iw], a.sn, a.kn, a.off, a.sd, a.win, ix);
            if (ok) tapmask |= 1u << (ih * nkw + iw);
            else all_taps = false;
          }
      }
    }
    float* __restrict__ out = a.out + (long long)s * a.Mper * a.cout;
    const float* __restrict__ aux = a.aux ? a.aux + (long long)s * a.Mper * a.cout : nullptr;
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // staging tile in the (now idle) stage buffers: [128 rows][npad + 4] floats, then per-row pixel / tap mask
    float* s_out = reinterpret_cast<float*>(gen);
    const int ldo = a.npad + 4;
    long long* s_pix = reinterpret_cast<long long*>(gen + TC_BM * ldo * 4);
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(gen + TC_BM * ldo * 4 + TC_BM * 8);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    TC_STAMP(4);
    if ((warp >> 2) == 0) {
      s_pix[row] = mval ? pix : -1;
      s_mask[row] = all_taps ? 0xFFFFFFFFu : tapmask;
    }
    for (int n0 = (warp >> 2) * 16; n0 < a.npad; n0 += 16 * NHALF) {
      uint32_t v[16];
      tmem_ld16(trow + n0, v);
      uint4* dst = reinterpret_cast<uint4*>(s_out + row * ldo + n0);
      dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
      dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
      dst[2] = make_uint4(v[8], v[9], v[10], v[11]);
      dst[3] = make_uint4(v[12], v[13], v[14], v[15]);
    }
    // all 8 epilogue warps: staging complete -> coalesced bias / time term / epilogue op / store
    asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");
    if (!(a.dbg & 8) && (a.cout & 3) == 0 && (TC_PROD % (a.cout >> 2)) == 0) {
      // fast path: a fixed float4 column per thread -> bias / time sums live in registers
      const int cn = a.cout >> 2;
      const int c4 = tid % cn, r0 = tid / cn, rstep = TC_PROD / cn;
      const float4 b4 = *reinterpret_cast<const float4*>(s_tb + c4 * 4);
      float4 t4 = *reinterpret_cast<const float4*>(s_tb + a.npad + c4 * 4);
      t4.x *= a.t; t4.y *= a.t; t4.z *= a.t; t4.w *= a.t;
#pragma unroll 1
      for (int rr = r0; rr < TC_BM; rr += rstep) {
        const long long pp = s_pix[rr];
        if (pp < 0) continue;
        const uint32_t mask = s_mask[rr];
        float4 q = *reinterpret_cast<const float4*>(s_out + rr * ldo + c4 * 4);
        q.x += b4.x; q.y += b4.y; q.z += b4.z; q.w += b4.w;
        if (a.has_time) {
          if (mask == 0xFFFFFFFFu) {
            q.x += t4.x; q.y += t4.y; q.z += t4.z; q.w += t4.w;
          } else {
#pragma unroll 1
            for (int t = 0; t < ntap; ++t)
              if ((mask >> t) & 1) {
                const float4 tt = *reinterpret_cast<const float4*>(s_tb + (t + 2) * a.npad + c4 * 4);
                q.x = fmaf(a.t, tt.x, q.x); q.y = fmaf(a.t, tt.y, q.y); q.z = fmaf(a.t, tt.z, q.z); q.w = fmaf(a.t, tt.w, q.w);
              }
          }
        }
        const long long o = pp * a.cout + c4 * 4;
        if (a.epi == BSDE_EPI_BIAS_ACT) {
          q.x = tc_act_fwd(a.act, q.x); q.y = tc_act_fwd(a.act, q.y); q.z = tc_act_fwd(a.act, q.z); q.w = tc_act_fwd(a.act, q.w);
        } else if (a.epi >= BSDE_EPI_EULER) {
          const float4 x = *reinterpret_cast<const float4*>(aux + o);
          if (a.epi == BSDE_EPI_DACT) {
            q.x = a.scale * q.x * tc_act_grad_out(a.act, x.x); q.y = a.scale * q.y * tc_act_grad_out(a.act, x.y);
            q.z = a.scale * q.z * tc_act_grad_out(a.act, x.z); q.w = a.scale * q.w * tc_act_grad_out(a.act, x.w);
          } else {
            q.x = fmaf(a.scale, q.x, x.x); q.y = fmaf(a.scale, q.y, x.y); q.z = fmaf(a.scale, q.z, x.z); q.w = fmaf(a.scale, q.w, x.w);
          }
        }
        *reinterpret_cast<float4*>(out + o) = q;
      }
    } else if (!(a.dbg & 8)) {
      const int cn = (a.cout & 3) == 0 ? (a.cout >> 2) : a.cout;  // columns in units of VW floats
      const int VW = (a.cout & 3) == 0 ? 4 : 1;
#pragma unroll 1
      for (int idx = tid; idx < TC_BM * cn; idx += TC_PROD) {
        const int rr = idx / cn, n = (idx - rr * cn) * VW;
        const long long pp = s_pix[rr];
        if (pp < 0) continue;
        const uint32_t mask = s_mask[rr];
        const long long o = pp * a.cout + n;
        if (a.dbg & 32) { *reinterpret_cast<float4*>(out + o) = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
        float q[4], x[4];
        for (int e = 0; e < VW; ++e) {
          float acc = s_out[rr * ldo + n + e] + s_tb[n + e];
          if (a.has_time) {
            float tsum;
            if (mask == 0xFFFFFFFFu) {
              tsum = s_tb[a.npad + n + e];
            } else {
              tsum = 0.f;
#pragma unroll 1
              for (int t = 0; t < ntap; ++t)
                if ((mask >> t) & 1) tsum += s_tb[(t + 2) * a.npad + n + e];
            }
            acc = fmaf(a.t, tsum, acc);
          }
          q[e] = acc;
        }
        if (a.epi >= BSDE_EPI_EULER) {
          if (VW == 4) {
            const float4 xv = *reinterpret_cast<const float4*>(aux + o);
            x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
          } else {
            x[0] = aux[o];
          }
        }
        for (int e = 0; e < VW; ++e) {
          switch (a.epi) {
            case BSDE_EPI_BIAS: break;
            case BSDE_EPI_BIAS_ACT: q[e] = tc_act_fwd(a.act, q[e]); break;
            case BSDE_EPI_DACT: q[e] = a.scale * q[e] * tc_act_grad_out(a.act, x[e]); break;
            default: q[e] = x[e] + a.scale * q[e]; break;  // EULER / ACCUM
          }
        }
        if ((a.dbg & 16) && q[0] != 12345.678f) continue;
        if (VW == 4) *reinterpret_cast<float4*>(out + o) = make_float4(q[0], q[1], q[2], q[3]);
        else out[o] = q[0];
      }
    }
  } else {
    // =============================== MMA issuer ===============================
    // whole warp, uniform operands, one elected lane issues (no per-MMA R2UR waterfall)
    {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.npad >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint64_t hi = um::desc_hi(16u, 1024u, 2);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
#pragma unroll 1
      for (int kb = 0; kb < NKB; ++kb) {
        const int st = kb % TC_STAGES;
        mbar_wait(full_bar(st), (kb / TC_STAGES) & 1);
        tc_fence_after();
        const uint32_t sA16 = (base + st * stage_bytes) >> 4;
        const uint32_t sB16 = sA16 + (TC_A_BYTES >> 4);
        if (!(a.dbg & 4)) {
#pragma unroll
          for (int k = 0; k < TC_KB / 8; ++k)
            um::umma_tf32_elect(tm, hi | (uint64_t)((sA16 + (uint32_t)(k * 2)) & 0x3FFFu), hi | (uint64_t)((sB16 + (uint32_t)(k * 2)) & 0x3FFFu),
                                idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        um::umma_commit_elect(empty_bar(st));
      }
      um::umma_commit_elect(tmem_full_bar);
    }
    TC_STAMP(3);
    __syncwarp();
  }
  TC_STAMP(5);
  tc_fence_before();
  __syncthreads();
  if (warp == TC_PROD / 32) tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  TC_STAMP(6);
}

// ---- host side ----------------------------------------------------------------------------------
static int tc_npad(int cout) { return (cout + 15) / 16 * 16; }

static void tc_class_info(const bsde_conv_layer* L, int cls, int* khv, int* kwv, int& nkh, int& nkw) {
  nkh = tc_valid_taps(L->ksize, L->sn, L->kn, L->off, L->sd, cls >> 1, khv);
  nkw = tc_valid_taps(L->ksize, L->sn, L->kn, L->off, L->sd, cls & 1, kwv);
}

bool tconv_tc_supported(const bsde_conv_layer* L, int epi) {
  (void)epi;
  if (!L) return false;
  if (L->ksize != 1 && L->ksize != 3) return false;
  if (L->cout > 128 || L->cout < 1) return false;
  if (L->sd == 2 && !(L->sn & 1)) return false;
  const int ncls = L->sd == 2 ? 4 : 1;
  for (int c = 0; c < ncls; ++c) {
    int khv[3], kwv[3], nkh, nkw;
    tc_class_info(L, c, khv, kwv, nkh, nkw);
    if (nkh * nkw < 1) return false;
  }
  return true;
}

size_t tconv_tc_ws_bytes(const bsde_conv_layer* L, int S) {
  const int npad = tc_npad(L->cout);
  const int ncls = L->sd == 2 ? 4 : 1;
  size_t fl = 0;
  for (int c = 0; c < ncls; ++c) {
    int khv[3], kwv[3], nkh, nkw;
    tc_class_info(L, c, khv, kwv, nkh, nkw);
    const int Kcls = nkh * nkw * L->cin;
    const int nkb = (Kcls + TC_KB - 1) / TC_KB;
    fl += (size_t)nkb * npad * TC_KB + 11 * npad + (size_t)nkb * (L->cin % 4 == 0 ? 8 : 32) * 2;
    fl = (fl + 3) & ~(size_t)3;
  }
  return fl * sizeof(float) * S;
}

int tconv_fwd_tc(const bsde_conv_layer* L, int S, int B, const float* in, const float* w,
                 long long w_sample_stride, float t, int act, int epi, float scale, const float* aux,
                 float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  BSDE_CHECK_ARG(tconv_tc_supported(L, epi), "layer not supported by the tensor-core path");
  BSDE_CHECK_ARG(!(epi >= BSDE_EPI_EULER) || aux != nullptr, "epilogue %d needs aux", epi);
  const size_t need = tconv_tc_ws_bytes(L, S);
  if (!ws || ws_bytes < need) {
    set_error("tconv_fwd(tc): workspace %zu < %zu bytes", ws_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  TcArgs a;
  memset(&a, 0, sizeof(a));
  const int npad = tc_npad(L->cout);
  const int ncls = L->sd == 2 ? 4 : 1;
  a.in = in; a.w = w; a.aux = aux; a.out = out; a.bimg = (const float*)ws;
  a.w_sample_stride = w_sample_stride; a.bimg_sample_stride = (long long)(need / sizeof(float) / S);
  a.S = S; a.B = B;
  a.cin = L->cin; a.cout = L->cout; a.ksize = L->ksize; a.hin = L->hin; a.win = L->win;
  a.hout = L->hout; a.wout = L->wout; a.sn = L->sn; a.kn = L->kn; a.off = L->off; a.sd = L->sd;
  a.has_time = L->has_time; a.time_row = L->time_first ? 0 : L->cin;
  a.w_off = L->w_off; a.b_off = L->b_off; a.ts = L->w_tap_stride; a.ks = L->w_k_stride; a.ns = L->w_n_stride;
  a.t = t; a.scale = scale; a.act = act; a.epi = epi;
  a.Kc = L->cin; a.Mper = B * L->hout * L->wout; a.npad = npad;
  a.tmem_cols = npad <= 32 ? 32 : (npad <= 64 ? 64 : 128);
  { const char* e = getenv("BSDE_TC_DBG"); a.dbg = e ? atoi(e) : 0; }
  // repack w(t_k) into tile images + tables: one launch covering every parity class
  {
    RepackArgs r;
    memset(&r, 0, sizeof(r));
    r.w = w; r.img = (float*)ws; r.w_sample_stride = w_sample_stride; r.img_sample_stride = a.bimg_sample_stride;
    r.ncls = ncls; r.Kc = L->cin; r.ksize = L->ksize; r.cout = L->cout; r.npad = npad;
    r.real_base = (L->has_time && L->time_first) ? 1 : 0; r.vec = L->cin % 4 == 0;
    r.sn = L->sn; r.kn = L->kn; r.off = L->off; r.sd = L->sd; r.win = L->win; r.cin = L->cin;
    r.has_time = L->has_time; r.time_row = a.time_row; r.use_bias = epi <= BSDE_EPI_EULER;
    r.w_off = L->w_off; r.b_off = L->b_off; r.ts = L->w_tap_stride; r.ks = L->w_k_stride; r.ns = L->w_n_stride;
    int off = 0, blk = 0;
    for (int c = 0; c < ncls; ++c) {
      RepackClass& rc = r.c[c];
      tc_class_info(L, c, rc.khv, rc.kwv, rc.nkh, rc.nkw);
      rc.Kcls = rc.nkh * rc.nkw * L->cin;
      rc.nkb = (rc.Kcls + TC_KB - 1) / TC_KB;
      rc.py = c >> 1; rc.px = c & 1; rc.blk0 = blk;
      blk += rc.nkb + 1;
      rc.cls_off = off; a.cls_off[c] = off;
      off += rc.nkb * npad * TC_KB;
      rc.tb_off = off; a.cls_tb_off[c] = off;
      off += 11 * npad;
      rc.tap_off = off; a.cls_tap_off[c] = off;
      off += rc.nkb * (r.vec ? 8 : 32) * 2;
      off = (off + 3) & ~3;  // keep every class block 16-byte aligned
    }
    tc_repack_kernel<<<dim3(blk, S), 256, 0, st>>>(r);
    BSDE_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  const int stage_bytes = TC_A_BYTES + npad * 128;
  int maxkb = 0;
  for (int c = 0; c < ncls; ++c) {
    int khv[3], kwv[3], nkh, nkw;
    tc_class_info(L, c, khv, kwv, nkh, nkw);
    maxkb = std::max(maxkb, (nkh * nkw * L->cin + TC_KB - 1) / TC_KB);
  }
  const size_t smem = 1024 + (size_t)TC_STAGES * stage_bytes + ((8 * (2 * TC_STAGES + 1) + 15) & ~15) + 16 + (size_t)11 * npad * sizeof(float) +
                      (size_t)maxkb * (L->cin % 4 == 0 ? 8 : 32) * 8;
  const int Mcls = L->sd == 2 ? B * ((L->hout + 1) / 2) * ((L->wout + 1) / 2) : a.Mper;
  dim3 grid((Mcls + TC_BM - 1) / TC_BM, ncls, S);
  // opt in to the large dynamic shared memory once per kernel (the attribute is per device)
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    BSDE_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    BSDE_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (L->cin % 4 == 0) conv_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(a);
  else conv_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(a);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace bsde

extern "C" int bsde_debug_set_tc_stamps(long long* d_buf) {
  return cudaMemcpyToSymbol(bsde::g_tc_dbg, &d_buf, sizeof(d_buf)) == cudaSuccess ? 0 : -3;
}
