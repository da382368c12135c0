// C-ABI surface of libbsde.so: error plumbing, K2 entry points, the x-path scan drivers (K2+K3),
// the generic flat-state step and the Philox increment materialiser.
#include <stdarg.h>

#include <algorithm>
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace bsde {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
static thread_local int g_sm_budget = 148;
int sm_budget() { return g_sm_budget; }
void set_sm_budget(int n) { g_sm_budget = n < 1 ? 1 : (n > 148 ? 148 : n); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// fp32 path (tconv_simt.cu)
int tconv_fwd_simt(const bsde_conv_layer* L, int S, int B, const float* in, const float* w,
                   long long w_sample_stride, float t, int act, int epi, float scale, const float* aux,
                   float* out, cudaStream_t st);
size_t tconv_wgrad_ws_simt(const bsde_conv_layer* L, int S, int B);
int tconv_wgrad_simt(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t,
                     float scale, float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes,
                     cudaStream_t st);
// tensor-core path (tconv_tc.cu)
bool tconv_tc_supported(const bsde_conv_layer* L, int epi);
size_t tconv_tc_ws_bytes(const bsde_conv_layer* L, int S);
int tconv_fwd_tc(const bsde_conv_layer* L, int S, int B, const float* in, const float* w,
                 long long w_sample_stride, float t, int act, int epi, float scale, const float* aux,
                 float* out, void* ws, size_t ws_bytes, cudaStream_t st);

// window-staged persistent tensor-core path (tconv_win.cu)
bool tconv_win_supported(const bsde_conv_layer* L, int S, int B);
size_t tconv_win_ws_bytes(const bsde_conv_layer* L, int S);
int tconv_fwd_win(const bsde_conv_layer* L, int S, int B, const float* in, const float* w, long long w_sample_stride,
                  float t, int act, int epi, float scale, const float* aux, float* out, void* ws, size_t ws_bytes,
                  cudaStream_t st, const float* prepacked, int prepacked_safe);
size_t tconv_win_img_floats(const bsde_conv_layer* L);
int tconv_win_repack(const bsde_conv_layer* L, int S, int nit, const float* w, long long w_sample_stride, long long w_iter_stride,
                     int epi, float* img, cudaStream_t st);

// taps-on-N path for small-cout stride-1 layers (csrc/tconv_tapn.cu)
bool tconv_tapn_supported(const bsde_conv_layer* L, int S, int B, int epi);
size_t tconv_tapn_img_floats(const bsde_conv_layer* L);
int tconv_tapn_repack(const bsde_conv_layer* L, int S, int nit, const float* w, long long w_sample_stride, long long w_iter_stride,
                      int epi, float* img, cudaStream_t st);
int tconv_fwd_tapn(const bsde_conv_layer* L, int S, int B, const float* in, const float* w, long long w_sample_stride,
                   float t, int act, int epi, float scale, const float* aux, float* out, void* ws, size_t ws_bytes,
                   cudaStream_t st, const float* prepacked);
bool tconv_wgrad_tc_supported(const bsde_conv_layer* L, int S, int B);
size_t tconv_wgrad_ws_tc(const bsde_conv_layer* L, int S, int B);
int tconv_wgrad_tc(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t, float scale,
                   float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes, cudaStream_t st);

// window-staged tensor-core weight gradient (wgrad_win.cu)
bool tconv_wgrad_win_supported(const bsde_conv_layer* L, int S, int B);
size_t tconv_wgrad_ws_win(const bsde_conv_layer* L, int S, int B);
int tconv_wgrad_win(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t, float scale,
                    float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes, cudaStream_t st);

// thin layers (64 <-> <= 7 channels, stride-1 3x3): taps on the N side (csrc/wgrad_thin.cu)
bool tconv_wgrad_thin_supported(const bsde_conv_layer* L, int S, int B);
size_t tconv_wgrad_ws_thin(const bsde_conv_layer* L, int S, int B);
int tconv_wgrad_thin(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t, float scale,
                     float* gw, long long gw_sample_stride, void* ws, size_t ws_bytes, cudaStream_t st);

// trainable swish of f_x (swish.cu): elementwise activation / derivative + per-channel beta gradient
size_t swish_partial_floats(int S, long long per_sample, int C);
int swish_forward(const float* z, const float* w, long long wss, long long beta_off, int S, long long per_sample, int C, float* a,
                  cudaStream_t st);
int swish_backward(float* dz, const float* z, const float* w, long long wss, long long beta_off, int S, long long per_sample, int C,
                   float scale, float* partial, float* gw, long long gss, cudaStream_t st);

static size_t wgrad_ws(const bsde_conv_layer* L, int S, int B, int math) {
  size_t b = tconv_wgrad_ws_simt(L, S, B);
  if (math == BSDE_MATH_TC && tconv_wgrad_thin_supported(L, S, B)) b = std::max(b, tconv_wgrad_ws_thin(L, S, B));
  if (math == BSDE_MATH_TC) b = std::max(b, tconv_wgrad_ws_win(L, S, B));
  if (math == BSDE_MATH_TC && tconv_wgrad_tc_supported(L, S, B)) b = std::max(b, tconv_wgrad_ws_tc(L, S, B));
  return b;
}

static int conv_wgrad(const bsde_conv_layer* L, int S, int B, const float* in, const float* dy, float t, float scale,
                      float* gw, long long gss, int math, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (math == BSDE_MATH_TC && tconv_wgrad_thin_supported(L, S, B))
    return tconv_wgrad_thin(L, S, B, in, dy, t, scale, gw, gss, ws, ws_bytes, st);
  if (math == BSDE_MATH_TC && tconv_wgrad_win_supported(L, S, B))
    return tconv_wgrad_win(L, S, B, in, dy, t, scale, gw, gss, ws, ws_bytes, st);
  if (math == BSDE_MATH_TC && tconv_wgrad_tc_supported(L, S, B))
    return tconv_wgrad_tc(L, S, B, in, dy, t, scale, gw, gss, ws, ws_bytes, st);
  return tconv_wgrad_simt(L, S, B, in, dy, t, scale, gw, gss, ws, ws_bytes, st);
}

static int conv_fwd(const bsde_conv_layer* L, int S, int B, const float* in, const float* w,
                    long long wss, float t, int act, int epi, float scale, const float* aux, float* out,
                    int math, void* tc_ws, size_t tc_ws_bytes, cudaStream_t st, const float* prepacked = nullptr,
                    int prepacked_safe = 0) {
  if (math == BSDE_MATH_TC && tconv_tapn_supported(L, S, B, epi))
    return tconv_fwd_tapn(L, S, B, in, w, wss, t, act, epi, scale, aux, out, tc_ws, tc_ws_bytes, st, prepacked);
  if (math == BSDE_MATH_TC && tconv_win_supported(L, S, B))
    return tconv_fwd_win(L, S, B, in, w, wss, t, act, epi, scale, aux, out, tc_ws, tc_ws_bytes, st, prepacked, prepacked_safe);
  if (math == BSDE_MATH_TC && tconv_tc_supported(L, epi))
    return tconv_fwd_tc(L, S, B, in, w, wss, t, act, epi, scale, aux, out, tc_ws, tc_ws_bytes, st);
  return tconv_fwd_simt(L, S, B, in, w, wss, t, act, epi, scale, aux, out, st);
}

// data-gradient descriptor of a forward layer: reduce over cout, produce cin, inverted spatial map
static bsde_conv_layer dgrad_layer(const bsde_conv_layer& F) {
  bsde_conv_layer D = F;
  D.cin = F.cout;
  D.cout = F.cin;
  D.hin = F.hout; D.win = F.wout; D.hout = F.hin; D.wout = F.win;
  // forward: i = (o*sn + k*kn + off)/sd  <=>  o = (i*sd - k*kn - off)/sn
  D.sn = F.sd; D.kn = -F.kn; D.off = -F.off; D.sd = F.sn;
  D.has_time = 0; D.time_first = 0;
  D.b_off = -1;
  // weight index = w_off + tap*ts + row(c)*ks + n*ns ; real channel c lives in row c + (time_first?1:0)
  D.w_off = F.w_off + ((F.has_time && F.time_first) ? F.w_k_stride : 0);
  D.w_k_stride = F.w_n_stride;  // reduction runs over the forward output channel
  D.w_n_stride = F.w_k_stride;  // output runs over the forward input channel
  return D;
}

// scratch of the tensor-core forward / data-gradient paths for one layer (either kernel may be chosen at run time)
static size_t conv_tc_ws(const bsde_conv_layer* L, int S) {
  size_t m = 0;
  if (tconv_tc_supported(L, 0)) m = std::max(m, tconv_tc_ws_bytes(L, S));
  m = std::max(m, tconv_win_ws_bytes(L, S));
  m = std::max(m, tconv_tapn_img_floats(L) * sizeof(float) * S);
  return m;
}

static size_t act_elems(const bsde_xpath_desc* d, int l) {  // output of layer l
  const bsde_conv_layer& L = d->layers[l];
  return (size_t)d->S * d->B * L.hout * L.wout * L.cout;
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// stored output of layer l: the activation a_l and, for the trainable swish (whose derivative needs the PRE-activation: it is not
// monotone), z_l right behind it
static size_t act_slot_bytes(const bsde_xpath_desc* d, int l) {
  return align_up(act_elems(d, l) * sizeof(float)) * (d->act == BSDE_ACT_SWISH ? 2 : 1);
}
static float* swish_z(const bsde_xpath_desc* d, int l, float* a) { return a + align_up(act_elems(d, l) * sizeof(float)) / sizeof(float); }

}  // namespace bsde

using namespace bsde;

extern "C" int bsde_version(void) { return BSDE_VERSION; }
extern "C" const char* bsde_last_error(void) { return g_err; }
extern "C" uint64_t bsde_launch_count(void) { return g_launches.load(); }

extern "C" int bsde_tconv_fwd(const bsde_conv_layer* L, int32_t S, int32_t B, const float* d_in,
                              const float* d_w, int64_t w_sample_stride, float t, int32_t act,
                              int32_t epilogue, float scale, const float* d_aux, float* d_out, int32_t math,
                              void* workspace, size_t workspace_bytes, void* stream) {
  BSDE_CHECK_ARG(d_in && d_w && d_out, "NULL tensor pointer");
  return conv_fwd(L, S, B, d_in, d_w, w_sample_stride, t, act, epilogue, scale, d_aux, d_out, math, workspace,
                  workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t bsde_tconv_fwd_workspace_bytes(const bsde_conv_layer* L, int32_t S, int32_t math) {
  if (!L || math != BSDE_MATH_TC) return 0;
  return conv_tc_ws(L, S);
}

extern "C" size_t bsde_tconv_wgrad_workspace_bytes(const bsde_conv_layer* L, int32_t S, int32_t B) {
  if (!L) return 0;
  return wgrad_ws(L, S, B, BSDE_MATH_TC);
}

extern "C" int bsde_tconv_wgrad(const bsde_conv_layer* L, int32_t S, int32_t B, const float* d_in,
                                const float* d_dy, float t, float scale, float* d_gw,
                                int64_t gw_sample_stride, int32_t math, void* workspace,
                                size_t workspace_bytes, void* stream) {
  BSDE_CHECK_ARG(d_in && d_dy && d_gw, "NULL tensor pointer");
  return conv_wgrad(L, S, B, d_in, d_dy, t, scale, d_gw, gw_sample_stride, math, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// x-path scan.  Workspace layout:  a_0..a_{n-2} (layer outputs), [bwd: dz_0..dz_{n-2}, wgrad scratch]
// ---------------------------------------------------------------------------------------------
static int xpath_check(const bsde_xpath_desc* d) {
  BSDE_CHECK_ARG(d != nullptr, "NULL descriptor");
  BSDE_CHECK_ARG(d->nlayers >= 2 && d->nlayers <= BSDE_MAX_LAYERS, "nlayers=%d unsupported", d->nlayers);
  BSDE_CHECK_ARG(d->S >= 1 && d->B >= 1 && d->nit >= 1, "S, B, nit must be positive");
  BSDE_CHECK_ARG((d->act >= BSDE_ACT_SOFTPLUS && d->act <= BSDE_ACT_ELU) || d->act == BSDE_ACT_RBF || d->act == BSDE_ACT_SWISH,
                 "fx activation %d unsupported", d->act);
  if (d->act == BSDE_ACT_SWISH)
    for (int l = 0; l + 1 < d->nlayers; ++l)
      BSDE_CHECK_ARG(d->beta_off[l] >= 0 && d->beta_off[l] + d->layers[l].cout <= d->P, "swish: beta_off[%d] outside w", l);
  const bsde_conv_layer& L0 = d->layers[0];
  const bsde_conv_layer& Ln = d->layers[d->nlayers - 1];
  BSDE_CHECK_ARG(L0.cin == Ln.cout && L0.hin == Ln.hout && L0.win == Ln.wout,
                 "fx must map its input shape to itself (brax.py:40)");
  BSDE_CHECK_ARG((int64_t)d->S * d->B * L0.hin * L0.win * L0.cin == d->x_numel, "x_numel mismatch");
  for (int l = 0; l + 1 < d->nlayers; ++l)
    BSDE_CHECK_ARG(d->layers[l].cout == d->layers[l + 1].cin && d->layers[l].hout == d->layers[l + 1].hin &&
                       d->layers[l].wout == d->layers[l + 1].win,
                   "layer %d output does not match layer %d input", l, l + 1);
  return 0;
}

static size_t xpath_tc_bytes(const bsde_xpath_desc* d, int backward) {
  if (d->math != BSDE_MATH_TC) return 0;
  size_t m = 0;
  for (int l = 0; l < d->nlayers; ++l) {
    m = std::max(m, conv_tc_ws(&d->layers[l], d->S));
    if (backward) {
      bsde_conv_layer D = dgrad_layer(d->layers[l]);
      m = std::max(m, conv_tc_ws(&D, d->S));
    }
  }
  return align_up(m);
}

// per-CTA partial sums of the swish beta gradient (they share the weight-gradient scratch: launches are stream-ordered)
static size_t xpath_swish_bytes(const bsde_xpath_desc* d) {
  if (d->act != BSDE_ACT_SWISH) return 0;
  size_t m = 0;
  for (int l = 0; l + 1 < d->nlayers; ++l) {
    const bsde_conv_layer& L = d->layers[l];
    m = std::max(m, swish_partial_floats(d->S, (long long)d->B * L.hout * L.wout * L.cout, L.cout) * sizeof(float));
  }
  return m;
}

// Weights of ALL scan iterations are repacked up front, one launch per layer form, into
// [form][iteration][sample][image]: the per-iteration conv launches then follow each other directly.
struct Prepack {
  float* img[2 * BSDE_MAX_LAYERS];     // forward forms 0..n-1, data-gradient forms n..2n-1 (NULL: not prepacked)
  size_t per_iter[2 * BSDE_MAX_LAYERS];  // floats per iteration (S * image)
};

// resident-image forms: the taps-on-N kernel where it applies (epilogue-dependent), else the window kernel
static bool fast_form(const bsde_conv_layer* L, int S, int B, int epi) {
  return tconv_tapn_supported(L, S, B, epi) || tconv_win_supported(L, S, B);
}
static size_t fast_img_floats(const bsde_conv_layer* L, int S, int B, int epi) {
  return tconv_tapn_supported(L, S, B, epi) ? tconv_tapn_img_floats(L) : tconv_win_img_floats(L);
}
static int fast_repack(const bsde_conv_layer* L, int S, int B, int nit, const float* w, long long wss, long long wis, int epi, float* img,
                       cudaStream_t st) {
  if (tconv_tapn_supported(L, S, B, epi)) return tconv_tapn_repack(L, S, nit, w, wss, wis, epi, img, st);
  return tconv_win_repack(L, S, nit, w, wss, wis, epi, img, st);
}

static size_t prepack_floats(const bsde_xpath_desc* d, int backward, bool need_fwd_forms) {
  if (d->math != BSDE_MATH_TC) return 0;
  size_t tot = 0;
  const int n = d->nlayers;
  for (int l = 0; l < n; ++l) {
    const int epi_f = l == n - 1 ? BSDE_EPI_EULER : (d->act == BSDE_ACT_SWISH ? BSDE_EPI_BIAS : BSDE_EPI_BIAS_ACT);
    if (need_fwd_forms && fast_form(&d->layers[l], d->S, d->B, epi_f))
      tot += (size_t)d->nit * d->S * fast_img_floats(&d->layers[l], d->S, d->B, epi_f);
    if (backward) {
      bsde_conv_layer D = dgrad_layer(d->layers[l]);
      const int epi_b = l == 0 ? BSDE_EPI_ACCUM : BSDE_EPI_DACT;
      if (fast_form(&D, d->S, d->B, epi_b)) tot += (size_t)d->nit * d->S * fast_img_floats(&D, d->S, d->B, epi_b);
    }
  }
  return tot;
}

static int prepack_all(const bsde_xpath_desc* d, int backward, bool need_fwd_forms, const float* d_wtraj, float* buf, Prepack& pp,
                       cudaStream_t st) {
  memset(&pp, 0, sizeof(pp));
  if (d->math != BSDE_MATH_TC) return 0;
  const int n = d->nlayers;
  const long long wss = (long long)(d->nit + 1) * d->P;
  for (int l = 0; l < n; ++l) {
    const int epi_f = l == n - 1 ? BSDE_EPI_EULER : (d->act == BSDE_ACT_SWISH ? BSDE_EPI_BIAS : BSDE_EPI_BIAS_ACT);
    if (need_fwd_forms && fast_form(&d->layers[l], d->S, d->B, epi_f)) {
      pp.img[l] = buf; pp.per_iter[l] = (size_t)d->S * fast_img_floats(&d->layers[l], d->S, d->B, epi_f);
      if (int e = fast_repack(&d->layers[l], d->S, d->B, d->nit, d_wtraj, wss, d->P, epi_f, buf, st)) return e;
      buf += (size_t)d->nit * pp.per_iter[l];
    }
    if (backward) {
      bsde_conv_layer D = dgrad_layer(d->layers[l]);
      const int epi_b = l == 0 ? BSDE_EPI_ACCUM : BSDE_EPI_DACT;
      if (fast_form(&D, d->S, d->B, epi_b)) {
        pp.img[n + l] = buf; pp.per_iter[n + l] = (size_t)d->S * fast_img_floats(&D, d->S, d->B, epi_b);
        if (int e = fast_repack(&D, d->S, d->B, d->nit, d_wtraj, wss, d->P, epi_b, buf, st)) return e;
        buf += (size_t)d->nit * pp.per_iter[n + l];
      }
    }
  }
  return 0;
}

// side stream + events of the two-stream reverse pass, one set per device (created on first use, never destroyed)
struct DualStreams {
  cudaStream_t side;
  cudaEvent_t start, d_done[BSDE_MAX_LAYERS], w_done[BSDE_MAX_LAYERS];
};
static bool dual_enabled() {
  static const int on = [] { const char* e = getenv("BSDE_DUAL"); return e ? atoi(e) : 1; }();
  return on != 0;
}
// SM budget of each stream: blocks whose position grid is small (B * H * W <= BSDE_DUAL_MAXPOS) run their launch pairs side by side
// on half of the SMs each; larger ones keep full-size grids (the pair then overlaps only in its tails)
static int dual_budget(long long positions) {
  static const int b = [] { const char* e = getenv("BSDE_DUAL_BUDGET"); return e ? atoi(e) : 74; }();
  static const long long maxpos = [] { const char* e = getenv("BSDE_DUAL_MAXPOS"); return e ? atoll(e) : 0LL; }();
  return positions <= maxpos ? b : 148;
}
static DualStreams* dual_streams() {
  static DualStreams pool[64];
  static bool made[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!made[dev]) {
    DualStreams& q = pool[dev];
    if (cudaStreamCreateWithFlags(&q.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    bool ok = cudaEventCreateWithFlags(&q.start, cudaEventDisableTiming) == cudaSuccess;
    for (int l = 0; l < BSDE_MAX_LAYERS; ++l)
      ok = ok && cudaEventCreateWithFlags(&q.d_done[l], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&q.w_done[l], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return nullptr;
    made[dev] = true;
  }
  return &pool[dev];
}


// ---------------------------------------------------------------------------------------------
// Deferred, iteration-batched weight gradients (reverse pass with stored activations).
//
// The weight gradient of layer l at scan iteration k, G_k = sum over positions of in_k (x) dY_k, is independent of every other
// iteration, and at the bench sizes a weight-gradient launch is dominated by per-launch fixed cost (tools/bench_tconv.py --B
// 32..2048, profiles/r02_summary.md: 35-55 us of a 58-115 us call do not depend on the batch: split-K partial tiles of up to
// 187 KB per CTA, their reduction launch, prologue / tail).  With the activations kept per iteration (remat = False) and dZ / gx
// kept per iteration too, the (iteration, sample) pairs of a layer are laid out contiguously with ONE stride -- activations are
// stored layer-major [layer][k][s] -- so a run of c iterations is exactly a launch with S * c "samples": c = 74 / S iterations
// per launch give two CTAs per pair instead of 148 / S, a sixth of the partial tiles and a c-th of the launches.  The kernels are
// unchanged; they run with t = 1, scale = 1 into a [k][s][P] buffer, and one finishing pass applies the per-iteration factors
// (t_k on the time rows, dt_k on the last layer) while transposing to the [s][k][P] layout that K1' reads.
// ---------------------------------------------------------------------------------------------
constexpr int FIN_INLINE = 128;
struct FinalizeArgs {
  const float* src;     // [nit][S][P]
  float* dst;           // [S][nit][P]
  const float* tk;      // [nit] device (grids longer than FIN_INLINE)
  const float* dtk;     // [nit] device
  int inl;              // the time grid travels in the kernel parameters (no host -> device copy: a pageable copy stalls the host)
  float tkv[FIN_INLINE], dtkv[FIN_INLINE];
  int S, nit, nlayers;
  long long P;
  long long w_off[BSDE_MAX_LAYERS], w_end[BSDE_MAX_LAYERS], b_off[BSDE_MAX_LAYERS];
  int Kc[BSDE_MAX_LAYERS], cout[BSDE_MAX_LAYERS], time_row[BSDE_MAX_LAYERS];
};

// one thread per (parameter p, sample s): the factor's structure (last layer -> dt_k, time row -> t_k) is found once, then the
// nit iterations are streamed [k][s][P] -> [s][k][P] (coalesced in p on both sides; no per-element 64-bit divisions)
__global__ void gw_finalize_kernel(FinalizeArgs a) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (p >= a.P) return;
  bool use_dt = false, use_t = false;
  for (int l = 0; l < a.nlayers; ++l) {
    const bool in_w = p >= a.w_off[l] && p < a.w_end[l];
    const bool in_b = p >= a.b_off[l] && p < a.b_off[l] + a.cout[l];
    if (in_w || in_b) {
      use_dt = l == a.nlayers - 1;
      use_t = in_w && a.time_row[l] >= 0 && (int)(((p - a.w_off[l]) / a.cout[l]) % a.Kc[l]) == a.time_row[l];
      break;
    }
  }
  const float* __restrict__ src = a.src + (long long)s * a.P + p;
  float* __restrict__ dst = a.dst + (long long)s * a.nit * a.P + p;
  const long long sstep = (long long)a.S * a.P;
#pragma unroll 4
  for (int k = 0; k < a.nit; ++k) {
    float f = 1.f;
    if (use_dt) f *= a.inl ? a.dtkv[k] : a.dtk[k];
    if (use_t) f *= a.inl ? a.tkv[k] : a.tk[k];
    dst[(long long)k * a.P] = src[(long long)k * sstep] * f;
  }
}

static bool defer_enabled() {
  static const int on = [] { const char* e = getenv("BSDE_DEFER"); return e ? atoi(e) : 1; }();
  return on != 0;
}
// iterations per batched weight-gradient launch: two CTAs per (iteration, sample) pair on 148 SMs
static int defer_chunk(const bsde_xpath_desc* d) {
  static const int ctas = [] { const char* e = getenv("BSDE_DEFER_CTAS"); return e ? atoi(e) : 2; }();   // CTAs per pair (tools)
  return std::max(1, std::min(d->nit, 148 / (std::max(1, ctas) * d->S)));
}
// scratch of the batched weight-gradient launches: the largest over the chunk sizes that occur (c and the remainder nit % c:
// a smaller chunk gets MORE splits per pair, so its partial tiles can be the larger ones)
static size_t defer_wgrad_bytes(const bsde_xpath_desc* d) {
  const int c = defer_chunk(d);
  size_t wg = 0;
  for (int cnt : {c, d->nit % c}) {
    if (cnt < 1) continue;
    for (int l = 0; l < d->nlayers; ++l) wg = std::max(wg, wgrad_ws(&d->layers[l], d->S * cnt, d->B, d->math));
  }
  return wg;
}
// extra floats of the deferred mode: dZ of every iteration, the gx trajectory, the [k][s][P] gradient buffer, (t_k, dt_k)
static size_t defer_extra_floats(const bsde_xpath_desc* d) {
  size_t f = 0;
  for (int l = 0; l + 1 < d->nlayers; ++l) f += (size_t)d->nit * act_elems(d, l);
  f += (size_t)(d->nit + 1) * (size_t)d->x_numel + (size_t)d->nit * d->S * (size_t)d->P + 2 * (size_t)d->nit + 64;
  return f;
}
// Activations are stored layer-major [layer][k][s] (instead of [k][layer]) when the deferred mode applies: a function of the
// descriptor only, so that bsde_xpath_fwd (which writes them) and bsde_xpath_bwd (which reads them) always agree.
static bool xpath_layer_major(const bsde_xpath_desc* d) {
  if (!defer_enabled() || d->math != BSDE_MATH_TC || d->act == BSDE_ACT_SWISH || d->nit < 2 || d->S > 37) return false;
  for (int l = 0; l + 1 < d->nlayers; ++l)
    if ((act_elems(d, l) * sizeof(float)) % 256 != 0) return false;          // slots must be exactly S * per-sample floats
  for (int l = 0; l < d->nlayers; ++l) {                                     // dense HWIO leaves, time row last (JAX front end)
    const bsde_conv_layer& L = d->layers[l];
    const int Kc = L.cin + (L.has_time ? 1 : 0);
    if (L.w_n_stride != 1 || L.w_k_stride != L.cout || L.w_tap_stride != (long long)Kc * L.cout || (L.has_time && L.time_first)) return false;
  }
  static const double cap_gb = [] { const char* e = getenv("BSDE_DEFER_MAX_GB"); return e ? atof(e) : 48.0; }();
  return (double)defer_extra_floats(d) * 4.0 <= cap_gb * 1e9;
}
static size_t act_base_floats(const bsde_xpath_desc* d, int l) {  // layer-major: start of layer l's [nit][S][...] block
  size_t off = 0;
  for (int j = 0; j < l; ++j) off += (size_t)d->nit * act_elems(d, j);
  return off;
}

extern "C" size_t bsde_xpath_workspace_bytes(const bsde_xpath_desc* d, int32_t backward) {
  if (!d || d->nlayers < 2 || d->nlayers > BSDE_MAX_LAYERS) return 0;
  size_t tot = xpath_tc_bytes(d, backward);
  for (int l = 0; l + 1 < d->nlayers; ++l) tot += act_slot_bytes(d, l) * (backward ? 2 : 1);
  if (backward) {
    size_t wg = 0;
    for (int l = 0; l < d->nlayers; ++l) {
      size_t b = wgrad_ws(&d->layers[l], d->S, d->B, d->math);
      if (b > wg) wg = b;
    }
    wg = std::max(wg, xpath_swish_bytes(d));
    tot += align_up(wg);
  }
  tot += align_up(prepack_floats(d, backward, true) * sizeof(float));
  if (backward == 2 && xpath_layer_major(d)) {   // reverse pass with stored activations: deferred weight gradients
    tot += align_up(defer_wgrad_bytes(d)) + align_up(defer_extra_floats(d) * sizeof(float));
  }
  return tot;
}

static size_t xpath_acts_floats(const bsde_xpath_desc* d) {  // stored layer outputs of one scan iteration
  size_t tot = 0;
  for (int l = 0; l + 1 < d->nlayers; ++l) tot += act_slot_bytes(d, l) / sizeof(float);
  return tot;
}

extern "C" size_t bsde_xpath_acts_bytes(const bsde_xpath_desc* d) {
  if (!d || d->nlayers < 2 || d->nlayers > BSDE_MAX_LAYERS) return 0;
  return xpath_acts_floats(d) * sizeof(float) * (size_t)d->nit;
}

static int xpath_recompute(const bsde_xpath_desc* d, const float* x, const float* wk, float t, float* const* a,
                           void* tc_ws, size_t tc_bytes, cudaStream_t st, const Prepack& pp, int k, int& nconv) {
  const int n = d->nlayers;
  const long long wss = (long long)(d->nit + 1) * d->P;
  const float* in = x;
  const bool sw = d->act == BSDE_ACT_SWISH;
  for (int l = 0; l + 1 < n; ++l) {
    // swish: the conv writes the pre-activation z_l (bias + time term, no activation), one elementwise pass applies beta's gate
    if (int e = conv_fwd(&d->layers[l], d->S, d->B, in, wk, wss, t, sw ? BSDE_ACT_NONE : d->act, sw ? BSDE_EPI_BIAS : BSDE_EPI_BIAS_ACT,
                         1.f, nullptr, sw ? swish_z(d, l, a[l]) : a[l], d->math, tc_ws, tc_bytes, st,
                         pp.img[l] ? pp.img[l] + (size_t)k * pp.per_iter[l] : nullptr, nconv++ > 0))
      return e;
    if (sw) {
      const bsde_conv_layer& L = d->layers[l];
      if (int e = swish_forward(swish_z(d, l, a[l]), wk, wss, d->beta_off[l], d->S, (long long)d->B * L.hout * L.wout * L.cout, L.cout,
                                a[l], st))
        return e;
    }
    in = a[l];
  }
  return 0;
}

extern "C" int bsde_xpath_fwd(const bsde_xpath_desc* d, const float* h_tk, const float* h_dtk,
                              const float* d_wtraj, float* d_xs, float* d_acts, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (int e = xpath_check(d)) return e;
  BSDE_CHECK_ARG(h_tk && h_dtk && d_wtraj && d_xs, "NULL pointer");
  const size_t need = bsde_xpath_workspace_bytes(d, 0);
  if (!workspace || workspace_bytes < need) {
    set_error("xpath_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n = d->nlayers;
  float* a[BSDE_MAX_LAYERS];
  char* ws = (char*)workspace;
  void* tc_ws = ws;
  const size_t tc_bytes = xpath_tc_bytes(d, 0);
  ws += tc_bytes;
  for (int l = 0; l + 1 < n; ++l) { a[l] = (float*)ws; ws += act_slot_bytes(d, l); }
  const long long wss = (long long)(d->nit + 1) * d->P;
  const size_t acts_step = xpath_acts_floats(d);
  const bool layer_major = xpath_layer_major(d);
  Prepack pp;
  if (int e = prepack_all(d, 0, true, d_wtraj, (float*)ws, pp, st)) return e;
  int nconv = 0;  // conv launches since the prepack: the first one must not read the images before its pdl_wait()
  for (int k = 0; k < d->nit; ++k) {
    if (d_acts && layer_major) {  // [layer][k][s]: every layer's (iteration, sample) pairs are contiguous with one stride
      for (int l = 0; l + 1 < n; ++l) a[l] = d_acts + act_base_floats(d, l) + (size_t)k * act_elems(d, l);
    } else if (d_acts) {  // keep this iteration's layer outputs for the reverse pass (remat = False)
      float* p = d_acts + (size_t)k * acts_step;
      for (int l = 0; l + 1 < n; ++l) { a[l] = p; p += act_slot_bytes(d, l) / sizeof(float); }
    }
    const float* xk = d_xs + (size_t)k * d->x_numel;
    float* xn = d_xs + (size_t)(k + 1) * d->x_numel;
    const float* wk = d_wtraj + (size_t)k * d->P;
    if (int e = xpath_recompute(d, xk, wk, h_tk[k], a, tc_ws, tc_bytes, st, pp, k, nconv)) return e;
    // last conv + Euler update: x_{k+1} = x_k + dt_k (conv(a_{n-2}) + b)   (sdeint.py:48-50, g_x = 0)
    if (int e = conv_fwd(&d->layers[n - 1], d->S, d->B, a[n - 2], wk, wss, h_tk[k], d->act, BSDE_EPI_EULER,
                         h_dtk[k], xk, xn, d->math, tc_ws, tc_bytes, st,
                         pp.img[n - 1] ? pp.img[n - 1] + (size_t)k * pp.per_iter[n - 1] : nullptr, nconv++ > 0))
      return e;
  }
  return 0;
}


static int xpath_bwd_deferred(const bsde_xpath_desc* d, const float* h_tk, const float* h_dtk, const float* d_wtraj, const float* d_xs,
                              const float* d_acts, float* d_gx, float* d_gw_traj, void* workspace, cudaStream_t st) {
  const int n = d->nlayers, nit = d->nit, S = d->S;
  char* ws = (char*)workspace;
  void* tc_ws = ws;
  const size_t tc_bytes = xpath_tc_bytes(d, 1);
  ws += tc_bytes;
  // (the per-iteration slots of the other mode stay unused: the layout of the common part is shared with bsde_xpath_workspace_bytes)
  for (int l = 0; l + 1 < n; ++l) ws += act_slot_bytes(d, l) * 2;
  {
    size_t wgmax = 0;
    for (int l = 0; l < n; ++l) wgmax = std::max(wgmax, wgrad_ws(&d->layers[l], S, d->B, d->math));
    wgmax = std::max(wgmax, xpath_swish_bytes(d));
    ws += align_up(wgmax);
  }
  Prepack pp;
  if (int e = prepack_all(d, 1, false, d_wtraj, (float*)ws, pp, st)) return e;
  ws += align_up(prepack_floats(d, 1, true) * sizeof(float));
  const int c = defer_chunk(d);
  void* wg_ws = ws;
  const size_t wg_bytes = align_up(defer_wgrad_bytes(d));
  ws += wg_bytes;
  float* fp = (float*)ws;
  float* dzs[BSDE_MAX_LAYERS];
  for (int l = 0; l + 1 < n; ++l) { dzs[l] = fp; fp += (size_t)nit * act_elems(d, l); }
  float* gxs = fp; fp += (size_t)(nit + 1) * (size_t)d->x_numel;
  float* gw_ks = fp; fp += (size_t)nit * S * (size_t)d->P;
  float* d_tk = fp; float* d_dtk = fp + nit;

  if (nit > FIN_INLINE) {
    BSDE_CUDA_OK(cudaMemcpyAsync(d_tk, h_tk, nit * sizeof(float), cudaMemcpyHostToDevice, st));
    BSDE_CUDA_OK(cudaMemcpyAsync(d_dtk, h_dtk, nit * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  BSDE_CUDA_OK(cudaMemcpyAsync(gxs + (size_t)nit * d->x_numel, d_gx, (size_t)d->x_numel * sizeof(float), cudaMemcpyDeviceToDevice, st));

  const long long wss = (long long)(nit + 1) * d->P;
  bsde_conv_layer D[BSDE_MAX_LAYERS];
  for (int l = 0; l < n; ++l) D[l] = dgrad_layer(d->layers[l]);
  DualStreams* ds = dual_enabled() ? dual_streams() : nullptr;
  cudaStream_t wst = ds ? ds->side : st;
  int nconv = 0;
  int k_hi = nit - 1;   // top of the chunk whose weight gradients are still to be launched

  for (int k = nit - 1; k >= 0; --k) {
    const float t = h_tk[k], dt = h_dtk[k];
    const float* wk = d_wtraj + (size_t)k * d->P;
    const float* gx_in = gxs + (size_t)(k + 1) * d->x_numel;
    float* gx_out = k == 0 ? d_gx : gxs + (size_t)k * d->x_numel;
    const float* dy = gx_in;
    for (int l = n - 1; l >= 1; --l) {
      const float sc = (l == n - 1) ? dt : 1.f;
      const float* al = d_acts + act_base_floats(d, l - 1) + (size_t)k * act_elems(d, l - 1);
      float* out = dzs[l - 1] + (size_t)k * act_elems(d, l - 1);
      if (int e = conv_fwd(&D[l], S, d->B, dy, wk, wss, t, d->act, BSDE_EPI_DACT, sc, al, out, d->math, tc_ws, tc_bytes, st,
                           pp.img[n + l] ? pp.img[n + l] + (size_t)k * pp.per_iter[n + l] : nullptr, nconv++ > 0))
        return e;
      dy = out;
    }
    // gx_k = gx_{k+1} + dgrad_0(dz_0)   (identity path of the Euler update); the trajectory of gx is kept for the last layer's weight gradient
    if (int e = conv_fwd(&D[0], S, d->B, dy, wk, wss, t, d->act, BSDE_EPI_ACCUM, n == 1 ? dt : 1.f, gx_in, gx_out, d->math, tc_ws, tc_bytes, st,
                         pp.img[n] ? pp.img[n] + (size_t)k * pp.per_iter[n] : nullptr, nconv++ > 0))
      return e;
    // a chunk of c iterations is complete: its weight gradients, one launch per layer over S * c (iteration, sample) pairs
    if (k_hi - k + 1 == c || k == 0) {
      const int k_lo = k, cnt = k_hi - k_lo + 1;
      if (ds) {
        BSDE_CUDA_OK(cudaEventRecord(ds->start, st));
        BSDE_CUDA_OK(cudaStreamWaitEvent(ds->side, ds->start, 0));
      }
      for (int l = n - 1; l >= 0; --l) {
        const float* lin = l == 0 ? d_xs + (size_t)k_lo * d->x_numel : d_acts + act_base_floats(d, l - 1) + (size_t)k_lo * act_elems(d, l - 1);
        const float* ldy = l == n - 1 ? gxs + (size_t)(k_lo + 1) * d->x_numel : dzs[l] + (size_t)k_lo * act_elems(d, l);
        if (int e = conv_wgrad(&d->layers[l], S * cnt, d->B, lin, ldy, 1.f, 1.f, gw_ks + (size_t)k_lo * S * d->P, d->P, d->math, wg_ws, wg_bytes, wst))
          return e;
      }
      k_hi = k - 1;
    }
  }
  if (ds) {
    BSDE_CUDA_OK(cudaEventRecord(ds->start, ds->side));
    BSDE_CUDA_OK(cudaStreamWaitEvent(st, ds->start, 0));
  }
  FinalizeArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.src = gw_ks; fa.dst = d_gw_traj; fa.tk = d_tk; fa.dtk = d_dtk; fa.S = S; fa.nit = nit; fa.nlayers = n; fa.P = d->P;
  fa.inl = nit <= FIN_INLINE;
  if (fa.inl) { for (int k = 0; k < nit; ++k) { fa.tkv[k] = h_tk[k]; fa.dtkv[k] = h_dtk[k]; } }
  for (int l = 0; l < n; ++l) {
    const bsde_conv_layer& L = d->layers[l];
    const int Kc = L.cin + (L.has_time ? 1 : 0);
    fa.w_off[l] = L.w_off; fa.w_end[l] = L.w_off + (long long)L.ksize * L.ksize * Kc * L.cout; fa.b_off[l] = L.b_off;
    fa.Kc[l] = Kc; fa.cout[l] = L.cout; fa.time_row[l] = L.has_time ? L.cin : -1;
  }
  gw_finalize_kernel<<<dim3((unsigned)((d->P + 255) / 256), (unsigned)S), 256, 0, st>>>(fa);
  BSDE_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int bsde_xpath_bwd(const bsde_xpath_desc* d, const float* h_tk, const float* h_dtk,
                              const float* d_wtraj, const float* d_xs, const float* d_acts, float* d_gx,
                              float* d_gw_traj, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = xpath_check(d)) return e;
  BSDE_CHECK_ARG(h_tk && h_dtk && d_wtraj && d_xs && d_gx && d_gw_traj, "NULL pointer");
  const bool deferred = d_acts != nullptr && xpath_layer_major(d);
  const size_t need = bsde_xpath_workspace_bytes(d, deferred ? 2 : 1);
  if (!workspace || workspace_bytes < need) {
    set_error("xpath_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  if (deferred) return xpath_bwd_deferred(d, h_tk, h_dtk, d_wtraj, d_xs, d_acts, d_gx, d_gw_traj, workspace, (cudaStream_t)stream);
  cudaStream_t st = (cudaStream_t)stream;
  const int n = d->nlayers;
  float* a[BSDE_MAX_LAYERS];
  float* dz[BSDE_MAX_LAYERS];
  char* ws = (char*)workspace;
  void* tc_ws = ws;
  const size_t tc_bytes = xpath_tc_bytes(d, 1);
  ws += tc_bytes;
  for (int l = 0; l + 1 < n; ++l) { a[l] = (float*)ws; ws += act_slot_bytes(d, l); }
  for (int l = 0; l + 1 < n; ++l) { dz[l] = (float*)ws; ws += align_up(act_elems(d, l) * sizeof(float)); }
  void* wg_ws = ws;
  size_t wgmax = 0;
  for (int l = 0; l < n; ++l) wgmax = std::max(wgmax, wgrad_ws(&d->layers[l], d->S, d->B, d->math));
  wgmax = std::max(wgmax, xpath_swish_bytes(d));
  const size_t wg_bytes = align_up(wgmax);
  ws += wg_bytes;
  Prepack pp;
  if (int e = prepack_all(d, 1, d_acts == nullptr, d_wtraj, (float*)ws, pp, st)) return e;
  int nconv = 0;
  const long long wss = (long long)(d->nit + 1) * d->P;
  const long long gss = (long long)d->nit * d->P;

  bsde_conv_layer D[BSDE_MAX_LAYERS];
  for (int l = 0; l < n; ++l) D[l] = dgrad_layer(d->layers[l]);

  // Two-stream reverse pass (tensor-core route): the weight gradient of layer l and the data gradient of layer l read the same dY
  // and are independent of each other, and every kernel on this path is a persistent launch with ONE CTA per SM whose time at
  // these sizes is dominated by per-launch latencies (prologue, weight image, tail) rather than by throughput.  The weight
  // gradients go to a side stream, both streams size their grids for half of the SMs (sm_budget), so the pair is co-resident.
  // Ordering by events: wgrad_l(k) after the launch that produced its dY; the launch that overwrites a buffer after its readers.
  DualStreams* ds = nullptr;
  if (d->math == BSDE_MATH_TC && d->act != BSDE_ACT_SWISH && dual_enabled()) ds = dual_streams();
  bool have_w[BSDE_MAX_LAYERS] = {false, false, false, false};
  if (ds) {
    BSDE_CUDA_OK(cudaEventRecord(ds->start, st));
    BSDE_CUDA_OK(cudaStreamWaitEvent(ds->side, ds->start, 0));
    set_sm_budget(dual_budget((long long)d->B * d->layers[0].hin * d->layers[0].win));
  }
  struct BudgetReset { bool on; ~BudgetReset() { if (on) set_sm_budget(148); } } budget_reset{ds != nullptr};
  cudaStream_t wst = ds ? ds->side : st;

  for (int k = d->nit - 1; k >= 0; --k) {
    const float t = h_tk[k], dt = h_dtk[k];
    const float* xk = d_xs + (size_t)k * d->x_numel;
    const float* wk = d_wtraj + (size_t)k * d->P;
    float* gwk = d_gw_traj + (size_t)k * d->P;
    if (d_acts) {
      float* p = const_cast<float*>(d_acts) + (size_t)k * xpath_acts_floats(d);
      for (int l = 0; l + 1 < n; ++l) { a[l] = p; p += act_slot_bytes(d, l) / sizeof(float); }
    } else {
      if (ds)   // the recompute overwrites the activations the previous iteration's weight gradients read
        for (int l = 0; l < n; ++l)
          if (have_w[l]) BSDE_CUDA_OK(cudaStreamWaitEvent(st, ds->w_done[l], 0));
      if (int e = xpath_recompute(d, xk, wk, t, a, tc_ws, tc_bytes, st, pp, k, nconv)) return e;
      if (ds) {
        BSDE_CUDA_OK(cudaEventRecord(ds->start, st));
        BSDE_CUDA_OK(cudaStreamWaitEvent(ds->side, ds->start, 0));
      }
    }
    // walk the stack backwards; `dy` is dL/d(conv output of layer l)
    const float* dy = d_gx;  // for the last layer dL/d(conv out) = dt * gx  (scale folded below)
    for (int l = n - 1; l >= 0; --l) {
      const float* lin = l == 0 ? xk : a[l - 1];
      const float sc = (l == n - 1) ? dt : 1.f;
      if (ds) BSDE_CUDA_OK(cudaStreamWaitEvent(ds->side, ds->d_done[l], 0));   // dY of layer l is complete (gx: the previous dgrad_0)
      if (int e = conv_wgrad(&d->layers[l], d->S, d->B, lin, dy, t, sc, gwk, gss, d->math, wg_ws, wg_bytes, wst)) return e;
      if (ds) { BSDE_CUDA_OK(cudaEventRecord(ds->w_done[l], ds->side)); have_w[l] = true; }
      if (l > 0) {
        // dz_{l-1} = sc * dgrad(dy) * act'(a_{l-1})
        const bool sw = d->act == BSDE_ACT_SWISH;
        if (ds && have_w[l - 1]) BSDE_CUDA_OK(cudaStreamWaitEvent(st, ds->w_done[l - 1], 0));   // readers of dz_{l-1} (previous iteration)
        if (int e = conv_fwd(&D[l], d->S, d->B, dy, wk, wss, t, sw ? BSDE_ACT_NONE : d->act, BSDE_EPI_DACT, sc, a[l - 1], dz[l - 1],
                             d->math, tc_ws, tc_bytes, st, pp.img[n + l] ? pp.img[n + l] + (size_t)k * pp.per_iter[n + l] : nullptr, nconv++ > 0))
          return e;
        if (sw) {  // dz <- dz * da/dz(z, beta) and dL/dbeta (beta is part of w: arch.py:172-174), from the stored pre-activation
          const bsde_conv_layer& Lp = d->layers[l - 1];
          if (int e = swish_backward(dz[l - 1], swish_z(d, l - 1, a[l - 1]), wk, wss, d->beta_off[l - 1], d->S,
                                     (long long)d->B * Lp.hout * Lp.wout * Lp.cout, Lp.cout, 1.f, (float*)wg_ws, gwk, gss, st))
            return e;
        }
        if (ds) BSDE_CUDA_OK(cudaEventRecord(ds->d_done[l - 1], st));
        dy = dz[l - 1];
      } else {
        // gx_k = gx_{k+1} + dgrad_0(dz_0)   (identity path of the Euler update)
        if (ds) BSDE_CUDA_OK(cudaStreamWaitEvent(st, ds->w_done[n - 1], 0));   // the last layer's weight gradient reads gx_{k+1}
        if (int e = conv_fwd(&D[0], d->S, d->B, dy, wk, wss, t, d->act, BSDE_EPI_ACCUM, sc, d_gx, d_gx,
                             d->math, tc_ws, tc_bytes, st, pp.img[n] ? pp.img[n] + (size_t)k * pp.per_iter[n] : nullptr, nconv++ > 0))
          return e;
        if (ds) BSDE_CUDA_OK(cudaEventRecord(ds->d_done[n - 1], st));
      }
    }
  }
  if (ds) {   // the caller's stream continues after the last weight gradient
    BSDE_CUDA_OK(cudaEventRecord(ds->start, ds->side));
    BSDE_CUDA_OK(cudaStreamWaitEvent(st, ds->start, 0));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Feed-forward drift net of the torch front end (make_y_net, torchbnn/_impl/models.py:25-54): ONE evaluation
// f = y_net(t, x; w) of a chain of time-dependent convs, and its reverse pass (data gradients + per-layer weight
// gradients).  The fixed-step solver (midpoint: two evaluations per step) drives these from the host side.
// ---------------------------------------------------------------------------------------------
static int chain_check(const bsde_chain_desc* d) {
  BSDE_CHECK_ARG(d != nullptr, "NULL descriptor");
  BSDE_CHECK_ARG(d->nlayers >= 1 && d->nlayers <= BSDE_CHAIN_MAX_LAYERS, "nlayers=%d unsupported", d->nlayers);
  BSDE_CHECK_ARG(d->S >= 1 && d->B >= 1, "S, B must be positive");
  BSDE_CHECK_ARG(d->act >= BSDE_ACT_SOFTPLUS && d->act <= BSDE_ACT_ELU, "activation %d unsupported", d->act);
  for (int l = 0; l + 1 < d->nlayers; ++l)
    BSDE_CHECK_ARG(d->layers[l].cout == d->layers[l + 1].cin && d->layers[l].hout == d->layers[l + 1].hin &&
                       d->layers[l].wout == d->layers[l + 1].win,
                   "layer %d output does not match layer %d input", l, l + 1);
  if (d->act_after[d->nlayers - 1]) {
    set_error("an activation after the last layer of the chain is not supported (models.py:35: last_activation=False there)");
    return BSDE_ERR_UNSUPPORTED;
  }
  return 0;
}

static size_t chain_act_elems(const bsde_chain_desc* d, int l) {
  const bsde_conv_layer& L = d->layers[l];
  return (size_t)d->S * d->B * L.hout * L.wout * L.cout;
}

static size_t chain_act_offset(const bsde_chain_desc* d, int l) {  // floats
  size_t off = 0;
  for (int i = 0; i < l; ++i) off += align_up(chain_act_elems(d, i) * sizeof(float)) / sizeof(float);
  return off;
}

static size_t chain_tc_bytes(const bsde_chain_desc* d, int backward) {
  if (d->math != BSDE_MATH_TC) return 0;
  size_t m = 0;
  for (int l = 0; l < d->nlayers; ++l) {
    m = std::max(m, conv_tc_ws(&d->layers[l], d->S));
    if (backward) {
      bsde_conv_layer D = dgrad_layer(d->layers[l]);
      m = std::max(m, conv_tc_ws(&D, d->S));
    }
  }
  return align_up(m);
}

static size_t chain_max_act_bytes(const bsde_chain_desc* d) {
  size_t m = 0;
  for (int l = 0; l < d->nlayers; ++l) m = std::max(m, align_up(chain_act_elems(d, l) * sizeof(float)));
  return m;
}

static size_t chain_wgrad_bytes(const bsde_chain_desc* d) {
  size_t m = 0;
  for (int l = 0; l < d->nlayers; ++l) m = std::max(m, wgrad_ws(&d->layers[l], d->S, d->B, d->math));
  return align_up(m);
}

extern "C" size_t bsde_chain_acts_floats(const bsde_chain_desc* d) {
  if (!d || d->nlayers < 1 || d->nlayers > BSDE_CHAIN_MAX_LAYERS) return 0;
  return chain_act_offset(d, d->nlayers);
}

extern "C" size_t bsde_chain_out_offset(const bsde_chain_desc* d) {
  if (!d || d->nlayers < 1 || d->nlayers > BSDE_CHAIN_MAX_LAYERS) return 0;
  return chain_act_offset(d, d->nlayers - 1);
}

extern "C" size_t bsde_chain_workspace_bytes(const bsde_chain_desc* d, int32_t backward) {
  if (!d || d->nlayers < 1 || d->nlayers > BSDE_CHAIN_MAX_LAYERS) return 0;
  size_t tot = chain_tc_bytes(d, backward);
  if (backward) tot += 2 * chain_max_act_bytes(d) + chain_wgrad_bytes(d);
  return tot + 256;
}

extern "C" int bsde_chain_fwd(const bsde_chain_desc* d, float t, const float* d_w, int64_t w_sample_stride, const float* d_x,
                              float* d_acts, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = chain_check(d)) return e;
  BSDE_CHECK_ARG(d_w && d_x && d_acts, "NULL pointer");
  const size_t need = bsde_chain_workspace_bytes(d, 0);
  if (!workspace || workspace_bytes < need) {
    set_error("chain_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  const size_t tc_bytes = chain_tc_bytes(d, 0);
  const float* in = d_x;
  size_t off = 0;
  for (int l = 0; l < d->nlayers; ++l) {
    float* out = d_acts + off;
    if (int e = conv_fwd(&d->layers[l], d->S, d->B, in, d_w, w_sample_stride, t, d->act,
                         d->act_after[l] ? BSDE_EPI_BIAS_ACT : BSDE_EPI_BIAS, 1.f, nullptr, out, d->math, workspace, tc_bytes,
                         (cudaStream_t)stream))
      return e;
    in = out;
    off += align_up(chain_act_elems(d, l) * sizeof(float)) / sizeof(float);
  }
  return 0;
}

extern "C" int bsde_chain_bwd(const bsde_chain_desc* d, float t, const float* d_w, int64_t w_sample_stride, const float* d_x,
                              const float* d_acts, const float* d_gf, float* d_gx, int32_t accumulate_gx, float* d_gw,
                              int64_t gw_sample_stride, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = chain_check(d)) return e;
  BSDE_CHECK_ARG(d_w && d_x && d_acts && d_gf && d_gx && d_gw, "NULL pointer");
  const size_t need = bsde_chain_workspace_bytes(d, 1);
  if (!workspace || workspace_bytes < need) {
    set_error("chain_bwd: workspace %zu < %zu bytes", workspace_bytes, need);
    return BSDE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  void* tc_ws = ws;
  const size_t tc_bytes = chain_tc_bytes(d, 1);
  ws += tc_bytes;
  const size_t mab = chain_max_act_bytes(d);
  float* dz[2] = {(float*)ws, (float*)(ws + mab)};
  ws += 2 * mab;
  void* wg_ws = ws;
  const size_t wg_bytes = chain_wgrad_bytes(d);
  const int n = d->nlayers;
  const float* dy = d_gf;  // dL/d(conv output of layer l)
  for (int l = n - 1; l >= 0; --l) {
    const float* lin = l == 0 ? d_x : d_acts + chain_act_offset(d, l - 1);
    if (int e = conv_wgrad(&d->layers[l], d->S, d->B, lin, dy, t, 1.f, d_gw, gw_sample_stride, d->math, wg_ws, wg_bytes, st)) return e;
    bsde_conv_layer D = dgrad_layer(d->layers[l]);
    if (l > 0) {
      float* out = dz[l & 1];
      const int epi = d->act_after[l - 1] ? BSDE_EPI_DACT : BSDE_EPI_BIAS;  // D.b_off < 0: BIAS is the bare accumulator
      if (int e = conv_fwd(&D, d->S, d->B, dy, d_w, w_sample_stride, t, d->act, epi, 1.f, lin, out, d->math, tc_ws, tc_bytes, st))
        return e;
      dy = out;
    } else {
      const int epi = accumulate_gx ? BSDE_EPI_ACCUM : BSDE_EPI_BIAS;
      if (int e = conv_fwd(&D, d->S, d->B, dy, d_w, w_sample_stride, t, d->act, epi, 1.f, d_gx, d_gx, d->math, tc_ws, tc_bytes, st))
        return e;
    }
  }
  return 0;
}

// The torch front end adds the drift to the state in FLAT NCHW ORDER (models.py:163,168: `fy.reshape(-1)`): the
// net's output (B, C2, H2, W2) is reinterpreted as the state's (B, C, H, W) per image.  Activations live in NHWC
// storage here, so that reinterpretation is an index permutation:  out[n][h][w][c] = base + a * src[n][h2][w2][c2]
// with  c*H*W + h*W + w  ==  c2*H2*W2 + h2*W2 + w2.   Swapping the two shapes gives the transpose (reverse pass).
namespace bsde {
__global__ void nchw_reinterp_axpy_kernel(long long total, int C, int H, int W, int C2, int H2, int W2, float a,
                                          const float* __restrict__ base, const float* __restrict__ src, float* __restrict__ out) {
  const int chw = C * H * W;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
    const long long n = o / chw;
    const int r = (int)(o - n * chw);
    const int c = r % C, pix = r / C;       // NHWC: r = (h*W + w)*C + c
    const int i = c * (H * W) + pix;        // flat NCHW index inside the image
    const int c2 = i / (H2 * W2), pix2 = i - c2 * (H2 * W2);
    const float v = a * __ldg(src + n * chw + (long long)pix2 * C2 + c2);
    out[o] = base ? base[o] + v : v;
  }
}
}  // namespace bsde

extern "C" int bsde_nchw_reinterp_axpy(int64_t nimg, int32_t C, int32_t H, int32_t W, int32_t C2, int32_t H2, int32_t W2, float a,
                                       const float* d_base, const float* d_src, float* d_out, void* stream) {
  BSDE_CHECK_ARG(nimg >= 0 && C >= 1 && H >= 1 && W >= 1 && C2 >= 1 && H2 >= 1 && W2 >= 1, "bad shape");
  BSDE_CHECK_ARG((int64_t)C * H * W == (int64_t)C2 * H2 * W2, "the two shapes must have the same number of elements per image");
  if (nimg == 0) return 0;
  BSDE_CHECK_ARG(d_src && d_out, "NULL tensor pointer");
  const long long total = (long long)nimg * C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  nchw_reinterp_axpy_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(total, C, H, W, C2, H2, W2, a, d_base, d_src, d_out);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// generic flat-state step (jaxsde/jaxsde/sdeint.py:36-50) and Philox increments
// ---------------------------------------------------------------------------------------------
namespace bsde {
__global__ void step_update_kernel(int method, long long n, float dt, const float* __restrict__ y,
                                   const float* __restrict__ f, const float* __restrict__ g,
                                   const float* __restrict__ dW, const float* __restrict__ gdg,
                                   const float* __restrict__ g2, float* __restrict__ y1) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = dW[i];
    float v = y[i] + dt * f[i];
    if (method == BSDE_EULER_HEUN && g2) v += 0.5f * (g[i] + g2[i]) * w;
    else v += g[i] * w;
    if (method == BSDE_MILSTEIN && gdg) v += 0.5f * gdg[i] * (w * w - dt);
    y1[i] = v;
  }
}

__global__ void philox_increments_kernel(uint64_t seed, uint32_t stream_id, int S, int nit, long long P,
                                         const float* __restrict__ dtk, float* __restrict__ out) {
  const long long total = (long long)S * nit * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % P;
    const long long r = i / P;
    const int k = (int)(r % nit), s = (int)(r / nit);
    float z4[4];   // the K1 increments stream: one Philox call per four samples (common.cuh: philox_normal4)
    philox_normal4(seed, (uint32_t)p, (uint32_t)k, (uint32_t)s >> 2, stream_id, z4);
    const int zc = s & 3;
    out[i] = sqrtf(fmaxf(dtk[k], 0.f)) * (zc == 0 ? z4[0] : zc == 1 ? z4[1] : zc == 2 ? z4[2] : z4[3]);
  }
}
}  // namespace bsde

// ---------------------------------------------------------------------------------------------
// Virtual Brownian tree (jaxsde/jaxsde/brownian.py:19-95) and the product-form step of the stochastic adjoint
// (jaxsde/jaxsde/sdeint.py:36-50 with g_prod / gdg_prod given as vectors, sde_vjp.py:13-47)
// ---------------------------------------------------------------------------------------------
namespace bsde {
__global__ void brownian_tree_kernel(uint64_t seed, uint32_t stream_id, long long n, int rep, int depth, float t0, float t1,
                                     float t, float* __restrict__ out) {
  const float span = t1 - t0;
  const float s = (t - t0) / span;  // scaled_brownian_motion, brownian.py:12-17
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // sample_with_rep (brownian.py:69-74): the last `rep` entries reuse the normals of the `rep` entries before them
    const uint32_t e = (uint32_t)((rep > 0 && i >= n - rep) ? i - rep : i);
    float tl = 0.f, xl = 0.f, tr = 1.f;
    float xr = philox_normal(seed, e, 0u, 0u, stream_id);  // B(1): make_standard_brownian_motion, brownian.py:77-88
    uint32_t node = 1u;                                     // heap index <-> the left / right key of random.split
    for (int lvl = 0; lvl < depth; ++lvl) {                 // virtual_brownian_tree, brownian.py:26-51
      const float tm = (tl + tr) * 0.5f;
      const float z = philox_normal(seed, e, node, 1u, stream_id);
      const float xm = (xl * (tr - tm) + xr * (tm - tl)) / (tr - tl) + sqrtf((tr - tm) * (tm - tl) / (tr - tl)) * z;  // :19-23
      if (s < tm) { tr = tm; xr = xm; node = 2u * node; }
      else { tl = tm; xl = xm; node = 2u * node + 1u; }
    }
    const float z = philox_normal(seed, e, node, 1u, stream_id);
    const float mean = (xl * (tr - s) + xr * (s - tl)) / (tr - tl);
    const float sd = sqrtf(fmaxf((tr - s) * (s - tl), 0.f) / (tr - tl));
    out[i] = sqrtf(span) * (mean + sd * z);
  }
}

__global__ void step_update_prod_kernel(long long n, float dt, const float* __restrict__ y, const float* __restrict__ f,
                                        const float* __restrict__ gp, const float* __restrict__ mp, float* __restrict__ y1) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = y[i] + dt * f[i] + gp[i];
    if (mp) v += 0.5f * mp[i];
    y1[i] = v;
  }
}
}  // namespace bsde

extern "C" int bsde_brownian_tree(uint64_t seed, uint32_t stream_id, int64_t n, int32_t rep, int32_t depth, float t0, float t1,
                                  float t, float* d_out, void* stream) {
  BSDE_CHECK_ARG(n >= 0 && rep >= 0 && 2 * (int64_t)rep <= n, "need 0 <= 2*rep <= n (n=%lld, rep=%d)", (long long)n, rep);
  BSDE_CHECK_ARG(depth >= 0 && depth <= 30, "depth must be in [0, 30], got %d", depth);
  BSDE_CHECK_ARG(t1 > t0, "need t1 > t0");
  BSDE_CHECK_ARG(n < (1LL << 32), "n must be < 2^32");
  if (n == 0) return 0;
  BSDE_CHECK_ARG(d_out != nullptr, "NULL tensor pointer");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  brownian_tree_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seed, stream_id, n, rep, depth, t0, t1, t, d_out);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// The same tree on JAX's threefry PRNG: jaxsde/jaxsde/brownian.py:19-95 with jax.random.{split,normal}
// (jax==0.1.77, requirements.txt:1; Threefry-2x32-20, uniform from the top 23 bits, XLA's float32 erf_inv) so that the path a JAX
// run saw can be queried here.  The keys along the root-to-leaf walk are the same for every element: the host derives them
// (22 Threefry blocks), the kernel draws one normal per element and level.
// ---------------------------------------------------------------------------------------------
namespace bsde {
#define BSDE_TF_MAX_DEPTH 30
struct TfPath {
  uint32_t end_key[2];
  uint32_t lvl_key[BSDE_TF_MAX_DEPTH + 1][2];  // key of the bridge draw at level l; [depth] = the final bridge point
};

__host__ __device__ __forceinline__ uint32_t tf_rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0]; x1 += ks[1];
#define TF_R(r) x0 += x1; x1 = tf_rotl(x1, r) ^ x0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if (i % 2 == 0) { TF_R(13) TF_R(15) TF_R(26) TF_R(6) } else { TF_R(17) TF_R(29) TF_R(16) TF_R(24) }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
#undef TF_R
}

// jax.random.split(key): threefry over iota(4) = lanes (0, 2) and (1, 3); left = lane-0 outputs, right = lane-1 outputs
static void tf_split(const uint32_t key[2], uint32_t left[2], uint32_t right[2]) {
  uint32_t a0 = 0, a1 = 2, b0 = 1, b1 = 3;
  threefry2x32(key[0], key[1], a0, a1);
  threefry2x32(key[0], key[1], b0, b1);
  left[0] = a0; left[1] = b0; right[0] = a1; right[1] = b1;
}

// element e of jax.random.normal(key, (n,)) in float32: _random_bits puts counts [0, half) on lane 0 and [half, 2 half) on lane 1
// (one zero appended when n is odd); uniform on [nextafter(-1, 0), 1) from the top 23 bits; sqrt(2) * erf_inv (XLA / Giles 2010)
__device__ __forceinline__ float tf_normal(uint32_t k0, uint32_t k1, long long n, long long e) {
  const long long half = (n + 1) >> 1;
  uint32_t x0, x1;
  const bool lane1 = e >= half;
  if (!lane1) { x0 = (uint32_t)e; x1 = (e + half < n) ? (uint32_t)(e + half) : 0u; }
  else { x0 = (uint32_t)(e - half); x1 = (uint32_t)e; }
  threefry2x32(k0, k1, x0, x1);
  const uint32_t bits = lane1 ? x1 : x0;
  const float lo = -0.99999994f;                                              // nextafter(-1, 0)
  const float fl = __fsub_rn(__uint_as_float((bits >> 9) | 0x3F800000u), 1.0f);
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(fl, __fsub_rn(1.0f, lo)), lo));
  float w = -log1pf(__fmul_rn(-u, u));
  const bool lt = w < 5.0f;
  w = lt ? __fsub_rn(w, 2.5f) : __fsub_rn(sqrtf(w), 3.0f);
  float p = lt ? 2.81022636e-08f : -0.000200214257f;
#define TF_C(a, b) p = __fadd_rn(lt ? a : b, __fmul_rn(p, w));
  TF_C(3.43273939e-07f, 0.000100950558f) TF_C(-3.5233877e-06f, 0.00134934322f) TF_C(-4.39150654e-06f, -0.00367342844f)
  TF_C(0.00021858087f, 0.00573950773f) TF_C(-0.00125372503f, -0.0076224613f) TF_C(-0.00417768164f, 0.00943887047f)
  TF_C(0.246640727f, 1.00167406f) TF_C(1.50140941f, 2.83297682f)
#undef TF_C
  const float r = fabsf(u) == 1.0f ? u * INFINITY : __fmul_rn(p, u);
  return __fmul_rn(1.41421354f, r);
}

__global__ void brownian_tree_threefry_kernel(TfPath path, long long n, int rep, int depth, float t0, float t1, float t,
                                              long long first, long long count, float* __restrict__ out) {
  const float span = __fsub_rn(t1, t0);
  const float s = __fdiv_rn(__fsub_rn(t, t0), span);
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < count; j += (long long)gridDim.x * blockDim.x) {
    const long long i = first + j;
    const long long e = (rep > 0 && i >= n - rep) ? i - rep : i;               // sample_with_rep, brownian.py:69-74
    float tl = 0.f, xl = 0.f, tr = 1.f;
    float xr = tf_normal(path.end_key[0], path.end_key[1], n, e);            // w_end, brownian.py:81-82
    for (int lvl = 0; lvl < depth; ++lvl) {                                   // virtual_brownian_tree, brownian.py:26-51
      const float tm = __fdiv_rn(__fadd_rn(tl, tr), 2.0f);
      const float z = tf_normal(path.lvl_key[lvl][0], path.lvl_key[lvl][1], n, e);
      const float xm = (xl * (tr - tm) + xr * (tm - tl)) / (tr - tl) + sqrtf((tr - tm) * (tm - tl) / (tr - tl)) * z;
      if (s < tm) { tr = tm; xr = xm; } else { tl = tm; xl = xm; }
    }
    const float z = tf_normal(path.lvl_key[depth][0], path.lvl_key[depth][1], n, e);
    const float mean = (xl * (tr - s) + xr * (s - tl)) / (tr - tl);
    const float sd = sqrtf((tr - s) * (s - tl) / (tr - tl));
    out[j] = sqrtf(span) * (mean + sd * z);
  }
}
}  // namespace bsde

extern "C" int bsde_threefry_split(uint32_t key0, uint32_t key1, int32_t num, uint32_t* h_out) {
  BSDE_CHECK_ARG(num >= 1 && h_out, "need num >= 1 and a host output array of 2*num words");
  // jax.random.split(key, num): threefry of iota(2 num), first half on lane 0, second half on lane 1, reshaped (num, 2)
  for (int32_t i = 0; i < num; ++i) {
    uint32_t x0 = (uint32_t)i, x1 = (uint32_t)(i + num);
    bsde::threefry2x32(key0, key1, x0, x1);
    h_out[i] = x0;
    h_out[num + i] = x1;
  }
  return 0;
}

extern "C" int bsde_brownian_tree_threefry(uint32_t key0, uint32_t key1, int64_t n, int32_t rep, int32_t depth, float t0, float t1,
                                           float t, int64_t first, int64_t count, float* d_out, void* stream) {
  BSDE_CHECK_ARG(n >= 0 && rep >= 0 && 2 * (int64_t)rep <= n, "need 0 <= 2*rep <= n (n=%lld, rep=%d)", (long long)n, rep);
  BSDE_CHECK_ARG(depth >= 0 && depth <= BSDE_TF_MAX_DEPTH, "depth must be in [0, %d], got %d", BSDE_TF_MAX_DEPTH, depth);
  BSDE_CHECK_ARG(t1 > t0, "need t1 > t0");
  BSDE_CHECK_ARG(n < (1LL << 32), "n must be < 2^32");
  BSDE_CHECK_ARG(first >= 0 && count >= 0 && first + count <= n, "slice [first, first+count) must lie inside [0, n)");
  if (count == 0) return 0;
  BSDE_CHECK_ARG(d_out != nullptr, "NULL tensor pointer");
  bsde::TfPath path;
  const uint32_t root[2] = {key0, key1};
  uint32_t cur[2], l[2], r[2];
  bsde::tf_split(root, path.end_key, cur);                     // end_rng, path_rng = random.split(rng), brownian.py:81
  const float s = (t - t0) / (t1 - t0);
  float tl = 0.f, tr = 1.f;
  for (int lvl = 0; lvl < depth; ++lvl) {
    path.lvl_key[lvl][0] = cur[0]; path.lvl_key[lvl][1] = cur[1];
    const float tm = (tl + tr) / 2.0f;
    bsde::tf_split(cur, l, r);
    if (s < tm) { tr = tm; cur[0] = l[0]; cur[1] = l[1]; } else { tl = tm; cur[0] = r[0]; cur[1] = r[1]; }
  }
  path.lvl_key[depth][0] = cur[0]; path.lvl_key[depth][1] = cur[1];
  long long blocks = (count + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bsde::brownian_tree_threefry_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(path, n, rep, depth, t0, t1, t, first, count,
                                                                                     d_out);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int bsde_step_update_prod(int64_t n, float dt, const float* d_y, const float* d_f, const float* d_gprod,
                                     const float* d_gdgprod, float* d_y1, void* stream) {
  BSDE_CHECK_ARG(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  BSDE_CHECK_ARG(d_y && d_f && d_gprod && d_y1, "NULL tensor pointer");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  step_update_prod_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(n, dt, d_y, d_f, d_gprod, d_gdgprod, d_y1);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int bsde_step_update(int32_t method, int64_t n, float dt, const float* d_y, const float* d_f,
                                const float* d_g, const float* d_dW, const float* d_gdg, const float* d_g2,
                                float* d_y1, void* stream) {
  BSDE_CHECK_ARG(method >= BSDE_EULER_MARUYAMA && method <= BSDE_EULER_HEUN, "Unknown method: %d", method);
  BSDE_CHECK_ARG(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  BSDE_CHECK_ARG(d_y && d_f && d_g && d_dW && d_y1, "NULL tensor pointer");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  step_update_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(method, n, dt, d_y, d_f, d_g, d_dW, d_gdg, d_g2, d_y1);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int bsde_philox_increments(uint64_t seed, uint32_t stream_id, int32_t S, int32_t nit, int64_t P,
                                      const float* d_dtk, float* d_out, void* stream) {
  BSDE_CHECK_ARG(S >= 1 && nit >= 1 && P >= 1, "S, nit, P must be positive");
  BSDE_CHECK_ARG(d_dtk && d_out, "NULL tensor pointer");
  const long long total = (long long)S * nit * P;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  philox_increments_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seed, stream_id, S, nit, P, d_dtk, d_out);
  count_launch();
  BSDE_CUDA_OK(cudaGetLastError());
  return 0;
}
