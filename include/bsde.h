/*
 * bsde.h -- C ABI of libbsde.so: the B200 (sm_100a) kernels behind bayeSDE's fixed-grid
 * SDE-BNN hot path.  Plain pointers and sizes only; no torch / C++ types cross this line.
 *
 * Conventions
 *   - every pointer named d_* / documented "device" is a CUDA device pointer owned by the
 *     CALLER (PyTorch).  The library never allocates or frees tensor memory; scratch space
 *     is passed in as `workspace` and sized with the *_workspace_bytes() queries.
 *   - every entry point is stream-ordered on `stream` (a cudaStream_t passed as void*),
 *     never synchronises the device, and returns 0 on success or a negative bsde_status.
 *     bsde_last_error() returns a thread-local message for the last failure.
 *   - all arithmetic is fp32 (the reference examples run in fp32).
 *
 * Reference interfaces replaced (paths relative to the xwinxu/bayeSDE tree):
 *   bsde_wsde_fwd/bwd        f_aug's dw/dkl rows + g_aug + the Euler/Milstein update of the
 *                            [w, kl] rows for all scan iterations:
 *                            brax/_impl/brax.py:49-78, brax/_impl/arch.py:34-74,
 *                            brax/_impl/diffeq_layers.py:73-96, jaxsde/jaxsde/sdeint.py:36-50,103-110
 *   bsde_tconv_fwd/wgrad     one ConcatConv2D (time channel appended, 3x3 / 1x1, SAME, stride 2,
 *                            transposed) with weights taken from the state w(t):
 *                            brax/_impl/diffeq_layers.py:124-150, torchbnn/_impl/diffeq_layers.py:75-166
 *   bsde_xpath_fwd/bwd       the x rows of the same scan (f_x conv stack + Euler update) for all
 *                            iterations, and its hand-written reverse pass with recomputation:
 *                            brax/_impl/arch.py:171-210, brax/_impl/brax.py:52,113,131-132,
 *                            jaxsde/jaxsde/sdeint.py:103-110,135
 *   bsde_step_update         ito_euler_step / ito_milstein_step on a flat state for arbitrary
 *                            user f, g: jaxsde/jaxsde/sdeint.py:36-50 (+ euler_heun, :73-74 stub)
 *   bsde_toy_em_fwd/bwd      em_solver + VariationalSDE drift/prior of the torch toy (config 2):
 *                            examples/torch/sdebnn_toy1d.py:19-60, torchbnn/_impl/diffeq_layers.py:178-188,
 *                            torchbnn/_impl/basic.py:298-312
 *   bsde_chain_fwd/bwd       one evaluation of SDENet's y_net drift (all groups + ConvDownsample) and its reverse:
 *                            torchbnn/_impl/models.py:25-54,159-169, torchbnn/_impl/diffeq_layers.py:40-62,75-166,230-311
 *   bsde_nchw_reinterp_axpy  the flat-state update y + a * fy.reshape(-1) of the torch front end: models.py:163,168
 *   bsde_brownian_tree       a queryable Brownian path for sdeint_ito (stochastic adjoint): jaxsde/jaxsde/brownian.py:19-95
 *   bsde_brownian_tree_threefry / bsde_threefry_split   the same tree on jax.random's threefry keys (replay of a JAX run's path):
 *                            jaxsde/jaxsde/brownian.py:69-95 + jax.random.{split,normal} (jax==0.1.77, requirements.txt:1)
 *   bsde_step_update_prod    the product-form step of the adjoint system: jaxsde/jaxsde/sde_vjp.py:13-47, sdeint.py:36-50
 *   bsde_philox_increments   replaces make_brownian_motion / virtual_brownian_tree queries
 *                            bm(next_t)-bm(curr_t): jaxsde/jaxsde/brownian.py:26-95, sdeint.py:108
 */
#ifndef BSDE_H_
#define BSDE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSDE_VERSION 200

typedef enum {
  BSDE_OK = 0,
  BSDE_ERR_INVALID_ARG = -1,
  BSDE_ERR_UNSUPPORTED = -2,
  BSDE_ERR_CUDA = -3,
  BSDE_ERR_WORKSPACE = -4
} bsde_status;

/* activations: brax/_impl/arch.py:31-32 ACTFNS (swish carries trainable state: unsupported here) */
/* BSDE_ACT_RBF = exp(-x^2) (brax/_impl/arch.py:31).  Its derivative -2 x a needs the sign of the pre-activation, which the output
 * a = exp(-x^2) does not determine: the conv epilogues store a with the sign of x in the LAST MANTISSA BIT (a perturbation of at most
 * one ulp of the stored activation) and the data-gradient epilogues read it back: x = (bit ? -1 : 1) * sqrt(-ln a). */
/* BSDE_ACT_SWISH = x * sigmoid(x * softplus(beta)) with a TRAINABLE beta per channel / hidden unit (brax/_impl/arch.py:14-29): for
 * f_x beta lives in the flat weight vector w(t) after the conv it follows (bsde_xpath_desc.beta_off), for f_w it is a parameter
 * vector of the hidden layer (bsde_fw_params.beta1 / betam). */
typedef enum { BSDE_ACT_SOFTPLUS = 0, BSDE_ACT_TANH = 1, BSDE_ACT_ELU = 2, BSDE_ACT_NONE = 3, BSDE_ACT_RBF = 4, BSDE_ACT_SWISH = 5 } bsde_act;

typedef enum { BSDE_EULER_MARUYAMA = 0, BSDE_MILSTEIN = 1, BSDE_EULER_HEUN = 2 } bsde_method;

/* conv arithmetic mode: exact fp32 FMA on CUDA cores, or tcgen05 tensor-core tiles
 * (kind::tf32 operands, fp32 accumulation in TMEM) */
typedef enum { BSDE_MATH_FP32 = 0, BSDE_MATH_TC = 1 } bsde_math;

int bsde_version(void);
const char* bsde_last_error(void);
/* number of kernels this library has launched in this process (bench bookkeeping) */
uint64_t bsde_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K1: weight-SDE path  (one SDEBNN block, S Monte-Carlo samples, all scan iterations)
 * ---------------------------------------------------------------------------------------- */

#define BSDE_FW_MAX_MID 4   /* ConcatSquashLinear layers strictly between the first and last  */
#define BSDE_FW_MAX_EDGE 4  /* max width of the first hidden layer d1 and the last hidden dL   */
#define BSDE_FW_MAX_HID 256 /* max width of a middle hidden layer                               */

/* f_w = MLP(fw_dims, xt=True): P -> d[0] -> ... -> d[n-1] -> P, every layer a
 * ConcatSquashLinear (W,b,w_t,w_tb,b_t).  The same struct shape carries the gradients. */
typedef struct {
  int32_t P;       /* w_dim */
  int32_t nhid;    /* len(fw_dims) >= 1 */
  int32_t dims[BSDE_FW_MAX_MID + 2]; /* fw_dims */
  int32_t act;     /* bsde_act between layers (fw_actfn) */
  /* first layer  P -> d1 :  W1 [P][d1] row-major, then b,w_t,w_tb,b_t [d1] each */
  float* W1; float* b1; float* wt1; float* wtb1; float* bt1;
  /* middle layers i=0..nhid-2 : d[i] -> d[i+1] :  W [d[i]][d[i+1]], b,w_t,w_tb,b_t [d[i+1]] */
  float* Wm[BSDE_FW_MAX_MID]; float* bm[BSDE_FW_MAX_MID]; float* wtm[BSDE_FW_MAX_MID];
  float* wtbm[BSDE_FW_MAX_MID]; float* btm[BSDE_FW_MAX_MID];
  /* last layer  dL -> P :  W4 [dL][P] row-major, b,w_t,w_tb,b_t [P] each */
  float* W4; float* b4; float* wt4; float* wtb4; float* bt4;
  /* act == BSDE_ACT_SWISH: trainable beta of the activation after the first hidden layer ([dims[0]]) and after middle layer i
   * ([dims[i+1]]); NULL otherwise */
  float* beta1; float* betam[BSDE_FW_MAX_MID];
  int32_t plain;   /* 1: every layer is a plain Linear (torchbnn `Linear`, time ignored: torchbnn/_impl/diffeq_layers.py:169-175,
                      make_w_net models.py:57-80): gate = 1, shift = 0; the w_t / w_tb / b_t pointers are not touched */
} bsde_fw_params;

typedef struct {
  int32_t S;          /* Monte-Carlo samples integrated together */
  int32_t nit;        /* scan iterations = len(ts)-1 */
  int32_t stl;        /* 0: g_kl = 0; 1: kl += u(w_stale, t_k) * dW_k  (brax.py:64-78, quirk Q2) */
  int32_t w0_per_sample; /* 0: w0 is [P] shared by all samples; 1: w0 is [S][P] */
  float sigma;        /* diff_coef */
  float kl_weight;    /* integrand is kl_weight * u^2 : 1 for brax (brax.py:61) */
  float ou_coef;      /* drift = f_w(w,t) + ou_coef*w : 0 for brax as shipped (quirk Q3), -1 for
                         torchbnn (models.py:165); u = (f_w + (ou_coef+1) w)/sigma (prior drift -w) */
  uint64_t seed;      /* Philox key when d_dW == NULL */
  uint32_t stream_id; /* Philox sub-stream (block index in the network) */
  uint32_t sample_offset; /* global index of local sample 0 in the Philox counter: ranks that shard the MC samples
                             (examples/jax/sdebnn_classification.py:218 vmaps over per-sample rngs) draw the same paths as one rank
                             integrating all of them */
  /* Optional stage tables, device arrays [nit] (all NULL = the Euler scan above).  They express fixed-step schemes whose
   * stages are Euler-like updates that start from the previous slot, e.g. torchsde's midpoint (call site
   * torchbnn/_impl/models.py:184-197;  y' = y + dt/2 f(t,y) + g I/2,  y1 = y + dt f(t + dt/2, y') + g I):
   *     w_{j+1} = w_{j - back_j} + dt_j f_w(w_j, t_j) + sigma * noise_j ,   kl += dtkl_j * kl_weight * u(w_j)^2
   *     noise_j = d_dW[s][j][p]   or   nscale_j * N(0,1)[Philox counter (p, nidx_j, s, stream_id)]
   * back_j in {0, 1, 2}, back_0 = 0 and no two consecutive non-zeros; back_j = 2 starts the update from the average
   * (w_{j-1} + w_j)/2 -- the corrector of torchsde's trapezoidal `heun` (y' = y + h f(t,y) + g I, y1 = y + h/2 (f(t,y) + f(t+h,y')) + g I):
   * dt = (h, h/2), dtkl = (h/2, h/2), nscale = (sqrt(h), sqrt(h)/2), nidx = (k, k), back = (0, 2).  midpoint: dt = (h/2, h), dtkl = (0, h), nscale = (sqrt(h)/2, sqrt(h)),
   * nidx = (k, k), back = (0, 1) for step k; the trajectory then holds w_k in slot 2k and the midpoint state in slot 2k+1. */
  const float* d_dtkl; const float* d_nscale; const int32_t* d_nidx; const int32_t* d_back;
  /* Optional output of bsde_wsde_fwd (NULL = not written): the kl ROWS of the state at every scan time, [S][nit+1][P]
   * (slot 0 = 0) -- what `ys[:, x_dim+P:]` of sdeint_ito_fixed_grid holds (jaxsde/jaxsde/sdeint.py:138; brax.py:110).  The
   * layer itself only consumes their sum at the final time (brax.py:119), which is d_kl. */
  float* d_kltraj;
} bsde_wsde_desc;

size_t bsde_wsde_workspace_bytes(const bsde_wsde_desc* d, const bsde_fw_params* fw);

/* Forward.  d_tk,d_dtk: [nit] device floats (t_k, dt_k of each scan iteration).
 * d_dW: [S][nit][P] supplied increments or NULL for in-kernel Philox.
 * d_w_stale: [P] (only read when stl=1).
 * Outputs: d_wtraj [S][nit+1][P] (w_0..w_nit), d_z1 nit*S*d1 floats (first-layer dots, opaque, kept for
 * the reverse pass), d_kl [S] (sum over P of the kl rows at the final time, brax.py:119). */
int bsde_wsde_fwd(const bsde_wsde_desc* d, const bsde_fw_params* fw, const float* d_w0,
                  const float* d_tk, const float* d_dtk, const float* d_dW, const float* d_w_stale,
                  float* d_wtraj, float* d_z1, float* d_kl, void* workspace, size_t workspace_bytes,
                  void* stream);

/* Reverse pass.  d_gw_traj: [S][nit][P] dL/dw_k contributions from the x rows (conv wgrads,
 * already multiplied by dt_k) or NULL; d_gkl [S] = dL/dkl_s; d_gwT [S][P] = dL/dw_nit or NULL.
 * Outputs: d_gw0 ([P] summed over samples, or [S][P] when w0_per_sample), `gfw` = gradient
 * buffers with the layout of `fw` (overwritten, summed over samples). */
int bsde_wsde_bwd(const bsde_wsde_desc* d, const bsde_fw_params* fw, const float* d_tk,
                  const float* d_dtk, const float* d_wtraj, const float* d_z1,
                  const float* d_gw_traj, const float* d_gkl, const float* d_gwT, float* d_gw0,
                  const bsde_fw_params* gfw, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2: time-dependent convolution with weights taken from a flat weight vector
 * ---------------------------------------------------------------------------------------- */

/* One ConcatConv2D-style layer.  Activations are always NHWC in device memory (the torch front
 * end keeps its NCHW tensors in channels-last storage).  The kernel layout is described by
 * strides: HWIO with the time channel LAST for the JAX front end, OIHW / IOHW with the time
 * channel FIRST for the torch front end.  Spatial map (per axis):
 *     num = o*sn + k*kn + off ;  valid iff num >= 0, num % sd == 0, num/sd < in_size ; i = num/sd
 * which covers SAME / explicit padding, stride 2, lhs-dilated (transposed) convs and all their
 * data gradients. */
typedef struct {
  int32_t cin, cout;         /* real channels (time channel excluded) */
  int32_t ksize;             /* 1 or 3 */
  int32_t hin, win, hout, wout;
  int32_t sn, kn, off, sd;   /* spatial map */
  int32_t has_time;          /* 1: a constant channel t takes part in the reduction */
  int32_t time_first;        /* weight row of the time channel: 0 -> index cin, 1 -> index 0 */
  int64_t w_off, b_off;      /* offsets into the flat weight vector; b_off < 0: no bias */
  int64_t w_tap_stride, w_k_stride, w_n_stride; /* weight index = w_off + tap*ts + kc*ks + n*ns */
} bsde_conv_layer;

typedef enum {
  BSDE_EPI_BIAS = 0,      /* y = acc + b                                  */
  BSDE_EPI_BIAS_ACT = 1,  /* y = act(acc + b)                             */
  BSDE_EPI_EULER = 2,     /* y = aux + scale*(acc + b)      (aux = x_k)   */
  BSDE_EPI_DACT = 3,      /* y = scale*acc*act'(aux)        (aux = act output at y's position) */
  BSDE_EPI_ACCUM = 4      /* y = aux + scale*acc            (aux may alias y) */
} bsde_epilogue;

/* y[n,...] = epilogue( sum_{tap,c} in(n, map(o,tap), c) * w[sample(n)][tap,c,:] ).
 * images n in [0, S*B); image n uses the weight vector d_w + (n / B) * w_sample_stride. */
int bsde_tconv_fwd(const bsde_conv_layer* L, int32_t S, int32_t B, const float* d_in,
                   const float* d_w, int64_t w_sample_stride, float t, int32_t act,
                   int32_t epilogue, float scale, const float* d_aux, float* d_out, int32_t math,
                   void* workspace, size_t workspace_bytes, void* stream);

/* scratch for the repacked weight tiles of the tensor-core path (0 for BSDE_MATH_FP32) */
size_t bsde_tconv_fwd_workspace_bytes(const bsde_conv_layer* L, int32_t S, int32_t math);

size_t bsde_tconv_wgrad_workspace_bytes(const bsde_conv_layer* L, int32_t S, int32_t B);

/* d_gw[s][w_off + ...] = scale * sum_{n in sample s, o} in(n, map(o,tap), c) * dy[n,o,:]  (and the
 * bias / time-channel rows), written with the layer's weight strides. */
int bsde_tconv_wgrad(const bsde_conv_layer* L, int32_t S, int32_t B, const float* d_in,
                     const float* d_dy, float t, float scale, float* d_gw, int64_t gw_sample_stride,
                     int32_t math, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2+K3: the x rows of one SDEBNN block over all scan iterations, and the reverse pass
 * ---------------------------------------------------------------------------------------- */

#define BSDE_MAX_LAYERS 4

typedef struct {
  int32_t S, B;            /* samples, images per sample */
  int32_t nit;             /* scan iterations */
  int32_t nlayers;         /* convs in f_x (arch.py:176-205: 4, 3 or 2) */
  int32_t act;             /* fx_actfn */
  int32_t math;            /* bsde_math */
  int64_t P;               /* w_dim: stride between w_k and w_{k+1} in d_wtraj */
  int64_t x_numel;         /* S*B*H*W*C */
  bsde_conv_layer layers[BSDE_MAX_LAYERS];
  int64_t beta_off[BSDE_MAX_LAYERS]; /* act == BSDE_ACT_SWISH: offset inside w of the beta vector (layers[l].cout entries) of the
                                        activation that follows conv l (arch.py:172-174; l < nlayers - 1); ignored otherwise */
} bsde_xpath_desc;

/* backward: 0 forward scan; 1 reverse pass that recomputes the activations (d_acts == NULL); 2 reverse pass with stored activations
 * (d_acts != NULL): when the deferred mode applies (tensor-core route, nit >= 2, JAX weight layout) it also holds dZ and gx of
 * every iteration and a [k][s][P] gradient buffer, so that the weight gradients of 74 / S iterations run as one launch per layer. */
size_t bsde_xpath_workspace_bytes(const bsde_xpath_desc* d, int32_t backward);

/* bytes of the optional per-iteration activation store (all nit iterations) */
size_t bsde_xpath_acts_bytes(const bsde_xpath_desc* d);

/* h_tk,h_dtk: HOST arrays [nit].  d_wtraj: [S][nit+1][P].  d_xs: [nit+1][x_numel]; slot 0 holds
 * x_0 on entry, slot k+1 is written with x_{k+1} = x_k + dt_k f_x(x_k, t_k; w_k).
 * d_acts: NULL (the reverse pass recomputes the conv intermediates from x_k: the jax.checkpoint /
 * --remat analogue, brax.py:131-132) or bsde_xpath_acts_bytes() bytes that receive them. */
int bsde_xpath_fwd(const bsde_xpath_desc* d, const float* h_tk, const float* h_dtk,
                   const float* d_wtraj, float* d_xs, float* d_acts, void* workspace,
                   size_t workspace_bytes, void* stream);

/* d_gx: [x_numel] holds dL/dx_nit on entry and dL/dx_0 on return.  d_gw_traj: [S][nit][P] receives
 * dL/dw_k through the convolutions (every element of every slot is written). */
int bsde_xpath_bwd(const bsde_xpath_desc* d, const float* h_tk, const float* h_dtk,
                   const float* d_wtraj, const float* d_xs, const float* d_acts, float* d_gx,
                   float* d_gw_traj, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * torch front end: the feed-forward drift net on the hidden state (make_y_net, torchbnn/_impl/models.py:25-54,
 * blocks of torchbnn/_impl/diffeq_layers.py:280-311 + ConvDownsample :230-239) -- one evaluation and its reverse
 * ---------------------------------------------------------------------------------------- */

#define BSDE_CHAIN_MAX_LAYERS 40

typedef struct {
  int32_t S, B;            /* weight samples, images per sample */
  int32_t nlayers;
  int32_t act;             /* bsde_act applied after the layers flagged in act_after */
  int32_t math;            /* bsde_math */
  int32_t act_after[BSDE_CHAIN_MAX_LAYERS];      /* 1: DiffEqWrapper(activation) follows conv l; must be 0 for the last */
  bsde_conv_layer layers[BSDE_CHAIN_MAX_LAYERS]; /* layer l's output shape is layer l+1's input shape */
} bsde_chain_desc;

size_t bsde_chain_workspace_bytes(const bsde_chain_desc* d, int32_t backward);
/* floats of the buffer that receives every layer's output (256-byte aligned slots), and the offset (floats) of the
 * last layer's output f inside it */
size_t bsde_chain_acts_floats(const bsde_chain_desc* d);
size_t bsde_chain_out_offset(const bsde_chain_desc* d);

/* f = y_net(t, x; w) (DiffEqSequential.forward with explicit params, diffeq_layers.py:49-56).  d_x: [S*B] NHWC images;
 * image n uses the flat weight vector d_w + (n / B) * w_sample_stride.  d_acts receives all layer outputs. */
int bsde_chain_fwd(const bsde_chain_desc* d, float t, const float* d_w, int64_t w_sample_stride, const float* d_x,
                   float* d_acts, void* workspace, size_t workspace_bytes, void* stream);

/* Reverse of one evaluation.  d_gf = dL/df (shape of the last layer's output).  d_gx receives dL/dx (added to its
 * content when accumulate_gx != 0); d_gw + s * gw_sample_stride receives dL/dw of sample s (every element that belongs
 * to a layer of the chain is written). */
int bsde_chain_bwd(const bsde_chain_desc* d, float t, const float* d_w, int64_t w_sample_stride, const float* d_x,
                   const float* d_acts, const float* d_gf, float* d_gx, int32_t accumulate_gx, float* d_gw,
                   int64_t gw_sample_stride, void* workspace, size_t workspace_bytes, void* stream);

/* SDENet.f adds the drift to the state in flat NCHW order (models.py:163,168 `fy.reshape(-1)`): with NHWC storage
 *     out[n][h][w][c] = (d_base ? d_base[n][h][w][c] : 0) + a * src[n][h2][w2][c2],   c*H*W + h*W + w == c2*H2*W2 + h2*W2 + w2
 * out/base: [nimg][H][W][C], src: [nimg][H2][W2][C2].  Swapping the shapes gives the transpose (reverse pass). */
int bsde_nchw_reinterp_axpy(int64_t nimg, int32_t C, int32_t H, int32_t W, int32_t C2, int32_t H2, int32_t W2, float a,
                            const float* d_base, const float* d_src, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * generic flat-state step and Brownian increments
 * ---------------------------------------------------------------------------------------- */

/* y1 = y + dt*f + g*dW  [+ 0.5*gdg*(dW^2 - dt)  (milstein, d_gdg = J_g (.) g, may be NULL)]
 *                       [heun: y + dt*f + 0.5*(g + g2)*dW, d_g2 = g(y + g dW, t+dt)]          */
int bsde_step_update(int32_t method, int64_t n, float dt, const float* d_y, const float* d_f,
                     const float* d_g, const float* d_dW, const float* d_gdg, const float* d_g2,
                     float* d_y1, void* stream);

/* out[s][k][p] = sqrt(max(dt_k,0)) * N(0,1) from Philox4x32-10 keyed (seed) with counter
 * (p, k, s, stream_id): the stream bsde_wsde_fwd consumes when d_dW == NULL. */
int bsde_philox_increments(uint64_t seed, uint32_t stream_id, int32_t S, int32_t nit, int64_t P,
                           const float* d_dtk, float* d_out, void* stream);

/* Virtual Brownian tree (jaxsde/jaxsde/brownian.py:19-95, make_brownian_motion :92-95): W(t) on [t0, t1] for n state entries,
 * O(1) memory, queryable at any t in any order.  Standard motion B on [0, 1] with B(1) ~ N(0, 1); `depth` bisection levels,
 * every midpoint a Brownian-bridge draw (:19-23); a final bridge point at s = (t - t0)/(t1 - t0) inside the leaf;
 * W(t) = sqrt(t1 - t0) * B(s).  The normals come from Philox4x32-10 keyed (seed) with counter (element, heap index of the tree
 * node, ., stream_id) in place of threefry key splitting (NOT bit-compatible with a JAX run); `rep` ties the last rep entries
 * to the rep entries before them (sample_with_rep, :69-74). */
int bsde_brownian_tree(uint64_t seed, uint32_t stream_id, int64_t n, int32_t rep, int32_t depth, float t0, float t1, float t,
                       float* d_out, void* stream);

/* The same tree driven by JAX's threefry PRNG, so that the Brownian path of a JAX run of the reference can be
 * queried here: `make_brownian_motion(t0, zeros(n), t1, rng, depth, rep)` (jaxsde/jaxsde/brownian.py:77-95) with rng = (key0, key1),
 * every draw `jax.random.normal(key, (n,))` in float32 and every descent `jax.random.split(key)` (jax==0.1.77: Threefry-2x32-20,
 * uniform from the top 23 bits on [nextafter(-1,0), 1), sqrt(2)*erf_inv with XLA's float32 polynomial).  The integer stream is
 * exact; the float transform follows the same operation order (log1p / sqrt to within 1-2 ulp of XLA's).  Writes the entries
 * [first, first+count) of the n-vector W(t) to d_out[0..count). */
int bsde_brownian_tree_threefry(uint32_t key0, uint32_t key1, int64_t n, int32_t rep, int32_t depth, float t0, float t1, float t,
                                int64_t first, int64_t count, float* d_out, void* stream);

/* jax.random.split(key, num) on the host: h_out receives num keys as [num][2] words (threefry over iota(2*num)). */
int bsde_threefry_split(uint32_t key0, uint32_t key1, int32_t num, uint32_t* h_out);

/* y1 = y + dt*f + gprod [+ 0.5*gdgprod]: ito_euler_step / ito_milstein_step (jaxsde/jaxsde/sdeint.py:36-50) when the
 * diffusion enters as the products g_prod(y,t,args,dW) and gdg_prod(y,t,args,dW^2 - dt) -- the form the augmented adjoint
 * system of sdeint_ito's reverse pass takes (jaxsde/jaxsde/sde_vjp.py:13-47,137-199). */
int bsde_step_update_prod(int64_t n, float dt, const float* d_y, const float* d_f, const float* d_gprod,
                          const float* d_gdgprod, float* d_y1, void* stream);

/* ------------------------------------------------------------------------------------------
 * config 2 (torch toy): em_solver + VariationalSDE of examples/torch/sdebnn_toy1d.py:19-60
 * ---------------------------------------------------------------------------------------- */

/* f = construct_odenet([1, h1, h2, 1]) (sdebnn_toy1d.py:35-41): ConcatLinear / TimeDependentSwish,
 * g = sigma, prior drift -theta * x.  torch nn.Linear layouts: W1 [h1][2], W2 [h2][h1+1], W3 [1][h2+1],
 * column 0 = time channel (torchbnn/_impl/utils.py:201-203).  beta1 [nt-1][h1], beta2 [nt-1][h2] are the
 * TimeDependentSwish gates beta(t_k) (torchbnn/_impl/basic.py:298-312), evaluated by the caller. */
typedef struct {
  int32_t n;      /* samples (independent scalar paths) */
  int32_t nt;     /* time points; nt-1 Euler-Maruyama steps */
  int32_t h1, h2; /* hidden widths, <= 128 */
  float sigma;    /* diffusion */
  float theta;    /* OU prior coefficient */
} bsde_toy_desc;

/* d_ts [nt]; d_z0 [n]; d_dW [nt-1][n] increments (already N(0,dt)) or NULL for Philox(seed).
 * Outputs: d_zt [nt][n] (z_0 .. z_T), d_kl [n] = sum_k dt_k u_k^2 / 2 with u evaluated at the NEW state. */
int bsde_toy_em_fwd(const bsde_toy_desc* d, const float* d_ts, const float* d_z0, const float* d_dW, uint64_t seed,
                    const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                    const float* b3, const float* beta1, const float* beta2, float* d_zt, float* d_kl,
                    void* stream);

/* Reverse pass.  d_gzt [nt][n] = dL/dz_t (or NULL), d_gkl [n] = dL/dkl (or NULL).  Outputs: d_gz0 [n] and the
 * per-(step, sample) layer gradients whose sums over (step, sample) are the parameter gradients:
 * DA1 [nt-1][n][h1], DA2 [nt-1][n][h2] (pre-activation gradients), DF [nt-1][n] (output gradient),
 * H1, H2 (hidden activations, same shapes), GB1, GB2 (dL/dbeta per step, sample, unit). */
int bsde_toy_em_bwd(const bsde_toy_desc* d, const float* d_ts, const float* d_zt, const float* d_gzt,
                    const float* d_gkl, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* W3, const float* b3, const float* beta1, const float* beta2, float* d_gz0,
                    float* d_DA1, float* d_DA2, float* d_DF, float* d_H1, float* d_H2, float* d_GB1,
                    float* d_GB2, void* stream);

/* ------------------------------------------------------------------------------------------
 * config 1 (JAX toy): SDELayer -- augmented_post_drift / augmented_prior_drift / augmented_diffusion integrated by _sdeint
 * brax/_impl/layers.py:54-157,164-248, brax/utils/sdeint.py:8-60, examples/jax/sdebnn_toy1d.py:32-95
 * ---------------------------------------------------------------------------------------- */

/* State row of one MC sample: [w (P), x (B*D), kl (1)] (ravel order of layers.py:116-118,159-162).  driftx = CSL(H) act CSL(H) act
 * CSL(D) on the rows of x with ITS WEIGHTS TAKEN FROM w (ravel_pytree order: per layer W [din][dout], b, w_t, w_tb, b_t, then the
 * (beta,) leaf of a trainable swish); driftw = P -> H -> H -> P ConcatSquashLineartx stack (bsde_fw_params with nhid = 2,
 * dims = {H, H}; W1 [P][H], Wm[0] [H][H], W4 [H][P]; beta1 / betam[0] for swish), optionally added to the prior drift
 * (`diff_drift`, layers.py:272-292); prior drift -w (`prior_ou`) or 0; diffusion sigma on the w rows (`diffw`), 0 elsewhere.
 * Per step (sdeint.py:13-22): eps ~ N(0, I_P);  u = (dw - prior) / sigma;
 *   w += h dw + sqrt(h) sigma eps;   x += h driftx(w, t, x);   kl += h (sum u^2 / 2 + sum u_sg eps)
 * where u_sg is u evaluated with stop_gradient(driftw parameters) when `stop_grad` (layers.py:81-112; a value-neutral change that
 * only redirects gradients).  `prior_solve` = 1 integrates augmented_prior_drift instead (dw = prior drift, u = dw / sigma). */
typedef struct {
  int32_t S;             /* MC samples (the reference vmaps apply_fn over rngs, sdebnn_toy1d.py:127): one CTA each */
  int32_t B, D, H;       /* rows of x, width of x (x_dim + aug_dim <= 16), hidden width of both nets (<= 128) */
  int32_t nt;            /* time points, nt - 1 steps */
  int32_t act;           /* bsde_act of both nets */
  int32_t driftw, diffw, prior_ou, diff_drift, stop_grad;   /* `setting` of sdebnn_toy1d.py:186-198 */
  int32_t prior_solve;
  int32_t w0_per_sample; /* d_w0 is [S][P] instead of [P] */
  float sigma;           /* diff_const */
  int64_t P;             /* w_dim */
  uint64_t seed;         /* in-kernel Philox stream (when d_eps is NULL): counter (p, step, sample / 4, stream_id) */
  uint32_t stream_id;
  uint32_t sample_offset;
} bsde_sdelayer_desc;

/* d_ts [nt]; d_w0 [P] or [S][P]; d_x0 [B*D] (the same inputs for every sample); d_eps [S][nt-1][P] unit normals of the w rows
 * (exact-comparison mode) or NULL; d_traj [S][nt][P + B*D + 1] receives every state, row 0 = the initial state
 * (`sdeint` returns rows 1.., `_sdeint` the last row). */
int bsde_sdelayer_fwd(const bsde_sdelayer_desc* d, const bsde_fw_params* fw, const float* d_ts, const float* d_w0,
                      const float* d_x0, const float* d_eps, float* d_traj, void* stream);

/* Reverse pass.  Cotangents: d_gtraj [S][nt][P + B*D + 1] (every state) or NULL, d_gT [S][P + B*D + 1] (final state only) or
 * NULL.  d_g0 [S][P + B*D + 1] receives dL/d(initial state) per sample.  gfw: accumulators of the driftw parameter gradients
 * with a LEADING SAMPLE DIMENSION ([S][P][H], [S][H], ...; zero them before the call; the caller sums over samples). */
int bsde_sdelayer_bwd(const bsde_sdelayer_desc* d, const bsde_fw_params* fw, const float* d_ts, const float* d_eps,
                      const float* d_traj, const float* d_gtraj, const float* d_gT, float* d_g0, const bsde_fw_params* gfw,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BSDE_H_ */
