/* vtb.h — C ABI of libvtb_b200.so: the B200-native (sm_100a) execution layer behind
 * vision_toolbox's ConvNormAct / Darknet / VoVNet hot path.
 *
 * The reference (gau-nernst/vision-toolbox) is pure Python: its "operator interface" for this path is
 * the torch.nn module stack built in vision_toolbox/components.py:13-46 (ConvNormAct = Conv2d(bias=False)
 * -> BatchNorm2d -> ReLU(inplace)), vision_toolbox/backbones/darknet.py:20-55 (residual / CSP blocks) and
 * vision_toolbox/backbones/vovnet.py:20-63 (eSE / OSA blocks).  Each entry point below replaces one torch
 * library call made from those lines; the comment on each function names the call site it stands in for.
 *
 * Conventions
 *  - plain C: pointers, ints, sizes.  No torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *  - the CALLER owns all device memory; the library never allocates, frees or synchronises.
 *  - activations are NHWC bf16 "views": base pointer + pixel pitch `ld` (elements) so a tensor may be a
 *    channel slice of a wider concat buffer (ld >= C, ld % 8 == 0, base 16-byte aligned).
 *  - parameters / statistics / parameter gradients are fp32 in the reference's own layouts (OIHW, [C]).
 *  - every function returns 0 on success, a negative VTB_E* code otherwise; vtb_last_error() gives text.
 *  - all functions are asynchronous on `stream` and re-entrant per stream, with ONE restriction: vtb_bn_bwd_fused without
 *    SyncBN peers is an ordinary (not cooperative) launch whose <= num_SMs blocks meet at a hand-rolled grid barrier.
 *    Everything else this library launches finishes without waiting for it, so within one model's streams the blocks
 *    always become co-resident; two such kernels in flight on two user streams of one device, or foreign work that holds
 *    SMs indefinitely, can starve the barrier.  VTB_BWD_COOP=1 selects the cooperative launch (co-residency guaranteed by
 *    the driver) for such setups; the SyncBN variant always uses it.
 *  - SyncBN assumes the same per-rank batch size on every rank (count * world); BatchNorm2d(momentum=None) is rejected.
 */
#ifndef VTB_H_
#define VTB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTB_OK 0
#define VTB_EINVAL (-1)   /* unsupported / inconsistent arguments */
#define VTB_ECUDA (-2)    /* CUDA runtime / driver error          */
#define VTB_ENODEV (-3)   /* no sm_100 device / driver too old    */

/* Geometry of one square convolution (nn.Conv2d(cin, cout, k, stride, padding=pad, bias=False),
 * components.py:26-35).  `cin` is the channel count of the NHWC operand (the 3-channel image is stored
 * padded to 16 channels; everything else is already a multiple of 16). */
typedef struct VtbConv {
  int n, h, w;  /* input batch / height / width */
  int cin, cout;
  int k, stride, pad;
} VtbConv;

const char* vtb_last_error(void);
int vtb_version(void);
/* number of SMs of the current device (grid sizing), <0 on error */
int vtb_num_sms(void);
/* total kernels launched by this library in this process (bench.py's gpu_launches claim) */
long long vtb_launch_count(void);

/* ---- shape / workspace queries (host only, no GPU needed) ---- */
int vtb_conv_out_hw(const VtbConv* c, int* ho, int* wo);
/* rows of the per-CTA statistics scratch written by vtb_conv_fprop (one per thread block): floats = rows * cout * 2 */
int vtb_conv_stats_rows(const VtbConv* c);
size_t vtb_conv_wgrad_workspace_bytes(const VtbConv* c);
/* The tiling the library will use for this geometry (introspection for tests / benchmarks: "did this case really run
 * multi-tile CTAs, two TMEM sets, K-split chains, split-K wgrad?").  op: 0 fprop, 1 dgrad (stride 2: the largest phase),
 * 2 wgrad.  info[8] = fprop/dgrad: {block_m, block_n, grid, tiles, max tiles per CTA, TMEM sets, ksplit, stages};
 * wgrad: {pixels per stage, columns per tile, CTAs per split, splits, pixel blocks, ksplit, stages, 0}. */
int vtb_conv_tiling_info(const VtbConv* c, int op, int* info);

/* ---- weights ----
 * Re-pack an OIHW fp32 master weight (nn.Conv2d.weight, components.py:26) into the two bf16 matrices the
 * tensor-core kernels consume: wf[cout][k*k][cin] (fprop / wgrad K-order) and wd[cin][k*k][cout] (dgrad).
 * cin_real <= c->cin is the channel count of the fp32 tensor (3 for the image stem). wd may be NULL. */
int vtb_pack_weight(const VtbConv* c, const float* w_oihw, int cin_real, void* wf, void* wd, void* stream);
/* The same re-pack for every convolution of a model in ONE launch.  The host code runs it at the start of every forward:
 * a parameter's torch version counter cannot tell whether the weights changed (fused optimizers such as
 * torch.optim.SGD(fused=True) update parameters without bumping it), so the bf16 operands are always rebuilt from the
 * fp32 masters (reads 4 B + writes 2 x 2 B per weight: ~35 us for CSPDarknet-53).
 * jobs_device: njobs VtbPackJob records in DEVICE memory, first_block = running sum of vtb_pack_job_blocks() over the
 * preceding jobs (first job: 0); total_blocks = the sum over all jobs; at most 256 jobs per launch. */
typedef struct VtbPackJob {
  const float* w;        /* OIHW fp32 master, [cout][cin_real][kk] */
  void* wf;              /* bf16 [cout][kk][cin] */
  void* wd;              /* bf16 [cin][kk][wd_ld] written at columns [wd_co_off, wd_co_off + cout), may be NULL */
  int cout, cin_real, cin, kk;
  int wd_ld, wd_co_off;  /* wd_ld >= cout: two convolutions that share their input can be packed side by side into ONE
                            dgrad operand (CSP conv1 | conv2, darknet.py:52-53); plain case: wd_ld = cout, wd_co_off = 0 */
  int wf_ld;             /* row pitch of wf in elements: kk * cin, or more (rows of the gathered-operand stem are padded
                            to a multiple of 16 columns; the pad columns are never written - keep them zero) */
  long long first_block;
  /* vtb_sgd_pack_weights only (ignored by vtb_pack_weights): gradient and momentum buffer in the master's OIHW layout */
  const float* g;
  float* m;
  float weight_decay;
} VtbPackJob;
long long vtb_pack_job_blocks(int cout, int cin, int kk);
int vtb_pack_weights(const VtbPackJob* jobs_device, int njobs, long long total_blocks, void* stream);

/* ---- optimizer step: replaces torch.optim.SGD (reference classifier.py:141-169: momentum 0.9, weight decay on conv /
 * linear weights only) ----
 * vtb_sgd_pack_weights: the SGD update of every convolution weight (g += weight_decay * w; m = momentum * m + g;
 *   w -= lr * m; zero-initialised m reproduces torch's first step) FUSED with the bf16 re-pack above: the operands of the
 *   next forward are cut from the updated master in the same pass, so no separate vtb_pack_weights launch runs.
 * vtb_sgd_step: the same update for plain tensors (BatchNorm weight / bias, the classifier head), one launch for a table
 *   of jobs; first_block = running sum of vtb_sgd_job_blocks(n) over the preceding jobs.
 * hyper_device: two floats in DEVICE memory {lr, momentum} (a learning-rate schedule updates them without re-capturing
 *   a CUDA graph). */
typedef struct VtbSgdJob {
  float* w;
  const float* g;
  float* m;
  long long n;
  float weight_decay;
  long long first_block;
} VtbSgdJob;
int vtb_sgd_pack_weights(const VtbPackJob* jobs_device, int njobs, long long total_blocks, const float* hyper_device,
                         void* stream);
long long vtb_sgd_job_blocks(long long n);
int vtb_sgd_step(const VtbSgdJob* jobs_device, int njobs, long long total_blocks, const float* hyper_device, void* stream);

/* ---- convolution: replaces aten::convolution (cuDNN) at components.py:26-35 ----
 * y[n,ho,wo,:] = conv(x)  (bf16, fp32 accumulate).
 * stats_partial != NULL: also accumulates per-channel sum / sum of squares of the bf16-rounded result
 *   (the quantities BatchNorm2d needs, components.py:36) into [vtb_conv_stats_rows][cout][2] floats
 *   (the call zeroes the buffer first).
 * scale/shift != NULL: fused eval-mode epilogue y = conv*scale[c] + shift[c], then ReLU if relu != 0,
 *   then + residual (bf16 NHWC, pitch ldr) if residual != NULL (darknet.py:28, vovnet.py:60-61). */
int vtb_conv_fprop(const VtbConv* c, const void* x, int ldx, const void* wf, void* y, int ldy, float* stats_partial,
                   const float* scale, const float* shift, int relu, const void* residual, int ldr, void* stream);

/* Training-mode fusion of the two library calls at components.py:26-36: the convolution above (with statistics)
 * PLUS vtb_bn_finalize below, executed by the last thread block of the convolution kernel to finish - BatchNorm2d's
 * statistics finalisation then costs no kernel launch.  `tickets`: >= 128 zero-initialised uint32 owned by the caller
 * (the kernel leaves them zero; one array can serve every layer launched on the same stream). */
struct VtbSyncBn;
typedef struct VtbBnTrain {
  double count;                    /* elements per channel: N*Ho*Wo (the GLOBAL count over all ranks under SyncBN) */
  const float* gamma;              /* norm.weight */
  const float* beta;               /* norm.bias */
  float eps, momentum;
  float* running_mean;             /* may be NULL (track_running_stats=False) */
  float* running_var;
  long long* num_batches_tracked;  /* may be NULL */
  float* mean;                     /* out [cout] */
  float* invstd;                   /* out [cout] */
  float* scale;                    /* out [cout]: gamma*invstd */
  float* shift;                    /* out [cout]: beta - mean*scale */
  unsigned int* tickets;
  const struct VtbSyncBn* sync;    /* NULL: single-GPU statistics; else the last block also exchanges the sums with all
                                      ranks over NVLink peer memory before finalising (see vtb_bn_sync_* below) */
  /* Two ConvNormAct units that read the SAME input (CSPDarknetStage conv1 | conv2, darknet.py:46-47,52-53) run as ONE
   * convolution over side-by-side packed weights: output channels >= split belong to the second unit and use its
   * parameter tensors below (indexed from 0).  split = 0: one unit.  mean/invstd/scale/shift stay [cout] arrays. */
  int split;
  const float* gamma2;
  const float* beta2;
  float* running_mean2;
  float* running_var2;
  long long* num_batches_tracked2;
  /* Fused normalise (optional, act_out != NULL; single unit only: split must be 0): once the statistics are final the
   * same launch also writes act_out = [relu](y*scale + shift) [+ act_residual] (NHWC bf16 views with pitches act_ld /
   * act_ldr) - components.py:36-39 (+ the residual add of darknet.py:28) without a separate vtb_bn_act launch: every
   * thread block re-reads the raw tiles it produced (L2) after its n-block's coefficients were published.  Needs all
   * thread blocks of the launch co-resident (the grid never exceeds the SM count; do not run another persistent kernel
   * that pins SMs next to it).  y is still written (BatchNorm backward reads it). */
  void* act_out;
  int act_ld;
  int act_relu;
  const void* act_residual;
  int act_ldr;
} VtbBnTrain;
int vtb_conv_fprop_bn(const VtbConv* c, const void* x, int ldx, const void* wf, void* y, int ldy, float* stats_partial,
                      const VtbBnTrain* bn, void* stream);

/* ---- convolution backward: replaces aten::convolution_backward (autograd of components.py:26-35) ----
 * dgrad: dx = conv_transpose(dy, w); accumulate != 0 adds into dx (gradient fan-in of residual / CSP /
 * OSA branches: darknet.py:28,53 ; vovnet.py:55,61) with bf16 rounding of each addend like autograd. */
int vtb_conv_dgrad(const VtbConv* c, const void* dy, int lddy, const void* wd, void* dx, int lddx, int accumulate,
                   void* stream);
/* The same dgrad for a 3x3 / stride-2 / pad-1 convolution with even H and W and few channels (cin <= 64), as ONE dense
 * GEMM over 2x2 super-pixels of dx instead of four output-parity phase launches that each stream the whole dy:
 * 16/9 of the minimal FLOPs (7 of 16 weight blocks are zero) but dy is read once - a win where the layer is HBM-bound.
 * dx must be dense (lddx == cin).  workspace: vtb_conv_dgrad_s2_workspace_bytes(c) bytes (0: geometry not supported, use
 * vtb_conv_dgrad), 128-byte aligned; it receives the merged weight matrix [4*cin][4*cout] cut from wd by this call. */
size_t vtb_conv_dgrad_s2_workspace_bytes(const VtbConv* c);
int vtb_conv_dgrad_s2(const VtbConv* c, const void* dy, int lddy, const void* wd, void* workspace, void* dx, int lddx,
                      int accumulate, void* stream);
/* dgrad that ALSO starts the BatchNorm(+ReLU) backward of the layer(s) that produced x (autograd chain
 * ConvolutionBackward0 -> ReluBackward0 -> NativeBatchNormBackward0 of components.py:26-39): when this call is the
 * LAST contribution to dx, dx is the complete gradient `g` of the producer's output.  The epilogue then reads the
 * producer's raw conv output y at the pixels it stores and reduces, per channel, sum(dz) and sum(dz * xhat) with
 * dz = g * (y*scale + shift > 0), xhat = (y - mean) * invstd; the last thread block finalises (and exchanges the sums
 * under SyncBN): dgamma, dbeta (overwritten, local sums) and coef[c][2] = (mean dz, mean dz*xhat) over `count`.
 * BatchNorm's backward for that layer is then ONE apply pass (vtb_bn_bwd_apply with this coef) - no reduction pass,
 * no grid barrier.  x may be the concatenation of two producers' outputs (CSPDarknetStage, darknet.py:53):
 * channels [0, split) belong to layer[0], [split, cin) to layer[1] (split = 0: one producer); split must be a multiple of
 * vtb_conv_dgrad_panel_w(c).  partial: vtb_conv_dgrad_stats_rows(c) * cin * 2 floats of scratch. */
typedef struct VtbBnBwdLayer {
  const void* y;          /* raw conv output of the producer, bf16 NHWC view on x's pixel lattice */
  int ldy;
  const float* scale;     /* gamma * invstd  (vtb_bn_finalize / vtb_conv_fprop_bn) */
  const float* shift;
  const float* mean;
  const float* invstd;
  int relu;               /* the producer's activation: 1 = ReLU (mask recomputed from y), 0 = none */
  float* dgamma;          /* out [c], may be NULL */
  float* dbeta;           /* out [c], may be NULL */
  float* coef;            /* out [c][2] */
} VtbBnBwdLayer;
typedef struct VtbDgradBn {
  int split;
  VtbBnBwdLayer layer[2];
  double count;           /* elements per channel of the producer's output (GLOBAL count under SyncBN) */
  float* partial;
  unsigned int* tickets;  /* as in VtbBnTrain */
  const struct VtbSyncBn* sync;
} VtbDgradBn;
int vtb_conv_dgrad_stats_rows(const VtbConv* c);
int vtb_conv_dgrad_panel_w(const VtbConv* c);
int vtb_conv_dgrad_bn(const VtbConv* c, const void* dy, int lddy, const void* wd, void* dx, int lddx, int accumulate,
                      const VtbDgradBn* bn, void* stream);
/* wgrad: dw_oihw (fp32, [cout][cin_real][k][k]) (+)= sum over pixels dy * im2col(x).
 * workspace: vtb_conv_wgrad_workspace_bytes(c) bytes of scratch. Deterministic (no atomics). */
int vtb_conv_wgrad(const VtbConv* c, const void* dy, int lddy, const void* x, int ldx, void* workspace,
                   float* dw_oihw, int cin_real, int accumulate, void* stream);
/* wgrad of two side-by-side units (see VtbBnTrain.split): rows [0, split) of the weight gradient go to dw_a, rows
 * [split, cout) to dw_b (both OIHW fp32, each indexed from its own channel 0). */
int vtb_conv_wgrad_pair(const VtbConv* c, const void* dy, int lddy, const void* x, int ldx, void* workspace,
                        float* dw_a, float* dw_b, int split, int cin_real, int accumulate, void* stream);

/* ---- BatchNorm2d training forward: replaces aten::native_batch_norm at components.py:36 ----
 * The conv epilogue leaves per-CTA partial sums; these calls finish the job.
 *  vtb_bn_stats_reduce : partial[rows][c][2] -> sums[c][2] (double). Only needed for SyncBN, where `sums`
 *                        is all-reduced across ranks before vtb_bn_finalize (configs/base.yaml:22).
 *  vtb_bn_finalize     : from `partial` (single GPU) or `sums` (exactly one non-NULL) and the element count
 *                        per channel (`count`, the GLOBAL N*H*W under SyncBN): mean, invstd = 1/sqrt(var+eps)
 *                        (biased var), scale = gamma*invstd, shift = beta - mean*scale, and the running-stat
 *                        update running = (1-momentum)*running + momentum*stat with UNBIASED var, plus
 *                        num_batches_tracked += 1 (int64). running_* / num_batches_tracked may be NULL.
 *  vtb_bn_eval_affine  : eval mode, scale/shift from the running statistics. */
int vtb_bn_stats_reduce(const float* partial, int rows, int c, double* sums, void* stream);
int vtb_bn_finalize(const float* partial, int rows, const double* sums, double count, int c, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
                    void* stream);
int vtb_bn_eval_affine(int c, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, void* stream);

/* ---- SyncBatchNorm (configs/base.yaml:22 `sync_batchnorm: true` -> torch.nn.SyncBatchNorm's per-layer
 * all_gather / all_reduce, torch/nn/modules/_functions.py:39-170) as ONE kernel per exchange over NVLink peer memory.
 * Every rank owns a zero-initialised buffer of vtb_bn_sync_buffer_bytes() that all peers have mapped
 * (peer_buffers[r] = rank r's buffer as seen from THIS process, e.g. torch.distributed._symmetric_memory buffer_ptrs).
 * The kernel reduces the local partial rows, pushes the fp64 sums to every peer, waits for all peers and finalises with
 * the GLOBAL element count (`count` = sum over ranks of N*H*W).  All ranks must issue the same sequence of calls.
 *  vtb_bn_sync_finalize     : forward, same outputs as vtb_bn_finalize.
 *  vtb_bn_sync_bwd_finalize : backward, same outputs as vtb_bn_bwd_finalize (dgamma/dbeta from LOCAL sums, coef from
 *                             global sums); local_scratch: 2*c doubles. */
#define VTB_SYNC_MAX_RANKS 8
#define VTB_SYNC_MAX_CHANNELS 2048
typedef struct VtbSyncBn {
  int rank, world;
  void* peer_buffers[VTB_SYNC_MAX_RANKS];
} VtbSyncBn;
size_t vtb_bn_sync_buffer_bytes(void);
int vtb_bn_sync_finalize(const float* partial, int rows, int c, const VtbSyncBn* sync, double count,
                         const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                         float* running_var, long long* num_batches_tracked, float* mean, float* invstd, float* scale,
                         float* shift, void* stream);
int vtb_bn_sync_bwd_finalize(const float* partial, int rows, int c, const VtbSyncBn* sync, double count,
                             float* dgamma, float* dbeta, int accumulate, float* coef, double* local_scratch,
                             void* stream);

/* out = [relu](y*scale + shift) [+ residual]: the BatchNorm normalise + nn.ReLU(inplace) of
 * components.py:36-39 and the post-activation residual add of darknet.py:28 / vovnet.py:60-61 in one pass.
 * `out` may be a channel slice of a concat buffer (replaces torch.cat, darknet.py:53 / vovnet.py:55). */
int vtb_bn_act(const void* y, int ldy, long long pixels, int c, const float* scale, const float* shift, int relu,
               const void* residual, int ldr, void* out, int ldo, void* stream);

/* ---- BatchNorm2d + ReLU backward: replaces threshold_backward + native_batch_norm_backward ----
 * dz = dout * [y*scale+shift > 0]  (ReLU mask recomputed from the saved conv output y)
 *  reduce  : partial[rows][c][2] = per-block sums of (dz, dz*xhat), rows = vtb_bn_bwd_rows(pixels, c)
 *  finalize: dgamma (+)= sum dz*xhat, dbeta (+)= sum dz (LOCAL sums: `local_sums` if given, else the reduced
 *            input), coef[c][2] = (sum dz, sum dz*xhat)/count. Pass `sums_out` to only publish the local
 *            double sums (SyncBN step 1: all-reduce them, then call again with `sums` = reduced values).
 *  apply   : dy = scale * (dz - coef0 - xhat*coef1), bf16. */
int vtb_bn_bwd_rows(long long pixels, int c);
int vtb_bn_bwd_reduce(const void* dout, int lddo, const void* y, int ldy, long long pixels, int c,
                      const float* scale, const float* shift, const float* mean, const float* invstd, int relu,
                      float* partial, void* stream);
int vtb_bn_bwd_finalize(const float* partial, int rows, const double* sums, const double* local_sums, double count,
                        int c, float* dgamma, float* dbeta, int accumulate, float* coef, double* sums_out,
                        void* stream);
int vtb_bn_bwd_apply(const void* dout, int lddo, const void* y, int ldy, long long pixels, int c, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int relu, const float* coef,
                     void* dy, int lddy, void* stream);

/* The three calls above in ONE cooperative launch: reduce -> grid barrier -> finalize (-> SyncBN exchange over NVLink
 * peer memory when `peers` != NULL; `count` is then the GLOBAL element count) -> apply.
 * partial: vtb_bn_bwd_fused_rows(pixels, c) * c * 2 floats of scratch (+ 2*c floats when peers != NULL);
 * sync: >= 256 zero-initialised uint32 owned by the caller (left zero); dgamma / dbeta may be NULL.  c % 16 == 0. */
int vtb_bn_bwd_fused_rows(long long pixels, int c);
int vtb_bn_bwd_fused(const void* dout, int lddo, const void* y, int ldy, long long pixels, int c, const float* scale,
                     const float* shift, const float* mean, const float* invstd, int relu, double count,
                     float* partial, float* dgamma, float* dbeta, int accumulate, unsigned int* sync, void* dy,
                     int lddy, const struct VtbSyncBn* peers, void* stream);

/* dst (+)= src on bf16 NHWC views: gradient fan-out of the residual add (darknet.py:28) when it cannot be
 * aliased, and injection of incoming feature-map gradients. */
int vtb_grad_add(void* dst, int ldd, const void* src, int lds, long long pixels, int c, int accumulate, void* stream);

/* The image's FIRST convolution (stems: darknet.py:74,109, vovnet.py:85) as a 1x1 GEMM over a gathered operand:
 * vtb_im2col_input writes out[n][ho][wo][kp] bf16, column t*c + ci = x[n][ci][ho*s-p+kh][wo*s-p+kw] (t = kh*k + kw, zero
 * padding, columns >= k*k*c zero); the convolution is then called with VtbConv{n, ho, wo, cin = kp, cout, 1, 1, 0} and a
 * weight packed by a VtbPackJob{cin_real = cin = c, kk = k*k, wf_ld = kp}.  Its weight gradient comes out in the same
 * (tap, ci) column order, [cout][k*k*c]; vtb_dw_from_col permutes it to OIHW.  Used when the image needs no gradient.
 * The host side uses it for k = 3, c = 3 (kp = 32), the case every stem of the path has and the one the parity tests cover. */
int vtb_im2col_input(const float* x, int n, int c, int h, int w, int k, int stride, int pad, void* out, int kp,
                     void* stream);
int vtb_dw_from_col(const float* dw_col, int cout, int c, int kk, float* dw_oihw, int accumulate, void* stream);

/* ---- feature-pyramid fuse: replaces nn.Upsample(scale_factor=2 | 0.5, mode="nearest") + the "sum" aggregate of the
 * reference necks (necks.py:16-20, 66, 69-79) ----
 * out[n,y,x,:] = (a ? a[n,y,x,:] : 0) + b[n, src(y), src(x), :];  up != 0: b is (hb, wb) = (h/2, w/2), src(i) = i >> 1
 * (top-down FPN); up == 0: h = hb/2, w = wb/2, src(i) = 2i (bottom-up path of PAN).  a == NULL: the plain resize (the
 * "concat" aggregate writes it into a channel slice).  vtb_resize2_add_bwd: gb (+)= the transposed resize of gout
 * (the gradient of `a` is gout itself). */
int vtb_resize2_add(const void* a, int lda, const void* b, int ldb, int n, int h, int w, int c, int hb, int wb, int up,
                    void* out, int ldo, void* stream);
int vtb_resize2_add_bwd(const void* gout, int ldg, int n, int h, int w, int c, void* gb, int ldgb, int hb, int wb, int up,
                        int accumulate, void* stream);

/* ---- input side of the training step: RandomMixup / RandomCutmix (reference extras.py:14-109, classifier.py:86-87) ----
 * out[i] = mix(x[i], x[i-1]) on NCHW fp32 batches (the reference pairs image i with the batch rolled by one).
 * params_device: six floats in DEVICE memory {mode, lambda, x1, y1, x2, y2} - mode 0: copy, 1: mixup
 * (x*lambda + x_prev*(1-lambda), rounded like the reference's two in-place multiplies and one add), 2: cutmix (the box
 * rows [y1,y2) x columns [x1,x2) comes from x_prev).  The decision and its parameters never visit the host. */
int vtb_mix_images(const float* x, float* out, int n, int c, int h, int w, const float* params_device, void* stream);

/* NCHW fp32 (the layout model(x) receives, tests/test_backbones.py:21) -> NHWC bf16, channels zero-padded
 * to cpad (the 3-channel image is stored with 16 channels for the tensor-core stem). */
int vtb_nchw_to_nhwc(const float* x, int n, int c, int h, int w, void* out, int cpad, void* stream);

/* ---- VoVNet-only ops ----
 * MaxPool2d(3, 2, 1) (vovnet.py:94): -inf padding, first maximum wins ties in backward. */
/* idx (optional, n*ho*wo*c bytes): the window position of the first maximum of every output element, i.e. what
 * aten::max_pool2d_with_indices keeps for backward; with it the backward needs neither x nor a 36-load recomputation. */
int vtb_maxpool3s2_fwd(const void* x, int ldx, int n, int h, int w, int c, void* out, int ldo, void* idx, void* stream);
int vtb_maxpool3s2_bwd(const void* x, int ldx, int n, int h, int w, int c, const void* dout, int lddo, void* dx,
                       int lddx, int accumulate, const void* idx, void* stream);
/* ESEBlock (vovnet.py:20-28): out = x * hardsigmoid(W * mean_hw(x) + b) [+ residual] (vovnet.py:58-61).
 * weight [c][c] fp32 (the (C,C,1,1) Conv2d weight), bias [c]; pool/z/gate: [n][c] fp32 saved for backward.
 * bwd scratch: 3*n*c floats. dweight/dbias are fp32 parameter gradients. */
int vtb_ese_fwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* bias,
                const void* residual, int ldr, void* out, int ldo, float* pool, float* z, float* gate, void* stream);
int vtb_ese_bwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* pool, const float* z,
                const float* gate, const void* dout, int lddo, void* dx, int lddx, int accumulate_dx, float* dweight,
                float* dbias, int accumulate_dw, float* scratch, void* stream);

/* ---- classifier head + loss of the training step (reference classifier.py:59-64, 92: nn.AdaptiveAvgPool2d(1), nn.Flatten,
 * nn.Linear(C, K), F.cross_entropy(label_smoothing)) on the last feature map f (NHWC bf16 view, hw pixels per image).
 * fwd: pooled [n][c] fp32 (spatial mean), logits [n][k] = pooled . weight^T + bias (fp32), row_loss [n], loss[0] = mean,
 *      dlogits [n][k] = d loss / d logits (saved for backward).  labels: int64 class indices.  4 launches.
 * bwd: dweight [k][c] (+)=, dbias [k] (+)=, df (NHWC bf16 view, may be NULL) = d loss / d f, all scaled by *gscale (the
 *      upstream gradient of the scalar loss; NULL = 1: it rides in the GEMMs' alpha).  scratch: n*k + n*c floats (only the
 *      last n*c are used).  4 launches.  The three GEMMs run fp32 FMA tiles (BM x 64 x 16, BM = 32 | 64). */
int vtb_head_ce_fwd(const void* f, int ldf, int n, int hw, int c, const float* weight, const float* bias, int k,
                    const long long* labels, float label_smoothing, float* pooled, float* logits, float* dlogits,
                    float* row_loss, float* loss, void* stream);
int vtb_head_ce_bwd(const float* pooled, const float* dlogits, const float* weight, int n, int hw, int c, int k,
                    const float* gscale, float* dweight, float* dbias, int accumulate, void* df, int lddf, float* scratch,
                    void* stream);

/* ---- fp32 parity mode ------------------------------------------------------------------------------------------------
 * BASELINE north_star: "forward feature maps and gradients within 1e-4 relative in fp32 mode" - the reference run WITHOUT
 * autocast (components.py:26-39 in fp32).  Same dataflow and view conventions as above, but every activation / gradient
 * view is fp32 (ld in elements, 4-byte aligned, any channel count), weights are read straight from the OIHW fp32 master
 * (no packing), contractions are FMA loops on the CUDA cores in a fixed order, statistics are accumulated in fp64.
 * BatchNorm finalisation reuses vtb_bn_finalize / vtb_bn_bwd_finalize with their `sums` (double) inputs, so SyncBN in this
 * mode is one all-reduce of `sums` between the two calls.  Selected from Python with vision_toolbox_b200.precision("fp32"). */
int vtb_f32_nchw_to_nhwc(const float* x, int n, int c, int h, int w, float* out, int cpad, void* stream);
/* aten::convolution / convolution_backward at components.py:26-35, fp32. cin_real <= c->cin as in vtb_pack_weight. */
int vtb_f32_conv_fprop(const VtbConv* c, const float* x, int ldx, const float* w_oihw, int cin_real, float* y, int ldy,
                       void* stream);
int vtb_f32_conv_dgrad(const VtbConv* c, const float* dy, int lddy, const float* w_oihw, int cin_real, float* dx, int lddx,
                       int accumulate, void* stream);
size_t vtb_f32_conv_wgrad_workspace_bytes(const VtbConv* c);
int vtb_f32_conv_wgrad(const VtbConv* c, const float* dy, int lddy, const float* x, int ldx, void* workspace, float* dw_oihw,
                       int cin_real, int accumulate, void* stream);
/* BatchNorm2d statistics (components.py:36): sums[c][2] = (sum y, sum y*y) in double; partial: vtb_f32_bn_rows(pixels, c)
 * * c * 2 doubles of scratch.  Follow with vtb_bn_finalize(NULL, 0, sums, ...). */
int vtb_f32_bn_rows(long long pixels, int c);
int vtb_f32_bn_stats(const float* y, int ldy, long long pixels, int c, double* partial, double* sums, void* stream);
/* out = [relu]((y - mean) * invstd * gamma + beta) [+ residual]  (components.py:36-39, darknet.py:28, vovnet.py:60-61) */
int vtb_f32_bn_act(const float* y, int ldy, long long pixels, int c, const float* mean, const float* invstd, const float* gamma,
                   const float* beta, int relu, const float* residual, int ldr, float* out, int ldo, void* stream);
/* BatchNorm2d + ReLU backward: sums[c][2] = (sum dz, sum dz*xhat) in double -> vtb_bn_bwd_finalize(NULL, 0, sums, ...) ->
 * dy = gamma * invstd * (dz - coef0 - xhat * coef1). */
int vtb_f32_bn_bwd_reduce(const float* dout, int lddo, const float* y, int ldy, long long pixels, int c, const float* mean,
                          const float* invstd, const float* gamma, const float* beta, int relu, double* partial, double* sums,
                          void* stream);
int vtb_f32_bn_bwd_apply(const float* dout, int lddo, const float* y, int ldy, long long pixels, int c, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, int relu, const float* coef, float* dy,
                         int lddy, void* stream);
/* fp32 twins of vtb_grad_add / vtb_maxpool3s2_* / vtb_ese_* (identical signatures; idx is required by the pool backward) */
int vtb_f32_grad_add(void* dst, int ldd, const void* src, int lds, long long pixels, int c, int accumulate, void* stream);
int vtb_f32_maxpool3s2_fwd(const void* x, int ldx, int n, int h, int w, int c, void* out, int ldo, void* idx, void* stream);
int vtb_f32_maxpool3s2_bwd(const void* x, int ldx, int n, int h, int w, int c, const void* dout, int lddo, void* dx, int lddx,
                           int accumulate, const void* idx, void* stream);
int vtb_f32_ese_fwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* bias, const void* residual,
                    int ldr, void* out, int ldo, float* pool, float* z, float* gate, void* stream);
int vtb_f32_ese_bwd(const void* x, int ldx, int n, int hw, int c, const float* weight, const float* pool, const float* z,
                    const float* gate, const void* dout, int lddo, void* dx, int lddx, int accumulate_dx, float* dweight,
                    float* dbias, int accumulate_dw, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VTB_H_ */
