/*
 * sivae.h -- C ABI of the B200-native Soft-IntroVAE training engine (libsivae_b200.so).
 *
 * The reference (taldatech/soft-intro-vae-pytorch) has no native code and no FFI: its hot path is the
 * Python function train_soft_intro_vae() calling torch.nn modules.  This header is therefore the
 * boundary a maintainer would bind with ctypes from the reference's own module
 * (see INTEGRATION.md); each entry point names the reference lines it replaces, relative to
 * soft_intro_vae/train_soft_intro_vae.py unless stated otherwise.
 *
 * Conventions
 *   - plain C types only; all tensor arguments are raw DEVICE pointers (fp32 unless noted) owned by the
 *     caller (torch allocations on the Python side).  The library never allocates or frees caller
 *     memory; its scratch lives in the caller-provided workspace (sivae_bind_workspace).  Two stated exceptions:
 *     sivae_jpeg_decode_batch reads the COMPRESSED files from host pointers, and the libraries resolved at run time
 *     (NCCL for the gradient all-reduce, nvJPEG for the opt-in decode) manage their own internal buffers.
 *   - every function returns 0 on success, <0 for an engine error, >0 for a cudaError_t; the text is
 *     available from sivae_last_error().  No C++ exception crosses the boundary.
 *   - all work is enqueued on the caller's stream (void* = cudaStream_t); nothing synchronises except
 *     where stated.  One engine per GPU/process; an engine is not thread-safe.
 *   - image tensors at the boundary are NCHW like the reference; inside the engine everything is NHWC.
 *   - net ids: 0 = encoder, 1 = decoder, 2 = target decoder (bootstrap variant only).
 */
#ifndef SIVAE_H_
#define SIVAE_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sivae_engine sivae_engine;

enum { SIVAE_NET_ENCODER = 0, SIVAE_NET_DECODER = 1, SIVAE_NET_TARGET = 2 };
/* conv_backend:
   AUTO (default) = tcgen05 with compensated 16-bit FORWARD operands: activations and filters are stored as bf16 hi + lo
     pairs ("split32": 16 significand bits, fp32 exponent range) and every product of a forward conv is three kind::f16 MMAs
     (lo*hi + hi*lo + hi*hi, fp32 accumulate) = 1.5x the tf32 tensor time; every logged scalar (ELBO / KL / exp-ELBO) stays
     within 1e-4 of the reference.  dgrad / wgrad of the residual blocks run kind::f16 on plain bf16 operands (gradients
     within 4e-3..8e-3 median of the fp64 oracle; SIVAE_BWD16=0: kind::tf32), the image-facing convs (stem, predict) kind::tf32.
     Needs every channel count % 32 == 0 and image_size % 16 == 0;
     otherwise AUTO degrades to TF32 below;
   SIMT = exact fp32 CUDA-core path with fp64-chunked accumulation (on-device reference);
   TCGEN05 = single-kernel entry points only: plain kind::tf32 tensor-core kernel or error;
   TC3X = every operand split into two tf32 parts, three tf32 convs per product (round-1 compensated mode, ~3x conv time);
   TF32 = round-1 default: kind::tf32 forward and backward (scalars 1e-4..2e-3 from the reference), SIMT where the shape
     is not tensor-core eligible */
enum { SIVAE_CONV_AUTO = 0, SIVAE_CONV_SIMT = 1, SIVAE_CONV_TCGEN05 = 2, SIVAE_CONV_TC3X = 3, SIVAE_CONV_TF32 = 4 };
enum { SIVAE_T_CONV = 0, SIVAE_T_BN_WEIGHT = 1, SIVAE_T_BN_BIAS = 2, SIVAE_T_LINEAR = 3, SIVAE_T_BIAS = 4 };

/* SoftIntroVAE(cdim, zdim, channels, image_size) -- ctor :173-184; Encoder :79-109; Decoder :126-159 */
typedef struct {
  int cdim, zdim, image_size, n_channels;
  int channels[16];
  int max_batch;     /* largest batch the workspace is sized for */
  int variant;       /* 0 = standard, 1 = bootstrap (soft_intro_vae_bootstrap/...:192-217) */
  int conv_backend;  /* SIVAE_CONV_* */
  int cond_dim;      /* 0 = unconditional; > 0: SoftIntroVAE(conditional=True, cond_dim): the encoder fc takes [features | cond]
                        rows, the decoder fc [z | cond] rows (:106-109, :139-143).  Inference entry points only
                        (sivae_encode_cond / sivae_decode_cond): the reference's training step passes no condition
                        (:559-561), so a conditional model cannot be trained by it either (F.linear shape error) */
} sivae_config;

/* beta_kl, beta_rec, beta_neg, gamma_r kwargs (:340-341) and scale = 1/(ch*image_size^2) (:456) */
typedef struct { float beta_kl, beta_rec, beta_neg, gamma_r, scale; } sivae_hyper;

/* One parameter tensor, in the reference's registration (= state_dict / optimiser) order. */
typedef struct {
  char name[96];        /* key relative to the net prefix, e.g. "main.res_in_16.conv1.weight" */
  int kind;             /* SIVAE_T_* */
  long long offset;     /* in floats, into the net's flat parameter buffer */
  long long numel;
  int shape[4];         /* logical (reference) shape: conv [Cout,Cin,kh,kw] (memory order Cout,kh,kw,Cin =
                           torch.channels_last), linear [out,in], vectors [C] */
  int ndim;
} sivae_tensor_info;

/* One BatchNorm2d: running_mean at bn_offset, running_var at bn_offset + channels (flat float buffer),
   num_batches_tracked at index `index` of the int64 buffer. */
typedef struct { char name[96]; int channels; long long bn_offset; int index; } sivae_bn_info;

const char* sivae_last_error(void);
int sivae_version(void);

/* model construction (:442) */
int sivae_create(const sivae_config* cfg, sivae_engine** out);
void sivae_destroy(sivae_engine* e);
int sivae_num_tensors(const sivae_engine* e, int net);
int sivae_tensor(const sivae_engine* e, int net, int i, sivae_tensor_info* out);
long long sivae_param_count(const sivae_engine* e, int net);
int sivae_num_bn(const sivae_engine* e, int net);
int sivae_bn(const sivae_engine* e, int net, int i, sivae_bn_info* out);
long long sivae_bn_floats(const sivae_engine* e, int net);
long long sivae_workspace_bytes(const sivae_engine* e);

/* memory hand-over: flat fp32 buffers of sivae_param_count floats each (params, grads, Adam m, Adam v),
   BN running stats (sivae_bn_floats floats) and num_batches_tracked (int64[sivae_num_bn]).
   grads/m/v may be NULL for a net that is never optimised (target decoder). */
int sivae_bind_net(sivae_engine* e, int net, float* params, float* grads, float* adam_m, float* adam_v,
                   float* bn_running, long long* bn_nbt);
int sivae_bind_workspace(sivae_engine* e, void* ws, long long bytes);
/* tell the engine that parameters of `net` were modified by the caller (load_state_dict, target copy) so
   derived operand copies (tf32-rounded / transposed filters) are refreshed before next use */
int sivae_params_changed(sivae_engine* e, int net);

/* Update-E half of the introspective iteration (:551-588): all forwards, the loss and the backward that
   fills the encoder's flat grad buffer.  real: [B,cdim,S,S] NCHW; noise: [B,z] (:547); eps: [3,B,z] the
   reparameterisation draws in the order of :560,:567,:568.  stats (device, 16 floats) receives
   [0]=loss_rec [1]=lossE_real_kl [2]=expelbo_rec [3]=expelbo_fake [4]=lossE [15]=nan flag
   [14]=bce domain flag (a reconstruction outside [0, 1], see sivae_set_recon_loss). */
int sivae_e_step(sivae_engine* e, const float* real_nchw, const float* noise, const float* eps, int batch,
                 const sivae_hyper* hp, float* stats, void* stream);
/* Update-D half (:591-623); reuses real, noise and z of the preceding sivae_e_step (:597-598).
   eps: [2,B,z] (:602,:605).  stats[5]=loss_rec [6]=lossD_rec_kl [7]=lossD_fake_kl [8]=loss_rec_rec
   [9]=loss_fake_rec [10]=lossD [15]=nan flag (the isnan guard of :625). */
int sivae_d_step(sivae_engine* e, const float* eps, const sivae_hyper* hp, float* stats, void* stream);
/* Opt-in (default off; env SIVAE_REUSE_DEC=1): the reference recomputes fake = decoder(noise) and rec = decoder(z) at
   :597-598 although the decoder has not changed since :557,:561 (only optimizer_e stepped in between).  With on != 0
   sivae_d_step takes both passes' activations from the preceding sivae_e_step and replays only their BatchNorm
   running_mean / running_var / num_batches_tracked updates, in the reference's order, from the saved batch statistics:
   losses, gradients, parameters and buffers stay bit-identical to the recomputing path.  Falls back to recomputing when
   the decoder parameters were touched between the two halves. */
int sivae_set_reuse_decoder_passes(sivae_engine* e, int on);
int sivae_get_reuse_decoder_passes(const sivae_engine* e);
/* recon_loss_type kwarg of train_soft_intro_vae (:339) = loss_type of every calc_reconstruction_loss call of the step
   (:268-294; :563,573,576,599,610,612; VAE warm-up :520).  MSE (default): per-sample sum of squared errors, 'mean' = mean over
   the batch (:282-287).  L1 / BCE: F.l1_loss / F.binary_cross_entropy on the [B, D] views (:288-291): 'mean' divides by B*D,
   the 'none' form is summed per sample by the caller (:574-578).  BCE needs reconstructions inside [0, 1] (the reference raises
   inside F.binary_cross_entropy otherwise): the step then sets stats[14] = 1 and the Python boundary raises RuntimeError. */
enum { SIVAE_LOSS_MSE = 0, SIVAE_LOSS_L1 = 1, SIVAE_LOSS_BCE = 2 };
int sivae_set_recon_loss(sivae_engine* e, int loss_type);
int sivae_get_recon_loss(const sivae_engine* e);
/* vanilla VAE warm-up step (:512-536): grads of encoder AND decoder. eps: [B,z].
   stats[11]=loss_rec [12]=loss_kl [13]=loss */
int sivae_vae_step(sivae_engine* e, const float* real_nchw, const float* eps, int batch, const sivae_hyper* hp,
                   float* stats, void* stream);
/* optim.Adam.step (:450-451, :589, :624): p -= lr/bc1 * m/(sqrt(v)/sqrt(bc2)+1e-8) on the flat buffers with
   g = grad*grad_scale (grad_scale = 1/world_size after the NCCL all-reduce). Keeps its own step counter. */
int sivae_adam_step(sivae_engine* e, int net, float lr, float grad_scale, void* stream);
int sivae_adam_set_step(sivae_engine* e, int net, long long step);
long long sivae_adam_get_step(const sivae_engine* e, int net);

/* Data parallel (the reference's only DP precedent is style_soft_intro_vae/launcher.py:26-33 + DistributedDataParallel,
   train_style_soft_intro_vae.py:154-161): one process per GPU, rank-local batches and BatchNorm statistics, ONE sum-all-reduce of
   the flat encoder gradient buffer after the E backward and ONE of the decoder buffer after the D backward.  The library owns the
   collective: raw ncclAllReduce on the step's own stream (capturable in a CUDA graph together with the kernels around it).
   NCCL is resolved at run time with dlopen("libnccl.so.2") -- inside a torch process that is torch's own copy.
     sivae_comm_unique_id: rank 0 creates the 128-byte ncclUniqueId, the caller ships it to the other ranks (any side channel);
     sivae_comm_init:      every rank, with its CUDA device current: ncclCommInitRank.  The communicator is process-global (one
                           process = one GPU): later engines attach to it with id128 == NULL; it is destroyed only by
     sivae_comm_finalize   (all ranks, after their last step; never implicitly -- an un-finalized process leaves it to the OS);
     sivae_comm_global_world: world size of the process-global communicator, 0 if none;
     sivae_allreduce_attach: alternative -- borrow an existing ncclComm_t (not destroyed by the engine); NULL detaches;
     sivae_allreduce_grads: in-place sum-all-reduce of the net's flat gradient buffer (no-op without a communicator);
     sivae_iteration:      E half, all-reduce, Adam(encoder, grad_scale 1/world), D half, all-reduce, Adam(decoder) in one call. */
int sivae_comm_unique_id(unsigned char* out128);
int sivae_comm_init(sivae_engine* e, const unsigned char* id128, int world_size, int rank);
int sivae_comm_global_world(void);
int sivae_comm_finalize(void);
int sivae_allreduce_attach(sivae_engine* e, void* nccl_comm);
int sivae_comm_world(const sivae_engine* e);
int sivae_allreduce_grads(sivae_engine* e, int net, void* stream);
int sivae_iteration(sivae_engine* e, const float* real_nchw, const float* noise, const float* eps5, int batch,
                    const sivae_hyper* hp, float lr_e, float lr_d, float* stats, void* stream);

/* model.encode / model.decode / model.sample (:203-223): mu, logvar: [B,z]; out: [B,cdim,S,S] NCHW.
   train != 0 uses batch statistics and moves the running stats like the reference in model.train(). */
int sivae_encode(sivae_engine* e, const float* x_nchw, int batch, float* mu, float* logvar, int train, void* stream);
int sivae_decode(sivae_engine* e, int net, const float* z, int batch, float* out_nchw, int train, void* stream);
/* the same for a conditional model (sivae_config.cond_dim > 0): Encoder.forward(x, o_cond) (:116-122) / Decoder.forward(z, y_cond)
   (:161-169); cond: [B, cond_dim] device rows concatenated to the fc input.  The plain entry points above (and every training
   step) FAIL on a conditional engine, as the reference's fc layers reject the un-concatenated input. */
int sivae_encode_cond(sivae_engine* e, const float* x_nchw, const float* cond, int batch, float* mu, float* logvar, int train,
                      void* stream);
int sivae_decode_cond(sivae_engine* e, int net, const float* z, const float* cond, int batch, float* out_nchw, int train,
                      void* stream);

/* decoder outputs of the last half step as NCHW [B,cdim,S,S]: slot 0 = fake, 1 = rec, 2 = rec_rec, 3 = rec_fake
   (what the reference keeps in the Python variables of the same names, used for the sample grid :641-646) */
int sivae_last_image(sivae_engine* e, int slot, float* out_nchw, void* stream);
int sivae_last_batch(const sivae_engine* e);
/* a CUDA-graph replay of a step does not run the library's host code: the caller records the replayed step's batch size */
int sivae_set_last_batch(sivae_engine* e, int batch);

/* instrumentation for bench.py: number of kernels this library has enqueued so far; optional CUDA-event timing of
   every convolution launch on its stream, summed per kernel class by sivae_profile_read (which synchronises):
   out[class*3 + {0,1,2}] = {milliseconds, algorithmic FLOPs, launches}; class 0 = tcgen05 kind::tf32 conv fwd/dgrad,
   1 = tcgen05 kind::tf32 wgrad, 2 = CUDA-core conv fwd/dgrad, 3 = CUDA-core wgrad, 4 = fused loss pass (the FLOP slot holds
   its algorithmic BYTES: five images read once), 5 = tcgen05 forward conv on split32 operands (3 kind::f16 MMAs per product),
   6 = tcgen05 bf16 dgrad, 7 = tcgen05 bf16 wgrad; out must hold 24 doubles */
unsigned long long sivae_launch_count(void);
int sivae_profile_enable(int on);
int sivae_profile_read(double* out);
/* text table "class N H W Cin Cout k launches ms gflop" per distinct conv shape; call before sivae_profile_read */
int sivae_profile_dump(char* buf, int len);

/* ---- single-kernel entry points (unit parity tests; NHWC activations, [Cout][kh][kw][Cin] filters) ---- */
/* nn.Conv2d(k, stride 1, pad k/2) forward (:51,56,60,89,159); bias/addend may be NULL; y = conv + bias + addend */
int sivae_conv2d_fwd(const float* x, const float* w, const float* bias, const float* addend, float* y,
                     int N, int H, int W, int Cin, int Cout, int ksize, int backend, void* stream);
/* dgrad of the same conv: dx = conv_transpose(dy, w) + addend */
int sivae_conv2d_dgrad(const float* dy, const float* w, const float* addend, float* dx,
                       int N, int H, int W, int Cin, int Cout, int ksize, int backend, void* workspace,
                       long long ws_bytes, void* stream);
/* wgrad: dw (+)= sum_pixels dy (x) x ; accumulate != 0 adds to dw */
int sivae_conv2d_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                       int ksize, int accumulate, int backend, void* workspace, long long ws_bytes, void* stream);
/* train-mode BatchNorm2d + LeakyReLU(0.2) [+ residual add] [+ AvgPool2d(2) | nearest x2] (:65-75, :90-92,98,155)
   mode 0 none, 1 pool, 2 upsample.  Writes mean/invstd (2*C floats) and updates running stats. */
int sivae_bn_act_fwd(const float* t, const float* identity, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, long long* nbt, float* mean_invstd, float* out,
                     int N, int H, int W, int C, int mode, int train, void* workspace, long long ws_bytes, void* stream);
/* its backward: dt (and g = grad w.r.t. the identity branch, may be NULL), dgamma/dbeta (may be NULL) */
int sivae_bn_act_bwd(const float* dout, const float* t, const float* identity, const float* gamma, const float* beta,
                     const float* mean_invstd, float* dt, float* g, float* dgamma, float* dbeta, int accumulate,
                     int N, int H, int W, int C, int mode, void* workspace, long long ws_bytes, void* stream);
/* the same pair with the forward's SIGN BYTES: sign_mask (N*H*W*C/4 bytes, one per float4 of t; bit j = pre-activation of
   component j was positive) is written by the forward; given to the backward, neither of its passes reads `identity`
   (6.1 instead of 8 full-tensor passes for a residual block's second BatchNorm).  NULL = the plain functions above. */
int sivae_bn_act_fwd_m(const float* t, const float* identity, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, long long* nbt, float* mean_invstd, float* out,
                       int N, int H, int W, int C, int mode, int train, void* workspace, long long ws_bytes,
                       unsigned char* sign_mask, void* stream);
int sivae_bn_act_bwd_m(const float* dout, const float* t, const float* identity, const float* gamma, const float* beta,
                       const float* mean_invstd, float* dt, float* g, float* dgamma, float* dbeta, int accumulate,
                       int N, int H, int W, int C, int mode, void* workspace, long long ws_bytes,
                       const unsigned char* sign_mask, void* stream);
/* fused loss pass (:563-586 / :599-620): per-sample squared-error sums of (rec-real), (rec_rec-rec),
   (rec_fake-fake) in one read of the five images; out: [B,3] */
int sivae_mse3(const float* real, const float* rec, const float* rec_rec, const float* fake, const float* rec_fake,
               float* out, int batch, long long per_sample, void* workspace, long long ws_bytes, void* stream);
/* calc_kl(reduce='none') (:231-251) and reparameterize (:254-265) on [B,2z] encoder outputs */
int sivae_kl_reparam(const float* mu_logvar, const float* eps, float* z, float* kl, int batch, int zdim, void* stream);
int sivae_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float grad_scale,
                    long long step, void* stream);
/* nn.Linear (encoder fc :109,:121; decoder fc + ReLU :145-148,:166-167): y[B,O] = x[B,F] . w[O,F]^T + b (relu != 0: then
   ReLU), and its input gradient dx[B,F] = dy[B,O] . w[O,F] (workspace: sivae_linear_dgrad_workspace_bytes) */
int sivae_linear_fwd(const float* x, const float* w, const float* b, float* y, int batch, int in_features, int out_features,
                     int relu, void* stream);
int sivae_linear_dgrad(const float* dy, const float* w, float* dx, int batch, int in_features, int out_features,
                       void* workspace, long long ws_bytes, void* stream);
long long sivae_linear_dgrad_workspace_bytes(int batch, int in_features, int out_features);


/* ---- image batch assembly: the loader that feeds the step -------------------------------------
   soft_intro_vae/dataset.py:12-47 load_image as the image configs call it (train_soft_intro_vae.py:388-392, 400-404,
   415-417: input_height=None, crop_height=None, output_height=S, is_mirror=True) + transforms.ToTensor (:66-68, 75):
   ImageOps.mirror -> Image.resize((S,S), Image.BICUBIC) -> uint8/255.  Bit-exact with Pillow's fixed-point resampler
   (Resample.c); JPEG/PNG decoding stays with the caller.
   sivae_resample_coeffs: Pillow's precompute_coeffs + normalize_coeffs_8bpc for one axis (host only, no GPU needed):
   bounds [out_size][2] = (first source index, tap count), kk [out_size][*ksize] 22-bit fixed-point taps.
   A "plan" is the pair of coefficient tables of one (in_h,in_w)->(out_h,out_w) geometry in caller-owned device memory
   (sivae_image_plan_bytes / _init once per geometry).  sivae_image_batch_u8: src_hwc [B,in_h,in_w,C] decoded 8-bit
   pixels (C = 1 or 3, 4-byte aligned base), mirror [B] bytes (NULL = none), out_nchw [B,C,out_h,out_w] float32. */
int sivae_resample_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk, long long kk_capacity);
long long sivae_image_plan_bytes(int in_h, int in_w, int out_h, int out_w);
int sivae_image_plan_init(int in_h, int in_w, int out_h, int out_w, void* plan_dev, long long plan_bytes, void* stream);
int sivae_image_batch_u8(const unsigned char* src_hwc, const unsigned char* mirror, int batch, int in_h, int in_w,
                         int channels, int out_h, int out_w, const void* plan_dev, float* out_nchw, void* stream);
/* load_image in full (dataset.py:12-47: mirror -> [resize to input size, :29-30] -> [ImageOps.crop, :32-44] -> resize to the
   output size, :46): the resize reads the win_h x win_w window whose origin (x, y) inside each [src_h, src_w] source image is
   win_xy[b] (device int32 [B][2]; NULL = the whole image) -- Pillow resamples the cropped COPY, so nothing outside the window
   contributes -- mirrors it first when mirror[b] != 0, and writes EITHER out_nchw (float32 [B,C,out_h,out_w] = ToTensor) OR
   out_u8_hwc (uint8 [B,out_h,out_w,C]: the intermediate image of the two-stage resize, fed to a second call).  plan = that of
   (win_h, win_w) -> (out_h, out_w).  A mirror BEFORE a crop (:26-27 precede :32) is the crop window with its left / right
   margins swapped, mirrored. */
int sivae_image_batch_u8_ex(const unsigned char* src_hwc, const unsigned char* mirror, const int* win_xy, int batch, int src_h,
                            int src_w, int channels, int win_h, int win_w, int out_h, int out_w, const void* plan_dev,
                            float* out_nchw, unsigned char* out_u8_hwc, void* stream);

/* RGB JPEG files (dataset.py:20-24: Image.open(file), mode RGB) decoded by nvJPEG -- the CUDA toolkit's LIBRARY,
   resolved at run time with dlopen("libnvjpeg.so.12") (SIVAE_NVJPEG_LIB overrides) -- straight into the device batch
   [B,height,width,3] that sivae_image_batch_u8 reads: Huffman decoding on the host inside nvJPEG,
   IDCT / chroma upsampling / colour conversion on the GPU in stream order; returns after the stream has consumed the
   compressed bytes.  data[i] / lengths[i]: HOST pointers and sizes of the compressed files, all of the same height x width.
   OPT-IN and NOT bit-exact with the reference: Pillow decodes with libjpeg-turbo, whose IDCT / colour rounding and "fancy"
   4:2:0 chroma upsampling differ from nvJPEG's by a few grey levels; the default loader therefore decodes with Pillow on the host.
   Grey / CMYK files and is_gray data sets stay on the Pillow path.
   -9: nvJPEG unavailable; -8: not a 3-component JPEG nvJPEG decodes, or size mismatch.  sivae_jpeg_info: header only. */
int sivae_jpeg_info(const unsigned char* data, long long length, int* height, int* width, int* components);
int sivae_jpeg_decode_batch(const unsigned char* const* data, const long long* lengths, int batch, int height, int width,
                            unsigned char* out_hwc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIVAE_H_ */
