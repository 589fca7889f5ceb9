// Internal launch API shared by the translation units of libsivae_b200.so.  All pointers are device
// pointers, all launches go to `st`.  NHWC activations, [Cout][kh][kw][Cin] filters.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <string>

namespace sivae {

constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;
constexpr float kSlope = 0.2f;

extern unsigned long long g_launches;

enum ResampleMode { RS_NONE = 0, RS_POOL = 1, RS_UP = 2 };

struct ConvShape {
  int N, H, W, Cin, Cout, k;   // stride 1, pad k/2
  long long pixels() const { return (long long)N * H * W; }
  long long ktot() const { return (long long)k * k * Cin; }
};

// ---------------- layout ----------------
void launch_nchw_to_nhwc(const float* in, float* out, int N, int C, int H, int W, cudaStream_t st);
void launch_nhwc_to_nchw(const float* in, float* out, int N, int C, int H, int W, cudaStream_t st);
void launch_fill(float* p, float v, long long n, cudaStream_t st);
void launch_round_tf32(const float* in, float* out, long long n, cudaStream_t st);
// compensated (3xTF32) operand split: hi = tf32(x), lo = tf32(x - hi); hi may alias in (16-byte aligned pointers)
void launch_split_tf32(const float* in, float* hi, float* lo, long long n, cudaStream_t st);
// fp32 -> split32 (split32.cuh); n floats, n % 32 == 0, groups of 32 aligned with the tensor's channel groups.  out may alias in.
void launch_split32(const float* in, float* out, long long n, cudaStream_t st);
constexpr int FMT_TF32_ = 0, FMT_SPLIT_ = 1, FMT_BF16_ = 2;   // operand formats of the tensor-core forward / dgrad kernels (mirrors conv_tc.cu)
// wd[ci][k-1-r][k-1-s][co] = w[co][r][s][ci]  (filters for dgrad as a forward conv); optional tf32 rounding
void launch_pack_dgrad_filter(const float* w, float* wd, int Cout, int Cin, int k, bool round_tf32, cudaStream_t st);
void launch_pack_dgrad_filter_bf16(const float* w, void* wd_bf16, int Cout, int Cin, int k, cudaStream_t st);
// fp32 -> plain bf16 (round to nearest even); n % 4 == 0
void launch_to_bf16(const float* in, void* out, long long n, cudaStream_t st);

// ---------------- SIMT fp32 implicit-GEMM convolution (exact path; also stem / predict shapes) ----------
// y[p][co] = sum_{tap,ci} x[p+tap][ci] * w[co][tap][ci] + bias[co] + addend[p][co]
void launch_conv_fwd_simt(const float* x, const float* w, const float* bias, const float* addend, float* y,
                          const ConvShape& s, cudaStream_t st);
// dw[co][tap][ci] (+)= sum_p dy[p][co] * x[p+tap][ci];  scratch: split partials
size_t conv_wgrad_simt_scratch_bytes(const ConvShape& s);
void launch_conv_wgrad_simt(const float* x, const float* dy, float* dw, const ConvShape& s, bool accumulate,
                            void* scratch, size_t scratch_bytes, cudaStream_t st);
// out[c] (+)= sum_rows in[row][c]
size_t colsum_scratch_bytes(int C);
void launch_colsum(const float* in, float* out, long long rows, int C, bool accumulate, void* scratch, cudaStream_t st);

// ---------------- cdim-facing 5x5 convolutions on CUDA cores (conv_narrow.cu) ----------------
// row-separable tensor-core form of the image-facing 5x5 convs (conv_tc.cu): c <= 3 channels on the input or output side
bool conv_rowsep_in_supported(const ConvShape& s);
bool conv_rowsep_out_supported(const ConvShape& s);
long long conv_rowsep_scratch_floats(const ConvShape& s);
// round_tf32 = false: unrounded values (to be converted to split32 by launch_split32)
void launch_rowsep_filter_expand(const float* f, float* out, int Co, int c, cudaStream_t st, bool round_tf32 = true);     // [Co][5][5][c] -> [Co][5][32]
void launch_rowsep_filter_gather(const float* f, float* out, int c, int Ci, cudaStream_t st, bool round_tf32 = true);     // [c][5][5][Ci] -> [16][5][Ci]
// fmt (conv_tc.cu): FMT_TF32 = 0 (x, filter fp32 pre-rounded to tf32) or FMT_SPLIT = 1 (x, filter in the split32 format)
int launch_conv_rowsep_in(const float* x, const float* we, const float* bias, const float* addend, float* y, const ConvShape& s,
                          float* scratch, float* stats, cudaStream_t st, int fmt = 0);
bool conv_rowsep_wgrad_supported(int H, int W, int c, int wide, int k);
size_t conv_rowsep_wgrad_scratch_bytes(int N, int H, int W, int wide);
// mode 1: stem (narrow = x [N,H,W,c], wide = dy [N,H,W,wide]); mode 0: predict (narrow = dy, wide = x); dw [Cout][5][5][Cin]
int launch_conv_rowsep_wgrad(const float* narrow_t, const float* wide_t, float* dw, int N, int H, int W, int c, int wide, int mode,
                             bool accumulate, float* expand_scratch, void* scratch, size_t scratch_bytes, cudaStream_t st);
int launch_conv_rowsep_out(const float* x, const float* wg, const float* bias, const float* addend, float* y, const ConvShape& s,
                           float* scratch, cudaStream_t st, int fmt = 0);
bool conv_narrow_in_supported(const ConvShape& s);            // Cin <= 4: stem forward, predict dgrad
// wt = filter transposed to [tap][Cin][Cout] (launch_narrow_transpose of the [Cout][tap][Cin] filter)
void launch_narrow_transpose(const float* w, float* wt, int Cout, int k, int A, cudaStream_t st);
void launch_conv_narrow_in_fwd(const float* x, const float* wt, const float* bias, const float* addend, float* y,
                               const ConvShape& s, cudaStream_t st);
bool conv_narrow_corr_supported(int wideC, int narrowA, int k);
size_t conv_narrow_corr_scratch_bytes(int wideC, int k);
// mode 0: dw[a][tap][c] (predict wgrad: nar = dy, wide = x); mode 1: dw[c][tap][a] (stem wgrad: nar = x, wide = dy)
void launch_conv_narrow_corr(const float* nar, const float* wide, float* dw, int N, int H, int W, int A, int C, int k, int mode,
                             bool accumulate, void* scratch, size_t scratch_bytes, cudaStream_t st);

// ---------------- tcgen05 TF32 implicit-GEMM convolution (conv_tc.cu) ----------------
bool conv_tc_supported_fwd(const ConvShape& s, int fmt = 0);
// returns cudaError / driver error code (0 ok)
// fmt = FMT_SPLIT: x and w are split32 tensors ([32 bf16 hi | 32 bf16 lo] per 32-channel group, split32.cuh) and every
// product is computed as lo*hi + hi*lo + hi*hi with three kind::f16 MMAs -- fp32-class accuracy at 1.5x the tf32 tensor time
int launch_conv_fwd_tc(const float* x, const float* w, const float* bias, const float* addend, float* y,
                       const ConvShape& s, cudaStream_t st, float* stats = nullptr, void* scratch = nullptr,
                       size_t scratch_bytes = 0, int fmt = 0);
// split-K partial tensor the v2 kernel wants for few-tile layers (0: no split for this shape)
size_t conv_tc_splitk_scratch_bytes(const ConvShape& s);
// fused train-mode BN statistics: if > 0, passing `stats` ([parts][2][Cout] floats) to launch_conv_fwd_tc makes the
// epilogue emit per-warp column sums / sums of squares; finish with launch_bn_stats_from_parts
int conv_tc_stats_parts(const ConvShape& s);
bool conv_tc_supported_wgrad(const ConvShape& s);
size_t conv_wgrad_tc_scratch_bytes(const ConvShape& s);
int launch_conv_wgrad_tc(const float* x, const float* dy, float* dw, const ConvShape& s, bool accumulate,
                         void* scratch, size_t scratch_bytes, cudaStream_t st);
// bf16 wgrad (k = 1 or 3, Cin % 64 == 0, Cout % 64 == 0): x, dy plain bf16 NHWC tensors, dw fp32
bool conv_tc_supported_wgrad16(const ConvShape& s);
size_t conv_wgrad_tc16_scratch_bytes(const ConvShape& s);
int launch_conv_wgrad_tc16(const void* x, const void* dy, float* dw, const ConvShape& s, bool accumulate,
                           void* scratch, size_t scratch_bytes, cudaStream_t st);

// ---------------- train-mode BatchNorm (+LeakyReLU, +residual, +pool / upsample) ----------------
size_t bn_scratch_bytes(long long rows, int C);
// batch statistics of t[rows][C] -> mean_invstd[0..C) = mean, [C..2C) = invstd; running-stat EMA + nbt++
void launch_bn_stats(const float* t, long long rows, int C, float* mean_invstd, float* running_mean,
                     float* running_var, long long* nbt, void* scratch, size_t scratch_bytes, cudaStream_t st,
                     double* ema_save = nullptr);
size_t bn_parts_scratch_bytes(int nparts, int C);
void launch_bn_stats_from_parts(float* part, int nparts, long long rows, int C, float* mean_invstd, float* running_mean,
                                float* running_var, long long* nbt, cudaStream_t st, double* ema_save = nullptr);
// ema_save ([mean C | unbiased var C] doubles, nullable) receives the exact EMA inputs of the update; replaying them with
// launch_bn_ema_replay over a whole net (`ema` / `running` in the net's BN buffer layout, n floats, n_bn layers) applies
// the running-statistics side effect of one more identical train-mode forward without recomputing it
void launch_bn_ema_replay(const double* ema, float* running, long long n, long long* nbt, int n_bn, cudaStream_t st);
// eval-mode: mean_invstd from running stats
void launch_bn_eval_stats(const float* running_mean, const float* running_var, int C, float* mean_invstd, cudaStream_t st);
// out = resample(lrelu(bn(t) + identity)); identity may be null. (N,H,W) are the dims of t.
// sign_mask (nullable): receives one byte per float4 of t with the signs of the pre-activation (bit j = component j > 0)
// outs (nullable; needs C % 32 == 0): the output in the split32 operand format (from the unrounded value); with outs the
// fp32 copy `out` is optional (null = not written) and tf32-rounded.  idn_split: `identity` is a split32 tensor.
void launch_bn_act_fwd(const float* t, const float* identity, const float* mean_invstd, const float* gamma,
                       const float* beta, float* out, int N, int H, int W, int C, int mode, bool round_tf32,
                       cudaStream_t st, unsigned char* sign_mask = nullptr, float* outs = nullptr, bool idn_split = false,
                       void* outh = nullptr);      // outh (with outs only): plain bf16 copy of the output = the bf16 wgrad operand
// backward. dout has the shape of the resampled output.  sums: 2*C floats scratch inside `scratch`.
void launch_bn_act_bwd(const float* dout, const float* t, const float* identity, const float* mean_invstd,
                       const float* gamma, const float* beta, float* dt, float* g, float* dgamma, float* dbeta,
                       bool accumulate, int N, int H, int W, int C, int mode, bool round_tf32, void* scratch,
                       size_t scratch_bytes, cudaStream_t st, const unsigned char* sign_mask = nullptr, int out16 = 0);
// out16: bit 0 = dt, bit 1 = g are written as plain bf16 tensors (operands of the bf16 dgrad / wgrad kernels)
// with sign_mask (written by launch_bn_act_fwd) neither pass reads `identity`: 6.1 instead of 8 full-tensor passes

// ---------------- linear ----------------
// y[B][O] = act(x[B][F] . w[O][F]^T + b[O]);  relu optional
void launch_linear_fwd(const float* x, const float* w, const float* b, float* y, int B, int F, int O, bool relu,
                       cudaStream_t st);
// dx[B][F] = dy[B][O] . w[O][F]
size_t linear_dgrad_scratch_bytes(int B, int F, int O);
void launch_linear_dgrad(const float* dy, const float* w, float* dx, int B, int F, int O, void* scratch, cudaStream_t st);
// dw[O][F] (+)= dy^T . x ; db[O] (+)= sum_b dy
void launch_linear_wgrad(const float* x, const float* dy, float* dw, float* db, int B, int F, int O, bool accumulate,
                         cudaStream_t st);
// dy *= (y > 0)  (ReLU backward through the saved post-activation)
void launch_relu_bwd(const float* y, float* dy, long long n, cudaStream_t st);

// ---------------- losses ----------------
// recon_loss_type of calc_reconstruction_loss (:268-294); = SIVAE_LOSS_* of include/sivae.h
enum { SIVAE_LOSS_MSE_ = 0, SIVAE_LOSS_L1_ = 1, SIVAE_LOSS_BCE_ = 2 };
size_t mse3_scratch_bytes(int B, long long per_sample);
// per-sample SUMS over the image of the element error (mse: squared, l1: absolute, bce: binary cross entropy) for the pairs
// (real, rec), (rec, rec_rec), (fake, rec_fake) = (target, reconstruction); bad_flag (bce, nullable): set to 1 if a
// reconstruction lies outside [0, 1] (F.binary_cross_entropy raises there)
void launch_mse3(const float* real, const float* rec, const float* rec_rec, const float* fake, const float* rec_fake,
                 float* out /*[B][3]*/, int B, long long per_sample, void* scratch, size_t scratch_bytes, cudaStream_t st,
                 int loss_type = SIVAE_LOSS_MSE_, int* bad_flag = nullptr);
// mu_logvar [B][2z]; z = mu + eps*exp(.5 lv); kl[b] = -.5 sum(1 + lv - mu^2 - e^lv)
void launch_kl_reparam(const float* mu_logvar, const float* eps, float* z, float* kl, int B, int zdim, cudaStream_t st);
// d(mu_logvar)[B][2z] = dz-path + ckl[b]*KL-path.   dz may be null (no decoder path), ckl may be null
void launch_latent_bwd(const float* mu_logvar, const float* eps, const float* dz, const float* ckl, float ckl_const,
                       float* dml, int B, int zdim, cudaStream_t st);

struct StepCoefs;   // device-side per-sample coefficients, see loss kernels
// E-step scalar assembly (:563-586): consumes mse[B][3] and kl_real/kl_rec/kl_fake [B] each
void launch_e_loss_finalize(const float* mse, const float* kl_real, const float* kl_rec, const float* kl_fake, int B,
                            float beta_kl, float beta_rec, float beta_neg, float scale, float* stats,
                            float* coef /*[4][B]: c_rec, c_rr, c_rf, (unused)*/, float* ckl_rec, float* ckl_fake,
                            cudaStream_t st, float mean_div = 1.f, const int* bad_flag = nullptr);
// mean_div: factor on the batch mean of the per-sample sums for the reduction='mean' terms: 1 (mse, :282-287) or
// 1 / (cdim*S*S) (l1 / bce: F.*_loss(reduction='mean') divides by B*D, :288-291); bad_flag -> stats[14]
// D-step scalar assembly (:599-620)
void launch_d_loss_finalize(const float* mse, const float* kl_rec, const float* kl_fake, int B, float beta_kl,
                            float beta_rec, float gamma_r, float scale, float* stats, cudaStream_t st, float mean_div = 1.f,
                            const int* bad_flag = nullptr);
// vae-step scalar assembly (:520-523)
void launch_vae_loss_finalize(const float* mse, const float* kl, int B, float beta_kl, float beta_rec, float* stats,
                              cudaStream_t st, float mean_div = 1.f, const int* bad_flag = nullptr);
// gradient seeds on images (NHWC, per_sample floats each):
//   d_rec      = a_rec[b]*(rec-real) + a_t[b]*(rec_rec-rec)*(-1)      (a_* may be per-sample arrays or constants)
//   d_rec_rec  = a_t[b]*(rec_rec-rec)
//   d_rec_fake = a_f[b]*(rec_fake-fake);   d_fake = -a_f[b]*(rec_fake-fake) if d_fake != null
// a_rec is a constant; a_t / a_f are per-sample arrays when *_arr != null else constants.
// l1 / bce: (r - x) is replaced by half the element derivative w.r.t. the reconstruction, and the target-side terms (the
// "* (-1)" above) by half the derivative w.r.t. the target (see rec_g / rec_gx in kernels.cu)
void launch_loss_seed(const float* real, const float* rec, const float* rec_rec, const float* fake,
                      const float* rec_fake, float a_rec, const float* a_t_arr, float a_t, const float* a_f_arr,
                      float a_f, bool target_grad_rec, float* d_rec, float* d_rec_rec, float* d_rec_fake, float* d_fake,
                      int B, long long per_sample, cudaStream_t st, int loss_type = SIVAE_LOSS_MSE_);

// ---------------- optimiser ----------------
void launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float grad_scale,
                 float b1, float b2, float eps, long long* step_dev, float* coef_dev, cudaStream_t st);

// ---------------- image batch assembly (image.cu): mirror + Pillow-exact bicubic resize + ToTensor ----------------

// ---------------- JPEG decode through nvJPEG (jpeg.cu; opt-in, not bit-exact with Pillow) ----------------
int jpeg_decode_batch(const unsigned char* const* data, const long long* lengths, int batch, int height, int width,
                      unsigned char* out_hwc, cudaStream_t st, std::string* msg);
int jpeg_info(const unsigned char* data, long long length, int* height, int* width, int* components, std::string* msg);
}  // namespace sivae
#include <vector>
namespace sivae {
int resample_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk);   // returns ksize
size_t image_plan_bytes(int in_h, int in_w, int out_h, int out_w);
int image_plan_init(int in_h, int in_w, int out_h, int out_w, void* plan_dev, cudaStream_t st);
// src [B][src_h][src_w][ch]; the in_h x in_w window at win_xy[b] (nullable: whole image) is mirrored (flag) and resized to
// out_h x out_w; writes dst (float32 NCHW = ToTensor) or dst_u8 (8-bit NHWC), exactly one non-null
int launch_image_batch(const unsigned char* src, const unsigned char* mirror, int B, int in_h, int in_w, int ch, int out_h,
                       int out_w, const void* plan_dev, float* dst, cudaStream_t st, int src_h = 0, int src_w = 0,
                       const int* win_xy = nullptr, unsigned char* dst_u8 = nullptr);

// ---------------- JPEG decode through nvJPEG (jpeg.cu; opt-in, not bit-exact with Pillow) ----------------
int jpeg_decode_batch(const unsigned char* const* data, const long long* lengths, int batch, int height, int width,
                      unsigned char* out_hwc, cudaStream_t st, std::string* msg);
int jpeg_info(const unsigned char* data, long long length, int* height, int* width, int* components, std::string* msg);

}  // namespace sivae
