// Native runtime of the Soft-IntroVAE step: model description (parameter / buffer layout in the reference's
// registration order), workspace carving, the forward / hand-written backward traversal of encoder and decoder
// passes, the E-step / D-step / VAE-step graphs of train_soft_intro_vae.py:512-624 and the C ABI (include/sivae.h).
#include "../../include/sivae.h"
#include "kernels.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>      // types and enums only: the functions are resolved with dlsym (no link-time dependency on NCCL)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace sivae;

static thread_local std::string g_err;
extern "C" const char* sivae_last_error(void) { return g_err.c_str(); }
extern "C" int sivae_version(void) { return 200; }   // round 2: sivae_config grew (cond_dim), new entry points

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CHECK_CUDA_RET()                                                             \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) return fail((int)_e, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

namespace {

struct Conv { int cin = 0, cout = 0, k = 0; long long w_off = -1, b_off = -1, wd_off = -1, wr_off = -1, wn_off = -1, wg_off = -1,
              wl_off = -1, wdl_off = -1;   /* 3xTF32 mode: low parts of the forward / dgrad-packed filter */
              long long wdh_off = -1;      /* plain bf16 form of the dgrad-packed filter (bf16 backward of the residual blocks) */ };
struct Bn { int c = 0; long long g_off = -1, b_off = -1, rm_off = -1; int idx = -1; };
struct Lin { int fin = 0, fout = 0; long long w_off = -1, b_off = -1; };
struct Block { int inc = 0, outc = 0, size = 0, mode = RS_NONE; bool expand = false; Conv ce, c1, c2; Bn bn1, bn2; };

struct Net {
  int id = 0;
  bool enc = false;
  Conv stem; Bn stem_bn;    // encoder only
  Conv predict;             // decoder only
  Lin fc;
  std::vector<Block> blocks;
  std::vector<sivae_tensor_info> tinfo;
  std::vector<sivae_bn_info> binfo;
  long long n_params = 0, bn_floats = 0, derived_floats = 0;
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr, *bn = nullptr, *derived = nullptr;
  long long* nbt = nullptr;
  bool dirty = true;
  bool comp = false;               // 3xTF32 mode: keep the low parts of the filters in `derived`
  long long* step_dev = nullptr;   // device-side Adam step counter + 2 floats of bias-correction scratch (workspace)
  float* coef_dev = nullptr;
  long long step_pending = 0;      // value to load into step_dev at the next workspace bind
  bool present = false;
};

// a1s / outs / xs (split-forward mode): the split32 copies (split32.cuh) of a1 / out / the block input that the FORWARD convs
// consume; the fp32 tensors of the same name are the tf32-rounded wgrad operands, written only in passes whose net gets
// parameter gradients
struct BlockAct { float *id = nullptr, *t1 = nullptr, *a1 = nullptr, *t2 = nullptr, *out = nullptr, *mi1 = nullptr, *mi2 = nullptr; const float* x = nullptr;
                  float *a1s = nullptr, *outs = nullptr; const float* xs = nullptr;
                  unsigned char* m2 = nullptr;   /* sign bytes of the block's final pre-activation (one per float4 of t2) */ };
struct EncPass {
  const float* img = nullptr;
  float *t0 = nullptr, *mi0 = nullptr, *a0 = nullptr, *a0s = nullptr, *feat = nullptr, *ml = nullptr, *z = nullptr, *kl = nullptr;
  std::vector<BlockAct> blk;
};
struct DecPass {
  const float* zin = nullptr;
  float *h = nullptr, *x0 = nullptr, *x0s = nullptr, *x0h = nullptr, *y = nullptr;
  double* ema = nullptr;           // EMA inputs (batch mean | unbiased var) of every BN of the pass, in the net's BN buffer layout
  const void* net = nullptr;       // the weight set the slot's activations were computed with (pass re-use check)
  std::vector<BlockAct> blk;
};

struct Bump {
  char* base = nullptr; size_t off = 0;
  template <class T> T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

}  // namespace

struct sivae_engine {
  sivae_config cfg;
  Net nets[3];
  int C_last = 0, hw_last = 0;     // conv_output_size = (C_last, hw_last, hw_last)
  long long feat = 0;              // C_last*hw_last^2
  bool tc = false;                 // tcgen05 backend in use (activations pre-rounded to tf32)
  bool fast = false;               // cdim-facing narrow CUDA-core kernels in use (false: generic exact SIMT everywhere)
  bool comp = false;               // compensated tensor-core mode (SIVAE_CONV_TC3X): 3 tf32 MMAs per product on split operands
  bool rnd = false;                // producers round stored activations / gradients to tf32 (plain tensor-core mode only)
  bool fsplit = false;             // default tensor-core mode: FORWARD convs consume split32 operands (bf16 hi + lo, three
                                   // kind::f16 MMAs per product: fp32-class products, ELBO / KL within 1e-4 of the reference);
                                   // dgrad / wgrad stay kind::tf32 on the tf32-rounded fp32 tensors, unless:
  bool bwd16 = false;              // (with fsplit) dgrad / wgrad of the residual blocks run kind::f16 on plain bf16 operands: the
                                   // BatchNorm backward writes bf16 gradients, the forward keeps bf16 copies of its activations
  bool bn_mask = true;             // residual BN+LeakyReLU forward stores sign bytes so its backward skips the identity re-read (SIVAE_BN_MASK=0: off)
  float* split[4] = {nullptr, nullptr, nullptr, nullptr};   // comp: hi / lo parts of the two conv operands, max activation size each
  // workspace
  void* ws = nullptr; size_t ws_bytes = 0, ws_need = 0;
  EncPass ep[3]; DecPass dp[4];
  float *real = nullptr, *noise = nullptr, *z_keep = nullptr;
  float* sb[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // backward scratch, max activation size each
  float *d_rec = nullptr, *d_rec_rec = nullptr, *d_rec_fake = nullptr, *d_fake = nullptr;
  float *dml = nullptr, *dz = nullptr, *dfeat = nullptr, *dfeat2 = nullptr, *coef = nullptr, *ckl_a = nullptr, *ckl_b = nullptr, *mse = nullptr;
  float *out_tmp = nullptr;
  float *featc = nullptr, *zc = nullptr;   // conditional models: [feat | cond] and [z | cond] rows, the fc inputs of :118-119, :163-165
  void* sk = nullptr; size_t sk_bytes = 0;   // split-K partial tensor of the few-tile tensor-core convs
  float *rs = nullptr;             // scratch of the row-separable image-facing convs (B*S*S*32 floats)
  void* red = nullptr; size_t red_bytes = 0;
  int cur_batch = 0;
  bool have_e_state = false;
  // data parallel: NCCL communicator the gradient all-reduces run on, enqueued on the step's own stream
  ncclComm_t comm = nullptr; int world = 1; bool own_comm = false;
  // overlap of the encoder-gradient all-reduce + Adam(encoder) with the two decoder passes that open the D half (they do not
  // depend on the encoder): a library-owned side stream forked from / joined back into the step's stream with events
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool dp_overlap = true;
  bool reuse_dec = false;          // D half re-uses the E half's fake / rec decoder passes (SIVAE_REUSE_DEC=1 or sivae_set_option)
  bool e_dec_valid = false;        // dp[0] / dp[1] hold D(noise) / D(z) of the current decoder weights
  int loss_type = SIVAE_LOSS_MSE_; // recon_loss_type of the step (sivae_set_recon_loss): mse | l1 | bce (:268-294)
  // the 'mean' reduction of the l1 / bce branches divides by B*D, the mse branch by B (:282-291)
  float mean_div() const { return loss_type == SIVAE_LOSS_MSE_ ? 1.f : 1.f / (float)((long long)cfg.cdim * cfg.image_size * cfg.image_size); }
  int* bad_flag() const { return loss_type == SIVAE_LOSS_BCE_ ? reinterpret_cast<int*>(coef + 3 * (long long)cfg.max_batch) : nullptr; }
};

// -------------------------------------------------------------------------------------------------------------
// model description
// -------------------------------------------------------------------------------------------------------------
static void add_tensor(Net& n, const std::string& name, int kind, long long numel, std::initializer_list<int> shape, long long* off_out) {
  sivae_tensor_info ti;
  memset(&ti, 0, sizeof(ti));
  snprintf(ti.name, sizeof(ti.name), "%s", name.c_str());
  ti.kind = kind;
  ti.offset = n.n_params;
  ti.numel = numel;
  ti.ndim = (int)shape.size();
  int i = 0;
  for (int s : shape) ti.shape[i++] = s;
  *off_out = n.n_params;
  // keep every tensor 16-byte aligned in the flat buffer (float4 kernels, TMA): round sizes up to 4 floats
  n.n_params += (numel + 3) / 4 * 4;
  n.tinfo.push_back(ti);
}
static void add_conv(Net& n, const std::string& name, Conv& c, int cin, int cout, int k, bool bias) {
  c.cin = cin; c.cout = cout; c.k = k;
  add_tensor(n, name + ".weight", SIVAE_T_CONV, (long long)cout * cin * k * k, {cout, cin, k, k}, &c.w_off);
  if (bias) add_tensor(n, name + ".bias", SIVAE_T_BIAS, cout, {cout}, &c.b_off);
  c.wd_off = n.derived_floats; n.derived_floats += (long long)cout * cin * k * k;
  c.wr_off = n.derived_floats; n.derived_floats += (long long)cout * cin * k * k;
  // narrow-side filter forms: wn = [tap][narrow][wide] (CUDA-core kernels) or the row-expanded [wide][5][32] form;
  // wg = the 16-row gather form [16][5][wide] of the row-separable tensor-core path (conv_tc.cu)
  const bool narrow = (cin <= 4 || cout <= 4) && k == 5;
  const long long wide = cin > cout ? cin : cout;
  c.wn_off = n.derived_floats; n.derived_floats += narrow ? wide * 160 : (long long)cout * cin * k * k;
  if (narrow) { c.wg_off = n.derived_floats; n.derived_floats += wide * 80; }
  c.wdh_off = n.derived_floats; n.derived_floats += ((long long)cout * cin * k * k + 7) / 8 * 4;
  if (n.comp) {
    c.wl_off = n.derived_floats; n.derived_floats += (long long)cout * cin * k * k;
    c.wdl_off = n.derived_floats; n.derived_floats += (long long)cout * cin * k * k;
  }
}
static void add_bn(Net& n, const std::string& name, Bn& b, int c) {
  b.c = c;
  add_tensor(n, name + ".weight", SIVAE_T_BN_WEIGHT, c, {c}, &b.g_off);
  add_tensor(n, name + ".bias", SIVAE_T_BN_BIAS, c, {c}, &b.b_off);
  sivae_bn_info bi;
  memset(&bi, 0, sizeof(bi));
  snprintf(bi.name, sizeof(bi.name), "%s", name.c_str());
  bi.channels = c;
  bi.bn_offset = n.bn_floats;
  bi.index = (int)n.binfo.size();
  b.rm_off = n.bn_floats;
  b.idx = bi.index;
  n.bn_floats += 2LL * c;
  n.binfo.push_back(bi);
}
static void add_block(Net& n, const std::string& p, Block& b, int inc, int outc, int size, int mode) {
  b.inc = inc; b.outc = outc; b.size = size; b.mode = mode; b.expand = inc != outc;
  if (b.expand) add_conv(n, p + ".conv_expand", b.ce, inc, outc, 1, false);
  add_conv(n, p + ".conv1", b.c1, inc, outc, 3, false);
  add_bn(n, p + ".bn1", b.bn1, outc);
  add_conv(n, p + ".conv2", b.c2, outc, outc, 3, false);
  add_bn(n, p + ".bn2", b.bn2, outc);
}
static void add_lin(Net& n, const std::string& name, Lin& l, int fin, int fout) {
  l.fin = fin; l.fout = fout;
  add_tensor(n, name + ".weight", SIVAE_T_LINEAR, (long long)fin * fout, {fout, fin}, &l.w_off);
  add_tensor(n, name + ".bias", SIVAE_T_BIAS, fout, {fout}, &l.b_off);
}

// Encoder.__init__ (:79-109) -- same module order, names `res_in_{sz}`
static void build_encoder(sivae_engine* e, Net& n) {
  const sivae_config& c = e->cfg;
  n.enc = true; n.present = true;
  int cc = c.channels[0];
  add_conv(n, "main.0", n.stem, c.cdim, cc, 5, false);
  add_bn(n, "main.1", n.stem_bn, cc);
  int sz = c.image_size / 2;
  for (int i = 1; i < c.n_channels; ++i) {
    Block b;
    add_block(n, "main.res_in_" + std::to_string(sz), b, cc, c.channels[i], sz, RS_POOL);
    n.blocks.push_back(b);
    cc = c.channels[i]; sz /= 2;
  }
  Block b;
  add_block(n, "main.res_in_" + std::to_string(sz), b, cc, cc, sz, RS_NONE);
  n.blocks.push_back(b);
  e->C_last = cc; e->hw_last = sz; e->feat = (long long)cc * sz * sz;
  add_lin(n, "fc", n.fc, (int)e->feat + c.cond_dim, 2 * c.zdim);     // conditional: Linear(num_fc_features + cond_dim, 2z), :106-107
}
// Decoder.__init__ (:126-159) -- names count from 4, real size starts at conv_output_size
static void build_decoder(sivae_engine* e, Net& n) {
  const sivae_config& c = e->cfg;
  n.enc = false; n.present = true;
  add_lin(n, "fc.0", n.fc, c.zdim + c.cond_dim, (int)e->feat);       // conditional: Linear(zdim + cond_dim, ...), :139-143
  int cc = c.channels[c.n_channels - 1];
  int name_sz = 4, real = e->hw_last;
  for (int i = c.n_channels - 1; i >= 0; --i) {
    Block b;
    add_block(n, "main.res_in_" + std::to_string(name_sz), b, cc, c.channels[i], real, RS_UP);
    n.blocks.push_back(b);
    cc = c.channels[i]; name_sz *= 2; real *= 2;
  }
  Block b;
  add_block(n, "main.res_in_" + std::to_string(name_sz), b, cc, cc, real, RS_NONE);
  n.blocks.push_back(b);
  add_conv(n, "main.predict", n.predict, cc, c.cdim, 5, true);
}

static long long max_act_elems(const sivae_engine* e) {
  const sivae_config& c = e->cfg;
  long long B = c.max_batch, S = c.image_size;
  long long m = B * S * S * c.channels[0];
  for (int ni = 0; ni < 2; ++ni)
    for (const Block& b : e->nets[ni].blocks) {
      long long v = B * b.size * b.size * (long long)(b.outc > b.inc ? b.outc : b.inc);
      if (b.mode == RS_UP) v = B * 4LL * b.size * b.size * b.outc;
      if (v > m) m = v;
    }
  return m;
}

static size_t reduce_scratch_bytes(const sivae_engine* e) {
  const sivae_config& c = e->cfg;
  long long B = c.max_batch, S = c.image_size;
  size_t m = bn_scratch_bytes(B * S * S, c.channels[0]);
  size_t v = mse3_scratch_bytes((int)B, (long long)c.cdim * S * S);
  if (v > m) m = v;
  v = colsum_scratch_bytes(c.cdim);
  if (v > m) m = v;
  v = linear_dgrad_scratch_bytes((int)B, (int)e->feat, 2 * c.zdim);
  if (v > m) m = v;
  v = linear_dgrad_scratch_bytes((int)B, c.zdim, (int)e->feat);
  if (v > m) m = v;
  auto upd_conv = [&](const Conv& cv, int size) {
    if (cv.k == 0) return;
    ConvShape s{(int)B, size, size, cv.cin, cv.cout, cv.k};
    size_t a = conv_wgrad_simt_scratch_bytes(s);
    if (a > m) m = a;
    if (conv_tc_supported_wgrad(s)) { size_t t = conv_wgrad_tc_scratch_bytes(s); if (t > m) m = t; }
    if (conv_tc_supported_wgrad16(s)) { size_t t = conv_wgrad_tc16_scratch_bytes(s); if (t > m) m = t; }
    { int parts = conv_tc_stats_parts(s); if (parts > 0) { size_t t = bn_parts_scratch_bytes(parts, cv.cout); if (t > m) m = t; } }
    if ((cv.cin <= 3 || cv.cout <= 3) && cv.k == 5) {
      const int wide = cv.cin > cv.cout ? cv.cin : cv.cout, c = cv.cin > cv.cout ? cv.cout : cv.cin;
      if (conv_rowsep_wgrad_supported(size, size, c, wide, cv.k)) { size_t t = conv_rowsep_wgrad_scratch_bytes((int)B, size, size, wide); if (t > m) m = t; }
    }
    if (cv.cin <= 4 && conv_narrow_corr_supported(cv.cout, cv.cin, cv.k)) { size_t t = conv_narrow_corr_scratch_bytes(cv.cout, cv.k); if (t > m) m = t; }
    if (cv.cout <= 4 && conv_narrow_corr_supported(cv.cin, cv.cout, cv.k)) { size_t t = conv_narrow_corr_scratch_bytes(cv.cin, cv.k); if (t > m) m = t; }
  };
  for (int ni = 0; ni < 2; ++ni) {
    const Net& n = e->nets[ni];
    for (const Block& b : n.blocks) {
      size_t a = bn_scratch_bytes(B * b.size * b.size, b.outc);
      if (a > m) m = a;
      upd_conv(b.ce, b.size); upd_conv(b.c1, b.size); upd_conv(b.c2, b.size);
    }
  }
  upd_conv(e->nets[0].stem, (int)S);
  upd_conv(e->nets[1].predict, (int)S);
  return m;
}

// carve the workspace; with base == nullptr only computes the size
static size_t carve(sivae_engine* e, char* base) {
  const sivae_config& c = e->cfg;
  Bump bp; bp.base = base;
  const long long B = c.max_batch, S = c.image_size, z = c.zdim;
  for (int ni = 0; ni < 3; ++ni)
    if (e->nets[ni].present) {
      e->nets[ni].derived = bp.take<float>(e->nets[ni].derived_floats);
      e->nets[ni].step_dev = bp.take<long long>(2);
      e->nets[ni].coef_dev = bp.take<float>(4);
    }
  e->real = bp.take<float>(B * S * S * c.cdim);
  e->noise = bp.take<float>(B * z);
  e->z_keep = bp.take<float>(B * z);
  const Net& en = e->nets[0];
  const Net& dn = e->nets[1];
  for (int s = 0; s < 3; ++s) {
    EncPass& p = e->ep[s];
    p.t0 = bp.take<float>(B * S * S * c.channels[0]);
    p.mi0 = bp.take<float>(2 * c.channels[0]);
    p.a0 = bp.take<float>(B * (S / 2) * (S / 2) * c.channels[0]);
    p.a0s = e->fsplit ? bp.take<float>(B * (S / 2) * (S / 2) * c.channels[0]) : nullptr;
    p.blk.assign(en.blocks.size(), BlockAct());
    for (size_t i = 0; i < en.blocks.size(); ++i) {
      const Block& b = en.blocks[i];
      long long full = B * b.size * b.size * b.outc;
      long long os = b.mode == RS_POOL ? b.size / 2 : b.size;
      BlockAct& a = p.blk[i];
      if (b.expand) a.id = bp.take<float>(full);
      a.t1 = bp.take<float>(full); a.a1 = bp.take<float>(full); a.t2 = bp.take<float>(full);
      a.out = bp.take<float>(B * os * os * b.outc);
      if (e->fsplit) { a.a1s = bp.take<float>(full); a.outs = (i + 1 < en.blocks.size()) ? bp.take<float>(B * os * os * b.outc) : nullptr; }
      a.mi1 = bp.take<float>(2 * b.outc); a.mi2 = bp.take<float>(2 * b.outc);
      a.m2 = e->bn_mask ? bp.take<unsigned char>(full / 4) : nullptr;
    }
    p.feat = bp.take<float>(B * e->feat);
    p.ml = bp.take<float>(B * 2 * z);
    p.z = bp.take<float>(B * z);
    p.kl = bp.take<float>(B);
  }
  for (int s = 0; s < 4; ++s) {
    DecPass& p = e->dp[s];
    p.h = bp.take<float>(B * e->feat);
    p.x0 = bp.take<float>(B * e->feat);
    p.x0s = e->fsplit ? bp.take<float>(B * e->feat) : nullptr;
    p.x0h = e->bwd16 ? bp.take<float>((B * e->feat + 1) / 2) : nullptr;
    p.blk.assign(dn.blocks.size(), BlockAct());
    for (size_t i = 0; i < dn.blocks.size(); ++i) {
      const Block& b = dn.blocks[i];
      long long full = B * b.size * b.size * b.outc;
      long long os = b.mode == RS_UP ? b.size * 2 : b.size;
      BlockAct& a = p.blk[i];
      if (b.expand) a.id = bp.take<float>(full);
      a.t1 = bp.take<float>(full); a.a1 = bp.take<float>(full); a.t2 = bp.take<float>(full);
      a.out = bp.take<float>(B * os * os * b.outc);
      if (e->fsplit) { a.a1s = bp.take<float>(full); a.outs = bp.take<float>(B * os * os * b.outc); }
      a.mi1 = bp.take<float>(2 * b.outc); a.mi2 = bp.take<float>(2 * b.outc);
      a.m2 = e->bn_mask ? bp.take<unsigned char>(full / 4) : nullptr;
    }
    p.y = bp.take<float>(B * S * S * c.cdim);
    p.ema = bp.take<double>(dn.bn_floats);
  }
  long long ma = max_act_elems(e);
  for (int i = 0; i < 5; ++i) e->sb[i] = bp.take<float>(ma);
  if (e->comp) for (int i = 0; i < 4; ++i) e->split[i] = bp.take<float>(ma);
  long long img = B * S * S * c.cdim;
  e->d_rec = bp.take<float>(img); e->d_rec_rec = bp.take<float>(img); e->d_rec_fake = bp.take<float>(img); e->d_fake = bp.take<float>(img);
  e->out_tmp = bp.take<float>(img);
  if (c.cond_dim > 0) { e->featc = bp.take<float>(B * (e->feat + c.cond_dim)); e->zc = bp.take<float>(B * (z + c.cond_dim)); }
  e->rs = bp.take<float>(B * S * S * 32);        // row-expanded image / 16-column partial image of the row-separable convs
  e->dml = bp.take<float>(B * 2 * z); e->dz = bp.take<float>(B * z);
  e->dfeat = bp.take<float>(B * e->feat); e->dfeat2 = bp.take<float>(B * e->feat);
  e->coef = bp.take<float>(4 * B); e->ckl_a = bp.take<float>(B); e->ckl_b = bp.take<float>(B); e->mse = bp.take<float>(3 * B);
  {
    size_t m = 0;
    auto upd = [&](const Conv& cv, int size) {
      if (cv.k == 0 || !e->tc) return;
      size_t a = conv_tc_splitk_scratch_bytes(ConvShape{(int)B, size, size, cv.cin, cv.cout, cv.k});
      size_t b2 = conv_tc_splitk_scratch_bytes(ConvShape{(int)B, size, size, cv.cout, cv.cin, cv.k});
      if (a > m) m = a;
      if (b2 > m) m = b2;
    };
    for (int ni = 0; ni < 2; ++ni)
      for (const Block& b : e->nets[ni].blocks) { upd(b.ce, b.size); upd(b.c1, b.size); upd(b.c2, b.size); }
    e->sk_bytes = m;
    e->sk = m ? bp.take<char>(m) : nullptr;
  }
  e->red_bytes = reduce_scratch_bytes(e);
  e->red = bp.take<char>(e->red_bytes);
  return bp.off + 256;
}

// -------------------------------------------------------------------------------------------------------------
// conv dispatch
// -------------------------------------------------------------------------------------------------------------

// ---- optional per-kernel-class timing (CUDA events on the launching stream), used by bench.py's roofline ---------
// PC_TC_FWD3 = forward convs on split32 operands (three kind::f16 MMAs per product); PC_TC_FWD = kind::tf32 forward / dgrad
// PC_TC_DG16 / PC_TC_WG16 = dgrad / wgrad of the residual blocks on plain bf16 operands (kind::f16, one MMA per product)
enum { PC_TC_FWD = 0, PC_TC_WGRAD = 1, PC_SIMT_FWD = 2, PC_SIMT_WGRAD = 3, PC_LOSS = 4, PC_TC_FWD3 = 5, PC_TC_DG16 = 6, PC_TC_WG16 = 7,
       PC_COUNT = 8, PC_BN_FWD = 8, PC_BN_BWD = 9 };      // classes >= PC_COUNT appear in sivae_profile_dump only (bytes in the flops slot)
struct ProfRec { cudaEvent_t a, b; int cls; double flops; ConvShape shape; };
struct Prof {
  bool on = false;
  std::vector<ProfRec> recs;
  size_t used = 0;
};
static Prof g_prof;
struct ProfScope {
  cudaStream_t st; ProfRec* r = nullptr;
  ProfScope(int cls, const ConvShape& s, cudaStream_t st_) : st(st_) {
    if (!g_prof.on) return;
    if (g_prof.used == g_prof.recs.size()) {
      ProfRec n; cudaEventCreate(&n.a); cudaEventCreate(&n.b); n.cls = 0; n.flops = 0;
      g_prof.recs.push_back(n);
    }
    r = &g_prof.recs[g_prof.used++];
    r->cls = cls;
    r->flops = 2.0 * (double)s.pixels() * (double)s.Cout * (double)s.ktot();
    r->shape = s;
    cudaEventRecord(r->a, st);
  }
  ~ProfScope() { if (r) cudaEventRecord(r->b, st); }
};
// fused loss pass: `flops` carries the algorithmic BYTES (5 images read once)
struct ProfLoss {
  ProfScope ps;
  ProfLoss(int B, long long per, cudaStream_t st) : ps(PC_LOSS, ConvShape{B, 1, 1, 1, 1, 1}, st) {
    if (ps.r) ps.r->flops = 5.0 * (double)B * (double)per * 4.0;
  }
};
// BatchNorm + activation (+ residual, + resample) passes: `flops` carries the algorithmic bytes.
// passes = tensors of B*s*s*C floats touched (a pooled tensor counts 1/4, an upsampled one 4)
struct ProfElem {
  ProfScope ps;
  ProfElem(int cls, int B, int s, int C, int mode, double passes, cudaStream_t st) : ps(cls, ConvShape{B, s, s, C, C, mode}, st) {
    if (ps.r) ps.r->flops = passes * (double)B * s * s * C * 4.0;
  }
};
static double resampled(int mode) { return mode == RS_POOL ? 0.25 : (mode == RS_UP ? 4.0 : 1.0); }
// which implementation serves a convolution (exact = SIMT-only engine; narrow = cdim-facing CUDA-core kernels)
static bool fwd_on_tc(const sivae_engine* e, const ConvShape& s) { return e->tc && conv_tc_supported_fwd(s); }
static bool fwd_on_rowsep_in(const sivae_engine* e, const ConvShape& s) { return e->tc && !e->comp && e->fast && e->rs && conv_rowsep_in_supported(s); }
static bool fwd_on_rowsep_out(const sivae_engine* e, const ConvShape& s) { return e->tc && !e->comp && e->fast && e->rs && conv_rowsep_out_supported(s); }
static bool fwd_on_narrow(const sivae_engine* e, const ConvShape& s) { return e->fast && !fwd_on_rowsep_in(e, s) && conv_narrow_in_supported(s); }

static int refresh_derived(sivae_engine* e, Net& n, cudaStream_t st) {
  if (!n.dirty) return 0;
  auto one = [&](const Conv& c, int size = 8) {
    if (c.k == 0) return;
    // dgrad filters (a dgrad is a forward conv over dy): rounded to tf32 only when the tensor core consumes them
    ConvShape sd{1, size, size, c.cout, c.cin, c.k};
    const long long wn = (long long)c.cout * c.cin * c.k * c.k;
    const bool d_tc = fwd_on_tc(e, sd) && !fwd_on_narrow(e, sd);
    // (bf16 backward: the residual blocks' dgrads read only the bf16 form below; the tf32 form is needed by the image-facing convs)
    if (!(e->bwd16 && c.k != 5))
      launch_pack_dgrad_filter(n.params + c.w_off, n.derived + c.wd_off, c.cout, c.cin, c.k, d_tc && !e->comp, st);
    if (d_tc && e->comp) launch_split_tf32(n.derived + c.wd_off, n.derived + c.wd_off, n.derived + c.wdl_off, wn, st);
    if (e->bwd16 && c.k != 5) launch_pack_dgrad_filter_bf16(n.params + c.w_off, n.derived + c.wdh_off, c.cout, c.cin, c.k, st);
    ConvShape sf{1, size, size, c.cin, c.cout, c.k};
    if (fwd_on_tc(e, sf) && !fwd_on_narrow(e, sf)) {
      if (e->comp) launch_split_tf32(n.params + c.w_off, n.derived + c.wr_off, n.derived + c.wl_off, wn, st);
      else if (e->fsplit) launch_split32(n.params + c.w_off, n.derived + c.wr_off, wn, st);     // forward filters: split32
      else launch_round_tf32(n.params + c.w_off, n.derived + c.wr_off, wn, st);
    }
    // narrow (cdim-facing) kernels read the filter transposed to [tap][narrow][wide]
    if (fwd_on_narrow(e, sf)) launch_narrow_transpose(n.params + c.w_off, n.derived + c.wn_off, c.cout, c.k, c.cin, st);
    else if (fwd_on_narrow(e, sd)) launch_narrow_transpose(n.derived + c.wd_off, n.derived + c.wn_off, c.cin, c.k, c.cout, st);
    // row-separable tensor-core forms of the image-facing 5x5 convs
    // (split-forward mode: the forms a FORWARD conv reads are converted to split32 in place; dgrad forms stay tf32)
    if (fwd_on_rowsep_in(e, sf)) {
      launch_rowsep_filter_expand(n.params + c.w_off, n.derived + c.wn_off, c.cout, c.cin, st, !e->fsplit);
      if (e->fsplit) launch_split32(n.derived + c.wn_off, n.derived + c.wn_off, (long long)c.cout * 160, st);
    } else if (fwd_on_rowsep_in(e, sd)) launch_rowsep_filter_expand(n.derived + c.wd_off, n.derived + c.wn_off, c.cin, c.cout, st);
    if (c.wg_off >= 0) {
      if (fwd_on_rowsep_out(e, sf)) {
        launch_rowsep_filter_gather(n.params + c.w_off, n.derived + c.wg_off, c.cout, c.cin, st, !e->fsplit);
        if (e->fsplit) launch_split32(n.derived + c.wg_off, n.derived + c.wg_off, (long long)80 * c.cin, st);
      } else if (fwd_on_rowsep_out(e, sd)) launch_rowsep_filter_gather(n.derived + c.wd_off, n.derived + c.wg_off, c.cin, c.cout, st);
    }
  };
  if (n.enc) one(n.stem, e->cfg.image_size); else one(n.predict, e->cfg.image_size);
  for (const Block& b : n.blocks) { one(b.ce); one(b.c1); one(b.c2); }
  n.dirty = false;
  CHECK_CUDA_RET();
  return 0;
}
// y = conv(x, filt) (+bias) (+addend).  w_master: fp32 filter [Cout][k][k][Cin]; w_tc: its tf32-rounded copy
// fmt = FMT_SPLIT_ (split-forward mode, forward convs only): x (except the fp32 image of a narrow-input conv), w_tc, w_narrow
// and w_gather are split32 tensors; only tensor-core kernels can serve such a call
static int conv_any(sivae_engine* e, const ConvShape& s, const float* x, const float* w_master, const float* w_tc,
                    const float* bias, const float* addend, float* y, cudaStream_t st, float* stats = nullptr,
                    const float* w_narrow = nullptr, const float* w_gather = nullptr, const float* w_lo = nullptr,
                    int fmt = FMT_TF32_) {
  const int pc_fwd = fmt == FMT_SPLIT_ ? PC_TC_FWD3 : (fmt == FMT_BF16_ ? PC_TC_DG16 : PC_TC_FWD);
  if (fwd_on_rowsep_in(e, s) && w_narrow) {
    ProfScope ps(pc_fwd, s, st);
    int r = launch_conv_rowsep_in(x, w_narrow, bias, addend, y, s, e->rs, stats, st, fmt);
    if (r) return fail(r, "row-separable tcgen05 conv launch failed");
  } else if (fwd_on_rowsep_out(e, s) && w_gather) {
    ProfScope ps(pc_fwd, s, st);
    int r = launch_conv_rowsep_out(x, w_gather, bias, addend, y, s, e->rs, st, fmt);
    if (r) return fail(r, "row-separable tcgen05 conv launch failed");
  } else if (fmt != FMT_TF32_) {
    if (!e->tc || !conv_tc_supported_fwd(s, fmt)) return fail(-7, "split32 / bf16 conv: shape not served by a tensor-core kernel");
    ProfScope ps(pc_fwd, s, st);
    int r = launch_conv_fwd_tc(x, w_tc, bias, addend, y, s, st, stats, e->sk, e->sk_bytes, fmt);
    if (r) return fail(r, "tcgen05 conv launch failed");
  } else if (fwd_on_narrow(e, s) && w_narrow) {
    ProfScope ps(PC_SIMT_FWD, s, st);
    launch_conv_narrow_in_fwd(x, w_narrow, bias, addend, y, s, st);
  } else if (fwd_on_tc(e, s) && e->comp) {
    // 3xTF32: x = xh + xl, w = wh + wl (every part an exact tf32 operand); y = xl*wh + xh*wl + xh*wh (+bias +addend), the
    // three tensor-core convs chained through the in-place addend path, small terms first.  Dropped: xl*wl (2^-22).
    if (!w_lo || !e->split[0]) return fail(-4, "compensated conv needs the split filter and the split scratch");
    ProfScope ps(PC_TC_FWD, s, st);
    float *xh = e->split[0], *xl = e->split[1];
    launch_split_tf32(x, xh, xl, s.pixels() * s.Cin, st);
    int r = launch_conv_fwd_tc(xl, w_tc, bias, addend, y, s, st, nullptr, e->sk, e->sk_bytes);
    if (!r) r = launch_conv_fwd_tc(xh, w_lo, nullptr, y, y, s, st, nullptr, e->sk, e->sk_bytes);
    if (!r) r = launch_conv_fwd_tc(xh, w_tc, nullptr, y, y, s, st, nullptr, e->sk, e->sk_bytes);
    if (r) return fail(r, "tcgen05 conv launch failed");
  } else if (fwd_on_tc(e, s)) {
    ProfScope ps(PC_TC_FWD, s, st);
    int r = launch_conv_fwd_tc(x, w_tc, bias, addend, y, s, st, stats, e->sk, e->sk_bytes);
    if (r) return fail(r, "tcgen05 conv launch failed");
  } else {
    ProfScope ps(PC_SIMT_FWD, s, st);
    launch_conv_fwd_simt(x, w_master, bias, addend, y, s, st);
  }
  return 0;
}
static int conv_fwd(sivae_engine* e, Net& n, const Conv& c, const float* x, float* y, const float* addend, int B, int size, cudaStream_t st) {
  ConvShape s{B, size, size, c.cin, c.cout, c.k};
  const float* bias = c.b_off >= 0 ? n.params + c.b_off : nullptr;
  return conv_any(e, s, x, n.params + c.w_off, n.derived + c.wr_off, bias, addend, y, st, nullptr, n.derived + c.wn_off,
                  c.wg_off >= 0 ? n.derived + c.wg_off : nullptr, c.wl_off >= 0 ? n.derived + c.wl_off : nullptr,
                  e->fsplit ? FMT_SPLIT_ : FMT_TF32_);
}
// t = conv(x, W) followed by the BatchNorm batch statistics of t (train) or the running statistics (eval).  On the tensor
// core path the statistics come out of the conv epilogue (no second pass over t).
static int conv_bn_stats(sivae_engine* e, Net& n, const Conv& c, const Bn& bn, const float* x, float* t, float* mi, int B, int size,
                         bool train, cudaStream_t st, double* ema = nullptr) {
  ConvShape s{B, size, size, c.cin, c.cout, c.k};
  const long long rows = (long long)B * size * size;
  int parts = (train && e->tc && !e->comp && (fwd_on_rowsep_in(e, s) || (!fwd_on_narrow(e, s) && fwd_on_tc(e, s)))) ? conv_tc_stats_parts(s) : 0;
  if (parts > 0 && bn_parts_scratch_bytes(parts, c.cout) > e->red_bytes) parts = 0;
  float* sp = parts > 0 ? (float*)e->red : nullptr;
  TRY(conv_any(e, s, x, n.params + c.w_off, n.derived + c.wr_off, nullptr, nullptr, t, st, sp, n.derived + c.wn_off,
               c.wg_off >= 0 ? n.derived + c.wg_off : nullptr, c.wl_off >= 0 ? n.derived + c.wl_off : nullptr,
               e->fsplit ? FMT_SPLIT_ : FMT_TF32_));
  if (parts > 0)
    launch_bn_stats_from_parts(sp, parts, rows, bn.c, mi, n.bn + bn.rm_off, n.bn + bn.rm_off + bn.c, n.nbt + bn.idx, st,
                               ema ? ema + bn.rm_off : nullptr);
  else if (train)
    launch_bn_stats(t, rows, bn.c, mi, n.bn + bn.rm_off, n.bn + bn.rm_off + bn.c, n.nbt + bn.idx, e->red, e->red_bytes, st,
                    ema ? ema + bn.rm_off : nullptr);
  else
    launch_bn_eval_stats(n.bn + bn.rm_off, n.bn + bn.rm_off + bn.c, bn.c, mi, st);
  return 0;
}
// dx = conv_transpose(dy, W) (+addend): a forward conv over dy with the packed dgrad filters
// dy16: dy is a plain bf16 tensor (written by the BatchNorm backward with out16) -> kind::f16 dgrad with the bf16 filter
static int conv_dgrad(sivae_engine* e, Net& n, const Conv& c, const float* dy, float* dx, const float* addend, int B, int size, cudaStream_t st,
                      bool dy16 = false) {
  ConvShape s{B, size, size, c.cout, c.cin, c.k};
  if (dy16)
    return conv_any(e, s, dy, nullptr, n.derived + c.wdh_off, nullptr, addend, dx, st, nullptr, nullptr, nullptr, nullptr, FMT_BF16_);
  return conv_any(e, s, dy, n.derived + c.wd_off, n.derived + c.wd_off, nullptr, addend, dx, st, nullptr, n.derived + c.wn_off,
                  c.wg_off >= 0 ? n.derived + c.wg_off : nullptr, c.wdl_off >= 0 ? n.derived + c.wdl_off : nullptr);
}
// op16: x and dy are plain bf16 tensors
static int conv_wgrad(sivae_engine* e, Net& n, const Conv& c, const float* x, const float* dy, int B, int size, cudaStream_t st,
                      bool op16 = false) {
  ConvShape s{B, size, size, c.cin, c.cout, c.k};
  float* dw = n.grads + c.w_off;
  if (op16) {
    ProfScope ps(PC_TC_WG16, s, st);
    int r = launch_conv_wgrad_tc16(x, dy, dw, s, true, e->red, e->red_bytes, st);
    if (r) return fail(r, "tcgen05 bf16 conv wgrad launch failed");
    return 0;
  }
  if (e->tc && !e->comp && e->fast && e->rs && c.cin <= 3 && conv_rowsep_wgrad_supported(size, size, c.cin, c.cout, c.k)) {          // stem
    ProfScope ps(PC_TC_WGRAD, s, st);
    int r = launch_conv_rowsep_wgrad(x, dy, dw, B, size, size, c.cin, c.cout, 1, true, e->rs, e->red, e->red_bytes, st);
    if (r) return fail(r, "row-separable tcgen05 wgrad launch failed");
  } else if (e->tc && !e->comp && e->fast && e->rs && c.cout <= 3 && conv_rowsep_wgrad_supported(size, size, c.cout, c.cin, c.k)) {  // predict
    ProfScope ps(PC_TC_WGRAD, s, st);
    int r = launch_conv_rowsep_wgrad(dy, x, dw, B, size, size, c.cout, c.cin, 0, true, e->rs, e->red, e->red_bytes, st);
    if (r) return fail(r, "row-separable tcgen05 wgrad launch failed");
  } else if (e->fast && c.cin <= 4 && conv_narrow_corr_supported(c.cout, c.cin, c.k)) {            // stem
    ProfScope ps(PC_SIMT_WGRAD, s, st);
    launch_conv_narrow_corr(x, dy, dw, B, size, size, c.cin, c.cout, c.k, 1, true, e->red, e->red_bytes, st);
  } else if (e->fast && c.cout <= 4 && conv_narrow_corr_supported(c.cin, c.cout, c.k)) {     // predict
    ProfScope ps(PC_SIMT_WGRAD, s, st);
    launch_conv_narrow_corr(dy, x, dw, B, size, size, c.cout, c.cin, c.k, 0, true, e->red, e->red_bytes, st);
  } else if (e->tc && e->comp && conv_tc_supported_wgrad(s)) {
    // 3xTF32: dw += xl (x) dyh + xh (x) dyl + xh (x) dyh
    if (!e->split[0]) return fail(-4, "compensated wgrad needs the split scratch");
    ProfScope ps(PC_TC_WGRAD, s, st);
    float *xh = e->split[0], *xl = e->split[1], *dh = e->split[2], *dl = e->split[3];
    launch_split_tf32(x, xh, xl, s.pixels() * s.Cin, st);
    launch_split_tf32(dy, dh, dl, s.pixels() * s.Cout, st);
    int r = launch_conv_wgrad_tc(xl, dh, dw, s, true, e->red, e->red_bytes, st);
    if (!r) r = launch_conv_wgrad_tc(xh, dl, dw, s, true, e->red, e->red_bytes, st);
    if (!r) r = launch_conv_wgrad_tc(xh, dh, dw, s, true, e->red, e->red_bytes, st);
    if (r) return fail(r, "tcgen05 conv wgrad launch failed");
  } else if (e->tc && conv_tc_supported_wgrad(s)) {
    ProfScope ps(PC_TC_WGRAD, s, st);
    int r = launch_conv_wgrad_tc(x, dy, dw, s, true, e->red, e->red_bytes, st);
    if (r) return fail(r, "tcgen05 conv wgrad launch failed");
  } else {
    ProfScope ps(PC_SIMT_WGRAD, s, st);
    launch_conv_wgrad_simt(x, dy, dw, s, true, e->red, e->red_bytes, st);
  }
  return 0;
}

// -------------------------------------------------------------------------------------------------------------
// forward passes
// -------------------------------------------------------------------------------------------------------------

// ResidualBlock.forward (:65-75) + the AvgPool2d / Upsample that follows it in `main` (:98,155)
// x: the fp32 block input (wgrad operand; unused by the forward in split-forward mode), xs: its split32 copy (split-forward
// mode).  keep: this pass will run wgrad, so the fp32 tf32-rounded copies of a1 / out are written next to the split32 ones.
static int block_forward(sivae_engine* e, Net& n, const Block& b, BlockAct& a, const float* x, const float* xs, int B, bool train,
                         bool keep, bool last, cudaStream_t st, double* ema = nullptr) {
  a.x = x; a.xs = xs;
  const int s = b.size;
  const bool sp = e->fsplit;
  const float* xin = sp ? xs : x;            // what the forward convs read
  const float* idn = xin;
  if (b.expand) { TRY(conv_fwd(e, n, b.ce, xin, a.id, nullptr, B, s, st)); idn = a.id; }
  TRY(conv_bn_stats(e, n, b.c1, b.bn1, xin, a.t1, a.mi1, B, s, train, st, ema));
  // wgrad operand copy of a1 (keep): fp32 tf32-rounded, or plain bf16 in the same buffer (bwd16)
  const bool h16 = e->bwd16;
  { ProfElem pe(PC_BN_FWD, B, s, b.outc, RS_NONE, 2.0 + ((sp && keep) ? (h16 ? 0.5 : 1.0) : 0.0), st);
    launch_bn_act_fwd(a.t1, nullptr, a.mi1, n.params + b.bn1.g_off, n.params + b.bn1.b_off, (sp && (!keep || h16)) ? nullptr : a.a1, B, s, s,
                      b.outc, RS_NONE, e->rnd, st, nullptr, sp ? a.a1s : nullptr, false, (sp && keep && h16) ? a.a1 : nullptr); }
  TRY(conv_bn_stats(e, n, b.c2, b.bn2, sp ? a.a1s : a.a1, a.t2, a.mi2, B, s, train, st, ema));
  // copies of `out` next to the split32 one: fp32 for the fc layer (last encoder block: no split32 twin) and for the
  // row-separable tf32 wgrad of `predict` (last decoder block, keep); otherwise (keep) the next block's wgrad operand:
  // fp32 tf32-rounded, or plain bf16 in the same buffer (bwd16)
  const bool f32_out = !sp || !a.outs || (keep && (!h16 || last));
  const bool h_out = sp && a.outs && keep && h16 && !last;
  { ProfElem pe(PC_BN_FWD, B, s, b.outc, 4 + b.mode, 2.0 + resampled(b.mode) * ((sp && a.outs) ? (f32_out ? 2.0 : (h_out ? 1.5 : 1.0)) : 1.0), st);
    launch_bn_act_fwd(a.t2, idn, a.mi2, n.params + b.bn2.g_off, n.params + b.bn2.b_off, f32_out ? a.out : nullptr, B, s, s, b.outc,
                      b.mode, (sp && !a.outs) ? false : e->rnd, st, a.m2, sp ? a.outs : nullptr, sp && !b.expand, h_out ? a.out : nullptr); }
  return 0;
}

// Encoder.forward (:116-122): img is NHWC [B,S,S,cdim]; result p.ml = [B,2z] (mu | logvar)
// keep: the pass will be followed by a backward WITH parameter gradients (see block_forward)
// [a | b] rows: out[B][wa + wb] (torch.cat(dim=1) of :119, :165)
static void concat_rows(const float* a, int wa, const float* b, int wb, float* out, int B, cudaStream_t st) {
  const size_t pitch = sizeof(float) * (size_t)(wa + wb);
  cudaMemcpy2DAsync(out, pitch, a, sizeof(float) * wa, sizeof(float) * wa, B, cudaMemcpyDeviceToDevice, st);
  cudaMemcpy2DAsync(out + wa, pitch, b, sizeof(float) * wb, sizeof(float) * wb, B, cudaMemcpyDeviceToDevice, st);
}
static const char* kCondMissing =
    "conditional model called without a condition: its fc layer takes [features | cond_dim] rows (the reference raises the "
    "matmul shape error of F.linear here, :118-120 / :163-166; its training step never passes one, :559-561)";
// cond (conditional models only): [B, cond_dim] rows concatenated to the fc input (:118-119)
static int enc_forward(sivae_engine* e, Net& n, EncPass& p, const float* img, int B, bool train, bool keep, cudaStream_t st,
                       const float* cond = nullptr) {
  const sivae_config& c = e->cfg;
  const int S = c.image_size;
  const bool sp = e->fsplit;
  p.img = img;
  TRY(conv_bn_stats(e, n, n.stem, n.stem_bn, img, p.t0, p.mi0, B, S, train, st));
  { ProfElem pe(PC_BN_FWD, B, S, n.stem.cout, RS_POOL, 1.25 + ((sp && keep) ? (e->bwd16 ? 0.125 : 0.25) : 0.0), st);
    launch_bn_act_fwd(p.t0, nullptr, p.mi0, n.params + n.stem_bn.g_off, n.params + n.stem_bn.b_off, (sp && (!keep || e->bwd16)) ? nullptr : p.a0,
                      B, S, S, n.stem.cout, RS_POOL, e->rnd, st, nullptr, sp ? p.a0s : nullptr, false, (sp && keep && e->bwd16) ? p.a0 : nullptr); }
  const float* x = p.a0;
  const float* xs = p.a0s;
  for (size_t i = 0; i < n.blocks.size(); ++i) {
    TRY(block_forward(e, n, n.blocks[i], p.blk[i], x, xs, B, train, keep, i + 1 == n.blocks.size(), st));
    x = p.blk[i].out; xs = p.blk[i].outs;
  }
  // .view(B, -1) of the NCHW tensor (:117)
  launch_nhwc_to_nchw(x, p.feat, B, e->C_last, e->hw_last, e->hw_last, st);
  const float* fc_in = p.feat;
  if (c.cond_dim > 0) {
    if (!cond) return fail(-7, kCondMissing);
    concat_rows(p.feat, (int)e->feat, cond, c.cond_dim, e->featc, B, st);
    fc_in = e->featc;
  }
  launch_linear_fwd(fc_in, n.params + n.fc.w_off, n.params + n.fc.b_off, p.ml, B, n.fc.fin, n.fc.fout, false, st);
  CHECK_CUDA_RET();
  return 0;
}

// Decoder.forward (:161-169): z [B,zdim] -> p.y NHWC [B,S,S,cdim]
static int dec_forward(sivae_engine* e, Net& n, DecPass& p, const float* z, int B, bool train, bool keep, cudaStream_t st,
                       const float* cond = nullptr) {
  const sivae_config& c = e->cfg;
  if (c.cond_dim > 0) {                           // z = torch.cat([z, y_cond], dim=1), :163-165
    if (!cond) return fail(-7, kCondMissing);
    concat_rows(z, c.zdim, cond, c.cond_dim, e->zc, B, st);
    z = e->zc;
  }
  p.zin = z;
  launch_linear_fwd(z, n.params + n.fc.w_off, n.params + n.fc.b_off, p.h, B, n.fc.fin, n.fc.fout, true, st);
  launch_nchw_to_nhwc(p.h, p.x0, B, e->C_last, e->hw_last, e->hw_last, st);
  if (e->fsplit) launch_split32(p.x0, p.x0s, (long long)B * e->feat, st);       // from the unrounded values
  if (e->bwd16) { if (keep) launch_to_bf16(p.x0, p.x0h, (long long)B * e->feat, st); }
  else if (e->rnd) launch_round_tf32(p.x0, p.x0, (long long)B * e->feat, st);
  const float* x = e->bwd16 ? p.x0h : p.x0;
  const float* xs = p.x0s;
  p.net = (train && (keep || !e->fsplit)) ? &n : nullptr;       // pass re-use needs the wgrad operands of this pass
  for (size_t i = 0; i < n.blocks.size(); ++i) {
    TRY(block_forward(e, n, n.blocks[i], p.blk[i], x, xs, B, train, keep, i + 1 == n.blocks.size(), st, train ? p.ema : nullptr));
    x = p.blk[i].out; xs = p.blk[i].outs;
  }
  TRY(conv_fwd(e, n, n.predict, e->fsplit ? xs : x, p.y, nullptr, B, c.image_size, st));
  CHECK_CUDA_RET();
  return 0;
}

// -------------------------------------------------------------------------------------------------------------
// backward passes (hand-written autograd of the above; dgrad-only when wgrad == false)
// -------------------------------------------------------------------------------------------------------------
// dout: gradient w.r.t. the block's (resampled) output; writes the gradient w.r.t. the block input into dx
static int block_backward(sivae_engine* e, Net& n, const Block& b, BlockAct& a, const float* dout, float* dx, bool wgrad, int B, cudaStream_t st) {
  const int s = b.size;
  float* DT = e->sb[2];
  float* G2 = e->sb[3];
  float* DA1 = e->sb[4];
  const float* idn = b.expand ? a.id : a.x;
  float* g = n.grads;
  // bwd16: the conv operands among the gradients (dt of both BatchNorms; the identity-branch gradient when it feeds conv_expand)
  // are written as plain bf16; the identity-branch gradient of a block without conv_expand stays the fp32 addend of dx
  const bool h16 = e->bwd16;
  { // reduce pass reads dout, t2 and the identity (or, with the forward's sign bytes, 1/16 of a pass instead of it); the apply
    // pass reads them again and writes dt and the identity-path gradient
    ProfElem pe(PC_BN_BWD, B, s, b.outc, 4 + b.mode, 2.0 * (resampled(b.mode) + (a.m2 ? 1.0625 : 2.0)) + (h16 ? (b.expand ? 1.0 : 1.5) : 2.0), st);
    launch_bn_act_bwd(dout, a.t2, idn, a.mi2, n.params + b.bn2.g_off, n.params + b.bn2.b_off, DT, G2,
                      wgrad ? g + b.bn2.g_off : nullptr, wgrad ? g + b.bn2.b_off : nullptr, true, B, s, s, b.outc, b.mode, e->rnd,
                      e->red, e->red_bytes, st, a.m2, h16 ? (1 | (b.expand ? 2 : 0)) : 0); }
  if (wgrad) TRY(conv_wgrad(e, n, b.c2, a.a1, DT, B, s, st, h16));
  TRY(conv_dgrad(e, n, b.c2, DT, DA1, nullptr, B, s, st, h16));
  { ProfElem pe(PC_BN_BWD, B, s, b.outc, RS_NONE, h16 ? 4.5 : 5.0, st);
    launch_bn_act_bwd(DA1, a.t1, nullptr, a.mi1, n.params + b.bn1.g_off, n.params + b.bn1.b_off, DT, nullptr,
                      wgrad ? g + b.bn1.g_off : nullptr, wgrad ? g + b.bn1.b_off : nullptr, true, B, s, s, b.outc, RS_NONE, e->rnd,
                      e->red, e->red_bytes, st, nullptr, h16 ? 1 : 0); }
  if (wgrad) {
    TRY(conv_wgrad(e, n, b.c1, a.x, DT, B, s, st, h16));
    if (b.expand) TRY(conv_wgrad(e, n, b.ce, a.x, G2, B, s, st, h16));
  }
  if (b.expand) {
    TRY(conv_dgrad(e, n, b.c1, DT, dx, nullptr, B, s, st, h16));
    TRY(conv_dgrad(e, n, b.ce, G2, dx, dx, B, s, st, h16));
  } else {
    TRY(conv_dgrad(e, n, b.c1, DT, dx, G2, B, s, st, h16));
  }
  return 0;
}

// dml: gradient w.r.t. the fc output [B,2z].  d_img (nullable): receives d loss / d input image (NHWC), plus
// d_img_addend if given (may alias d_img).
static int enc_backward(sivae_engine* e, Net& n, EncPass& p, const float* dml, bool wgrad, float* d_img, const float* d_img_addend, int B, cudaStream_t st) {
  const sivae_config& c = e->cfg;
  const int S = c.image_size;
  if (wgrad) launch_linear_wgrad(p.feat, dml, n.grads + n.fc.w_off, n.grads + n.fc.b_off, B, n.fc.fin, n.fc.fout, true, st);
  launch_linear_dgrad(dml, n.params + n.fc.w_off, e->dfeat, B, n.fc.fin, n.fc.fout, e->red, st);
  float* cur = e->sb[0];
  float* nxt = e->sb[1];
  launch_nchw_to_nhwc(e->dfeat, cur, B, e->C_last, e->hw_last, e->hw_last, st);
  for (int i = (int)n.blocks.size() - 1; i >= 0; --i) {
    TRY(block_backward(e, n, n.blocks[i], p.blk[i], cur, nxt, wgrad, B, st));
    float* t = cur; cur = nxt; nxt = t;
  }
  // stem: conv5x5 + BN + LeakyReLU + AvgPool (:89-92)
  float* DT = e->sb[2];
  { ProfElem pe(PC_BN_BWD, B, S, n.stem.cout, RS_POOL, 2.0 * 1.25 + 1.0, st);
    launch_bn_act_bwd(cur, p.t0, nullptr, p.mi0, n.params + n.stem_bn.g_off, n.params + n.stem_bn.b_off, DT, nullptr,
                      wgrad ? n.grads + n.stem_bn.g_off : nullptr, wgrad ? n.grads + n.stem_bn.b_off : nullptr, true, B, S, S,
                      n.stem.cout, RS_POOL, e->rnd, e->red, e->red_bytes, st); }
  if (wgrad) TRY(conv_wgrad(e, n, n.stem, p.img, DT, B, S, st));
  if (d_img) TRY(conv_dgrad(e, n, n.stem, DT, d_img, d_img_addend, B, S, st));
  CHECK_CUDA_RET();
  return 0;
}

// dy: gradient w.r.t. the decoder output (NHWC).  dz (nullable): gradient w.r.t. the latent input.
static int dec_backward(sivae_engine* e, Net& n, DecPass& p, const float* dy, bool wgrad, float* dz, int B, cudaStream_t st) {
  const sivae_config& c = e->cfg;
  const int S = c.image_size;
  float* cur = e->sb[0];
  float* nxt = e->sb[1];
  const float* xlast = p.blk.back().out;
  if (wgrad) {
    launch_colsum(dy, n.grads + n.predict.b_off, (long long)B * S * S, c.cdim, true, e->red, st);
    TRY(conv_wgrad(e, n, n.predict, xlast, dy, B, S, st));
  }
  TRY(conv_dgrad(e, n, n.predict, dy, cur, nullptr, B, S, st));
  for (int i = (int)n.blocks.size() - 1; i >= 0; --i) {
    TRY(block_backward(e, n, n.blocks[i], p.blk[i], cur, nxt, wgrad, B, st));
    float* t = cur; cur = nxt; nxt = t;
  }
  // view + ReLU + fc (:146-147, :166-167)
  launch_nhwc_to_nchw(cur, e->dfeat2, B, e->C_last, e->hw_last, e->hw_last, st);
  launch_relu_bwd(p.h, e->dfeat2, (long long)B * e->feat, st);
  if (wgrad) launch_linear_wgrad(p.zin, e->dfeat2, n.grads + n.fc.w_off, n.grads + n.fc.b_off, B, n.fc.fin, n.fc.fout, true, st);
  if (dz) launch_linear_dgrad(e->dfeat2, n.params + n.fc.w_off, dz, B, n.fc.fin, n.fc.fout, e->red, st);
  CHECK_CUDA_RET();
  return 0;
}

// -------------------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------------------
extern "C" int sivae_create(const sivae_config* cfg, sivae_engine** out) {
  if (!cfg || !out) return fail(-1, "null argument");
  if (cfg->n_channels < 1 || cfg->n_channels > 16) return fail(-2, "n_channels must be in [1,16]");
  if (cfg->cdim < 1 || cfg->cdim > 8 || cfg->zdim < 1 || cfg->max_batch < 1) return fail(-2, "bad cdim (1..8) / zdim / max_batch");
  int S = cfg->image_size;
  if (S < 2 || (S % (1 << cfg->n_channels)) != 0) return fail(-2, "image_size must be divisible by 2^len(channels)");
  for (int i = 0; i < cfg->n_channels; ++i)
    if (cfg->channels[i] < 4 || cfg->channels[i] % 4 != 0 || cfg->channels[i] > 1024) return fail(-2, "channels must be multiples of 4 in [4,1024]");
  if ((cfg->cdim * S * S) % 4 != 0) return fail(-2, "cdim*image_size^2 must be a multiple of 4");
  if (cfg->conv_backend < SIVAE_CONV_AUTO || cfg->conv_backend > SIVAE_CONV_TF32) return fail(-2, "bad conv_backend");
  if (cfg->cond_dim < 0 || cfg->cond_dim > 65536) return fail(-2, "cond_dim must be in [0,65536] (0 = unconditional)");
  sivae_engine* e = new sivae_engine();
  e->cfg = *cfg;
  e->comp = cfg->conv_backend == SIVAE_CONV_TC3X;
  { const char* v = getenv("SIVAE_BN_MASK"); e->bn_mask = !(v && v[0] == '0'); }
  { const char* v = getenv("SIVAE_REUSE_DEC"); e->reuse_dec = v && v[0] == '1'; }
  { const char* v = getenv("SIVAE_DP_OVERLAP"); e->dp_overlap = !(v && v[0] == '0'); }
  for (int i = 0; i < 3; ++i) { e->nets[i].id = i; e->nets[i].comp = e->comp; }
  build_encoder(e, e->nets[0]);
  build_decoder(e, e->nets[1]);
  if (cfg->variant == 1) build_decoder(e, e->nets[2]);
  e->tc = cfg->conv_backend != SIVAE_CONV_SIMT;
  e->fast = cfg->conv_backend != SIVAE_CONV_SIMT;
  e->rnd = e->tc && !e->comp;
  // split-forward mode (the default): every forward conv must be served by a tensor-core kernel -- stem / predict by the
  // row-separable form, the blocks by the implicit-GEMM kernels -- and every activation must have whole 32-channel groups
  e->fsplit = cfg->conv_backend == SIVAE_CONV_AUTO;
  { const char* v = getenv("SIVAE_FWD_SPLIT"); if (v && v[0] == '0') e->fsplit = false; }
  if (e->fsplit) {
    e->rs = (float*)1;                       // fwd_on_rowsep_* test for "scratch present"; carve() sets the real pointer
    bool ok = fwd_on_rowsep_in(e, ConvShape{1, S, S, cfg->cdim, cfg->channels[0], 5}) &&
              fwd_on_rowsep_out(e, ConvShape{1, S, S, cfg->channels[0], cfg->cdim, 5});
    for (int ni = 0; ni < 2 && ok; ++ni)
      for (const Block& b : e->nets[ni].blocks) {
        ok = ok && b.inc % 32 == 0 && b.outc % 32 == 0 && fwd_on_tc(e, ConvShape{1, b.size, b.size, b.inc, b.outc, 3}) &&
             fwd_on_tc(e, ConvShape{1, b.size, b.size, b.outc, b.outc, 3});
      }
    e->rs = nullptr;
    e->fsplit = ok;
    // AUTO promises every logged scalar within 1e-4 of the reference.  An architecture the split-forward kernels do not take
    // (channel counts that are not multiples of 32 -- celeb1024 starts at 16 --, image sizes the row-separable stem / predict
    // do not tile -- 28x28 --) therefore degrades to the compensated 3xTF32 mode (same bound, ~3x the conv time; exact fp32
    // kernels wherever a shape is not tensor-core eligible), NOT to plain TF32 (2e-3).  SIVAE_AUTO_FALLBACK=tf32 keeps the
    // faster, looser round-1 behaviour.
    if (!ok && cfg->conv_backend == SIVAE_CONV_AUTO && !getenv("SIVAE_FWD_SPLIT")) {
      const char* v = getenv("SIVAE_AUTO_FALLBACK");
      if (!(v && strcmp(v, "tf32") == 0)) {
        sivae_config c2 = *cfg;
        c2.conv_backend = SIVAE_CONV_TC3X;
        delete e;
        return sivae_create(&c2, out);
      }
    }
  }
  if (e->fsplit) {
    bool ok = true;
    { const char* v = getenv("SIVAE_BWD16"); if (v && v[0] == '0') ok = false; }
    for (int ni = 0; ni < 2 && ok; ++ni)
      for (const Block& b : e->nets[ni].blocks) {
        const ConvShape s1{1, b.size, b.size, b.inc, b.outc, 3}, s2{1, b.size, b.size, b.outc, b.outc, 3};
        ok = ok && b.inc % 64 == 0 && b.outc % 64 == 0 && conv_tc_supported_wgrad16(s1) && conv_tc_supported_wgrad16(s2) &&
             conv_tc_supported_fwd(ConvShape{1, b.size, b.size, b.outc, b.inc, 3}, FMT_BF16_) &&
             conv_tc_supported_fwd(ConvShape{1, b.size, b.size, b.outc, b.outc, 3}, FMT_BF16_);
      }
    e->bwd16 = ok;
  }
  if (e->fsplit) e->bn_mask = true;          // the backward must not need the fp32 identity tensor (absent in dgrad-only passes)
  e->ws_need = carve(e, nullptr);
  *out = e;
  return 0;
}
// ---- NCCL, resolved at run time from the libnccl the process already has (torch's bundled one) or the system's ----------
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("SIVAE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    if (!nm || !nm[0]) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return nullptr;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.CommCount = (decltype(api.CommCount))dlsym(h, "ncclCommCount");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.CommCount && api.GetErrorString;
  return api.ok ? &api : nullptr;
}
int nccl_fail(const NcclApi* a, ncclResult_t r, const char* what) {
  return fail(-20 - (int)r, std::string(what) + ": " + (a && a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}
}  // namespace

extern "C" int sivae_comm_unique_id(unsigned char* out128) {
  NcclApi* a = nccl_api();
  if (!a) return fail(-10, "libnccl.so.2 not found (set SIVAE_NCCL_LIB)");
  if (!out128) return fail(-1, "null argument");
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(a, r, "ncclGetUniqueId");
  memcpy(out128, &id, sizeof(id));
  return 0;
}
// The library's communicator is PROCESS-GLOBAL (one process = one GPU = one rank): created once by the first
// sivae_comm_init, shared by every engine of the process, and destroyed only by an explicit sivae_comm_finalize -- never from
// an engine's destructor (ncclCommDestroy from garbage collection or interpreter shutdown runs at rank-dependent times and
// after CUDA teardown: both hang).  A process that exits without finalizing simply leaves the cleanup to the OS.
static ncclComm_t g_comm = nullptr;
static int g_comm_world = 0, g_comm_rank = -1;
static void comm_release(sivae_engine* e) { e->comm = nullptr; e->world = 1; e->own_comm = false; }
extern "C" int sivae_comm_global_world(void) { return g_comm ? g_comm_world : 0; }
// id128 == NULL: attach the process-global communicator that an earlier call created (same world / rank)
extern "C" int sivae_comm_init(sivae_engine* e, const unsigned char* id128, int world, int rank) {
  NcclApi* a = nccl_api();
  if (!a) return fail(-10, "libnccl.so.2 not found (set SIVAE_NCCL_LIB)");
  if (!e || world < 1 || rank < 0 || rank >= world) return fail(-1, "bad argument");
  comm_release(e);
  if (g_comm) {
    if (g_comm_world != world || g_comm_rank != rank) return fail(-11, "the process-global communicator has another world size / rank");
  } else {
    if (!id128) return fail(-11, "no process-global communicator yet: pass the ncclUniqueId");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    ncclResult_t r = a->CommInitRank(&c, world, id, rank);          // on the calling thread's current device
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclCommInitRank");
    g_comm = c; g_comm_world = world; g_comm_rank = rank;
  }
  e->comm = g_comm; e->world = world; e->own_comm = false;
  return 0;
}
// explicit teardown (after the last step of every engine, all ranks): synchronises the device, then ncclCommDestroy
extern "C" int sivae_comm_finalize(void) {
  if (!g_comm) return 0;
  NcclApi* a = nccl_api();
  cudaDeviceSynchronize();
  ncclResult_t r = a ? a->CommDestroy(g_comm) : ncclSuccess;
  g_comm = nullptr; g_comm_world = 0; g_comm_rank = -1;
  if (r != ncclSuccess) return nccl_fail(a, r, "ncclCommDestroy");
  return 0;
}
extern "C" int sivae_allreduce_attach(sivae_engine* e, void* nccl_comm) {
  NcclApi* a = nccl_api();
  if (!a) return fail(-10, "libnccl.so.2 not found (set SIVAE_NCCL_LIB)");
  if (!e) return fail(-1, "null engine");
  comm_release(e);
  if (!nccl_comm) return 0;                                         // detach
  int n = 0;
  ncclResult_t r = a->CommCount((ncclComm_t)nccl_comm, &n);
  if (r != ncclSuccess) return nccl_fail(a, r, "ncclCommCount");
  e->comm = (ncclComm_t)nccl_comm; e->world = n; e->own_comm = false;
  return 0;
}
extern "C" int sivae_comm_world(const sivae_engine* e) { return e ? e->world : -1; }
extern "C" int sivae_allreduce_grads(sivae_engine* e, int net, void* stream) {
  if (!e || net < 0 || net > 2 || !e->nets[net].present) return fail(-1, "bad net id");
  Net& n = e->nets[net];
  if (!n.grads) return fail(-4, "grads not bound");
  if (!e->comm || e->world <= 1) return 0;
  NcclApi* a = nccl_api();
  ncclResult_t r = a->AllReduce(n.grads, n.grads, (size_t)n.n_params, ncclFloat32, ncclSum, e->comm, (cudaStream_t)stream);
  if (r != ncclSuccess) return nccl_fail(a, r, "ncclAllReduce");
  g_launches += 1;
  return 0;
}
extern "C" void sivae_destroy(sivae_engine* e) {
  if (e) {
    comm_release(e);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->side) cudaStreamDestroy(e->side);
  }
  delete e;
}
static Net* get_net(sivae_engine* e, int net) {
  if (!e || net < 0 || net > 2 || !e->nets[net].present) return nullptr;
  return &e->nets[net];
}
extern "C" int sivae_num_tensors(const sivae_engine* e, int net) {
  Net* n = get_net(const_cast<sivae_engine*>(e), net);
  return n ? (int)n->tinfo.size() : -1;
}
extern "C" int sivae_tensor(const sivae_engine* e, int net, int i, sivae_tensor_info* out) {
  Net* n = get_net(const_cast<sivae_engine*>(e), net);
  if (!n || i < 0 || i >= (int)n->tinfo.size() || !out) return fail(-1, "bad tensor index");
  *out = n->tinfo[i];
  return 0;
}
extern "C" long long sivae_param_count(const sivae_engine* e, int net) {
  Net* n = get_net(const_cast<sivae_engine*>(e), net);
  return n ? n->n_params : -1;
}
extern "C" int sivae_num_bn(const sivae_engine* e, int net) {
  Net* n = get_net(const_cast<sivae_engine*>(e), net);
  return n ? (int)n->binfo.size() : -1;
}
extern "C" int sivae_bn(const sivae_engine* e, int net, int i, sivae_bn_info* out) {
  Net* n = get_net(const_cast<sivae_engine*>(e), net);
  if (!n || i < 0 || i >= (int)n->binfo.size() || !out) return fail(-1, "bad bn index");
  *out = n->binfo[i];
  return 0;
}
extern "C" long long sivae_bn_floats(const sivae_engine* e, int net) {
  Net* n = get_net(const_cast<sivae_engine*>(e), net);
  return n ? n->bn_floats : -1;
}
extern "C" long long sivae_workspace_bytes(const sivae_engine* e) { return e ? (long long)e->ws_need : -1; }

extern "C" int sivae_bind_net(sivae_engine* e, int net, float* params, float* grads, float* m, float* v, float* bn, long long* nbt) {
  Net* n = get_net(e, net);
  if (!n) return fail(-1, "bad net id");
  if (!params || !bn || !nbt) return fail(-1, "params / bn buffers must not be null");
  n->params = params; n->grads = grads; n->m = m; n->v = v; n->bn = bn; n->nbt = nbt;
  n->dirty = true;
  return 0;
}
extern "C" int sivae_bind_workspace(sivae_engine* e, void* ws, long long bytes) {
  if (!e || !ws) return fail(-1, "null argument");
  if ((size_t)bytes < e->ws_need) return fail(-3, "workspace too small");
  if (((uintptr_t)ws & 255) != 0) return fail(-3, "workspace must be 256-byte aligned");
  e->ws = ws; e->ws_bytes = (size_t)bytes;
  carve(e, (char*)ws);
  for (int i = 0; i < 3; ++i) {
    e->nets[i].dirty = true;
    if (e->nets[i].present) cudaMemcpy(e->nets[i].step_dev, &e->nets[i].step_pending, sizeof(long long), cudaMemcpyHostToDevice);
  }
  e->have_e_state = false;
  e->e_dec_valid = false;
  return 0;
}
extern "C" int sivae_params_changed(sivae_engine* e, int net) {
  Net* n = get_net(e, net);
  if (!n) return fail(-1, "bad net id");
  n->dirty = true;
  return 0;
}
static int check_ready(sivae_engine* e, int batch) {
  if (!e) return fail(-1, "null engine");
  if (!e->ws) return fail(-4, "workspace not bound");
  for (int i = 0; i < 3; ++i)
    if (e->nets[i].present && !e->nets[i].params) return fail(-4, "net not bound");
  if (batch < 1 || batch > e->cfg.max_batch) return fail(-5, "batch out of range");
  return 0;
}

extern "C" int sivae_e_step(sivae_engine* e, const float* real_nchw, const float* noise, const float* eps, int B,
                            const sivae_hyper* hp, float* stats, void* stream) {
  TRY(check_ready(e, B));
  if (!real_nchw || !noise || !eps || !hp || !stats) return fail(-1, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const sivae_config& c = e->cfg;
  const int S = c.image_size, z = c.zdim;
  const long long per = (long long)c.cdim * S * S;
  const bool boot = c.variant == 1;
  Net& en = e->nets[0];
  Net& dn = e->nets[1];
  Net& tn = boot ? e->nets[2] : e->nets[1];
  if (!en.grads) return fail(-4, "encoder grads not bound");
  TRY(refresh_derived(e, en, st)); TRY(refresh_derived(e, dn, st));
  if (boot) TRY(refresh_derived(e, tn, st));
  launch_nchw_to_nhwc(real_nchw, e->real, B, c.cdim, S, S, st);
  cudaMemcpyAsync(e->noise, noise, sizeof(float) * B * z, cudaMemcpyDeviceToDevice, st);
  const float *eps1 = eps, *eps2 = eps + (long long)B * z, *eps3 = eps + 2LL * B * z;
  EncPass &E1 = e->ep[0], &E2 = e->ep[1], &E3 = e->ep[2];
  DecPass &D1 = e->dp[0], &D2 = e->dp[1], &D3 = e->dp[2], &D4 = e->dp[3];
  // forwards in the reference's per-net order (BN running stats are order dependent): :557-568
  // (keep flags: the encoder passes get parameter gradients in this half; the decoder passes are dgrad-only -- unless the D
  // half is going to re-use D1 / D2, whose wgrad operands must then exist)
  const bool kd = e->reuse_dec;
  TRY(dec_forward(e, dn, D1, e->noise, B, true, kd, st));             // fake
  TRY(enc_forward(e, en, E1, e->real, B, true, true, st));
  launch_kl_reparam(E1.ml, eps1, e->z_keep, E1.kl, B, z, st);          // z (kept for the D half, :598)
  TRY(dec_forward(e, dn, D2, e->z_keep, B, true, kd, st));            // rec
  TRY(enc_forward(e, en, E2, D2.y, B, true, true, st));               // model(rec.detach())
  launch_kl_reparam(E2.ml, eps2, E2.z, E2.kl, B, z, st);
  TRY(dec_forward(e, tn, D3, E2.z, B, true, false, st));              // rec_rec
  TRY(enc_forward(e, en, E3, D1.y, B, true, true, st));               // model(fake.detach())
  launch_kl_reparam(E3.ml, eps3, E3.z, E3.kl, B, z, st);
  TRY(dec_forward(e, tn, D4, E3.z, B, true, false, st));              // rec_fake
  // losses :563-586
  const int lt = e->loss_type;
  const float md = e->mean_div();
  { ProfLoss pl(B, per, st); launch_mse3(e->real, D2.y, D3.y, D1.y, D4.y, e->mse, B, per, e->red, e->red_bytes, st, lt, e->bad_flag()); }
  launch_e_loss_finalize(e->mse, E1.kl, E2.kl, E3.kl, B, hp->beta_kl, hp->beta_rec, hp->beta_neg, hp->scale, stats,
                         e->coef, e->ckl_a, e->ckl_b, st, md, e->bad_flag());
  const float a_rec = 2.f * hp->scale * hp->beta_rec / (float)B * md;
  launch_loss_seed(e->real, D2.y, D3.y, D1.y, D4.y, a_rec, e->coef + B, 0.f, e->coef + 2 * B, 0.f, /*rec not detached :573*/ true,
                   e->d_rec, e->d_rec_rec, e->d_rec_fake, nullptr, B, per, st, lt);
  // backward :587-588 -- only the encoder accumulates parameter grads; decoders are dgrad-only
  cudaMemsetAsync(en.grads, 0, sizeof(float) * en.n_params, st);
  TRY(dec_backward(e, tn, D3, e->d_rec_rec, false, e->dz, B, st));
  launch_latent_bwd(E2.ml, eps2, e->dz, e->ckl_a, 0.f, e->dml, B, z, st);
  TRY(enc_backward(e, en, E2, e->dml, true, nullptr, nullptr, B, st));
  TRY(dec_backward(e, tn, D4, e->d_rec_fake, false, e->dz, B, st));
  launch_latent_bwd(E3.ml, eps3, e->dz, e->ckl_b, 0.f, e->dml, B, z, st);
  TRY(enc_backward(e, en, E3, e->dml, true, nullptr, nullptr, B, st));
  TRY(dec_backward(e, dn, D2, e->d_rec, false, e->dz, B, st));
  launch_latent_bwd(E1.ml, eps1, e->dz, nullptr, hp->scale * hp->beta_kl / (float)B, e->dml, B, z, st);
  TRY(enc_backward(e, en, E1, e->dml, true, nullptr, nullptr, B, st));
  CHECK_CUDA_RET();
  e->cur_batch = B;
  e->have_e_state = true;
  e->e_dec_valid = true;
  return 0;
}

// join (nullable): an event recorded on ANOTHER stream after the encoder's Adam step; the D half then starts with its two
// encoder-independent decoder passes, waits for the event and only then touches the encoder's (refreshed) filters
static int d_step_impl(sivae_engine* e, const float* eps, const sivae_hyper* hp, float* stats, void* stream, cudaEvent_t join);
extern "C" int sivae_d_step(sivae_engine* e, const float* eps, const sivae_hyper* hp, float* stats, void* stream) {
  return d_step_impl(e, eps, hp, stats, stream, nullptr);
}
static int d_step_impl(sivae_engine* e, const float* eps, const sivae_hyper* hp, float* stats, void* stream, cudaEvent_t join) {
  if (!e || !e->have_e_state) return fail(-6, "sivae_d_step requires a preceding sivae_e_step");
  const int B = e->cur_batch;
  TRY(check_ready(e, B));
  if (!eps || !hp || !stats) return fail(-1, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const sivae_config& c = e->cfg;
  const int S = c.image_size, z = c.zdim;
  const long long per = (long long)c.cdim * S * S;
  const bool boot = c.variant == 1;
  Net& en = e->nets[0];
  Net& dn = e->nets[1];
  Net& tn = boot ? e->nets[2] : e->nets[1];
  if (!dn.grads) return fail(-4, "decoder grads not bound");
  // pass re-use (opt-in): the decoder weights have not moved since the E half (only Adam(encoder) ran), so fake = D(noise)
  // and rec = D(z) (:597-598) are bit for bit the E half's passes (:557,:561), whose activations still sit in dp[0] / dp[1]
  const bool reuse = e->reuse_dec && e->e_dec_valid && !dn.dirty && e->dp[0].net == &dn && e->dp[1].net == &dn;
  e->e_dec_valid = false;
  if (!join) TRY(refresh_derived(e, en, st));
  TRY(refresh_derived(e, dn, st));
  if (boot) TRY(refresh_derived(e, tn, st));
  const float *eps4 = eps, *eps5 = eps + (long long)B * z;
  EncPass &E4 = e->ep[0], &E5 = e->ep[1];
  DecPass &D5 = e->dp[0], &D6 = e->dp[1], &D7 = e->dp[2], &D8 = e->dp[3];
  if (reuse) {
    // the only new effect of recomputing them: one more running-statistics update per pass, in this order
    launch_bn_ema_replay(D5.ema, dn.bn, dn.bn_floats, dn.nbt, (int)dn.binfo.size(), st);
    launch_bn_ema_replay(D6.ema, dn.bn, dn.bn_floats, dn.nbt, (int)dn.binfo.size(), st);
  } else {
    TRY(dec_forward(e, dn, D5, e->noise, B, true, true, st));         // fake :597
    TRY(dec_forward(e, dn, D6, e->z_keep, B, true, true, st));        // rec  :598
  }
  if (join) {                         // the encoder's all-reduce + Adam step ran beside the two passes above
    cudaStreamWaitEvent(st, join, 0);
    TRY(refresh_derived(e, en, st));
  }
  TRY(enc_forward(e, en, E4, D6.y, B, true, false, st));              // :601  (encoder: dgrad-only in this half)
  launch_kl_reparam(E4.ml, eps4, E4.z, E4.kl, B, z, st);
  TRY(enc_forward(e, en, E5, D5.y, B, true, false, st));              // :604
  launch_kl_reparam(E5.ml, eps5, E5.z, E5.kl, B, z, st);
  TRY(dec_forward(e, tn, D7, E4.z, B, true, !boot, st));              // rec_rec :607 (bootstrap: frozen target decoder)
  TRY(dec_forward(e, tn, D8, E5.z, B, true, !boot, st));              // rec_fake :608
  const int lt = e->loss_type;
  const float md = e->mean_div();
  { ProfLoss pl(B, per, st); launch_mse3(e->real, D6.y, D7.y, D5.y, D8.y, e->mse, B, per, e->red, e->red_bytes, st, lt, e->bad_flag()); }
  launch_d_loss_finalize(e->mse, E4.kl, E5.kl, B, hp->beta_kl, hp->beta_rec, hp->gamma_r, hp->scale, stats, st, md, e->bad_flag());
  const float a_rec = 2.f * hp->scale * hp->beta_rec / (float)B * md;
  const float a_t = hp->scale * hp->gamma_r * hp->beta_rec / (float)B * md;     // 2 * (scale * gamma_r/2 * beta_rec / B)
  const float ckl = hp->scale * 0.5f * hp->beta_kl / (float)B;
  // standard: targets detached (:610-613); bootstrap: nothing detached (bootstrap :635-641)
  launch_loss_seed(e->real, D6.y, D7.y, D5.y, D8.y, a_rec, nullptr, a_t, nullptr, a_t, boot, e->d_rec, e->d_rec_rec,
                   e->d_rec_fake, boot ? e->d_fake : nullptr, B, per, st, lt);
  cudaMemsetAsync(dn.grads, 0, sizeof(float) * dn.n_params, st);
  if (!boot) {
    TRY(dec_backward(e, dn, D7, e->d_rec_rec, true, nullptr, B, st));
    TRY(dec_backward(e, dn, D8, e->d_rec_fake, true, nullptr, B, st));
    launch_latent_bwd(E4.ml, eps4, nullptr, nullptr, ckl, e->dml, B, z, st);
    TRY(enc_backward(e, en, E4, e->dml, false, e->d_rec, e->d_rec, B, st));
    launch_latent_bwd(E5.ml, eps5, nullptr, nullptr, ckl, e->dml, B, z, st);
    TRY(enc_backward(e, en, E5, e->dml, false, e->d_fake, nullptr, B, st));
  } else {
    TRY(dec_backward(e, tn, D7, e->d_rec_rec, false, e->dz, B, st));
    launch_latent_bwd(E4.ml, eps4, e->dz, nullptr, ckl, e->dml, B, z, st);
    TRY(enc_backward(e, en, E4, e->dml, false, e->d_rec, e->d_rec, B, st));
    TRY(dec_backward(e, tn, D8, e->d_rec_fake, false, e->dz, B, st));
    launch_latent_bwd(E5.ml, eps5, e->dz, nullptr, ckl, e->dml, B, z, st);
    TRY(enc_backward(e, en, E5, e->dml, false, e->d_fake, e->d_fake, B, st));
  }
  TRY(dec_backward(e, dn, D6, e->d_rec, true, nullptr, B, st));
  TRY(dec_backward(e, dn, D5, e->d_fake, true, nullptr, B, st));
  CHECK_CUDA_RET();
  return 0;
}

extern "C" int sivae_vae_step(sivae_engine* e, const float* real_nchw, const float* eps, int B, const sivae_hyper* hp,
                              float* stats, void* stream) {
  TRY(check_ready(e, B));
  if (!real_nchw || !eps || !hp || !stats) return fail(-1, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const sivae_config& c = e->cfg;
  const int S = c.image_size, z = c.zdim;
  const long long per = (long long)c.cdim * S * S;
  // bootstrap variant: model(real_batch) decodes with the frozen TARGET decoder (bootstrap trainer :196-217, default
  // target=True; warm-up call at :546), so only the encoder receives gradients -- the decoder pass is dgrad-only
  const bool boot = c.variant == 1;
  Net& en = e->nets[0];
  Net& dn = boot ? e->nets[2] : e->nets[1];
  if (!en.grads || (!boot && !dn.grads)) return fail(-4, "grads not bound");
  TRY(refresh_derived(e, en, st)); TRY(refresh_derived(e, dn, st));
  launch_nchw_to_nhwc(real_nchw, e->real, B, c.cdim, S, S, st);
  EncPass& E1 = e->ep[0];
  DecPass& D1 = e->dp[0];
  TRY(enc_forward(e, en, E1, e->real, B, true, true, st));            // model(real_batch) :518
  launch_kl_reparam(E1.ml, eps, E1.z, E1.kl, B, z, st);
  TRY(dec_forward(e, dn, D1, E1.z, B, true, !boot, st));
  launch_mse3(e->real, D1.y, nullptr, nullptr, nullptr, e->mse, B, per, e->red, e->red_bytes, st, e->loss_type, e->bad_flag());
  launch_vae_loss_finalize(e->mse, E1.kl, B, hp->beta_kl, hp->beta_rec, stats, st, e->mean_div(), e->bad_flag());
  launch_loss_seed(e->real, D1.y, nullptr, nullptr, nullptr, 2.f * hp->beta_rec / (float)B * e->mean_div(), nullptr, 0.f, nullptr, 0.f, false,
                   e->d_rec, nullptr, nullptr, nullptr, B, per, st, e->loss_type);
  cudaMemsetAsync(en.grads, 0, sizeof(float) * en.n_params, st);
  if (!boot) cudaMemsetAsync(dn.grads, 0, sizeof(float) * dn.n_params, st);
  TRY(dec_backward(e, dn, D1, e->d_rec, !boot, e->dz, B, st));
  launch_latent_bwd(E1.ml, eps, e->dz, nullptr, hp->beta_kl / (float)B, e->dml, B, z, st);
  TRY(enc_backward(e, en, E1, e->dml, true, nullptr, nullptr, B, st));
  CHECK_CUDA_RET();
  e->have_e_state = false;
  e->e_dec_valid = false;
  return 0;
}

extern "C" int sivae_set_reuse_decoder_passes(sivae_engine* e, int on) {
  if (!e) return fail(-1, "null engine");
  e->reuse_dec = on != 0;
  return 0;
}
extern "C" int sivae_get_reuse_decoder_passes(const sivae_engine* e) { return e ? (e->reuse_dec ? 1 : 0) : -1; }
// recon_loss_type kwarg of train_soft_intro_vae (:339) -> calc_reconstruction_loss(loss_type=...) at :563,573,576,599,610,612
extern "C" int sivae_set_recon_loss(sivae_engine* e, int loss_type) {
  if (!e) return fail(-1, "null engine");
  if (loss_type < SIVAE_LOSS_MSE_ || loss_type > SIVAE_LOSS_BCE_) return fail(-2, "recon loss type must be SIVAE_LOSS_MSE / _L1 / _BCE");
  e->loss_type = loss_type;
  return 0;
}
extern "C" int sivae_get_recon_loss(const sivae_engine* e) { return e ? e->loss_type : -1; }

extern "C" int sivae_adam_step(sivae_engine* e, int net, float lr, float grad_scale, void* stream) {
  Net* n = get_net(e, net);
  if (!n) return fail(-1, "bad net id");
  if (!n->grads || !n->m || !n->v) return fail(-4, "optimiser buffers not bound");
  if (!n->step_dev) return fail(-4, "workspace not bound");
  launch_adam(n->params, n->grads, n->m, n->v, n->n_params, lr, grad_scale, 0.9f, 0.999f, 1e-8f, n->step_dev, n->coef_dev,
              (cudaStream_t)stream);
  n->dirty = true;
  CHECK_CUDA_RET();
  // The derived operand copies (split32 / bf16 / tf32 filters) are refreshed HERE, on the same stream, not lazily at their next
  // use: the lazy form made the kernel sequence of a step depend on a host flag, and a CUDA graph captured right after an
  // inference call (the trainer's iteration-0 sample grid: sivae_decode refreshes the decoder and clears its flag) then
  // lacked the decoder refresh -- every replay ran the decoder on the filters of the capture iteration.  Found by
  // tests/test_gpu_step.py::test_resume_from_checkpoint_is_bit_identical_to_uninterrupted_training.
  if (e->ws) TRY(refresh_derived(e, *n, (cudaStream_t)stream));
  return 0;
}
// One whole introspective iteration (:551-624) as a single enqueue: E half, [all-reduce of the encoder gradients],
// Adam(encoder), D half, [all-reduce of the decoder gradients], Adam(decoder) -- the collectives are raw ncclAllReduce calls on
// the same stream (sivae_comm_init / sivae_allreduce_attach), so the iteration is capturable as ONE CUDA graph under data
// parallelism too.  eps: [5,B,z] in the draw order of :560,:567,:568,:602,:605.  Gradients are averaged (1 / world in Adam).
extern "C" int sivae_iteration(sivae_engine* e, const float* real_nchw, const float* noise, const float* eps5, int B,
                               const sivae_hyper* hp, float lr_e, float lr_d, float* stats, void* stream) {
  if (!e || !eps5) return fail(-1, "null argument");
  const float gs = 1.f / (float)(e->world > 0 ? e->world : 1);
  TRY(sivae_e_step(e, real_nchw, noise, eps5, B, hp, stats, stream));
  if (e->comm && e->world > 1 && e->dp_overlap && !e->reuse_dec) {
    if (!e->side) {
      if (cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess)
        return fail(-3, "could not create the side stream of the data-parallel overlap");
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaEventRecord(e->ev_fork, st);
    cudaStreamWaitEvent(e->side, e->ev_fork, 0);
    TRY(sivae_allreduce_grads(e, SIVAE_NET_ENCODER, e->side));
    TRY(sivae_adam_step(e, SIVAE_NET_ENCODER, lr_e, gs, e->side));
    cudaEventRecord(e->ev_join, e->side);
    TRY(d_step_impl(e, eps5 + 3LL * B * e->cfg.zdim, hp, stats, stream, e->ev_join));
  } else {
    TRY(sivae_allreduce_grads(e, SIVAE_NET_ENCODER, stream));
    TRY(sivae_adam_step(e, SIVAE_NET_ENCODER, lr_e, gs, stream));
    TRY(sivae_d_step(e, eps5 + 3LL * B * e->cfg.zdim, hp, stats, stream));
  }
  TRY(sivae_allreduce_grads(e, SIVAE_NET_DECODER, stream));
  TRY(sivae_adam_step(e, SIVAE_NET_DECODER, lr_d, gs, stream));
  return 0;
}
extern "C" int sivae_adam_set_step(sivae_engine* e, int net, long long step) {
  Net* n = get_net(e, net);
  if (!n) return fail(-1, "bad net id");
  n->step_pending = step;
  if (n->step_dev) cudaMemcpy(n->step_dev, &step, sizeof(long long), cudaMemcpyHostToDevice);
  return 0;
}
// synchronises (reads the device counter)
extern "C" long long sivae_adam_get_step(const sivae_engine* e, int net) {
  Net* n = get_net(const_cast<sivae_engine*>(e), net);
  if (!n) return -1;
  if (!n->step_dev) return n->step_pending;
  long long s = 0;
  cudaDeviceSynchronize();
  cudaMemcpy(&s, n->step_dev, sizeof(long long), cudaMemcpyDeviceToHost);
  return s;
}

static int encode_impl(sivae_engine* e, const float* x_nchw, const float* cond, int B, float* mu, float* logvar, int train, void* stream);
extern "C" int sivae_encode(sivae_engine* e, const float* x_nchw, int B, float* mu, float* logvar, int train, void* stream) {
  return encode_impl(e, x_nchw, nullptr, B, mu, logvar, train, stream);
}
// Encoder.forward(x, o_cond) of a conditional model (:116-122): cond = [B, cond_dim] device rows
extern "C" int sivae_encode_cond(sivae_engine* e, const float* x_nchw, const float* cond, int B, float* mu, float* logvar, int train,
                                 void* stream) {
  if (!e || e->cfg.cond_dim <= 0) return fail(-2, "sivae_encode_cond needs an engine created with cond_dim > 0");
  if (!cond) return fail(-1, "null argument");
  return encode_impl(e, x_nchw, cond, B, mu, logvar, train, stream);
}
static int encode_impl(sivae_engine* e, const float* x_nchw, const float* cond, int B, float* mu, float* logvar, int train, void* stream) {
  TRY(check_ready(e, B));
  if (!x_nchw || !mu || !logvar) return fail(-1, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const sivae_config& c = e->cfg;
  Net& en = e->nets[0];
  TRY(refresh_derived(e, en, st));
  launch_nchw_to_nhwc(x_nchw, e->out_tmp, B, c.cdim, c.image_size, c.image_size, st);
  EncPass& p = e->ep[2];
  TRY(enc_forward(e, en, p, e->out_tmp, B, train != 0, false, st, cond));
  cudaMemcpy2DAsync(mu, sizeof(float) * c.zdim, p.ml, sizeof(float) * 2 * c.zdim, sizeof(float) * c.zdim, B, cudaMemcpyDeviceToDevice, st);
  cudaMemcpy2DAsync(logvar, sizeof(float) * c.zdim, p.ml + c.zdim, sizeof(float) * 2 * c.zdim, sizeof(float) * c.zdim, B, cudaMemcpyDeviceToDevice, st);
  CHECK_CUDA_RET();
  return 0;
}
static int decode_impl(sivae_engine* e, int net, const float* z, const float* cond, int B, float* out_nchw, int train, void* stream);
extern "C" int sivae_decode(sivae_engine* e, int net, const float* z, int B, float* out_nchw, int train, void* stream) {
  return decode_impl(e, net, z, nullptr, B, out_nchw, train, stream);
}
// Decoder.forward(z, y_cond) of a conditional model (:161-169)
extern "C" int sivae_decode_cond(sivae_engine* e, int net, const float* z, const float* cond, int B, float* out_nchw, int train,
                                 void* stream) {
  if (!e || e->cfg.cond_dim <= 0) return fail(-2, "sivae_decode_cond needs an engine created with cond_dim > 0");
  if (!cond) return fail(-1, "null argument");
  return decode_impl(e, net, z, cond, B, out_nchw, train, stream);
}
static int decode_impl(sivae_engine* e, int net, const float* z, const float* cond, int B, float* out_nchw, int train, void* stream) {
  TRY(check_ready(e, B));
  Net* n = get_net(e, net);
  if (!n || n->enc) return fail(-1, "bad decoder net id");
  if (!z || !out_nchw) return fail(-1, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const sivae_config& c = e->cfg;
  TRY(refresh_derived(e, *n, st));
  DecPass& p = e->dp[3];
  TRY(dec_forward(e, *n, p, z, B, train != 0, false, st, cond));
  launch_nhwc_to_nchw(p.y, out_nchw, B, c.cdim, c.image_size, c.image_size, st);
  CHECK_CUDA_RET();
  return 0;
}

extern "C" unsigned long long sivae_launch_count(void) { return sivae::g_launches; }
extern "C" int sivae_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.used = 0;
  return 0;
}
// out[class][3] = {milliseconds, flops, launches} for class 0 = tcgen05 conv fwd/dgrad, 1 = tcgen05 wgrad,
// 2 = SIMT conv fwd/dgrad, 3 = SIMT wgrad.  Synchronises the device.
// per-(class, shape) table of the recorded launches as text lines "cls N H W Cin Cout k launches ms gflop"; call BEFORE
// sivae_profile_read (which resets the recorder).  Returns the number of bytes written (truncated to len-1).
extern "C" int sivae_profile_dump(char* buf, int len) {
  if (!buf || len < 2) return fail(-1, "bad buffer");
  cudaDeviceSynchronize();
  struct Agg { int cls; ConvShape s; int n; double ms, fl; };
  std::vector<Agg> aggs;
  for (size_t i = 0; i < g_prof.used; ++i) {
    const ProfRec& r = g_prof.recs[i];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    bool found = false;
    for (Agg& a : aggs)
      if (a.cls == r.cls && a.s.N == r.shape.N && a.s.H == r.shape.H && a.s.W == r.shape.W && a.s.Cin == r.shape.Cin &&
          a.s.Cout == r.shape.Cout && a.s.k == r.shape.k) { a.n++; a.ms += ms; a.fl += r.flops; found = true; break; }
    if (!found) aggs.push_back(Agg{r.cls, r.shape, 1, ms, r.flops});
  }
  int off = 0;
  for (const Agg& a : aggs) {
    int w = snprintf(buf + off, len - off, "%d %d %d %d %d %d %d %d %.4f %.3f\n", a.cls, a.s.N, a.s.H, a.s.W, a.s.Cin, a.s.Cout,
                     a.s.k, a.n, a.ms, a.fl * 1e-9);
    if (w < 0 || w >= len - off) break;
    off += w;
  }
  buf[off] = 0;
  return off;
}
extern "C" int sivae_profile_read(double* out) {
  if (!out) return fail(-1, "null argument");
  cudaDeviceSynchronize();
  for (int i = 0; i < PC_COUNT * 3; ++i) out[i] = 0.0;
  for (size_t i = 0; i < g_prof.used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.recs[i].a, g_prof.recs[i].b) != cudaSuccess) continue;
    if (g_prof.recs[i].cls >= PC_COUNT) continue;
    out[g_prof.recs[i].cls * 3 + 0] += ms;
    out[g_prof.recs[i].cls * 3 + 1] += g_prof.recs[i].flops;
    out[g_prof.recs[i].cls * 3 + 2] += 1.0;
  }
  g_prof.used = 0;
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_last_batch(const sivae_engine* e) { return e ? e->cur_batch : -1; }
// CUDA-graph replays of a step do not run this library's host code: the caller records the batch size of the replayed step
extern "C" int sivae_set_last_batch(sivae_engine* e, int batch) {
  if (!e || batch < 1 || batch > e->cfg.max_batch) return fail(-5, "batch out of range");
  e->cur_batch = batch;
  return 0;
}
extern "C" int sivae_last_image(sivae_engine* e, int slot, float* out_nchw, void* stream) {
  if (!e || !e->ws || slot < 0 || slot > 3 || !out_nchw || e->cur_batch < 1) return fail(-1, "bad argument / no step run yet");
  const sivae_config& c = e->cfg;
  launch_nhwc_to_nchw(e->dp[slot].y, out_nchw, e->cur_batch, c.cdim, c.image_size, c.image_size, (cudaStream_t)stream);
  CHECK_CUDA_RET();
  return 0;
}

// ---- single-kernel entry points ------------------------------------------------------------------------------
// library-owned scratch for the transposed narrow filter of the single-kernel test entry points (the engine keeps its
// own copy in the workspace)
static float* lib_scratch(int slot, size_t floats) {
  static float* buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  static size_t cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (floats > cap[slot]) {
    if (buf[slot]) cudaFree(buf[slot]);
    if (cudaMalloc(&buf[slot], floats * sizeof(float)) != cudaSuccess) { buf[slot] = nullptr; cap[slot] = 0; return nullptr; }
    cap[slot] = floats;
  }
  return buf[slot];
}
static float* narrow_scratch(size_t floats) { return lib_scratch(0, floats); }
// AUTO backend of the single-kernel entry points: prepare whatever derived filter / scratch the engine's choice needs
// split (forward entry point only): where a tensor-core kernel serves the shape, hand it split32 operands like the engine's
// default mode does -- *x / *w_tc are redirected to library-owned split32 copies and *fmt is set to FMT_SPLIT_
static int auto_prepare(sivae_engine* tmp, const ConvShape& s, const float* filt, const float** wn, const float** wg, cudaStream_t st,
                        bool split = false, const float** x = nullptr, const float** w_tc = nullptr, int* fmt = nullptr) {
  *wn = *wg = nullptr;
  tmp->rs = lib_scratch(1, (size_t)conv_rowsep_scratch_floats(s));
  tmp->sk_bytes = conv_tc_supported_fwd(s) ? conv_tc_splitk_scratch_bytes(s) : 0;
  tmp->sk = tmp->sk_bytes ? (void*)lib_scratch(2, (tmp->sk_bytes + 3) / 4) : nullptr;
  if (!tmp->sk) tmp->sk_bytes = 0;
  const int wide = s.Cin > s.Cout ? s.Cin : s.Cout;
  auto split_x = [&]() -> int {
    float* xs = lib_scratch(5, (size_t)(s.pixels() * s.Cin));
    if (!xs) return fail(-3, "cudaMalloc of the split32 scratch failed");
    launch_split32(*x, xs, s.pixels() * s.Cin, st);
    *x = xs; *fmt = FMT_SPLIT_;
    return 0;
  };
  if (fwd_on_rowsep_in(tmp, s)) {
    float* buf = narrow_scratch((size_t)wide * 160);
    if (!buf) return fail(-3, "cudaMalloc of the narrow-filter scratch failed");
    launch_rowsep_filter_expand(filt, buf, s.Cout, s.Cin, st, !split);
    if (split) { launch_split32(buf, buf, (long long)s.Cout * 160, st); *fmt = FMT_SPLIT_; }     // x stays the fp32 image
    *wn = buf;
  } else if (fwd_on_rowsep_out(tmp, s)) {
    float* buf = narrow_scratch((size_t)wide * 160);
    if (!buf) return fail(-3, "cudaMalloc of the narrow-filter scratch failed");
    launch_rowsep_filter_gather(filt, buf, s.Cout, s.Cin, st, !split);
    if (split) { launch_split32(buf, buf, (long long)80 * s.Cin, st); TRY(split_x()); }
    *wg = buf;
  } else if (split && !fwd_on_narrow(tmp, s) && fwd_on_tc(tmp, s)) {
    float* ws = lib_scratch(3, (size_t)(s.Cout * s.ktot()));
    if (!ws) return fail(-3, "cudaMalloc of the split32 scratch failed");
    launch_split32(filt, ws, (long long)s.Cout * s.ktot(), st);
    *w_tc = ws;
    TRY(split_x());
  } else if (fwd_on_narrow(tmp, s)) {
    float* buf = narrow_scratch((size_t)s.Cout * s.Cin * s.k * s.k);
    if (!buf) return fail(-3, "cudaMalloc of the narrow-filter scratch failed");
    launch_narrow_transpose(filt, buf, s.Cout, s.k, s.Cin, st);
    *wn = buf;
  }
  return 0;
}
// compensated (3xTF32) conv of the single-kernel entry points: unrounded fp32 x and filter, library-owned split scratch
static int conv3x_entry(const ConvShape& s, const float* x, const float* filt, const float* bias, const float* addend, float* y,
                        cudaStream_t st) {
  if (!conv_tc_supported_fwd(s)) return fail(-7, "shape not supported by the tcgen05 conv kernel");
  sivae_engine tmp; tmp.tc = tmp.comp = true; tmp.fast = false;
  const long long wn = (long long)s.Cout * s.ktot();
  float* wh = lib_scratch(3, (size_t)wn);
  float* wl = lib_scratch(4, (size_t)wn);
  tmp.split[0] = lib_scratch(5, (size_t)(s.pixels() * s.Cin));
  tmp.split[1] = lib_scratch(6, (size_t)(s.pixels() * s.Cin));
  tmp.sk_bytes = conv_tc_splitk_scratch_bytes(s);
  tmp.sk = tmp.sk_bytes ? (void*)lib_scratch(2, (tmp.sk_bytes + 3) / 4) : nullptr;
  if (!wh || !wl || !tmp.split[0] || !tmp.split[1] || (tmp.sk_bytes && !tmp.sk)) return fail(-3, "cudaMalloc of the split scratch failed");
  launch_split_tf32(filt, wh, wl, wn, st);
  return conv_any(&tmp, s, x, filt, wh, bias, addend, y, st, nullptr, nullptr, nullptr, wl);
}
// backend: SIVAE_CONV_SIMT = generic exact fp32 kernel; SIVAE_CONV_TCGEN05 = tensor-core kernel or error -7;
// SIVAE_CONV_TC3X = compensated tensor-core conv (3 tf32 MMAs per product on split operands, fp32-class accuracy);
// SIVAE_CONV_AUTO = what the engine would pick for this shape (narrow CUDA-core kernel, tensor core, generic)
extern "C" int sivae_conv2d_fwd(const float* x, const float* w, const float* bias, const float* addend, float* y, int N, int H,
                                int W, int Cin, int Cout, int k, int backend, void* stream) {
  ConvShape s{N, H, W, Cin, Cout, k};
  cudaStream_t st = (cudaStream_t)stream;
  if (backend == SIVAE_CONV_TCGEN05) {
    if (!conv_tc_supported_fwd(s)) return fail(-7, "shape not supported by the tcgen05 conv kernel");
    size_t skb = conv_tc_splitk_scratch_bytes(s);
    void* sk = skb ? (void*)lib_scratch(2, (skb + 3) / 4) : nullptr;
    int r = launch_conv_fwd_tc(x, w, bias, addend, y, s, st, nullptr, sk, sk ? skb : 0);
    if (r) return fail(r, "tcgen05 conv launch failed");
  } else if (backend == SIVAE_CONV_AUTO || backend == SIVAE_CONV_TF32) {
    // AUTO: what the engine's default mode runs for a FORWARD conv of this shape (split32 operands on the tensor core);
    // TF32: the round-1 choice (kind::tf32 on the caller's values)
    sivae_engine tmp; tmp.tc = tmp.fast = true;
    const float *wn, *wg, *xin = x, *wtc = w;
    int fmt = FMT_TF32_;
    TRY(auto_prepare(&tmp, s, w, &wn, &wg, st, backend == SIVAE_CONV_AUTO, &xin, &wtc, &fmt));
    TRY(conv_any(&tmp, s, xin, w, wtc, bias, addend, y, st, nullptr, wn, wg, nullptr, fmt));
  } else if (backend == SIVAE_CONV_TC3X) {
    TRY(conv3x_entry(s, x, w, bias, addend, y, st));
  } else {
    launch_conv_fwd_simt(x, w, bias, addend, y, s, st);
  }
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_conv2d_dgrad(const float* dy, const float* w, const float* addend, float* dx, int N, int H, int W, int Cin,
                                  int Cout, int k, int backend, void* workspace, long long ws_bytes, void* stream) {
  size_t need = (size_t)Cout * Cin * k * k * sizeof(float);
  if (!workspace || (size_t)ws_bytes < need) return fail(-3, "workspace too small for the packed dgrad filter");
  cudaStream_t st = (cudaStream_t)stream;
  float* wd = (float*)workspace;
  ConvShape s{N, H, W, Cout, Cin, k};
  if (backend == SIVAE_CONV_AUTO && k != 5 && conv_tc_supported_fwd(s, FMT_BF16_) && (Cin & 3) == 0) {
    // what the engine's default mode runs for the dgrad of a residual-block conv: plain bf16 dy and filter, kind::f16
    void* dyh = lib_scratch(6, (size_t)(s.pixels() * s.Cin + 1) / 2);
    void* wdh = lib_scratch(7, (size_t)((long long)Cout * Cin * k * k + 1) / 2);
    if (!dyh || !wdh) return fail(-3, "cudaMalloc of the bf16 scratch failed");
    launch_to_bf16(dy, dyh, s.pixels() * s.Cin, st);
    launch_pack_dgrad_filter_bf16(w, wdh, Cout, Cin, k, st);
    sivae_engine tmp16; tmp16.tc = tmp16.fast = true;
    tmp16.sk_bytes = conv_tc_splitk_scratch_bytes(s);
    tmp16.sk = tmp16.sk_bytes ? (void*)lib_scratch(2, (tmp16.sk_bytes + 3) / 4) : nullptr;
    if (!tmp16.sk) tmp16.sk_bytes = 0;
    TRY(conv_any(&tmp16, s, (const float*)dyh, nullptr, (const float*)wdh, nullptr, addend, dx, st, nullptr, nullptr, nullptr, nullptr, FMT_BF16_));
    CHECK_CUDA_RET();
    return 0;
  }
  if (backend == SIVAE_CONV_TF32) backend = SIVAE_CONV_AUTO;      // TF32: the round-1 choice (kind::tf32 backward)
  sivae_engine tmp; tmp.tc = tmp.fast = (backend == SIVAE_CONV_AUTO);
  const bool on_tc = backend == SIVAE_CONV_TCGEN05 || (backend == SIVAE_CONV_AUTO && !fwd_on_narrow(&tmp, s) && fwd_on_tc(&tmp, s));
  launch_pack_dgrad_filter(w, wd, Cout, Cin, k, on_tc, st);
  if (backend == SIVAE_CONV_TC3X) {
    TRY(conv3x_entry(s, dy, wd, nullptr, addend, dx, st));
  } else if (backend == SIVAE_CONV_TCGEN05) {
    if (!conv_tc_supported_fwd(s)) return fail(-7, "shape not supported by the tcgen05 conv kernel");
    size_t skb = conv_tc_splitk_scratch_bytes(s);
    void* sk = skb ? (void*)lib_scratch(2, (skb + 3) / 4) : nullptr;
    int r = launch_conv_fwd_tc(dy, wd, nullptr, addend, dx, s, st, nullptr, sk, sk ? skb : 0);
    if (r) return fail(r, "tcgen05 conv launch failed");
  } else if (backend == SIVAE_CONV_AUTO) {
    const float *wn, *wg;
    TRY(auto_prepare(&tmp, s, wd, &wn, &wg, st));
    TRY(conv_any(&tmp, s, dy, wd, wd, nullptr, addend, dx, st, nullptr, wn, wg));
  } else {
    launch_conv_fwd_simt(dy, wd, nullptr, addend, dx, s, st);
  }
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_conv2d_wgrad(const float* x, const float* dy, float* dw, int N, int H, int W, int Cin, int Cout, int k,
                                  int accumulate, int backend, void* workspace, long long ws_bytes, void* stream) {
  ConvShape s{N, H, W, Cin, Cout, k};
  cudaStream_t st = (cudaStream_t)stream;
  const bool acc = accumulate != 0;
  if (backend == SIVAE_CONV_AUTO && k != 5 && conv_tc_supported_wgrad16(s)) {
    // the default mode's wgrad of a residual-block conv: plain bf16 x and dy, kind::f16
    if ((size_t)ws_bytes < conv_wgrad_tc16_scratch_bytes(s)) return fail(-3, "workspace too small");
    void* xh = lib_scratch(6, (size_t)(s.pixels() * s.Cin + 1) / 2);
    void* dyh = lib_scratch(7, (size_t)(s.pixels() * s.Cout + 1) / 2);
    if (!xh || !dyh) return fail(-3, "cudaMalloc of the bf16 scratch failed");
    launch_to_bf16(x, xh, s.pixels() * s.Cin, st);
    launch_to_bf16(dy, dyh, s.pixels() * s.Cout, st);
    int r = launch_conv_wgrad_tc16(xh, dyh, dw, s, acc, workspace, (size_t)ws_bytes, st);
    if (r) return fail(r, "tcgen05 bf16 wgrad launch failed");
    CHECK_CUDA_RET();
    return 0;
  }
  if (backend == SIVAE_CONV_TF32) backend = SIVAE_CONV_AUTO;      // TF32: the round-1 choice (kind::tf32 backward)
  const bool rs_stem = backend == SIVAE_CONV_AUTO && Cin <= 3 && conv_rowsep_wgrad_supported(H, W, Cin, Cout, k);
  const bool rs_pred = backend == SIVAE_CONV_AUTO && Cout <= 3 && conv_rowsep_wgrad_supported(H, W, Cout, Cin, k);
  if (rs_stem || rs_pred) {
    const int wide = rs_stem ? Cout : Cin, c = rs_stem ? Cin : Cout;
    if ((size_t)ws_bytes < conv_rowsep_wgrad_scratch_bytes(N, H, W, wide)) return fail(-3, "workspace too small");
    float* xe = lib_scratch(1, (size_t)conv_rowsep_scratch_floats(s));
    if (!xe) return fail(-3, "cudaMalloc of the row-expansion scratch failed");
    int r = launch_conv_rowsep_wgrad(rs_stem ? x : dy, rs_stem ? dy : x, dw, N, H, W, c, wide, rs_stem ? 1 : 0, acc, xe, workspace,
                                     (size_t)ws_bytes, st);
    if (r) return fail(r, "row-separable tcgen05 wgrad launch failed");
  } else if (backend == SIVAE_CONV_AUTO && Cin <= 4 && conv_narrow_corr_supported(Cout, Cin, k)) {
    if ((size_t)ws_bytes < conv_narrow_corr_scratch_bytes(Cout, k)) return fail(-3, "workspace too small");
    launch_conv_narrow_corr(x, dy, dw, N, H, W, Cin, Cout, k, 1, acc, workspace, (size_t)ws_bytes, st);
  } else if (backend == SIVAE_CONV_AUTO && Cout <= 4 && conv_narrow_corr_supported(Cin, Cout, k)) {
    if ((size_t)ws_bytes < conv_narrow_corr_scratch_bytes(Cin, k)) return fail(-3, "workspace too small");
    launch_conv_narrow_corr(dy, x, dw, N, H, W, Cout, Cin, k, 0, acc, workspace, (size_t)ws_bytes, st);
  } else if (backend == SIVAE_CONV_TC3X) {
    if (!conv_tc_supported_wgrad(s)) return fail(-7, "shape not supported by the tcgen05 wgrad kernel");
    if ((size_t)ws_bytes < conv_wgrad_tc_scratch_bytes(s)) return fail(-3, "workspace too small");
    float* sp[4];
    for (int i = 0; i < 4; ++i) {
      sp[i] = lib_scratch(4 + i, (size_t)(s.pixels() * (i < 2 ? s.Cin : s.Cout)));
      if (!sp[i]) return fail(-3, "cudaMalloc of the split scratch failed");
    }
    launch_split_tf32(x, sp[0], sp[1], s.pixels() * s.Cin, st);
    launch_split_tf32(dy, sp[2], sp[3], s.pixels() * s.Cout, st);
    if (!acc) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)s.Cout * s.ktot(), st);
    int r = launch_conv_wgrad_tc(sp[1], sp[2], dw, s, true, workspace, (size_t)ws_bytes, st);
    if (!r) r = launch_conv_wgrad_tc(sp[0], sp[3], dw, s, true, workspace, (size_t)ws_bytes, st);
    if (!r) r = launch_conv_wgrad_tc(sp[0], sp[2], dw, s, true, workspace, (size_t)ws_bytes, st);
    if (r) return fail(r, "tcgen05 wgrad launch failed");
  } else if (backend == SIVAE_CONV_TCGEN05 || (backend == SIVAE_CONV_AUTO && conv_tc_supported_wgrad(s))) {
    if (!conv_tc_supported_wgrad(s)) return fail(-7, "shape not supported by the tcgen05 wgrad kernel");
    if ((size_t)ws_bytes < conv_wgrad_tc_scratch_bytes(s)) return fail(-3, "workspace too small");
    int r = launch_conv_wgrad_tc(x, dy, dw, s, acc, workspace, (size_t)ws_bytes, st);
    if (r) return fail(r, "tcgen05 wgrad launch failed");
  } else {
    if ((size_t)ws_bytes < conv_wgrad_simt_scratch_bytes(s)) return fail(-3, "workspace too small");
    launch_conv_wgrad_simt(x, dy, dw, s, acc, workspace, (size_t)ws_bytes, st);
  }
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_bn_act_fwd(const float* t, const float* identity, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, long long* nbt, float* mean_invstd, float* out, int N, int H, int W, int C,
                                int mode, int train, void* workspace, long long ws_bytes, void* stream) {
  if (C % 4 != 0) return fail(-2, "C must be a multiple of 4");
  long long rows = (long long)N * H * W;
  if ((size_t)ws_bytes < bn_scratch_bytes(rows, C)) return fail(-3, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (train) launch_bn_stats(t, rows, C, mean_invstd, running_mean, running_var, nbt, workspace, (size_t)ws_bytes, st);
  else launch_bn_eval_stats(running_mean, running_var, C, mean_invstd, st);
  launch_bn_act_fwd(t, identity, mean_invstd, gamma, beta, out, N, H, W, C, mode, false, st);
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_bn_act_bwd(const float* dout, const float* t, const float* identity, const float* gamma, const float* beta,
                                const float* mean_invstd, float* dt, float* g, float* dgamma, float* dbeta, int accumulate, int N,
                                int H, int W, int C, int mode, void* workspace, long long ws_bytes, void* stream) {
  if (C % 4 != 0) return fail(-2, "C must be a multiple of 4");
  if ((size_t)ws_bytes < bn_scratch_bytes((long long)N * H * W, C)) return fail(-3, "workspace too small");
  launch_bn_act_bwd(dout, t, identity, mean_invstd, gamma, beta, dt, g, dgamma, dbeta, accumulate != 0, N, H, W, C, mode, false,
                    workspace, (size_t)ws_bytes, (cudaStream_t)stream);
  CHECK_CUDA_RET();
  return 0;
}
// variants with the forward's sign bytes (N*H*W*C/4 bytes): the backward then never reads `identity`
extern "C" int sivae_bn_act_fwd_m(const float* t, const float* identity, const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, long long* nbt, float* mean_invstd, float* out, int N, int H, int W, int C,
                                  int mode, int train, void* workspace, long long ws_bytes, unsigned char* sign_mask, void* stream) {
  if (C % 4 != 0) return fail(-2, "C must be a multiple of 4");
  long long rows = (long long)N * H * W;
  if ((size_t)ws_bytes < bn_scratch_bytes(rows, C)) return fail(-3, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (train) launch_bn_stats(t, rows, C, mean_invstd, running_mean, running_var, nbt, workspace, (size_t)ws_bytes, st);
  else launch_bn_eval_stats(running_mean, running_var, C, mean_invstd, st);
  launch_bn_act_fwd(t, identity, mean_invstd, gamma, beta, out, N, H, W, C, mode, false, st, sign_mask);
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_bn_act_bwd_m(const float* dout, const float* t, const float* identity, const float* gamma, const float* beta,
                                  const float* mean_invstd, float* dt, float* g, float* dgamma, float* dbeta, int accumulate, int N,
                                  int H, int W, int C, int mode, void* workspace, long long ws_bytes, const unsigned char* sign_mask,
                                  void* stream) {
  if (C % 4 != 0) return fail(-2, "C must be a multiple of 4");
  if ((size_t)ws_bytes < bn_scratch_bytes((long long)N * H * W, C)) return fail(-3, "workspace too small");
  launch_bn_act_bwd(dout, t, identity, mean_invstd, gamma, beta, dt, g, dgamma, dbeta, accumulate != 0, N, H, W, C, mode, false,
                    workspace, (size_t)ws_bytes, (cudaStream_t)stream, sign_mask);
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_mse3(const float* real, const float* rec, const float* rec_rec, const float* fake, const float* rec_fake,
                          float* out, int B, long long per_sample, void* workspace, long long ws_bytes, void* stream) {
  if ((size_t)ws_bytes < mse3_scratch_bytes(B, per_sample)) return fail(-3, "workspace too small");
  launch_mse3(real, rec, rec_rec, fake, rec_fake, out, B, per_sample, workspace, (size_t)ws_bytes, (cudaStream_t)stream);
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_kl_reparam(const float* mu_logvar, const float* eps, float* z, float* kl, int B, int zdim, void* stream) {
  launch_kl_reparam(mu_logvar, eps, z, kl, B, zdim, (cudaStream_t)stream);
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float grad_scale,
                               long long step, void* stream) {
  static long long* sdev = nullptr;
  static float* cdev = nullptr;
  if (!sdev) { cudaMalloc(&sdev, 16); cudaMalloc(&cdev, 16); }     // 32 bytes of library-owned scratch for this test entry point
  long long prev = step - 1;
  cudaMemcpyAsync(sdev, &prev, sizeof(long long), cudaMemcpyHostToDevice, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  launch_adam(p, g, m, v, n, lr, grad_scale, 0.9f, 0.999f, 1e-8f, sdev, cdev, (cudaStream_t)stream);
  CHECK_CUDA_RET();
  return 0;
}

// nn.Linear forward / input gradient as stand-alone calls (unit parity of the fc kernels)
extern "C" int sivae_linear_fwd(const float* x, const float* w, const float* b, float* y, int B, int F, int O, int relu, void* stream) {
  if (!x || !w || !y || B < 1 || F < 1 || O < 1) return fail(-1, "bad argument");
  launch_linear_fwd(x, w, b, y, B, F, O, relu != 0, (cudaStream_t)stream);
  CHECK_CUDA_RET();
  return 0;
}
extern "C" int sivae_linear_dgrad(const float* dy, const float* w, float* dx, int B, int F, int O, void* workspace, long long ws_bytes,
                                  void* stream) {
  if (!dy || !w || !dx || !workspace || B < 1 || F < 1 || O < 1) return fail(-1, "bad argument");
  if ((size_t)ws_bytes < linear_dgrad_scratch_bytes(B, F, O)) return fail(-3, "workspace too small");
  launch_linear_dgrad(dy, w, dx, B, F, O, workspace, (cudaStream_t)stream);
  CHECK_CUDA_RET();
  return 0;
}
extern "C" long long sivae_linear_dgrad_workspace_bytes(int B, int F, int O) { return (long long)linear_dgrad_scratch_bytes(B, F, O); }

// ---- image batch assembly (image.cu) -----------------------------------------------------------------------------
extern "C" int sivae_resample_coeffs(int in_size, int out_size, int* ksize, int* bounds, int* kk, long long kk_capacity) {
  if (in_size < 1 || out_size < 1 || !ksize || !bounds || !kk) return fail(-1, "bad argument");
  std::vector<int> b, k;
  const int ks = resample_coeffs(in_size, out_size, b, k);
  *ksize = ks;
  if ((long long)k.size() > kk_capacity) return fail(-3, "coefficient buffer too small");
  memcpy(bounds, b.data(), sizeof(int) * b.size());
  memcpy(kk, k.data(), sizeof(int) * k.size());
  return 0;
}
extern "C" long long sivae_image_plan_bytes(int in_h, int in_w, int out_h, int out_w) {
  if (in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1) return -1;
  return (long long)image_plan_bytes(in_h, in_w, out_h, out_w);
}
extern "C" int sivae_image_plan_init(int in_h, int in_w, int out_h, int out_w, void* plan_dev, long long plan_bytes, void* stream) {
  if (in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1 || !plan_dev) return fail(-1, "bad argument");
  if ((size_t)plan_bytes < image_plan_bytes(in_h, in_w, out_h, out_w)) return fail(-3, "plan buffer too small");
  int r = image_plan_init(in_h, in_w, out_h, out_w, plan_dev, (cudaStream_t)stream);
  if (r) return fail(r, "coefficient upload failed");
  return 0;
}
extern "C" int sivae_image_batch_u8(const unsigned char* src_hwc, const unsigned char* mirror, int batch, int in_h, int in_w,
                                    int channels, int out_h, int out_w, const void* plan_dev, float* out_nchw, void* stream) {
  if (!src_hwc || !plan_dev || !out_nchw) return fail(-1, "null argument");
  int r = launch_image_batch(src_hwc, mirror, batch, in_h, in_w, channels, out_h, out_w, plan_dev, out_nchw, (cudaStream_t)stream);
  if (r == -2) return fail(-2, "bad image batch arguments (channels must be 1 or 3, batch <= 65535, source 4-byte aligned)");
  if (r == -7) return fail(-7, "down-scaling factor too large for the shared-memory staging of the resize kernel");
  if (r) return fail(r, cudaGetErrorString((cudaError_t)r));
  return 0;
}
// load_image in full (dataset.py:12-47): the resize reads the win_h x win_w window at win_xy[b] = (x, y) of each [src_h, src_w]
// source image (ImageOps.crop, :32-44; win_xy == NULL: the whole image) and writes EITHER out_nchw (float32, = ToTensor of the
// final image) OR out_u8_hwc (the 8-bit image itself: first stage of the two-stage resize, :29-30).  plan: (win_h, win_w) ->
// (out_h, out_w).
extern "C" int sivae_image_batch_u8_ex(const unsigned char* src_hwc, const unsigned char* mirror, const int* win_xy, int batch,
                                       int src_h, int src_w, int channels, int win_h, int win_w, int out_h, int out_w,
                                       const void* plan_dev, float* out_nchw, unsigned char* out_u8_hwc, void* stream) {
  if (!src_hwc || !plan_dev || (!out_nchw && !out_u8_hwc)) return fail(-1, "null argument");
  int r = launch_image_batch(src_hwc, mirror, batch, win_h, win_w, channels, out_h, out_w, plan_dev, out_nchw, (cudaStream_t)stream,
                             src_h, src_w, win_xy, out_u8_hwc);
  if (r == -2) return fail(-2, "bad image batch arguments (channels 1 or 3, batch <= 65535, source 4-byte aligned, window inside the "
                               "source, exactly one output)");
  if (r == -7) return fail(-7, "down-scaling factor too large for the shared-memory staging of the resize kernel");
  if (r) return fail(r, cudaGetErrorString((cudaError_t)r));
  return 0;
}

// ---- JPEG decode in front of the batch-assembly kernel (jpeg.cu): nvJPEG, opt-in ------------------------------------
extern "C" int sivae_jpeg_info(const unsigned char* data, long long length, int* height, int* width, int* components) {
  if (!data || length <= 0 || !height || !width || !components) return fail(-1, "null argument");
  std::string msg;
  int r = jpeg_info(data, length, height, width, components, &msg);
  if (r) return fail(r, msg);
  return 0;
}
extern "C" int sivae_jpeg_decode_batch(const unsigned char* const* data, const long long* lengths, int batch, int height, int width,
                                       unsigned char* out_hwc, void* stream) {
  if (!data || !lengths || !out_hwc || batch < 1 || height < 1 || width < 1) return fail(-1, "null / empty argument");
  std::string msg;
  int r = jpeg_decode_batch(data, lengths, batch, height, width, out_hwc, (cudaStream_t)stream, &msg);
  if (r) return fail(r, msg);
  return 0;
}
