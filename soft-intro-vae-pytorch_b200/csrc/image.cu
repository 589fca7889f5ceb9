// image.cu -- batch assembly of decoded 8-bit images on the GPU (the loader that feeds the step):
//   ImageOps.mirror  ->  Image.resize((S, S), Image.BICUBIC)  ->  transforms.ToTensor()
// i.e. soft_intro_vae/dataset.py:26-27, :46 and :66-68,75 as the image configs call load_image
// (train_soft_intro_vae.py:388-392, 400-404, 415-417: input_height=None, crop_height=None, output_height=S).
// The arithmetic is Pillow's (src/libImaging/Resample.c; not part of the reference tree): fixed-point (22 fractional bits)
// bicubic coefficients computed in doubles on the host, a horizontal pass that rounds and saturates to 8 bits, then a
// vertical pass that does the same; ToTensor is uint8 / 255 in float32.  All of it integer / correctly-rounded, so the
// result is bit-exact (tests/test_gpu_image.py against oracle/image_oracle.py, which is pinned to Pillow itself).
//
// One kernel per batch: a CTA owns a TW x TH tile of one output image, stages the source window it needs (coalesced
// aligned 32-bit loads) in shared memory, runs the horizontal pass into a shared 8-bit intermediate (never written to
// HBM) and the vertical pass straight into the NCHW float output.  HBM traffic = source bytes (x halo overlap, absorbed
// by L2) + output floats.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "kernels.h"

namespace sivae {

constexpr int IMG_PRECISION_BITS = 32 - 8 - 2;     // Resample.c PRECISION_BITS

// Resample.c bicubic_filter (Keys, a = -0.5)
static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Resample.c precompute_coeffs (box = the whole axis) + normalize_coeffs_8bpc.  bounds: [out][2] = (first source index,
// tap count); kk: [out][ksize] fixed-point taps, unused tail zero.  Same operation order in IEEE doubles as Pillow.
int resample_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = (int)ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> w((size_t)ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      double k = w[x];
      if (ww != 0.0) k /= ww;
      kk[(size_t)xx * ksize + x] = k < 0 ? (int)(-0.5 + k * (1 << IMG_PRECISION_BITS)) : (int)(0.5 + k * (1 << IMG_PRECISION_BITS));
    }
    bounds[(size_t)xx * 2] = xmin;
    bounds[(size_t)xx * 2 + 1] = xmax;
  }
  return ksize;
}

namespace {

struct ImgArgs {
  const unsigned char* src;      // [B][src_h][src_w][CH] decoded pixels
  const unsigned char* src_end;  // one past the last source byte
  const unsigned char* mirror;   // [B] flags (nullable): 1 = ImageOps.mirror of the WINDOW before the resize
  const int* win_xy;             // [B][2] (nullable = 0,0): origin (x, y) of the Hin x Win window inside each source image --
                                 // ImageOps.crop (dataset.py:32-44); the resize sees only the window, like Pillow's cropped copy
  float* dst;                    // [B][CH][Hout][Wout] float32 = ToTensor, or (when dst_u8 != null, dst == null):
  unsigned char* dst_u8;         // [B][Hout][Wout][CH] the 8-bit image itself (first stage of load_image's two-stage resize, :29-30)
  const int *bx, *kx, *by, *ky;  // device coefficient tables (plan)
  int B, Hin, Win, Hout, Wout, ksx, ksy, TW, TH, pitch_in, max_rows;
  int src_pitch;                 // bytes per source row (src_w * CH)
  long long img_stride;          // bytes per source image
};

__device__ __forceinline__ int clip8(int acc) {     // Resample.c clip8: saturating lookup of acc >> PRECISION_BITS
  int v = acc >> IMG_PRECISION_BITS;
  return min(max(v, 0), 255);
}

constexpr int IMG_RG = 4;        // source rows one thread carries through the horizontal pass (taps loaded once for all)

// MIR: the tile's image is mirrored (block-uniform; ImageOps.mirror folded into the horizontal pass' tap direction)
template <int CH, bool MIR>
__device__ __forceinline__ void image_tile(const ImgArgs& a, unsigned char* img_smem) {
  // tables: taps padded to a multiple of 4 per output (zero-filled) so the tap loops run in unrolled groups of 4
  const int ksx4 = (a.ksx + 3) & ~3, ksy4 = (a.ksy + 3) & ~3;
  int* kx_s = reinterpret_cast<int*>(img_smem);          // [TW][ksx4]
  int* ky_s = kx_s + a.TW * ksx4;                         // [TH][ksy4]
  int* bx_s = ky_s + a.TH * ksy4;                         // [TW][2]
  int* by_s = bx_s + 2 * a.TW;                            // [TH][2]
  unsigned char* in_s = reinterpret_cast<unsigned char*>(by_s + 2 * a.TH) + 16;   // [max_rows][pitch_in] source window (+16 B slack either side:
  unsigned char* mid_s = in_s + (size_t)a.max_rows * a.pitch_in + 16;              //  zero-weight padded taps may read just outside a row)
                                                                                   // [max_rows + 3][TW*CH] after the horizontal pass
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx0 = blockIdx.x * a.TW, ty0 = blockIdx.y * a.TH, b = blockIdx.z;
  const int tw = min(a.TW, a.Wout - tx0), th = min(a.TH, a.Hout - ty0);
  for (int i = tid; i < tw * ksx4; i += 256) {
    const int x = i / ksx4, k = i - x * ksx4;
    kx_s[i] = k < a.ksx ? __ldg(a.kx + (size_t)(tx0 + x) * a.ksx + k) : 0;
  }
  for (int i = tid; i < th * ksy4; i += 256) {
    const int y = i / ksy4, k = i - y * ksy4;
    ky_s[i] = k < a.ksy ? __ldg(a.ky + (size_t)(ty0 + y) * a.ksy + k) : 0;
  }
  for (int i = tid; i < 2 * tw; i += 256) bx_s[i] = __ldg(a.bx + 2 * tx0 + i);
  for (int i = tid; i < 2 * th; i += 256) by_s[i] = __ldg(a.by + 2 * ty0 + i);
  __syncthreads();
  // source window of the tile (first index and first+count are both non-decreasing in the output coordinate)
  const int x0 = bx_s[0], x1 = bx_s[2 * (tw - 1)] + bx_s[2 * (tw - 1) + 1];
  const int y0 = by_s[0], y1 = by_s[2 * (th - 1)] + by_s[2 * (th - 1) + 1];
  const int span = x1 - x0, R = y1 - y0;
  const int sx0 = MIR ? a.Win - x1 : x0;                  // mirrored window [x0,x1) = source [Win-x1, Win-x0) reversed
  const int wx = a.win_xy ? a.win_xy[2 * b] : 0, wy = a.win_xy ? a.win_xy[2 * b + 1] : 0;
  const int rowb = a.src_pitch;
  const unsigned char* img = a.src + (size_t)b * a.img_stride + (size_t)wy * rowb + (size_t)wx * CH;   // origin of the crop window
  const int nbytes = span * CH;
  const unsigned char* win = img + (size_t)y0 * rowb + (size_t)sx0 * CH;           // first byte of the tile window's first row
  const int g0 = (int)(reinterpret_cast<uintptr_t>(win) & 3);                      // its misalignment; row r: (g0 + r*rowb) & 3

  // ---- stage: one warp per source row, aligned 32-bit words (the row start is rounded down to a word), asynchronous
  //      global->shared copies (LDGSTS): every row of the window is in flight at once, no register round trip --------------
  for (int r = warp; r < R; r += 8) {
    const int g = (g0 + r * (rowb & 3)) & 3;
    const unsigned char* ap = win + (size_t)r * rowb - g;
    const int nwords = (g + nbytes + 3) >> 2;
    unsigned char* drow = in_s + (size_t)r * a.pitch_in;
    if (ap + 4 * (size_t)nwords <= a.src_end) {
      const unsigned dsh = (unsigned)__cvta_generic_to_shared(drow);
      for (int w = lane; w < nwords; w += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dsh + 4u * (unsigned)w), "l"(ap + 4 * (size_t)w) : "memory");
    } else {                                              // last row of the last image: never read past the tensor
      for (int w = lane; w < nwords; w += 32) {
        unsigned v = 0;
        for (int j = 0; j < 4; ++j)
          if (ap + 4 * w + j < a.src_end) v |= (unsigned)ap[4 * w + j] << (8 * j);
        reinterpret_cast<unsigned*>(drow)[w] = v;
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // ---- horizontal pass (ImagingResampleHorizontal_8bpc): lane = output column, warp = groups of IMG_RG source rows ----
  const int ngroups = (R + IMG_RG - 1) / IMG_RG;
  const int mid_pitch = a.TW * CH;
  for (int x = lane; x < tw; x += 32) {
    const int xmin = bx_s[2 * x], ngrp4 = (bx_s[2 * x + 1] + 3) >> 2;
    const int* kp = kx_s + x * ksx4;
    const int p = xmin - x0;                              // first tap, in window coordinates of the (mirrored) image
    const int off0 = (MIR ? span - 1 - p : p) * CH;
    for (int rg = warp; rg < ngroups; rg += 8) {
      int acc[IMG_RG][CH];
      const unsigned char* rowp[IMG_RG];
#pragma unroll
      for (int r = 0; r < IMG_RG; ++r) {
        const int row = min(rg * IMG_RG + r, R - 1);      // rows past the window repeat the last one, result discarded
        rowp[r] = in_s + row * a.pitch_in + ((g0 + row * (rowb & 3)) & 3) + off0;
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[r][c] = 1 << (IMG_PRECISION_BITS - 1);
      }
      for (int k4 = 0; k4 < ngrp4; ++k4) {
        const int4 cf = *reinterpret_cast<const int4*>(kp + 4 * k4);
        const int coef[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
        for (int r = 0; r < IMG_RG; ++r) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[r][c] += (int)rowp[r][(MIR ? -k : k) * CH + c] * coef[k];
          rowp[r] += MIR ? -4 * CH : 4 * CH;
        }
      }
#pragma unroll
      for (int r = 0; r < IMG_RG; ++r) {
        const int row = rg * IMG_RG + r;
        if (row < R) {
#pragma unroll
          for (int c = 0; c < CH; ++c) mid_s[row * mid_pitch + x * CH + c] = (unsigned char)clip8(acc[r][c]);
        }
      }
    }
  }
  __syncthreads();

  // ---- vertical pass (ImagingResampleVertical_8bpc) + ToTensor: lane = output column, warp = output rows ----------
  for (int x = lane; x < tw; x += 32) {
    for (int yy = warp; yy < th; yy += 8) {
      const int ymin = by_s[2 * yy] - y0, ngrp4 = (by_s[2 * yy + 1] + 3) >> 2;
      const int* kp = ky_s + yy * ksy4;
      const unsigned char* mp = mid_s + ymin * mid_pitch + x * CH;
      int acc[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = 1 << (IMG_PRECISION_BITS - 1);
      for (int k4 = 0; k4 < ngrp4; ++k4) {
        const int4 cf = *reinterpret_cast<const int4*>(kp + 4 * k4);
        const int coef[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int c = 0; c < CH; ++c) acc[c] += (int)mp[k * mid_pitch + c] * coef[k];
        mp += 4 * mid_pitch;
      }
      if (a.dst_u8) {                                     // the resized 8-bit image, HWC like the source (block-uniform branch)
#pragma unroll
        for (int c = 0; c < CH; ++c)
          a.dst_u8[(((size_t)b * a.Hout + ty0 + yy) * a.Wout + tx0 + x) * CH + c] = (unsigned char)clip8(acc[c]);
      } else {
#pragma unroll
        for (int c = 0; c < CH; ++c)
          a.dst[(((size_t)b * CH + c) * a.Hout + ty0 + yy) * a.Wout + tx0 + x] = __fdiv_rn((float)clip8(acc[c]), 255.f);   // ToTensor: .div(255)
      }
    }
  }
}

template <int CH>
__global__ void __launch_bounds__(256) k_image_batch(ImgArgs a) {
  extern __shared__ __align__(16) unsigned char img_smem[];
  if (a.mirror != nullptr && a.mirror[blockIdx.z] != 0) image_tile<CH, true>(a, img_smem);
  else image_tile<CH, false>(a, img_smem);
}

struct ImgPlan {
  int ksx = 0, ksy = 0;
  std::vector<int> bx, kx, by, ky;
};
struct TileCfg { int TW = 0, TH = 0, pitch_in = 0, max_rows = 0; size_t smem = 0; };

std::mutex g_plan_mu;
std::map<std::tuple<int, int, int, int>, ImgPlan> g_plans;

const ImgPlan& get_plan(int in_h, int in_w, int out_h, int out_w) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto key = std::make_tuple(in_h, in_w, out_h, out_w);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) return it->second;
  ImgPlan p;
  p.ksx = resample_coeffs(in_w, out_w, p.bx, p.kx);
  p.ksy = resample_coeffs(in_h, out_h, p.by, p.ky);
  return g_plans.emplace(key, std::move(p)).first->second;
}

// largest tile whose staging fits the shared-memory budget (two CTAs per SM)
bool pick_tile(const ImgPlan& p, int out_h, int out_w, int ch, TileCfg* out) {
  static const int cand[][2] = {{32, 16}, {32, 8}, {16, 8}, {16, 4}, {8, 4}, {8, 2}, {4, 2}, {4, 1}};
  const size_t budget = 100 * 1024;
  for (auto& c : cand) {
    const int TW = c[0], TH = c[1];
    int max_span = 0, max_rows = 0;
    for (int x0 = 0; x0 < out_w; x0 += TW) {
      const int xl = (x0 + TW < out_w ? x0 + TW : out_w) - 1;
      const int s = p.bx[2 * xl] + p.bx[2 * xl + 1] - p.bx[2 * x0];
      if (s > max_span) max_span = s;
    }
    for (int y0 = 0; y0 < out_h; y0 += TH) {
      const int yl = (y0 + TH < out_h ? y0 + TH : out_h) - 1;
      const int r = p.by[2 * yl] + p.by[2 * yl + 1] - p.by[2 * y0];
      if (r > max_rows) max_rows = r;
    }
    const int pitch = ((max_span * ch + 3 + 3) / 4) * 4;
    const int ksx4 = (p.ksx + 3) & ~3, ksy4 = (p.ksy + 3) & ~3;
    const size_t smem = sizeof(int) * ((size_t)TW * ksx4 + (size_t)TH * ksy4 + 2 * TW + 2 * TH) + 16 +
                        (size_t)max_rows * pitch + 16 + (size_t)(max_rows + 3) * TW * ch;
    if (smem <= budget) {
      out->TW = TW; out->TH = TH; out->pitch_in = pitch; out->max_rows = max_rows; out->smem = smem;
      return true;
    }
  }
  return false;
}

}  // namespace

size_t image_plan_bytes(int in_h, int in_w, int out_h, int out_w) {
  const ImgPlan& p = get_plan(in_h, in_w, out_h, out_w);
  return sizeof(int) * (p.bx.size() + p.kx.size() + p.by.size() + p.ky.size());
}

int image_plan_init(int in_h, int in_w, int out_h, int out_w, void* plan_dev, cudaStream_t st) {
  const ImgPlan& p = get_plan(in_h, in_w, out_h, out_w);          // lives in the process-wide cache: stable source for the async copies
  int* d = static_cast<int*>(plan_dev);
  cudaMemcpyAsync(d, p.bx.data(), sizeof(int) * p.bx.size(), cudaMemcpyHostToDevice, st); d += p.bx.size();
  cudaMemcpyAsync(d, p.kx.data(), sizeof(int) * p.kx.size(), cudaMemcpyHostToDevice, st); d += p.kx.size();
  cudaMemcpyAsync(d, p.by.data(), sizeof(int) * p.by.size(), cudaMemcpyHostToDevice, st); d += p.by.size();
  cudaMemcpyAsync(d, p.ky.data(), sizeof(int) * p.ky.size(), cudaMemcpyHostToDevice, st);
  return (int)cudaGetLastError();
}

// returns 0, a cudaError_t (> 0), -2 (bad argument) or -7 (down-scaling factor too large for the shared-memory staging)
// src: [B][src_h][src_w][ch]; the resize reads the in_h x in_w window at win_xy[b] (nullable: the whole image, then
// src_h == in_h, src_w == in_w); exactly one of dst (float NCHW) / dst_u8 (8-bit NHWC) is written.  The plan is that of
// (in_h, in_w) -> (out_h, out_w).  The caller guarantees that every window lies inside its image.
int launch_image_batch(const unsigned char* src, const unsigned char* mirror, int B, int in_h, int in_w, int ch, int out_h,
                       int out_w, const void* plan_dev, float* dst, cudaStream_t st, int src_h, int src_w, const int* win_xy,
                       unsigned char* dst_u8) {
  if (src_h <= 0) src_h = in_h;
  if (src_w <= 0) src_w = in_w;
  if (B < 1 || in_h < 1 || in_w < 1 || out_h < 1 || out_w < 1 || (ch != 1 && ch != 3)) return -2;
  if (B > 65535 || (reinterpret_cast<uintptr_t>(src) & 3) != 0) return -2;
  if (in_h > src_h || in_w > src_w || (dst == nullptr) == (dst_u8 == nullptr)) return -2;
  if (!win_xy && (in_h != src_h || in_w != src_w)) return -2;
  const ImgPlan& p = get_plan(in_h, in_w, out_h, out_w);
  TileCfg t;
  if (!pick_tile(p, out_h, out_w, ch, &t)) return -7;
  ImgArgs a;
  a.src = src; a.src_end = src + (size_t)B * src_h * src_w * ch; a.mirror = mirror; a.dst = dst; a.dst_u8 = dst_u8;
  a.win_xy = win_xy; a.src_pitch = src_w * ch; a.img_stride = (long long)src_h * src_w * ch;
  const int* d = static_cast<const int*>(plan_dev);
  a.bx = d; d += p.bx.size();
  a.kx = d; d += p.kx.size();
  a.by = d; d += p.by.size();
  a.ky = d;
  a.B = B; a.Hin = in_h; a.Win = in_w; a.Hout = out_h; a.Wout = out_w; a.ksx = p.ksx; a.ksy = p.ksy;
  a.TW = t.TW; a.TH = t.TH; a.pitch_in = t.pitch_in; a.max_rows = t.max_rows;
  dim3 grid((out_w + t.TW - 1) / t.TW, (out_h + t.TH - 1) / t.TH, B);
  g_launches += 1;
  if (ch == 3) {
    static size_t attr3 = 0;
    if (t.smem > 48 * 1024 && t.smem > attr3) {
      cudaFuncSetAttribute(k_image_batch<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem);
      attr3 = t.smem;
    }
    k_image_batch<3><<<grid, 256, t.smem, st>>>(a);
  } else {
    static size_t attr1 = 0;
    if (t.smem > 48 * 1024 && t.smem > attr1) {
      cudaFuncSetAttribute(k_image_batch<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t.smem);
      attr1 = t.smem;
    }
    k_image_batch<1><<<grid, 256, t.smem, st>>>(a);
  }
  return (int)cudaGetLastError();
}

}  // namespace sivae
