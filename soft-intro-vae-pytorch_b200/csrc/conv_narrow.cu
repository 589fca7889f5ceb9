// fp32 CUDA-core kernels for the two image-facing 5x5 convolutions whose narrow side (cdim = 1..4 channels) does not
// fill a tensor-core tile: the encoder stem (cdim -> c0, reference :89) and the decoder's `predict` (c0 -> cdim, :159).
//   k_narrow_in_fwd   y[p][co] = sum_{tap,a} x[p+tap][a] * w[tap][a][co]          stem forward, predict dgrad
//   k_narrow_corr     G[tap][a][c] = sum_p nar[p][a] * wide[p+tap][c]             stem wgrad, predict wgrad
// Since the row-separable tensor-core form exists (conv_tc.cu: launch_conv_rowsep_in/_out/_wgrad, 0.29 / 0.29 / 0.4 ms
// against 0.72 / 0.92 / 0.94 ms here at 32x256x256) these kernels are the FALLBACK for shapes it does not take: image
// sizes that are not multiples of 16 x 8, cdim = 4, or a tensor-core-less build of the engine's "fast" mode.
// Both are FMA-bound direct convolutions with the halo tile staged in shared memory; exact fp32 (no tf32 rounding).
//   forward: block = 8 x 32 output pixels, lane = (output-channel quad, pixel parity), filter pre-transposed to
//            [tap][a][Cout] so a quad is one conflict-free LDS.128 reused over 8 pixels (96 FMA per 11 LDS);
//   correlation: thread = (wide channel, filter row) with a KS-wide sliding register window over the pixel row, fully
//            unrolled (15 FMA per 2 LDS), persistent blocks, partials folded by k_narrow_corr_reduce in fixed order.
#include "kernels.h"

namespace sivae {

static inline unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------------------------
// narrow-in forward.  Block = 8 rows x 32 cols of output pixels, one warp per row.  Lane = (channel quad cq = lane & 15,
// pixel parity ph = lane >> 4): a lane accumulates 4 output channels for 8 pixels at a time, so the filter quad is one
// conflict-free LDS.128 reused over 8 pixels (96 FMA per 11 LDS) and the 16 lanes of a pixel write 256 contiguous bytes
// (the first version stored 16 B per lane at a 256 B stride: store-bound at ~0.5 TB/s, profiles/r01d_layers_H_tmastore.md)
// ---------------------------------------------------------------------------------------------------------------
constexpr int NI_TH = 8, NI_TW = 32;
template <int KS>
__global__ void __launch_bounds__(256) k_narrow_in_fwd(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* addend, float* y, int N,
                                                       int H, int W, int A, int Cout) {
  extern __shared__ float sm[];
  constexpr int HTH = NI_TH + KS - 1, HTW = NI_TW + KS - 1;
  float* sx = sm;                                   // [HTH][HTW][4]  (channels padded to 4: one LDS.128 per pixel)
  float* sw = sm + HTH * HTW * 4;                   // [tap][a][Cout]
  const int tiles_w = (W + NI_TW - 1) / NI_TW, tiles_h = (H + NI_TH - 1) / NI_TH;
  const int tile = blockIdx.x;
  const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
  const int w0 = tw * NI_TW, h0 = th * NI_TH, pad = KS / 2;
  constexpr int taps = KS * KS;
  // w is PRE-TRANSPOSED to [tap][a][Cout] (launch_narrow_transpose): straight 128-bit copy.  (Transposing here cost a
  // 32-way bank-conflicted store + three integer divisions per element in every one of the 8192 blocks: 70 % of the
  // kernel, profiles/r01g_prof_narrow.md.)
  for (int i = threadIdx.x; i < (Cout * taps * A) >> 2; i += 256)
    reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  for (int i = threadIdx.x; i < HTH * HTW * 4; i += 256) {
    int a = i & 3, cx = (i >> 2) % HTW, cy = i / (4 * HTW);
    int hh = h0 + cy - pad, ww = w0 + cx - pad;
    sx[i] = (a < A && hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(((long long)n * H + hh) * W + ww) * A + a] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  const int cq = lane & 15, ph = lane >> 4;
  const int ho = h0 + row;
  for (int c0 = 0; c0 < Cout; c0 += 64) {
    const int co = c0 + 4 * cq;
    const bool cvalid = co < Cout;
    const int cos = cvalid ? co : 0;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int r = 0; r < KS; ++r)
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const float* wp = sw + (r * KS + s) * A * Cout + cos;
          float4 wq[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) wq[a] = a < A ? *reinterpret_cast<const float4*>(wp + a * Cout) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float* xp = sx + ((row + r) * HTW + ph + 16 * g + s) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 xv = *reinterpret_cast<const float4*>(xp + 8 * i);      // pixel ph + 2*(8g + i), tap column s
            acc[i][0] = fmaf(xv.x, wq[0].x, fmaf(xv.y, wq[1].x, fmaf(xv.z, wq[2].x, fmaf(xv.w, wq[3].x, acc[i][0]))));
            acc[i][1] = fmaf(xv.x, wq[0].y, fmaf(xv.y, wq[1].y, fmaf(xv.z, wq[2].y, fmaf(xv.w, wq[3].y, acc[i][1]))));
            acc[i][2] = fmaf(xv.x, wq[0].z, fmaf(xv.y, wq[1].z, fmaf(xv.z, wq[2].z, fmaf(xv.w, wq[3].z, acc[i][2]))));
            acc[i][3] = fmaf(xv.x, wq[0].w, fmaf(xv.y, wq[1].w, fmaf(xv.z, wq[2].w, fmaf(xv.w, wq[3].w, acc[i][3]))));
          }
        }
      if (cvalid && ho < H) {
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) bq = *reinterpret_cast<const float4*>(bias + co);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int wo = w0 + ph + 2 * (8 * g + i);
          if (wo < W) {
            const long long pix = ((long long)n * H + ho) * W + wo;
            float4 o = make_float4(acc[i][0] + bq.x, acc[i][1] + bq.y, acc[i][2] + bq.z, acc[i][3] + bq.w);
            if (addend) {
              float4 q = *reinterpret_cast<const float4*>(addend + pix * Cout + co);
              o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            *reinterpret_cast<float4*>(y + pix * Cout + co) = o;
          }
        }
      }
    }
  }
}

// wt[tap][a][co] = w[co][tap][a]
__global__ void k_narrow_transpose(const float* __restrict__ w, float* __restrict__ wt, int Cout, int taps, int A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * taps * A) return;
  int a = i % A, t = (i / A) % taps, co = i / (A * taps);
  wt[(t * A + a) * Cout + co] = w[i];
}
void launch_narrow_transpose(const float* w, float* wt, int Cout, int k, int A, cudaStream_t st) {
  g_launches += 1;
  int n = Cout * k * k * A;
  k_narrow_transpose<<<cdivu(n, 256), 256, 0, st>>>(w, wt, Cout, k * k, A);
}
bool conv_narrow_in_supported(const ConvShape& s) {
  return s.Cin >= 1 && s.Cin <= 4 && s.Cout % 4 == 0 && s.Cout <= 256 && (s.k == 5 || s.k == 3) && s.N * (long long)s.H * s.W > 0;
}
// wt: the filter transposed to [tap][Cin][Cout] by launch_narrow_transpose
void launch_conv_narrow_in_fwd(const float* x, const float* wt, const float* bias, const float* addend, float* y,
                               const ConvShape& s, cudaStream_t st) {
  const float* w = wt;
  g_launches += 1;
  const int HTH = NI_TH + s.k - 1, HTW = NI_TW + s.k - 1;
  size_t shmem = ((size_t)HTH * HTW * 4 + (size_t)s.Cout * s.k * s.k * s.Cin) * sizeof(float);
  unsigned grid = (unsigned)(cdivu(s.W, NI_TW) * cdivu(s.H, NI_TH) * s.N);
  if (s.k == 5) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_narrow_in_fwd<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
    k_narrow_in_fwd<5><<<grid, 256, shmem, st>>>(x, w, bias, addend, y, s.N, s.H, s.W, s.Cin, s.Cout);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_narrow_in_fwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; }
    k_narrow_in_fwd<3><<<grid, 256, shmem, st>>>(x, w, bias, addend, y, s.N, s.H, s.W, s.Cin, s.Cout);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// narrow x wide correlation (wgrad of both image-facing layers)
//   G[tap][a][c] = sum_{n,h,w} nar[n,h,w][a] * wide[n, h + r - pad, w + s - pad][c]
// Persistent blocks loop over pixel tiles of TR x 32.  Thread = (wide channel c, filter row r): it owns the KS taps
// (r, 0..KS-1) and walks along the image row with a KS-wide sliding window of wide[.][c] kept in registers, so each
// step costs one new shared load + one broadcast load of the narrow pixel for KS*A FMAs.  Per-block partial G, reduced
// in fixed order (deterministic).
// ---------------------------------------------------------------------------------------------------------------
constexpr int NC_TW = 32;
template <int C, int KS>
struct NarrowCorrCfg {
  static constexpr int THREADS = C * KS;
  static constexpr int TAPS = KS * KS;
  static constexpr int TR = (C <= 64) ? 8 : 4;                    // tile rows
  static constexpr int HTH = TR + KS - 1, HTW = NC_TW + KS - 1;
  static constexpr size_t SMEM = ((size_t)HTH * HTW * C + (size_t)TR * NC_TW * 4) * sizeof(float);
};
template <int C, int KS, bool A4>
__global__ void __launch_bounds__(C* KS) k_narrow_corr(const float* __restrict__ nar, const float* __restrict__ wide,
                                                       float* __restrict__ part, int N, int H, int W, int A) {
  using CF = NarrowCorrCfg<C, KS>;
  extern __shared__ float sm[];
  float* swd = sm;                                  // [HTH][HTW][C]
  float* snr = sm + CF::HTH * CF::HTW * C;          // [TR][32][4]
  const int c = threadIdx.x % C, r = threadIdx.x / C;
  const int pad = KS / 2;
  float acc[KS][4];
#pragma unroll
  for (int i = 0; i < KS; ++i)
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[i][a] = 0.f;
  const int tiles_w = (W + NC_TW - 1) / NC_TW, tiles_h = (H + CF::TR - 1) / CF::TR;
  const long long n_tiles = (long long)tiles_w * tiles_h * N;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w), th = (int)((tile / tiles_w) % tiles_h), n = (int)(tile / ((long long)tiles_w * tiles_h));
    const int w0 = tw * NC_TW, h0 = th * CF::TR;
    __syncthreads();
    for (int i = threadIdx.x; i < CF::HTH * CF::HTW * (C / 4); i += CF::THREADS) {
      int c4 = i % (C / 4), cx = (i / (C / 4)) % CF::HTW, cy = i / ((C / 4) * CF::HTW);
      int hh = h0 + cy - pad, ww = w0 + cx - pad;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(reinterpret_cast<const float4*>(wide + (((long long)n * H + hh) * W + ww) * C) + c4);
      reinterpret_cast<float4*>(swd)[i] = v;
    }
    for (int i = threadIdx.x; i < CF::TR * NC_TW * 4; i += CF::THREADS) {
      int a = i & 3, px = (i >> 2) % NC_TW, py = i / (4 * NC_TW);
      int hh = h0 + py, ww = w0 + px;
      snr[i] = (a < A && hh < H && ww < W) ? nar[(((long long)n * H + hh) * W + ww) * A + a] : 0.f;
    }
    __syncthreads();
    for (int py = 0; py < CF::TR; ++py) {
      const float* wrow = swd + ((py + r) * CF::HTW) * C + c;       // wide[h0+py+r-pad][w0-pad ...][c]
      float win[KS];
#pragma unroll
      for (int s = 0; s < KS - 1; ++s) win[s + 1] = wrow[s * C];    // preload: after the first shift win[0..KS-2] hold cols 0..KS-2
#pragma unroll
      for (int px = 0; px < NC_TW; ++px) {          // fully unrolled: the window shift becomes register renaming
#pragma unroll
        for (int s = 0; s < KS - 1; ++s) win[s] = win[s + 1];
        win[KS - 1] = wrow[(px + KS - 1) * C];
        const float4 nv = *reinterpret_cast<const float4*>(snr + (py * NC_TW + px) * 4);   // broadcast
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          acc[s][0] = fmaf(nv.x, win[s], acc[s][0]);
          acc[s][1] = fmaf(nv.y, win[s], acc[s][1]);
          acc[s][2] = fmaf(nv.z, win[s], acc[s][2]);
          if (A4) acc[s][3] = fmaf(nv.w, win[s], acc[s][3]);
        }
      }
    }
  }
  // part[block][tap][a(4)][C]
  float* dst = part + (long long)blockIdx.x * CF::TAPS * 4 * C;
#pragma unroll
  for (int s = 0; s < KS; ++s)
#pragma unroll
    for (int a = 0; a < 4; ++a) dst[((r * KS + s) * 4 + a) * C + c] = acc[s][a];
}
// mode 0 (predict wgrad): dw[a][tap][c]   = G[tap][a][c]             (nar = dy (A = Cout), wide = x)
// mode 1 (stem wgrad):    dw[c][tap][a]   = G[mirror(tap)][a][c]     (nar = x  (A = Cin),  wide = dy)
__global__ void k_narrow_corr_reduce(const float* __restrict__ part, float* __restrict__ dw, int nblk, int taps, int A, int C,
                                     int mode, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;     // over taps*A*C outputs in G order
  if (i >= taps * A * C) return;
  int c = i % C, a = (i / C) % A, t = i / (C * A);
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += (double)part[((long long)b * taps + t) * 4 * C + a * C + c];
  long long o = mode == 0 ? ((long long)a * taps + t) * C + c : ((long long)c * taps + (taps - 1 - t)) * A + a;
  dw[o] = accumulate ? (float)((double)dw[o] + s) : (float)s;
}

static int narrow_corr_blocks() { return 148; }      // 1 persistent block per SM (110-147 KB of shared memory each)
bool conv_narrow_corr_supported(int wideC, int narrowA, int k) {
  return (wideC == 32 || wideC == 64 || wideC == 128) && narrowA >= 1 && narrowA <= 4 && k == 5;
}
size_t conv_narrow_corr_scratch_bytes(int wideC, int k) { return (size_t)narrow_corr_blocks() * k * k * 4 * wideC * sizeof(float); }

template <int C>
static void launch_corr_t(const float* nar, const float* wide, float* part, int N, int H, int W, int A, int nblk, cudaStream_t st) {
  using CF = NarrowCorrCfg<C, 5>;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_narrow_corr<C, 5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM);
    cudaFuncSetAttribute(k_narrow_corr<C, 5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::SMEM);
    attr = true;
  }
  if (A == 4) k_narrow_corr<C, 5, true><<<nblk, CF::THREADS, CF::SMEM, st>>>(nar, wide, part, N, H, W, A);
  else k_narrow_corr<C, 5, false><<<nblk, CF::THREADS, CF::SMEM, st>>>(nar, wide, part, N, H, W, A);
}
void launch_conv_narrow_corr(const float* nar, const float* wide, float* dw, int N, int H, int W, int A, int C, int k, int mode,
                             bool accumulate, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  g_launches += 2;
  float* part = (float*)scratch;
  long long tiles = (long long)cdivu(W, NC_TW) * cdivu(H, C <= 64 ? 8 : 4) * N;
  int nblk = (int)(tiles < narrow_corr_blocks() ? tiles : narrow_corr_blocks());
  if (C == 32) launch_corr_t<32>(nar, wide, part, N, H, W, A, nblk, st);
  else if (C == 64) launch_corr_t<64>(nar, wide, part, N, H, W, A, nblk, st);
  else launch_corr_t<128>(nar, wide, part, N, H, W, A, nblk, st);
  int total = k * k * A * C;
  k_narrow_corr_reduce<<<cdivu(total, 128), 128, 0, st>>>(part, dw, nblk, k * k, A, C, mode, accumulate ? 1 : 0);
}

}  // namespace sivae
