// Memory-bound kernels of the Soft-IntroVAE step (layout, train-mode BatchNorm + LeakyReLU + residual + pool /
// upsample and their backward, linear layers, the fused loss pass, Adam) and the exact fp32 SIMT implicit-GEMM
// convolution used for the 3-channel stem / predict layers and as the on-device cross-check of the tcgen05 path.
// Reference semantics: soft_intro_vae/train_soft_intro_vae.py (lines cited per kernel).
#include "kernels.h"
#include "split32.cuh"
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>

namespace sivae {

unsigned long long g_launches = 0;   // number of kernels enqueued by this library (bench.py reports it)
static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

__device__ __forceinline__ float round_tf32_dev(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float lrelu(float y) { return y > 0.f ? y : kSlope * y; }

// =====================================================================================================
// layout
// =====================================================================================================
__global__ void k_nchw_to_nhwc(const float* __restrict__ in, float* __restrict__ out, int N, int C, int H, int W) {
  long long total = (long long)N * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the NHWC output (coalesced writes); reads are strided by H*W but C is small / tensor is tiny
    int c = (int)(i % C);
    long long p = i / C;
    int w = (int)(p % W);
    long long q = p / W;
    int h = (int)(q % H);
    int n = (int)(q / H);
    out[i] = in[(((long long)n * C + c) * H + h) * W + w];
  }
}
__global__ void k_nhwc_to_nchw(const float* __restrict__ in, float* __restrict__ out, int N, int C, int H, int W) {
  long long total = (long long)N * C * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the NCHW output
    int w = (int)(i % W);
    long long q = i / W;
    int h = (int)(q % H);
    q /= H;
    int c = (int)(q % C);
    int n = (int)(q / C);
    out[i] = in[(((long long)n * H + h) * W + w) * C + c];
  }
}
void launch_nchw_to_nhwc(const float* in, float* out, int N, int C, int H, int W, cudaStream_t st) {
  g_launches += 1;
  long long total = (long long)N * C * H * W;
  if (total == 0) return;
  k_nchw_to_nhwc<<<min(cdiv(total, 256), 148u * 16), 256, 0, st>>>(in, out, N, C, H, W);
}
void launch_nhwc_to_nchw(const float* in, float* out, int N, int C, int H, int W, cudaStream_t st) {
  g_launches += 1;
  long long total = (long long)N * C * H * W;
  if (total == 0) return;
  k_nhwc_to_nchw<<<min(cdiv(total, 256), 148u * 16), 256, 0, st>>>(in, out, N, C, H, W);
}
__global__ void k_fill(float* p, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill(float* p, float v, long long n, cudaStream_t st) {
  g_launches += 1;
  if (n <= 0) return;
  k_fill<<<min(cdiv(n, 256), 148u * 8), 256, 0, st>>>(p, v, n);
}
__global__ void k_round_tf32(const float* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = round_tf32_dev(in[i]);
}
void launch_round_tf32(const float* in, float* out, long long n, cudaStream_t st) {
  g_launches += 1;
  if (n <= 0) return;
  k_round_tf32<<<min(cdiv(n, 256), 148u * 8), 256, 0, st>>>(in, out, n);
}
// 3xTF32 operand split (compensated tensor-core mode): hi = rna_tf32(x), lo = rna_tf32(x - hi).  x - hi is exact in fp32
// (13 significant bits), so hi + lo carries 21-22 bits of x and both parts are exactly representable tf32 operands.
__global__ void k_split_tf32(const float* in, float* hi, float* __restrict__ lo, long long n) {   // hi may alias in
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    float4 h, l;
    h.x = round_tf32_dev(v.x); h.y = round_tf32_dev(v.y); h.z = round_tf32_dev(v.z); h.w = round_tf32_dev(v.w);
    l.x = round_tf32_dev(v.x - h.x); l.y = round_tf32_dev(v.y - h.y); l.z = round_tf32_dev(v.z - h.z); l.w = round_tf32_dev(v.w - h.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const float v = in[i], h = round_tf32_dev(v);
    hi[i] = h;
    lo[i] = round_tf32_dev(v - h);
  }
}
void launch_split_tf32(const float* in, float* hi, float* lo, long long n, cudaStream_t st) {
  g_launches += 1;
  if (n <= 0) return;
  k_split_tf32<<<min(cdiv((n >> 2) + 1, 256), 148u * 16), 256, 0, st>>>(in, hi, lo, n);
}
// fp32 -> split32 (split32.cuh).  In-place safe: the 8 lanes that own one 128-byte group have all read their float4 before
// any of them writes (warps stay converged: the loop bound is rounded up to a whole warp).
__global__ void k_split32(const float* in, float* out, long long n4) {
  const long long n4r = (n4 + 31) / 32 * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4r; i += (long long)gridDim.x * blockDim.x) {
    const bool ok = i < n4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) v = reinterpret_cast<const float4*>(in)[i];
    __syncwarp();
    if (ok) split32_store4(out + (i >> 3) * 32, (uint32_t)(i & 7), v);
  }
}
void launch_split32(const float* in, float* out, long long n, cudaStream_t st) {
  g_launches += 1;
  if (n <= 0) return;
  k_split32<<<min(cdiv(n >> 2, 256), 148u * 16), 256, 0, st>>>(in, out, n >> 2);
}
// four floats -> four bf16 (round to nearest even), 8 bytes
__device__ __forceinline__ uint2 pack_bf16x4(const float4& v) {
  const uint32_t a = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.x)), b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.y));
  const uint32_t c = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.z)), d = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.w));
  return make_uint2(a | (b << 16), c | (d << 16));
}
__global__ void k_to_bf16(const float* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    out[i] = pack_bf16x4(reinterpret_cast<const float4*>(in)[i]);
}
void launch_to_bf16(const float* in, void* out, long long n, cudaStream_t st) {
  g_launches += 1;
  if (n <= 0) return;
  k_to_bf16<<<min(cdiv(n >> 2, 256), 148u * 16), 256, 0, st>>>(in, (uint2*)out, n >> 2);
}
// bf16 form of the dgrad-packed filter (the dgrad of the residual blocks in the default mode)
__global__ void k_pack_dgrad_bf16(const float* __restrict__ w, __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int k) {
  long long total = (long long)Cout * Cin * k * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % Cout);
    long long q = i / Cout;
    int s2 = (int)(q % k);
    q /= k;
    int r2 = (int)(q % k);
    int ci = (int)(q / k);
    wd[i] = __float2bfloat16_rn(w[(((long long)co * k + (k - 1 - r2)) * k + (k - 1 - s2)) * Cin + ci]);
  }
}
void launch_pack_dgrad_filter_bf16(const float* w, void* wd, int Cout, int Cin, int k, cudaStream_t st) {
  g_launches += 1;
  long long total = (long long)Cout * Cin * k * k;
  k_pack_dgrad_bf16<<<min(cdiv(total, 256), 148u * 8), 256, 0, st>>>(w, (__nv_bfloat16*)wd, Cout, Cin, k);
}
__global__ void k_pack_dgrad(const float* __restrict__ w, float* __restrict__ wd, int Cout, int Cin, int k, int rnd) {
  long long total = (long long)Cout * Cin * k * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes wd[ci][r'][s'][co]
    int co = (int)(i % Cout);
    long long q = i / Cout;
    int s2 = (int)(q % k);
    q /= k;
    int r2 = (int)(q % k);
    int ci = (int)(q / k);
    float v = w[(((long long)co * k + (k - 1 - r2)) * k + (k - 1 - s2)) * Cin + ci];
    wd[i] = rnd ? round_tf32_dev(v) : v;
  }
}
void launch_pack_dgrad_filter(const float* w, float* wd, int Cout, int Cin, int k, bool rnd, cudaStream_t st) {
  g_launches += 1;
  long long total = (long long)Cout * Cin * k * k;
  k_pack_dgrad<<<min(cdiv(total, 256), 148u * 8), 256, 0, st>>>(w, wd, Cout, Cin, k, rnd ? 1 : 0);
}

// =====================================================================================================
// SIMT implicit-GEMM convolution, fp32 exact.  GEMM view: M = N*H*W pixels, N = Cout, K = k*k*Cin.
// 128x64 tile, BK = 16, 256 threads, 8x4 micro-tile.
// =====================================================================================================
constexpr int CS_BM = 128, CS_BN = 64, CS_BK = 16;

__global__ void __launch_bounds__(256) k_conv_fwd_simt(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* addend,
                                                       float* y, int N, int H, int W, int Cin, int Cout, int ks) {
  __shared__ float As[CS_BK][CS_BM + 4];
  __shared__ float Bs[CS_BK][CS_BN + 4];
  const int tid = threadIdx.x;
  const long long M = (long long)N * H * W;
  const int Ktot = ks * ks * Cin;
  const int pad = ks / 2;
  const long long m0 = (long long)blockIdx.x * CS_BM;
  const int n0 = blockIdx.y * CS_BN;

  // A loader: thread -> (kk = tid%16, rows tid/16 + 16*j)
  const int a_kk = tid & 15;
  const int a_r0 = tid >> 4;
  int an[8], ah[8], aw[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    long long m = m0 + a_r0 + 16 * j;
    if (m < M) {
      int wq = (int)(m % W);
      long long q = m / W;
      ah[j] = (int)(q % H);
      an[j] = (int)(q / H);
      aw[j] = wq;
    } else {
      an[j] = -1; ah[j] = 0; aw[j] = 0;
    }
  }
  // B loader: thread -> (kk = tid%16, cols tid/16 + 16*j), j<4
  const int b_kk = tid & 15;
  const int b_c0 = tid >> 4;

  const int tx = tid & 15;   // column group (4 cols)
  const int ty = tid >> 4;   // row group (8 rows)
  // exact path: every 16-deep K chunk is accumulated in fp32 and folded into an fp64 running sum, so the result is
  // correct to fp32 round-off independent of K (this kernel is the on-device reference of the tensor-core path)
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < Ktot; k0 += CS_BK) {
    {
      int k = k0 + a_kk;
      int tap = 0, ci = 0, r = 0, s = 0;
      bool kvalid = k < Ktot;
      if (kvalid) { tap = k / Cin; ci = k - tap * Cin; r = tap / ks; s = tap - r * ks; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = 0.f;
        if (kvalid && an[j] >= 0) {
          int hh = ah[j] + r - pad, ww = aw[j] + s - pad;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[(((long long)an[j] * H + hh) * W + ww) * Cin + ci];
        }
        As[a_kk][a_r0 + 16 * j] = v;
      }
    }
    {
      int k = k0 + b_kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int co = n0 + b_c0 + 16 * j;
        Bs[b_kk][b_c0 + 16 * j] = (k < Ktot && co < Cout) ? w[(long long)co * Ktot + k] : 0.f;
      }
    }
    __syncthreads();
    float part[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < CS_BK; ++kk) {
      float a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += (double)part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= Cout) continue;
      double v = acc[i][j];
      if (bias) v += (double)bias[co];
      if (addend) v += (double)addend[m * Cout + co];
      y[m * Cout + co] = (float)v;
    }
  }
}
void launch_conv_fwd_simt(const float* x, const float* w, const float* bias, const float* addend, float* y,
                          const ConvShape& s, cudaStream_t st) {
  g_launches += 1;
  long long M = s.pixels();
  if (M == 0) return;
  dim3 grid(cdiv(M, CS_BM), cdiv(s.Cout, CS_BN));
  k_conv_fwd_simt<<<grid, 256, 0, st>>>(x, w, bias, addend, y, s.N, s.H, s.W, s.Cin, s.Cout, s.k);
}

// wgrad: D[co][k] = sum_p dy[p][co] * A(p,k).  64(co) x 64(k) tile, pixel chunks of 16, split over pixels.
constexpr int WG_BM = 64, WG_BN = 64, WG_BP = 16;
__global__ void __launch_bounds__(256) k_conv_wgrad_simt(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ part, int N, int H, int W, int Cin,
                                                         int Cout, int ks, long long pix_per_split) {
  __shared__ float Ds[WG_BP][WG_BM + 4];   // dy tile  [pixel][co]
  __shared__ float Xs[WG_BP][WG_BN + 4];   // im2col tile [pixel][k]
  const int tid = threadIdx.x;
  const long long M = (long long)N * H * W;
  const int Ktot = ks * ks * Cin;
  const int pad = ks / 2;
  const int co0 = blockIdx.y * WG_BM;
  const int k0 = blockIdx.x * WG_BN;
  const long long p_begin = (long long)blockIdx.z * pix_per_split;
  const long long p_end = min(M, p_begin + pix_per_split);

  // loaders: thread -> column (tid%64), pixel rows tid/64 + 4*j (j<4)
  const int lc = tid & 63;
  const int lr0 = tid >> 6;
  const int kcol = k0 + lc;
  int tap = 0, ci = 0, r = 0, s = 0;
  const bool kvalid = kcol < Ktot;
  if (kvalid) { tap = kcol / Cin; ci = kcol - tap * Cin; r = tap / ks; s = tap - r * ks; }
  const int co_l = co0 + lc;

  const int tx = tid & 15, ty = tid >> 4;   // 4x4 micro tile: rows (co) ty*4.., cols (k) tx*4..
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (long long p0 = p_begin; p0 < p_end; p0 += WG_BP) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int pr = lr0 + 4 * j;
      long long p = p0 + pr;
      float dv = 0.f, xv = 0.f;
      if (p < p_end) {
        if (co_l < Cout) dv = dy[p * Cout + co_l];
        if (kvalid) {
          int wq = (int)(p % W);
          long long q = p / W;
          int hq = (int)(q % H);
          int nq = (int)(q / H);
          int hh = hq + r - pad, ww = wq + s - pad;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) xv = x[(((long long)nq * H + hh) * W + ww) * Cin + ci];
        }
      }
      Ds[pr][lc] = dv;
      Xs[pr][lc] = xv;
    }
    __syncthreads();
    float part4[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part4[i][j] = 0.f;
#pragma unroll
    for (int pp = 0; pp < WG_BP; ++pp) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ds[pp][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Xs[pp][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part4[i][j] = fmaf(a[i], b[j], part4[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += (double)part4[i][j];
    __syncthreads();
  }
  float* dst = part + (long long)blockIdx.z * Cout * Ktot;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < Ktot) dst[(long long)co * Ktot + k] = (float)acc[i][j];
    }
  }
}
__global__ void k_splitk_reduce(const float* __restrict__ part, float* __restrict__ out, long long n, int splits, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += (double)part[(long long)z * n + i];   // fixed order: deterministic
    out[i] = accumulate ? (float)((double)out[i] + s) : (float)s;
  }
}
static int wgrad_simt_splits(const ConvShape& s) {
  long long tiles = (long long)cdiv(s.ktot(), WG_BN) * cdiv(s.Cout, WG_BM);
  long long want = (148 * 4 + tiles - 1) / tiles;
  long long maxs = (s.pixels() + 255) / 256;
  long long sp = want < maxs ? want : maxs;
  if (sp < 1) sp = 1;
  if (sp > 256) sp = 256;
  return (int)sp;
}
size_t conv_wgrad_simt_scratch_bytes(const ConvShape& s) {
  return (size_t)wgrad_simt_splits(s) * s.Cout * s.ktot() * sizeof(float);
}
void launch_conv_wgrad_simt(const float* x, const float* dy, float* dw, const ConvShape& s, bool accumulate,
                            void* scratch, size_t scratch_bytes, cudaStream_t st) {
  g_launches += 2;
  int splits = wgrad_simt_splits(s);
  long long M = s.pixels();
  long long pps = (M + splits - 1) / splits;
  pps = (pps + WG_BP - 1) / WG_BP * WG_BP;
  dim3 grid(cdiv(s.ktot(), WG_BN), cdiv(s.Cout, WG_BM), splits);
  float* part = (float*)scratch;
  k_conv_wgrad_simt<<<grid, 256, 0, st>>>(x, dy, part, s.N, s.H, s.W, s.Cin, s.Cout, s.k, pps);
  long long n = (long long)s.Cout * s.ktot();
  k_splitk_reduce<<<min(cdiv(n, 256), 148u * 8), 256, 0, st>>>(part, dw, n, splits, accumulate ? 1 : 0);
}

// out[c] (+)= sum_rows in[row][c] for a small channel count (predict bias gradient, C = cdim <= 8): two-stage,
// fixed-order (deterministic).  scratch: COLSUM_BLOCKS * C floats.
constexpr int COLSUM_BLOCKS = 148 * 2;
__global__ void __launch_bounds__(256) k_colsum_partial(const float* __restrict__ in, float* __restrict__ part, long long rows, int C) {
  __shared__ float sh[8][8];
  float acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(rows, r0 + per);
  for (long long r = r0 + threadIdx.x; r < r1; r += 256)
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) acc[c] += in[r * C + c];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[wid][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sh[k][threadIdx.x];
    part[(size_t)blockIdx.x * C + threadIdx.x] = v;
  }
}
__global__ void k_colsum_final(const float* __restrict__ part, float* __restrict__ out, int nblk, int C, int accumulate) {
  int c = threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += (double)part[(size_t)b * C + c];
  out[c] = accumulate ? out[c] + (float)s : (float)s;
}
size_t colsum_scratch_bytes(int C) { return (size_t)COLSUM_BLOCKS * C * sizeof(float); }
void launch_colsum(const float* in, float* out, long long rows, int C, bool accumulate, void* scratch, cudaStream_t st) {
  g_launches += 2;
  int nblk = (int)min((long long)COLSUM_BLOCKS, (rows + 255) / 256);
  if (nblk < 1) nblk = 1;
  k_colsum_partial<<<nblk, 256, 0, st>>>(in, (float*)scratch, rows, C);
  k_colsum_final<<<1, 32, 0, st>>>((const float*)scratch, out, nblk, C, accumulate ? 1 : 0);
}

// =====================================================================================================
// BatchNorm2d in train mode (:58,62,90; eps 1e-5, momentum 0.1): batch statistics over N*H*W
// =====================================================================================================
// rows handled by one block of the per-channel reductions: sized so that the grid is ~4 waves of 148 SMs even for the
// low-resolution layers (a fixed 1024 rows/block left the 4x4..16x16 layers on 1..8 SMs)
constexpr int BN_MAX_BLOCKS = 148 * 4;      // (148 * 8 blocks + a 4x unrolled row loop measured SLOWER: 21.9 vs 20.6 ms of BN backward
                                            // per config-H step, profiles/r02e_bench_H_bnreduce_grid.json)
static int bn_rows_per_block(long long rows) {
  long long r = (rows + BN_MAX_BLOCKS - 1) / BN_MAX_BLOCKS;
  if (r < 16) r = 16;
  return (int)((r + 15) / 16 * 16);
}
static int bn_nblocks(long long rows) { return (int)cdiv(rows, bn_rows_per_block(rows)); }
size_t bn_scratch_bytes(long long rows, int C) {
  // partial sums: [nblk][2][C] floats, plus 2*C floats of finalized sums for the backward
  return ((size_t)bn_nblocks(rows) * 2 * C + 2 * (size_t)C) * sizeof(float);
}

// each block: rpb rows; threads: cvec = C/4 lanes over channels x (256/cvec) row lanes
__global__ void __launch_bounds__(256) k_bn_stats_partial(const float* __restrict__ t, long long rows, int C,
                                                          float* __restrict__ part, int rpb) {
  extern __shared__ float sh[];   // [rl][2][C]
  const int cvec = C >> 2;
  const int rl_n = 256 / cvec;
  const int cl = threadIdx.x % cvec;
  const int rl = threadIdx.x / cvec;
  float4 s = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  if (rl < rl_n) {
    long long r0 = (long long)blockIdx.x * rpb;
    long long r1 = min(rows, r0 + rpb);
    for (long long r = r0 + rl; r < r1; r += rl_n) {
      float4 v = __ldg(reinterpret_cast<const float4*>(t + r * C) + cl);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    float* d = sh + (size_t)rl * 2 * C;
    reinterpret_cast<float4*>(d)[cl] = s;
    reinterpret_cast<float4*>(d + C)[cl] = q;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float a = 0.f;
    for (int k = 0; k < rl_n; ++k) a += sh[(size_t)k * 2 * C + i];
    part[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}
// one warp per channel: lanes stride over the per-block partials (fp64), shuffle-reduce, lane 0 finalises
__global__ void k_bn_stats_finalize(const float* __restrict__ part, int nblk, long long rows, int C,
                                    float* __restrict__ mean_invstd, float* running_mean, float* running_var,
                                    long long* nbt, double* __restrict__ ema_save) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c == 0 && lane == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    s += (double)part[(size_t)b * 2 * C + c];
    q += (double)part[(size_t)b * 2 * C + C + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane != 0) return;
  double mean = s / (double)rows;
  double var = q / (double)rows - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_invstd[c] = (float)mean;
  mean_invstd[C + c] = (float)(1.0 / sqrt(var + (double)kBnEps));
  if (running_mean) {
    double unb = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[c] = (float)((1.0 - kBnMomentum) * (double)running_mean[c] + kBnMomentum * mean);
    running_var[c] = (float)((1.0 - kBnMomentum) * (double)running_var[c] + kBnMomentum * unb);
    if (ema_save) { ema_save[c] = mean; ema_save[C + c] = unb; }     // the exact EMA inputs, for launch_bn_ema_replay
  }
}
// A second train-mode forward of the same net over the same input with unchanged weights (the D half's fake / rec,
// :597-598, repeat the E half's :557,:561) sees the same batch statistics: its only new effect is one more EMA update of
// every running buffer and num_batches_tracked.  `ema` mirrors the net's BN buffer layout ([mean C | unbiased var C] per
// layer) and holds the doubles k_bn_stats_finalize used, so the replayed update is bit-identical to recomputing the pass.
__global__ void k_bn_ema_replay(const double* __restrict__ ema, float* __restrict__ running, long long n,
                                long long* __restrict__ nbt, int n_bn) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) running[i] = (float)((1.0 - kBnMomentum) * (double)running[i] + kBnMomentum * ema[i]);
  if (i < n_bn) nbt[i] += 1;
}
void launch_bn_ema_replay(const double* ema, float* running, long long n, long long* nbt, int n_bn, cudaStream_t st) {
  g_launches += 1;
  long long m = n > n_bn ? n : n_bn;
  k_bn_ema_replay<<<cdiv(m, 256), 256, 0, st>>>(ema, running, n, nbt, n_bn);
}
void launch_bn_stats(const float* t, long long rows, int C, float* mean_invstd, float* running_mean,
                     float* running_var, long long* nbt, void* scratch, size_t scratch_bytes, cudaStream_t st, double* ema_save) {
  g_launches += 2;
  int nblk = bn_nblocks(rows);
  float* part = (float*)scratch;
  int cvec = C / 4;
  int rl_n = 256 / cvec;
  size_t shmem = (size_t)rl_n * 2 * C * sizeof(float);
  k_bn_stats_partial<<<nblk, 256, shmem, st>>>(t, rows, C, part, bn_rows_per_block(rows));
  k_bn_stats_finalize<<<cdiv(C, 8), 256, 0, st>>>(part, nblk, rows, C, mean_invstd, running_mean, running_var, nbt, ema_save);
}
// coalesced pre-reduction of many partial rows: out[b][i] = sum_{r in slab b} part[r][i], i over 2*C
__global__ void __launch_bounds__(256) k_parts_reduce(const float* __restrict__ part, float* __restrict__ out, int nparts, int twoC, int slab) {
  const int r0 = blockIdx.x * slab, r1 = min(nparts, r0 + slab);
  for (int i = threadIdx.x; i < twoC; i += 256) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += part[(size_t)r * twoC + i];
    out[(size_t)blockIdx.x * twoC + i] = s;
  }
}
size_t bn_parts_scratch_bytes(int nparts, int C) {
  return ((size_t)nparts + (size_t)(nparts + 63) / 64) * 2 * C * sizeof(float);
}
// part: [nparts][2][C] written by the conv epilogue; the pre-reduced rows are stored right behind it (same scratch)
void launch_bn_stats_from_parts(float* part, int nparts, long long rows, int C, float* mean_invstd, float* running_mean,
                                float* running_var, long long* nbt, cudaStream_t st, double* ema_save) {
  const float* src = part;
  int n = nparts;
  if (nparts > 512) {
    const int slab = 64;
    const int nb = (nparts + slab - 1) / slab;
    float* out = part + (size_t)nparts * 2 * C;
    g_launches += 1;
    k_parts_reduce<<<nb, 256, 0, st>>>(part, out, nparts, 2 * C, slab);
    src = out;
    n = nb;
  }
  g_launches += 1;
  k_bn_stats_finalize<<<cdiv(C, 8), 256, 0, st>>>(src, n, rows, C, mean_invstd, running_mean, running_var, nbt, ema_save);
}
__global__ void k_bn_eval_stats(const float* rm, const float* rv, int C, float* mi) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mi[c] = rm[c]; mi[C + c] = rsqrtf(rv[c] + kBnEps); }
}
void launch_bn_eval_stats(const float* rm, const float* rv, int C, float* mi, cudaStream_t st) {
  g_launches += 1;
  k_bn_eval_stats<<<cdiv(C, 128), 128, 0, st>>>(rm, rv, C, mi);
}

// out = resample(lrelu(bn(t) + identity)).  One thread per float4 of the OUTPUT for NONE/POOL, per float4 of the
// INPUT for UP.  (:65-75 ResidualBlock.forward, :90-92 stem, :98 AvgPool2d(2), :155 Upsample nearest x2)
// SPLIT: the output is (also) written in the split32 operand format of the forward convolutions to `outs`, from the
// UNROUNDED value; `out` (fp32, tf32-rounded if ROUND: the wgrad operand) is skipped when null; a split32 identity tensor
// (the block input when the block has no conv_expand) is read with idn_split != 0.
template <int MODE, bool ROUND, bool SPLIT>
__global__ void __launch_bounds__(256) k_bn_act_fwd(const float* __restrict__ t, const float* __restrict__ idn,
                                                    const float* __restrict__ mi, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float* __restrict__ out, int N,
                                                    int H, int W, int C, unsigned char* __restrict__ mask,
                                                    float* __restrict__ outs, int idn_split, uint2* __restrict__ outh) {
  // 32-bit index arithmetic (tensor sizes are < 2^31 float4s; checked by the launcher): 64-bit div/mod per element
  // would make this HBM-bound kernel instruction-bound
  const unsigned cvec = (unsigned)C >> 2;
  const unsigned Ho = MODE == RS_POOL ? H / 2 : H, Wo = MODE == RS_POOL ? W / 2 : W;
  const unsigned total = (unsigned)N * Ho * Wo * cvec;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % cvec;
    const unsigned p = i / cvec;
    unsigned w = 0, h = 0, n = 0;
    if (MODE != RS_NONE) {
      w = p % Wo;
      const unsigned q = p / Wo;
      h = q % Ho;
      n = q / Ho;
    }
    float4 mean = __ldg(reinterpret_cast<const float4*>(mi) + c4);
    float4 istd = __ldg(reinterpret_cast<const float4*>(mi + C) + c4);
    float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 sc = make_float4(g.x * istd.x, g.y * istd.y, g.z * istd.z, g.w * istd.w);
    float4 sf = make_float4(b.x - mean.x * sc.x, b.y - mean.y * sc.y, b.z - mean.z * sc.z, b.w - mean.w * sc.w);
    auto eval = [&](long long pix) -> float4 {
      float4 v = __ldg(reinterpret_cast<const float4*>(t + pix * C) + c4);
      float4 y = make_float4(fmaf(v.x, sc.x, sf.x), fmaf(v.y, sc.y, sf.y), fmaf(v.z, sc.z, sf.z), fmaf(v.w, sc.w, sf.w));
      if (idn) {
        float4 d = idn_split ? split32_load4(idn + pix * C, c4) : __ldg(reinterpret_cast<const float4*>(idn + pix * C) + c4);
        y.x += d.x; y.y += d.y; y.z += d.z; y.w += d.w;
      }
      // sign of the pre-activation, one byte per float4: the backward needs nothing else of y, so it can skip
      // re-reading the identity tensor (two of its eight full-tensor passes)
      if (mask)
        mask[pix * cvec + c4] = (unsigned char)((y.x > 0.f ? 1 : 0) | (y.y > 0.f ? 2 : 0) | (y.z > 0.f ? 4 : 0) | (y.w > 0.f ? 8 : 0));
      return make_float4(lrelu(y.x), lrelu(y.y), lrelu(y.z), lrelu(y.w));
    };
    auto rnd4 = [](float4 y) {
      if (ROUND) { y.x = round_tf32_dev(y.x); y.y = round_tf32_dev(y.y); y.z = round_tf32_dev(y.z); y.w = round_tf32_dev(y.w); }
      return y;
    };
    if (MODE == RS_NONE) {
      long long pix = p;
      const float4 y = eval(pix);
      if (SPLIT) split32_store4(outs + pix * C, c4, y);
      if (SPLIT && outh) outh[pix * cvec + c4] = pack_bf16x4(y);            // plain bf16 copy: the wgrad operand
      if (!SPLIT || out) reinterpret_cast<float4*>(out + pix * C)[c4] = rnd4(y);
    } else if (MODE == RS_POOL) {
      long long p00 = ((long long)n * H + 2 * h) * W + 2 * w;
      float4 a = eval(p00), b2 = eval(p00 + 1), c2 = eval(p00 + W), d2 = eval(p00 + W + 1);
      const float4 y = make_float4((a.x + b2.x + c2.x + d2.x) * 0.25f, (a.y + b2.y + c2.y + d2.y) * 0.25f,
                                   (a.z + b2.z + c2.z + d2.z) * 0.25f, (a.w + b2.w + c2.w + d2.w) * 0.25f);
      long long po = ((long long)n * Ho + h) * Wo + w;
      if (SPLIT) split32_store4(outs + po * C, c4, y);
      if (SPLIT && outh) outh[po * cvec + c4] = pack_bf16x4(y);
      if (!SPLIT || out) reinterpret_cast<float4*>(out + po * C)[c4] = rnd4(y);
    } else {
      long long pix = ((long long)n * H + h) * W + w;
      const float4 y = eval(pix);
      long long po = ((long long)n * (2 * H) + 2 * h) * (2 * W) + 2 * w;
      if (SPLIT) {
        uint2 hi, lo;
        split32_pack4(y, hi, lo);
        const uint32_t off = split32_off4(c4);
        const long long pp[4] = {po, po + 1, po + 2 * W, po + 2 * W + 1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          char* d = reinterpret_cast<char*>(outs + pp[k] * C) + off;
          *reinterpret_cast<uint2*>(d) = hi;
          *reinterpret_cast<uint2*>(d + 64) = lo;
        }
        if (outh) {
          const uint2 hb = pack_bf16x4(y);
#pragma unroll
          for (int k = 0; k < 4; ++k) outh[pp[k] * cvec + c4] = hb;
        }
      }
      if (!SPLIT || out) {
        const float4 yr = rnd4(y);
        reinterpret_cast<float4*>(out + po * C)[c4] = yr;
        reinterpret_cast<float4*>(out + (po + 1) * C)[c4] = yr;
        reinterpret_cast<float4*>(out + (po + 2 * W) * C)[c4] = yr;
        reinterpret_cast<float4*>(out + (po + 2 * W + 1) * C)[c4] = yr;
      }
    }
  }
}
void launch_bn_act_fwd(const float* t, const float* identity, const float* mi, const float* gamma, const float* beta,
                       float* out, int N, int H, int W, int C, int mode, bool rnd, cudaStream_t st, unsigned char* mask,
                       float* outs, bool idn_split, void* outh) {
  g_launches += 1;
  int Ho = mode == RS_POOL ? H / 2 : H, Wo = mode == RS_POOL ? W / 2 : W;
  long long total = (long long)N * Ho * Wo * (C / 4);
  if (total == 0) return;
  unsigned grid = min(cdiv(total, 256), 148u * 32);
  const int is = idn_split ? 1 : 0;
#define LAUNCH(M, R, S) k_bn_act_fwd<M, R, S><<<grid, 256, 0, st>>>(t, identity, mi, gamma, beta, out, N, H, W, C, mask, outs, is, (uint2*)outh)
  if (outs) {           // split32 output (C % 32 == 0), always together with tf32 rounding of the optional fp32 copy
    if (mode == RS_NONE) LAUNCH(RS_NONE, true, true);
    else if (mode == RS_POOL) LAUNCH(RS_POOL, true, true);
    else LAUNCH(RS_UP, true, true);
  } else if (mode == RS_NONE) { if (rnd) LAUNCH(RS_NONE, true, false); else LAUNCH(RS_NONE, false, false); }
  else if (mode == RS_POOL) { if (rnd) LAUNCH(RS_POOL, true, false); else LAUNCH(RS_POOL, false, false); }
  else { if (rnd) LAUNCH(RS_UP, true, false); else LAUNCH(RS_UP, false, false); }
#undef LAUNCH
}

// ---- backward -----------------------------------------------------------------------------------------
// upstream gradient at full resolution pixel (n,h,w): NONE: dout; POOL: dout[n,h/2,w/2]/4; UP: sum of the 4 children
template <int MODE>
__device__ __forceinline__ float4 upstream(const float* __restrict__ dout, long long row, int n, int h, int w, int H, int W, int C, int c4) {
  if (MODE == RS_NONE) {
    return __ldg(reinterpret_cast<const float4*>(dout + row * C) + c4);
  } else if (MODE == RS_POOL) {
    float4 v = __ldg(reinterpret_cast<const float4*>(dout + (((long long)n * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1)) * C) + c4);
    return make_float4(v.x * 0.25f, v.y * 0.25f, v.z * 0.25f, v.w * 0.25f);
  } else {
    long long po = ((long long)n * (2 * H) + 2 * h) * (2 * W) + 2 * w;
    float4 a = __ldg(reinterpret_cast<const float4*>(dout + po * C) + c4);
    float4 b = __ldg(reinterpret_cast<const float4*>(dout + (po + 1) * C) + c4);
    float4 c = __ldg(reinterpret_cast<const float4*>(dout + (po + 2 * W) * C) + c4);
    float4 d = __ldg(reinterpret_cast<const float4*>(dout + (po + 2 * W + 1) * C) + c4);
    return make_float4(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z, a.w + b.w + c.w + d.w);
  }
}
__device__ __forceinline__ float lrelu_grad(float y, float d) { return y > 0.f ? d : kSlope * d; }
// the same through the forward's sign byte (bit j = pre-activation of component j was positive)
__device__ __forceinline__ float4 lrelu_grad_mask(unsigned m, const float4& d) {
  return make_float4((m & 1u) ? d.x : kSlope * d.x, (m & 2u) ? d.y : kSlope * d.y, (m & 4u) ? d.z : kSlope * d.z, (m & 8u) ? d.w : kSlope * d.w);
}

template <int MODE>
__global__ void __launch_bounds__(256) k_bn_bwd_reduce(const float* __restrict__ dout, const float* __restrict__ t,
                                                       const float* __restrict__ idn, const float* __restrict__ mi,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int N, int H, int W, int C, float* __restrict__ part, int rpb,
                                                       const unsigned char* __restrict__ mask) {
  extern __shared__ float sh[];
  const int cvec = C >> 2;
  const int rl_n = 256 / cvec;
  const int cl = threadIdx.x % cvec;
  const int rl = threadIdx.x / cvec;
  const long long rows = (long long)N * H * W;
  float4 sg = make_float4(0, 0, 0, 0), sx = make_float4(0, 0, 0, 0);
  if (rl < rl_n) {
    float4 mean = __ldg(reinterpret_cast<const float4*>(mi) + cl);
    float4 istd = __ldg(reinterpret_cast<const float4*>(mi + C) + cl);
    float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cl);
    float4 be = __ldg(reinterpret_cast<const float4*>(beta) + cl);
    long long r0 = (long long)blockIdx.x * rpb;
    long long r1 = min(rows, r0 + rpb);
    for (long long r = r0 + rl; r < r1; r += rl_n) {
      int w = 0, h = 0, n = 0;
      if (MODE != RS_NONE) {
        const unsigned ru = (unsigned)r;
        w = (int)(ru % (unsigned)W);
        const unsigned q = ru / (unsigned)W;
        h = (int)(q % (unsigned)H);
        n = (int)(q / (unsigned)H);
      }
      float4 d = upstream<MODE>(dout, r, n, h, w, H, W, C, cl);
      float4 v = __ldg(reinterpret_cast<const float4*>(t + r * C) + cl);
      float4 xh = make_float4((v.x - mean.x) * istd.x, (v.y - mean.y) * istd.y, (v.z - mean.z) * istd.z, (v.w - mean.w) * istd.w);
      float4 g;
      if (mask) {
        g = lrelu_grad_mask(__ldg(mask + r * cvec + cl), d);
      } else {
        float4 y = make_float4(fmaf(xh.x, ga.x, be.x), fmaf(xh.y, ga.y, be.y), fmaf(xh.z, ga.z, be.z), fmaf(xh.w, ga.w, be.w));
        if (idn) {
          float4 e = __ldg(reinterpret_cast<const float4*>(idn + r * C) + cl);
          y.x += e.x; y.y += e.y; y.z += e.z; y.w += e.w;
        }
        g = make_float4(lrelu_grad(y.x, d.x), lrelu_grad(y.y, d.y), lrelu_grad(y.z, d.z), lrelu_grad(y.w, d.w));
      }
      sg.x += g.x; sg.y += g.y; sg.z += g.z; sg.w += g.w;
      sx.x = fmaf(g.x, xh.x, sx.x); sx.y = fmaf(g.y, xh.y, sx.y); sx.z = fmaf(g.z, xh.z, sx.z); sx.w = fmaf(g.w, xh.w, sx.w);
    }
    float* dd = sh + (size_t)rl * 2 * C;
    reinterpret_cast<float4*>(dd)[cl] = sg;
    reinterpret_cast<float4*>(dd + C)[cl] = sx;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float a = 0.f;
    for (int k = 0; k < rl_n; ++k) a += sh[(size_t)k * 2 * C + i];
    part[(size_t)blockIdx.x * 2 * C + i] = a;
  }
}
__global__ void k_bn_bwd_finalize(const float* __restrict__ part, int nblk, long long rows, int C,
                                  float* __restrict__ sums, float* dgamma, float* dbeta, int accumulate) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = lane; b < nblk; b += 32) {
    s += (double)part[(size_t)b * 2 * C + c];
    q += (double)part[(size_t)b * 2 * C + C + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane != 0) return;
  sums[c] = (float)(s / (double)rows);       // mean of g
  sums[C + c] = (float)(q / (double)rows);   // mean of g * xhat
  if (dgamma) {
    dgamma[c] = accumulate ? dgamma[c] + (float)q : (float)q;
    dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
  }
}
// out16 bit 0: dt is written as plain bf16 (the operand of the bf16 dgrad / wgrad); bit 1: so is gout
template <int MODE, bool ROUND>
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const float* __restrict__ dout, const float* __restrict__ t,
                                                      const float* __restrict__ idn, const float* __restrict__ mi,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ sums, float* __restrict__ dt,
                                                      float* __restrict__ gout, int N, int H, int W, int C,
                                                      const unsigned char* __restrict__ mask, int out16) {
  const unsigned cvec = (unsigned)C >> 2;
  const unsigned total = (unsigned)N * H * W * cvec;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cvec);
    const unsigned r = i / cvec;
    int w = 0, h = 0, n = 0;
    if (MODE != RS_NONE) {
      w = (int)(r % (unsigned)W);
      const unsigned q = r / (unsigned)W;
      h = (int)(q % (unsigned)H);
      n = (int)(q / (unsigned)H);
    }
    float4 mean = __ldg(reinterpret_cast<const float4*>(mi) + c4);
    float4 istd = __ldg(reinterpret_cast<const float4*>(mi + C) + c4);
    float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 mg = __ldg(reinterpret_cast<const float4*>(sums) + c4);
    float4 mx = __ldg(reinterpret_cast<const float4*>(sums + C) + c4);
    float4 d = upstream<MODE>(dout, (long long)r, n, h, w, H, W, C, c4);
    float4 v = __ldg(reinterpret_cast<const float4*>(t + r * C) + c4);
    float4 xh = make_float4((v.x - mean.x) * istd.x, (v.y - mean.y) * istd.y, (v.z - mean.z) * istd.z, (v.w - mean.w) * istd.w);
    float4 g;
    if (mask) {
      g = lrelu_grad_mask(__ldg(mask + (size_t)r * cvec + c4), d);
    } else {
      float4 y = make_float4(fmaf(xh.x, ga.x, be.x), fmaf(xh.y, ga.y, be.y), fmaf(xh.z, ga.z, be.z), fmaf(xh.w, ga.w, be.w));
      if (idn) {
        float4 e = __ldg(reinterpret_cast<const float4*>(idn + r * C) + c4);
        y.x += e.x; y.y += e.y; y.z += e.z; y.w += e.w;
      }
      g = make_float4(lrelu_grad(y.x, d.x), lrelu_grad(y.y, d.y), lrelu_grad(y.z, d.z), lrelu_grad(y.w, d.w));
    }
    float4 o;
    o.x = ga.x * istd.x * (g.x - mg.x - xh.x * mx.x);
    o.y = ga.y * istd.y * (g.y - mg.y - xh.y * mx.y);
    o.z = ga.z * istd.z * (g.z - mg.z - xh.z * mx.z);
    o.w = ga.w * istd.w * (g.w - mg.w - xh.w * mx.w);
    if (out16 & 1) {
      reinterpret_cast<uint2*>(dt)[(size_t)r * cvec + c4] = pack_bf16x4(o);
    } else {
      if (ROUND) {
        o.x = round_tf32_dev(o.x); o.y = round_tf32_dev(o.y); o.z = round_tf32_dev(o.z); o.w = round_tf32_dev(o.w);
      }
      reinterpret_cast<float4*>(dt + r * C)[c4] = o;
    }
    if (gout) {
      if (out16 & 2) {
        reinterpret_cast<uint2*>(gout)[(size_t)r * cvec + c4] = pack_bf16x4(g);
      } else {
        if (ROUND) { g.x = round_tf32_dev(g.x); g.y = round_tf32_dev(g.y); g.z = round_tf32_dev(g.z); g.w = round_tf32_dev(g.w); }
        reinterpret_cast<float4*>(gout + r * C)[c4] = g;
      }
    }
  }
}
// Small maps (rows <= BN_SMALL_ROWS: the 4x4 / 8x8 layers, 50+ launches per step each): the three launches above are pure
// latency there.  BatchNorm backward is channel-local, so one block per float4 channel group does everything: pass 1 over
// all rows (256 row lanes) for sum(g), sum(g*xhat); block reduction (shuffles, then fp64 over the 8 warps, fixed order);
// parameter gradients; pass 2 re-reads the rows (L1 / L2 hits) and writes dt (and the identity-branch gradient).
constexpr long long BN_SMALL_ROWS = 4096;
template <int MODE, bool ROUND>
__global__ void __launch_bounds__(256) k_bn_bwd_small(const float* __restrict__ dout, const float* __restrict__ t,
                                                      const float* __restrict__ idn, const float* __restrict__ mi,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ dt, float* __restrict__ gout, float* dgamma,
                                                      float* dbeta, int accumulate, int N, int H, int W, int C,
                                                      const unsigned char* __restrict__ mask, int out16) {
  const int c4 = blockIdx.x;
  const unsigned cvec = (unsigned)C >> 2;
  const unsigned rows = (unsigned)N * H * W;
  const float4 mean = __ldg(reinterpret_cast<const float4*>(mi) + c4);
  const float4 istd = __ldg(reinterpret_cast<const float4*>(mi + C) + c4);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  auto elem = [&](unsigned r, float4& g, float4& xh) {
    int w = 0, h = 0, n = 0;
    if (MODE != RS_NONE) {
      w = (int)(r % (unsigned)W);
      const unsigned q = r / (unsigned)W;
      h = (int)(q % (unsigned)H);
      n = (int)(q / (unsigned)H);
    }
    const float4 d = upstream<MODE>(dout, (long long)r, n, h, w, H, W, C, c4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(t + (size_t)r * C) + c4);
    xh = make_float4((v.x - mean.x) * istd.x, (v.y - mean.y) * istd.y, (v.z - mean.z) * istd.z, (v.w - mean.w) * istd.w);
    if (mask) {
      g = lrelu_grad_mask(__ldg(mask + (size_t)r * cvec + c4), d);
    } else {
      float4 y = make_float4(fmaf(xh.x, ga.x, be.x), fmaf(xh.y, ga.y, be.y), fmaf(xh.z, ga.z, be.z), fmaf(xh.w, ga.w, be.w));
      if (idn) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(idn + (size_t)r * C) + c4);
        y.x += e.x; y.y += e.y; y.z += e.z; y.w += e.w;
      }
      g = make_float4(lrelu_grad(y.x, d.x), lrelu_grad(y.y, d.y), lrelu_grad(y.z, d.z), lrelu_grad(y.w, d.w));
    }
  };
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // sum g (4), sum g*xhat (4)
  for (unsigned r = threadIdx.x; r < rows; r += 256) {
    float4 g, xh;
    elem(r, g, xh);
    acc[0] += g.x; acc[1] += g.y; acc[2] += g.z; acc[3] += g.w;
    acc[4] = fmaf(g.x, xh.x, acc[4]); acc[5] = fmaf(g.y, xh.y, acc[5]); acc[6] = fmaf(g.z, xh.z, acc[6]); acc[7] = fmaf(g.w, xh.w, acc[7]);
  }
  __shared__ double red[8][8];
  __shared__ float fin[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = (double)v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int k = threadIdx.x;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][k];
    fin[k] = (float)(s / (double)rows);                            // mean of g / of g*xhat
    const int c = c4 * 4 + (k & 3);
    if (dgamma) {
      if (k < 4) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
      else dgamma[c] = accumulate ? dgamma[c] + (float)s : (float)s;
    }
  }
  __syncthreads();
  const float4 mg = make_float4(fin[0], fin[1], fin[2], fin[3]), mx = make_float4(fin[4], fin[5], fin[6], fin[7]);
  for (unsigned r = threadIdx.x; r < rows; r += 256) {
    float4 g, xh;
    elem(r, g, xh);
    float4 o;
    o.x = ga.x * istd.x * (g.x - mg.x - xh.x * mx.x);
    o.y = ga.y * istd.y * (g.y - mg.y - xh.y * mx.y);
    o.z = ga.z * istd.z * (g.z - mg.z - xh.z * mx.z);
    o.w = ga.w * istd.w * (g.w - mg.w - xh.w * mx.w);
    if (out16 & 1) {
      reinterpret_cast<uint2*>(dt)[(size_t)r * cvec + c4] = pack_bf16x4(o);
    } else {
      if (ROUND) {
        o.x = round_tf32_dev(o.x); o.y = round_tf32_dev(o.y); o.z = round_tf32_dev(o.z); o.w = round_tf32_dev(o.w);
      }
      reinterpret_cast<float4*>(dt + (size_t)r * C)[c4] = o;
    }
    if (gout) {
      if (out16 & 2) {
        reinterpret_cast<uint2*>(gout)[(size_t)r * cvec + c4] = pack_bf16x4(g);
      } else {
        if (ROUND) { g.x = round_tf32_dev(g.x); g.y = round_tf32_dev(g.y); g.z = round_tf32_dev(g.z); g.w = round_tf32_dev(g.w); }
        reinterpret_cast<float4*>(gout + (size_t)r * C)[c4] = g;
      }
    }
  }
}
static bool bn_small_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SIVAE_BN_SMALL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
void launch_bn_act_bwd(const float* dout, const float* t, const float* identity, const float* mi, const float* gamma,
                       const float* beta, float* dt, float* g, float* dgamma, float* dbeta, bool accumulate, int N,
                       int H, int W, int C, int mode, bool rnd, void* scratch, size_t scratch_bytes, cudaStream_t st,
                       const unsigned char* mask, int out16) {
  long long rows = (long long)N * H * W;
  if (rows == 0) return;
  int cvec = C / 4;
  if (rows <= BN_SMALL_ROWS && bn_small_enabled()) {
    g_launches += 1;
#define LAUNCH(M, R) k_bn_bwd_small<M, R><<<cvec, 256, 0, st>>>(dout, t, identity, mi, gamma, beta, dt, g, dgamma, dbeta, accumulate ? 1 : 0, N, H, W, C, mask, out16)
    if (mode == RS_NONE) { if (rnd) LAUNCH(RS_NONE, true); else LAUNCH(RS_NONE, false); }
    else if (mode == RS_POOL) { if (rnd) LAUNCH(RS_POOL, true); else LAUNCH(RS_POOL, false); }
    else { if (rnd) LAUNCH(RS_UP, true); else LAUNCH(RS_UP, false); }
#undef LAUNCH
    return;
  }
  g_launches += 3;
  int nblk = bn_nblocks(rows);
  const int rpb = bn_rows_per_block(rows);
  float* part = (float*)scratch;
  float* sums = part + (size_t)nblk * 2 * C;
  int rl_n = 256 / cvec;
  size_t shmem = (size_t)rl_n * 2 * C * sizeof(float);
  if (mode == RS_NONE) k_bn_bwd_reduce<RS_NONE><<<nblk, 256, shmem, st>>>(dout, t, identity, mi, gamma, beta, N, H, W, C, part, rpb, mask);
  else if (mode == RS_POOL) k_bn_bwd_reduce<RS_POOL><<<nblk, 256, shmem, st>>>(dout, t, identity, mi, gamma, beta, N, H, W, C, part, rpb, mask);
  else k_bn_bwd_reduce<RS_UP><<<nblk, 256, shmem, st>>>(dout, t, identity, mi, gamma, beta, N, H, W, C, part, rpb, mask);
  k_bn_bwd_finalize<<<cdiv(C, 8), 256, 0, st>>>(part, nblk, rows, C, sums, dgamma, dbeta, accumulate ? 1 : 0);
  long long total = rows * cvec;
  unsigned grid = min(cdiv(total, 256), 148u * 32);
#define LAUNCH(M, R) k_bn_bwd_apply<M, R><<<grid, 256, 0, st>>>(dout, t, identity, mi, gamma, beta, sums, dt, g, N, H, W, C, mask, out16)
  if (mode == RS_NONE) { if (rnd) LAUNCH(RS_NONE, true); else LAUNCH(RS_NONE, false); }
  else if (mode == RS_POOL) { if (rnd) LAUNCH(RS_POOL, true); else LAUNCH(RS_POOL, false); }
  else { if (rnd) LAUNCH(RS_UP, true); else LAUNCH(RS_UP, false); }
#undef LAUNCH
}

// =====================================================================================================
// nn.Linear (encoder fc :109, decoder fc + ReLU :145-148).  Batch is tiny (<= 128): weight-streaming kernels.
// =====================================================================================================
constexpr int LF_OT = 4, LF_BT = 8;
__global__ void __launch_bounds__(256) k_linear_fwd(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ b, float* __restrict__ y, int B, int F,
                                                    int O, int relu) {
  const int o0 = blockIdx.x * LF_OT, b0 = blockIdx.y * LF_BT;
  float acc[LF_OT][LF_BT];
#pragma unroll
  for (int i = 0; i < LF_OT; ++i)
#pragma unroll
    for (int j = 0; j < LF_BT; ++j) acc[i][j] = 0.f;
  for (int f = threadIdx.x; f < F; f += 256) {
    float wv[LF_OT], xv[LF_BT];
#pragma unroll
    for (int i = 0; i < LF_OT; ++i) wv[i] = (o0 + i < O) ? __ldg(w + (long long)(o0 + i) * F + f) : 0.f;
#pragma unroll
    for (int j = 0; j < LF_BT; ++j) xv[j] = (b0 + j < B) ? __ldg(x + (long long)(b0 + j) * F + f) : 0.f;
#pragma unroll
    for (int i = 0; i < LF_OT; ++i)
#pragma unroll
      for (int j = 0; j < LF_BT; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
  }
  __shared__ float red[8][LF_OT * LF_BT];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < LF_OT; ++i)
#pragma unroll
    for (int j = 0; j < LF_BT; ++j) {
      float v = acc[i][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[wid][i * LF_BT + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < LF_OT * LF_BT) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    int i = threadIdx.x / LF_BT, j = threadIdx.x % LF_BT;
    if (o0 + i < O && b0 + j < B) {
      if (b) v += b[o0 + i];
      if (relu) v = fmaxf(v, 0.f);
      y[(long long)(b0 + j) * O + o0 + i] = v;
    }
  }
}
// y[b][o] = x[b][:] . w[o][:] for 8 outputs x 32 batch rows per block: warp w owns batch rows 4w..4w+3 against all 8 weight
// rows (the 8 warps read the same weight lines: one trip to L2, the rest L1 hits), lanes stride the reduction dimension
// with 128-bit loads.  The 32 per-lane partials are reduced with a transposing butterfly (31 shuffles instead of 160;
// lane l ends up owning sum l), fixed order -> deterministic.
constexpr int LF2_OT = 8, LF2_BT = 32;
__global__ void __launch_bounds__(256) k_linear_fwd_v2(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ y, int B, int F,
                                                       int O, int relu) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int o0 = blockIdx.x * LF2_OT, b0 = blockIdx.y * LF2_BT + wid * 4;
  float acc[LF2_OT * 4];
#pragma unroll
  for (int i = 0; i < LF2_OT * 4; ++i) acc[i] = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int f = lane * 4; f < F; f += 128) {
    float4 xv[4], wv[LF2_OT];
#pragma unroll
    for (int j = 0; j < 4; ++j) xv[j] = (b0 + j < B) ? __ldg(reinterpret_cast<const float4*>(x + (long long)(b0 + j) * F + f)) : z4;
#pragma unroll
    for (int i = 0; i < LF2_OT; ++i) wv[i] = (o0 + i < O) ? __ldg(reinterpret_cast<const float4*>(w + (long long)(o0 + i) * F + f)) : z4;
#pragma unroll
    for (int i = 0; i < LF2_OT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = acc[i * 4 + j];
        a = fmaf(wv[i].x, xv[j].x, a); a = fmaf(wv[i].y, xv[j].y, a);
        a = fmaf(wv[i].z, xv[j].z, a); a = fmaf(wv[i].w, xv[j].w, a);
        acc[i * 4 + j] = a;
      }
  }
  // transposing butterfly: after the step with mask m, a lane with bit m set carries the upper half of the index range
#pragma unroll
  for (int m = 16, n = 32; m > 0; m >>= 1, n >>= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = up ? acc[i + n / 2] : acc[i];
      const float send = up ? acc[i] : acc[i + n / 2];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  const int i = lane >> 2, j = lane & 3;           // lane l owns (output l/4, batch row l%4)
  if (o0 + i < O && b0 + j < B) {
    float v = acc[0];
    if (b) v += b[o0 + i];
    if (relu) v = fmaxf(v, 0.f);
    y[(long long)(b0 + j) * O + o0 + i] = v;
  }
}
void launch_linear_fwd(const float* x, const float* w, const float* b, float* y, int B, int F, int O, bool relu, cudaStream_t st) {
  g_launches += 1;
  // long reductions (encoder fc, F = 8192) stay on the many-small-blocks kernel: with 8 weight rows per block the v2 kernel
  // has only 4 KB of unique weight bytes in flight per SM per iteration and is latency-bound there (85 vs 68 us, cold L2);
  // short ones (decoder fc, F = z) are reduction-overhead-bound on the old kernel (31 us on v2)
  if (F <= 2048 && (F & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
    dim3 grid(cdiv(O, LF2_OT), cdiv(B, LF2_BT));
    k_linear_fwd_v2<<<grid, 256, 0, st>>>(x, w, b, y, B, F, O, relu ? 1 : 0);
    return;
  }
  dim3 grid(cdiv(O, LF_OT), cdiv(B, LF_BT));
  k_linear_fwd<<<grid, 256, 0, st>>>(x, w, b, y, B, F, O, relu ? 1 : 0);
}
constexpr int LD_BT = 16, LD_OC = 128;
// dx[B][F] = dy[B][O] . w[O][F]; a thread owns 4 consecutive columns f (128-bit, coalesced weight loads) for 16 batch rows;
// O is split over blockIdx.z in chunks of LD_OC (partials in scratch, reduced in fixed order -> deterministic)
__global__ void __launch_bounds__(256) k_linear_dgrad(const float* __restrict__ dy, const float* __restrict__ w,
                                                      float* __restrict__ part, int B, int F, int O) {
  __shared__ __align__(16) float sdy[LD_OC][LD_BT];   // [o][b]: the 16 batch values of one o are 4 x 128-bit broadcast reads
  const int b0 = blockIdx.y * LD_BT;
  const int f = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int o0 = blockIdx.z * LD_OC;
  const int oc = min(LD_OC, O - o0);
  for (int i = threadIdx.x; i < LD_BT * oc; i += 256) {
    int j = i / oc, o = i - j * oc;
    sdy[o][j] = (b0 + j < B) ? dy[(long long)(b0 + j) * O + o0 + o] : 0.f;
  }
  __syncthreads();
  if (f >= F) return;
  float4 acc[LD_BT];
#pragma unroll
  for (int j = 0; j < LD_BT; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* wp = w + (long long)o0 * F + f;
#pragma unroll 2
  for (int o = 0; o < oc; ++o) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + (long long)o * F));
#pragma unroll
    for (int j4 = 0; j4 < LD_BT / 4; ++j4) {
      const float4 d = *reinterpret_cast<const float4*>(&sdy[o][j4 * 4]);
      const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4& a = acc[j4 * 4 + q];
        a.x = fmaf(dv[q], wv.x, a.x); a.y = fmaf(dv[q], wv.y, a.y); a.z = fmaf(dv[q], wv.z, a.z); a.w = fmaf(dv[q], wv.w, a.w);
      }
    }
  }
  float* dst = part + (long long)blockIdx.z * B * F;
#pragma unroll
  for (int j = 0; j < LD_BT; ++j)
    if (b0 + j < B) *reinterpret_cast<float4*>(dst + (long long)(b0 + j) * F + f) = acc[j];
}
// generic fallback (F not a multiple of 4): one column per thread
__global__ void __launch_bounds__(256) k_linear_dgrad_scalar(const float* __restrict__ dy, const float* __restrict__ w,
                                                             float* __restrict__ part, int B, int F, int O) {
  __shared__ float sdy[LD_OC][LD_BT];
  const int b0 = blockIdx.y * LD_BT;
  const int f = blockIdx.x * 256 + threadIdx.x;
  const int o0 = blockIdx.z * LD_OC;
  const int oc = min(LD_OC, O - o0);
  for (int i = threadIdx.x; i < LD_BT * oc; i += 256) {
    int j = i / oc, o = i - j * oc;
    sdy[o][j] = (b0 + j < B) ? dy[(long long)(b0 + j) * O + o0 + o] : 0.f;
  }
  __syncthreads();
  if (f >= F) return;
  float acc[LD_BT];
#pragma unroll
  for (int j = 0; j < LD_BT; ++j) acc[j] = 0.f;
  for (int o = 0; o < oc; ++o) {
    float wv = __ldg(w + (long long)(o0 + o) * F + f);
#pragma unroll
    for (int j = 0; j < LD_BT; ++j) acc[j] = fmaf(sdy[o][j], wv, acc[j]);
  }
  float* dst = part + (long long)blockIdx.z * B * F;
#pragma unroll
  for (int j = 0; j < LD_BT; ++j)
    if (b0 + j < B) dst[(long long)(b0 + j) * F + f] = acc[j];
}
__global__ void k_linear_dgrad_reduce(const float* __restrict__ part, float* __restrict__ dx, long long n, int splits) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(long long)z * n + i];
    dx[i] = s;
  }
}
size_t linear_dgrad_scratch_bytes(int B, int F, int O) { return (size_t)cdiv(O, LD_OC) * B * F * sizeof(float); }
void launch_linear_dgrad(const float* dy, const float* w, float* dx, int B, int F, int O, void* scratch, cudaStream_t st) {
  g_launches += 2;
  const int splits = (int)cdiv(O, LD_OC);
  if ((F & 3) == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(scratch)) & 15) == 0) {
    dim3 grid(cdiv(F, 1024), cdiv(B, LD_BT), splits);
    k_linear_dgrad<<<grid, 256, 0, st>>>(dy, w, (float*)scratch, B, F, O);
  } else {
    dim3 grid(cdiv(F, 256), cdiv(B, LD_BT), splits);
    k_linear_dgrad_scalar<<<grid, 256, 0, st>>>(dy, w, (float*)scratch, B, F, O);
  }
  long long n = (long long)B * F;
  k_linear_dgrad_reduce<<<min(cdiv(n, 256), 148u * 4), 256, 0, st>>>((const float*)scratch, dx, n, splits);
}
constexpr int LW_OT = 8;
__global__ void __launch_bounds__(256) k_linear_wgrad(const float* __restrict__ x, const float* __restrict__ dy,
                                                      float* __restrict__ dw, float* __restrict__ db, int B, int F,
                                                      int O, int accumulate) {
  extern __shared__ float sdy[];   // [B][LW_OT]
  const int o0 = blockIdx.y * LW_OT;
  for (int i = threadIdx.x; i < B * LW_OT; i += 256) {
    int bb = i / LW_OT, oo = i - bb * LW_OT;
    sdy[i] = (o0 + oo < O) ? dy[(long long)bb * O + o0 + oo] : 0.f;
  }
  __syncthreads();
  int f = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x == 0 && db && threadIdx.x < LW_OT && o0 + threadIdx.x < O) {
    float s = 0.f;
    for (int bb = 0; bb < B; ++bb) s += sdy[bb * LW_OT + threadIdx.x];
    db[o0 + threadIdx.x] = accumulate ? db[o0 + threadIdx.x] + s : s;
  }
  if (f >= F) return;
  float acc[LW_OT];
#pragma unroll
  for (int i = 0; i < LW_OT; ++i) acc[i] = 0.f;
  for (int bb = 0; bb < B; ++bb) {
    float xv = __ldg(x + (long long)bb * F + f);
#pragma unroll
    for (int i = 0; i < LW_OT; ++i) acc[i] = fmaf(sdy[bb * LW_OT + i], xv, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < LW_OT; ++i)
    if (o0 + i < O) {
      long long idx = (long long)(o0 + i) * F + f;
      dw[idx] = accumulate ? dw[idx] + acc[i] : acc[i];
    }
}
void launch_linear_wgrad(const float* x, const float* dy, float* dw, float* db, int B, int F, int O, bool accumulate, cudaStream_t st) {
  g_launches += 1;
  dim3 grid(cdiv(F, 256), cdiv(O, LW_OT));
  size_t shmem = (size_t)B * LW_OT * sizeof(float);
  k_linear_wgrad<<<grid, 256, shmem, st>>>(x, dy, dw, db, B, F, O, accumulate ? 1 : 0);
}
__global__ void k_relu_bwd(const float* __restrict__ y, float* __restrict__ dy, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (!(y[i] > 0.f)) dy[i] = 0.f;
}
void launch_relu_bwd(const float* y, float* dy, long long n, cudaStream_t st) {
  g_launches += 1;
  if (n <= 0) return;
  k_relu_bwd<<<min(cdiv(n, 256), 148u * 8), 256, 0, st>>>(y, dy, n);
}

// =====================================================================================================
// Fused loss pass.  calc_reconstruction_loss('mse', reduction none) for the three image pairs of one half
// iteration in ONE read of the five images (:563,573,576 / :599,610,612), vectorised 128-bit loads,
// warp-shuffle + block reduction, deterministic two-stage sum (no atomics).
// =====================================================================================================
constexpr int MSE_CHUNK = 8192;   // floats per block
static int mse_blocks_per_sample(long long per_sample) { return (int)cdiv(per_sample, MSE_CHUNK); }
size_t mse3_scratch_bytes(int B, long long per_sample) {
  return (size_t)B * mse_blocks_per_sample(per_sample) * 3 * sizeof(float);
}
__device__ __forceinline__ float sq4(float4 a, float4 b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
  return fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
}
// per-element reconstruction error f(x = target, r = reconstruction) of calc_reconstruction_loss (:268-294) for the loss
// types that go through F.l1_loss / F.binary_cross_entropy (:288-291).  bce: ATen's kernel, (x - 1) max(log1p(-r), -100)
// - x max(log r, -100); a reconstruction outside [0, 1] is an error there (RuntimeError on CPU, device assert on CUDA): flagged.
template <int LT> __device__ __forceinline__ float rec_elem(float x, float r, int& bad) {
  if (LT == SIVAE_LOSS_L1_) return fabsf(r - x);
  if (!(r >= 0.f && r <= 1.f)) bad = 1;
  return (x - 1.f) * fmaxf(log1pf(-r), -100.f) - x * fmaxf(logf(r), -100.f);
}
template <int LT> __device__ __forceinline__ float rec_elem4(float4 x, float4 r, int& bad) {
  return (rec_elem<LT>(x.x, r.x, bad) + rec_elem<LT>(x.y, r.y, bad)) + (rec_elem<LT>(x.z, r.z, bad) + rec_elem<LT>(x.w, r.w, bad));
}
// LT = SIVAE_LOSS_MSE_: the squared error (the only type the reference's CLI passes); _L1_ / _BCE_: see rec_elem
template <int LT>
__global__ void __launch_bounds__(256) k_mse3_partial(const float* __restrict__ real, const float* __restrict__ rec,
                                                      const float* __restrict__ rec_rec, const float* __restrict__ fake,
                                                      const float* __restrict__ rec_fake, float* __restrict__ part,
                                                      long long per_sample, int nblk, int* __restrict__ bad_flag) {
  const int b = blockIdx.y;
  const long long base = (long long)b * per_sample;
  const long long e0 = (long long)blockIdx.x * MSE_CHUNK;
  const long long e1 = min(per_sample, e0 + MSE_CHUNK);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  int bad = 0;
  const bool vec = ((per_sample & 3) == 0);
  if (vec) {
    for (long long e = e0 + 4 * threadIdx.x; e < e1; e += 4 * 256) {
      float4 r = __ldg(reinterpret_cast<const float4*>(rec + base + e));
      if (LT == SIVAE_LOSS_MSE_) {
        if (real) s0 += sq4(r, __ldg(reinterpret_cast<const float4*>(real + base + e)));
        if (rec_rec) s1 += sq4(__ldg(reinterpret_cast<const float4*>(rec_rec + base + e)), r);
        if (fake) s2 += sq4(__ldg(reinterpret_cast<const float4*>(rec_fake + base + e)), __ldg(reinterpret_cast<const float4*>(fake + base + e)));
      } else {      // (target, reconstruction) pairs of :563/:599, :573/:610, :576/:612
        if (real) s0 += rec_elem4<LT>(__ldg(reinterpret_cast<const float4*>(real + base + e)), r, bad);
        if (rec_rec) s1 += rec_elem4<LT>(r, __ldg(reinterpret_cast<const float4*>(rec_rec + base + e)), bad);
        if (fake) s2 += rec_elem4<LT>(__ldg(reinterpret_cast<const float4*>(fake + base + e)), __ldg(reinterpret_cast<const float4*>(rec_fake + base + e)), bad);
      }
    }
  } else {
    for (long long e = e0 + threadIdx.x; e < e1; e += 256) {
      float r = rec[base + e];
      if (LT == SIVAE_LOSS_MSE_) {
        if (real) { float d = r - real[base + e]; s0 = fmaf(d, d, s0); }
        if (rec_rec) { float d = rec_rec[base + e] - r; s1 = fmaf(d, d, s1); }
        if (fake) { float d = rec_fake[base + e] - fake[base + e]; s2 = fmaf(d, d, s2); }
      } else {
        if (real) s0 += rec_elem<LT>(real[base + e], r, bad);
        if (rec_rec) s1 += rec_elem<LT>(r, rec_rec[base + e], bad);
        if (fake) s2 += rec_elem<LT>(fake[base + e], rec_fake[base + e], bad);
      }
    }
  }
  if (LT == SIVAE_LOSS_BCE_ && bad && bad_flag) *bad_flag = 1;      // every writer stores the same value
  __shared__ float red[8][3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[wid][0] = s0; red[wid][1] = s1; red[wid][2] = s2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    part[((long long)b * nblk + blockIdx.x) * 3 + threadIdx.x] = v;
  }
}
__global__ void k_mse3_final(const float* __restrict__ part, float* __restrict__ out, int B, int nblk) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;   // over B*3
  if (i >= B * 3) return;
  int b = i / 3, j = i - 3 * b;
  double s = 0.0;
  for (int k = 0; k < nblk; ++k) s += (double)part[((long long)b * nblk + k) * 3 + j];
  out[i] = (float)s;
}
// loss_type: SIVAE_LOSS_* (kernels.h); bad_flag (bce only, nullable): device int set to 1 if a reconstruction lies outside [0, 1]
void launch_mse3(const float* real, const float* rec, const float* rec_rec, const float* fake, const float* rec_fake,
                 float* out, int B, long long per_sample, void* scratch, size_t scratch_bytes, cudaStream_t st, int loss_type,
                 int* bad_flag) {
  g_launches += 2;
  int nblk = mse_blocks_per_sample(per_sample);
  float* part = (float*)scratch;
  dim3 grid(nblk, B);
  if (loss_type == SIVAE_LOSS_L1_)
    k_mse3_partial<SIVAE_LOSS_L1_><<<grid, 256, 0, st>>>(real, rec, rec_rec, fake, rec_fake, part, per_sample, nblk, nullptr);
  else if (loss_type == SIVAE_LOSS_BCE_) {
    if (bad_flag) cudaMemsetAsync(bad_flag, 0, sizeof(int), st);
    k_mse3_partial<SIVAE_LOSS_BCE_><<<grid, 256, 0, st>>>(real, rec, rec_rec, fake, rec_fake, part, per_sample, nblk, bad_flag);
  } else
    k_mse3_partial<SIVAE_LOSS_MSE_><<<grid, 256, 0, st>>>(real, rec, rec_rec, fake, rec_fake, part, per_sample, nblk, nullptr);
  k_mse3_final<<<cdiv(B * 3, 128), 128, 0, st>>>(part, out, B, nblk);
}

// calc_kl(reduce='none') (:231-251, mu_o = logvar_o = 0) + reparameterize (:254-265); one warp per sample
__global__ void k_kl_reparam(const float* __restrict__ ml, const float* __restrict__ eps, float* __restrict__ z,
                             float* __restrict__ kl, int B, int zd) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* mu = ml + (long long)b * 2 * zd;
  const float* lv = mu + zd;
  float s = 0.f;
  for (int j = lane; j < zd; j += 32) {
    float m = mu[j], l = lv[j];
    s += 1.f + l - m * m - expf(l);
    if (z) z[(long long)b * zd + j] = fmaf(eps[(long long)b * zd + j], expf(0.5f * l), m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0 && kl) kl[b] = -0.5f * s;
}
void launch_kl_reparam(const float* ml, const float* eps, float* z, float* kl, int B, int zd, cudaStream_t st) {
  g_launches += 1;
  k_kl_reparam<<<cdiv(B, 4), 128, 0, st>>>(ml, eps, z, kl, B, zd);
}
__global__ void k_latent_bwd(const float* __restrict__ ml, const float* __restrict__ eps, const float* __restrict__ dz,
                             const float* __restrict__ ckl, float ckl_const, float* __restrict__ dml, int B, int zd) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * zd) return;
  int b = (int)(i / zd), j = (int)(i - (long long)b * zd);
  float m = ml[(long long)b * 2 * zd + j], l = ml[(long long)b * 2 * zd + zd + j];
  float c = ckl ? ckl[b] : ckl_const;
  float dmu = c * m;
  float dlv = c * 0.5f * (expf(l) - 1.f);
  if (dz) {
    float d = dz[i];
    dmu += d;
    dlv += d * eps[i] * 0.5f * expf(0.5f * l);
  }
  dml[(long long)b * 2 * zd + j] = dmu;
  dml[(long long)b * 2 * zd + zd + j] = dlv;
}
void launch_latent_bwd(const float* ml, const float* eps, const float* dz, const float* ckl, float ckl_const,
                       float* dml, int B, int zd, cudaStream_t st) {
  g_launches += 1;
  long long n = (long long)B * zd;
  k_latent_bwd<<<cdiv(n, 256), 256, 0, st>>>(ml, eps, dz, ckl, ckl_const, dml, B, zd);
}

// block-wide sum of one value per thread (blockDim = 256), result valid in all threads
__device__ float block_sum_256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) r += sh[k];
  return r;
}
// E-step scalars (:563-586) and the per-sample backward coefficients of the exp-ELBO terms
__global__ void __launch_bounds__(256) k_e_loss_finalize(const float* __restrict__ mse, const float* __restrict__ kl_real,
                                                         const float* __restrict__ kl_rec, const float* __restrict__ kl_fake,
                                                         int B, float beta_kl, float beta_rec, float beta_neg, float scale,
                                                         float* stats, float* coef, float* ckl_rec, float* ckl_fake,
                                                         float md, const int* __restrict__ bad) {
  __shared__ float sh[8];
  float s_r = 0.f, s_k = 0.f, s_er = 0.f, s_ef = 0.f;
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += 256) {
    float r = mse[b * 3 + 0], rr = mse[b * 3 + 1], rf = mse[b * 3 + 2];
    float er = expf(-2.f * scale * (beta_rec * rr + beta_neg * kl_rec[b]));
    float ef = expf(-2.f * scale * (beta_rec * rf + beta_neg * kl_fake[b]));
    s_r += r; s_k += kl_real[b]; s_er += er; s_ef += ef;
    // d lossE / d rr_b = 0.25/B * er * (-2 scale beta_rec); image seed uses 2x that (d/dx of (x-y)^2)
    float c_rr = 0.25f * invB * er * (-2.f * scale * beta_rec);
    float c_rf = 0.25f * invB * ef * (-2.f * scale * beta_rec);
    coef[1 * B + b] = 2.f * c_rr;
    coef[2 * B + b] = 2.f * c_rf;
    ckl_rec[b] = 0.25f * invB * er * (-2.f * scale * beta_neg);
    ckl_fake[b] = 0.25f * invB * ef * (-2.f * scale * beta_neg);
  }
  // md: 1 for mse (mean over the batch of per-sample sums, :282-287); 1/D for l1 / bce (F.*_loss(reduction='mean'), :288-291)
  float loss_rec = block_sum_256(s_r, sh) * invB * md;
  float kl = block_sum_256(s_k, sh) * invB;
  float e_rec = block_sum_256(s_er, sh) * invB;
  float e_fake = block_sum_256(s_ef, sh) * invB;
  if (threadIdx.x == 0) {
    float lossE = scale * (beta_rec * loss_rec + beta_kl * kl) + 0.25f * (e_rec + e_fake);
    stats[0] = loss_rec; stats[1] = kl; stats[2] = e_rec; stats[3] = e_fake; stats[4] = lossE;
    stats[15] = (lossE != lossE) ? 1.f : 0.f;
    stats[14] = (bad && *bad) ? 1.f : 0.f;       // bce: reconstruction outside [0, 1]
  }
}
void launch_e_loss_finalize(const float* mse, const float* kl_real, const float* kl_rec, const float* kl_fake, int B,
                            float beta_kl, float beta_rec, float beta_neg, float scale, float* stats, float* coef,
                            float* ckl_rec, float* ckl_fake, cudaStream_t st, float mean_div, const int* bad_flag) {
  g_launches += 1;
  k_e_loss_finalize<<<1, 256, 0, st>>>(mse, kl_real, kl_rec, kl_fake, B, beta_kl, beta_rec, beta_neg, scale, stats, coef, ckl_rec, ckl_fake,
                                       mean_div, bad_flag);
}
__global__ void __launch_bounds__(256) k_d_loss_finalize(const float* __restrict__ mse, const float* __restrict__ kl_rec,
                                                         const float* __restrict__ kl_fake, int B, float beta_kl,
                                                         float beta_rec, float gamma_r, float scale, float* stats,
                                                         float md, const int* __restrict__ bad) {
  __shared__ float sh[8];
  float s_r = 0.f, s_rr = 0.f, s_rf = 0.f, s_kr = 0.f, s_kf = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    s_r += mse[b * 3]; s_rr += mse[b * 3 + 1]; s_rf += mse[b * 3 + 2]; s_kr += kl_rec[b]; s_kf += kl_fake[b];
  }
  const float invB = 1.f / (float)B;
  float loss_rec = block_sum_256(s_r, sh) * invB * md, lrr = block_sum_256(s_rr, sh) * invB * md, lrf = block_sum_256(s_rf, sh) * invB * md;
  float kr = block_sum_256(s_kr, sh) * invB, kf = block_sum_256(s_kf, sh) * invB;
  if (threadIdx.x == 0) {
    float lossD = scale * (loss_rec * beta_rec + (kr + kf) * 0.5f * beta_kl + gamma_r * 0.5f * beta_rec * (lrr + lrf));
    stats[5] = loss_rec; stats[6] = kr; stats[7] = kf; stats[8] = lrr; stats[9] = lrf; stats[10] = lossD;
    if (lossD != lossD) stats[15] = 1.f;
    if (bad && *bad) stats[14] = 1.f;
  }
}
void launch_d_loss_finalize(const float* mse, const float* kl_rec, const float* kl_fake, int B, float beta_kl,
                            float beta_rec, float gamma_r, float scale, float* stats, cudaStream_t st, float mean_div,
                            const int* bad_flag) {
  g_launches += 1;
  k_d_loss_finalize<<<1, 256, 0, st>>>(mse, kl_rec, kl_fake, B, beta_kl, beta_rec, gamma_r, scale, stats, mean_div, bad_flag);
}
__global__ void __launch_bounds__(256) k_vae_loss_finalize(const float* __restrict__ mse, const float* __restrict__ kl,
                                                           int B, float beta_kl, float beta_rec, float* stats, float md,
                                                           const int* __restrict__ bad) {
  __shared__ float sh[8];
  float s_r = 0.f, s_k = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) { s_r += mse[b * 3]; s_k += kl[b]; }
  const float invB = 1.f / (float)B;
  float lr = block_sum_256(s_r, sh) * invB * md, lk = block_sum_256(s_k, sh) * invB;
  if (threadIdx.x == 0) {
    float loss = beta_rec * lr + beta_kl * lk;
    stats[11] = lr; stats[12] = lk; stats[13] = loss;
    stats[15] = (loss != loss) ? 1.f : 0.f;
    stats[14] = (bad && *bad) ? 1.f : 0.f;
  }
}
void launch_vae_loss_finalize(const float* mse, const float* kl, int B, float beta_kl, float beta_rec, float* stats, cudaStream_t st,
                              float mean_div, const int* bad_flag) {
  g_launches += 1;
  k_vae_loss_finalize<<<1, 256, 0, st>>>(mse, kl, B, beta_kl, beta_rec, stats, mean_div, bad_flag);
}

// Half the derivatives of the per-element reconstruction error f(x = target, r = reconstruction) -- "half" because the
// coefficients a_* carry the factor 2 of d(r - x)^2 / dr (mse: g = r - x, gx = -g; both exact scalings):
//   l1  (F.l1_loss backward = sign):                g = sign(r - x) / 2,                      gx = -g
//   bce (ATen binary_cross_entropy_backward):       g = (r - x) / max((1 - r) r, 1e-12) / 2,  gx = -logit(r) / 2 (target grad)
template <int LT> __device__ __forceinline__ float rec_g(float x, float r) {
  if (LT == SIVAE_LOSS_L1_) { float d = r - x; return d > 0.f ? 0.5f : (d < 0.f ? -0.5f : 0.f); }
  if (LT == SIVAE_LOSS_BCE_) return 0.5f * (r - x) / fmaxf((1.f - r) * r, 1e-12f);
  return r - x;
}
template <int LT> __device__ __forceinline__ float rec_gx(float x, float r) {
  if (LT == SIVAE_LOSS_BCE_) return 0.5f * (log1pf(-r) - logf(r));
  return -rec_g<LT>(x, r);
}
template <int LT> __device__ __forceinline__ float4 rec_g4(float a, float4 x, float4 r) {
  return make_float4(a * rec_g<LT>(x.x, r.x), a * rec_g<LT>(x.y, r.y), a * rec_g<LT>(x.z, r.z), a * rec_g<LT>(x.w, r.w));
}
template <int LT> __device__ __forceinline__ float4 rec_gx4(float a, float4 x, float4 r) {
  return make_float4(a * rec_gx<LT>(x.x, r.x), a * rec_gx<LT>(x.y, r.y), a * rec_gx<LT>(x.z, r.z), a * rec_gx<LT>(x.w, r.w));
}
template <int LT>
__global__ void __launch_bounds__(256) k_loss_seed(const float* __restrict__ real, const float* __restrict__ rec,
                                                   const float* __restrict__ rec_rec, const float* __restrict__ fake,
                                                   const float* __restrict__ rec_fake, float a_rec,
                                                   const float* __restrict__ a_t_arr, float a_t_c,
                                                   const float* __restrict__ a_f_arr, float a_f_c, int tgt_rec,
                                                   float* __restrict__ d_rec, float* __restrict__ d_rec_rec,
                                                   float* __restrict__ d_rec_fake, float* __restrict__ d_fake,
                                                   long long per_sample4) {
  const int b = blockIdx.y;
  const float a_t = a_t_arr ? a_t_arr[b] : a_t_c;
  const float a_f = a_f_arr ? a_f_arr[b] : a_f_c;
  const long long base = (long long)b * per_sample4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample4; i += (long long)gridDim.x * blockDim.x) {
    float4 r = __ldg(reinterpret_cast<const float4*>(rec) + base + i);
    float4 x = __ldg(reinterpret_cast<const float4*>(real) + base + i);
    if (LT == SIVAE_LOSS_MSE_) {
      float4 o = make_float4(a_rec * (r.x - x.x), a_rec * (r.y - x.y), a_rec * (r.z - x.z), a_rec * (r.w - x.w));
      if (rec_rec) {
        float4 q = __ldg(reinterpret_cast<const float4*>(rec_rec) + base + i);
        float4 d = make_float4(a_t * (q.x - r.x), a_t * (q.y - r.y), a_t * (q.z - r.z), a_t * (q.w - r.w));
        if (d_rec_rec) reinterpret_cast<float4*>(d_rec_rec)[base + i] = d;
        if (tgt_rec) { o.x -= d.x; o.y -= d.y; o.z -= d.z; o.w -= d.w; }
      }
      if (d_rec) reinterpret_cast<float4*>(d_rec)[base + i] = o;
      if (fake) {
        float4 f = __ldg(reinterpret_cast<const float4*>(fake) + base + i);
        float4 q = __ldg(reinterpret_cast<const float4*>(rec_fake) + base + i);
        float4 d = make_float4(a_f * (q.x - f.x), a_f * (q.y - f.y), a_f * (q.z - f.z), a_f * (q.w - f.w));
        if (d_rec_fake) reinterpret_cast<float4*>(d_rec_fake)[base + i] = d;
        if (d_fake) reinterpret_cast<float4*>(d_fake)[base + i] = make_float4(-d.x, -d.y, -d.z, -d.w);
      }
    } else {
      float4 o = rec_g4<LT>(a_rec, x, r);                            // d / d rec of f(real, rec)
      if (rec_rec) {
        float4 q = __ldg(reinterpret_cast<const float4*>(rec_rec) + base + i);
        if (d_rec_rec) reinterpret_cast<float4*>(d_rec_rec)[base + i] = rec_g4<LT>(a_t, r, q);     // f(rec, rec_rec) w.r.t. rec_rec
        if (tgt_rec) { float4 t = rec_gx4<LT>(a_t, r, q); o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }   // ... w.r.t. its target rec
      }
      if (d_rec) reinterpret_cast<float4*>(d_rec)[base + i] = o;
      if (fake) {
        float4 f = __ldg(reinterpret_cast<const float4*>(fake) + base + i);
        float4 q = __ldg(reinterpret_cast<const float4*>(rec_fake) + base + i);
        if (d_rec_fake) reinterpret_cast<float4*>(d_rec_fake)[base + i] = rec_g4<LT>(a_f, f, q);
        if (d_fake) reinterpret_cast<float4*>(d_fake)[base + i] = rec_gx4<LT>(a_f, f, q);
      }
    }
  }
}
void launch_loss_seed(const float* real, const float* rec, const float* rec_rec, const float* fake, const float* rec_fake,
                      float a_rec, const float* a_t_arr, float a_t, const float* a_f_arr, float a_f, bool target_grad_rec,
                      float* d_rec, float* d_rec_rec, float* d_rec_fake, float* d_fake, int B, long long per_sample,
                      cudaStream_t st, int loss_type) {
  g_launches += 1;
  long long ps4 = per_sample / 4;   // per_sample = cdim*S*S with even S: multiple of 4
  dim3 grid(min(cdiv(ps4, 256), 148u * 4), B);
  if (loss_type == SIVAE_LOSS_L1_)
    k_loss_seed<SIVAE_LOSS_L1_><<<grid, 256, 0, st>>>(real, rec, rec_rec, fake, rec_fake, a_rec, a_t_arr, a_t, a_f_arr, a_f,
                                                      target_grad_rec ? 1 : 0, d_rec, d_rec_rec, d_rec_fake, d_fake, ps4);
  else if (loss_type == SIVAE_LOSS_BCE_)
    k_loss_seed<SIVAE_LOSS_BCE_><<<grid, 256, 0, st>>>(real, rec, rec_rec, fake, rec_fake, a_rec, a_t_arr, a_t, a_f_arr, a_f,
                                                       target_grad_rec ? 1 : 0, d_rec, d_rec_rec, d_rec_fake, d_fake, ps4);
  else
    k_loss_seed<SIVAE_LOSS_MSE_><<<grid, 256, 0, st>>>(real, rec, rec_rec, fake, rec_fake, a_rec, a_t_arr, a_t, a_f_arr, a_f,
                                                       target_grad_rec ? 1 : 0, d_rec, d_rec_rec, d_rec_fake, d_fake, ps4);
}

// =====================================================================================================
// torch.optim.Adam (:450-451): betas (.9,.999), eps 1e-8, bias-corrected, no weight decay -- flat buffers
// =====================================================================================================
// coef[0] = lr / (1 - b1^step), coef[1] = sqrt(1 - b2^step); the step counter lives on the device so that a captured
// CUDA graph of the iteration replays with the correct bias correction
__global__ void k_adam_prep(long long* step, float lr, float b1, float b2, float* coef) {
  long long t = *step + 1;
  *step = t;
  double bc1 = 1.0 - pow((double)b1, (double)t);
  double bc2 = 1.0 - pow((double)b2, (double)t);
  coef[0] = (float)((double)lr / bc1);
  coef[1] = (float)sqrt(bc2);
}
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long long n, const float* __restrict__ coef,
                                              float grad_scale, float b1, float b2, float eps) {
  const float step_size = coef[0], sqrt_bc2 = coef[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    float mi = m[i] + (gi - m[i]) * (1.f - b1);            // exp_avg.lerp_(grad, 1-beta1)
    float vi = fmaf(gi * gi, 1.f - b2, v[i] * b2);         // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / sqrt_bc2 + eps;             // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p[i] = p[i] - step_size * (mi / denom);
  }
}
// step_dev: device int64 counter (incremented here); coef_dev: 2 floats of device scratch
void launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float grad_scale, float b1,
                 float b2, float eps, long long* step_dev, float* coef_dev, cudaStream_t st) {
  g_launches += 2;
  if (n <= 0) return;
  k_adam_prep<<<1, 1, 0, st>>>(step_dev, lr, b1, b2, coef_dev);
  k_adam<<<min(cdiv(n, 256), 148u * 16), 256, 0, st>>>(p, g, m, v, n, coef_dev, grad_scale, b1, b2, eps);
}

}  // namespace sivae
