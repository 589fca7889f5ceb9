// tcgen05 (5th-gen tensor core) TF32 implicit-GEMM convolution for sm_100a.
//
//   forward / dgrad :  D[M = 128 pixels][N = Cout tile] = sum_{tap, ci}  X[pixel + tap][ci] * Wp[co][tap][ci]
//       A tile  = 128 pixels x 32 channels (one filter tap), fetched by ONE 4-D TMA box {32c, bw, bh, bn} of the
//                 NHWC activation at the tap-shifted coordinate; the zero padding of the convolution is the TMA
//                 out-of-bounds fill, so there is no im2col buffer and no bounds logic in the kernel.
//       B tile  = BLOCK_N filters x 32 channels, 2-D TMA box of the packed filter matrix [Cout][kh*kw*Cin].
//       both K-major, 128-byte swizzle; UMMA 128 x BLOCK_N x 8 (kind::tf32), fp32 accumulator in TMEM.
//   wgrad           :  D[M = 128 (tap,ci)][N = Cout tile] = sum_pixels X[pixel + tap][ci] * dY[pixel][co]
//       the reduction runs over pixels, which is the SLOW dimension of both NHWC operands, so both operands are
//       MN-major: A = four {32c x 32 pixel} TMA boxes (tap-shifted, OOB zero fill), B = BLOCK_N/32 boxes of dY.
//       split over pixel ranges, partial tiles reduced in fixed order (deterministic).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> registers -> global).  smem ring of STAGES {A,B} slots with full/empty
// mbarriers; tcgen05.commit releases slots and signals the epilogue.
// Operands are expected to be pre-rounded to TF32 by their producers (the tensor core truncates fp32 -> tf32).
#include "kernels.h"
#include "split32.cuh"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace sivae {

// ---------------------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// NHWC tensor [N][H][W][C] -> 4-D map, box {128 bytes of channels, bw, bh, bn}, 128B swizzle, OOB -> 0.
// bf16 = false: fp32 elements (32 channels per box row; also the split32 format, which has the same byte geometry);
// bf16 = true: plain bf16 elements (64 channels per box row)
static int make_map_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, bool bf16 = false) {
  EncodeTiledFn f = encode_fn();
  if (!f) return -101;
  const cuuint64_t es_b = bf16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * es_b, (cuuint64_t)W * C * es_b, (cuuint64_t)H * W * C * es_b};
  cuuint32_t box[4] = {bf16 ? 64u : 32u, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = f(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides,
                 box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -102;
}
static int make_store_map(CUtensorMap* m, float* base, int N, int H, int W, int C, int sw, int sh, int sn) {
  return make_map_nhwc(m, base, N, H, W, C, sw, sh, sn);
}
// row-major matrix [rows][cols] (fp32 / split32, or plain bf16) -> 2-D map, box {128 bytes, box_rows}
static int make_map_2d(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows, bool bf16 = false) {
  EncodeTiledFn f = encode_fn();
  if (!f) return -101;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {bf16 ? 64u : 32u, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = f(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides,
                 box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -102;
}

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  if (mbar_try_wait(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(addr, parity)) {
    if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz: far beyond any legitimate wait in these kernels
      printf("sivae: mbarrier wait timed out (block %d,%d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives row (lane base + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// Staging tiles per epilogue warp.  With ONE tile every chunk waited for the previous TMA store to finish reading it:
// ~3.3k clk per 32x32 chunk, which made the epilogue (not the MMA loop) the bottleneck of every output-heavy layer
// (profiles/r01h_layers_H_rowsep_wgrad.md: 1x1 convs at 1.4 TB/s of output, 64->64 3x3 at 15k clk per 256-pixel item).
constexpr int EPI_NBUF = 2;
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Epilogue of one 32-row x 32-column accumulator chunk held by a warp (lane = row): + bias + addend, then either
//  (a) staged through a 4 KB shared tile (128B-swizzled, conflict-free 16-byte stores) and written by ONE TMA store of the
//      box {32 ch, bw, bh, bn} -- full-line coalesced writes, batch tail clipped by the tensor map; or
//  (b) direct per-thread stores (narrow outputs whose channel count is not a multiple of 4, e.g. the 3-channel image).
// Path (a) exists because direct stores (32 rows x 16 B per instruction, each its own L2 transaction) capped every
// output-heavy layer at ~1.1 TB/s (profiles/r01c_layers_H_halo.md: 1x1 convs 35-126 TFLOP/s, 64-channel 3x3 at 0.45 ms).
// explicit shared-space accesses for the staging tiles: their addresses are derived from a 1024B-aligned (integer
// round-up) base, so the compiler no longer knows the state space and would emit GENERIC ld/st (ST.E.128 / LD.E in the
// SASS of the first versions) for every staging access
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}
struct EpiOut { const float* bias; const float* addend; float* y; int Cout; float* stats; int amode; };   // stats: BN partials base or null
struct EpiState { uint32_t n = 0; bool pref = false; };     // per-warp count of staged chunks (selects the staging tile);
                                                             // pref: the addend of chunk n is already on its way (mode 3)
// Addend mode 3: the addend tile of a chunk is fetched by TMA straight into the staging tile that chunk will use (same
// box and swizzle as the output store), one chunk ahead -- no registers, no exposed global-load latency.  abar = this
// warp's EPI_NBUF mbarriers.  Issued right after the previous chunk's store (or, for the first chunk of an item, before
// the accumulator is ready).
__device__ __forceinline__ void epi_prefetch(const float* addend, int amode, EpiState& es, uint8_t* stage0, uint64_t* abar,
                                             const CUtensorMap* map_add, int col, int cw, int ch, int cn, int lane) {
  if (!addend || (amode & 3) != 3 || es.pref) return;
  const uint32_t b = es.n % EPI_NBUF;
  if (lane == 0) {
    bulk_wait_read<EPI_NBUF - 1>();       // the store that last used this tile has finished reading it
    mbar_expect_tx(&abar[b], 4096);
    tma_load_4d(stage0 + b * 4096, map_add, &abar[b], col, cw, ch, cn);
  }
  es.pref = true;
}
// srow >= 0 (TMA path only): also emit the per-channel sum / sum of squares of this warp's 32 rows (train-mode BatchNorm
// statistics, fused so that the conv output is not re-read): stats[srow][0][c] = sum, stats[srow][1][c] = sum of squares
// addend (residual / accumulated gradient; may alias the output) of one 32x32 chunk: COALESCED loads -- instruction j covers
// rows 4j..4j+3 of this warp's 32 rows (lane -> row 4j + lane/8, 16-byte piece lane%8 = four full 128 B lines per
// instruction); epi_chunk passes them through the staging tile.  The earlier lane-per-row loads (32 lines x 16 B per
// instruction) made every dgrad-with-addend launch ~2.5x slower than the same conv without it, and loading inside epi_chunk
// still exposed one global-load latency per chunk (profiles/r01i_probe_conv_bw.txt): callers prefetch chunk i+1 here
// before they process chunk i.
__device__ __forceinline__ void epi_load_addend(const EpiOut& o, long long pix, bool valid, int col, int lane, float4 (&a)[8]) {
  if (!o.addend) return;
  const long long mypix = valid ? pix : -1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long prow = __shfl_sync(0xffffffffu, mypix, 4 * j + (lane >> 3));
    const int cc = col + ((lane & 7) << 2);
    a[j] = (prow >= 0 && cc < o.Cout) ? *reinterpret_cast<const float4*>(o.addend + prow * o.Cout + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void epi_chunk(uint32_t (&v)[32], const EpiOut& o, long long pix, bool valid, int col, bool tma,
                                          uint8_t* stage0, EpiState& es, const CUtensorMap* map_y, int cw, int ch, int cn, int lane,
                                          long long srow = -1, uint64_t* abar = nullptr, const CUtensorMap* map_add = nullptr) {
  if (tma) {
    const bool tma_add = o.addend && (o.amode & 3) == 3;
    if (tma_add) epi_prefetch(o.addend, o.amode, es, stage0, abar, map_add, col, cw, ch, cn, lane);     // no-op if already requested
    const uint32_t buf = es.n % EPI_NBUF, par = (es.n / EPI_NBUF) & 1;
    uint8_t* stage = stage0 + buf * 4096;
    const uint32_t stage_s = smem_u32(stage);
    ++es.n;
    es.pref = false;
    float4 al[8];
    const bool staged_add = o.addend && (o.amode & 3) != 0;
    if (o.addend && (o.amode & 3) == 1) epi_load_addend(o, pix, valid, col, lane, al);      // coalesced, not prefetched
    if (tma_add) {
      mbar_wait(&abar[buf], par);                    // the addend tile has landed in this staging tile
    } else {
      if (lane == 0) bulk_wait_read<EPI_NBUF - 1>();   // the store issued EPI_NBUF chunks ago has finished reading this tile
      __syncwarp();
    }
    if (staged_add && !tma_add) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = 4 * j + (lane >> 3);
        sts128(stage_s + row * 128 + (((lane & 7) ^ (row & 7)) << 4), al[j]);
      }
      __syncwarp();
    }
    const float* add0 = (o.addend && (o.amode & 3) == 0 && valid) ? o.addend + pix * o.Cout + col : nullptr;   // mode 0: lane-per-row loads
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 q = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      const uint32_t slot = stage_s + lane * 128 + ((j ^ (lane & 7)) << 4);
      if (col + 4 * j < o.Cout) {
        if (o.bias) {
          float4 b = __ldg(reinterpret_cast<const float4*>(o.bias + col + 4 * j));
          q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
        }
        if (staged_add) {
          const float4 d = lds128(slot);
          q.x += d.x; q.y += d.y; q.z += d.z; q.w += d.w;
        } else if (add0) {
          const float4 d = *reinterpret_cast<const float4*>(add0 + 4 * j);
          q.x += d.x; q.y += d.y; q.z += d.z; q.w += d.w;
        }
      }
      sts128(slot, q);
    }
    if (!(o.amode & 8)) fence_proxy_async();      // bit 3 of amode: timing experiment only (SIVAE_TC_NOFENCE=1, results invalid)
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(map_y, stage, col, cw, ch, cn);
      bulk_commit();
    }
    if (o.stats && srow >= 0) {
      // lane = column: walk the 32 staged rows (swizzled 16-byte chunks -> conflict-free)
      float sm = 0.f, sq = 0.f;
      const int cj = lane >> 2, ce = (lane & 3) << 2;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const float x = lds32(stage_s + r * 128 + ((cj ^ (r & 7)) << 4) + ce);
        sm += x;
        sq = fmaf(x, x, sq);
      }
      if (col + lane < o.Cout) {
        float* d = o.stats + srow * 2 * o.Cout + col + lane;
        d[0] = sm;
        d[o.Cout] = sq;
      }
    }
    return;
  }
  if (!valid) return;
  float* dst = o.y + pix * o.Cout + col;
  const float* add = o.addend ? o.addend + pix * o.Cout + col : nullptr;
  if ((o.Cout & 3) != 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (col + j < o.Cout) {
        float x = __uint_as_float(v[j]);
        if (o.bias) x += __ldg(o.bias + col + j);
        if (add) x += add[j];
        dst[j] = x;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (col + j < o.Cout) {
        float4 q = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        if (o.bias) {
          float4 b = __ldg(reinterpret_cast<const float4*>(o.bias + col + j));
          q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
        }
        if (add) {
          float4 a = *reinterpret_cast<const float4*>(add + j);
          q.x += a.x; q.y += a.y; q.z += a.z; q.w += a.w;
        }
        *reinterpret_cast<float4*>(dst + j) = q;
      }
    }
  }
}
// Straight-line epilogue of a FULL 32x32 chunk on the TMA-store path: no bias, all 32 columns valid, addend either absent
// or TMA-prefetched into the staging tile (mode 3).  The generic epi_chunk above spends ~2.9k clk per chunk on the
// 64-channel layers (336 SASS instructions, a runtime branch and a constant-bank reload per 16-byte piece, instruction
// fetch stalls from jumping over the unused modes) and is the critical path there: the epilogue warps never wait for an
// accumulator while the tensor pipe idles at 43 % (profiles/r01o_prof_halo2.md).  Same arithmetic, same staging layout.
// Statistics: lane (g = lane/8, p = lane%8) sums the 16-byte piece p of rows 8g..8g+7 (conflict-free under the swizzle),
// two xor-shuffles fold the four row groups, lanes 0..7 write one float4 of sums and one of squares.
template <bool ADD, bool STATS>
__device__ __forceinline__ void epi_chunk_fast(uint32_t (&v)[32], const EpiOut& o, int col, uint8_t* stage0, EpiState& es,
                                               const CUtensorMap* map_y, int cw, int ch, int cn, int lane, long long srow,
                                               uint64_t* abar, const CUtensorMap* map_add) {
  if (ADD) epi_prefetch(o.addend, 3, es, stage0, abar, map_add, col, cw, ch, cn, lane);     // no-op if already requested
  const uint32_t buf = es.n % EPI_NBUF, par = (es.n / EPI_NBUF) & 1;
  uint8_t* stage = stage0 + buf * 4096;
  const uint32_t stage_s = smem_u32(stage);
  ++es.n;
  es.pref = false;
  if (ADD) {
    mbar_wait(&abar[buf], par);                      // the addend tile has landed in this staging tile
  } else {
    bulk_wait_read<EPI_NBUF - 1>();                  // per-thread groups: a no-op for the lanes that never issue a store
    __syncwarp();
  }
  const uint32_t mine = stage_s + lane * 128;
  const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 q = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    const uint32_t slot = mine + ((j ^ sw) << 4);
    if (ADD) {
      const float4 d = lds128(slot);
      q.x += d.x; q.y += d.y; q.z += d.z; q.w += d.w;
    }
    sts128(slot, q);
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(map_y, stage, col, cw, ch, cn);
    bulk_commit();
  }
  if (STATS) {
    const uint32_t g = (uint32_t)lane >> 3, pc = (uint32_t)lane & 7;
    const uint32_t rbase = stage_s + g * 1024;       // rows 8g .. 8g+7
    float4 sm = make_float4(0.f, 0.f, 0.f, 0.f), sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = lds128(rbase + i * 128 + ((pc ^ (uint32_t)i) << 4));
      sm.x += x.x; sm.y += x.y; sm.z += x.z; sm.w += x.w;
      sq.x = fmaf(x.x, x.x, sq.x); sq.y = fmaf(x.y, x.y, sq.y); sq.z = fmaf(x.z, x.z, sq.z); sq.w = fmaf(x.w, x.w, sq.w);
    }
#pragma unroll
    for (int o2 = 8; o2 <= 16; o2 <<= 1) {
      sm.x += __shfl_xor_sync(0xffffffffu, sm.x, o2); sm.y += __shfl_xor_sync(0xffffffffu, sm.y, o2);
      sm.z += __shfl_xor_sync(0xffffffffu, sm.z, o2); sm.w += __shfl_xor_sync(0xffffffffu, sm.w, o2);
      sq.x += __shfl_xor_sync(0xffffffffu, sq.x, o2); sq.y += __shfl_xor_sync(0xffffffffu, sq.y, o2);
      sq.z += __shfl_xor_sync(0xffffffffu, sq.z, o2); sq.w += __shfl_xor_sync(0xffffffffu, sq.w, o2);
    }
    if (g == 0) {
      float* d = o.stats + srow * 2 * o.Cout + col + 4 * (int)pc;
      *reinterpret_cast<float4*>(d) = sm;
      *reinterpret_cast<float4*>(d + o.Cout) = sq;
    }
  }
}
// epilogue variant of a launch (warp-uniform, fixed for the whole kernel): 0 = generic epi_chunk; otherwise
// 1 | ADD << 1 | STATS << 2 selects epi_chunk_fast<ADD, STATS>.  SIVAE_TC_FASTEPI=0 (host side) forces 0.
__device__ __forceinline__ int epi_variant(const EpiOut& o, bool tma, int fast_ok) {
  if (!fast_ok || !tma || o.bias || (o.Cout & 31) != 0) return 0;
  if (o.addend && (o.amode & 3) != 3) return 0;
  return 1 | (o.addend ? 2 : 0) | (o.stats ? 4 : 0);
}
template <int EM>
__device__ __forceinline__ void epi_do(uint32_t (&v)[32], const EpiOut& o, long long pix, bool valid, int col, bool tma,
                                       uint8_t* stage0, EpiState& es, const CUtensorMap* map_y, int cw, int ch, int cn, int lane,
                                       long long srow, uint64_t* abar, const CUtensorMap* map_add) {
  if (EM == 0) epi_chunk(v, o, pix, valid, col, tma, stage0, es, map_y, cw, ch, cn, lane, srow, abar, map_add);
  else epi_chunk_fast<(EM & 2) != 0, (EM & 4) != 0>(v, o, col, stage0, es, map_y, cw, ch, cn, lane, srow, abar, map_add);
}
// run `body(std::integral_constant<int, EM>)` for the launch's epilogue variant
template <class F>
__device__ __forceinline__ void epi_dispatch(int em, F&& body) {
  switch (em) {
    case 1: body(std::integral_constant<int, 1>{}); break;
    case 3: body(std::integral_constant<int, 3>{}); break;
    case 5: body(std::integral_constant<int, 5>{}); break;
    case 7: body(std::integral_constant<int, 7>{}); break;
    default: body(std::integral_constant<int, 0>{}); break;
  }
}
// shared-memory matrix descriptor (sm_100 UMMA): start address, leading / stride byte offsets (>>4), version 1,
// layout type 2 = SWIZZLE_128B (16-byte swizzle atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms: the only layout the
// tensor core accepts for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor: D fp32, A/B tf32, majors, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with bf16 A/B operands (format code 1): K = 16 elements = the same 32 bytes per MMA as 8 tf32 values
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate);
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate);

// ---------------------------------------------------------------------------------------------------------------
// Operand formats of the forward / dgrad kernels (template parameter FMT).  Both formats occupy the SAME bytes: one
// 128-byte swizzle row = one pixel (or filter row) x 32 channels.
//   FMT_TF32  : 32 fp32 values, pre-rounded to tf32; 4 MMAs (K = 8) of kind::tf32 per row.
//   FMT_SPLIT : "split32" = [32 x bf16 hi | 32 x bf16 lo] with hi = bf16(x), lo = bf16(x - hi) (16 significand bits, fp32
//               exponent range).  Per 16-channel K step three kind::f16 MMAs: lo*hi + hi*lo + hi*hi (the lo*lo term, 2^-18
//               relative, is dropped): 6 MMAs per row at twice the tf32 issue rate = 1.5x the tf32 tensor time for
//               fp32-class products.  TMA maps, staging bytes, descriptors' row geometry and epilogues are identical.
//   FMT_BF16  : 64 plain bf16 values (one 128-byte row = 64 channels); 4 MMAs (K = 16) of kind::f16 per row = half the tf32
//               tensor time per channel.  Used by the dgrad of the residual blocks: its operand rounding (8 significand
//               bits) moves the gradients by less than the forward's own round-off does (profiles/r02a_split_formats.md).
// a_addr / b_addr: shared-memory byte address of the first row of the operand tile (row start, K offset 0).
// ---------------------------------------------------------------------------------------------------------------
constexpr int FMT_TF32 = 0, FMT_SPLIT = 1, FMT_BF16 = 2;
template <int FMT> __host__ __device__ constexpr uint32_t make_idesc_fmt(int M, int N) {
  return FMT == FMT_TF32 ? make_idesc_tf32(M, N, 0, 0) : make_idesc_bf16(M, N, 0, 0);
}
template <int FMT> __host__ __device__ constexpr int fmt_csh() { return FMT == FMT_BF16 ? 6 : 5; }   // log2(channels per 128-byte row)
__device__ __forceinline__ uint64_t make_smem_desc_bo(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset);
template <int FMT, bool PAIR>
__device__ __forceinline__ void umma_row(uint32_t tmem_d, uint32_t a_addr, uint32_t a_sbo, uint32_t a_bo, uint32_t b_addr,
                                         uint32_t idesc, bool fresh) {
  auto mma = [&](uint32_t ao, uint32_t bo, uint32_t acc) {
    const uint64_t ad = make_smem_desc_bo(a_addr + ao, a_sbo, a_bo);
    const uint64_t bd = make_smem_desc_bo(b_addr + bo, 1024, 0);
    if constexpr (FMT != FMT_TF32) {
      if constexpr (PAIR) umma_f16_2cta(tmem_d, ad, bd, idesc, acc); else umma_f16(tmem_d, ad, bd, idesc, acc);
    } else {
      if constexpr (PAIR) umma_tf32_2cta(tmem_d, ad, bd, idesc, acc); else umma_tf32(tmem_d, ad, bd, idesc, acc);
    }
  };
  if constexpr (FMT == FMT_SPLIT) {
#pragma unroll
    for (uint32_t kk = 0; kk < 2; ++kk) {
      mma(64 + kk * 32, kk * 32, (fresh && kk == 0) ? 0u : 1u);      // a_lo * b_hi
      mma(kk * 32, 64 + kk * 32, 1u);                               // a_hi * b_lo
      mma(kk * 32, kk * 32, 1u);                                    // a_hi * b_hi
    }
  } else {
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) mma(k * 32, k * 32, (fresh && k == 0) ? 0u : 1u);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ---------------------------------------------------------------------------------------------------------------
struct FwdParams {
  int N, H, W, Cin, Cout, ks;
  int bw, bh, bn;            // pixel tile = bn x bh x bw = 128
  int tiles_w, tiles_h;      // W/bw, H/bh
  int sbh, sbn;              // TMA-store box of one warp's 32 rows: {32 ch, bw, sbh, sbn}
  int tma_store;             // 1: epilogue through shared memory + TMA store; 0: direct stores
  const float* bias;
  const float* addend;
  float* y;
  float* stats;              // BN partial sums [m_tiles*4][2][Cout] or null
  int ksplit, kb_per, npad;  // split-K (v2 kernel): k-blocks per split; partial tensor batch stride (tiles_n * bn)
};
constexpr int TC_A_BYTES = 128 * 128;       // 128 rows x 32 fp32

template <int BLOCK_N, int STAGES>
struct FwdSmem {
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(192) k_conv_fwd_tc(const __grid_constant__ CUtensorMap map_x,
                                                     const __grid_constant__ CUtensorMap map_w, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = FwdSmem<BLOCK_N, STAGES>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, tn = tile / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
  const int col0 = blockIdx.y * BLOCK_N;
  const int cchunks = p.Cin >> 5;
  const int num_kb = p.ks * p.ks * cchunks;
  const int pad = p.ks >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], SM::STAGE_BYTES);
        const int tap = kb / cchunks, c0 = (kb - tap * cchunks) << 5;
        const int r = tap / p.ks, s = tap - r * p.ks;
        uint8_t* sa = smem + st * SM::STAGE_BYTES;
        tma_load_4d(sa, &map_x, &full[st], c0, w0 + s - pad, h0 + r - pad, n0);
        tma_load_2d(sa + TC_A_BYTES, &map_w, &full[st], tap * p.Cin + c0, col0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(128, BLOCK_N, 0, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int st = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + st * SM::STAGE_BYTES);
        const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // K-major, 128B swizzle: rows 128 B apart, 8-row groups 1024 B apart; advance 8 tf32 (32 B) per MMA
          uint64_t ad = make_smem_desc(sa + k * 32, 0, 1024);
          uint64_t bd = make_smem_desc(sb + k * 32, 0, 1024);
          umma_tf32(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);                      // slot free once these MMAs have read it
        if (kb == num_kb - 1) umma_commit(tmem_full); // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quarter (warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dw = row % p.bw, dh = (row / p.bw) % p.bh, dn = row / (p.bw * p.bh);
    const int n = n0 + dn;
    const bool valid = n < p.N;
    const long long pix = ((long long)n * p.H + (h0 + dh)) * p.W + (w0 + dw);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (valid && (p.Cout & 3) != 0) {
        // narrow output (e.g. the 3-channel predict layer / stem dgrad): scalar stores of the first Cout columns
        float* dst = p.y + pix * p.Cout + col0 + c;
        const float* add = p.addend ? p.addend + pix * p.Cout + col0 + c : nullptr;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (col0 + c + j < p.Cout) {
            float o = __uint_as_float(v[j]);
            if (p.bias) o += __ldg(p.bias + col0 + c + j);
            if (add) o += add[j];
            dst[j] = o;
          }
        }
      } else if (valid) {
        float* dst = p.y + pix * p.Cout + col0 + c;
        const float* add = p.addend ? p.addend + pix * p.Cout + col0 + c : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (col0 + c + j < p.Cout) {
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (p.bias) {
              float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c + j));
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (add) {
              float4 a = *reinterpret_cast<const float4*>(add + j);
              o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            *reinterpret_cast<float4*>(dst + j) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc<BLOCK_N>(tmem_base); }
}

static void pick_tile(int H, int W, int* bw, int* bh, int* bn) {
  int w = 1;
  while (w * 2 <= 16 && W % (w * 2) == 0) w *= 2;
  int h = 1;
  while (h * 2 <= 128 / w && H % (h * 2) == 0) h *= 2;
  *bw = w; *bh = h; *bn = 128 / (w * h);
}

bool conv_tc_supported_fwd(const ConvShape& s, int fmt) {
  // Cout need not fill a UMMA tile: filter rows beyond Cout are TMA out-of-bounds zeros and the epilogue masks them
  if (s.Cin % (fmt == FMT_BF16 ? 64 : 32) != 0 || s.Cout < 1) return false;
  if (s.k != 1 && s.k != 3 && s.k != 5) return false;
  int bw, bh, bn;
  pick_tile(s.H, s.W, &bw, &bh, &bn);
  if (bn > 256) return false;
  return true;
}

template <int BLOCK_N, int STAGES>
static int launch_fwd_t(const CUtensorMap& mx, const CUtensorMap& mw, const FwdParams& p, int m_tiles, cudaStream_t st) {
  using SM = FwdSmem<BLOCK_N, STAGES>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_fwd_tc<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid(m_tiles, (p.Cout + BLOCK_N - 1) / BLOCK_N);
  g_launches += 1;
  k_conv_fwd_tc<BLOCK_N, STAGES><<<grid, 192, SM::TOTAL, st>>>(mx, mw, p);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// forward / dgrad kernel v2: persistent (one CTA per SM loops over output tiles), double-buffered TMEM accumulator
// so that the epilogue of tile i overlaps the mainloop of tile i+1; BLOCK_N up to 256 (A tile re-read halves).
// Tile order: n-tiles of one pixel tile are adjacent (they share the A tile through L2).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int BLOCK_N, int STAGES>
struct Fwd2Smem {
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGE_OFF = STAGES * STAGE_BYTES;          // 4 warps x EPI_NBUF x 4 KB epilogue staging tiles
  static constexpr int BAR_OFF = STAGE_OFF + 4 * EPI_NBUF * 4096;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 4 + 4 * EPI_NBUF) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};
template <int BLOCK_N, int STAGES, int FMT>
__global__ void __launch_bounds__(192, 1) k_conv_fwd_tc2(const __grid_constant__ CUtensorMap map_x,
                                                         const __grid_constant__ CUtensorMap map_w,
                                                         const __grid_constant__ CUtensorMap map_y,
                                                         const __grid_constant__ CUtensorMap map_add, const FwdParams p,
                                                         const int n_tiles, const int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  using SM = Fwd2Smem<BLOCK_N, STAGES>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* abar = tmem_empty + 2;           // [4 warps][EPI_NBUF] addend-tile barriers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(abar + 4 * EPI_NBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CSH = fmt_csh<FMT>();
  const int cchunks = p.Cin >> CSH;
  const int num_kb = p.ks * p.ks * cchunks;
  const int pad = p.ks >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    for (int i = 0; i < 4 * EPI_NBUF; ++i) mbar_init(&abar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SM::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;                                          // global k-block counter of this CTA (ring position)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles, rest = tile / n_tiles;
        const int z = rest % p.ksplit, mt = rest / p.ksplit;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn, col0 = nt * BLOCK_N;
        const int kb0 = z * p.kb_per, kb1 = min(num_kb, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect_tx(&full[st], SM::STAGE_BYTES);
          const int tap = kb / cchunks, c0 = (kb - tap * cchunks) << CSH;
          const int r = tap / p.ks, s = tap - r * p.ks;
          uint8_t* sa = smem + st * SM::STAGE_BYTES;
          tma_load_4d(sa, &map_x, &full[st], c0, w0 + s - pad, h0 + r - pad, n0);
          tma_load_2d(sa + TC_A_BYTES, &map_w, &full[st], tap * p.Cin + c0, col0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {               // one thread runs the whole issue loop (see k_conv_halo)
      constexpr uint32_t idesc = make_idesc_fmt<FMT>(128, BLOCK_N);
      int it = 0, lt = 0;                                    // lt = local tile counter
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);    // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
        const int z = (tile / n_tiles) % p.ksplit;
        const int kb0 = z * p.kb_per, kb1 = min(num_kb, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int st = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[st], ph);
          const uint32_t sa = smem_u32(smem + st * SM::STAGE_BYTES);
          const uint32_t sb = sa + TC_A_BYTES;
          umma_row<FMT, false>(tmem_d, sa, 1024, 0, sb, idesc, kb == kb0);
          umma_commit(&empty[st]);
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dw = row % p.bw, dh = (row / p.bw) % p.bh, dn = row / (p.bw * p.bh);
    const int r0 = q * 32;                                  // first row of this warp: TMA-store box origin
    const int sdh = (r0 / p.bw) % p.bh, sdn = r0 / (p.bw * p.bh);
    uint8_t* stage = smem + SM::STAGE_OFF + q * (EPI_NBUF * 4096);
    EpiState es;
    const EpiOut eo{p.bias, p.addend, p.y, p.Cout, p.stats, p.tma_store >> 4};
    const bool tma = (p.tma_store & 1) != 0;
    epi_dispatch(epi_variant(eo, tma, (p.tma_store >> 12) & 1), [&](auto emc) {
    constexpr int EM = decltype(emc)::value;
    int lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const int nt = tile % n_tiles, rest = tile / n_tiles;
      const int z = rest % p.ksplit, mt = rest / p.ksplit;
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn + z * p.npad, col0 = nt * BLOCK_N;   // split z -> its own batch slab
      const int n = n0 + dn;
      const bool valid = n < p.N;
      const long long pix = ((long long)n * p.H + (h0 + dh)) * p.W + (w0 + dw);
      if (tma) epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0, w0, h0 + sdh, n0 + sdn, lane);
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        if (col0 + c >= p.Cout) break;                     // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c, v);
        epi_do<EM>(v, eo, pix, valid, col0 + c, tma, stage, es, &map_y, w0, h0 + sdh, n0 + sdn, lane, (long long)mt * 4 + q,
                   abar + q * EPI_NBUF, &map_add);
        if (tma && eo.addend && (c + 32 < BLOCK_N) && (col0 + c + 32 < p.Cout))
          epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0 + c + 32, w0, h0 + sdh, n0 + sdn, lane);
      }
      // this warp is done reading the accumulator: release it to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    });
    if (tma && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc<SM::TMEM_COLS>(tmem_base); }
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
template <int BLOCK_N, int STAGES, int FMT>
static int launch_fwd2_t(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const CUtensorMap& madd,
                         const FwdParams& p, int m_tiles, cudaStream_t st) {
  using SM = Fwd2Smem<BLOCK_N, STAGES>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_fwd_tc2<BLOCK_N, STAGES, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int n_tiles = (p.Cout + BLOCK_N - 1) / BLOCK_N;
  const int total = m_tiles * n_tiles * p.ksplit;
  const int grid = total < num_sms() ? total : num_sms();
  g_launches += 1;
  k_conv_fwd_tc2<BLOCK_N, STAGES, FMT><<<grid, 192, SM::TOTAL, st>>>(mx, mw, my, madd, p, n_tiles, total);
  return (int)cudaGetLastError();
}
// ---------------------------------------------------------------------------------------------------------------
// forward / dgrad kernel v3 (3x3 only): HALO REUSE.  The v1/v2 kernels fetch one A tile per filter tap, i.e. every input
// pixel crosses L2->SM nine times; ncu shows them pinned at the ~10 TB/s L2 read ceiling (profiles/r01a_prof_fwd.md).
// Here one TMA box {32c, 16w, 16T+2 rows} = the pixel tile PLUS its halo is loaded once per 32-channel chunk and
// all nine taps are fed from it: for tap (r,s) the A operand of the UMMA is the same shared-memory tile addressed from
// row (r*16 + s): output pixel (h, w<8) -> smem row (h+r)*16 + (w+s), i.e. 8-row groups 2048 B apart (SBO) whose
// start is shifted by s rows.  (The swizzle is a function of absolute smem address bits, so no base_offset is needed
// as long as the tile itself is 1024B-aligned -- verified on hardware, see halo_mode().)
// T = 2 vertically stacked 16x8 pixel tiles share every B (filter) tile, halving filter traffic as well.
// Two rings: A (halo, consumed for a whole chunk = 9 taps) and B (one filter tap x 32 channels).
// ---------------------------------------------------------------------------------------------------------------
static int fwd_kernel_version();
static bool fwd_kernel_version_is2() { return fwd_kernel_version() != 1; }
// how the epilogue fetches an addend tile.  0: lane-per-row loads (first version), 1: coalesced loads through the staging
// tile, (2: register-prefetched variant, removed: slower, profiles/r01i_probe_conv_bw.txt), 3 (default): TMA load
// straight into the staging tile, one chunk ahead.  Measured extra time of the in-place addend, 32x256x256 64->64 3x3:
// +0.21 / +0.095..0.23 / +0.41 / +0.084 ms; 1x1 64->128 at 128x128: +0.19 / +0.137 / +0.36 / +0.063 ms.
static int addend_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_ADDEND");
    v = e ? atoi(e) : 3;
    if (v == 2 || v < 0 || v > 3) v = 1;
    const char* nf = getenv("SIVAE_TC_NOFENCE");
    if (nf && nf[0] == '1') v |= 8;
  }
  return v;
}
// SIVAE_TC_FASTEPI (default 1): straight-line epilogue variants (epi_chunk_fast) where the launch qualifies
static int fast_epi() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_FASTEPI");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}
static bool tma_store_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_TMASTORE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
__device__ __forceinline__ uint64_t make_smem_desc_bo(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;     // SWIZZLE_128B
  return d;
}
struct HaloParams {
  int N, H, W, Cin, Cout;
  int tma_store;
  int tiles_w, tiles_h;        // W/8, H/16
  int n_tiles, total;          // Cout tiles, total work items
  int base_offset_mode;        // 1: descriptor base_offset = s (documented semantics); 0: always 0 (experiment)
  const float* bias;
  const float* addend;
  float* y;
  float* stats;                // BN partial sums [m_items*T*4][2][Cout] or null
};
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES, int KH, int KW>
struct HaloSmem {
  static constexpr int ROWS = 16 + KH - 1;           // one halo box per 16x8 pixel tile
  // box width in pixels = 8 output columns + column halo.  The 8-pixel row groups of the A operand sit BW*128 bytes apart
  // (SBO): any multiple of 16 B works because the swizzle is a function of absolute address bits, so the box is exactly
  // as wide as the halo needs (10 for 3x3: 37 % less L2->SM traffic and shared memory than the first 16-wide boxes).
  static constexpr int BW = 8 + KW - 1;
  static constexpr int BOX_TX = ROWS * BW * 128;                       // bytes one TMA box delivers
  static constexpr int BOX_BYTES = (BOX_TX + 1023) / 1024 * 1024;      // boxes stay 1024B-aligned
  static constexpr int A_BYTES = T * BOX_BYTES;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int B_OFF = A_STAGES * A_BYTES;
  static constexpr int STAGE_OFF = B_OFF + B_STAGES * B_BYTES;     // 4 warps x EPI_NBUF x 4 KB epilogue staging tiles
  static constexpr int BAR_OFF = STAGE_OFF + 4 * EPI_NBUF * 4096;
  static constexpr int NBAR = 2 * A_STAGES + 2 * B_STAGES + 4 + 4 * EPI_NBUF;
  static constexpr int TOTAL = BAR_OFF + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 2 * T * BLOCK_N;
};
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES, int KH, int KW, int FMT>
__global__ void __launch_bounds__(192, 1) k_conv_halo(const __grid_constant__ CUtensorMap map_x,
                                                      const __grid_constant__ CUtensorMap map_w,
                                                      const __grid_constant__ CUtensorMap map_y,
                                                      const __grid_constant__ CUtensorMap map_add, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = HaloSmem<BLOCK_N, T, A_STAGES, B_STAGES, KH, KW>;
  constexpr int TAPS = KH * KW, PADH = KH / 2, PADW = KW / 2, BW = SM::BW;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* tmem_full = b_empty + B_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint64_t* abar = tmem_empty + 2;            // [4 warps][EPI_NBUF] addend-tile barriers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(abar + 4 * EPI_NBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CSH = fmt_csh<FMT>();
  const int cchunks = p.Cin >> CSH;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    for (int i = 0; i < 4 * EPI_NBUF; ++i) mbar_init(&abar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SM::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int ai = 0, bi = 0;
      for (int item = blockIdx.x; item < p.total; item += gridDim.x) {
        const int nt = item % p.n_tiles, mt = item / p.n_tiles;
        const int col0 = nt * BLOCK_N;
        for (int ch = 0; ch < cchunks; ++ch) {
          {
            const int st = ai % A_STAGES;
            mbar_wait(&a_empty[st], ((ai / A_STAGES) & 1) ^ 1);
            mbar_expect_tx(&a_full[st], T * SM::BOX_TX);
#pragma unroll
            for (int t = 0; t < T; ++t) {        // the T pixel tiles of an item are consecutive in (n, tile row, tile column) order
              const int lin = mt * T + t;
              const int tw = lin % p.tiles_w, th = (lin / p.tiles_w) % p.tiles_h, n = lin / (p.tiles_w * p.tiles_h);
              tma_load_4d(smem + st * SM::A_BYTES + t * SM::BOX_BYTES, &map_x, &a_full[st], ch << CSH, tw * 8 - PADW, th * 16 - PADH, n);
            }
            ++ai;
          }
          for (int tap = 0; tap < TAPS; ++tap, ++bi) {
            const int st = bi % B_STAGES;
            mbar_wait(&b_empty[st], ((bi / B_STAGES) & 1) ^ 1);
            mbar_expect_tx(&b_full[st], SM::B_BYTES);
            tma_load_2d(smem + SM::B_OFF + st * SM::B_BYTES, &map_w, &b_full[st], tap * p.Cin + (ch << CSH), col0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ONE elected thread runs the whole issue loop (barrier waits included): no warp-wide election / re-convergence per
    // filter tap between the MMAs
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_fmt<FMT>(128, BLOCK_N);
      int ai = 0, bi = 0, lt = 0;
      for (int item = blockIdx.x; item < p.total; item += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * T * BLOCK_N);
        for (int ch = 0; ch < cchunks; ++ch, ++ai) {
          const int ast = ai % A_STAGES;
          mbar_wait(&a_full[ast], (ai / A_STAGES) & 1);
          const uint32_t sa = smem_u32(smem + ast * SM::A_BYTES);
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap, ++bi) {
            const int bst = bi % B_STAGES;
            mbar_wait(&b_full[bst], (bi / B_STAGES) & 1);
            const uint32_t sb = smem_u32(smem + SM::B_OFF + bst * SM::B_BYTES);
            const int r = tap / KW, s = tap - KW * r;
            const uint32_t bo = p.base_offset_mode ? (uint32_t)s : 0u;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const uint32_t arow = sa + (uint32_t)(t * SM::BOX_BYTES + (r * BW + s) * 128);
              umma_row<FMT, false>(tmem_d + (uint32_t)(t * BLOCK_N), arow, BW * 128, bo, sb, idesc, ch == 0 && tap == 0);
            }
            umma_commit(&b_empty[bst]);
            if (tap == TAPS - 1) {
              umma_commit(&a_empty[ast]);
              if (ch == cchunks - 1) umma_commit(&tmem_full[acc]);
            }
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dh = row >> 3, dw = row & 7;
    uint8_t* stage = smem + SM::STAGE_OFF + q * (EPI_NBUF * 4096);
    EpiState es;
    const EpiOut eo{p.bias, p.addend, p.y, p.Cout, p.stats, p.tma_store >> 4};
    const bool tma = (p.tma_store & 1) != 0;
    epi_dispatch(epi_variant(eo, tma, (p.tma_store >> 12) & 1), [&](auto emc) {
    constexpr int EM = decltype(emc)::value;
    int lt = 0;
    for (int item = blockIdx.x; item < p.total; item += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const int nt = item % p.n_tiles, mt = item / p.n_tiles;
      const int col0 = nt * BLOCK_N;
      // chunk sequence of an item: (t, c) for t < T, c < BLOCK_N step 32; the addend tile of chunk i+1 is requested right
      // after chunk i's store (that of chunk 0 before the accumulator is ready)
      int ncol = (p.Cout - col0 + 31) / 32;
      if (ncol > BLOCK_N / 32) ncol = BLOCK_N / 32;
      // tile t of the item is linear pixel tile mt*T + t: one division per item, then carry increments (the divisions by
      // runtime tile counts cost more instructions than the chunk epilogue itself)
      const int lin0 = mt * T;
      int tw = lin0 % p.tiles_w, th, n;
      { const int r = lin0 / p.tiles_w; th = r % p.tiles_h; n = r / p.tiles_h; }
      if (tma) epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0, tw * 8, th * 16 + 4 * q, n, lane);
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const int w0 = tw * 8, h0 = th * 16, n_cur = n;
        if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++n; } }       // (tw, th, n) = next tile
        const long long pix = ((long long)n_cur * p.H + (h0 + dh)) * p.W + (w0 + dw);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * T + t) * BLOCK_N);
#pragma unroll 1
        for (int ci = 0; ci < ncol; ++ci) {
          const int c = ci * 32;
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          // this warp's 32 rows = image rows h0+4q .. +3, columns w0 .. w0+7  -> store box {32 ch, 8, 4, 1}
          epi_do<EM>(v, eo, pix, true, col0 + c, tma, stage, es, &map_y, w0, h0 + 4 * q, n_cur, lane, (long long)(lin0 + t) * 4 + q,
                     abar + q * EPI_NBUF, &map_add);
          if ((EM == 0 && tma && eo.addend) || (EM & 2)) {
            if (ci + 1 < ncol) {
              epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0 + c + 32, w0, h0 + 4 * q, n_cur, lane);
            } else if (t + 1 < T) {
              epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0, tw * 8, th * 16 + 4 * q, n, lane);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    });
    if (tma && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc<SM::TMEM_COLS>(tmem_base); }
}
// 0 = off (v2 kernel), 2 = on with descriptor base_offset = 0 (default), 1 = on with base_offset = s.
// Measured on B200 (tests/tc_probe.py, profiles/r01b_halo_probe.txt): the tensor core swizzles on ABSOLUTE shared-memory
// address bits, so a start address shifted by s rows inside a 1024B-aligned tile needs base_offset = 0; base_offset = s
// (the documented formula for a matrix whose own base is unaligned) gives wrong results here.
static int halo_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_HALO");
    v = e ? atoi(e) : 2;
  }
  return v;
}
static bool halo_supported(const ConvShape& s) {
  return (s.k == 3 || s.k == 5) && s.Cin % 32 == 0 && s.Cout >= 1 && s.W % 8 == 0 && s.H % 16 == 0;
}
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES, int KH, int KW, int FMT>
static int launch_halo_f(const float* x, const float* w, const float* bias, const float* addend, float* y, const ConvShape& s,
                         float* stats, cudaStream_t st) {
  using SM = HaloSmem<BLOCK_N, T, A_STAGES, B_STAGES, KH, KW>;
  static_assert(SM::TOTAL <= 232448, "shared memory budget exceeded");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_halo<BLOCK_N, T, A_STAGES, B_STAGES, KH, KW, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  HaloParams p;
  p.N = s.N; p.H = s.H; p.W = s.W; p.Cin = s.Cin; p.Cout = s.Cout;
  p.tiles_w = s.W / 8; p.tiles_h = s.H / 16;
  p.n_tiles = (s.Cout + BLOCK_N - 1) / BLOCK_N;
  p.total = (p.tiles_w * p.tiles_h * s.N / T) * p.n_tiles;      // callers pick T = 2 only when the tile count is even
  p.base_offset_mode = halo_mode() == 1 ? 1 : 0;
  p.bias = bias; p.addend = addend; p.y = y;
  p.tma_store = ((s.Cout & 3) == 0 && tma_store_enabled()) ? (1 | (addend_mode() << 4) | (fast_epi() << 12)) : 0;
  p.stats = p.tma_store ? stats : nullptr;
  CUtensorMap mx, mw, my;
  int r = make_map_nhwc(&mx, x, s.N, s.H, s.W, s.Cin, SM::BW, SM::ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B, FMT == FMT_BF16);
  if (r) return r;
  r = make_map_2d(&mw, w, s.Cout, (long long)KH * KW * s.Cin, BLOCK_N, FMT == FMT_BF16);
  if (r) return r;
  my = mx;
  if (p.tma_store) {
    r = make_store_map(&my, y, s.N, s.H, s.W, s.Cout, 8, 4, 1);
    if (r) return r;
  }
  CUtensorMap madd = my;
  if (p.tma_store && addend && (addend_mode() & 3) == 3) {
    r = make_store_map(&madd, (float*)addend, s.N, s.H, s.W, s.Cout, 8, 4, 1);
    if (r) return r;
  }
  const int grid = p.total < num_sms() ? p.total : num_sms();
  g_launches += 1;
  k_conv_halo<BLOCK_N, T, A_STAGES, B_STAGES, KH, KW, FMT><<<grid, 192, SM::TOTAL, st>>>(mx, mw, my, madd, p);
  return (int)cudaGetLastError();
}
// fmt: FMT_TF32 (x, w fp32 pre-rounded to tf32) or FMT_SPLIT (x, w in the split32 format, same bytes)
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES, int KH, int KW>
static int launch_halo_t(const float* x, const float* w, const float* bias, const float* addend, float* y, const ConvShape& s,
                         float* stats, cudaStream_t st, int fmt) {
  if (fmt == FMT_SPLIT) return launch_halo_f<BLOCK_N, T, A_STAGES, B_STAGES, KH, KW, FMT_SPLIT>(x, w, bias, addend, y, s, stats, st);
  if (fmt == FMT_BF16) {         // plain bf16 operands: the dgrad of the residual blocks (3x3 only)
    if constexpr (KH == 3 && KW == 3) return launch_halo_f<BLOCK_N, T, A_STAGES, B_STAGES, KH, KW, FMT_BF16>(x, w, bias, addend, y, s, stats, st);
    else return -9;
  }
  return launch_halo_f<BLOCK_N, T, A_STAGES, B_STAGES, KH, KW, FMT_TF32>(x, w, bias, addend, y, s, stats, st);
}
// ---------------------------------------------------------------------------------------------------------------
// forward / dgrad kernel v4 (3x3, Cout >= 128): CTA PAIR, tcgen05.mma.cta_group::2, M = 256.
// The single-CTA TF32 mainloop is shared-memory-bandwidth bound (a 128xNx8 tf32 UMMA reads 4 KB of A + 32 N bytes of B:
// 64 clk at N = 128, exactly the MMA's own 64 clk, before any TMA write).  A CTA pair (cluster of 2 = one TPC) issues
// ONE 256 x N x 8 instruction from the leader: each CTA supplies its own 128 pixel rows of A and only HALF of the B tile
// (N/2 filter rows), so at N = 256 a CTA reads 8 KB per 128 clk of tensor work and loads half the filter bytes.
// Each CTA: its own pixel tile (halo box, exactly as k_conv_halo with T = 1), rows [r*N/2, (r+1)*N/2) of every B tile,
// its own 128 x N fp32 accumulator (double buffered) and epilogue.  Barriers: "full" barriers live in the leader and
// collect the TMA bytes of BOTH CTAs (cp.async.bulk.tensor.cta_group::2 with the leader's barrier address); "empty" and
// "accumulator ready" are multicast tcgen05.commit arrivals to the same barrier offset in both CTAs; "accumulator
// drained" = 8 epilogue-warp arrivals (4 local, 4 remote via mapa) on the leader.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();      // .aligned: the warp must be converged (single-thread role loops diverge it)
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Remote arrive on the leader's "accumulator drained" barrier.  What it orders is TMEM traffic only (the warp's tcgen05.ld
// have completed: wait::ld + tcgen05.fence::before_thread_sync), so no memory release at cluster scope is needed; the
// `.release.cluster` form compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the arrive -- 10 % of the epilogue
// warps' time in the ncu capture of the 64-channel layers (profiles/r01o_prof_halo2.md).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {      // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES>
struct Halo2Smem {
  static constexpr int ROWS = 18, BW = 10;
  static constexpr int BOX_TX = ROWS * BW * 128;                  // bytes of one halo box
  static constexpr int BOX_BYTES = (BOX_TX + 1023) / 1024 * 1024;
  static constexpr int A_BYTES = T * BOX_BYTES;                   // T pixel tiles per CTA share every filter half-tile
  static constexpr int B_BYTES = (BLOCK_N / 2) * 128;             // this CTA's half of a filter tile
  static constexpr int B_OFF = A_STAGES * A_BYTES;
  static constexpr int STAGE_OFF = B_OFF + B_STAGES * B_BYTES;
  static constexpr int BAR_OFF = STAGE_OFF + 4 * EPI_NBUF * 4096;
  static constexpr int NBAR = 2 * A_STAGES + 2 * B_STAGES + 4 + 4 * EPI_NBUF;
  static constexpr int TOTAL = BAR_OFF + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 2 * T * BLOCK_N;
};
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    k_conv_halo2(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_add, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = Halo2Smem<BLOCK_N, T, A_STAGES, B_STAGES>;
  constexpr int TAPS = 9;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* tmem_full = b_empty + B_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]  (used in the leader only)
  uint64_t* abar = tmem_empty + 2;            // [4 warps][EPI_NBUF] addend-tile barriers (CTA-local)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(abar + 4 * EPI_NBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int CSH = fmt_csh<FMT>();       // log2(channels per 128-byte row)
  const int cchunks = p.Cin >> CSH;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
    for (int i = 0; i < 4 * EPI_NBUF; ++i) mbar_init(&abar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<SM::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers are initialised before any remote signal can arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int ai = 0, bi = 0;
      for (int item = cluster_id; item < p.total; item += nclusters) {
        const int nt = item % p.n_tiles, mp = item / p.n_tiles;
        const int col0 = nt * BLOCK_N;
        for (int ch = 0; ch < cchunks; ++ch) {
          {
            const int st = ai % A_STAGES;
            mbar_wait(&a_empty[st], ((ai / A_STAGES) & 1) ^ 1);
            if (leader) mbar_expect_tx(&a_full[st], 2 * T * SM::BOX_TX);        // bytes of both CTAs' boxes
            const uint32_t bar = mapa_rank(smem_u32(&a_full[st]), 0);
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const int lin = (mp * 2 + (int)rank) * T + t;
              const int tw = lin % p.tiles_w, th = (lin / p.tiles_w) % p.tiles_h, n = lin / (p.tiles_w * p.tiles_h);
              tma_load_4d_2cta(smem + st * SM::A_BYTES + t * SM::BOX_BYTES, &map_x, bar, ch << CSH, tw * 8 - 1, th * 16 - 1, n);
            }
            ++ai;
          }
          for (int tap = 0; tap < TAPS; ++tap, ++bi) {
            const int st = bi % B_STAGES;
            mbar_wait(&b_empty[st], ((bi / B_STAGES) & 1) ^ 1);
            if (leader) mbar_expect_tx(&b_full[st], 2 * SM::B_BYTES);
            tma_load_2d_2cta(smem + SM::B_OFF + st * SM::B_BYTES, &map_w, mapa_rank(smem_u32(&b_full[st]), 0), tap * p.Cin + (ch << CSH),
                             col0 + (int)rank * (BLOCK_N / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_fmt<FMT>(256, BLOCK_N);
      int ai = 0, bi = 0, lt = 0;
      for (int item = cluster_id; item < p.total; item += nclusters, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * T * BLOCK_N);
        for (int ch = 0; ch < cchunks; ++ch, ++ai) {
          const int ast = ai % A_STAGES;
          mbar_wait(&a_full[ast], (ai / A_STAGES) & 1);
          const uint32_t sa = smem_u32(smem + ast * SM::A_BYTES);
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap, ++bi) {
            const int bst = bi % B_STAGES;
            mbar_wait(&b_full[bst], (bi / B_STAGES) & 1);
            const uint32_t sb = smem_u32(smem + SM::B_OFF + bst * SM::B_BYTES);
            const int r = tap / 3, s = tap - 3 * r;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const uint32_t arow = sa + (uint32_t)(t * SM::BOX_BYTES + (r * SM::BW + s) * 128);
              umma_row<FMT, true>(tmem_d + (uint32_t)(t * BLOCK_N), arow, SM::BW * 128, 0, sb, idesc, ch == 0 && tap == 0);
            }
            umma_commit_2cta(&b_empty[bst]);
            if (tap == TAPS - 1) {
              umma_commit_2cta(&a_empty[ast]);
              if (ch == cchunks - 1) umma_commit_2cta(&tmem_full[acc]);
            }
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int dh = row >> 3, dw = row & 7;
    uint8_t* stage = smem + SM::STAGE_OFF + q * (EPI_NBUF * 4096);
    EpiState es;
    const EpiOut eo{p.bias, p.addend, p.y, p.Cout, p.stats, p.tma_store >> 4};
    const bool tma = (p.tma_store & 1) != 0;
    epi_dispatch(epi_variant(eo, tma, (p.tma_store >> 12) & 1), [&](auto emc) {
    constexpr int EM = decltype(emc)::value;
    int lt = 0;
    for (int item = cluster_id; item < p.total; item += nclusters, ++lt) {
      const int acc = lt & 1;
      const int nt = item % p.n_tiles, mp = item / p.n_tiles;
      const int col0 = nt * BLOCK_N;
      int ncol = (p.Cout - col0 + 31) / 32;
      if (ncol > BLOCK_N / 32) ncol = BLOCK_N / 32;
      // this CTA's tile t of the item is linear pixel tile (2 mp + rank) T + t: one division per item, then carry increments
      const int lin0 = (mp * 2 + (int)rank) * T;
      int tw = lin0 % p.tiles_w, th, n;
      { const int r = lin0 / p.tiles_w; th = r % p.tiles_h; n = r / p.tiles_h; }
      if (tma) epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0, tw * 8, th * 16 + 4 * q, n, lane);
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        const int w0 = tw * 8, h0 = th * 16, n_cur = n;
        if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++n; } }       // (tw, th, n) = next tile
        const long long pix = ((long long)n_cur * p.H + (h0 + dh)) * p.W + (w0 + dw);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * T + t) * BLOCK_N);
#pragma unroll 1
        for (int ci = 0; ci < ncol; ++ci) {
          const int c = ci * 32;
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c, v);
          epi_do<EM>(v, eo, pix, true, col0 + c, tma, stage, es, &map_y, w0, h0 + 4 * q, n_cur, lane, (long long)(lin0 + t) * 4 + q,
                     abar + q * EPI_NBUF, &map_add);
          if ((EM == 0 && tma && eo.addend) || (EM & 2)) {
            if (ci + 1 < ncol) {
              epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0 + c + 32, w0, h0 + 4 * q, n_cur, lane);
            } else if (t + 1 < T) {
              epi_prefetch(eo.addend, eo.amode, es, stage, abar + q * EPI_NBUF, &map_add, col0, tw * 8, th * 16 + 4 * q, n, lane);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty[acc]), 0));      // accumulator drained -> leader's MMA warp
    }
    });
    if (tma && lane == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // no CTA of the pair leaves (or frees TMEM) while the other may still signal it
  if (warp == 1) { __syncwarp(); tmem_dealloc_2cta<SM::TMEM_COLS>(tmem_base); }
}
static int two_cta_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_2CTA");      // 0: off, 1: Cout >= 256 only, 2 (default): every 3x3 layer with Cout > 32
    v = e ? atoi(e) : 2;
  }
  return v;
}
static bool halo2_eligible(const ConvShape& s) {
  // measured (profiles/r01k_probe_2cta.txt): N = 256 pairs run the 256/512-channel layers at 660-800 TFLOP/s against
  // 530-665 single-CTA; at N = 128 the pair (T = 1 per CTA) is no faster than the single-CTA T = 2 kernel (622 vs 648)
  // mode 2 adds T = 2 pairs for the 128- and 64-channel layers (half the filter bytes and 6 resp. 5 KB instead of 8 / 6 KB
  // of operand reads per MMA and CTA)
  const int min_cout = two_cta_mode() >= 2 ? 33 : 256;
  const long long tiles = (long long)s.N * (s.H / 16) * (s.W / 8);
  return two_cta_mode() != 0 && s.k == 3 && s.Cin % 32 == 0 && s.Cout >= min_cout && (s.Cout & 3) == 0 && s.W % 8 == 0 && s.H % 16 == 0 &&
         tiles % (s.Cout >= 256 ? 2 : 4) == 0 && tma_store_enabled();
}
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES, int FMT>
static int launch_halo2_f(const float* x, const float* w, const float* bias, const float* addend, float* y, const ConvShape& s,
                          float* stats, cudaStream_t st) {
  using SM = Halo2Smem<BLOCK_N, T, A_STAGES, B_STAGES>;
  static_assert(SM::TOTAL <= 232448, "shared memory budget exceeded");
  static_assert(SM::TMEM_COLS <= 512, "TMEM budget exceeded");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_halo2<BLOCK_N, T, A_STAGES, B_STAGES, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  HaloParams p;
  p.N = s.N; p.H = s.H; p.W = s.W; p.Cin = s.Cin; p.Cout = s.Cout;
  p.tiles_w = s.W / 8; p.tiles_h = s.H / 16;
  p.n_tiles = (s.Cout + BLOCK_N - 1) / BLOCK_N;
  p.total = (p.tiles_w * p.tiles_h * s.N / (2 * T)) * p.n_tiles;    // items = (2T consecutive pixel tiles, n-tile)
  p.base_offset_mode = 0;
  p.bias = bias; p.addend = addend; p.y = y;
  p.tma_store = 1 | (addend_mode() << 4) | (fast_epi() << 12);
  p.stats = stats;
  CUtensorMap mx, mw, my;
  int r = make_map_nhwc(&mx, x, s.N, s.H, s.W, s.Cin, SM::BW, SM::ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B, FMT == FMT_BF16);
  if (r) return r;
  r = make_map_2d(&mw, w, s.Cout, (long long)9 * s.Cin, BLOCK_N / 2, FMT == FMT_BF16);
  if (r) return r;
  r = make_store_map(&my, y, s.N, s.H, s.W, s.Cout, 8, 4, 1);
  if (r) return r;
  CUtensorMap madd = my;
  if (addend && (addend_mode() & 3) == 3) {
    r = make_store_map(&madd, (float*)addend, s.N, s.H, s.W, s.Cout, 8, 4, 1);
    if (r) return r;
  }
  int nclusters = num_sms() / 2;
  if (p.total < nclusters) nclusters = p.total;
  g_launches += 1;
  k_conv_halo2<BLOCK_N, T, A_STAGES, B_STAGES, FMT><<<2 * nclusters, 192, SM::TOTAL, st>>>(mx, mw, my, madd, p);
  return (int)cudaGetLastError();
}
template <int BLOCK_N, int T, int A_STAGES, int B_STAGES>
static int launch_halo2_t(const float* x, const float* w, const float* bias, const float* addend, float* y, const ConvShape& s,
                          float* stats, cudaStream_t st, int fmt) {
  if (fmt == FMT_SPLIT) return launch_halo2_f<BLOCK_N, T, A_STAGES, B_STAGES, FMT_SPLIT>(x, w, bias, addend, y, s, stats, st);
  if (fmt == FMT_BF16) return launch_halo2_f<BLOCK_N, T, A_STAGES, B_STAGES, FMT_BF16>(x, w, bias, addend, y, s, stats, st);
  return launch_halo2_f<BLOCK_N, T, A_STAGES, B_STAGES, FMT_TF32>(x, w, bias, addend, y, s, stats, st);
}
static int launch_halo2(const float* x, const float* w, const float* bias, const float* addend, float* y, const ConvShape& s,
                        float* stats, cudaStream_t st, int fmt) {
  if (s.Cout >= 256) return launch_halo2_t<256, 1, 3, 6>(x, w, bias, addend, y, s, stats, st, fmt);
  if (s.Cout > 64) return launch_halo2_t<128, 2, 2, 8>(x, w, bias, addend, y, s, stats, st, fmt);
  return launch_halo2_t<64, 2, 3, 8>(x, w, bias, addend, y, s, stats, st, fmt);
}
static int launch_halo(const float* x, const float* w, const float* bias, const float* addend, float* y, const ConvShape& s,
                       float* stats, cudaStream_t st, int fmt) {
  if (halo2_eligible(s)) return launch_halo2(x, w, bias, addend, y, s, stats, st, fmt);
  const bool two = (((long long)s.N * (s.H / 16) * (s.W / 8)) % 2 == 0);
  if (s.k == 5) {     // image-facing 5x5 with a narrow output (predict forward, stem dgrad): N tile of 32
    return two ? launch_halo_t<32, 2, 2, 8, 5, 5>(x, w, bias, addend, y, s, stats, st, fmt) : launch_halo_t<32, 1, 3, 8, 5, 5>(x, w, bias, addend, y, s, stats, st, fmt);
  }
  if (s.Cout > 64) return two ? launch_halo_t<128, 2, 2, 5, 3, 3>(x, w, bias, addend, y, s, stats, st, fmt) : launch_halo_t<128, 1, 3, 7, 3, 3>(x, w, bias, addend, y, s, stats, st, fmt);
  return two ? launch_halo_t<64, 2, 3, 6, 3, 3>(x, w, bias, addend, y, s, stats, st, fmt) : launch_halo_t<64, 1, 3, 8, 3, 3>(x, w, bias, addend, y, s, stats, st, fmt);
}
// ---------------------------------------------------------------------------------------------------------------
// Row-separable form of the two image-facing 5x5 convolutions (c = cdim <= 3 channels on one side).
// A 5x5 conv with a c-channel side wastes the tensor core (K or N of 3) and, as a direct CUDA-core kernel, sits at
// ~28 TFLOP/s (profiles/r01f_layers_H.md: 0.72-1.05 ms per launch at 256x256).  Both directions become a 5x1 (rows only)
// implicit GEMM for the halo kernel above plus one cheap streaming pass:
//   narrow INPUT  (stem forward, predict dgrad):  xe[n,h,w,j] = x[n,h,w-2+j/c, j%c]  (j < 5c, zero-padded to 32 channels:
//       in NHWC the 5-pixel window of a c-channel row is 5c CONTIGUOUS floats), then  y = conv5x1(xe, We),
//       We[co][r][j] = F[co][r][j/c][j%c]  -- again 5c contiguous floats of the original filter.
//   narrow OUTPUT (predict forward, stem dgrad):  P = conv5x1(x, Wg) with 16 output columns (s, co),
//       Wg[s*c+co][r][ci] = F[co][r][s][ci],  then  y[n,h,w,co] = bias + addend + sum_s P[n,h,w+s-2, s*c+co].
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rs_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// rev = 1: mirrored column shift, xe[n,h,w,s*c+ch] = x[n,h,w+2-s,ch] (the dP/dy expansion used by the predict wgrad)
// (all three helpers: the channel quad q of a thread is loop-invariant -- the grid stride is a multiple of 8 -- so the per-channel
// index arithmetic is hoisted, and pixel indices are 32-bit.  They run at 1.8-2.7 TB/s under ncu, profiles/r02d_prof_rowsep.md;
// removing the 64-bit `pix % W` of the first version did not change that (r02_cmd20.sh): half of the 32 expanded channels are
// zero padding and each thread moves 16 bytes per four scalar loads -- the next step is 16-channel rows)
__global__ void k_rowsep_expand_rev(const float* __restrict__ x, float* __restrict__ xe, long long rows, int W, int c) {
  const long long total = rows * W * 8;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int q = (int)(i0 & 7);
  int so[4], cho[4];
  bool ok[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = q * 4 + e;
    so[e] = j / c; cho[e] = j - so[e] * c; ok[e] = j < 5 * c;
  }
  for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x) {
    const unsigned pix = (unsigned)(i >> 3);
    const int w = (int)(pix % (unsigned)W);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ws = w + 2 - so[e];
      v[e] = (ok[e] && ws >= 0 && ws < W) ? rs_tf32(__ldg(x + ((long long)pix + 2 - so[e]) * c + cho[e])) : 0.f;
    }
    *reinterpret_cast<float4*>(xe + (long long)pix * 32 + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
template <bool SPLIT>      // SPLIT: xe in the split32 format (forward operand), unrounded source values
__global__ void k_rowsep_expand(const float* __restrict__ x, float* __restrict__ xe, long long rows, int W, int c) {
  // one thread per (pixel, channel quad): 8 threads cover the 32 expanded channels of a pixel
  const long long total = rows * W * 8;
  const int rowlen = W * c;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int q = (int)(i0 & 7);
  const bool active = q * 4 < 5 * c;
  for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x) {
    const unsigned pix = (unsigned)(i >> 3);
    const unsigned row = pix / (unsigned)W;
    const int w = (int)(pix - row * (unsigned)W);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      const float* src = x + (long long)row * rowlen;
      const int j0 = (w - 2) * c + q * 4;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + e;
        v[e] = (q * 4 + e < 5 * c && j >= 0 && j < rowlen) ? (SPLIT ? __ldg(src + j) : rs_tf32(__ldg(src + j))) : 0.f;
      }
      o = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (SPLIT) split32_store4(xe + (long long)pix * 32, (uint32_t)q, o);
    else *reinterpret_cast<float4*>(xe + (long long)pix * 32 + q * 4) = o;
  }
}
__global__ void k_rowsep_gather(const float* __restrict__ P, const float* __restrict__ bias, const float* __restrict__ addend,
                                float* __restrict__ y, long long rows, int W, int c) {
  const long long total = rows * W;
  for (long long pl = blockIdx.x * (long long)blockDim.x + threadIdx.x; pl < total; pl += (long long)gridDim.x * blockDim.x) {
    const unsigned pix = (unsigned)pl;
    const int w = (int)(pix % (unsigned)W);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const int ws = w + s - 2;
      if (ws < 0 || ws >= W) continue;
      const float* src = P + ((long long)pix + s - 2) * 16 + s * c;
      for (int co = 0; co < c; ++co) acc[co] += __ldg(src + co);
    }
    for (int co = 0; co < c; ++co) {
      float v = acc[co];
      if (bias) v += __ldg(bias + co);
      if (addend) v += addend[(long long)pix * c + co];
      y[(long long)pix * c + co] = v;
    }
  }
}
// F[Co][5][5][c] -> We[Co][5][32] (tf32-rounded, zero padded)
__global__ void k_rowsep_filter_expand(const float* __restrict__ f, float* __restrict__ out, int Co, int c, int rnd) {
  const int total = Co * 5 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 31, cr = i >> 5;
    const float v = j < 5 * c ? f[(long long)cr * 5 * c + j] : 0.f;
    out[i] = rnd ? rs_tf32(v) : v;
  }
}
// F[c][5][5][Ci] -> Wg[16][5][Ci]: row s*c+co, (tf32-rounded, unused rows zero)
__global__ void k_rowsep_filter_gather(const float* __restrict__ f, float* __restrict__ out, int c, int Ci, int rnd) {
  const int total = 16 * 5 * Ci;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Ci, r = (i / Ci) % 5, row = i / (5 * Ci);
    float v = 0.f;
    if (row < 5 * c) {
      const int s = row / c, co = row - s * c;
      v = f[(((long long)co * 5 + r) * 5 + s) * Ci + ci];
      if (rnd) v = rs_tf32(v);
    }
    out[i] = v;
  }
}
static bool rowsep_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_ROWSEP");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0 && halo_mode() != 0 && tma_store_enabled() && fwd_kernel_version_is2();
}
bool conv_rowsep_in_supported(const ConvShape& s) {
  return rowsep_enabled() && s.k == 5 && s.Cin >= 1 && s.Cin <= 3 && s.Cout >= 32 && (s.Cout & 3) == 0 && s.W % 8 == 0 && s.H % 16 == 0;
}
bool conv_rowsep_out_supported(const ConvShape& s) {
  return rowsep_enabled() && s.k == 5 && s.Cout >= 1 && s.Cout <= 3 && s.Cin % 32 == 0 && s.W % 8 == 0 && s.H % 16 == 0;
}
long long conv_rowsep_scratch_floats(const ConvShape& s) { return (long long)s.N * s.H * s.W * 32; }
long long conv_rowsep_filter_floats(const ConvShape& s) { return s.Cin <= 3 ? (long long)s.Cout * 160 : (long long)80 * s.Cin; }
void launch_rowsep_filter_expand(const float* f, float* out, int Co, int c, cudaStream_t st, bool round_tf32) {
  g_launches += 1;
  k_rowsep_filter_expand<<<(Co * 160 + 255) / 256, 256, 0, st>>>(f, out, Co, c, round_tf32 ? 1 : 0);
}
void launch_rowsep_filter_gather(const float* f, float* out, int c, int Ci, cudaStream_t st, bool round_tf32) {
  g_launches += 1;
  k_rowsep_filter_gather<<<(80 * Ci + 255) / 256, 256, 0, st>>>(f, out, c, Ci, round_tf32 ? 1 : 0);
}
// y[N,H,W,Cout] = conv5x5(x[N,H,W,c], F) (+bias)(+addend); we = expanded filter; scratch >= N*H*W*32 floats
int launch_conv_rowsep_in(const float* x, const float* we, const float* bias, const float* addend, float* y, const ConvShape& s,
                          float* scratch, float* stats, cudaStream_t st, int fmt) {
  const long long rows = (long long)s.N * s.H;
  g_launches += 1;
  if (fmt == FMT_SPLIT) k_rowsep_expand<true><<<num_sms() * 8, 256, 0, st>>>(x, scratch, rows, s.W, s.Cin);
  else k_rowsep_expand<false><<<num_sms() * 8, 256, 0, st>>>(x, scratch, rows, s.W, s.Cin);
  ConvShape e{s.N, s.H, s.W, 32, s.Cout, 5};
  const bool two = (((long long)s.N * (s.H / 16) * (s.W / 8)) % 2 == 0);
  if (two) return launch_halo_t<64, 2, 2, 6, 5, 1>(scratch, we, bias, addend, y, e, stats, st, fmt);
  return launch_halo_t<64, 1, 3, 8, 5, 1>(scratch, we, bias, addend, y, e, stats, st, fmt);
}
// y[N,H,W,c] = conv5x5(x[N,H,W,Cin], F) (+bias)(+addend); wg = gather-form filter; scratch >= N*H*W*16 floats
int launch_conv_rowsep_out(const float* x, const float* wg, const float* bias, const float* addend, float* y, const ConvShape& s,
                           float* scratch, cudaStream_t st, int fmt) {
  ConvShape e{s.N, s.H, s.W, s.Cin, 16, 5};
  const bool two = (((long long)s.N * (s.H / 16) * (s.W / 8)) % 2 == 0);
  int r = two ? launch_halo_t<32, 2, 2, 8, 5, 1>(x, wg, nullptr, nullptr, scratch, e, nullptr, st, fmt)
                          : launch_halo_t<32, 1, 3, 8, 5, 1>(x, wg, nullptr, nullptr, scratch, e, nullptr, st, fmt);
  if (r) return r;
  g_launches += 1;
  k_rowsep_gather<<<num_sms() * 8, 256, 0, st>>>(scratch, bias, addend, y, (long long)s.N * s.H, s.W, s.Cout);
  return (int)cudaGetLastError();
}

static bool halo_eligible(const ConvShape& s) {
  if (!halo_supported(s)) return false;
  if (s.k == 5) return s.Cout <= 32;          // wide 5x5 outputs do not occur in this model
  return s.Cout >= 32;
}

static int fwd_kernel_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_FWD");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

static int splitk_count(const ConvShape& s);
// number of BN partial rows the fused-statistics epilogue produces for this shape (0: not available, use launch_bn_stats)
int conv_tc_stats_parts(const ConvShape& s) {
  if (conv_rowsep_in_supported(s)) return s.N * (s.H / 16) * (s.W / 8) * 4;
  if (!conv_tc_supported_fwd(s) || (s.Cout & 3) != 0 || !tma_store_enabled() || fwd_kernel_version() == 1) return 0;
  if (halo_mode() != 0 && halo_eligible(s)) return s.N * (s.H / 16) * (s.W / 8) * 4;      // items * T * 4 warps (T cancels)
  if (splitk_count(s) > 1) return 0;                // split-K layers: statistics from the reduced output (launch_bn_stats)
  int bw, bh, bn;
  pick_tile(s.H, s.W, &bw, &bh, &bn);
  return (s.W / bw) * (s.H / bh) * ((s.N + bn - 1) / bn) * 4;
}
// ---- split-K for few-tile layers (4x4 / 8x8 feature maps): tiles x splits CTAs each accumulate a slice of the
// (tap, channel-chunk) loop into their own slab of a partial tensor; one pass sums the slabs in fixed order (+bias, +addend).
__global__ void k_splitk_reduce(const float4* __restrict__ part, const float* __restrict__ bias, const float4* __restrict__ addend,
                                float4* __restrict__ y, long long n4, long long slab4, int splits, int cvec) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = part[i];
    for (int z = 1; z < splits; ++z) {
      const float4 b = part[z * slab4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % cvec));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (addend) {
      const float4 b = addend[i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    y[i] = a;
  }
}
struct SplitKPlan { int block_n, ksplit, kb_per, npad; };
static bool splitk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_SPLITK");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
static SplitKPlan splitk_plan(const ConvShape& s, int fmt = FMT_TF32) {
  SplitKPlan pl{0, 1, 0, 0};
  int bw, bh, bn;
  pick_tile(s.H, s.W, &bw, &bh, &bn);
  const int tiles_n = (s.N + bn - 1) / bn;
  const int m_tiles = (s.W / bw) * (s.H / bh) * tiles_n;
  // few-tile layers: keep the WIDEST N tile (a narrow tile re-streams the A tiles Cout/block_n times and pins the kernel
  // at the L2->SM ceiling: 8x8x512->512 with 64-wide tiles moved 442 MB per launch) and fill the SMs by splitting K instead
  int block_n = s.Cout > 128 ? 256 : (s.Cout > 64 ? 128 : (s.Cout > 32 ? 64 : 32));
  const int num_kb = s.k * s.k * (s.Cin / (fmt == FMT_BF16 ? 64 : 32));      // k-blocks = (tap, 128-byte channel chunk)
  pl.kb_per = num_kb;
  pl.npad = tiles_n * bn;
  const long long wide_tiles = (long long)m_tiles * ((s.Cout + block_n - 1) / block_n);
  if (splitk_enabled() && tma_store_enabled() && (s.Cout & 3) == 0 && wide_tiles * 2 <= num_sms() && num_kb >= 16) {
    int ks = (int)(num_sms() / wide_tiles);
    if (ks > num_kb / 8) ks = num_kb / 8;
    if (ks > 1) {
      pl.kb_per = (num_kb + ks - 1) / ks;
      pl.ksplit = (num_kb + pl.kb_per - 1) / pl.kb_per;     // every split owns at least one k-block
    }
  }
  if (pl.ksplit == 1)
    while (block_n > 64 && (long long)m_tiles * ((s.Cout + block_n - 1) / block_n) < num_sms()) block_n >>= 1;
  pl.block_n = block_n;
  return pl;
}
static int splitk_count(const ConvShape& s) { return splitk_plan(s).ksplit; }
size_t conv_tc_splitk_scratch_bytes(const ConvShape& s) {
  if (!conv_tc_supported_fwd(s) || (halo_mode() != 0 && halo_eligible(s)) || fwd_kernel_version() == 1) return 0;
  SplitKPlan pl = splitk_plan(s);
  return pl.ksplit > 1 ? (size_t)pl.ksplit * pl.npad * s.H * s.W * s.Cout * sizeof(float) : 0;
}
int launch_conv_fwd_tc(const float* x, const float* w, const float* bias, const float* addend, float* y, const ConvShape& s,
                       cudaStream_t st, float* stats, void* scratch, size_t scratch_bytes, int fmt) {
  FwdParams p;
  const bool b16 = fmt == FMT_BF16;
  if (b16 && s.Cin % 64 != 0) return -9;
  p.ksplit = 1; p.kb_per = s.k * s.k * (s.Cin / (b16 ? 64 : 32)); p.npad = 0;
  p.N = s.N; p.H = s.H; p.W = s.W; p.Cin = s.Cin; p.Cout = s.Cout; p.ks = s.k;
  pick_tile(s.H, s.W, &p.bw, &p.bh, &p.bn);
  p.tiles_w = s.W / p.bw; p.tiles_h = s.H / p.bh;
  p.bias = bias; p.addend = addend; p.y = y;
  p.sbh = p.sbn = 1; p.tma_store = 0; p.stats = nullptr;
  const int tiles_n = (s.N + p.bn - 1) / p.bn;
  const int m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  if (fwd_kernel_version() != 1 && halo_mode() != 0 && halo_eligible(s)) return launch_halo(x, w, bias, addend, y, s, stats, st, fmt);
  if (b16 && fwd_kernel_version() == 1) return -9;
  CUtensorMap mx, mw;
  int r = make_map_nhwc(&mx, x, s.N, s.H, s.W, s.Cin, p.bw, p.bh, p.bn, CU_TENSOR_MAP_SWIZZLE_128B, b16);
  if (r) return r;
  if (fwd_kernel_version() == 1) {
    if (fmt != FMT_TF32) return -9;             // the first-generation kernel (SIVAE_TC_FWD=1, experiments only) is tf32-only
    const int block_n = s.Cout > 64 ? 128 : (s.Cout > 32 ? 64 : 32);
    r = make_map_2d(&mw, w, s.Cout, (long long)s.k * s.k * s.Cin, block_n);
    if (r) return r;
    if (block_n == 128) return launch_fwd_t<128, 3>(mx, mw, p, m_tiles, st);
    if (block_n == 64) return launch_fwd_t<64, 4>(mx, mw, p, m_tiles, st);
    return launch_fwd_t<32, 4>(mx, mw, p, m_tiles, st);
  }
  // v2: widest N tile that the layer fills; few-tile problems prefer narrower tiles (more CTAs busy) and split K
  SplitKPlan pl = splitk_plan(s, fmt);
  const int block_n = pl.block_n;
  if (pl.ksplit > 1 && (!scratch || scratch_bytes < (size_t)pl.ksplit * pl.npad * s.H * s.W * s.Cout * sizeof(float))) { pl.ksplit = 1; pl.kb_per = p.kb_per; }
  r = make_map_2d(&mw, w, s.Cout, (long long)s.k * s.k * s.Cin, block_n, b16);
  if (r) return r;
  // epilogue store path: TMA store of each warp's 32-row slab (needs 16-byte aligned channel rows)
  CUtensorMap my = mx;
  if (b16 && ((s.Cout & 3) != 0 || !tma_store_enabled())) return -9;      // (my = mx is only a placeholder for fp32 inputs)
  p.tma_store = ((s.Cout & 3) == 0 && tma_store_enabled()) ? (1 | (addend_mode() << 4) | (fast_epi() << 12)) : 0;
  p.stats = p.tma_store ? stats : nullptr;
  if (p.tma_store) {
    p.sbh = (32 / p.bw) < p.bh ? (32 / p.bw) : p.bh;
    p.sbn = 32 / (p.bw * p.sbh);
    if (pl.ksplit > 1) {        // raw partial sums into the scratch tensor [ksplit * npad][H][W][Cout]
      p.ksplit = pl.ksplit; p.kb_per = pl.kb_per; p.npad = pl.npad;
      p.bias = nullptr; p.addend = nullptr; p.stats = nullptr; p.y = (float*)scratch;
      p.N = pl.ksplit * pl.npad;       // every slab row is "valid" for the epilogue
      r = make_store_map(&my, (float*)scratch, pl.ksplit * pl.npad, s.H, s.W, s.Cout, p.bw, p.sbh, p.sbn);
    } else {
      r = make_store_map(&my, y, s.N, s.H, s.W, s.Cout, p.bw, p.sbh, p.sbn);
    }
    if (r) return r;
  }
  CUtensorMap madd = my;
  if (p.tma_store && p.addend && (addend_mode() & 3) == 3) {
    r = make_store_map(&madd, (float*)p.addend, s.N, s.H, s.W, s.Cout, p.bw, p.sbh, p.sbn);
    if (r) return r;
  }
  if (fmt == FMT_BF16) {
    if (block_n == 256) r = launch_fwd2_t<256, 4, FMT_BF16>(mx, mw, my, madd, p, m_tiles, st);
    else if (block_n == 128) r = launch_fwd2_t<128, 5, FMT_BF16>(mx, mw, my, madd, p, m_tiles, st);
    else if (block_n == 64) r = launch_fwd2_t<64, 7, FMT_BF16>(mx, mw, my, madd, p, m_tiles, st);
    else r = launch_fwd2_t<32, 8, FMT_BF16>(mx, mw, my, madd, p, m_tiles, st);
  } else if (fmt == FMT_SPLIT) {
    if (block_n == 256) r = launch_fwd2_t<256, 4, FMT_SPLIT>(mx, mw, my, madd, p, m_tiles, st);
    else if (block_n == 128) r = launch_fwd2_t<128, 5, FMT_SPLIT>(mx, mw, my, madd, p, m_tiles, st);
    else if (block_n == 64) r = launch_fwd2_t<64, 7, FMT_SPLIT>(mx, mw, my, madd, p, m_tiles, st);
    else r = launch_fwd2_t<32, 8, FMT_SPLIT>(mx, mw, my, madd, p, m_tiles, st);
  } else {
    if (block_n == 256) r = launch_fwd2_t<256, 4, FMT_TF32>(mx, mw, my, madd, p, m_tiles, st);
    else if (block_n == 128) r = launch_fwd2_t<128, 5, FMT_TF32>(mx, mw, my, madd, p, m_tiles, st);
    else if (block_n == 64) r = launch_fwd2_t<64, 7, FMT_TF32>(mx, mw, my, madd, p, m_tiles, st);
    else r = launch_fwd2_t<32, 8, FMT_TF32>(mx, mw, my, madd, p, m_tiles, st);
  }
  if (r || pl.ksplit <= 1) return r;
  const long long n4 = (long long)s.N * s.H * s.W * s.Cout / 4, slab4 = (long long)pl.npad * s.H * s.W * s.Cout / 4;
  unsigned blocks = (unsigned)((n4 + 255) / 256);
  if (blocks > 148u * 8) blocks = 148u * 8;
  g_launches += 1;
  k_splitk_reduce<<<blocks, 256, 0, st>>>((const float4*)scratch, bias, (const float4*)addend, (float4*)y, n4, slab4, pl.ksplit,
                                          s.Cout / 4);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad kernel: D[m = tap*Cin + ci][n = co] = sum_pixels X[pixel + tap][ci] * dY[pixel][co]   (both MN-major)
// ---------------------------------------------------------------------------------------------------------------
struct WgParams {
  int N, H, W, Cin, Cout, ks;
  int pw, ph, pn;              // pixel block (K block) = pn x ph x pw = 32 pixels
  int tiles_w, tiles_h;        // W/pw, H/ph
  long long kb_total;          // number of pixel blocks
  long long kb_per_split;
  int Ktot;                    // ks*ks*Cin
  float* part;                 // [splits][Cout][Ktot]
};
constexpr int WG_CHUNK_BYTES = 32 * 128;    // 32 pixel rows x 32 channels fp32 (one TMA box)

template <int BLOCK_N, int STAGES>
struct WgSmem {
  static constexpr int A_BYTES = 4 * WG_CHUNK_BYTES;
  static constexpr int B_BYTES = (BLOCK_N / 32) * WG_CHUNK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(192) k_conv_wgrad_tc(const __grid_constant__ CUtensorMap map_x,
                                                       const __grid_constant__ CUtensorMap map_dy, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = WgSmem<BLOCK_N, STAGES>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;            // first (tap,ci) row of this tile
  const int col0 = blockIdx.y * BLOCK_N;      // first co
  const long long kb_begin = (long long)blockIdx.z * p.kb_per_split;
  long long kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.kb_total) kb_end = p.kb_total;
  const int num_kb = (int)(kb_end - kb_begin);
  const int pad = p.ks >> 1;
  // number of valid 32-row chunks of A in this tile
  int a_chunks = (p.Ktot - m0 + 31) / 32;
  if (a_chunks > 4) a_chunks = 4;
  int b_chunks = (p.Cout - col0 + 31) / 32;
  if (b_chunks > BLOCK_N / 32) b_chunks = BLOCK_N / 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < num_kb; ++i) {
        const int st = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], (uint32_t)(a_chunks + b_chunks) * WG_CHUNK_BYTES);
        const long long kb = kb_begin + i;
        const int tw = (int)(kb % p.tiles_w);
        const int th = (int)((kb / p.tiles_w) % p.tiles_h);
        const int tn = (int)(kb / ((long long)p.tiles_w * p.tiles_h));
        const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.pn;
        uint8_t* sa = smem + st * SM::STAGE_BYTES;
        for (int j = 0; j < a_chunks; ++j) {
          const int m = m0 + 32 * j;
          const int tap = m / p.Cin, c0 = m - tap * p.Cin;
          const int r = tap / p.ks, s = tap - r * p.ks;
          tma_load_4d(sa + j * WG_CHUNK_BYTES, &map_x, &full[st], c0, w0 + s - pad, h0 + r - pad, n0);
        }
        uint8_t* sb = sa + SM::A_BYTES;
        for (int j = 0; j < b_chunks; ++j)
          tma_load_4d(sb + j * WG_CHUNK_BYTES, &map_dy, &full[st], col0 + 32 * j, w0, h0, n0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(128, BLOCK_N, 1, 1);
    if (elect_one())                   // one thread runs the whole issue loop (see k_conv_halo)
    for (int i = 0; i < num_kb; ++i) {
      const int st = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full[st], ph);
      {
        const uint32_t sa = smem_u32(smem + st * SM::STAGE_BYTES);
        const uint32_t sb = sa + SM::A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // MN-major fp32: SWIZZLE_128B_BASE32B.  32 MN elements (128 B) contiguous per K row, K rows 128 B apart,
          // swizzle atom = 4 K rows (512 B = SBO), next 32-wide MN chunk one TMA box (4096 B) further (LBO);
          // one MMA consumes 8 K rows = two atoms; advance 1024 B per MMA
          uint64_t ad = make_smem_desc(sa + k * 1024, WG_CHUNK_BYTES, 512, 1);
          uint64_t bd = make_smem_desc(sb + k * 1024, WG_CHUNK_BYTES, 512, 1);
          umma_tf32(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
        if (i == num_kb - 1) umma_commit(tmem_full);
      }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool valid = m < p.Ktot && num_kb > 0;
    float* dst = p.part + (long long)blockIdx.z * p.Cout * p.Ktot;
    if (num_kb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (m < p.Ktot) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = col0 + c + j;
          if (co < p.Cout) dst[(long long)co * p.Ktot + m] = valid ? __uint_as_float(v[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc<BLOCK_N>(tmem_base); }
}

__global__ void k_wg_reduce(const float* __restrict__ part, float* __restrict__ out, long long n, int splits, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(long long)z * n + i];
    out[i] = accumulate ? out[i] + s : s;
  }
}

static void pick_pixel_block(int H, int W, int* pw, int* ph, int* pn) {
  int w = 1;
  while (w * 2 <= 32 && W % (w * 2) == 0) w *= 2;
  int h = 1;
  while (h * 2 <= 32 / w && H % (h * 2) == 0) h *= 2;
  *pw = w; *ph = h; *pn = 32 / (w * h);
}
static int wg_block_n(int Cout) { return Cout > 128 ? 256 : (Cout > 64 ? 128 : 64); }
static void wg_plan(const ConvShape& s, int* splits, long long* kb_total, long long* kb_per_split) {
  int pw, ph, pn;
  pick_pixel_block(s.H, s.W, &pw, &ph, &pn);
  long long kbt = (long long)(s.W / pw) * (s.H / ph) * ((s.N + pn - 1) / pn);
  int bn = wg_block_n(s.Cout);
  long long tiles = (long long)((s.ktot() + 127) / 128) * ((s.Cout + bn - 1) / bn);
  long long want = (148 * 2 + tiles - 1) / tiles;
  long long maxs = (kbt + 15) / 16;       // at least 16 pixel blocks per CTA
  long long sp = want < maxs ? want : maxs;
  if (sp < 1) sp = 1;
  if (sp > 512) sp = 512;
  long long per = (kbt + sp - 1) / sp;
  sp = (kbt + per - 1) / per;
  *splits = (int)sp; *kb_total = kbt; *kb_per_split = per;
}
bool conv_tc_supported_wgrad(const ConvShape& s) {
  if (s.Cin % 32 != 0 || s.Cout % 32 != 0) return false;
  if (s.k != 1 && s.k != 3) return false;
  int pw, ph, pn;
  pick_pixel_block(s.H, s.W, &pw, &ph, &pn);
  if (pn > 256) return false;
  return true;
}
struct Wg2Plan;
static bool wgrad_halo_supported(const ConvShape& s);
static size_t wg2_scratch_bytes(const ConvShape& s);
size_t conv_wgrad_tc_scratch_bytes(const ConvShape& s) {
  if (wgrad_halo_supported(s)) return wg2_scratch_bytes(s);
  int splits; long long kbt, per;
  wg_plan(s, &splits, &kbt, &per);
  return (size_t)splits * s.Cout * s.ktot() * sizeof(float);
}
template <int BLOCK_N, int STAGES>
static int launch_wg_t(const CUtensorMap& mx, const CUtensorMap& mdy, const WgParams& p, int splits, cudaStream_t st) {
  using SM = WgSmem<BLOCK_N, STAGES>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad_tc<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid((p.Ktot + 127) / 128, (p.Cout + BLOCK_N - 1) / BLOCK_N, splits);
  g_launches += 2;
  k_conv_wgrad_tc<BLOCK_N, STAGES><<<grid, 192, SM::TOTAL, st>>>(mx, mdy, p);
  return (int)cudaGetLastError();
}
// ---------------------------------------------------------------------------------------------------------------
// wgrad kernel v2 (3x3, H % 16 == 0, W % 8 == 0): HALO REUSE in the MN-major formulation.
// Work item = one 16x8 pixel tile (K = 128 pixels).  Per item the CTA loads, for each of its G 32-channel input chunks,
// ONE halo box of x {32c, 16w, 18h} (36 KB) and the dy tile {32c, 8w, 16h} per 32 output channels.  For filter row r
// and image row h one UMMA (M = 128, N = N_TILE, K = 8 pixels) is issued whose A operand is the x halo addressed at row
// (h + r) * 16 with FOUR MN chunks 128 B (= one pixel = one filter column s) apart:
//        D[(s, ci)][co] += sum_w x[h + r - 1][w + s - 1][ci] * dy[h][w][co]           s = 0..3 (s = 3 is a phantom tap)
// so the three real taps of a filter row come out of one instruction and x is read from L2 once instead of 9 times
// (v1 sat at 75-100 TFLOP/s: L2 + TMA-issue bound, profiles/r01c_layers_H_halo.md).  Accumulators for all
// (chunk, r) pairs stay in TMEM (G * 3 * N_TILE <= 512 columns) across the CTA's whole pixel range; split over pixel
// tiles, partials reduced in fixed order.
// ---------------------------------------------------------------------------------------------------------------
struct Wg2Params {
  int N, H, W, Cin, Cout, Ktot;
  int tiles_w, tiles_h;
  long long items_total, items_per_split;
  float* part;                 // [splits][Cout][Ktot]
};
constexpr int WG2_DY_BYTES = 16 * 8 * 128;     // dy box (32 channels)
// KH x KW taps: 3x3 for the residual blocks; 5x1 for the row-separable image-facing convs (x = row-expanded 32-channel
// image, only MN chunk 0 is a real tap there -- chunks 1..3 are phantoms that cost tensor time but no extra traffic)
template <int N_TILE, int G, int STAGES, int KH, int KW>
struct Wg2Smem {
  static constexpr int BW = 8 + KW - 1;                          // box exactly as wide as the column halo needs
  static constexpr int X_TX = (16 + KH - 1) * BW * 128;          // bytes of one x halo box
  static constexpr int X_BYTES = (X_TX + 1023) / 1024 * 1024;
  static constexpr int A_BYTES = G * X_BYTES;
  static constexpr int B_BYTES = (N_TILE / 32) * WG2_DY_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = (G * KH * N_TILE) <= 128 ? 128 : ((G * KH * N_TILE) <= 256 ? 256 : 512);
};
template <int N_TILE, int G, int STAGES, int KH, int KW>
__global__ void __launch_bounds__(192, 1) k_conv_wgrad_halo(const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_dy, const Wg2Params p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = Wg2Smem<N_TILE, G, STAGES, KH, KW>;
  static_assert(G * KH * N_TILE <= 512, "TMEM budget");
  constexpr int WG2_X_BYTES = SM::X_BYTES, BW = SM::BW;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = blockIdx.x;                  // group of G input-channel chunks
  const int col0 = blockIdx.y * N_TILE;       // first output channel
  const long long it_begin = (long long)blockIdx.z * p.items_per_split;
  long long it_end = it_begin + p.items_per_split;
  if (it_end > p.items_total) it_end = p.items_total;
  const int num_items = (int)(it_end > it_begin ? it_end - it_begin : 0);
  int b_chunks = (p.Cout - col0 + 31) / 32;
  if (b_chunks > N_TILE / 32) b_chunks = N_TILE / 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SM::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < num_items; ++i) {
        const int st = i % STAGES;
        mbar_wait(&empty[st], ((i / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[st], (uint32_t)(G * SM::X_TX + b_chunks * WG2_DY_BYTES));
        const long long item = it_begin + i;
        const int tw = (int)(item % p.tiles_w), th = (int)((item / p.tiles_w) % p.tiles_h);
        const int n = (int)(item / ((long long)p.tiles_w * p.tiles_h));
        const int w0 = tw * 8, h0 = th * 16;
        uint8_t* sa = smem + st * SM::STAGE_BYTES;
        for (int g = 0; g < G; ++g)
          tma_load_4d(sa + g * WG2_X_BYTES, &map_x, &full[st], (cg * G + g) * 32, w0 - KW / 2, h0 - KH / 2, n);
        uint8_t* sb = sa + SM::A_BYTES;
        for (int j = 0; j < b_chunks; ++j)
          tma_load_4d(sb + j * WG2_DY_BYTES, &map_dy, &full[st], col0 + 32 * j, w0, h0, n);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(128, N_TILE, 1, 1);
    if (elect_one())                   // one thread runs the whole issue loop (see k_conv_halo)
    for (int i = 0; i < num_items; ++i) {
      const int st = i % STAGES;
      mbar_wait(&full[st], (i / STAGES) & 1);
      {
        const uint32_t sa = smem_u32(smem + st * SM::STAGE_BYTES);
        const uint32_t sb = sa + SM::A_BYTES;
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
          for (int r = 0; r < KH; ++r) {
            const uint32_t tmem_d = tmem_base + (uint32_t)((g * KH + r) * N_TILE);
#pragma unroll 4
            for (int h = 0; h < 16; ++h) {
              // A: x halo rows (h + r) * 16 + w; MN chunks = filter columns s, 128 B (one pixel) apart; K atoms 512 B
              uint64_t ad = make_smem_desc(sa + g * WG2_X_BYTES + (uint32_t)((h + r) * BW * 128), 128, 512, 1);
              // B: dy rows h * 8 + w; MN chunks = 32-channel boxes
              uint64_t bd = make_smem_desc(sb + (uint32_t)(h * 8 * 128), WG2_DY_BYTES, 512, 1);
              umma_tf32(tmem_d, ad, bd, idesc, (i > 0 || h > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty[st]);
        if (i == num_items - 1) umma_commit(tmem_full);
      }
    }
  } else {
    const int q = warp & 3;          // = filter column s of this warp's 32 accumulator rows; s == 3 is the phantom tap
    if (num_items > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    if (q < KW) {
      float* dst = p.part + (long long)blockIdx.z * p.Cout * p.Ktot;
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        const int ci = (cg * G + g) * 32 + lane;
#pragma unroll 1
        for (int r = 0; r < KH; ++r) {
          const int kidx = (r * KW + q) * p.Cin + ci;
#pragma unroll 1
          for (int c = 0; c < N_TILE; c += 32) {
            if (col0 + c >= p.Cout) break;
            uint32_t v[32];
            if (num_items > 0) {
              tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g * KH + r) * N_TILE + c), v);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = col0 + c + j;
              if (co < p.Cout) dst[(long long)co * p.Ktot + kidx] = __uint_as_float(v[j]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc<SM::TMEM_COLS>(tmem_base); }
}
static int wgrad_halo_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SIVAE_TC_WGHALO");
    v = e ? atoi(e) : 1;
  }
  return v;
}
static bool wgrad_halo_supported(const ConvShape& s) {
  return wgrad_halo_mode() != 0 && s.k == 3 && s.H % 16 == 0 && s.W % 8 == 0 && s.Cin % 32 == 0 && s.Cout % 32 == 0;
}
struct Wg2Plan { int n_tile, g, cgroups, ntiles, splits; long long items, per; };
static Wg2Plan wg2_plan(const ConvShape& s) {
  Wg2Plan pl;
  pl.n_tile = s.Cout >= 128 ? 128 : 64;
  pl.g = (pl.n_tile == 64 && s.Cin % 64 == 0) ? 2 : 1;
  pl.cgroups = s.Cin / (32 * pl.g);
  pl.ntiles = (s.Cout + pl.n_tile - 1) / pl.n_tile;
  pl.items = (long long)s.N * (s.H / 16) * (s.W / 8);
  // splits: minimise (waves of 1-CTA-per-SM blocks) x (pixel tiles per CTA + fixed per-CTA cost).  Rounding the split count UP to cover all
  // SMs (the first version) put e.g. 256->256 at 64x64 on 160 CTAs = two waves of 103 items instead of one wave of 114.
  const long long pairs = (long long)pl.cgroups * pl.ntiles;
  long long want = 1, best = -1;
  for (long long w = 1; w <= pl.items && w <= 2 * num_sms(); ++w) {
    const long long per = (pl.items + w - 1) / w;
    const long long waves = (pairs * ((pl.items + per - 1) / per) + num_sms() - 1) / num_sms();
    const long long cost = waves * (per + 16);        // + fixed cost per CTA (accumulator write-out, reduce) in item units
    if (best < 0 || cost < best) { best = cost; want = w; }
  }
  pl.per = (pl.items + want - 1) / want;
  pl.splits = (int)((pl.items + pl.per - 1) / pl.per);
  return pl;
}
static size_t wg2_scratch_bytes(const ConvShape& s) { return (size_t)wg2_plan(s).splits * s.Cout * s.ktot() * sizeof(float); }
template <int N_TILE, int G, int STAGES, int KH = 3, int KW = 3>
static int launch_wg2_t(const float* x, const float* dy, const ConvShape& s, const Wg2Plan& pl, float* part, cudaStream_t st) {
  using SM = Wg2Smem<N_TILE, G, STAGES, KH, KW>;
  static_assert(SM::TOTAL <= 232448, "shared memory budget exceeded");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad_halo<N_TILE, G, STAGES, KH, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  Wg2Params p;
  p.N = s.N; p.H = s.H; p.W = s.W; p.Cin = s.Cin; p.Cout = s.Cout; p.Ktot = KH * KW * s.Cin;
  p.tiles_w = s.W / 8; p.tiles_h = s.H / 16;
  p.items_total = pl.items; p.items_per_split = pl.per;
  p.part = part;
  CUtensorMap mx, mdy;
  int r = make_map_nhwc(&mx, x, s.N, s.H, s.W, s.Cin, SM::BW, 16 + KH - 1, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (r) return r;
  r = make_map_nhwc(&mdy, dy, s.N, s.H, s.W, s.Cout, 8, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (r) return r;
  dim3 grid(pl.cgroups, pl.ntiles, pl.splits);
  g_launches += 2;
  k_conv_wgrad_halo<N_TILE, G, STAGES, KH, KW><<<grid, 192, SM::TOTAL, st>>>(mx, mdy, p);
  return (int)cudaGetLastError();
}

int launch_conv_wgrad_tc(const float* x, const float* dy, float* dw, const ConvShape& s, bool accumulate, void* scratch,
                         size_t scratch_bytes, cudaStream_t st) {
  if (wgrad_halo_supported(s)) {
    Wg2Plan pl = wg2_plan(s);
    if (scratch_bytes < (size_t)pl.splits * s.Cout * s.ktot() * sizeof(float)) return -103;
    int r2;
    if (pl.n_tile == 128) r2 = launch_wg2_t<128, 1, 2>(x, dy, s, pl, (float*)scratch, st);
    else if (pl.g == 2) r2 = launch_wg2_t<64, 2, 2>(x, dy, s, pl, (float*)scratch, st);
    else r2 = launch_wg2_t<64, 1, 3>(x, dy, s, pl, (float*)scratch, st);
    if (r2) return r2;
    long long n2 = (long long)s.Cout * s.ktot();
    unsigned blocks2 = (unsigned)((n2 + 255) / 256);
    if (blocks2 > 148u * 8) blocks2 = 148u * 8;
    k_wg_reduce<<<blocks2, 256, 0, st>>>((const float*)scratch, dw, n2, pl.splits, accumulate ? 1 : 0);
    return (int)cudaGetLastError();
  }
  WgParams p;
  p.N = s.N; p.H = s.H; p.W = s.W; p.Cin = s.Cin; p.Cout = s.Cout; p.ks = s.k;
  pick_pixel_block(s.H, s.W, &p.pw, &p.ph, &p.pn);
  p.tiles_w = s.W / p.pw; p.tiles_h = s.H / p.ph;
  int splits;
  wg_plan(s, &splits, &p.kb_total, &p.kb_per_split);
  p.Ktot = (int)s.ktot();
  p.part = (float*)scratch;
  if (scratch_bytes < (size_t)splits * s.Cout * s.ktot() * sizeof(float)) return -103;
  CUtensorMap mx, mdy;
  int r = make_map_nhwc(&mx, x, s.N, s.H, s.W, s.Cin, p.pw, p.ph, p.pn, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (r) return r;
  r = make_map_nhwc(&mdy, dy, s.N, s.H, s.W, s.Cout, p.pw, p.ph, p.pn, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (r) return r;
  const int bn = wg_block_n(s.Cout);
  if (bn == 256) r = launch_wg_t<256, 3>(mx, mdy, p, splits, st);
  else if (bn == 128) r = launch_wg_t<128, 4>(mx, mdy, p, splits, st);
  else r = launch_wg_t<64, 4>(mx, mdy, p, splits, st);
  if (r) return r;
  long long n = (long long)s.Cout * s.ktot();
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148u * 8) blocks = 148u * 8;
  k_wg_reduce<<<blocks, 256, 0, st>>>(p.part, dw, n, splits, accumulate ? 1 : 0);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 wgrad (residual blocks in the default mode): x = the bf16 copy of the forward activation, dy = the bf16 gradient that
// the BatchNorm backward writes; plain NHWC bf16, 64 channels per 128-byte pixel row, both operands MN-major with the
// ordinary 128B swizzle (16-bit MN-major needs no special atom), K = 16 pixels per kind::f16 MMA = half the tf32 tensor time.
//
// k_conv_wgrad_halo16 (3x3, H % 16 == 0, W % 8 == 0): one 16x8 pixel tile per item, ONE x halo box {64c, 10w, 18h} per item
// and dy tile {64c, 8w, 16h}; for filter row r and the image-row PAIR (h, h+1) one MMA with
//     A = x halo from pixel row (h + r) * 10 + 2p: two MN chunks 128 B (one pixel = one filter column) apart -> taps s = 2p, 2p+1
//         of 64 input channels (M = 128), two K groups (the 8 pixels of image row h + r, resp. h + r + 1) 1280 B apart;
//     B = dy tile from pixel row h * 8: 64 output channels, K groups 1024 B apart.
// p = 0 gives taps 0,1; p = 1 gives tap 2 plus a phantom tap (its accumulator rows are never read).  Accumulators of all
// (r, p) stay in TMEM (6 x 64 columns) over the CTA's pixel range; split over pixel tiles, partials reduced in fixed order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG16_DY_BYTES = 16 * 8 * 128;
template <int STAGES>
struct Wg16Smem {
  static constexpr int BW = 10;
  static constexpr int X_TX = 18 * BW * 128;
  static constexpr int X_BYTES = (X_TX + 1023) / 1024 * 1024 + 1024;     // + slack: the phantom tap reads 1 pixel past the box
  static constexpr int STAGE_BYTES = X_BYTES + WG16_DY_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 512;      // 3 filter rows x 2 tap pairs x 64 columns = 384
};
template <int STAGES>
__global__ void __launch_bounds__(192, 1) k_conv_wgrad_halo16(const __grid_constant__ CUtensorMap map_x,
                                                              const __grid_constant__ CUtensorMap map_dy, const Wg2Params p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = Wg16Smem<STAGES>;
  constexpr int BW = SM::BW, NT = 64;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cg = blockIdx.x;                  // 64-channel input chunk
  const int col0 = blockIdx.y * NT;           // first output channel
  const long long it_begin = (long long)blockIdx.z * p.items_per_split;
  long long it_end = it_begin + p.items_per_split;
  if (it_end > p.items_total) it_end = p.items_total;
  const int num_items = (int)(it_end > it_begin ? it_end - it_begin : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<SM::TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < num_items; ++i) {
        const int st = i % STAGES;
        mbar_wait(&empty[st], ((i / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[st], (uint32_t)(SM::X_TX + WG16_DY_BYTES));
        const long long item = it_begin + i;
        const int tw = (int)(item % p.tiles_w), th = (int)((item / p.tiles_w) % p.tiles_h);
        const int n = (int)(item / ((long long)p.tiles_w * p.tiles_h));
        uint8_t* sa = smem + st * SM::STAGE_BYTES;
        tma_load_4d(sa, &map_x, &full[st], cg * 64, tw * 8 - 1, th * 16 - 1, n);
        tma_load_4d(sa + SM::X_BYTES, &map_dy, &full[st], col0, tw * 8, th * 16, n);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, NT, 1, 1);
    if (elect_one())
    for (int i = 0; i < num_items; ++i) {
      const int st = i % STAGES;
      mbar_wait(&full[st], (i / STAGES) & 1);
      const uint32_t sa = smem_u32(smem + st * SM::STAGE_BYTES);
      const uint32_t sb = sa + SM::X_BYTES;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const uint32_t tmem_d = tmem_base + (uint32_t)((r * 2 + pr) * NT);
#pragma unroll
          for (int hp = 0; hp < 8; ++hp) {
            const int h = 2 * hp;
            // MN-major, 128B swizzle: LBO = distance of the 64-wide MN chunks, SBO = distance of the 8-row K groups
            const uint64_t ad = make_smem_desc(sa + (uint32_t)(((h + r) * BW + 2 * pr) * 128), 128, BW * 128, 2);
            const uint64_t bd = make_smem_desc(sb + (uint32_t)(h * 8 * 128), WG16_DY_BYTES, 1024, 2);
            umma_f16(tmem_d, ad, bd, idesc, (i > 0 || hp > 0) ? 1u : 0u);
          }
        }
      }
      umma_commit(&empty[st]);
      if (i == num_items - 1) umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;          // accumulator rows 32q .. 32q+31: tap (2 pr + q / 2), input channel (q & 1) * 32 + lane
    if (num_items > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    float* dst = p.part + (long long)blockIdx.z * p.Cout * p.Ktot;
    const int ci = cg * 64 + (q & 1) * 32 + lane;
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
#pragma unroll 1
      for (int pr = 0; pr < 2; ++pr) {
        const int s_tap = 2 * pr + (q >> 1);
        if (s_tap > 2) continue;                    // phantom tap
        const int kidx = (r * 3 + s_tap) * p.Cin + ci;
#pragma unroll 1
        for (int c = 0; c < NT; c += 32) {
          if (col0 + c >= p.Cout) break;
          uint32_t v[32];
          if (num_items > 0) {
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((r * 2 + pr) * NT + c), v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int co = col0 + c + j;
            if (co < p.Cout) dst[(long long)co * p.Ktot + kidx] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc<SM::TMEM_COLS>(tmem_base); }
}
// generic bf16 wgrad (1x1 convs, maps smaller than a 16x8 tile): K block = 32 pixels (two MMAs), A = two 64-row (tap, ci) chunks
// (Cin % 64 == 0 keeps a chunk inside one tap), B = BLOCK_N / 64 chunks of dy
template <int BLOCK_N, int STAGES>
struct Wg16tSmem {
  static constexpr int A_BYTES = 2 * WG_CHUNK_BYTES;
  static constexpr int B_BYTES = (BLOCK_N / 64) * WG_CHUNK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
};
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(192) k_conv_wgrad_tc16(const __grid_constant__ CUtensorMap map_x,
                                                         const __grid_constant__ CUtensorMap map_dy, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  using SM = Wg16tSmem<BLOCK_N, STAGES>;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;            // first (tap, ci) row of this tile
  const int col0 = blockIdx.y * BLOCK_N;
  const long long kb_begin = (long long)blockIdx.z * p.kb_per_split;
  long long kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.kb_total) kb_end = p.kb_total;
  const int num_kb = (int)(kb_end > kb_begin ? kb_end - kb_begin : 0);
  const int pad = p.ks >> 1;
  int a_chunks = (p.Ktot - m0 + 63) / 64;
  if (a_chunks > 2) a_chunks = 2;
  int b_chunks = (p.Cout - col0 + 63) / 64;
  if (b_chunks > BLOCK_N / 64) b_chunks = BLOCK_N / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<BLOCK_N>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < num_kb; ++i) {
        const int st = i % STAGES;
        mbar_wait(&empty[st], ((i / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full[st], (uint32_t)(a_chunks + b_chunks) * WG_CHUNK_BYTES);
        const long long kb = kb_begin + i;
        const int tw = (int)(kb % p.tiles_w);
        const int th = (int)((kb / p.tiles_w) % p.tiles_h);
        const int tn = (int)(kb / ((long long)p.tiles_w * p.tiles_h));
        const int w0 = tw * p.pw, h0 = th * p.ph, n0 = tn * p.pn;
        uint8_t* sa = smem + st * SM::STAGE_BYTES;
        for (int j = 0; j < a_chunks; ++j) {
          const int m = m0 + 64 * j;
          const int tap = m / p.Cin, c0 = m - tap * p.Cin;
          const int r = tap / p.ks, s = tap - r * p.ks;
          tma_load_4d(sa + j * WG_CHUNK_BYTES, &map_x, &full[st], c0, w0 + s - pad, h0 + r - pad, n0);
        }
        uint8_t* sb = sa + SM::A_BYTES;
        for (int j = 0; j < b_chunks; ++j)
          tma_load_4d(sb + j * WG_CHUNK_BYTES, &map_dy, &full[st], col0 + 64 * j, w0, h0, n0);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
    if (elect_one())
    for (int i = 0; i < num_kb; ++i) {
      const int st = i % STAGES;
      mbar_wait(&full[st], (i / STAGES) & 1);
      const uint32_t sa = smem_u32(smem + st * SM::STAGE_BYTES);
      const uint32_t sb = sa + SM::A_BYTES;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        // 16 pixels (K) per MMA = two 8-row groups 1024 B apart; MN chunks (64 rows / channels) one box (4096 B) apart
        const uint64_t ad = make_smem_desc(sa + k * 2048, WG_CHUNK_BYTES, 1024, 2);
        const uint64_t bd = make_smem_desc(sb + k * 2048, WG_CHUNK_BYTES, 1024, 2);
        umma_f16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty[st]);
      if (i == num_kb - 1) umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool valid = m < p.Ktot && num_kb > 0 && (q >> 1) < a_chunks;
    float* dst = p.part + (long long)blockIdx.z * p.Cout * p.Ktot;
    if (num_kb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (m < p.Ktot) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = col0 + c + j;
          if (co < p.Cout) dst[(long long)co * p.Ktot + m] = valid ? __uint_as_float(v[j]) : 0.f;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc<BLOCK_N>(tmem_base); }
}
bool conv_tc_supported_wgrad16(const ConvShape& s) {
  if (s.Cin % 64 != 0 || s.Cout % 64 != 0 || (s.k != 1 && s.k != 3)) return false;
  int pw, ph, pn;
  pick_pixel_block(s.H, s.W, &pw, &ph, &pn);
  return pn <= 256;
}
static bool wgrad_halo16_supported(const ConvShape& s) {
  return wgrad_halo_mode() != 0 && s.k == 3 && s.H % 16 == 0 && s.W % 8 == 0;
}
static Wg2Plan wg16_plan(const ConvShape& s) {
  Wg2Plan pl;
  pl.n_tile = 64; pl.g = 1;
  pl.cgroups = s.Cin / 64;
  pl.ntiles = s.Cout / 64;
  pl.items = (long long)s.N * (s.H / 16) * (s.W / 8);
  const long long pairs = (long long)pl.cgroups * pl.ntiles;
  long long want = 1, best = -1;
  for (long long w = 1; w <= pl.items && w <= 2 * num_sms(); ++w) {
    const long long per = (pl.items + w - 1) / w;
    const long long waves = (pairs * ((pl.items + per - 1) / per) + num_sms() - 1) / num_sms();
    const long long cost = waves * (per + 16);
    if (best < 0 || cost < best) { best = cost; want = w; }
  }
  pl.per = (pl.items + want - 1) / want;
  pl.splits = (int)((pl.items + pl.per - 1) / pl.per);
  return pl;
}
static void wg16t_plan(const ConvShape& s, int* bn, int* splits, long long* kb_total, long long* kb_per_split) {
  int pw, ph, pn;
  pick_pixel_block(s.H, s.W, &pw, &ph, &pn);
  const long long kbt = (long long)(s.W / pw) * (s.H / ph) * ((s.N + pn - 1) / pn);
  *bn = s.Cout > 128 ? 256 : (s.Cout > 64 ? 128 : 64);
  const long long tiles = (long long)((s.ktot() + 127) / 128) * ((s.Cout + *bn - 1) / *bn);
  long long want = (num_sms() * 2 + tiles - 1) / tiles;
  const long long maxs = (kbt + 7) / 8;       // at least 8 pixel blocks per CTA
  long long sp = want < maxs ? want : maxs;
  if (sp < 1) sp = 1;
  if (sp > 512) sp = 512;
  const long long per = (kbt + sp - 1) / sp;
  sp = (kbt + per - 1) / per;
  *splits = (int)sp; *kb_total = kbt; *kb_per_split = per;
}
size_t conv_wgrad_tc16_scratch_bytes(const ConvShape& s) {
  if (wgrad_halo16_supported(s)) return (size_t)wg16_plan(s).splits * s.Cout * s.ktot() * sizeof(float);
  int bn, splits; long long kbt, per;
  wg16t_plan(s, &bn, &splits, &kbt, &per);
  return (size_t)splits * s.Cout * s.ktot() * sizeof(float);
}
template <int BLOCK_N, int STAGES>
static int launch_wg16t(const CUtensorMap& mx, const CUtensorMap& mdy, const WgParams& p, int splits, cudaStream_t st) {
  using SM = Wg16tSmem<BLOCK_N, STAGES>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad_tc16<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid((p.Ktot + 127) / 128, (p.Cout + BLOCK_N - 1) / BLOCK_N, splits);
  g_launches += 2;
  k_conv_wgrad_tc16<BLOCK_N, STAGES><<<grid, 192, SM::TOTAL, st>>>(mx, mdy, p);
  return (int)cudaGetLastError();
}
// x: bf16 NHWC [N,H,W,Cin], dy: bf16 NHWC [N,H,W,Cout]; dw fp32 [Cout][k][k][Cin] (+= if accumulate)
int launch_conv_wgrad_tc16(const void* x, const void* dy, float* dw, const ConvShape& s, bool accumulate, void* scratch,
                           size_t scratch_bytes, cudaStream_t st) {
  if (!conv_tc_supported_wgrad16(s)) return -9;
  const long long n = (long long)s.Cout * s.ktot();
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148u * 8) blocks = 148u * 8;
  CUtensorMap mx, mdy;
  if (wgrad_halo16_supported(s)) {
    using SM = Wg16Smem<4>;
    static_assert(SM::TOTAL <= 232448, "shared memory budget exceeded");
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad_halo16<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
      if (e != cudaSuccess) return (int)e;
      attr = true;
    }
    const Wg2Plan pl = wg16_plan(s);
    if (scratch_bytes < (size_t)pl.splits * n * sizeof(float)) return -103;
    Wg2Params p;
    p.N = s.N; p.H = s.H; p.W = s.W; p.Cin = s.Cin; p.Cout = s.Cout; p.Ktot = 9 * s.Cin;
    p.tiles_w = s.W / 8; p.tiles_h = s.H / 16;
    p.items_total = pl.items; p.items_per_split = pl.per;
    p.part = (float*)scratch;
    int r = make_map_nhwc(&mx, x, s.N, s.H, s.W, s.Cin, SM::BW, 18, 1, CU_TENSOR_MAP_SWIZZLE_128B, true);
    if (r) return r;
    r = make_map_nhwc(&mdy, dy, s.N, s.H, s.W, s.Cout, 8, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B, true);
    if (r) return r;
    dim3 grid(pl.cgroups, pl.ntiles, pl.splits);
    g_launches += 2;
    k_conv_wgrad_halo16<4><<<grid, 192, SM::TOTAL, st>>>(mx, mdy, p);
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return (int)ce;
    k_wg_reduce<<<blocks, 256, 0, st>>>((const float*)scratch, dw, n, pl.splits, accumulate ? 1 : 0);
    return (int)cudaGetLastError();
  }
  WgParams p;
  p.N = s.N; p.H = s.H; p.W = s.W; p.Cin = s.Cin; p.Cout = s.Cout; p.ks = s.k;
  pick_pixel_block(s.H, s.W, &p.pw, &p.ph, &p.pn);
  p.tiles_w = s.W / p.pw; p.tiles_h = s.H / p.ph;
  int bn, splits;
  wg16t_plan(s, &bn, &splits, &p.kb_total, &p.kb_per_split);
  p.Ktot = (int)s.ktot();
  p.part = (float*)scratch;
  if (scratch_bytes < (size_t)splits * n * sizeof(float)) return -103;
  int r = make_map_nhwc(&mx, x, s.N, s.H, s.W, s.Cin, p.pw, p.ph, p.pn, CU_TENSOR_MAP_SWIZZLE_128B, true);
  if (r) return r;
  r = make_map_nhwc(&mdy, dy, s.N, s.H, s.W, s.Cout, p.pw, p.ph, p.pn, CU_TENSOR_MAP_SWIZZLE_128B, true);
  if (r) return r;
  if (bn == 256) r = launch_wg16t<256, 4>(mx, mdy, p, splits, st);
  else if (bn == 128) r = launch_wg16t<128, 5>(mx, mdy, p, splits, st);
  else r = launch_wg16t<64, 6>(mx, mdy, p, splits, st);
  if (r) return r;
  k_wg_reduce<<<blocks, 256, 0, st>>>(p.part, dw, n, splits, accumulate ? 1 : 0);
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad of the row-separable image-facing convs: the 5x1 halo kernel over (A side = row-expanded narrow tensor with its
// 4-row halo, B side = the wide tensor), then one reduce that folds the split partials T[wide][5][32] into dW[Cout][5][5][Cin]:
//   mode 1 (stem,    narrow = x, wide = dy):   dW[co][r][s][ci] += T[co][r][s*c+ci]
//   mode 0 (predict, narrow = dy, wide = x):   dW[co][r][s][ci] += T[ci][4-r][s*c+co]     (mirrored expansion, see above)
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_rowsep_wg_reduce(const float* __restrict__ part, float* __restrict__ dw, int splits, int c, int wide, int mode,
                                   int accumulate) {
  const int total = c * 25 * wide;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    long long t;
    if (mode == 1) {            // dw index: ((co*5 + r)*5 + s)*c + ci  = co*25c + r*5c + j
      const int j = i % (5 * c), r = (i / (5 * c)) % 5, co = i / (25 * c);
      t = ((long long)co * 5 + r) * 32 + j;
    } else {                    // dw index: ((co*5 + r)*5 + s)*wide + ci
      const int ci = i % wide, s = (i / wide) % 5, r = (i / (5 * wide)) % 5, co = i / (25 * wide);
      t = ((long long)ci * 5 + (4 - r)) * 32 + s * c + co;
    }
    const long long stride = (long long)wide * 160;
    float acc = 0.f;
    for (int k = 0; k < splits; ++k) acc += part[k * stride + t];
    dw[i] = accumulate ? dw[i] + acc : acc;
  }
}
static Wg2Plan rowsep_wg_plan(int N, int H, int W, int wide) {
  Wg2Plan pl;
  pl.n_tile = 64; pl.g = 1; pl.cgroups = 1;
  pl.ntiles = (wide + 63) / 64;
  pl.items = (long long)N * (H / 16) * (W / 8);
  long long want = (num_sms() + pl.ntiles - 1) / pl.ntiles;
  if (want > pl.items) want = pl.items;
  if (want < 1) want = 1;
  pl.per = (pl.items + want - 1) / want;
  pl.splits = (int)((pl.items + pl.per - 1) / pl.per);
  return pl;
}
bool conv_rowsep_wgrad_supported(int H, int W, int c, int wide, int k) {
  return rowsep_enabled() && wgrad_halo_mode() != 0 && k == 5 && c >= 1 && c <= 3 && wide % 32 == 0 && W % 8 == 0 && H % 16 == 0;
}
size_t conv_rowsep_wgrad_scratch_bytes(int N, int H, int W, int wide) {
  return (size_t)rowsep_wg_plan(N, H, W, wide).splits * wide * 160 * sizeof(float);
}
int launch_conv_rowsep_wgrad(const float* narrow_t, const float* wide_t, float* dw, int N, int H, int W, int c, int wide, int mode,
                             bool accumulate, float* expand_scratch, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  Wg2Plan pl = rowsep_wg_plan(N, H, W, wide);
  if (scratch_bytes < (size_t)pl.splits * wide * 160 * sizeof(float)) return -103;
  const long long rows = (long long)N * H;
  g_launches += 2;
  if (mode == 1) k_rowsep_expand<false><<<num_sms() * 8, 256, 0, st>>>(narrow_t, expand_scratch, rows, W, c);
  else k_rowsep_expand_rev<<<num_sms() * 8, 256, 0, st>>>(narrow_t, expand_scratch, rows, W, c);
  ConvShape e{N, H, W, 32, wide, 5};
  int r = launch_wg2_t<64, 1, 4, 5, 1>(expand_scratch, wide_t, e, pl, (float*)scratch, st);
  if (r) return r;
  const int total = c * 25 * wide;
  k_rowsep_wg_reduce<<<(total + 255) / 256, 256, 0, st>>>((const float*)scratch, dw, pl.splits, c, wide, mode, accumulate ? 1 : 0);
  return (int)cudaGetLastError();
}

}  // namespace sivae
