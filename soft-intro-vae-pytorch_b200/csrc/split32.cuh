// "split32": the compensated 16-bit operand format of the forward convolutions.
// A tensor whose innermost (channel) dimension is a multiple of 32 keeps its fp32 byte geometry; every aligned group of 32
// channels (128 bytes) holds  [32 x bf16 hi | 32 x bf16 lo]  with hi = bf16_rn(x), lo = bf16_rn(x - hi):  hi + lo carries 16
// significand bits of x (relative error <= 2^-17) over the full fp32 exponent range, and the tensor core consumes hi and lo
// as separate K slices of the same 128-byte swizzle row (conv_tc.cu, FMT_SPLIT).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

namespace sivae {

__device__ __forceinline__ void split32_one(float x, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
  hi = (uint32_t)__bfloat16_as_ushort(h);
  lo = (uint32_t)__bfloat16_as_ushort(l);
}
// four consecutive channels -> 8 bytes of hi parts + 8 bytes of lo parts
__device__ __forceinline__ void split32_pack4(const float4& v, uint2& hi, uint2& lo) {
  uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
  split32_one(v.x, h0, l0); split32_one(v.y, h1, l1); split32_one(v.z, h2, l2); split32_one(v.w, h3, l3);
  hi = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
  lo = make_uint2(l0 | (l1 << 16), l2 | (l3 << 16));
}
__device__ __forceinline__ float4 split32_unpack4(const uint2& hi, const uint2& lo) {
  float4 r;
  r.x = __uint_as_float(hi.x << 16) + __uint_as_float(lo.x << 16);
  r.y = __uint_as_float(hi.x & 0xFFFF0000u) + __uint_as_float(lo.x & 0xFFFF0000u);
  r.z = __uint_as_float(hi.y << 16) + __uint_as_float(lo.y << 16);
  r.w = __uint_as_float(hi.y & 0xFFFF0000u) + __uint_as_float(lo.y & 0xFFFF0000u);
  return r;
}
// byte offset, inside a row of C channels, of the hi parts of channels [4*c4, 4*c4 + 4); the lo parts sit 64 bytes further
__device__ __forceinline__ uint32_t split32_off4(uint32_t c4) { return (c4 >> 3) * 128u + (c4 & 7u) * 8u; }
// store / load one float4 (channels 4*c4 ..) of the row that starts at `row` (a split32 tensor, passed as its fp32-sized buffer)
__device__ __forceinline__ void split32_store4(float* row, uint32_t c4, const float4& v) {
  uint2 hi, lo;
  split32_pack4(v, hi, lo);
  char* p = reinterpret_cast<char*>(row) + split32_off4(c4);
  *reinterpret_cast<uint2*>(p) = hi;
  *reinterpret_cast<uint2*>(p + 64) = lo;
}
__device__ __forceinline__ float4 split32_load4(const float* row, uint32_t c4) {
  const char* p = reinterpret_cast<const char*>(row) + split32_off4(c4);
  const uint2 hi = __ldg(reinterpret_cast<const uint2*>(p));
  const uint2 lo = __ldg(reinterpret_cast<const uint2*>(p + 64));
  return split32_unpack4(hi, lo);
}

}  // namespace sivae
