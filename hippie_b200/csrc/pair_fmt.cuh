// 16-bit pair planes: x ~= hi + lo with hi = round16(x), lo = round16(x - hi).
//   kPairF16  : fp16 pair, ~22 mantissa bits for |x| >= 2^-3 (absolute error <= 2^-25 below); saturates at 65504
//   kPairBF16 : bf16 pair, 16 mantissa bits, fp32 exponent range (tools/pair_test; the engine uses fp16 pairs throughout)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace hp {

enum { kPairF16 = 0, kPairBF16 = 1 };
constexpr float kWeightPairScale = 256.f;  // weights are stored as fp16 pairs of w * 2^8 (tools/pair_precision.py)

template <int FMT>
__device__ __forceinline__ void pair_split(float x, uint16_t& h, uint16_t& l) {
  if (FMT == kPairF16) {
    uint16_t hh, ll;
    float back;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hh) : "f"(x));
    asm("cvt.f32.f16 %0, %1;" : "=f"(back) : "h"(hh));
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(ll) : "f"(x - back));
    h = hh, l = ll;
  } else {
    const __nv_bfloat16 hb = __float2bfloat16_rn(x);
    const __nv_bfloat16 lb = __float2bfloat16_rn(x - __bfloat162float(hb));
    h = __bfloat16_as_ushort(hb), l = __bfloat16_as_ushort(lb);
  }
}

constexpr float kPairF16Max = 65520.f;  // |x| >= this rounds past the largest finite fp16: cvt.satfinite clamps it

// 4 consecutive channels -> fp16 pair planes (hi at p, lo at p + ps), 8 bytes per plane.  `flags` (optional): device
// word that receives `bit` when a value saturates the hi plane (the product then silently loses its magnitude).
__device__ __forceinline__ void store_pair4(uint16_t* p, int64_t ps, int64_t idx, float4 v, unsigned* flags = nullptr,
                                            unsigned bit = 4u) {
  if (flags && !(fabsf(v.x) < kPairF16Max && fabsf(v.y) < kPairF16Max && fabsf(v.z) < kPairF16Max && fabsf(v.w) < kPairF16Max))
    atomicOr(flags, bit);  // also catches NaN / Inf
  uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
  pair_split<kPairF16>(v.x, h0, l0), pair_split<kPairF16>(v.y, h1, l1);
  pair_split<kPairF16>(v.z, h2, l2), pair_split<kPairF16>(v.w, h3, l3);
  *reinterpret_cast<uint2*>(p + idx) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
  *reinterpret_cast<uint2*>(p + ps + idx) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
}

}  // namespace hp
