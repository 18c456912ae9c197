// OCP microscaling helpers shared by the MXFP8 quantiser (quant.cu) and the fp8 wire format of the
// fused FedAvg collective (fedavg.cu): e4m3 elements with one UE8M0 power-of-two scale per 32 elements.
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <cstdint>

namespace b200 {

// shared exponent = floor(log2(amax)) - emax(e4m3 = 8), clamped to the UE8M0 range
// (bumped by one when amax / 2^e would exceed 448 = max e4m3, so the block maximum never saturates)
__device__ __forceinline__ int mx_exponent(float amax) {
  if (!(amax > 0.f)) return -127;
  const uint32_t bits = __float_as_uint(amax);
  int e = static_cast<int>((bits >> 23) & 0xFF) - 127 - 8;
  if ((bits & 0x7FFFFFu) > 0x600000u) e += 1;   // mantissa > 1.75  <=>  amax * 2^-e > 448
  return e < -127 ? -127 : (e > 127 ? 127 : e);
}
__device__ __forceinline__ float exp2_int(int e) {   // 2^e for e in [-127, 127]
  if (e <= -127) return __uint_as_float(0x00400000u);  // 2^-127 (denormal)
  return __uint_as_float(static_cast<uint32_t>(e + 127) << 23);
}
__device__ __forceinline__ uint16_t to_e4m3x2(float a, float b) {
  return static_cast<uint16_t>(__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3));
}
__device__ __forceinline__ float2 from_e4m3x2(uint16_t v) {
  const __half2_raw h = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>(v), __NV_E4M3);
  return __half22float2(*reinterpret_cast<const __half2*>(&h));
}

}  // namespace b200
