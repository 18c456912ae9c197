// CUDA-core fallback GEMM for shapes the TMA path cannot take (row pitch not a multiple of 16
// bytes, e.g. the 10-class classifier head) -- tiny problems only.  Same contract as
// b200_gemm_bf16: D = act(alpha * A B^T + bias) with K-major or MN-major bf16 operands.
#define B200_TU_TAG 6
#include "launch.h"
#include "pdl.cuh"
#include <cuda_bf16.h>

namespace b200 {

__device__ __forceinline__ float simt_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    return 0.5f * v * (1.f + tanhf(k0 * (v + k1 * v * v * v)));
  }
  return v;
}

// 32x32 output tile per CTA, 32-deep K slices through shared memory.
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, void* __restrict__ D,
                 const float* __restrict__ bias, int M, int N, int K, long long lda, long long ldb, long long ldd,
                 int a_mn, int b_mn, int out_fp32, int act, int accumulate, float alpha) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float sa[32][33];
  __shared__ float sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + i * 8;
      // sa[r][tx] = A(m0 + r, k0 + tx); choose the coalesced index order per major
      if (!a_mn) {
        const int m = m0 + r, k = k0 + tx;
        sa[r][tx] = (m < M && k < K) ? __bfloat162float(A[static_cast<long long>(m) * lda + k]) : 0.f;
      } else {
        const int m = m0 + tx, k = k0 + r;
        sa[tx][r] = (m < M && k < K) ? __bfloat162float(A[static_cast<long long>(k) * lda + m]) : 0.f;
      }
      if (!b_mn) {
        const int n = n0 + r, k = k0 + tx;
        sb[r][tx] = (n < N && k < K) ? __bfloat162float(B[static_cast<long long>(n) * ldb + k]) : 0.f;
      } else {
        const int n = n0 + tx, k = k0 + r;
        sb[tx][r] = (n < N && k < K) ? __bfloat162float(B[static_cast<long long>(k) * ldb + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float bv = sb[tx][k];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sa[ty + i * 8][k], bv, acc[i]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n >= N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + i * 8;
    if (m >= M) continue;
    float v = acc[i] * alpha;
    if (bias != nullptr) v += bias[n];
    v = simt_act(v, act);
    const long long o = static_cast<long long>(m) * ldd + n;
    if (out_fp32) {
      float* d = reinterpret_cast<float*>(D);
      d[o] = accumulate ? d[o] + v : v;
    } else {
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(D);
      d[o] = __float2bfloat16_rn(accumulate ? __bfloat162float(d[o]) + v : v);
    }
  }
}

}  // namespace b200

extern "C" int b200_gemm_simt(const void* a, const void* b, void* d, const float* bias, int M, int N, int K,
                              long long lda, long long ldb, long long ldd, int a_mn, int b_mn, int out_fp32, int act,
                              int accumulate, float alpha, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  b200::launch_pdl(b200::gemm_simt_kernel, grid, 256, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), d, bias, M, N, K, lda,
      ldb, ldd, a_mn, b_mn, out_fp32, act, accumulate, alpha);
  return static_cast<int>(cudaGetLastError());
}

B200_TRACE_REGISTER(gemm_simt)
