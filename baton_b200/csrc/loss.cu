// Fused loss kernels (K7): forward + backward in one pass, with the batch-mean loss accumulated
// into a device scalar that the trainer reads ONCE per epoch (the reference does float(loss) -- a
// host sync -- on every batch: utils.py:88).
#define B200_TU_TAG 11
#include "launch.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <bool IN_FP32>
__device__ __forceinline__ float load_logit(const void* p, long long i) {
  if constexpr (IN_FP32)
    return reinterpret_cast<const float*>(p)[i];
  else
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}

// one warp per row: loss_row = logsumexp(z) - z[target]; dz = (softmax(z) - onehot) * grad_scale
// loss_acc[0] += sum(loss_row) * grad_scale ; loss_acc[1] += #correct (argmax == target)
template <bool IN_FP32, bool OUT_FP32>
__global__ void __launch_bounds__(256)
softmax_xent_kernel(const void* __restrict__ logits, const long long* __restrict__ target, void* __restrict__ dlogits,
                    float* __restrict__ loss_acc, long long rows, int C, long long ld, float grad_scale) {
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  float my_loss = 0.f, my_hit = 0.f;
  if (row < rows) {
    const long long base = row * ld;
    float m = -INFINITY;
    int am = 0;
    for (int c = lane; c < C; c += 32) {
      const float z = load_logit<IN_FP32>(logits, base + c);
      if (z > m) { m = z; am = c; }
    }
    // warp argmax (ties -> lowest index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __expf(load_logit<IN_FP32>(logits, base + c) - m);
    s = wsum(s);
    const float lse = m + __logf(s);
    const int t = static_cast<int>(target[row]);
    const float inv = 1.f / s;
    for (int c = lane; c < C; c += 32) {
      const float z = load_logit<IN_FP32>(logits, base + c);
      const float g = (__expf(z - m) * inv - (c == t ? 1.f : 0.f)) * grad_scale;
      if (dlogits != nullptr) {
        if constexpr (OUT_FP32)
          reinterpret_cast<float*>(dlogits)[base + c] = g;
        else
          reinterpret_cast<__nv_bfloat16*>(dlogits)[base + c] = __float2bfloat16_rn(g);
      }
    }
    if (lane == 0) {
      my_loss = (lse - load_logit<IN_FP32>(logits, base + t)) * grad_scale;
      my_hit = (am == t) ? 1.f : 0.f;
    }
  }
  // block-level combine -> one atomic pair per block
  __shared__ float sl[8], sh[8];
  if (lane == 0) { sl[threadIdx.x >> 5] = my_loss; sh[threadIdx.x >> 5] = my_hit; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { a += sl[w]; b += sh[w]; }
    atomicAdd(loss_acc, a);
    atomicAdd(loss_acc + 1, b);
  }
}

// MSE: loss_acc[0] += sum((p - t)^2) * grad_scale ; dp = 2 * (p - t) * grad_scale   (grad_scale = 1/numel)
template <bool IN_FP32, bool OUT_FP32>
__global__ void __launch_bounds__(256)
mse_kernel(const void* __restrict__ pred, const float* __restrict__ target, void* __restrict__ dpred,
           float* __restrict__ loss_acc, long long n, float grad_scale) {
  griddep_launch_dependents();
  griddep_wait();
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = load_logit<IN_FP32>(pred, i) - target[i];
    acc = fmaf(d, d, acc);
    if (dpred != nullptr) {
      const float g = 2.f * d * grad_scale;
      if constexpr (OUT_FP32)
        reinterpret_cast<float*>(dpred)[i] = g;
      else
        reinterpret_cast<__nv_bfloat16*>(dpred)[i] = __float2bfloat16_rn(g);
    }
  }
  acc = wsum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) a += s[w];
    atomicAdd(loss_acc, a * grad_scale);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_softmax_xent(const void* logits, int logits_fp32, const long long* target, void* dlogits,
                                 int dl_fp32, float* loss_acc, long long rows, int C, long long ld, float grad_scale,
                                 cudaStream_t stream) {
  if (rows <= 0) return 0;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
#define XENT(A, B) launch_pdl(softmax_xent_kernel<A, B>, grid, 256, 0, stream, logits, target, dlogits, loss_acc, rows, C, ld, grad_scale)
  if (logits_fp32) { if (dl_fp32) XENT(true, true); else XENT(true, false); }
  else             { if (dl_fp32) XENT(false, true); else XENT(false, false); }
#undef XENT
  return static_cast<int>(cudaGetLastError());
}

extern "C" int b200_mse(const void* pred, int pred_fp32, const float* target, void* dpred, int dp_fp32,
                        float* loss_acc, long long n, float grad_scale, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long g = (n + 255) / 256;
  if (g > 148 * 4) g = 148 * 4;
  const unsigned grid = static_cast<unsigned>(g);
#define MSE(A, B) launch_pdl(mse_kernel<A, B>, grid, 256, 0, stream, pred, target, dpred, loss_acc, n, grad_scale)
  if (pred_fp32) { if (dp_fp32) MSE(true, true); else MSE(true, false); }
  else           { if (dp_fp32) MSE(false, true); else MSE(false, false); }
#undef MSE
  return static_cast<int>(cudaGetLastError());
}

B200_TRACE_REGISTER(loss)
