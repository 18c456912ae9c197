// HBM-bound elementwise / reduction kernels over flat buffers: fused arena SGD (K4), multi-source
// weighted sum (manager-side FedAvg, K1 fallback), casts, batch row gather (K8), column sums
// (bias gradients), ReLU / GELU pieces.  All use 16-byte vectors and grid-stride loops sized to
// 148 SMs x a few resident CTAs.
#define B200_TU_TAG 8
#include "launch.h"
#include "pdl.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int EW_THREADS = 256;
static inline int ew_grid(long long n_vec, int max_ctas = 148 * 8) {
  long long g = (n_vec + EW_THREADS - 1) / EW_THREADS;
  if (g < 1) g = 1;
  if (g > max_ctas) g = max_ctas;
  return static_cast<int>(g);
}

// ------------------------------------------------------------------ fused SGD over the arena
// One pass over {w, g, m}: g' = g + wd*w ; m = mu*m + (1-damp)*g' ; step = nesterov ? g' + mu*m : m ;
// w -= lr*step ; g = 0 (so split-K wgrad GEMMs can red.add into it next step) ; bf16 shadow = bf16(w).
// Hyper-parameters come from device memory so a captured CUDA graph can be replayed with a new lr.
// K4 "emit the upload copy" (SURVEY 2.6): the LAST step of a local epoch can also write this client's wire copy for the
// round-end collective -- bf16 / fp32 of (w_new - global) * scale -- while w_new is still in registers, so the
// collective's own pack phase (one more read of theta + global, ~10 B / element) disappears.  The wire address is
// read from a device word (`wire_slot`): the collective double-buffers its wire by round parity and the captured
// epoch graph must follow without being re-captured.  Elements [n, n_pack) are float BUFFERS (BatchNorm running
// statistics): not optimised, only packed.
struct SgdPack {
  const unsigned long long* wire_slot;   // device word holding the wire base address, or nullptr = no pack
  const float* global_w;                 // delta upload: wire = w - global_w; nullptr: wire = w
  const float* scale;                    // device scalar multiplied into the wire value (NVLS: n_k; P2P: 1)
  long long n_pack;                      // elements to pack (>= n)
  int wire_fp32;                         // 0: bf16 wire, 1: fp32 wire
};

__global__ void __launch_bounds__(EW_THREADS)
fused_sgd_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ mom,
                 __nv_bfloat16* __restrict__ wb, long long n, const float* __restrict__ hyper, int zero_grad,
                 int nesterov, const SgdPack pk) {
  griddep_launch_dependents();
  griddep_wait();
  const float lr = hyper[0], mu = hyper[1], wd = hyper[2], damp = hyper[3];
  const long long nv = n >> 2;
  uint8_t* wire = nullptr;
  float pscale = 1.f;
  if (pk.wire_slot != nullptr) {
    wire = reinterpret_cast<uint8_t*>(*pk.wire_slot);
    if (pk.scale != nullptr) pscale = *pk.scale;
  }
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[i];
    float4 gv = reinterpret_cast<float4*>(g)[i];
    gv.x = fmaf(wd, wv.x, gv.x); gv.y = fmaf(wd, wv.y, gv.y);
    gv.z = fmaf(wd, wv.z, gv.z); gv.w = fmaf(wd, wv.w, gv.w);
    float4 st = gv;
    if (mom != nullptr) {
      float4 mv = reinterpret_cast<float4*>(mom)[i];
      const float od = 1.f - damp;
      mv.x = fmaf(mu, mv.x, od * gv.x); mv.y = fmaf(mu, mv.y, od * gv.y);
      mv.z = fmaf(mu, mv.z, od * gv.z); mv.w = fmaf(mu, mv.w, od * gv.w);
      reinterpret_cast<float4*>(mom)[i] = mv;
      if (nesterov) {
        st.x = fmaf(mu, mv.x, gv.x); st.y = fmaf(mu, mv.y, gv.y);
        st.z = fmaf(mu, mv.z, gv.z); st.w = fmaf(mu, mv.w, gv.w);
      } else {
        st = mv;
      }
    }
    wv.x = fmaf(-lr, st.x, wv.x); wv.y = fmaf(-lr, st.y, wv.y);
    wv.z = fmaf(-lr, st.z, wv.z); wv.w = fmaf(-lr, st.w, wv.w);
    reinterpret_cast<float4*>(w)[i] = wv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wb != nullptr) reinterpret_cast<uint2*>(wb)[i] = make_uint2(pack_bf16x2(wv.x, wv.y), pack_bf16x2(wv.z, wv.w));
    if (wire != nullptr) {
      float4 d = wv;
      if (pk.global_w != nullptr) {
        const float4 gl = reinterpret_cast<const float4*>(pk.global_w)[i];
        d.x -= gl.x; d.y -= gl.y; d.z -= gl.z; d.w -= gl.w;
      }
      d.x *= pscale; d.y *= pscale; d.z *= pscale; d.w *= pscale;
      if (pk.wire_fp32) reinterpret_cast<float4*>(wire)[i] = d;
      else reinterpret_cast<uint2*>(wire)[i] = make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
    }
  }
  if (wire != nullptr) {      // float buffers behind the parameters: pack only (n and n_pack are multiples of 8)
    const long long npv = pk.n_pack >> 2;
    for (long long i = nv + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npv;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      float4 d = reinterpret_cast<const float4*>(w)[i];
      if (pk.global_w != nullptr) {
        const float4 gl = reinterpret_cast<const float4*>(pk.global_w)[i];
        d.x -= gl.x; d.y -= gl.y; d.z -= gl.z; d.w -= gl.w;
      }
      d.x *= pscale; d.y *= pscale; d.z *= pscale; d.w *= pscale;
      if (pk.wire_fp32) reinterpret_cast<float4*>(wire)[i] = d;
      else reinterpret_cast<uint2*>(wire)[i] = make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
    }
  }
  // scalar tail (n is normally padded to a multiple of 4 by the arena)
  if (blockIdx.x == 0) {
    for (long long i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float gv = fmaf(wd, w[i], g[i]);
      float st = gv;
      if (mom != nullptr) {
        float mv = fmaf(mu, mom[i], (1.f - damp) * gv);
        mom[i] = mv;
        st = nesterov ? fmaf(mu, mv, gv) : mv;
      }
      float wv = fmaf(-lr, st, w[i]);
      w[i] = wv;
      if (zero_grad) g[i] = 0.f;
      if (wb != nullptr) wb[i] = __float2bfloat16_rn(wv);
    }
  }
}

// ------------------------------------------------------------------ logical-client fold (time-sliced clients on one GPU)
// One pass per co-resident logical client:  acc (+)= n_k * (theta - global)   and, if another client follows on this GPU,
// reset the replica to the global model (theta = global, bf16 shadow, momentum = 0).  mode: 0 = accumulate, 1 = first
// client (acc = ...), 2 = finish: theta = global + acc / total (no accumulate; `nk` carries 1 / total).
__global__ void __launch_bounds__(EW_THREADS)
fold_client_kernel(float* __restrict__ acc, float* __restrict__ theta, const float* __restrict__ global_w,
                   __nv_bfloat16* __restrict__ wb, float* __restrict__ mom, long long n_mom, long long n, float nk,
                   int mode, int reset) {
  griddep_launch_dependents();
  griddep_wait();
  const long long nv = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(global_w)[i];
    if (mode == 2) {
      const float4 a = reinterpret_cast<const float4*>(acc)[i];
      reinterpret_cast<float4*>(theta)[i] = make_float4(fmaf(a.x, nk, g.x), fmaf(a.y, nk, g.y), fmaf(a.z, nk, g.z),
                                                        fmaf(a.w, nk, g.w));
      continue;
    }
    const float4 t = reinterpret_cast<const float4*>(theta)[i];
    float4 a = mode == 1 ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(acc)[i];
    a.x = fmaf(nk, t.x - g.x, a.x); a.y = fmaf(nk, t.y - g.y, a.y);
    a.z = fmaf(nk, t.z - g.z, a.z); a.w = fmaf(nk, t.w - g.w, a.w);
    reinterpret_cast<float4*>(acc)[i] = a;
    if (reset) {
      reinterpret_cast<float4*>(theta)[i] = g;
      if (wb != nullptr) reinterpret_cast<uint2*>(wb)[i] = make_uint2(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w));
      if (mom != nullptr && (i << 2) < n_mom) reinterpret_cast<float4*>(mom)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------ multi-source weighted sum
struct WsumArgs {
  const void* src[B200_MAX_RANKS];
  float w[B200_MAX_RANKS];
  int n_src;
};
template <bool BF16>
__global__ void __launch_bounds__(EW_THREADS) weighted_sum_kernel(void* __restrict__ dst, WsumArgs a, long long n) {
  griddep_launch_dependents();
  griddep_wait();
  constexpr int VEC = BF16 ? 8 : 4;
  const long long nv = n / VEC;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    for (int k = 0; k < a.n_src; ++k) {
      const uint4 u = reinterpret_cast<const uint4*>(a.src[k])[i];
      const float wk = a.w[k];
      if constexpr (BF16) {
        const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
        acc[0] = fmaf(wk, p0.x, acc[0]); acc[1] = fmaf(wk, p0.y, acc[1]);
        acc[2] = fmaf(wk, p1.x, acc[2]); acc[3] = fmaf(wk, p1.y, acc[3]);
        acc[4] = fmaf(wk, p2.x, acc[4]); acc[5] = fmaf(wk, p2.y, acc[5]);
        acc[6] = fmaf(wk, p3.x, acc[6]); acc[7] = fmaf(wk, p3.y, acc[7]);
      } else {
        acc[0] = fmaf(wk, __uint_as_float(u.x), acc[0]); acc[1] = fmaf(wk, __uint_as_float(u.y), acc[1]);
        acc[2] = fmaf(wk, __uint_as_float(u.z), acc[2]); acc[3] = fmaf(wk, __uint_as_float(u.w), acc[3]);
      }
    }
    uint4 o;
    if constexpr (BF16) {
      o = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                     pack_bf16x2(acc[6], acc[7]));
    } else {
      o = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
    }
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0) {
    for (long long i = nv * VEC + threadIdx.x; i < n; i += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < a.n_src; ++k) {
        const float v = BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.src[k])[i])
                             : reinterpret_cast<const float*>(a.src[k])[i];
        acc = fmaf(a.w[k], v, acc);
      }
      if (BF16)
        reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(acc);
      else
        reinterpret_cast<float*>(dst)[i] = acc;
    }
  }
}

// ------------------------------------------------------------------ casts
__global__ void __launch_bounds__(EW_THREADS) cast_f32_bf16_kernel(const float* __restrict__ s,
                                                                    __nv_bfloat16* __restrict__ d, long long n) {
  griddep_launch_dependents();
  griddep_wait();
  const long long nv = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(s)[i];
    reinterpret_cast<uint2*>(d)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  if (blockIdx.x == 0)
    for (long long i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void __launch_bounds__(EW_THREADS) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ s,
                                                                    float* __restrict__ d, long long n) {
  griddep_launch_dependents();
  griddep_wait();
  const long long nv = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint2 u = reinterpret_cast<const uint2*>(s)[i];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    reinterpret_cast<float4*>(d)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
  if (blockIdx.x == 0)
    for (long long i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) d[i] = __bfloat162float(s[i]);
}

// ------------------------------------------------------------------ batch gather (X[idx]) on a resident shard
__global__ void __launch_bounds__(EW_THREADS)
gather_rows_kernel(const uint4* __restrict__ src, const long long* __restrict__ idx, uint4* __restrict__ dst,
                   long long n_rows, int row_vecs) {
  griddep_launch_dependents();
  griddep_wait();
  const long long total = n_rows * row_vecs;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / row_vecs;
    const int c = static_cast<int>(i - r * row_vecs);
    dst[i] = __ldg(src + idx[r] * row_vecs + c);
  }
}
__global__ void gather_i64_kernel(const long long* __restrict__ src, const long long* __restrict__ idx,
                                  long long* __restrict__ dst, long long n) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = src[idx[i]];
}

// ------------------------------------------------------------------ column sum of a bf16 [rows, cols] matrix
// (bias gradient).  Block = 32 x 8: 32 consecutive columns, 8 row lanes; grid.x over column groups,
// grid.y over row chunks; partial sums are combined with fp32 atomics.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, long long rows, int cols) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < cols)
    for (long long r = static_cast<long long>(blockIdx.y) * 8 + ty; r < rows; r += static_cast<long long>(gridDim.y) * 8)
      acc += __bfloat162float(x[r * cols + c]);
  s[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += s[j][tx];
    atomicAdd(out + c, t);
  }
}

// 16-byte version (cols % 8 == 0): block = 32 column groups (256 columns) x 8 row lanes, 4 rows in flight per
// thread; grid.x over column chunks, grid.y over row chunks
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const uint4* __restrict__ x, float* __restrict__ out, long long rows, int cols8, int rows_per_cta) {
  griddep_launch_dependents();
  griddep_wait();
  __shared__ float s[8][32][9];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int cg = blockIdx.x * 32 + tx;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > rows) r1 = rows;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (cg < cols8) {
    for (long long r = r0 + ty; r < r1; r += 32) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + 8 * u < r1) v[u] = __ldcs(x + (r + 8 * u) * cols8 + cg);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + 8 * u < r1) {
          const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            a[2 * j] += f.x;
            a[2 * j + 1] += f.y;
          }
        }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s[ty][tx][j] = a[j];
  __syncthreads();
  // 256 threads -> 256 columns of this chunk
  const int col = threadIdx.x, g = col >> 3, j = col & 7;
  if (blockIdx.x * 32 + g < cols8) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += s[y][g][j];
    atomicAdd(out + (static_cast<long long>(blockIdx.x) * 32 + g) * 8 + j, t);
  }
}

// ------------------------------------------------------------------ small bf16 elementwise ops
__global__ void __launch_bounds__(EW_THREADS)
add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o, long long nv, int relu) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 x = a[i], y = b[i];
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 p = unpack_bf16x2(xs[j]), q = unpack_bf16x2(ys[j]);
      float u = p.x + q.x, v = p.y + q.y;
      if (relu) { u = fmaxf(u, 0.f); v = fmaxf(v, 0.f); }
      r[j] = pack_bf16x2(u, v);
    }
    o[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}
// dx = dy * (y > 0)
__global__ void __launch_bounds__(EW_THREADS)
relu_bwd_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy, uint4* __restrict__ dx, long long nv) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 a = y[i], b = dy[i];
    const uint32_t as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 p = unpack_bf16x2(as[j]), q = unpack_bf16x2(bs[j]);
      r[j] = pack_bf16x2(p.x > 0.f ? q.x : 0.f, p.y > 0.f ? q.y : 0.f);
    }
    dx[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}
__device__ __forceinline__ float gelu_f(float v) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * v * (1.f + tanhf(k0 * (v + k1 * v * v * v)));
}
__device__ __forceinline__ float gelu_grad_f(float v) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (v + k1 * v * v * v);
  const float t = tanhf(u);
  return 0.5f * (1.f + t) + 0.5f * v * (1.f - t * t) * k0 * (1.f + 3.f * k1 * v * v);
}
__global__ void __launch_bounds__(EW_THREADS)
gelu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16_rn(gelu_f(__bfloat162float(x[i])));
}
__global__ void __launch_bounds__(EW_THREADS)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                __nv_bfloat16* __restrict__ dx, long long n) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dx[i] = __float2bfloat16_rn(__bfloat162float(dy[i]) * gelu_grad_f(__bfloat162float(x[i])));
}
// 16-byte versions (n % 8 == 0, aligned): 8 elements per thread, streaming loads
__global__ void __launch_bounds__(EW_THREADS)
gelu_vec_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long nv) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 u = __ldcs(x + i);
    const uint32_t in[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = unpack_bf16x2(in[j]);
      o[j] = pack_bf16x2(gelu_f(v.x), gelu_f(v.y));
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void __launch_bounds__(EW_THREADS)
gelu_bwd_vec_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx, long long nv) {
  griddep_launch_dependents();
  griddep_wait();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nv;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 u = __ldcs(x + i), g = __ldcs(dy + i);
    const uint32_t xs[4] = {u.x, u.y, u.z, u.w}, gs[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 v = unpack_bf16x2(xs[j]), d = unpack_bf16x2(gs[j]);
      o[j] = pack_bf16x2(d.x * gelu_grad_f(v.x), d.y * gelu_grad_f(v.y));
    }
    dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
// dst[r, 0:kp] = src[r, 0:k] zero padded (weights whose K is not a multiple of 8, e.g. 7x7x3 = 147)
__global__ void __launch_bounds__(EW_THREADS)
pad_rows_kernel(const __nv_bfloat16* __restrict__ s, __nv_bfloat16* __restrict__ d, long long rows, int k, int kp,
                const uint32_t* __restrict__ flags, const uint32_t* __restrict__ epoch_word, long long elem_off, int granule) {
  // gated mode: do NOT let the dependent GEMM become resident early -- its large CTAs would sit on the SMs while this
  // kernel spins, and the collective that publishes the flags might be the one still waiting for those SMs
  if (flags == nullptr) griddep_launch_dependents();
  griddep_wait();
  if (flags != nullptr) {
    // bcast_gemm staging: the source is a slice of the bf16 arena that the FedAvg collective may still be writing on
    // another stream -- acquire the arrival flags of the granules under it first (bounded spin)
    if (threadIdx.x == 0) {
      const uint32_t need = *reinterpret_cast<const volatile uint32_t*>(epoch_word);
      for (long long t = elem_off / granule; t <= (elem_off + rows * k - 1) / granule; ++t) {
        unsigned long long spins = 0;
        while (static_cast<int32_t>(ld_acquire_sys(flags + t) - need) < 0)
          if (++spins > (1ull << 26)) break;
      }
    }
    __syncthreads();
  }
  const long long total = rows * kp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / kp;
    const int c = static_cast<int>(i - r * kp);
    d[i] = c < k ? s[r * k + c] : __float2bfloat16_rn(0.f);
  }
}

// embedding backward: grad_table[idx[r], :] += dy[r, :]  (fp32 atomics; rows of 8-element vectors)
__global__ void __launch_bounds__(EW_THREADS)
embedding_bwd_kernel(const uint4* __restrict__ dy, const long long* __restrict__ idx, float* __restrict__ grad,
                     long long n_rows, int row_vecs) {
  griddep_launch_dependents();
  griddep_wait();
  const long long total = n_rows * row_vecs;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / row_vecs;
    const int c = static_cast<int>(i - r * row_vecs);
    const uint4 v = dy[i];
    float* g = grad + (idx[r] * row_vecs + c) * 8;
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), cc = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + 4), "f"(cc.x), "f"(cc.y), "f"(d.x), "f"(d.y) : "memory");
  }
}

}  // namespace b200

using namespace b200;
#define RET_LAST() return static_cast<int>(cudaGetLastError())

// max_ctas > 0 caps the grid (grid-stride kernel): an optimizer slice that runs BESIDE latency-bound compute kernels on
// another stream must leave SM slots free for them
extern "C" int b200_fused_sgd(float* w, float* g, float* mom, void* w_bf16, long long n, const float* hyper,
                              int zero_grad, int nesterov, int max_ctas, const unsigned long long* wire_slot,
                              const float* pack_global, const float* pack_scale, long long n_pack, int wire_fp32,
                              cudaStream_t stream) {
  if (n <= 0) return 0;
  SgdPack pk;
  pk.wire_slot = wire_slot; pk.global_w = pack_global; pk.scale = pack_scale;
  pk.n_pack = n_pack > n ? n_pack : n; pk.wire_fp32 = wire_fp32;
  if (wire_slot != nullptr && ((n & 7) || (pk.n_pack & 7))) return -2;
  launch_pdl(fused_sgd_kernel, max_ctas > 0 ? ew_grid(n >> 2, max_ctas) : ew_grid(n >> 2), EW_THREADS, 0, stream, w, g, mom, reinterpret_cast<__nv_bfloat16*>(w_bf16), n,
                                                                hyper, zero_grad, nesterov, pk);
  RET_LAST();
}
extern "C" int b200_fold_client(float* acc, float* theta, const float* global_w, void* w_bf16, float* mom, long long n_mom,
                                long long n, float nk, int mode, int reset, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n & 3) return -2;
  launch_pdl(fold_client_kernel, ew_grid(n >> 2), EW_THREADS, 0, stream, acc, theta, global_w,
             reinterpret_cast<__nv_bfloat16*>(w_bf16), mom, n_mom, n, nk, mode, reset);
  RET_LAST();
}
extern "C" int b200_weighted_sum(void* dst, const void* const* srcs, const float* weights, int n_src, long long n,
                                 int dtype, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n_src > B200_MAX_RANKS || n_src < 1) return -2;
  WsumArgs a;
  a.n_src = n_src;
  for (int k = 0; k < n_src; ++k) { a.src[k] = srcs[k]; a.w[k] = weights[k]; }
  if (dtype == 1)
    launch_pdl(weighted_sum_kernel<true>, ew_grid(n / 8), EW_THREADS, 0, stream, dst, a, n);
  else
    launch_pdl(weighted_sum_kernel<false>, ew_grid(n / 4), EW_THREADS, 0, stream, dst, a, n);
  RET_LAST();
}
extern "C" int b200_cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  launch_pdl(cast_f32_bf16_kernel, ew_grid(n >> 2), EW_THREADS, 0, stream, src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  RET_LAST();
}
extern "C" int b200_cast_bf16_f32(const void* src, float* dst, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  launch_pdl(cast_bf16_f32_kernel, ew_grid(n >> 2), EW_THREADS, 0, stream, reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  RET_LAST();
}
extern "C" int b200_gather_rows(const void* src, const long long* idx, void* dst, long long n_rows,
                                long long row_bytes, cudaStream_t stream) {
  if (n_rows <= 0) return 0;
  if (row_bytes % 16) return -2;
  const int rv = static_cast<int>(row_bytes / 16);
  launch_pdl(gather_rows_kernel, ew_grid(n_rows * rv), EW_THREADS, 0, stream, reinterpret_cast<const uint4*>(src), idx,
                                                                      reinterpret_cast<uint4*>(dst), n_rows, rv);
  RET_LAST();
}
extern "C" int b200_gather_rows_i64(const long long* src, const long long* idx, long long* dst, long long n,
                                    cudaStream_t stream) {
  if (n <= 0) return 0;
  launch_pdl(gather_i64_kernel, ew_grid(n, 64), EW_THREADS, 0, stream, src, idx, dst, n);
  RET_LAST();
}
extern "C" int b200_colsum(const void* x, float* out, long long rows, int cols, int accumulate, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * cols, stream);
  if (cols % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int cols8 = cols / 8;
    const unsigned gx = static_cast<unsigned>((cols8 + 31) / 32);
    // ~2 waves of CTAs over the machine, at least one 32-row pass each
    long long want = (2 * 148 + gx - 1) / gx;
    long long rpc = (rows + want - 1) / want;
    rpc = (rpc + 31) / 32 * 32;
    if (rpc < 32) rpc = 32;
    dim3 gridv(gx, static_cast<unsigned>((rows + rpc - 1) / rpc));
    launch_pdl(colsum_vec_kernel, gridv, 256, 0, stream, reinterpret_cast<const uint4*>(x), out, rows, cols8,
               static_cast<int>(rpc));
    RET_LAST();
  }
  long long gy = (rows + 63) / 64;
  if (gy > 64) gy = 64;
  dim3 grid((cols + 31) / 32, static_cast<unsigned>(gy));
  launch_pdl(colsum_kernel, grid, 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x), out, rows, cols);
  RET_LAST();
}
extern "C" int b200_add_bf16(const void* a, const void* b, void* out, long long n, int relu, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n % 8) return -2;
  launch_pdl(add_bf16_kernel, ew_grid(n / 8), EW_THREADS, 0, stream, reinterpret_cast<const uint4*>(a),
                                                             reinterpret_cast<const uint4*>(b),
                                                             reinterpret_cast<uint4*>(out), n / 8, relu);
  RET_LAST();
}
extern "C" int b200_relu_bwd_bf16(const void* y, const void* dy, void* dx, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n % 8) return -2;
  launch_pdl(relu_bwd_kernel, ew_grid(n / 8), EW_THREADS, 0, stream, reinterpret_cast<const uint4*>(y),
                                                             reinterpret_cast<const uint4*>(dy),
                                                             reinterpret_cast<uint4*>(dx), n / 8);
  RET_LAST();
}
extern "C" int b200_gelu_bf16(const void* x, void* y, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    launch_pdl(gelu_vec_kernel, ew_grid(n / 8), EW_THREADS, 0, stream, reinterpret_cast<const uint4*>(x),
               reinterpret_cast<uint4*>(y), n / 8);
    RET_LAST();
  }
  launch_pdl(gelu_kernel, ew_grid(n), EW_THREADS, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x),
                                                     reinterpret_cast<__nv_bfloat16*>(y), n);
  RET_LAST();
}
extern "C" int b200_gelu_bwd_bf16(const void* x, const void* dy, void* dx, long long n, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
    launch_pdl(gelu_bwd_vec_kernel, ew_grid(n / 8), EW_THREADS, 0, stream, reinterpret_cast<const uint4*>(x),
               reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), n / 8);
    RET_LAST();
  }
  launch_pdl(gelu_bwd_kernel, ew_grid(n), EW_THREADS, 0, stream, reinterpret_cast<const __nv_bfloat16*>(x),
                                                         reinterpret_cast<const __nv_bfloat16*>(dy),
                                                         reinterpret_cast<__nv_bfloat16*>(dx), n);
  RET_LAST();
}
extern "C" int b200_pad_rows_bf16(const void* src, void* dst, long long rows, int k, int kp, const uint32_t* flags,
                                  const uint32_t* epoch_word, long long elem_off, int granule, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (flags != nullptr && (epoch_word == nullptr || granule <= 0)) return -2;
  launch_pdl(pad_rows_kernel, ew_grid(rows * kp), EW_THREADS, 0, stream, reinterpret_cast<const __nv_bfloat16*>(src),
                                                                 reinterpret_cast<__nv_bfloat16*>(dst), rows, k, kp, flags,
                                                                 epoch_word, elem_off, granule);
  RET_LAST();
}

extern "C" int b200_embedding_bwd(const void* dy, const long long* idx, float* grad, long long n_rows, int width,
                                  cudaStream_t stream) {
  if (n_rows <= 0) return 0;
  if (width % 8) return -2;
  launch_pdl(embedding_bwd_kernel, ew_grid(n_rows * (width / 8)), EW_THREADS, 0, stream,
             reinterpret_cast<const uint4*>(dy), idx, grad, n_rows, width / 8);
  RET_LAST();
}

B200_TRACE_REGISTER(elementwise)
