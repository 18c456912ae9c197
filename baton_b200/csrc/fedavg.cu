// Fused FedAvg collective over NVLink 5 / NVSwitch -- ONE persistent kernel per round that does
//
//   phase 0  pack      wire_r[t]  = cast( w_r * (theta_r[t] - global[t]) )   (delta mode)
//                                   cast( w_r * theta_r[t] )                 (weights mode)
//   barrier  (per-CTA flags in peer-mapped signal pads, st.release.sys / ld.acquire.sys)
//   phase 1  reduce    owner(t) pulls tile t from every participant with 16 B peer loads over
//            + bcast   NVLink (or ONE multimem.ld_reduce: the switch adds the replicas), sums in
//                      fp32 in fixed rank order, casts, and pushes the result into tile t of every
//                      live replica's wire buffer (peer stores, or ONE multimem.st replicated by
//                      the switch).  In place: owner(t) is the only reader and writer of tile t.
//   barrier
//   phase 2  apply     global += result ; theta = global ; bf16 shadow = bf16(theta) ; momentum = 0
//                      (the reference's load_state_dict, worker.py:98, with no extra pass), then
//                      publish a per-tile arrival flag so the next forward's first GEMM
//                      (gemm_tcgen05, flag-gated TMA producer) can start on tile 0 while the rest
//                      of the arena is still in flight.
//
// This replaces the reference's upload (worker.py:108-118), CPU reduce (manager.py:119-126),
// broadcast (manager.py:77-86) and load_state_dict (worker.py:98).  No NCCL call on this path.
//
// Tile t (tile_elems elements) is owned by the (t mod A)-th live rank and handled by CTA
// ((t div A) mod G) on EVERY rank in every phase, so a per-CTA cross-GPU barrier is enough:
// CTA b only ever consumes data produced by CTA b of some rank.
//
// Participation: weights[k] == 0 -> rank k is not read (P2P) / packs zeros (NVLS);
// alive_mask bit k == 0 -> rank k is neither read, written nor waited for (dead process), so a
// dead peer cannot hang the collective the way a blocking NCCL call would.
#include "ptx.cuh"
#include "launch.h"

namespace b200 {

constexpr int FEDAVG_THREADS = 512;

// Per-CTA barrier across the live ranks.  pads[k] is rank k's signal pad (peer-mapped); slot
// layout: pad[(cta * B200_MAX_RANKS + src_rank)].  Epochs only grow, so no reset races.
__device__ __forceinline__ bool cta_barrier_all_ranks(const FedAvgArgs& a, uint32_t epoch, int* status) {
  __syncthreads();
  const int t = threadIdx.x;
  bool ok = true;
  if (t < a.world && ((a.alive_mask >> t) & 1u)) {
    fence_sys();
    st_release_sys(a.pads[t] + (static_cast<size_t>(blockIdx.x) * B200_MAX_RANKS + a.rank), epoch);
    const uint32_t* mine = a.pads[a.rank] + (static_cast<size_t>(blockIdx.x) * B200_MAX_RANKS + t);
    unsigned long long spins = 0;
    const unsigned long long limit = a.timeout_cycles_log2 > 0 ? (1ull << a.timeout_cycles_log2) : ~0ull;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > limit) {
        ok = false;
        if (status != nullptr) atomicExch(status, 1 + t);
        break;
      }
    }
  }
  __syncthreads();
  return ok;
}

template <bool WIRE_BF16>
struct Wire;
template <>
struct Wire<true> {  // 8 bf16 per 16 B
  static constexpr int VEC = 8;
  __device__ static void unpack(const uint4& u, float (&f)[8]) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  __device__ static uint4 pack(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
  }
  __device__ static uint4 mc_reduce(const void* p) { return multimem_ld_reduce_bf16x8(p); }
};
template <>
struct Wire<false> {  // 4 fp32 per 16 B
  static constexpr int VEC = 4;
  __device__ static void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y);
    f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ static uint4 pack(const float (&f)[4]) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  __device__ static uint4 mc_reduce(const void* p) {
    float4 r = multimem_ld_reduce_f32x4(p);
    return make_uint4(__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z), __float_as_uint(r.w));
  }
};

template <bool WIRE_BF16>
__global__ void __launch_bounds__(FEDAVG_THREADS, 1) fedavg_allreduce_kernel(const __grid_constant__ FedAvgArgs a) {
  using W = Wire<WIRE_BF16>;
  constexpr int VEC = W::VEC;
  const int G = gridDim.x;
  // live ranks, in rank order; A = number alive; my position among them
  __shared__ uint8_t* s_wire[B200_MAX_RANKS];
  __shared__ long long* s_int[B200_MAX_RANKS];
  __shared__ float s_w[B200_MAX_RANKS];
  int A = 0, my_pos = -1;
  for (int k = 0; k < a.world; ++k)
    if ((a.alive_mask >> k) & 1u) {
      if (k == a.rank) my_pos = A;
      if (threadIdx.x == 0) {
        s_wire[A] = reinterpret_cast<uint8_t*>(a.wire[k]);
        s_int[A] = a.int_wire[k];
        s_w[A] = a.weights[k];
      }
      ++A;
    }
  if (my_pos < 0 || A == 0) return;
  __syncthreads();
  const long long n = a.n;
  const int T = a.tile_elems;  // multiple of VEC * FEDAVG_THREADS is not required; multiple of VEC is
  const long long n_tiles = (n + T - 1) / T;
  const float my_w = a.weights[a.rank];
  const size_t esz = WIRE_BF16 ? 2 : 4;
  uint8_t* my_wire = reinterpret_cast<uint8_t*>(a.wire[a.rank]);

  // ---------------------------------------------------------------- phase 0: pack + prescale + cast
  // P2P mode applies the weight on the reader side (full-precision upload); NVLS mode needs the
  // scaled value on the wire because the switch can only add.
  const float pack_scale = a.use_nvls ? my_w : 1.0f;
  if (my_w != 0.f || a.use_nvls) {
    // same tile -> CTA map as the other phases: CTA b owns q = b, b+G, ... and tiles q*A .. q*A+A-1
    for (long long t = static_cast<long long>(blockIdx.x) * A; t < n_tiles;
         t = ((t + 1) % A == 0) ? (t + 1 + static_cast<long long>(G - 1) * A) : (t + 1)) {
      const long long base = t * T;
      const int len = static_cast<int>((n - base) < T ? (n - base) : T);
      for (int i = threadIdx.x * VEC; i < len; i += FEDAVG_THREADS * VEC) {
        float f[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          float4 th = *reinterpret_cast<const float4*>(a.theta + base + i + j);
          if (a.delta) {
            float4 g = *reinterpret_cast<const float4*>(a.global_w + base + i + j);
            th.x -= g.x; th.y -= g.y; th.z -= g.z; th.w -= g.w;
          }
          f[j] = th.x * pack_scale; f[j + 1] = th.y * pack_scale;
          f[j + 2] = th.z * pack_scale; f[j + 3] = th.w * pack_scale;
        }
        *reinterpret_cast<uint4*>(my_wire + (base + i) * esz) = W::pack(f);
      }
    }
  }
  // integer side arena (num_batches_tracked ...): publish the local values
  if (a.n_int > 0 && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < a.n_int; i += FEDAVG_THREADS) a.int_wire[a.rank][i] = a.int_local[i];
  }
  if (!cta_barrier_all_ranks(a, a.epoch + 1, a.status)) return;

  // ---------------------------------------------------------------- phase 1: reduce + broadcast
  // tile t belongs to live rank alive[t % A]; on that rank CTA ((t / A) % G) handles it
  for (long long t = my_pos + static_cast<long long>(blockIdx.x) * A; t < n_tiles; t += static_cast<long long>(G) * A) {
    const long long base = t * T;
    const int len = static_cast<int>((n - base) < T ? (n - base) : T);
    for (int i = threadIdx.x * VEC; i < len; i += FEDAVG_THREADS * VEC) {
      const size_t off = (base + i) * esz;
      uint4 out;
      if (a.use_nvls) {
        out = W::mc_reduce(reinterpret_cast<const uint8_t*>(a.wire_mc) + off);  // switch adds the replicas
        multimem_st_v4(reinterpret_cast<uint8_t*>(a.wire_mc) + off, out);       // switch replicates the store
      } else {
        // issue all peer loads first (MLP), then accumulate in fixed rank order (deterministic)
        uint4 v[B200_MAX_RANKS];
#pragma unroll
        for (int k = 0; k < B200_MAX_RANKS; ++k) {
          if (k < A && s_w[k] != 0.f) v[k] = ld_volatile_v4(s_wire[k] + off);
        }
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll
        for (int k = 0; k < B200_MAX_RANKS; ++k) {
          if (k < A) {
            const float w = s_w[k];
            if (w != 0.f) {
              float f[VEC];
              W::unpack(v[k], f);
#pragma unroll
              for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w, f[j], acc[j]);
            }
          }
        }
        out = W::pack(acc);
#pragma unroll
        for (int k = 0; k < B200_MAX_RANKS; ++k) {
          if (k < A) st_na_v4(s_wire[k] + off, out);
        }
      }
    }
  }
  if (!cta_barrier_all_ranks(a, a.epoch + 2, a.status)) return;

  // ---------------------------------------------------------------- phase 2: running-mean apply
  // CTA b applies exactly the tiles CTA b of the owners produced: (t / A) % G == b
  for (long long q = blockIdx.x; q * A < n_tiles; q += G) {
    for (int r = 0; r < A; ++r) {
      const long long t = q * A + r;
      if (t >= n_tiles) break;
      const long long base = t * T;
      const int len = static_cast<int>((n - base) < T ? (n - base) : T);
      for (int i = threadIdx.x * VEC; i < len; i += FEDAVG_THREADS * VEC) {
        float f[VEC];
        W::unpack(ld_volatile_v4(my_wire + (base + i) * esz), f);
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          float4 nw;
          if (a.delta) {
            float4 g = *reinterpret_cast<const float4*>(a.global_w + base + i + j);
            nw = make_float4(g.x + f[j], g.y + f[j + 1], g.z + f[j + 2], g.w + f[j + 3]);
            *reinterpret_cast<float4*>(a.global_w + base + i + j) = nw;
          } else {
            nw = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            if (a.global_w != nullptr) *reinterpret_cast<float4*>(a.global_w + base + i + j) = nw;
          }
          *reinterpret_cast<float4*>(a.theta + base + i + j) = nw;
          if (a.momentum != nullptr)
            *reinterpret_cast<float4*>(a.momentum + base + i + j) = make_float4(0.f, 0.f, 0.f, 0.f);
          if (a.theta_bf16 != nullptr) {
            uint2 o = make_uint2(pack_bf16x2(nw.x, nw.y), pack_bf16x2(nw.z, nw.w));
            *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(a.theta_bf16) + (base + i + j) * 2) = o;
          }
        }
      }
      if (a.tile_flags != nullptr) {
        __syncthreads();  // every thread's stores of this tile are done
        if (threadIdx.x == 0) {
          __threadfence();
          st_release_sys(a.tile_flags + t, a.flag_value);
        }
      }
    }
  }
  // integer side arena: max over participants (BatchNorm step counters only ever grow)
  if (a.n_int > 0 && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < a.n_int; i += FEDAVG_THREADS) {
      long long m = a.int_local[i];
      for (int k = 0; k < A; ++k)
        if (s_w[k] != 0.f) {
          long long v = *reinterpret_cast<volatile long long*>(s_int[k] + i);
          m = v > m ? v : m;
        }
      a.int_local[i] = m;
    }
  }
  // closing barrier: nobody may start the next round's phase 0 (overwriting its wire buffer, which
  // peers pushed results into) or exit and let the host reuse int_wire while a peer still reads it
  cta_barrier_all_ranks(a, a.epoch + 3, a.status);
}

// stand-alone cross-GPU barrier on the signal pads (one CTA): used to fence host-side phases
__global__ void flag_barrier_kernel(FedAvgArgs a, int slot) {
  const int t = threadIdx.x;
  if (t < a.world && ((a.alive_mask >> t) & 1u)) {
    fence_sys();
    st_release_sys(a.pads[t] + (static_cast<size_t>(slot) * B200_MAX_RANKS + a.rank), a.epoch);
    const uint32_t* mine = a.pads[a.rank] + (static_cast<size_t>(slot) * B200_MAX_RANKS + t);
    while (static_cast<int32_t>(ld_acquire_sys(mine) - a.epoch) < 0) {
    }
  }
}

}  // namespace b200

extern "C" int b200_fedavg_allreduce(const FedAvgArgs* args, int n_ctas, cudaStream_t stream) {
  using namespace b200;
  if (args->world > B200_MAX_RANKS || args->n % 8 != 0 || args->tile_elems % 8 != 0) return -2;
  if (n_ctas < 1) n_ctas = 1;
  if (args->wire_bf16)
    fedavg_allreduce_kernel<true><<<n_ctas, FEDAVG_THREADS, 0, stream>>>(*args);
  else
    fedavg_allreduce_kernel<false><<<n_ctas, FEDAVG_THREADS, 0, stream>>>(*args);
  return static_cast<int>(cudaGetLastError());
}

extern "C" int b200_flag_barrier(uint32_t* const* pads, int rank, int world, uint32_t alive_mask, uint32_t epoch,
                                 int slot, cudaStream_t stream) {
  using namespace b200;
  if (world > B200_MAX_RANKS) return -2;
  FedAvgArgs a = {};
  for (int k = 0; k < world; ++k) a.pads[k] = pads[k];
  a.rank = rank; a.world = world; a.alive_mask = alive_mask; a.epoch = epoch;
  flag_barrier_kernel<<<1, 32, 0, stream>>>(a, slot);
  return static_cast<int>(cudaGetLastError());
}
