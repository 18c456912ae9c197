// bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (accumulator in TMEM) -> tcgen05.ld epilogue with fused bias / activation / cast.
//
//     D[M,N] = act( alpha * A[M,K] * B[N,K]^T + bias[N] )          (A, B bf16; accumulate fp32)
//
// Each operand may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]), so the
// three training GEMMs need no transposes:
//     fwd    Y  = X  W^T      A = X   (K-major)   B = W  (K-major)
//     dgrad  dX = dY W        A = dY  (K-major)   B = W  (MN-major)
//     wgrad  dW = dY^T X      A = dY  (MN-major)  B = X  (MN-major)
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (warp q reads TMEM lanes 32q..32q+31).
// The pipeline depth is a runtime value sized to the K extent, so short-K GEMMs ask for little
// shared memory and several CTAs share an SM (one CTA's epilogue overlaps another's main loop).
//
// Split-K, two flavours:
//   * atomic   (fp32 output, gradient accumulation): gridDim.z slices red.global.add into D;
//   * cluster  (any output): the z-slices of one output tile form a thread-block CLUSTER; each CTA
//     parks its fp32 partial tile in its own shared memory, and after a cluster barrier CTA r
//     reduces rows [r*128/S, (r+1)*128/S) of all S partials through distributed shared memory
//     (ld.shared::cluster), applies the epilogue and stores -- no workspace, no second kernel.
//     This is what makes the deep layers of a ResNet (M = 128, K = 4608) use more than 8 SMs.
//
// Flag-gated variant ("bcast_gemm", K3 in SURVEY.md 2.6): the producer acquires per-arena-tile
// arrival flags (published by the FedAvg kernel with st.release) before issuing the TMA loads of
// a weight tile, so the first GEMM of a round consumes the new global weights tile by tile while
// the rest of the model is still landing over NVLink.
#define B200_TU_TAG 1
#include "ptx.cuh"
#include "launch.h"
#include "pdl.cuh"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int MAX_STAGES = 8;

struct GemmParams {
  int M, N, K;
  void* D;
  long long ldd;          // leading dimension of D in elements
  const float* bias;      // [N] or nullptr
  int out_fp32;           // 1: D is fp32, 0: D is bf16
  int act;                // 0 none, 1 relu, 2 gelu(tanh)
  int epi_staged;         // persistent kernels: 1 = smem-staged row-coalesced stores, 0 = direct per-lane stores
  float* col_stats;       // optional [2N]: += column sums / sums of squares of the (bf16-rounded) output (BatchNorm)
  int a_mn, b_mn;         // operand majors
  int k_tiles_per_split;  // split-K: k tiles handled by one z-slice
  int atomic_out;         // 1: red.add fp32 into D
  int cluster_k;          // > 1: z-slices form a cluster of this size and reduce through DSMEM
  int stages;             // pipeline depth (1..MAX_STAGES)
  const uint32_t* tile_flags;  // optional arrival flags, one per arena tile (bcast_gemm)
  uint32_t flag_epoch;         // value a flag must reach before the data under it may be loaded
  long long flag_elem_off;     // arena element offset of B[0,0]
  int flag_tile_elems;         // arena elements covered by one flag
  long long flag_bias_off;     // arena element offset of bias[0], or -1
  long long ldb;               // row pitch of B (elements)
  float alpha;
  // strided-batched mode (attention): blockIdx.z = outer * batch_inner + inner; operands come from
  // 4-D tensor maps (col, row, inner, outer); D is offset by outer * d_outer + inner * d_inner elements
  int batched, batch_inner;
  int batch_count;         // persistent batched mode: number of z slices
  long long d_outer, d_inner;
  // implicit-GEMM convolution (CONV template modes; appended last so existing field offsets do not move)
  int conv_ho, conv_wo;    // output image size
  int conv_stride, conv_pad;
  int conv_kw;             // filter width (tap = r * kw + s)
  int conv_cin;            // channels of the im2col-gathered tensor (multiple of 64): Cin forward, Cout for dgrad
  int conv_taps;           // dgrad: KH * KW
  int conv_ncol;           // dgrad: Cin of the convolution (column pitch of one tap inside a weight row)
  const uint32_t* flag_epoch_ptr;   // bcast_gemm inside a captured graph: the required flag value lives in device memory
};

// Wait until every arrival flag covering arena elements [e0, e1] has reached `need` (published by the FedAvg kernel with
// st.release.sys).  Bounded: a collective that died must not hang the consumer (it then reads what is there).
__device__ __forceinline__ void wait_arrival_flags(const uint32_t* flags, long long e0, long long e1, int granule,
                                                   uint32_t need) {
  for (long long t = e0 / granule; t <= e1 / granule; ++t) {
    unsigned long long spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(flags + t) - need) < 0) {
      if (++spins > (1ull << 26)) break;
    }
  }
}

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PART_PITCH = BN + 4;                  // floats; +4 keeps float4 alignment, skews banks
  static constexpr int PART_BYTES = BM * PART_PITCH * 4;     // fp32 partial tile for the cluster reduce
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float t = tanhf(k0 * (v + k1 * v * v * v));
    return 0.5f * v * (1.f + t);
  }
  return v;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(cta_rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(remote)
               : "memory");
  return v;
}

// bias + activation + cast + store of `NV` consecutive output columns of one row
// Column sums of a 32 x 32 register block held one ROW per lane.  Butterfly: at every step a lane keeps one
// half of its remaining columns and trades the other half with its partner, so after 5 steps (31 shuffles)
// lane j owns the complete sum of column j.
#define COLSUM_STEP(OFF, HALF)                                                  \
  {                                                                             \
    const bool upper = (lane & (OFF)) != 0;                                     \
    _Pragma("unroll") for (int i = 0; i < (HALF); ++i) {                        \
      const float keep = upper ? t[i + (HALF)] : t[i];                          \
      const float send = upper ? t[i] : t[i + (HALF)];                          \
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, (OFF));                  \
    }                                                                           \
  }
__device__ __forceinline__ float warp_colsum32(float (&t)[32]) {
  const uint32_t lane = lane_id();
  COLSUM_STEP(16, 16) COLSUM_STEP(8, 8) COLSUM_STEP(4, 4) COLSUM_STEP(2, 2) COLSUM_STEP(1, 1)
  return t[0];
}
#undef COLSUM_STEP
// BatchNorm batch statistics fused into the producing GEMM: every lane of the warp must call this.  Rows
// beyond M hold exact zeros (TMA zero-fills out-of-range operand rows; no bias / activation in this mode).
__device__ __forceinline__ void accumulate_col_stats(const GemmParams& p, int col0, const float (&v)[32]) {
  float s[32], q[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float r = __bfloat162float(__float2bfloat16_rn(v[j]));   // what the consumer will read back
    s[j] = r;
    q[j] = r * r;
  }
  const float cs = warp_colsum32(s), cq = warp_colsum32(q);
  const int col = col0 + static_cast<int>(lane_id());
  if (col < p.N) {
    atomicAdd(p.col_stats + col, cs);
    atomicAdd(p.col_stats + p.N + col, cq);
  }
}

// Same butterflies, but the warp's column sums go to shared memory (`sbuf[0:BN]` sums, `sbuf[BN:2BN]` sums of squares of
// ONE warp): the four epilogue warps of a CTA are combined there and the CTA issues ONE global atomic per statistic
// instead of four -- with 256 CTAs (ResNet stem) the atomics of a launch pile up on 2 N addresses and serialise in L2.
__device__ __forceinline__ void stage_col_stats(float* sbuf, int BNv, int c, const float (&v)[32]) {
  float s[32], q[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float r = __bfloat162float(__float2bfloat16_rn(v[j]));
    s[j] = r;
    q[j] = r * r;
  }
  const float cs = warp_colsum32(s), cq = warp_colsum32(q);
  const int lane = static_cast<int>(lane_id());
  sbuf[c + lane] = cs;
  sbuf[BNv + c + lane] = cq;
}

// alpha / bias / activation of one row chunk.  The (activation, bias) combination is resolved ONCE per chunk
// with warp-uniform branches into fully unrolled straight-line code; the per-element runtime switch this
// replaces cost ~900 SASS instructions per 32-column chunk (ncu: profiles/r1_ncu_gemm_qkv_epilogue.txt).
template <int NV, int ACT, bool BIAS>
__device__ __forceinline__ void transform_chunk_t(const GemmParams& p, int col0, float (&v)[NV]) {
  float b[NV];
  if (BIAS) {
    if (col0 + NV <= p.N && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) {      // every lane reads the same addresses: one broadcast request each
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) b[j] = (col0 + j) < p.N ? __ldg(p.bias + col0 + j) : 0.f;
    }
  }
  const float alpha = p.alpha;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float x = v[j] * alpha;
    if (BIAS) x += b[j];
    if (ACT == 1) x = fmaxf(x, 0.f);
    if (ACT == 2) {
      const float k0 = 0.7978845608028654f, k1 = 0.044715f;
      x = 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
    }
    v[j] = x;
  }
}
template <int NV>
__device__ __forceinline__ void transform_chunk(const GemmParams& p, int col0, float (&v)[NV]) {
  if (p.bias == nullptr) {
    if (p.act == 0) transform_chunk_t<NV, 0, false>(p, col0, v);
    else if (p.act == 1) transform_chunk_t<NV, 1, false>(p, col0, v);
    else transform_chunk_t<NV, 2, false>(p, col0, v);
  } else {
    if (p.act == 0) transform_chunk_t<NV, 0, true>(p, col0, v);
    else if (p.act == 1) transform_chunk_t<NV, 1, true>(p, col0, v);
    else transform_chunk_t<NV, 2, true>(p, col0, v);
  }
}

template <int NV>
__device__ __forceinline__ void store_row_chunk(const GemmParams& p, int row, int col0, float (&v)[NV], bool vec_ok,
                                                size_t d_off = 0) {
  transform_chunk<NV>(p, col0, v);
  const bool full = (col0 + NV <= p.N);
  if (p.out_fp32) {
    float* d = reinterpret_cast<float*>(p.D) + d_off + static_cast<size_t>(row) * p.ldd + col0;
    if (p.atomic_out) {
      if (full && vec_ok) {
#pragma unroll
        for (int j = 0; j < NV; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + j), "f"(v[j]), "f"(v[j + 1]),
                       "f"(v[j + 2]), "f"(v[j + 3])
                       : "memory");
      } else {
        _Pragma("unroll") for (int j = 0; j < NV; ++j) if (col0 + j < p.N) atomicAdd(d + j, v[j]);
      }
    } else if (full && vec_ok) {
#pragma unroll
      for (int j = 0; j < NV; j += 4) *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      _Pragma("unroll") for (int j = 0; j < NV; ++j) if (col0 + j < p.N) d[j] = v[j];
    }
  } else {
    __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + d_off + static_cast<size_t>(row) * p.ldd + col0;
    if (full && vec_ok) {
#pragma unroll
      for (int j = 0; j < NV; j += 8) {
        uint4 o;
        o.x = pack_bf16x2(v[j], v[j + 1]);
        o.y = pack_bf16x2(v[j + 2], v[j + 3]);
        o.z = pack_bf16x2(v[j + 4], v[j + 5]);
        o.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(d + j) = o;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < NV; ++j) if (col0 + j < p.N) d[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// ---- fixed-depth pipeline, one CTA per SM: the default path ----
// CONV: 0 = plain GEMM; 1 = implicit-GEMM conv forward (A = im2col(x) gathered by TMA im2col, k-tile = one filter
// tap x 64 input channels); 2 = implicit wgrad (B = im2col(x) MN-major, k-tile = 64 output pixels, every 64-wide
// N atom = one tap x 64 channels); 3 = implicit dgrad of a stride-1 convolution: dx = conv(dy, flipped w) -- A =
// im2col(dy) gathered by TMA im2col with pad' = k - 1 - pad (k-tile = one flipped tap x 64 OUTPUT channels), B = the
// [64 cout] x [BN cin] slab of that tap inside the channels_last weight matrix, loaded MN-major (no weight transpose,
// no col2im).  See csrc/im2col_tma.cu for the tensor maps.
template <int BN, int STAGES, int CONV = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 2)      // <= 128 registers: two shallow-ring CTAs may share an SM
gemm_bf16_fixed_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16B aligned: realign to the 1024B the 128B swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  griddep_launch_dependents();  // PDL: the next kernel may start its prologue now
  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int k_tiles_total = (p.K + BK - 1) / BK;
  const int bz_outer = p.batched ? static_cast<int>(blockIdx.z) / p.batch_inner : 0;
  const int bz_inner = p.batched ? static_cast<int>(blockIdx.z) % p.batch_inner : 0;
  const int kt_begin = p.batched ? 0 : blockIdx.z * p.k_tiles_per_split;
  int kt_end = kt_begin + p.k_tiles_per_split;
  if (kt_end > k_tiles_total) kt_end = k_tiles_total;
  const int num_kt = kt_end - kt_begin;  // host guarantees >= 1 for every launched z

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, BN);  // BN fp32 accumulator columns (power of two >= 32)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // PDL: everything above overlapped the previous kernel; its results are visible from here

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (p.tile_flags != nullptr) {
        // bcast_gemm: wait until the FedAvg kernel has published every arena tile under the rows
        // [n0, n0+BN) of the (K-major) weight matrix this CTA is about to TMA-load
        const int rows_here = (p.N - n0) < BN ? (p.N - n0) : BN;
        const uint32_t need = p.flag_epoch_ptr != nullptr ? *reinterpret_cast<const volatile uint32_t*>(p.flag_epoch_ptr)
                                                          : p.flag_epoch;
        wait_arrival_flags(p.tile_flags, p.flag_elem_off + static_cast<long long>(n0) * p.ldb,
                           p.flag_elem_off + static_cast<long long>(n0 + rows_here) * p.ldb - 1, p.flag_tile_elems, need);
        if (p.flag_bias_off >= 0)  // the bias slice the epilogue of this CTA will add
          wait_arrival_flags(p.tile_flags, p.flag_bias_off + n0, p.flag_bias_off + n0 + rows_here - 1,
                             p.flag_tile_elems, need);
        fence_proxy_async_all();  // order the acquires before the async-proxy (TMA) reads of global memory
      }
      TRACE_POINT();  // fixed: producer starts issuing TMA
      for (int i = 0; i < num_kt; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        const int k0 = (kt_begin + i) * BK;
        if constexpr (CONV == 2) {
          const int live = (p.N - n0 + 63) / 64 < BN / 64 ? (p.N - n0 + 63) / 64 : BN / 64;
          mbar_expect_tx(&full_bar[s], L::A_BYTES + live * 8192);
        } else {
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        }
        if (p.batched) {
          if (!p.a_mn) {
            tma_load_4d(sa, &tmA, &full_bar[s], k0, m0, bz_inner, bz_outer);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_4d(sa + j * 8192, &tmA, &full_bar[s], m0 + j * 64, k0, bz_inner, bz_outer);
          }
          if (!p.b_mn) {
            tma_load_4d(sb, &tmB, &full_bar[s], k0, n0, bz_inner, bz_outer);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, k0, bz_inner, bz_outer);
          }
        } else {
        if constexpr (CONV == 1 || CONV == 3) {
          // k-tile -> (filter tap, 64-channel block); base pixel of this CTA's 128 output pixels in input coords
          const int kt = kt_begin + i;
          const int cblocks = p.conv_cin >> 6;
          const int tap = kt / cblocks, cb = kt - tap * cblocks;
          const int fr = tap / p.conv_kw, fs = tap - fr * p.conv_kw;
          const int q0 = m0 % p.conv_wo, t0 = m0 / p.conv_wo;
          tma_load_im2col_4d(sa, &tmA, &full_bar[s], cb * 64, q0 * p.conv_stride - p.conv_pad,
                             (t0 % p.conv_ho) * p.conv_stride - p.conv_pad, t0 / p.conv_ho, fs, fr);
        } else if (!p.a_mn) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);  // box [64 k][128 rows]
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)  // box [64 m][64 k rows] per MN atom
            tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], m0 + j * 64, k0);
        }
        if constexpr (CONV == 2) {
          // k-tile = 64 output pixels starting at k0; N atom j = (tap, channel block) of column n0 + 64 j
          const int q0 = k0 % p.conv_wo, t0 = k0 / p.conv_wo;
          const int w0 = q0 * p.conv_stride - p.conv_pad, h0 = (t0 % p.conv_ho) * p.conv_stride - p.conv_pad;
          const int img = t0 / p.conv_ho;
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) {
            const int col = n0 + j * 64;
            if (col < p.N) {          // atoms past the last tap are never stored: leave them (expect_tx below)
              const int tap = col / p.conv_cin, c = col - tap * p.conv_cin;
              const int fr = tap / p.conv_kw, fs = tap - fr * p.conv_kw;
              tma_load_im2col_4d(sb + j * 8192, &tmB, &full_bar[s], c, w0, h0, img, fs, fr);
            }
          }
        } else if constexpr (CONV == 3) {
          const int kt = kt_begin + i;
          const int cblocks = p.conv_cin >> 6;
          const int tap = kt / cblocks, cb = kt - tap * cblocks;
          const int wcol = (p.conv_taps - 1 - tap) * p.conv_ncol + n0;      // flipped tap, cin block of this CTA
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], wcol + j * 64, cb * 64);
        } else if (!p.b_mn) {
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);  // box [64 k][BN rows]
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, k0);
        }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    for (int i = 0; i < num_kt; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        if (i == 0) TRACE_POINT();  // fixed: first k-tile landed in smem
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: 8-row groups are 1024 B apart (SBO), K advance = 32 B inside the swizzle row.
          // MN-major: 64-wide MN atoms are 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO),
          //           K advance of 16 rows = 2048 B.
          const uint64_t ad = p.a_mn ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t bd = p.b_mn ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
          tc_mma_f16(tmem_base, ad, bd, idesc, (i | k) != 0);
        }
        tc_commit(&empty_bar[s]);                       // frees the smem slot when the MMAs retire
        if (i == num_kt - 1) tc_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
    }
  }
  {
    // ===================== epilogue: ALL EIGHT WARPS =====================
    // A warp can only read the TMEM lane quarter (warp % 4), but two warps may share a quarter: once the producer / MMA /
    // allocator warps have finished their roles they join, so the BN accumulator columns are drained by eight warps
    // instead of four (warps 4-7 take the first half of the 32-column chunks, warps 0-3 the second).  The epilogue was
    // 1.2 us (2.0 us with fused BatchNorm statistics) of a 5-7 us latency-bound conv GEMM (intra-kernel timeline,
    // profiles/r2_gemm_anatomy_*.txt).
    const int q = warp & 3;
    constexpr int NCHUNK = BN / 32;
    const int c_begin = (warp >= 4 ? 0 : (NCHUNK + 1) / 2) * 32, c_end = (warp >= 4 ? (NCHUNK + 1) / 2 : NCHUNK) * 32;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 128) TRACE_POINT();  // fixed: accumulator complete (epilogue starts)
    const int row = m0 + q * 32 + static_cast<int>(lane_id());
    const bool row_ok = row < p.M;
    const size_t elt = p.out_fp32 ? 4 : 2;
    uint8_t* drow = reinterpret_cast<uint8_t*>(p.D) +
                    (static_cast<size_t>(row) * p.ldd + static_cast<size_t>(bz_outer) * p.d_outer +
                     static_cast<size_t>(bz_inner) * p.d_inner) * elt;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((p.ldd * elt) % 16 == 0);
    // fused BatchNorm statistics: the operand ring is drained by now (the accumulator is complete), so its first bytes
    // hold the per-warp column sums: [4 warps][2 * BN] floats
    float* sstat = reinterpret_cast<float*>(smem) + q * 2 * BN;
    const bool want_stats = p.col_stats != nullptr;
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      const int col0 = n0 + c;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (col0 >= p.N) {                          // warp-uniform
        if (want_stats) { sstat[c + lane_id()] = 0.f; sstat[BN + c + lane_id()] = 0.f; }
        continue;
      }
      transform_chunk<32>(p, col0, v);
      if (want_stats) stage_col_stats(sstat, BN, c, v);
      if (!row_ok) continue;
      const bool full = (col0 + 32 <= p.N);
      if (p.atomic_out) {
        float* d = reinterpret_cast<float*>(drow) + col0;
        if (full && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + j), "f"(v[j]), "f"(v[j + 1]),
                         "f"(v[j + 2]), "f"(v[j + 3])
                         : "memory");
        } else {
          _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < p.N) atomicAdd(d + j, v[j]);
        }
      } else if (p.out_fp32) {
        float* d = reinterpret_cast<float*>(drow) + col0;
        if (full && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < p.N) d[j] = v[j];
        }
      } else {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(drow) + col0;
        if (full && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(v[j], v[j + 1]);
            o.y = pack_bf16x2(v[j + 2], v[j + 3]);
            o.z = pack_bf16x2(v[j + 4], v[j + 5]);
            o.w = pack_bf16x2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(d + j) = o;
          }
        } else {
          _Pragma("unroll") for (int j = 0; j < 32; ++j) if (col0 + j < p.N) d[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
    if (threadIdx.x == 128) TRACE_POINT();  // fixed: epilogue stores issued
    if (want_stats) {
      __syncthreads();                                     // all eight warps staged their column sums
      const float* all = reinterpret_cast<const float*>(smem);
      for (int i = threadIdx.x; i < 2 * BN; i += GEMM_THREADS) {
        const int col = i < BN ? i : i - BN;
        if (n0 + col < p.N)
          atomicAdd(p.col_stats + (i < BN ? 0 : p.N) + n0 + col,
                    all[i] + all[2 * BN + i] + all[4 * BN + i] + all[6 * BN + i]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRACE_POINT();  // fixed: all warps done (before TMEM dealloc / exit)
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}


// Epilogue of one 128 x BN accumulator stage for the persistent kernels.  TMEM rows land one per lane, so
// direct stores would scatter 16 B pieces over 32 different output rows per instruction; instead each warp
// stages 32 rows x 128 B in shared memory (pitch 144 B: conflict-free 16 B accesses) and writes them back
// with lanes running along the row -- every store instruction covers four complete 128 B row segments.
constexpr int EPI_PITCH = 144;
constexpr int EPI_WARP_BYTES = 32 * EPI_PITCH;

template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int q, int m0, int n0,
                                              uint8_t* stage, bool vec_ok, size_t d_off = 0) {
  const int lane = static_cast<int>(lane_id());
  const int row = m0 + q * 32 + lane;
  const int elt = p.out_fp32 ? 4 : 2;
  const bool fast = p.epi_staged && vec_ok && !p.atomic_out && (p.N % 8 == 0);
  if (!fast) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      const int col0 = n0 + c;
      if (col0 >= p.N) continue;                  // warp-uniform
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.col_stats != nullptr) accumulate_col_stats(p, col0, v);   // alpha == 1, no bias / act in this mode
      if (row >= p.M) continue;
      store_row_chunk<32>(p, row, col0, v, vec_ok, d_off);
    }
    return;
  }
  const int W = 128 / elt;                          // output columns per 128-byte pass
  uint8_t* srow = stage + lane * EPI_PITCH;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += W) {
    if (n0 + c0 >= p.N) break;                      // warp-uniform
    for (int cc = 0; cc < W; cc += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c0 + cc, r);
      tmem_ld_wait();
      const int col0 = n0 + c0 + cc;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      transform_chunk<32>(p, col0, v);
      if (p.col_stats != nullptr) accumulate_col_stats(p, col0, v);
      if (p.out_fp32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(srow + (cc + j) * 4) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(srow + (cc + j) * 2) =
              make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]), pack_bf16x2(v[j + 4], v[j + 5]),
                         pack_bf16x2(v[j + 6], v[j + 7]));
      }
    }
    __syncwarp();
    const int per16 = 16 / elt;                     // columns per 16-byte chunk
#pragma unroll
    for (int it = 0; it < 8; ++it) {                // 32 rows x 8 chunks = 256 chunks, 32 per instruction
      const int idx = it * 32 + lane;
      const int r = idx >> 3, ch = idx & 7;
      const int grow = m0 + q * 32 + r;
      const int col = n0 + c0 + ch * per16;
      if (grow < p.M && col < p.N) {
        const uint4 val = *reinterpret_cast<const uint4*>(stage + r * EPI_PITCH + ch * 16);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.D) +
                                  (d_off + static_cast<size_t>(grow) * p.ldd + col) * elt) = val;
      }
    }
    __syncwarp();
  }
}

// ---- persistent kernel: one CTA per SM walks the tile list; two TMEM accumulator stages let the
// ---- epilogue of tile i overlap the main loop of tile i+1 (large problems: >= one wave of tiles) ----
template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator stage complete (MMA -> epilogue)
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator stage drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(tmem_slot + 4);   // 4 warps x 32 rows x 144 B

  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_per_z = tiles_m * tiles_n;
  const int num_tiles = tiles_per_z * (p.batched ? p.batch_count : 1);   // batched: z-major tile list
  const int num_kt = (p.K + BK - 1) / BK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer: runs ahead across tile boundaries =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // consecutive CTAs share the same N block (the B tile stays hot in L2) and walk M
        const int z = tile / tiles_per_z, rem = tile - z * tiles_per_z;
        const int m0 = (rem % tiles_m) * BM, n0 = (rem / tiles_m) * BN;
        const int zo = p.batched ? z / p.batch_inner : 0, zi = p.batched ? z % p.batch_inner : 0;
        for (int kt = 0; kt < num_kt; ++kt) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          const int k0 = kt * BK;
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          if (p.batched) {
            if (!p.a_mn) {
              tma_load_4d(sa, &tmA, &full_bar[s], k0, m0, zi, zo);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_4d(sa + j * 8192, &tmA, &full_bar[s], m0 + j * 64, k0, zi, zo);
            }
            if (!p.b_mn) {
              tma_load_4d(sb, &tmB, &full_bar[s], k0, n0, zi, zo);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, k0, zi, zo);
            }
          } else {
            if (!p.a_mn) {
              tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], m0 + j * 64, k0);
            }
            if (!p.b_mn) {
              tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, k0);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    int s = 0;
    uint32_t ph = 0;
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      mbar_wait(&tempty_bar[acc], aph ^ 1);   // the epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kt = 0; kt < num_kt; ++kt) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = p.a_mn ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                       : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bd = p.b_mn ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                       : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
            tc_mma_f16(d_tmem, ad, bd, idesc, (kt | k) != 0);
          }
          tc_commit(&empty_bar[s]);
          if (kt == num_kt - 1) tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: drains stage `acc` while the MMA warp fills the other =====================
    const int q = warp & 3;
    const size_t elt = p.out_fp32 ? 4 : 2;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((p.ldd * elt) % 16 == 0);
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      const int z = tile / tiles_per_z, rem = tile - z * tiles_per_z;
      const int m0 = (rem % tiles_m) * BM, n0 = (rem / tiles_m) * BN;
      size_t d_off = 0;
      if (p.batched)
        d_off = static_cast<size_t>(z / p.batch_inner) * p.d_outer + static_cast<size_t>(z % p.batch_inner) * p.d_inner;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      epilogue_tile<BN>(p, tmem_base + acc * BN, q, m0, n0, epi_stage + q * EPI_WARP_BYTES,
                        vec_ok && ((d_off * elt) & 15) == 0, d_off);
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
}

// ---- 2-CTA persistent kernel: a CTA PAIR (cluster of 2, adjacent SMs) computes a 256 x BN tile with
// ---- tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and only HALF of the B tile, the
// ---- tensor cores of both SMs read both halves, so L2 -> SMEM traffic per FLOP drops by a third.
// ---- Leader CTA (cluster rank 0) issues every MMA; both CTAs run TMA producers and epilogues.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at the same smem offset in BOTH CTAs of the pair once the MMAs retire
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 0x3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const GemmParams p) {
  constexpr int A_BYTES = BM * BK * 2;            // this CTA's 128 rows of A
  constexpr int BH_BYTES = (BN / 2) * BK * 2;     // this CTA's half of the B tile
  constexpr int STAGE = A_BYTES + BH_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                   // both CTAs (multicast commit)
  uint64_t* tfull_bar = empty_bar + STAGES;                                  // [2] both CTAs (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;                                      // [2] leader only, 8 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(tmem_slot + 4);            // 4 warps x 32 rows x 144 B

  griddep_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kt = (p.K + BK - 1) / BK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);   // 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own A rows + own half of B =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        const int m0 = (tile % tiles_m) * 2 * BM + static_cast<int>(rank) * BM;
        const int n0 = (tile / tiles_m) * BN + static_cast<int>(rank) * (BN / 2);
        for (int kt = 0; kt < num_kt; ++kt) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = kt * BK;
          const uint32_t lbar = mapa_u32(smem_u32(&full_bar[s]), 0);   // the LEADER's barrier collects both CTAs' bytes
          if (leader) mbar_expect_tx(&full_bar[s], 2 * STAGE);
          if (!p.a_mn) {
            tma_load_2d_2sm(sa, &tmA, lbar, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(sa + j * 8192, &tmA, lbar, m0 + j * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_2sm(sb, &tmB, lbar, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_2sm(sb + j * 8192, &tmB, lbar, n0 + j * 64, k0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader only): M = 256 across the pair =====================
    const uint32_t idesc = umma_idesc_bf16(2 * BM, BN, p.a_mn, p.b_mn);
    int s = 0;
    uint32_t ph = 0;
    int t = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs, ++t) {
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      mbar_wait(&tempty_bar[acc], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kt = 0; kt < num_kt; ++kt) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + s * STAGE);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = p.a_mn ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                       : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bd = p.b_mn ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                       : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
            tc_mma_f16_2sm(d_tmem, ad, bd, idesc, (kt | k) != 0);
          }
          tc_commit_2sm(&empty_bar[s]);                       // frees the slot in BOTH CTAs
          if (kt == num_kt - 1) tc_commit_2sm(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): own 128 rows of the 256-row tile =====================
    const int q = warp & 3;
    const size_t elt = p.out_fp32 ? 4 : 2;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((p.ldd * elt) % 16 == 0);
    int t = 0;
    for (int tile = pair; tile < num_tiles; tile += n_pairs, ++t) {
      const int acc = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      const int m0 = (tile % tiles_m) * 2 * BM + static_cast<int>(rank) * BM;
      const int n0 = (tile / tiles_m) * BN;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      epilogue_tile<BN>(p, tmem_base + acc * BN, q, m0, n0, epi_stage + q * EPI_WARP_BYTES, vec_ok);
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // tell the leader's MMA warp
    }
  }

  tc_fence_before();
  cluster_sync_all();      // the peer's smem / TMEM stay valid until the leader's last MMA has retired
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
}

// ---- runtime-depth pipeline + cluster split-K (DSMEM reduce) ----
template <int BN, bool CLUSTER, int CONV = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16B aligned: realign to the 1024B the 128B swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = p.stages;
  const int ring_bytes = STAGES * L::STAGE_BYTES;
  const int data_bytes = (CLUSTER && L::PART_BYTES > ring_bytes) ? L::PART_BYTES : ring_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + data_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* cstat = reinterpret_cast<float*>(tmem_slot + 4);   // [4 warps][2 * BN] column statistics of this CTA's rows

  griddep_launch_dependents();  // PDL: the next kernel may start its prologue now
  const int warp = threadIdx.x >> 5;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int k_tiles_total = (p.K + BK - 1) / BK;
  const int kt_begin = blockIdx.z * p.k_tiles_per_split;
  int kt_end = kt_begin + p.k_tiles_per_split;
  if (kt_end > k_tiles_total) kt_end = k_tiles_total;
  const int num_kt = kt_end - kt_begin;  // host guarantees >= 1 for every launched z

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, BN);  // BN fp32 accumulator columns (power of two >= 32)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // PDL: everything above overlapped the previous kernel; its results are visible from here

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (p.tile_flags != nullptr) {
        // bcast_gemm: wait until the FedAvg kernel has published every arena tile under the rows
        // [n0, n0+BN) of the (K-major) weight matrix this CTA is about to TMA-load
        const int rows_here = (p.N - n0) < BN ? (p.N - n0) : BN;
        const uint32_t need = p.flag_epoch_ptr != nullptr ? *reinterpret_cast<const volatile uint32_t*>(p.flag_epoch_ptr)
                                                          : p.flag_epoch;
        wait_arrival_flags(p.tile_flags, p.flag_elem_off + static_cast<long long>(n0) * p.ldb,
                           p.flag_elem_off + static_cast<long long>(n0 + rows_here) * p.ldb - 1, p.flag_tile_elems, need);
        if (p.flag_bias_off >= 0)  // the bias slice the epilogue of this CTA will add
          wait_arrival_flags(p.tile_flags, p.flag_bias_off + n0, p.flag_bias_off + n0 + rows_here - 1,
                             p.flag_tile_elems, need);
        fence_proxy_async_all();  // order the acquires before the async-proxy (TMA) reads of global memory
      }
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < num_kt; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        const int k0 = (kt_begin + i) * BK;
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        if constexpr (CONV == 1 || CONV == 3) {
          const int kt = kt_begin + i;
          const int cblocks = p.conv_cin >> 6;
          const int tap = kt / cblocks, cb = kt - tap * cblocks;
          const int fr = tap / p.conv_kw, fs = tap - fr * p.conv_kw;
          const int q0 = m0 % p.conv_wo, t0 = m0 / p.conv_wo;
          tma_load_im2col_4d(sa, &tmA, &full_bar[s], cb * 64, q0 * p.conv_stride - p.conv_pad,
                             (t0 % p.conv_ho) * p.conv_stride - p.conv_pad, t0 / p.conv_ho, fs, fr);
        } else if (!p.a_mn) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);  // box [64 k][128 rows]
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)  // box [64 m][64 k rows] per MN atom
            tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], m0 + j * 64, k0);
        }
        if constexpr (CONV == 3) {
          const int kt = kt_begin + i;
          const int cblocks = p.conv_cin >> 6;
          const int tap = kt / cblocks, cb = kt - tap * cblocks;
          const int wcol = (p.conv_taps - 1 - tap) * p.conv_ncol + n0;
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], wcol + j * 64, cb * 64);
        } else if (!p.b_mn) {
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);  // box [64 k][BN rows]
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, k0);
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < num_kt; ++i) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: 8-row groups are 1024 B apart (SBO), K advance = 32 B inside the swizzle row.
          // MN-major: 64-wide MN atoms are 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO),
          //           K advance of 16 rows = 2048 B.
          const uint64_t ad = p.a_mn ? umma_smem_desc_sw128(sa + k * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(sa + k * 32, 16, 1024);
          const uint64_t bd = p.b_mn ? umma_smem_desc_sw128(sb + k * 2048, 8192, 1024)
                                     : umma_smem_desc_sw128(sb + k * 32, 16, 1024);
          tc_mma_f16(tmem_base, ad, bd, idesc, (i | k) != 0);
        }
        tc_commit(&empty_bar[s]);                       // frees the smem slot when the MMAs retire
        if (i == num_kt - 1) tc_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
  {
    // ===================== epilogue (phase 1): all eight warps (see the fixed kernel) =====================
    const int q = warp & 3;
    constexpr int NCHUNK = BN / 32;
    const int c_begin = (warp >= 4 ? 0 : (NCHUNK + 1) / 2) * 32, c_end = (warp >= 4 ? (NCHUNK + 1) / 2 : NCHUNK) * 32;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int lrow = q * 32 + static_cast<int>(lane_id());
    const int row = m0 + lrow;
    const size_t elt = p.out_fp32 ? 4 : 2;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((p.ldd * elt) % 16 == 0);
    float* part = reinterpret_cast<float*>(smem);  // cluster mode: reuse the (drained) operand ring
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      if constexpr (CLUSTER) {
        float4* dst = reinterpret_cast<float4*>(part + lrow * L::PART_PITCH + c);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          dst[j >> 2] = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                    __uint_as_float(r[j + 3]));
      } else {
        const int col0 = n0 + c;
        if (row >= p.M || col0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        store_row_chunk<32>(p, row, col0, v, vec_ok);
      }
    }
  }

  if constexpr (CLUSTER) {
    // ===================== cluster split-K: reduce the S partial tiles through DSMEM =====================
    const int S = p.cluster_k;
    cluster_sync_all();  // every CTA's partial tile is in its shared memory
    if (warp >= 4) {
      const uint32_t me = cluster_ctarank();
      const int rows_per = BM / S;                 // S in {2, 4, 8}
      constexpr int CG = BN / 8;                   // 8-column groups per row
      const int t = threadIdx.x - 128;             // 0..127
      const size_t elt = p.out_fp32 ? 4 : 2;
      const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((p.ldd * elt) % 16 == 0);
      const bool want_stats = p.col_stats != nullptr;
      // 128 % CG == 0: a thread always lands on the same 8 columns, so its statistics stay in registers
      float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, cq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int item = t; item < rows_per * CG; item += 128) {
        const int lrow = static_cast<int>(me) * rows_per + item / CG;
        const int c = (item % CG) * 8;
        const uint32_t laddr = smem_u32(smem) + static_cast<uint32_t>((lrow * L::PART_PITCH + c) * 4);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < S; ++r) {              // fixed order: deterministic sum
          const float4 a = ld_dsmem_f4(laddr, r), b = ld_dsmem_f4(laddr + 16, r);
          v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
          v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        const int row = m0 + lrow, col0 = n0 + c;
        if (want_stats) {                          // rows >= M are exact zeros (TMA zero fill)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float rnd = __bfloat162float(__float2bfloat16_rn(v[j]));
            cs[j] += rnd;
            cq[j] = fmaf(rnd, rnd, cq[j]);
          }
        }
        if (row < p.M && col0 < p.N) store_row_chunk<8>(p, row, col0, v, vec_ok);
      }
      if (want_stats) {
        // threads t, t + CG, t + 2 CG ... own the same 8 columns: fold the lanes of a warp with shuffles, park one
        // row of partials per warp in shared memory (plain stores), then one thread per column adds the four warps
        // and issues the global atomics.  (The previous shared-memory atomicAdd version serialised 16-way on a
        // CAS loop: +8 us per launch in the captured step, profiles/r2_mb_layers.txt.)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int o = CG; o < 32; o <<= 1) {
            cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], o);
            cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], o);
          }
        }
        const int wq = t >> 5, ln = t & 31;
        if (ln < CG) {                               // CG <= 32: lanes 0..CG-1 hold the warp's sums of columns ln*8..+7
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            cstat[wq * 2 * BN + ln * 8 + j] = cs[j];
            cstat[wq * 2 * BN + BN + ln * 8 + j] = cq[j];
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
        for (int i = t; i < 2 * BN; i += 128) {
          const int col = i < BN ? i : i - BN;
          if (n0 + col < p.N) {
            const float v = cstat[i] + cstat[2 * BN + i] + cstat[4 * BN + i] + cstat[6 * BN + i];
            atomicAdd(p.col_stats + (i < BN ? 0 : p.N) + n0 + col, v);
          }
        }
      }
    }
    cluster_sync_all();  // nobody leaves (and frees its smem) while a peer may still read it
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // cuTensorMapEncodeTiled is a DRIVER entry point: it needs a current context on the calling
  // thread.  Autograd worker threads may not have touched the runtime yet -> bind the primary
  // context once per thread (cudaFree(0) is the canonical no-op that does so).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row pitch `ld` elements;
// box = [box_cols (inner), box_rows], 128B swizzle (box_cols must be 64).
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                    int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

// 4-D bf16 tensor map: [outer][inner][rows][cols] with element strides; box = [box_cols, box_rows, 1, 1]
static int make_map4(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, long long inner,
                     long long s_inner, long long outer, long long s_outer, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return -1;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(inner),
                        static_cast<cuuint64_t>(outer)};
  cuuint64_t gstr[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(s_inner) * 2,
                        static_cast<cuuint64_t>(s_outer) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : static_cast<int>(r);
}

template <int BN, int STAGES, int CONV = 0>
static int launch_fixed(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid,
                      cudaStream_t stream) {
  constexpr int smem = STAGES * SmemLayout<BN>::STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_fixed_kernel<BN, STAGES, CONV>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  cudaError_t le = launch_pdl(gemm_bf16_fixed_kernel<BN, STAGES, CONV>, grid, GEMM_THREADS, smem, stream, ta, tb, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}


template <int BN, int STAGES>
static int launch_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_tiles,
                       cudaStream_t stream) {
  constexpr int stage = BM * BK * 2 + (BN / 2) * BK * 2;
  constexpr int smem = STAGES * stage + (2 * STAGES + 4) * 8 + 16 + 4 * EPI_WARP_BYTES + 1024;
  static bool configured = false;
  static int num_sms = 148;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_2cta_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  int pairs = num_sms / 2;
  if (pairs > num_tiles) pairs = num_tiles;
  cudaError_t le = launch_pdl(gemm_bf16_2cta_kernel<BN, STAGES>, dim3(2 * pairs), GEMM_THREADS, smem, stream, ta, tb, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

template <int BN, int STAGES>
static int launch_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_tiles,
                             cudaStream_t stream) {
  constexpr int smem = STAGES * SmemLayout<BN>::STAGE_BYTES + (2 * STAGES + 4) * 8 + 16 + 4 * EPI_WARP_BYTES + 1024;
  static bool configured = false;
  static int num_sms = 148;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_persistent_kernel<BN, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  cudaError_t le = launch_pdl(gemm_bf16_persistent_kernel<BN, STAGES>, dim3(grid), GEMM_THREADS, smem, stream, ta, tb, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

template <int BN, int CONV = 0>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid,
                      cudaStream_t stream) {
  using L = SmemLayout<BN>;
  constexpr int max_stages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  constexpr int max_smem = max_stages * L::STAGE_BYTES + (2 * MAX_STAGES + 1) * 8 + 16 + 8 * BN * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    const int cap = max_smem > L::PART_BYTES + 4096 ? max_smem : L::PART_BYTES + 4096;
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, false, CONV>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, true, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  int ring = p.stages * L::STAGE_BYTES;
  if (p.cluster_k > 1 && L::PART_BYTES > ring) ring = L::PART_BYTES;
  int smem = ring + (2 * MAX_STAGES + 1) * 8 + 16 + 8 * BN * 4 + 1024;
  // occupancy cap: CTAs of this kernel per SM (shared memory is the limiter we control)
  static int max_ctas = -1;
  if (max_ctas < 0) {
    const char* e = std::getenv("BATON_GEMM_CTAS_PER_SM");
    max_ctas = e != nullptr ? std::atoi(e) : 2;
    if (max_ctas < 1) max_ctas = 1;
  }
  // atomic epilogues (wgrad) measured slower with co-resident CTAs: keep those at one CTA per SM
  const int ctas_here = p.atomic_out ? 1 : max_ctas;
  const int floor_smem = (227 * 1024) / (ctas_here + 1) + 1024;   // > 1/(ctas+1) of the SM
  if (smem < floor_smem && floor_smem <= max_smem) smem = floor_smem;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p.cluster_k > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = p.cluster_k;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = p.cluster_k > 1 ? cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, true, CONV>, ta, tb, p)
                                   : cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, false, CONV>, ta, tb, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace b200

// D = act(alpha * A B^T + bias).  a/b: bf16 device pointers.
//   a_mn == 0: A is row-major [M, K] with pitch lda;  a_mn == 1: A is row-major [K, M] with pitch lda
//   b_mn == 0: B is row-major [N, K] with pitch ldb;  b_mn == 1: B is row-major [K, N] with pitch ldb
//   split_k > 1 with accumulate / fp32 atomic output -> atomic split-K; split_k < 0 -> cluster split-K of
//   size -split_k (2, 4 or 8) for any output type.
// Returns 0 on success, a CUDA / driver error code otherwise, -2 on unsupported alignment.
// tensor-map encoders for other translation units (attention.cu)
extern "C" int b200_encode_map2_bf16(void* map, const void* base, long long rows, long long cols, long long ld,
                                     int box_cols, int box_rows) {
  return b200::make_map(reinterpret_cast<CUtensorMap*>(map), base, rows, cols, ld, box_cols, box_rows);
}
extern "C" int b200_encode_map4_bf16(void* map, const void* base, long long rows, long long cols, long long ld,
                                     long long inner, long long s_inner, long long outer, long long s_outer,
                                     int box_cols, int box_rows) {
  return b200::make_map4(reinterpret_cast<CUtensorMap*>(map), base, rows, cols, ld, inner, s_inner, outer, s_outer,
                         box_cols, box_rows);
}

extern "C" int b200_gemm_bf16(const void* a, const void* b, void* d, const float* bias, int M, int N, int K,
                              long long lda, long long ldb, long long ldd, int a_mn, int b_mn, int out_fp32, int act,
                              int split_k, int accumulate, float alpha, const uint32_t* tile_flags,
                              uint32_t flag_epoch, long long flag_elem_off, int flag_tile_elems,
                              long long flag_bias_off, int force_bn, float* col_stats, const uint32_t* flag_epoch_ptr,
                              cudaStream_t stream) {
  using namespace b200;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  // fused BatchNorm statistics: plain single-pass GEMM only (no split-K partials, no bias / activation / scaling)
  if (col_stats != nullptr && (split_k > 1 || bias != nullptr || act != 0 || alpha != 1.0f || accumulate)) return -3;
  if ((lda % 8) || (ldb % 8) || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
    return -2;
  int bn = force_bn > 0 ? force_bn : (N > 128 ? 256 : (N > 64 ? 128 : 64));
  if (tile_flags != nullptr && (b_mn || flag_tile_elems <= 0)) return -4;
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn)
    rc = make_map(&ta, a, M, K, lda, BK, BM);
  else
    rc = make_map(&ta, a, K, M, lda, 64, BK);
  if (rc) return rc;
  if (!b_mn)
    rc = make_map(&tb, b, N, K, ldb, BK, bn);
  else
    rc = make_map(&tb, b, K, N, ldb, 64, BK);
  if (rc) return rc;

  const int k_tiles = (K + BK - 1) / BK;
  int cluster_k = 1;
  if (split_k < 0) {  // cluster split-K: every z-slice must own at least one k tile
    cluster_k = -split_k;
    if (cluster_k != 2 && cluster_k != 4 && cluster_k != 8) return -5;
    while (cluster_k > 1 && (cluster_k - 1) * ((k_tiles + cluster_k - 1) / cluster_k) >= k_tiles) cluster_k >>= 1;
    split_k = cluster_k;
  }
  if (split_k < 1) split_k = 1;
  if (split_k > k_tiles) split_k = k_tiles;
  int per = (k_tiles + split_k - 1) / split_k;
  if (cluster_k == 1) split_k = (k_tiles + per - 1) / per;  // no empty z-slices
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.D = d; p.ldd = ldd; p.bias = bias; p.out_fp32 = out_fp32; p.act = act;
  p.a_mn = a_mn; p.b_mn = b_mn; p.k_tiles_per_split = per;
  p.cluster_k = cluster_k;
  p.atomic_out = (accumulate || (split_k > 1 && cluster_k == 1)) ? 1 : 0;
  static int epi_staged = -1;
  if (epi_staged < 0) {
    const char* e = std::getenv("BATON_GEMM_EPI_STAGED");
    epi_staged = (e != nullptr && e[0] == '0') ? 0 : 1;   // default on: BERT-base round 553 -> 538 ms
  }
  p.epi_staged = epi_staged;
  p.col_stats = col_stats;
  p.tile_flags = tile_flags; p.flag_epoch = flag_epoch; p.alpha = alpha;
  p.flag_elem_off = flag_elem_off; p.flag_tile_elems = flag_tile_elems; p.ldb = ldb;
  p.flag_bias_off = (tile_flags != nullptr && bias != nullptr) ? flag_bias_off : -1;
  p.flag_epoch_ptr = tile_flags != nullptr ? flag_epoch_ptr : nullptr;
  p.batched = 0; p.batch_inner = 1; p.batch_count = 1; p.d_outer = 0; p.d_inner = 0;
  if (p.atomic_out && (!out_fp32 || bias != nullptr || act != 0)) return -3;
  const int max_stages = (bn == 256) ? 4 : (bn == 128 ? 6 : 8);
  p.stages = per < max_stages ? (per < 2 ? 2 : per) : max_stages;
  dim3 grid((N + bn - 1) / bn, (M + BM - 1) / BM, split_k);
  if (cluster_k > 1) {
    if (bn == 256) return launch_cfg<256>(ta, tb, p, grid, stream);
    if (bn == 128) return launch_cfg<128>(ta, tb, p, grid, stream);
    return launch_cfg<64>(ta, tb, p, grid, stream);
  }
  // large plain GEMMs (>= one wave of tiles, single K pass): persistent kernel with overlapped epilogue
  static int persistent_on = -1;
  if (persistent_on < 0) {
    const char* e = std::getenv("BATON_GEMM_PERSISTENT");
    persistent_on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const int num_tiles = static_cast<int>(grid.x * grid.y);
  static int two_cta_on = -1;
  if (two_cta_on < 0) {
    const char* e = std::getenv("BATON_GEMM_2CTA");
    two_cta_on = (e == nullptr) ? 2 : (e[0] == '1' ? 1 : 0);   // unset: automatic, 1: whenever legal, 0: never
  }
  // measured (profiles/r1_gemm_variants.md): with the lean epilogue CTA pairs win on every shape that fills the
  // machine (8192^3: 767 vs 801 us, BERT qkv 4096x2304x768: 21.9 vs 25.3 us, BERT-base round 521 vs 538 ms)
  const int tiles_2cta = ((M + 2 * BM - 1) / (2 * BM)) * ((N + 255) / 256);
  const bool two_cta_auto = two_cta_on == 2 && tiles_2cta >= 74;
  if ((two_cta_on == 1 || two_cta_auto) && split_k == 1 && tile_flags == nullptr && bn == 256 && num_tiles >= 148) {
    // CTA pairs: 256 x 256 tiles, each CTA TMA-loads its 128 rows of A and its 128-row half of B
    CUtensorMap tb2;
    int rc2 = !b_mn ? make_map(&tb2, b, N, K, ldb, BK, 128) : make_map(&tb2, b, K, N, ldb, 64, BK);
    if (rc2) return rc2;
    const int tiles2 = ((M + 2 * BM - 1) / (2 * BM)) * ((N + 255) / 256);
    return launch_2cta<256, 6>(ta, tb2, p, tiles2, stream);
  }
  if (persistent_on && split_k == 1 && tile_flags == nullptr && num_tiles >= 148 && bn >= 128) {
    if (bn == 256) return launch_persistent<256, 4>(ta, tb, p, num_tiles, stream);
    return launch_persistent<128, 6>(ta, tb, p, num_tiles, stream);
  }
  // short K loops (stem convolution: 3 k tiles, 1x1 shortcuts: 1-4) do not need a deep ring: a shallow one asks for
  // little shared memory, so two CTAs share an SM and a grid of > 148 tiles still runs as one wave
  const bool shallow = per <= 4 && !p.batched;
  if (bn == 256) return launch_fixed<256, 4>(ta, tb, p, grid, stream);
  if (bn == 128) return shallow ? launch_fixed<128, 3>(ta, tb, p, grid, stream) : launch_fixed<128, 6>(ta, tb, p, grid, stream);
  return shallow ? launch_fixed<64, 4>(ta, tb, p, grid, stream) : launch_fixed<64, 8>(ta, tb, p, grid, stream);
}

// Strided-batched GEMM (attention): for z = outer * n_inner + inner
//     D[z] = act(alpha * A[z] B[z]^T),  X[z] = X + outer * x_outer + inner * x_inner   (element strides)
// Same operand-major conventions as b200_gemm_bf16; every stride must be a multiple of 8 elements.
extern "C" int b200_gemm_bf16_batched(const void* a, const void* b, void* d, int M, int N, int K, long long lda,
                                      long long ldb, long long ldd, int a_mn, int b_mn, int out_fp32, int act,
                                      float alpha, int n_outer, int n_inner, long long a_outer, long long a_inner,
                                      long long b_outer, long long b_inner, long long d_outer, long long d_inner,
                                      int accumulate, cudaStream_t stream) {
  using namespace b200;
  if (M <= 0 || N <= 0 || K <= 0 || n_outer <= 0 || n_inner <= 0) return 0;
  if ((lda % 8) || (ldb % 8) || (a_outer % 8) || (a_inner % 8) || (b_outer % 8) || (b_inner % 8) ||
      (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15))
    return -2;
  // degenerate strides (size-1 dims) still need a non-zero multiple-of-16-byte stride for the encoder
  auto fix = [](long long s) { return s > 0 ? s : 8; };
  const int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn)
    rc = make_map4(&ta, a, M, K, lda, n_inner, fix(a_inner), n_outer, fix(a_outer), BK, BM);
  else
    rc = make_map4(&ta, a, K, M, lda, n_inner, fix(a_inner), n_outer, fix(a_outer), 64, BK);
  if (rc) return rc;
  if (!b_mn)
    rc = make_map4(&tb, b, N, K, ldb, n_inner, fix(b_inner), n_outer, fix(b_outer), BK, bn);
  else
    rc = make_map4(&tb, b, K, N, ldb, n_inner, fix(b_inner), n_outer, fix(b_outer), 64, BK);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.D = d; p.ldd = ldd; p.bias = nullptr; p.out_fp32 = out_fp32; p.act = act;
  p.a_mn = a_mn; p.b_mn = b_mn; p.k_tiles_per_split = (K + BK - 1) / BK; p.cluster_k = 1;
  p.atomic_out = accumulate ? 1 : 0;
  p.epi_staged = 0;
  p.col_stats = nullptr;
  p.tile_flags = nullptr; p.flag_epoch = 0; p.alpha = alpha; p.flag_elem_off = 0; p.flag_tile_elems = 0;
  p.ldb = ldb; p.flag_bias_off = -1; p.flag_epoch_ptr = nullptr; p.stages = 4;
  p.batched = 1; p.batch_inner = n_inner; p.d_outer = d_outer; p.d_inner = d_inner;
  if (p.atomic_out && !out_fp32) return -3;
  dim3 grid((N + bn - 1) / bn, (M + BM - 1) / BM, n_outer * n_inner);
  p.batch_count = n_outer * n_inner;
  // attention-sized batches are thousands of one- or two-k-tile problems: a CTA per problem is all
  // prologue (TMEM alloc, barrier init, first TMA round trip).  Persistent CTAs amortise that and overlap
  // the epilogue of problem i with the loads + MMAs of problem i+1.
  static int batched_persistent = -1;
  if (batched_persistent < 0) {
    const char* e = std::getenv("BATON_GEMM_BATCHED_PERSISTENT");
    batched_persistent = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const long long total_tiles = static_cast<long long>(grid.x) * grid.y * grid.z;
  if (batched_persistent && total_tiles >= 2 * 148 && total_tiles < (1ll << 30)) {
    p.epi_staged = 1;
    const int nt = static_cast<int>(total_tiles);
    if (bn == 256) return launch_persistent<256, 4>(ta, tb, p, nt, stream);
    if (bn == 128) return launch_persistent<128, 6>(ta, tb, p, nt, stream);
    return launch_persistent<64, 8>(ta, tb, p, nt, stream);
  }
  if (bn == 256) return launch_fixed<256, 4>(ta, tb, p, grid, stream);
  if (bn == 128) return launch_fixed<128, 6>(ta, tb, p, grid, stream);
  return launch_fixed<64, 8>(ta, tb, p, grid, stream);
}

// ---- implicit-GEMM convolution (validated on B200 in round 2; the default, BATON_CONV_IGEMM=0 turns it off) ----
extern "C" int b200_encode_map_im2col_bf16(void* map, const void* x, int N, int H, int W, int C, int KH, int KW,
                                            int stride, int pad, int channels, int pixels);

// forward: y[N*Ho*Wo, Cout] = im2col(x) w^T with x NHWC bf16 (Cin % 64 == 0), w [Cout, KH*KW*Cin] (channels_last)
extern "C" int b200_conv_igemm_fwd(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int KH,
                                   int KW, int stride, int pad, int Ho, int Wo, int cluster_k, int force_bn,
                                   float* col_stats, cudaStream_t stream) {
  using namespace b200;
  const long long M = static_cast<long long>(N) * Ho * Wo;
  const int K = KH * KW * Cin;
  if (M <= 0 || Cout <= 0) return 0;
  if (Cin % 64 != 0 || Cout % 8 != 0 || M > (1ll << 30) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return -2;
  const int bn = force_bn > 0 ? force_bn : (Cout > 128 ? 256 : (Cout > 64 ? 128 : 64));
  CUtensorMap ta, tb;
  int rc = b200_encode_map_im2col_bf16(&ta, x, N, H, W, Cin, KH, KW, stride, pad, 64, BM);
  if (rc) return rc;
  rc = make_map(&tb, w, Cout, K, K, BK, bn);
  if (rc) return rc;
  const int k_tiles = K / BK;
  if (cluster_k != 1 && cluster_k != 2 && cluster_k != 4 && cluster_k != 8) return -5;
  while (cluster_k > 1 && (cluster_k - 1) * ((k_tiles + cluster_k - 1) / cluster_k) >= k_tiles) cluster_k >>= 1;
  const int per = (k_tiles + cluster_k - 1) / cluster_k;
  GemmParams p;
  p.M = static_cast<int>(M); p.N = Cout; p.K = K; p.D = y; p.ldd = Cout; p.bias = nullptr; p.out_fp32 = 0; p.act = 0;
  p.a_mn = 0; p.b_mn = 0; p.k_tiles_per_split = per; p.cluster_k = cluster_k; p.atomic_out = 0;
  p.epi_staged = 0; p.col_stats = col_stats;
  p.tile_flags = nullptr; p.flag_epoch = 0; p.alpha = 1.0f; p.flag_elem_off = 0; p.flag_tile_elems = 0; p.ldb = K;
  p.flag_bias_off = -1; p.flag_epoch_ptr = nullptr;
  p.batched = 0; p.batch_inner = 1; p.batch_count = 1; p.d_outer = 0; p.d_inner = 0;
  const int max_stages = (bn == 256) ? 4 : (bn == 128 ? 6 : 8);
  p.stages = per < max_stages ? (per < 2 ? 2 : per) : max_stages;
  p.conv_ho = Ho; p.conv_wo = Wo; p.conv_stride = stride; p.conv_pad = pad; p.conv_kw = KW; p.conv_cin = Cin;
  p.conv_taps = KH * KW; p.conv_ncol = Cin;
  dim3 grid((Cout + bn - 1) / bn, static_cast<unsigned>((M + BM - 1) / BM), cluster_k);
  if (cluster_k > 1) {
    if (bn == 256) return launch_cfg<256, 1>(ta, tb, p, grid, stream);
    if (bn == 128) return launch_cfg<128, 1>(ta, tb, p, grid, stream);
    return launch_cfg<64, 1>(ta, tb, p, grid, stream);
  }
  const bool shallow = per <= 4;
  if (bn == 256) return launch_fixed<256, 4, 1>(ta, tb, p, grid, stream);
  if (bn == 128) return shallow ? launch_fixed<128, 3, 1>(ta, tb, p, grid, stream) : launch_fixed<128, 6, 1>(ta, tb, p, grid, stream);
  return shallow ? launch_fixed<64, 4, 1>(ta, tb, p, grid, stream) : launch_fixed<64, 8, 1>(ta, tb, p, grid, stream);
}

// input gradient of a STRIDE-1 convolution: dx[N*H*W, Cin] = im2col_{pad' = K-1-pad}(dy) * flip(w), dy NHWC bf16
// [N, Ho, Wo, Cout] (Cout % 64 == 0), w [Cout, KH*KW*Cin] channels_last (Cin % 64 == 0 so that a 64-wide N atom never
// straddles two taps).  No col buffer, no col2im, no weight transpose.
extern "C" int b200_conv_igemm_dgrad(const void* dy, const void* w, void* dx, int N, int H, int W, int Cin, int Cout,
                                     int KH, int KW, int pad, int Ho, int Wo, int cluster_k, int force_bn,
                                     cudaStream_t stream) {
  using namespace b200;
  const long long M = static_cast<long long>(N) * H * W;
  const int K = KH * KW * Cout;
  if (M <= 0 || Cin <= 0) return 0;
  const int padp = KH - 1 - pad, padq = KW - 1 - pad;
  if (Cout % 64 != 0 || Cin % 64 != 0 || M > (1ll << 30) || padp < 0 || padq < 0 || padp != padq ||
      Ho + 2 * padp - (KH - 1) != H || Wo + 2 * padq - (KW - 1) != W || (reinterpret_cast<uintptr_t>(dy) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(dx) & 15))
    return -2;
  const int bn = force_bn > 0 ? force_bn : (Cin > 128 ? 256 : (Cin > 64 ? 128 : 64));
  CUtensorMap ta, tb;
  int rc = b200_encode_map_im2col_bf16(&ta, dy, N, Ho, Wo, Cout, KH, KW, 1, padp, 64, BM);
  if (rc) return rc;
  rc = make_map(&tb, w, Cout, static_cast<long long>(KH) * KW * Cin, static_cast<long long>(KH) * KW * Cin, 64, BK);
  if (rc) return rc;
  const int k_tiles = K / BK;
  if (cluster_k != 1 && cluster_k != 2 && cluster_k != 4 && cluster_k != 8) return -5;
  while (cluster_k > 1 && (cluster_k - 1) * ((k_tiles + cluster_k - 1) / cluster_k) >= k_tiles) cluster_k >>= 1;
  const int per = (k_tiles + cluster_k - 1) / cluster_k;
  GemmParams p;
  p.M = static_cast<int>(M); p.N = Cin; p.K = K; p.D = dx; p.ldd = Cin; p.bias = nullptr; p.out_fp32 = 0; p.act = 0;
  p.a_mn = 0; p.b_mn = 1; p.k_tiles_per_split = per; p.cluster_k = cluster_k; p.atomic_out = 0;
  p.epi_staged = 0; p.col_stats = nullptr;
  p.tile_flags = nullptr; p.flag_epoch = 0; p.alpha = 1.0f; p.flag_elem_off = 0; p.flag_tile_elems = 0;
  p.ldb = static_cast<long long>(KH) * KW * Cin;
  p.flag_bias_off = -1; p.flag_epoch_ptr = nullptr;
  p.batched = 0; p.batch_inner = 1; p.batch_count = 1; p.d_outer = 0; p.d_inner = 0;
  const int max_stages = (bn == 256) ? 4 : (bn == 128 ? 6 : 8);
  p.stages = per < max_stages ? (per < 2 ? 2 : per) : max_stages;
  p.conv_ho = H; p.conv_wo = W; p.conv_stride = 1; p.conv_pad = padp; p.conv_kw = KW; p.conv_cin = Cout;
  p.conv_taps = KH * KW; p.conv_ncol = Cin;
  dim3 grid((Cin + bn - 1) / bn, static_cast<unsigned>((M + BM - 1) / BM), cluster_k);
  if (cluster_k > 1) {
    if (bn == 256) return launch_cfg<256, 3>(ta, tb, p, grid, stream);
    if (bn == 128) return launch_cfg<128, 3>(ta, tb, p, grid, stream);
    return launch_cfg<64, 3>(ta, tb, p, grid, stream);
  }
  if (bn == 256) return launch_fixed<256, 4, 3>(ta, tb, p, grid, stream);
  if (bn == 128) return launch_fixed<128, 6, 3>(ta, tb, p, grid, stream);
  return launch_fixed<64, 8, 3>(ta, tb, p, grid, stream);
}

// weight gradient: dw[Cout, KH*KW*Cin] (fp32, accumulated with red.add) += dy[N*Ho*Wo, Cout]^T im2col(x)
extern "C" int b200_conv_igemm_wgrad(const void* dy, const void* x, float* dw, int N, int H, int W, int Cin, int Cout,
                                     int KH, int KW, int stride, int pad, int Ho, int Wo, int split_k, int force_bn,
                                     cudaStream_t stream) {
  using namespace b200;
  const long long Mp = static_cast<long long>(N) * Ho * Wo;      // reduction length (output pixels)
  const int Kc = KH * KW * Cin;                                    // GEMM N
  if (Mp <= 0 || Cout <= 0) return 0;
  if (Cin % 64 != 0 || Cout % 8 != 0 || Mp > (1ll << 30) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(dw) & 15))
    return -2;
  const int bn = force_bn > 0 ? force_bn : (Kc > 128 ? 256 : (Kc > 64 ? 128 : 64));
  CUtensorMap ta, tb;
  int rc = make_map(&ta, dy, Mp, Cout, Cout, 64, BK);              // MN-major A: rows = pixels (K), cols = Cout (M)
  if (rc) return rc;
  rc = b200_encode_map_im2col_bf16(&tb, x, N, H, W, Cin, KH, KW, stride, pad, 64, 64);
  if (rc) return rc;
  const int k_tiles = static_cast<int>((Mp + BK - 1) / BK);
  if (split_k < 1) split_k = 1;
  if (split_k > k_tiles) split_k = k_tiles;
  const int per = (k_tiles + split_k - 1) / split_k;
  split_k = (k_tiles + per - 1) / per;
  GemmParams p;
  p.M = Cout; p.N = Kc; p.K = static_cast<int>(Mp); p.D = dw; p.ldd = Kc; p.bias = nullptr; p.out_fp32 = 1; p.act = 0;
  p.a_mn = 1; p.b_mn = 1; p.k_tiles_per_split = per; p.cluster_k = 1; p.atomic_out = 1;
  p.epi_staged = 0; p.col_stats = nullptr;
  p.tile_flags = nullptr; p.flag_epoch = 0; p.alpha = 1.0f; p.flag_elem_off = 0; p.flag_tile_elems = 0; p.ldb = Kc;
  p.flag_bias_off = -1; p.flag_epoch_ptr = nullptr;
  p.batched = 0; p.batch_inner = 1; p.batch_count = 1; p.d_outer = 0; p.d_inner = 0;
  p.stages = 4;
  p.conv_ho = Ho; p.conv_wo = Wo; p.conv_stride = stride; p.conv_pad = pad; p.conv_kw = KW; p.conv_cin = Cin;
  p.conv_taps = KH * KW; p.conv_ncol = Cin;
  dim3 grid((Kc + bn - 1) / bn, (Cout + BM - 1) / BM, split_k);
  if (bn == 256) return launch_fixed<256, 4, 2>(ta, tb, p, grid, stream);
  if (bn == 128) return launch_fixed<128, 6, 2>(ta, tb, p, grid, stream);
  return launch_fixed<64, 8, 2>(ta, tb, p, grid, stream);
}

B200_TRACE_REGISTER(gemm_tcgen05)
