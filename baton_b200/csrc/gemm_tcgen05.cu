// bf16 GEMM on the 5th-gen tensor cores: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (accumulator in TMEM) -> tcgen05.ld epilogue with fused bias / activation / cast.
//
//     D[M,N] = act( A[M,K] * B[N,K]^T + bias[N] )          (A, B bf16; accumulate fp32)
//
// Each operand may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]), so the
// three training GEMMs need no transposes:
//     fwd    Y  = X  W^T      A = X   (K-major)   B = W  (K-major)
//     dgrad  dX = dY W        A = dY  (K-major)   B = W  (MN-major, W stored [N_out... K_red] rows)
//     wgrad  dW = dY^T X      A = dY  (MN-major)  B = X  (MN-major)
//
// Warp roles (256 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
// lane), warp 2 = TMEM allocator, warps 4-7 = epilogue (warp q reads TMEM lanes 32q..32q+31).
// Split-K (gridDim.z) accumulates with fp32 red.global.add into a zero-initialised D.
//
// Flag-gated variant ("bcast_gemm", K3 in SURVEY.md 2.6): the producer acquires a per-N-tile
// arrival flag (written by the broadcast kernel with st.release.sys) before issuing the TMA
// loads of a B tile, so the first GEMM of a round consumes the new global weights tile by tile
// as they land over NVLink.
#include "ptx.cuh"
#include "launch.h"
#include "pdl.cuh"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;

struct GemmParams {
  int M, N, K;
  void* D;
  long long ldd;          // leading dimension of D in elements
  const float* bias;      // [N] or nullptr
  int out_fp32;           // 1: D is fp32, 0: D is bf16
  int act;                // 0 none, 1 relu, 2 gelu(tanh)
  int a_mn, b_mn;         // operand majors
  int k_tiles_per_split;  // split-K: k tiles handled by one z-slice
  int atomic_out;         // 1: red.add fp32 into D (split-K)
  const uint32_t* tile_flags;  // optional arrival flags, one per arena tile (bcast_gemm)
  uint32_t flag_epoch;         // value a flag must reach before the data under it may be loaded
  long long flag_elem_off;     // arena element offset of B[0,0]
  int flag_tile_elems;         // arena elements covered by one flag
  long long flag_bias_off;     // arena element offset of bias[0], or -1
  long long ldb;               // row pitch of B (elements)
  float alpha;
};

template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
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

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
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
        const long long e0 = p.flag_elem_off + static_cast<long long>(n0) * p.ldb;
        const long long e1 = p.flag_elem_off + static_cast<long long>(n0 + rows_here) * p.ldb - 1;
        for (long long t = e0 / p.flag_tile_elems; t <= e1 / p.flag_tile_elems; ++t) {
          while (ld_acquire_sys(p.tile_flags + t) < p.flag_epoch) {
          }
        }
        if (p.flag_bias_off >= 0) {  // the bias slice the epilogue of this CTA will add
          const long long b0 = p.flag_bias_off + n0, b1 = p.flag_bias_off + n0 + rows_here - 1;
          for (long long t = b0 / p.flag_tile_elems; t <= b1 / p.flag_tile_elems; ++t) {
            while (ld_acquire_sys(p.tile_flags + t) < p.flag_epoch) {
            }
          }
        }
        fence_proxy_async_all();  // order the acquires before the async-proxy (TMA) reads of global memory
      }
      for (int i = 0; i < num_kt; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        const int k0 = (kt_begin + i) * BK;
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        if (!p.a_mn) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);  // box [64 k][128 rows]
        } else {
#pragma unroll
          for (int j = 0; j < BM / 64; ++j)  // box [64 m][64 k rows] per MN atom
            tma_load_2d(sa + j * 8192, &tmA, &full_bar[s], m0 + j * 64, k0);
        }
        if (!p.b_mn) {
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);  // box [64 k][BN rows]
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[s], n0 + j * 64, k0);
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
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + static_cast<int>(lane_id());
    const bool row_ok = row < p.M;
    const size_t elt = p.out_fp32 ? 4 : 2;
    uint8_t* drow = reinterpret_cast<uint8_t*>(p.D) + static_cast<size_t>(row) * p.ldd * elt;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) && ((p.ldd * elt) % 16 == 0);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      const int col0 = n0 + c;
      if (!row_ok || col0 >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]) * p.alpha;
        if (p.bias != nullptr && (col0 + j) < p.N) x += p.bias[col0 + j];
        v[j] = apply_act(x, p.act);
      }
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
          for (int j = 0; j < 32 && col0 + j < p.N; ++j) atomicAdd(d + j, v[j]);
        }
      } else if (p.out_fp32) {
        float* d = reinterpret_cast<float*>(drow) + col0;
        if (full && vec_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          for (int j = 0; j < 32 && col0 + j < p.N; ++j) d[j] = v[j];
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
          for (int j = 0; j < 32 && col0 + j < p.N; ++j) d[j] = __float2bfloat16_rn(v[j]);
        }
      }
    }
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

template <int BN, int STAGES>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid,
                      cudaStream_t stream) {
  constexpr int smem = STAGES * SmemLayout<BN>::STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  cudaError_t le = launch_pdl(gemm_bf16_tcgen05_kernel<BN, STAGES>, grid, GEMM_THREADS, smem, stream, ta, tb, p);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace b200

// D = act(alpha * A B^T + bias).  a/b: bf16 device pointers.
//   a_mn == 0: A is row-major [M, K] with pitch lda;  a_mn == 1: A is row-major [K, M] with pitch lda
//   b_mn == 0: B is row-major [N, K] with pitch ldb;  b_mn == 1: B is row-major [K, N] with pitch ldb
// Returns 0 on success, a CUDA / driver error code otherwise, -2 on unsupported alignment.
extern "C" int b200_gemm_bf16(const void* a, const void* b, void* d, const float* bias, int M, int N, int K,
                              long long lda, long long ldb, long long ldd, int a_mn, int b_mn, int out_fp32, int act,
                              int split_k, int accumulate, float alpha, const uint32_t* tile_flags,
                              uint32_t flag_epoch,
                              long long flag_elem_off, int flag_tile_elems, long long flag_bias_off, int force_bn,
                              cudaStream_t stream) {
  using namespace b200;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
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
  if (split_k < 1) split_k = 1;
  if (split_k > k_tiles) split_k = k_tiles;
  int per = (k_tiles + split_k - 1) / split_k;
  split_k = (k_tiles + per - 1) / per;  // no empty z-slices
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.D = d; p.ldd = ldd; p.bias = bias; p.out_fp32 = out_fp32; p.act = act;
  p.a_mn = a_mn; p.b_mn = b_mn; p.k_tiles_per_split = per; p.atomic_out = (split_k > 1 || accumulate) ? 1 : 0;
  p.tile_flags = tile_flags; p.flag_epoch = flag_epoch; p.alpha = alpha;
  p.flag_elem_off = flag_elem_off; p.flag_tile_elems = flag_tile_elems; p.ldb = ldb;
  p.flag_bias_off = (tile_flags != nullptr && bias != nullptr) ? flag_bias_off : -1;
  if (p.atomic_out && (!out_fp32 || bias != nullptr || act != 0)) return -3;
  dim3 grid((N + bn - 1) / bn, (M + BM - 1) / BM, split_k);
  if (bn == 256) return launch_cfg<256, 4>(ta, tb, p, grid, stream);
  if (bn == 128) return launch_cfg<128, 6>(ta, tb, p, grid, stream);
  return launch_cfg<64, 8>(ta, tb, p, grid, stream);
}
