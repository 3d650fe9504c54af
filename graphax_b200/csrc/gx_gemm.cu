// gx_gemm.cu - dense x dense contraction on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// The reference's `_pure_dot_product_mul` (sparse/tensor.py:999-1059) contracts two dense edge
// Jacobians with `lax.dot_general`.  For the per-sample graphs of configs 1-4 those are <= 3x3 and live
// in registers (gx_skeleton.cuh); for tensor-valued vertices (config 5: weight Jacobians of an MLP over
// a batch, [B,H]^T [B,H'] products) they are real GEMMs and go here:
//
//     C[M,N] (+)= A[M,K] * B[N,K]^T        A, B bf16 row-major (K contiguous), fp32 accumulate in TMEM
//
// One CTA computes one 128 x 128 tile (optionally one K-split of it):
//   warp 0   TMA producer  - cp.async.bulk.tensor.2d of 128x64 A and B boxes (SWIZZLE_128B) into a
//                            kStages-deep shared-memory ring, completion on per-stage mbarriers
//   warp 1   MMA issuer    - allocates 128 TMEM columns; one elected lane issues tcgen05.mma
//                            (kind::f16, M128 N128 K16) four times per stage and tcgen05.commit's the
//                            stage back to the producer, finally signals the epilogue
//   warps 2-5 epilogue     - tcgen05.ld the fp32 accumulator (32 lanes x 32 columns at a time), then
//                            store fp32 / bf16 rows or red.global.add.f32 them (split-K)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <string>

#include "../../include/graphax_b200.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16;
constexpr int kStages = 3;   // 96 KB of operand ring: two CTAs per SM, one's epilogue overlaps the other's main loop
constexpr int kThreads = 192;
constexpr int kTmemCols = 128;
constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr size_t kSmemBytes = kStages * (kStageBytesA + kStageBytesB) + 1024 /*align*/ + 256 /*barriers*/;

enum { EPI_STORE_F32 = 0, EPI_STORE_BF16 = 1, EPI_ATOMIC_F32 = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nGG_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra GG_DONE;\nbra GG_WAIT;\nGG_DONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile: 8-row groups of 1024 B (SBO), LBO unused (=1), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand (stored [K, MN] with the M/N axis contiguous), 128-byte swizzle: the tile is two boxes of
// [64 K-rows x 64 MN-elements] (8 KB each); LBO = distance between the two 64-element MN atoms, SBO = distance
// between groups of 8 K-rows (8 x 128 B)
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(8192 >> 4) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D fp32, A/B bf16, M = 128, N = 128; bit 15 / 16 select MN-major A / B
constexpr uint32_t kInstrDescBase = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

struct GemmArgs {
  float* c_f32;
  __nv_bfloat16* c_bf16;
  int M, N, K, ldc, epilogue, k_per_split;
  int a_mn, b_mn;   // operand stored [K, M] / [K, N] (MN-major) instead of [M, K] / [N, K]
};

__global__ void __launch_bounds__(kThreads, 2)
gx_gemm_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const GemmArgs args) {
  extern __shared__ uint8_t gg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gg_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kStageBytesA;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * (kStageBytesA + kStageBytesB));
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * args.k_per_split;
  const int num_k_blocks = args.k_per_split / BK;

  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the allocating warp also owns the deallocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&empty_bar[s], ((kb / kStages) & 1) ^ 1);  // slot free (passes immediately the first round)
        mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
        const int k0 = k_begin + kb * BK;
        if (!args.a_mn) {
          tma_load_2d(smem_a + s * kStageBytesA, &map_a, k0, m0, &full_bar[s]);
        } else {   // two [64 k x 64 m] boxes
          tma_load_2d(smem_a + s * kStageBytesA, &map_a, m0, k0, &full_bar[s]);
          tma_load_2d(smem_a + s * kStageBytesA + 8192, &map_a, m0 + 64, k0, &full_bar[s]);
        }
        if (!args.b_mn) {
          tma_load_2d(smem_b + s * kStageBytesB, &map_b, k0, n0, &full_bar[s]);
        } else {
          tma_load_2d(smem_b + s * kStageBytesB, &map_b, n0, k0, &full_bar[s]);
          tma_load_2d(smem_b + s * kStageBytesB + 8192, &map_b, n0 + 64, k0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    for (int kb = 0; kb < num_k_blocks; ++kb) {
      const int s = kb % kStages;
      mbar_wait(&full_bar[s], (kb / kStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem_a + s * kStageBytesA), sb = smem_u32(smem_b + s * kStageBytesB);
        const uint64_t da = args.a_mn ? make_smem_desc_mn(sa) : make_smem_desc(sa);
        const uint64_t db = args.b_mn ? make_smem_desc_mn(sb) : make_smem_desc(sb);
        // one MMA consumes 16 K: K-major = 32 B inside the swizzle atom (+2 in the 16-byte address field),
        // MN-major = 16 rows of 128 B (+128)
        const uint32_t step_a = args.a_mn ? 128u : 2u, step_b = args.b_mn ? 128u : 2u;
        const uint32_t idesc = kInstrDescBase | (args.a_mn ? (1u << 15) : 0u) | (args.b_mn ? (1u << 16) : 0u);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_f16(tmem_base, da + step_a * k, db + step_b * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);                       // smem slot reusable once these MMAs have read it
        if (kb == num_k_blocks - 1) umma_commit(tmem_full_bar);  // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes 32*(w%4) .. 32*(w%4)+31 =====
    const int quad = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m0 + quad * 32 + lane;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < args.M) {
        const size_t off = (size_t)row * args.ldc + n0 + c;
        if (args.epilogue == EPI_STORE_F32) {
          float4* dst = reinterpret_cast<float4*>(args.c_f32 + off);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
        } else if (args.epilogue == EPI_STORE_BF16) {
          uint4* dst = reinterpret_cast<uint4*>(args.c_bf16 + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 2 * q]), __uint_as_float(v[8 * j + 2 * q + 1]));
              w[q] = *reinterpret_cast<uint32_t*>(&h);
            }
            dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(args.c_f32 + off + j, __uint_as_float(v[j]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// [rows, cols] bf16 -> [cols, rows] bf16 (operands whose contraction axis is not contiguous)
__global__ void gx_transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows,
                                         int cols) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

thread_local std::string g_gemm_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess) fn = (EncodeTiledFn)p;
  });
  return fn;
}

// row-major [rows, cols] bf16 matrix, box of box_rows x 64 elements (64 elements = one 128-byte swizzle row)
int make_map(CUtensorMap* map, const void* ptr, int rows, int k, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  cuuint32_t box[2] = {(cuuint32_t)64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

}  // namespace

extern "C" {

const char* gx_gemm_last_error(void) { return g_gemm_error.c_str(); }

// C[M,N] (+)= A[M,K] * B[N,K]^T; A, B bf16 row-major device pointers; C fp32 (c_bf16 == 0) or bf16 row-major.
// k_splits > 1 accumulates with fp32 atomics into C, which the caller must have zeroed (fp32 only).
int gx_gemm_bf16_tn(int M, int N, int K, const void* A, const void* B, void* C, int c_is_bf16, int k_splits,
                    void* stream) {
  return gx_gemm_bf16(M, N, K, A, 0, B, 0, C, c_is_bf16, k_splits, stream);
}

// General form: a_mn != 0 means A is stored [K, M] row-major (M contiguous), likewise b_mn for B = [K, N].
int gx_gemm_bf16(int M, int N, int K, const void* A, int a_mn, const void* B, int b_mn, void* C, int c_is_bf16,
                 int k_splits, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) {
    g_gemm_error = "gx_gemm_bf16_tn: bad argument";
    return GX_ERR_BAD_ARG;
  }
  if (M % BM || N % BN || K % BK || k_splits < 1 || (K / BK) % k_splits || (k_splits > 1 && c_is_bf16)) {
    g_gemm_error = "gx_gemm_bf16_tn: M, N must be multiples of 128, K of 64 x k_splits; split-K needs an fp32 C";
    return GX_ERR_UNSUPPORTED;
  }
  CUtensorMap map_a, map_b;
  const int ra = a_mn ? make_map(&map_a, A, K, M, BK) : make_map(&map_a, A, M, K, BM);
  const int rb = b_mn ? make_map(&map_b, B, K, N, BK) : make_map(&map_b, B, N, K, BN);
  if (ra || rb) {
    g_gemm_error = "gx_gemm_bf16_tn: cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)";
    return GX_ERR_CUDA;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  });
  if (attr_err != cudaSuccess) {
    g_gemm_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(attr_err);
    return GX_ERR_CUDA;
  }
  GemmArgs args;
  args.c_f32 = c_is_bf16 ? nullptr : static_cast<float*>(C);
  args.c_bf16 = c_is_bf16 ? static_cast<__nv_bfloat16*>(C) : nullptr;
  args.M = M;
  args.N = N;
  args.K = K;
  args.ldc = N;
  args.epilogue = k_splits > 1 ? EPI_ATOMIC_F32 : (c_is_bf16 ? EPI_STORE_BF16 : EPI_STORE_F32);
  args.k_per_split = K / k_splits;
  args.a_mn = a_mn ? 1 : 0;
  args.b_mn = b_mn ? 1 : 0;
  dim3 grid(N / BN, M / BM, k_splits);
  gx_gemm_tn_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(map_a, map_b, args);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_gemm_error = std::string("gx_gemm_tn_kernel launch: ") + cudaGetErrorString(e);
    return GX_ERR_CUDA;
  }
  return GX_OK;
}

int gx_transpose_bf16(int rows, int cols, const void* in, void* out, void* stream) {
  if (rows <= 0 || cols <= 0 || !in || !out) {
    g_gemm_error = "gx_transpose_bf16: bad argument";
    return GX_ERR_BAD_ARG;
  }
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  gx_transpose_bf16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), rows, cols);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_gemm_error = std::string("gx_transpose_bf16 launch: ") + cudaGetErrorString(e);
    return GX_ERR_CUDA;
  }
  return GX_OK;
}

}  // extern "C"
