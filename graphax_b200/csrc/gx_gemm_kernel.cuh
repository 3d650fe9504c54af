// gx_gemm_kernel.cuh - device side of the dense x dense contraction on the 5th-generation tensor cores
// (tcgen05 + TMEM + TMA).  Included by gx_gemm.cu (compiled into the library with the plain store / bf16 /
// split-K epilogues) and compiled at run time with NVRTC in front of a generated epilogue (GX_GEMM_FUSED:
// graphax_b200/dense.py emits `struct GxEpi`, `gx_epi_chunk` and `gx_epi_finish`), so that the elementwise
// half of the reference's products (`_pure_broadcast_mul`, sparse/tensor.py:826-866, and the elemental
// partials, primitives.py:150-205) runs on the accumulator while it is still in registers.
//
//     C[M,N] (+)= op(A) * op(B)^T     A, B bf16; K-major ([M,K] / [N,K] row-major) or MN-major ([K,M] / [K,N])
//
// One CTA computes one 128 x 128 tile (optionally one K-split of it):
//   warp 0   TMA producer  - cp.async.bulk.tensor.2d of 128x64 A and B boxes (SWIZZLE_128B) into a
//                            kStages-deep shared-memory ring, completion on per-stage mbarriers
//   warp 1   MMA issuer    - allocates 128 TMEM columns; one elected lane issues tcgen05.mma
//                            (kind::f16, M128 N128 K16) four times per stage and tcgen05.commit's the
//                            stage back to the producer, finally signals the epilogue
//   warps 2-9 epilogue     - tcgen05.ld the fp32 accumulator (32 lanes x 32 columns at a time; two warps per
//                            TMEM lane quadrant, each owning 64 of the 128 columns), then the epilogue: store
//                            fp32 / bf16 rows, red.global.add.v4.f32 (split-K), or the fused one
// The prologue (barrier init, TMEM allocation, descriptor prefetch) may overlap the tail of the previous
// kernel on the stream (programmatic dependent launch); nothing touches global memory before
// griddepcontrol.wait.
#pragma once

#ifdef GX_NVRTC
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
struct alignas(64) CUtensorMap {
  uint64_t opaque[16];
};
#else
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace gxg {

constexpr int BM = 128, BN = 128, BK = 64, UMMA_K = 16;
// operand ring depth: 3 stages = 96 KB, two CTAs per SM (one's epilogue overlaps the other's main loop); grids that
// fill the GPU only once (<= one CTA per SM) use 6 stages so that twice as many bytes are in flight per SM
constexpr int kStagesShallow = 3, kStagesDeep = 6;
// CTA-pair variant (cta_group::2, below): a stage is 16 KB of A + 8 KB of B, so the same shared memory holds 4 / 8 stages;
// the stage count also names the variant on the host side (3 / 6 = one CTA per tile, 4 / 8 = CTA pairs)
constexpr int kStagesPairShallow = 4, kStagesPairDeep = 8;
#ifdef GX_NVRTC
#define GX_HD __device__
#else
#define GX_HD __host__ __device__
#endif
GX_HD constexpr bool stages_deep(int stages) { return stages == kStagesDeep || stages == kStagesPairDeep; }
GX_HD constexpr bool stages_pair(int stages) { return stages == kStagesPairShallow || stages == kStagesPairDeep; }
constexpr int kThreads = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int kTmemCols = 128;
constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
#ifndef GX_NVRTC
constexpr uint32_t smem_bytes(int stages) {
  return stages * (kStageBytesA + (stages_pair(stages) ? kStageBytesB / 2 : kStageBytesB)) + 1024 /*align*/ + 256 /*barriers*/;
}
#endif

enum { EPI_STORE_F32 = 0, EPI_STORE_BF16 = 1, EPI_ATOMIC_F32 = 2 };

struct GemmArgs {
  float* c_f32;
  uint16_t* c_bf16;
  int M, N, K, ldc, epilogue, k_per_split;
  int a_mn, b_mn;   // operand stored [K, M] / [K, N] (MN-major) instead of [M, K] / [N, K]
  // optional phase stamps of the single-tile kernels (tools/gemm_trace.py): 8 x u64 per CTA - %globaltimer at
  // entry, after griddepcontrol.wait, first operand stage landed, accumulator complete, epilogue done, exit; slot 6: %smid
  unsigned long long* trace;
  // optional: `zero_n` 16-byte words at `zero_ptr` are cleared by the (otherwise idle) epilogue warps of the single-tile
  // kernels while the main loop runs - the step's accumulator arena, so that no separate memset has to be launched
  // and waited for (gx_gemm_zero_with_next_launch)
  float4* zero_ptr;
  unsigned zero_n;
};
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int cta, int slot) {
  if (trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    trace[(size_t)cta * 8 + slot] = t;
  }
}

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
// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC compute one 256 x 128 tile.  Each
// loads its own 128 rows of A and HALF of the B tile (64 of the 128 N rows); one MMA (M256 N128 K16), issued by the
// leader CTA, reads both shared memories and writes both tensor memories: 24 KB of operands per CTA and k-block
// instead of 32 KB.  The MLP-sized contractions are bound by exactly that traffic.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// the pair's loads complete on the LEADER's barrier (bar_cluster_addr: a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// fp32 -> bf16 bits, round to nearest even (NaN stays NaN)
__device__ __forceinline__ uint32_t f32_to_bf16_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {   // one F2FP instruction, round to nearest even
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t h) { return __uint_as_float(h << 16); }

// Sum v[j] over the 32 lanes of the warp for every j at once: recursive halving, 31 shuffles in total.
// Afterwards v[0] of lane l holds the sum over all lanes of the original v[l].
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// Epilogue staging.  tcgen05.ld hands every lane one ROW of the chunk (32 consecutive columns), so direct stores
// touch 32 different lines per instruction with 8-16 bytes each, and the LSU -> L2 transaction rate, not bandwidth,
// bounds the epilogue (measured: ~8 us per million partial-line stores).  Each warp therefore passes its 32 x 32
// chunk through a private 4 KB shared-memory tile (the operand ring is free once the accumulator is complete;
// 16-byte groups XOR-swizzled by the row so that both directions are conflict-free) and writes global memory
// with row-contiguous 128-bit accesses: four full 128-byte lines per instruction.
__device__ __forceinline__ void stage_rows(float* tile, const float (&y)[32], int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(tile + lane * 32 + ((q ^ (lane & 7)) << 2)) =
        make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
  __syncwarp();
}
// i-th read-back step (i = 0..7): this lane's 4 consecutive columns (lane & 7) * 4 .. of tile row 4 * i + (lane >> 3)
__device__ __forceinline__ float4 staged_quad(const float* tile, int i, int lane) {
  const int r = 4 * i + (lane >> 3);
  return *reinterpret_cast<const float4*>(tile + r * 32 + (((lane & 7) ^ (r & 7)) << 2));
}
// The reverse direction for [M, N] operands of a fused epilogue: row-contiguous 128-bit loads into the tile, then
// every lane picks up its own row.  (Per-lane row loads would hit 32 lines per instruction, and with two CTAs'
// operand rings resident there is too little L1 left to keep those lines until their other 7 quads are read.)
__device__ __forceinline__ void unstage_rows(const float* tile, float (&y)[32], int lane) {
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(tile + lane * 32 + ((q ^ (lane & 7)) << 2));
    y[4 * q] = t.x, y[4 * q + 1] = t.y, y[4 * q + 2] = t.z, y[4 * q + 3] = t.w;
  }
  __syncwarp();
}
__device__ __forceinline__ void load_rows_f32(float* tile, float (&y)[32], const float* base, int ld, int m_base, int M,
                                              int col0, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3), row = m_base + r;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < M) t = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * ld + col0 + (lane & 7) * 4));
    *reinterpret_cast<float4*>(tile + r * 32 + (((lane & 7) ^ (r & 7)) << 2)) = t;
  }
  unstage_rows(tile, y, lane);
}
// load_rows_f32 in two halves: the global loads (issued early, results kept in registers) and the pass through the tile
__device__ __forceinline__ void fetch_rows_f32(float4 (&r)[8], const float* base, int ld, int m_base, int M, int col0, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_base + 4 * i + (lane >> 3);
    r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < M) r[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * ld + col0 + (lane & 7) * 4));
  }
}
__device__ __forceinline__ void spread_rows_f32(float* tile, float (&y)[32], const float4 (&r)[8], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + (lane >> 3);
    *reinterpret_cast<float4*>(tile + rr * 32 + (((lane & 7) ^ (rr & 7)) << 2)) = r[i];
  }
  unstage_rows(tile, y, lane);
}
__device__ __forceinline__ void load_rows_bf16(float* tile, float (&y)[32], const uint16_t* base, int ld, int m_base, int M,
                                               int col0, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + (lane >> 3), row = m_base + r;
    uint2 u = make_uint2(0u, 0u);
    if (row < M) u = __ldg(reinterpret_cast<const uint2*>(base + (size_t)row * ld + col0 + (lane & 7) * 4));
    *reinterpret_cast<float4*>(tile + r * 32 + (((lane & 7) ^ (r & 7)) << 2)) =
        make_float4(bf16_bits_to_f32(u.x & 0xffffu), bf16_bits_to_f32(u.x >> 16), bf16_bits_to_f32(u.y & 0xffffu),
                    bf16_bits_to_f32(u.y >> 16));
  }
  unstage_rows(tile, y, lane);
}
__device__ __forceinline__ void store_rows_f32(float* tile, const float (&y)[32], float* base, int ld, int m_base, int M,
                                               int col0, int lane) {
  stage_rows(tile, y, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_base + 4 * i + (lane >> 3);
    if (row < M) *reinterpret_cast<float4*>(base + (size_t)row * ld + col0 + (lane & 7) * 4) = staged_quad(tile, i, lane);
  }
  __syncwarp();
}
__device__ __forceinline__ void store_rows_bf16(float* tile, const float (&y)[32], uint16_t* base, int ld, int m_base, int M,
                                                int col0, int lane) {
  stage_rows(tile, y, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_base + 4 * i + (lane >> 3);
    const float4 t = staged_quad(tile, i, lane);
    if (row < M)
      *reinterpret_cast<uint2*>(base + (size_t)row * ld + col0 + (lane & 7) * 4) =
          make_uint2(pack_bf16x2(t.x, t.y), pack_bf16x2(t.z, t.w));
  }
  __syncwarp();
}
// both copies from one staging pass (either pointer may be null)
__device__ __forceinline__ void store_rows(float* tile, const float (&y)[32], float* o32, uint16_t* o16, int ld, int m_base,
                                           int M, int col0, int lane) {
  if (!o32 && !o16) return;
  stage_rows(tile, y, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_base + 4 * i + (lane >> 3);
    const float4 t = staged_quad(tile, i, lane);
    if (row < M) {
      const size_t off = (size_t)row * ld + col0 + (lane & 7) * 4;
      if (o32) *reinterpret_cast<float4*>(o32 + off) = t;
      if (o16) *reinterpret_cast<uint2*>(o16 + off) = make_uint2(pack_bf16x2(t.x, t.y), pack_bf16x2(t.z, t.w));
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void red_rows_f32(float* tile, const float (&y)[32], float* base, int ld, int m_base, int M,
                                             int col0, int lane) {
  stage_rows(tile, y, lane);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_base + 4 * i + (lane >> 3);
    const float4 t = staged_quad(tile, i, lane);
    if (row < M)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + (size_t)row * ld + col0 + (lane & 7) * 4),
                   "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w)
                   : "memory");
  }
  __syncwarp();
}

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
constexpr uint32_t kInstrDescPair = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | (((2 * BM) >> 4) << 24);   // M = 256

}  // namespace gxg

#ifndef GX_GEMM_KERNEL_NAME
#define GX_GEMM_KERNEL_NAME gx_gemm_tn_kernel
#endif

#ifdef GX_GEMM_FUSED
#define GX_GEMM_EPI_PARAM , const __grid_constant__ GxEpi epi
#define GX_GEMM_EPI_PARAM_REF , const GxEpi& epi
#else
#define GX_GEMM_EPI_PARAM
#define GX_GEMM_EPI_PARAM_REF
#endif

#ifdef GX_GEMM_FUSED
#define GX_GEMM_EPI_ARG , epi
#else
#define GX_GEMM_EPI_ARG
#endif

// tile_m / tile_n / split: which 128 x 128 tile (and which K split of it) this CTA computes; cta_linear / n_ctas: its
// place among the CTAs that share args.trace / args.zero_ptr
struct GxTileCoord {
  int tile_m, tile_n, split, cta_linear, n_ctas;
};
template <int kStages>
__device__ __forceinline__ GxTileCoord gx_tile_from_grid() {
  GxTileCoord c;
  c.tile_m = gxg::stages_pair(kStages) ? blockIdx.x : blockIdx.y;
  c.tile_n = gxg::stages_pair(kStages) ? blockIdx.y : blockIdx.x;
  c.split = blockIdx.z;
  c.cta_linear = (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
  c.n_ctas = (int)(gridDim.x * gridDim.y * gridDim.z);
  return c;
}
template <int kStages>
__device__ __forceinline__ void gx_gemm_body(const CUtensorMap& map_a, const CUtensorMap& map_b,
                                             const gxg::GemmArgs& args, const GxTileCoord& coord GX_GEMM_EPI_PARAM_REF) {
  using namespace gxg;
  extern __shared__ uint8_t gg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gg_smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr bool kPair = stages_pair(kStages), kDeep = stages_deep(kStages);
  constexpr uint32_t kBytesB = kPair ? kStageBytesB / 2 : kStageBytesB;   // a pair's CTA holds half of the B tile
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kStageBytesA;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * (kStageBytesA + kBytesB));
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  // pair: the grid is (tile rows, tile columns) and the cluster (2, 1, 1) - two consecutive tile rows (the launch is
  // rejected as a cluster misconfiguration unless the pair lies along x); rank 0 (the leader) issues the MMAs of both
  const int m0 = coord.tile_m * BM, n0 = coord.tile_n * BN;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int cta_linear = coord.cta_linear;
  if (threadIdx.x == 0) {
    trace_stamp(args.trace, cta_linear, 0);
    if (args.trace) {   // slot 6: the SM this CTA runs on
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      args.trace[(size_t)cta_linear * 8 + 6] = smid;
    }
  }
  const int k_begin = coord.split * args.k_per_split;
  const int k_len = args.K - k_begin < args.k_per_split ? args.K - k_begin : args.k_per_split;   // last split may be short
  const int num_k_blocks = (k_len + BK - 1) / BK;

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
  if (warp == 1) {  // the allocating warp also owns the deallocation (pair: warp 1 of both CTAs, collectively)
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                   "n"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                   "n"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  // everything above overlapped the previous kernel's tail; from here on its results are needed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) trace_stamp(args.trace, cta_linear, 1);

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&empty_bar[s], ((kb / kStages) & 1) ^ 1);  // slot free (passes immediately the first round)
        const int k0 = k_begin + kb * BK;
        if (kPair) {
          // both CTAs' bytes are counted on the leader's barrier (the only one the MMA thread waits on); the peer's
          // copies may complete before the leader has armed the phase: the pending arrival keeps it open
          if (leader) mbar_expect_tx(&full_bar[s], 2 * (kStageBytesA + kBytesB));
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
          const int nh = n0 + (int)cta_rank * (BN / 2);          // this CTA's half of the B tile
          if (!args.a_mn) {
            tma_load_2d_pair(smem_a + s * kStageBytesA, &map_a, k0, m0, bar);
          } else {
            tma_load_2d_pair(smem_a + s * kStageBytesA, &map_a, m0, k0, bar);
            tma_load_2d_pair(smem_a + s * kStageBytesA + 8192, &map_a, m0 + 64, k0, bar);
          }
          if (!args.b_mn) tma_load_2d_pair(smem_b + s * kBytesB, &map_b, k0, nh, bar);   // 64-row box (host: gxg_make_maps)
          else tma_load_2d_pair(smem_b + s * kBytesB, &map_b, nh, k0, bar);
          continue;
        }
        mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
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
    // ===== MMA issuer (pair: the leader's only; the peer's warp 1 just owns its half of the TMEM allocation) =====
    for (int kb = 0; kb < (leader ? num_k_blocks : 0); ++kb) {
      const int s = kb % kStages;
      mbar_wait(&full_bar[s], (kb / kStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        if (kb == 0) trace_stamp(args.trace, cta_linear, 2);
        const uint32_t sa = smem_u32(smem_a + s * kStageBytesA), sb = smem_u32(smem_b + s * kBytesB);
        const uint64_t da = args.a_mn ? make_smem_desc_mn(sa) : make_smem_desc(sa);
        const uint64_t db = args.b_mn ? make_smem_desc_mn(sb) : make_smem_desc(sb);
        // one MMA consumes 16 K: K-major = 32 B inside the swizzle atom (+2 in the 16-byte address field),
        // MN-major = 16 rows of 128 B (+128)
        const uint32_t step_a = args.a_mn ? 128u : 2u, step_b = args.b_mn ? 128u : 2u;
        const uint32_t idesc = (kPair ? kInstrDescPair : kInstrDescBase) | (args.a_mn ? (1u << 15) : 0u) | (args.b_mn ? (1u << 16) : 0u);
        if (kPair) {
          // descriptors are shared-memory OFFSETS: the same stage of both CTAs; rows 128-255 of D land in the peer's TMEM
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16_pair(tmem_base, da + step_a * k, db + step_b * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[s]);                       // frees the stage in both CTAs
          if (kb == num_k_blocks - 1) umma_commit_pair(tmem_full_bar);  // both epilogues may start
        } else {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16(tmem_base, da + step_a * k, db + step_b * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);                       // smem slot reusable once these MMAs have read it
          if (kb == num_k_blocks - 1) umma_commit(tmem_full_bar);  // accumulator complete
        }
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes 32*(w%4) .. 32*(w%4)+31 =====
    if (args.zero_n) {   // after griddepcontrol.wait: nobody adds into the region before this grid has completed
      const unsigned n_ctas = (unsigned)coord.n_ctas, per_cta = kThreads - 64;
      for (unsigned i = (unsigned)cta_linear * per_cta + (threadIdx.x - 64); i < args.zero_n; i += n_ctas * per_cta)
        args.zero_ptr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int quad = warp & 3;
    const int c_begin = ((warp - 2) >> 2) * (BN / 2);
    const int m_base = m0 + quad * 32;
    float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * 1024;   // 4 KB of the (then idle) operand ring per warp
    auto load_acc = [&](int c, float (&acc)[32]) {
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
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
    };
#ifdef GX_GEMM_FUSED
    // The memory operands of the fused expressions do not depend on this kernel's main loop: their loads are
    // issued while the tensor core is still working (first chunk: before the accumulator is even waited for;
    // second chunk: as well when registers allow - one CTA per SM -, else while the first chunk is evaluated).
    static_assert(BN / 2 == 64, "two 32-column chunks per epilogue warp");
#ifdef GX_GEMM_MIN_BLOCKS
    constexpr bool kRoomy = GX_GEMM_MIN_BLOCKS == 1;   // one CTA per SM: registers for both chunks' operands at once
#else
    constexpr bool kRoomy = kDeep;
#endif
    GxEpiState epi_state;
    gx_epi_begin(epi_state);
    GxEpiPre pre0, pre1;
    gx_epi_prefetch(epi, pre0, m_base, n0 + c_begin, args.M, args.ldc, lane);
    if (kRoomy) gx_epi_prefetch(epi, pre1, m_base, n0 + c_begin + 32, args.M, args.ldc, lane);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) trace_stamp(args.trace, cta_linear, 3);
    {
      float acc[32];
      load_acc(c_begin, acc);
      if (!kRoomy) gx_epi_prefetch(epi, pre1, m_base, n0 + c_begin + 32, args.M, args.ldc, lane);
      gx_epi_chunk(epi, epi_state, pre0, tile, m_base, n0 + c_begin, args.M, args.ldc, acc, lane);   // N parameter = row stride
    }
    {
      float acc[32];
      load_acc(c_begin + 32, acc);
      gx_epi_chunk(epi, epi_state, pre1, tile, m_base, n0 + c_begin + 32, args.M, args.ldc, acc, lane);
    }
    gx_epi_finish(epi, epi_state, lane);
#else
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 64) trace_stamp(args.trace, cta_linear, 3);
#pragma unroll 1
    for (int c = c_begin; c < c_begin + BN / 2; c += 32) {
      float acc[32];
      load_acc(c, acc);
      if (args.epilogue == EPI_STORE_F32)
        store_rows_f32(tile, acc, args.c_f32, args.ldc, m_base, args.M, n0 + c, lane);
      else if (args.epilogue == EPI_STORE_BF16)
        store_rows_bf16(tile, acc, args.c_bf16, args.ldc, m_base, args.M, n0 + c, lane);
      else
        red_rows_f32(tile, acc, args.c_f32, args.ldc, m_base, args.M, n0 + c, lane);
    }
#endif
    tc_fence_before();
    if (threadIdx.x == 64) trace_stamp(args.trace, cta_linear, 4);
  }
  __syncthreads();
  if (threadIdx.x == 64) trace_stamp(args.trace, cta_linear, 5);
  if (kPair) cluster_sync_all();   // neither CTA may leave (or free its TMEM) while the other still reads / writes it
  if (warp == 1) {
    tc_fence_after();
    if (kPair)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

#ifndef GX_NVRTC
// ---------------------------------------------------------------------------------------------------------------
// Persistent variant for grids of several waves: one CTA per SM walks over output tiles; the accumulator is
// double-buffered in TMEM (2 x 128 columns) so that the epilogue of tile i (warps 2-9, through their own 32 KB of
// staging tiles) overlaps the main loop of tile i+1, and the 6-stage operand ring keeps loading across tile
// boundaries.  Plain store epilogues only (fp32 / bf16).
// ---------------------------------------------------------------------------------------------------------------
namespace gxg {
constexpr uint32_t kRingBytesPersistent = 6 * (kStageBytesA + kStageBytesB);     // 192 KB: 6 x 32 KB or 4 x 48 KB
constexpr uint32_t kSmemPersistent = kRingBytesPersistent + 8 * 4096 + 1024 + 256;
// Tile order of the persistent CTAs: groups of kGroupM tile rows, column by column inside a group, so that the
// ~148 tiles in flight form a compact block of the output (16 x ~9 tiles) and re-use each other's A and B
// panels in L2 instead of streaming a whole operand per wave.
constexpr int kGroupM = 16;
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
  const int per_group = kGroupM * tiles_n;
  const int g = t / per_group, r = t - g * per_group;
  const int first = g * kGroupM;
  const int gsize = tiles_m - first < kGroupM ? tiles_m - first : kGroupM;
  tm = first + r % gsize;
  tn = r / gsize;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace gxg

// Work items of the persistent kernel.  With 256-wide tiles the tiles that do not fill a whole round of the grid
// are cut into two 128-wide halves: 4096^3 is 512 tiles = 3.46 rounds of 148 CTAs, i.e. 4 rounds; with the last 68
// tiles as 136 half tiles it is 3.5.  Item w < full: tile w, full width; else half (w - full) & 1 of tile
// full + (w - full) / 2.
struct GemmItem {
  int tile, half, width;
};
template <int BNP>
__device__ __forceinline__ GemmItem gemm_item(int w, int full) {
  if (BNP == 128 || w < full) return {w, 0, BNP};
  return {full + ((w - full) >> 1), (w - full) & 1, 128};
}

template <int BNP>   // tile width: 128, or 256 (A re-read half as often: 85 instead of 64 flop per operand byte)
__device__ __forceinline__ void gx_gemm_persistent_body(const CUtensorMap& map_a, const CUtensorMap& map_b,
                                                        const CUtensorMap& map_b_half, const gxg::GemmArgs& args) {
  using namespace gxg;
  constexpr uint32_t kStageBytesB = BNP * BK * 2;                        // shadows the 128-wide constant
  constexpr int kStages = kRingBytesPersistent / (kStageBytesA + kStageBytesB);   // 6 or 4
  constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((BNP >> 3) << 17) | ((BM >> 4) << 24);
  extern __shared__ uint8_t gg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gg_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kStageBytesA;
  float* staging = reinterpret_cast<float*>(smem + kStages * (kStageBytesA + kStageBytesB));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * (kStageBytesA + kStageBytesB) + 8 * 4096);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int tiles_m = (args.M + BM - 1) / BM, tiles_n = (args.N + BNP - 1) / BNP, n_tiles = tiles_m * tiles_n;
  const int num_k_blocks = (args.K + BK - 1) / BK;
  // tiles in whole rounds; the remainder is halved only if the halves still fit in one round (else nothing is won)
  const int rounds_full = (n_tiles / (int)gridDim.x) * (int)gridDim.x;
  const int full = (BNP == 128 || 2 * (n_tiles - rounds_full) > (int)gridDim.x) ? n_tiles : rounds_full;
  const int n_items = full + 2 * (n_tiles - full);

  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 8);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(2 * BNP)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer: the ring position runs on across tiles =====
    if (elect_one()) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const GemmItem item = gemm_item<BNP>(w, full);
        int tm, tn;
        tile_coords(item.tile, tiles_m, tiles_n, tm, tn);
        const int m0 = tm * BM, n0 = tn * BNP + item.half * 128;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&empty_bar[s], ((it / kStages) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], kStageBytesA + (uint32_t)item.width * BK * 2);
          const int k0 = kb * BK;
          if (!args.a_mn) {
            tma_load_2d(smem_a + s * kStageBytesA, &map_a, k0, m0, &full_bar[s]);
          } else {
            tma_load_2d(smem_a + s * kStageBytesA, &map_a, m0, k0, &full_bar[s]);
            tma_load_2d(smem_a + s * kStageBytesA + 8192, &map_a, m0 + 64, k0, &full_bar[s]);
          }
          if (!args.b_mn) {
            tma_load_2d(smem_b + s * kStageBytesB, item.width == BNP ? &map_b : &map_b_half, k0, n0, &full_bar[s]);
          } else {
            for (int q = 0; q < item.width / 64; ++q)
              tma_load_2d(smem_b + s * kStageBytesB + q * 8192, &map_b, n0 + 64 * q, k0, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator i & 1 =====
    uint32_t it = 0, i = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++i) {
      const int width = gemm_item<BNP>(w, full).width;
      const uint32_t buf = i & 1;
      mbar_wait(&tmem_empty_bar[buf], ((i >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
      tc_fence_after();
      for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&full_bar[s], (it / kStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem_a + s * kStageBytesA), sb = smem_u32(smem_b + s * kStageBytesB);
          const uint64_t da = args.a_mn ? make_smem_desc_mn(sa) : make_smem_desc(sa);
          const uint64_t db = args.b_mn ? make_smem_desc_mn(sb) : make_smem_desc(sb);
          const uint32_t step_a = args.a_mn ? 128u : 2u, step_b = args.b_mn ? 128u : 2u;
          const uint32_t idesc = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(width >> 3) << 17) | ((BM >> 4) << 24)) |
                                 (args.a_mn ? (1u << 15) : 0u) | (args.b_mn ? (1u << 16) : 0u);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_f16(tmem_base + buf * BNP, da + step_a * k, db + step_b * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full_bar[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue =====
    const int quad = warp & 3;
    float* tile = staging + (warp - 2) * 1024;
    uint32_t i = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++i) {
      const GemmItem item = gemm_item<BNP>(w, full);
      const uint32_t buf = i & 1;
      int tm, tn;
      tile_coords(item.tile, tiles_m, tiles_n, tm, tn);
      const int m_base = tm * BM + quad * 32, n0 = tn * BNP + item.half * 128;
      const int c_begin = ((warp - 2) >> 2) * (item.width / 2), c_end = c_begin + item.width / 2;
      mbar_wait(&tmem_full_bar[buf], (i >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 64) {
        float acc[2][32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + buf * BNP + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c + 32 * h);
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
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[h][j] = __uint_as_float(v[j]);
        }
        if (c + 64 >= c_end) {
          // the last of this warp's columns is in registers: hand the TMEM buffer back before the global stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col0 = n0 + c + 32 * h;
          if (args.epilogue == EPI_STORE_F32)
            store_rows_f32(tile, acc[h], args.c_f32, args.ldc, m_base, args.M, col0, lane);
          else
            store_rows_bf16(tile, acc[h], args.c_bf16, args.ldc, m_base, args.M, col0, lane);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BNP) : "memory");
  }
}

extern "C" __global__ void __launch_bounds__(gxg::kThreads, 1)
gx_gemm_tn_kernel_persistent(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                             const __grid_constant__ CUtensorMap map_b_half, const gxg::GemmArgs args) {
  gx_gemm_persistent_body<128>(map_a, map_b, map_b_half, args);
}
extern "C" __global__ void __launch_bounds__(gxg::kThreads, 1)
gx_gemm_tn_kernel_persistent_n256(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                  const __grid_constant__ CUtensorMap map_b_half, const gxg::GemmArgs args) {
  gx_gemm_persistent_body<256>(map_a, map_b, map_b_half, args);
}
#endif  // !GX_NVRTC

#ifndef GX_GEMM_STAGES
#define GX_GEMM_STAGES 3
#endif
#ifndef GX_GEMM_MIN_BLOCKS   /* resident CTAs per SM the register budget is sized for */
#define GX_GEMM_MIN_BLOCKS (gxg::stages_deep(GX_GEMM_STAGES) ? 1 : 2)
#endif
extern "C" __global__ void __launch_bounds__(gxg::kThreads, GX_GEMM_MIN_BLOCKS)
GX_GEMM_KERNEL_NAME(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const gxg::GemmArgs args GX_GEMM_EPI_PARAM) {
  gx_gemm_body<GX_GEMM_STAGES>(map_a, map_b, args, gx_tile_from_grid<GX_GEMM_STAGES>() GX_GEMM_EPI_ARG);
}
#ifndef GX_NVRTC
// the library's second instance: deep ring, one CTA per SM
extern "C" __global__ void __launch_bounds__(gxg::kThreads, 1)
gx_gemm_tn_kernel_deep(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const gxg::GemmArgs args) {
  gx_gemm_body<gxg::kStagesDeep>(map_a, map_b, args, gx_tile_from_grid<gxg::kStagesDeep>());
}
// CTA-pair instances (grid (tile rows, tile columns, splits) in clusters of (2, 1, 1); map_b has 64-row boxes)
extern "C" __global__ void __launch_bounds__(gxg::kThreads, 2)
gx_gemm_tn_kernel_pair(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const gxg::GemmArgs args) {
  gx_gemm_body<gxg::kStagesPairShallow>(map_a, map_b, args, gx_tile_from_grid<gxg::kStagesPairShallow>());
}
extern "C" __global__ void __launch_bounds__(gxg::kThreads, 1)
gx_gemm_tn_kernel_pair_deep(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                            const gxg::GemmArgs args) {
  gx_gemm_body<gxg::kStagesPairDeep>(map_a, map_b, args, gx_tile_from_grid<gxg::kStagesPairDeep>());
}
// Two independent contractions in ONE launch (1-D grid: the CTAs of the first problem, then those of the second):
// the weight Jacobians of adjacent layers, g2^T a1 and g1^T x - each alone is half a wave of split-K tiles that pays
// its own pipeline fill, epilogue and hand-over; together they fill the GPU with two CTAs per SM and the epilogue of
// one CTA runs under the main loop of its neighbour.
struct GxDualArgs {
  gxg::GemmArgs args[2];
  int n_ctas0;                 // CTAs of problem 0
  int tiles_m[2], tiles_n[2];
};
extern "C" __global__ void __launch_bounds__(gxg::kThreads, 2)
gx_gemm_tn_kernel_dual(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_b0,
                       const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                       const __grid_constant__ GxDualArgs dual) {
  const int which = (int)blockIdx.x >= dual.n_ctas0 ? 1 : 0;
  const int local = (int)blockIdx.x - (which ? dual.n_ctas0 : 0);
  GxTileCoord c;
  c.tile_n = local % dual.tiles_n[which];
  c.tile_m = (local / dual.tiles_n[which]) % dual.tiles_m[which];
  c.split = local / (dual.tiles_n[which] * dual.tiles_m[which]);
  c.cta_linear = (int)blockIdx.x;
  c.n_ctas = (int)gridDim.x;
  if (which) gx_gemm_body<gxg::kStagesShallow>(map_a1, map_b1, dual.args[1], c);
  else gx_gemm_body<gxg::kStagesShallow>(map_a0, map_b0, dual.args[0], c);
}
#endif
