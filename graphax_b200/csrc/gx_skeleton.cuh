// gx_skeleton.cuh - persistent, TMA-staged kernel around a plan-specialised body (sm_100a).
//
// One thread evaluates one sample: the whole elimination plan (primal, elemental partials, every
// J[s,p] += J[s,v]*J[v,p] product, reference core.py:155-272) runs in registers as straight-line
// code.  HBM is touched exactly twice per sample: the inputs (4*n_in bytes) and the dense Jacobian
// tile (4*n_rows*n_cols bytes).  Both sides are staged through shared memory with 1-D bulk
// asynchronous copies (cp.async.bulk = the TMA engine, SASS UBLKCP) so that
//   * global traffic is fully coalesced although every thread owns a strided [tile] row,
//   * the loads of tile i+1 and the stores of tile i-1 overlap the arithmetic of tile i, and
//   * no issue slots are spent on address arithmetic / LDG / STG for the 60 words per sample.
// CTAs are persistent (grid = SMs x resident CTAs).  Every WARP is an independent pipeline over tiles
// of 32 samples with its own stages and mbarriers: the steady state has no CTA-wide barrier, so a
// warp that finishes its tile early never waits for a slower sibling.  A partial last tile, or
// misaligned caller buffers, take a guarded non-bulk path through the same body.
//
// The including translation unit (emitted by graphax_b200/codegen.py) defines:
//   GX_N_ARR, GX_N_IN, GX_N_ROWS, GX_N_COLS, GX_HAS_PRIMAL, GX_THREADS, GX_MIN_BLOCKS, GX_OUT_STAGES
//   GX_SPT                        samples per thread: 1 (scalar FFMA/FMUL/FADD) or 2 (both samples packed in
//                                 64-bit register pairs and driven through FFMA2/FMUL2/FADD2, sm_100's
//                                 f32x2 pipe: half the issue slots for the same arithmetic)
//   GX_KERNEL_NAME                unique symbol of this kernel (gx_jacve_kernel for NVRTC)
//   GX_FOREACH_ARRAY(F)           F(index, scalars_per_sample, prefix_offset) for every input array
//   gx_body<kPrecise>(x, jrow, prow)  __device__ __forceinline__ function: the straight-line plan; reads the
//                                 sample's inputs x[k], writes tile entries jrow[pos] and outputs prow[row]
#pragma once

#ifndef GX_NVRTC
#include <cuda_runtime.h>
#include <stdint.h>
#else
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#endif

#define GX_TILE (GX_N_ROWS * GX_N_COLS)
#ifndef GX_IN_STAGES
#define GX_IN_STAGES 2 /* 0: inputs never touch shared memory - every lane loads its own sample's scalars for the NEXT
                             tile straight into registers from a point late in the current tile's body, where few
                             values are live (a few dozen bytes per sample: the loads of a warp cover whole sectors
                             of every input array, and their latency hides behind the rest of the body);
                          2: bulk-copied (TMA) into two shared-memory stages per warp, one tile ahead;
                          1: one such stage (wide tiles), the next tile requested as soon as the current one has
                             been copied to registers */
#endif
#ifndef GX_CHECKED_FAST_MATH
#define GX_CHECKED_FAST_MATH 1
#endif
#ifndef GX_SPT
#define GX_SPT 1
#endif
#define GX_WT (32 * GX_SPT) /* samples per warp tile */
#if GX_SPT != 1 && GX_IN_STAGES == 0
#error "the packed two-sample kernels stage their inputs: GX_IN_STAGES must be 1 or 2"
#endif
#define GX_MAX_ARRAYS_K 32

struct GxLaunchArgs {
  const float* in[GX_MAX_ARRAYS_K];
  float* jac;
  float* primal;
  long long batch;
  long long n_bulk_tiles;  // leading tiles that are full and 16B-aligned on every buffer
};

namespace {
namespace gxk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GX_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra GX_DONE;\n"
      "bra GX_WAIT;\n"
      "GX_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (TMA load)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy, tracked by bulk async-groups (TMA store)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// One lane of the (converged) warp. Unlike `lane == 0`, elect.sync tells the compiler that exactly one
// thread runs the guarded region, so uniform-datapath instructions (UBLKCP) need no election loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// Programmatic dependent launch: let the next grid on the stream start its prologue while this one
// drains, and make this grid wait for its predecessor before it touches global memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One warp = one pipeline. A warp tile is GX_WT consecutive samples; every warp owns its stages and
// mbarriers, so the only synchronisation in the steady state is __syncwarp().
constexpr int kWarps = GX_THREADS / 32;
constexpr int kInStageFloats = GX_N_IN * GX_WT;
constexpr int kOutStageFloats = (GX_TILE + (GX_HAS_PRIMAL ? GX_N_ROWS : 0)) * GX_WT;
constexpr int kWarpFloats = GX_IN_STAGES * kInStageFloats + GX_OUT_STAGES * kOutStageFloats + 4;  // + 2 mbarriers
constexpr int kSmemBytes = kWarps * kWarpFloats * 4;

}  // namespace gxk
}  // namespace

// shared memory, per warp: [ in stage 0 | in stage 1 | out stage 0 (tile rows, then primal rows) | out stage 1 | 2 mbarriers ]
extern "C" __global__ void __launch_bounds__(GX_THREADS, GX_MIN_BLOCKS)
GX_KERNEL_NAME(const __grid_constant__ GxLaunchArgs args) {
  using namespace gxk;
  extern __shared__ __align__(128) float gx_smem[];
  const int lane = threadIdx.x & 31;
  // broadcast from lane 0 so that the compiler knows the warp index (and everything derived from it:
  // tile index, stage pointers, bulk-copy addresses) is warp-uniform and keeps it on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  float* in_stage = gx_smem + warp * kWarpFloats;
  float* out_stage = in_stage + GX_IN_STAGES * kInStageFloats;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + GX_OUT_STAGES * kOutStageFloats);

  const long long n_tiles = (args.batch + GX_WT - 1) / GX_WT;
  const long long n_bulk = args.n_bulk_tiles;
  const long long stride = (long long)gridDim.x * kWarps;

  pdl_launch_dependents();
#if GX_IN_STAGES > 0
  if (lane == 0) {
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
#endif
  pdl_wait();  // everything above overlaps the tail of the previous kernel on the stream

  auto issue_loads = [&](long long tile, int stage) {
    // the warp's elected lane programs the TMA engine: one contiguous chunk per input array
    mbar_expect_tx(&full_bar[stage], (uint32_t)(GX_N_IN * GX_WT * 4));
#define GX_ISSUE(a, size, off)                                                                                  \
  bulk_g2s(in_stage + stage * kInStageFloats + (off) * GX_WT, args.in[a] + tile * (long long)(GX_WT * (size)), \
           (uint32_t)((size) * GX_WT * 4), &full_bar[stage]);
    GX_FOREACH_ARRAY(GX_ISSUE)
#undef GX_ISSUE
  };

  long long tile = (long long)blockIdx.x * kWarps + warp;
#if GX_IN_STAGES > 0
  if (tile < n_bulk && elect_one()) issue_loads(tile, 0);
#elif GX_SPT == 1
  // every lane reads its own sample (lanes past the end of the batch compute on 1.0 and store nothing)
  auto load_inputs = [&](float (&dst)[GX_N_IN], long long t) {
    const long long sample = t * GX_WT + lane;
    const bool ok = sample < args.batch;
#define GX_RD_DIRECT(a, size, off)                   \
  _Pragma("unroll") for (int c = 0; c < (size); ++c) \
      dst[(off) + c] = ok ? __ldg(args.in[a] + sample * (size) + c) : 1.0f;
    GX_FOREACH_ARRAY(GX_RD_DIRECT)
#undef GX_RD_DIRECT
  };
  float x[GX_N_IN], xn[GX_N_IN];
  load_inputs(x, tile);
#endif

  for (int it = 0; tile < n_tiles; ++it, tile += stride) {
    const int stage = GX_IN_STAGES == 2 ? (it & 1) : 0;
    const uint32_t in_parity = GX_IN_STAGES == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
    const int ostage = it % GX_OUT_STAGES;
    const bool bulk = tile < n_bulk;
    const long long next = tile + stride;
    // prefetch the next tile: its stage was last read one iteration ago, before the __syncwarp below
    if (GX_IN_STAGES == 2 && next < n_bulk && elect_one()) issue_loads(next, stage ^ 1);

    float* tile_s = out_stage + ostage * kOutStageFloats;
#if GX_SPT == 1
#if GX_IN_STAGES == 0
    // x holds this tile's inputs (loaded during the previous tile, or before the loop); the body calls `prefetch`
    // once, late, to start the loads of the next tile into xn
    auto prefetch = [&]() { load_inputs(xn, next); };
    auto no_prefetch = []() {};
    bulk_wait_read<GX_OUT_STAGES - 1>();
    __syncwarp();
#else
    float x[GX_N_IN];
    auto prefetch = []() {};
    auto no_prefetch = []() {};
    const long long sample = tile * GX_WT + lane;
    if (bulk) {
      mbar_wait(&full_bar[stage], in_parity);
      const float* s = in_stage + stage * kInStageFloats;
#define GX_RD_BULK(a, size, off) \
  _Pragma("unroll") for (int c = 0; c < (size); ++c) x[(off) + c] = s[(off) * GX_WT + lane * (size) + c];
      GX_FOREACH_ARRAY(GX_RD_BULK)
#undef GX_RD_BULK
    } else {
#define GX_RD_DIRECT(a, size, off)                       \
  _Pragma("unroll") for (int c = 0; c < (size); ++c)     \
      x[(off) + c] = sample < args.batch ? __ldg(args.in[a] + sample * (size) + c) : 1.0f;
      GX_FOREACH_ARRAY(GX_RD_DIRECT)
#undef GX_RD_DIRECT
    }
    // the out stage about to be written was handed to the TMA engine GX_OUT_STAGES iterations ago.  Bulk async-groups
    // are per-thread state and elect.sync does not promise which lane issued (and committed) the stores, so every
    // lane waits: a lane without outstanding groups falls straight through.
    bulk_wait_read<GX_OUT_STAGES - 1>();
    __syncwarp();
    // single input stage: every lane has its inputs in registers, the stage can take the next tile already
    if (GX_IN_STAGES == 1 && next < n_bulk && elect_one()) issue_loads(next, 0);
#endif
#if GX_CHECKED_FAST_MATH
    // branch-free reciprocals / square roots; a lane whose operand left their exact range flags the tile, which
    // is then re-evaluated with the IEEE intrinsics (identical results wherever both are defined)
    if (__any_sync(0xffffffffu, gx_body<false>(x, tile_s + lane * GX_TILE, tile_s + GX_WT * GX_TILE + lane * GX_N_ROWS,
                                                prefetch)))
      gx_body<true>(x, tile_s + lane * GX_TILE, tile_s + GX_WT * GX_TILE + lane * GX_N_ROWS, no_prefetch);
#else
    gx_body<true>(x, tile_s + lane * GX_TILE, tile_s + GX_WT * GX_TILE + lane * GX_N_ROWS, prefetch);
#endif
#if GX_IN_STAGES == 0
#pragma unroll
    for (int k = 0; k < GX_N_IN; ++k) x[k] = xn[k];
#endif
#else
    // two samples per thread: lane and lane + 32 of the warp tile, packed as (.x, .y)
    float2 x[GX_N_IN];
    const long long sample = tile * GX_WT + lane;
    if (bulk) {
      mbar_wait(&full_bar[stage], in_parity);
      const float* s = in_stage + stage * kInStageFloats;
#define GX_RD_BULK(a, size, off)                                                  \
  _Pragma("unroll") for (int c = 0; c < (size); ++c)                              \
      x[(off) + c] = make_float2(s[(off) * GX_WT + lane * (size) + c],           \
                                 s[(off) * GX_WT + (lane + 32) * (size) + c]);
      GX_FOREACH_ARRAY(GX_RD_BULK)
#undef GX_RD_BULK
    } else {
#define GX_RD_DIRECT(a, size, off)                                                                          \
  _Pragma("unroll") for (int c = 0; c < (size); ++c)                                                        \
      x[(off) + c] = make_float2(sample < args.batch ? __ldg(args.in[a] + sample * (size) + c) : 1.0f,      \
                                 sample + 32 < args.batch ? __ldg(args.in[a] + (sample + 32) * (size) + c) : 1.0f);
      GX_FOREACH_ARRAY(GX_RD_DIRECT)
#undef GX_RD_DIRECT
    }
    bulk_wait_read<GX_OUT_STAGES - 1>();
    __syncwarp();
    gx_body(x, tile_s + lane * GX_TILE, tile_s + (lane + 32) * GX_TILE,
            tile_s + GX_WT * GX_TILE + lane * GX_N_ROWS, tile_s + GX_WT * GX_TILE + (lane + 32) * GX_N_ROWS);
#endif

    if (bulk) {
      fence_async_smem();  // generic-proxy writes -> visible to the async proxy
      __syncwarp();
      if (elect_one()) {
        bulk_s2g(args.jac + tile * (long long)(GX_WT * GX_TILE), tile_s, (uint32_t)(GX_WT * GX_TILE * 4));
#if GX_HAS_PRIMAL
        if (args.primal != nullptr)
          bulk_s2g(args.primal + tile * (long long)(GX_WT * GX_N_ROWS), tile_s + GX_WT * GX_TILE,
                   (uint32_t)(GX_WT * GX_N_ROWS * 4));
#endif
        bulk_commit();
      }
    } else {
      __syncwarp();
      const long long base = tile * GX_WT;
      const int valid = (int)(args.batch - base < GX_WT ? args.batch - base : GX_WT);
      for (int i = lane; i < valid * GX_TILE; i += 32) args.jac[base * GX_TILE + i] = tile_s[i];
#if GX_HAS_PRIMAL
      if (args.primal != nullptr)
        for (int i = lane; i < valid * GX_N_ROWS; i += 32)
          args.primal[base * GX_N_ROWS + i] = tile_s[GX_WT * GX_TILE + i];
#endif
      __syncwarp();  // the stage may be rewritten by the next iteration
    }
  }
  bulk_wait_all();  // shared memory must outlive the stores that read it (every lane: see bulk_wait_read above)
}

#ifndef GX_NVRTC
// ---- ahead-of-time registration: build() compiles one such TU per named plan into the library ----
struct GxAotKernel {
  unsigned long long key;
  int threads, smem_bytes, n_in, n_rows, n_cols, has_primal, samples_per_thread;
  const void* func;
};
extern "C" void gx_register_aot(const GxAotKernel* k);
namespace {
struct GxRegistrar {
  GxRegistrar() {
    static GxAotKernel k = {GX_KEY, GX_THREADS, gxk::kSmemBytes, GX_N_IN, GX_N_ROWS, GX_N_COLS, GX_HAS_PRIMAL,
                            GX_SPT, (const void*)GX_KERNEL_NAME};
    gx_register_aot(&k);
  }
} gx_registrar_instance;
}  // namespace
#endif
