/*
 * graphax_b200 - C ABI of the B200-native jacve engine.
 *
 * graphax is pure Python on JAX: its hot path (`graphax.jacve`,
 * /root/reference/src/graphax/core.py:27-93 -> vertex_elimination_jaxpr, core.py:475-581) has no
 * FFI of its own.  This header is the boundary a `jax.ffi` custom call (or any other host) binds
 * to; each entry point cites the reference code it stands in for.  INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions: plain C, no exceptions across the boundary; 0 = success, negative = error
 * (message through gx_last_error(), thread local); the caller owns every buffer; launches are
 * stream-ordered and asynchronous; a plan is immutable after creation and may be launched from
 * several host threads.  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with GX_ERR_CUDA.
 */
#ifndef GRAPHAX_B200_H
#define GRAPHAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GX_ABI_VERSION 2
#define GX_MAX_ARRAYS 32

enum gx_status {
  GX_OK = 0,
  GX_ERR_BAD_BLOB = -1,   /* malformed / inconsistent elimination plan */
  GX_ERR_BAD_ARG = -2,    /* null pointer, negative batch, misuse */
  GX_ERR_CUDA = -3,       /* CUDA runtime / driver error (no device, launch failure, ...) */
  GX_ERR_JIT = -4,        /* NVRTC unavailable or compilation failed */
  GX_ERR_UNSUPPORTED = -5
};

enum gx_kernel_kind {
  GX_KERNEL_INTERPRETER = 0, /* generic persistent plan interpreter (always available; plans too large for
                                shared memory run from global memory with a per-plan scratch area, so two
                                launches of such a plan must be ordered on one stream) */
  GX_KERNEL_AOT = 1,         /* plan-specialised kernel compiled into this library by build() */
  GX_KERNEL_JIT = 2          /* plan-specialised kernel compiled at run time with NVRTC */
};

typedef struct gx_plan gx_plan; /* opaque */

typedef struct gx_info {
  uint32_t abi_version;
  uint32_t n_in_arrays;   /* number of function arguments */
  uint32_t n_out_arrays;  /* number of function outputs */
  uint32_t n_in_scalars;  /* f32 values read per sample */
  uint32_t n_rows;        /* output scalars per sample = rows of the Jacobian tile */
  uint32_t n_cols;        /* scalars of the argnums inputs = columns of the Jacobian tile */
  uint32_t n_ops;         /* plan length */
  uint32_t n_slots;       /* live values the interpreter keeps per sample */
  uint32_t has_primal;    /* plan also evaluates the function outputs (has_aux, core.py:578-579) */
  uint32_t kernel_kind;   /* enum gx_kernel_kind currently attached */
  uint32_t threads;       /* samples per CTA of that kernel */
  uint32_t smem_bytes;    /* dynamic shared memory per CTA */
  uint32_t ctas_per_sm;   /* occupancy the persistent grid is sized with */
  uint32_t regs_per_thread;
  uint32_t in_sizes[GX_MAX_ARRAYS];
  uint32_t out_sizes[GX_MAX_ARRAYS];
  uint64_t plan_key;      /* FNV-1a 64 of the plan blob; names AOT/JIT kernels */
  uint64_t launches;      /* kernels launched through this plan so far */
} gx_info;

/* Replaces: the trace-time work of vertex_elimination_jaxpr (core.py:522-533) being redone on every
 * call.  `blob` is the serialised elimination plan (graphax_b200/plan.py documents the layout): the
 * flat program that `_build_graph` (core.py:314-412) + `_eliminate_vertex` (core.py:155-272) emit
 * for one (function, order, argnums).  Validates the blob, uploads it to `device`, and attaches the
 * plan-specialised kernel compiled into the library for this blob if there is one. */
int gx_plan_create(const void* blob, size_t nbytes, int device, gx_plan** out);
void gx_plan_destroy(gx_plan* plan);
int gx_plan_info(const gx_plan* plan, gx_info* out);

/* Attach a plan-specialised kernel compiled at run time: `cuda_source` is the CUDA translation unit
 * graphax_b200/codegen.py emits for this plan (skeleton + straight-line body).  Needs libnvrtc.  */
int gx_plan_specialize(gx_plan* plan, const char* cuda_source, size_t nbytes, const char* arch);
/* Same, with explicit launch shape: threads per CTA (multiple of 32), resident CTAs per SM the register
 * allocation is bounded for (__launch_bounds__), 1 or 2 output stages per warp, and 1 or 2 samples per
 * thread (2 = both samples packed into FFMA2/FMUL2/FADD2 f32x2 instructions). */
int gx_plan_specialize_opts(gx_plan* plan, const char* cuda_source, size_t nbytes, const char* arch, int threads,
                            int min_blocks, int out_stages, int samples_per_thread);
/* Force a kernel kind for subsequent launches (GX_KERNEL_INTERPRETER is always valid).  Launching is thread-safe;
 * gx_plan_specialize* and gx_plan_select_kernel reconfigure the plan and must not run concurrently with launches
 * of the same plan (configure once, then launch from as many threads / streams as needed).  Plans too large for
 * shared memory run on an interpreter with one private scratch area per plan: their launches are chained with an
 * event, i.e. two streams never run the same such plan at the same time. */
int gx_plan_select_kernel(gx_plan* plan, int kernel_kind);

/* Replaces: the run-time half of `jax.vmap(graphax.jacve(f, order, argnums))(*xs)`, i.e. every
 * jnp/lax op the reference emits per call (elemental partials primitives.py:150-205, sparse_mul /
 * sparse_add tensor.py:334-397, densification core.py:555-562), as ONE kernel launch.
 *   in_ptrs[i]  device, f32, [batch, in_sizes[i]] row-major
 *   jac_out     device, f32, [batch, n_rows, n_cols] row-major (the caller slices it into the
 *               per-output / per-argnum blocks of core.py:564-568)
 *   primal_out  device, f32, [batch, n_rows], may be NULL
 *   stream      cudaStream_t (NULL = legacy default stream)                                         */
int gx_jacve_launch(gx_plan* plan, int64_t batch, const void* const* in_ptrs, void* jac_out,
                    void* primal_out, void* stream);

/* Same call with HOST buffers (the reference's user passes host arrays to a jitted function):
 * host->device copies, the launch and device->host copies are chunked and pipelined over internal
 * streams and pinned staging buffers; returns after the results are in `jac_out`.
 * `pinned` != 0 promises that all host buffers are page-locked (copied without staging).            */
int gx_jacve_host(gx_plan* plan, int64_t batch, const void* const* in_ptrs, void* jac_out,
                  void* primal_out, int pinned);

/* ---- dense x dense contraction (scope row (f)1: tensor-valued vertices, config 5) -----------------
 * Replaces: `lax.dot_general` inside `_pure_dot_product_mul` (sparse/tensor.py:1052-1057) when both edge
 * Jacobians are genuinely dense matrices, e.g. the weight Jacobians of an MLP over a batch.
 *   C[M,N] (+)= A[M,K] * B[N,K]^T,  A and B bf16 row-major (contraction axis contiguous), fp32 accumulate
 * on the tcgen05 tensor cores.  C is fp32 (c_is_bf16 = 0) or bf16, row-major.  M, N multiples of 128,
 * K a multiple of 64 * k_splits.  k_splits > 1 accumulates into a caller-zeroed fp32 C with atomics.
 * gx_transpose_bf16 provides [rows, cols] -> [cols, rows] for operands stored the other way round.       */
int gx_gemm_bf16_tn(int M, int N, int K, const void* A, const void* B, void* C, int c_is_bf16, int k_splits,
                    void* stream);
/* General operand layouts: a_mn != 0 means A is stored [K, M] row-major (the M axis contiguous, "MN-major"),
 * likewise b_mn for B stored [K, N]; no transposed copies are needed for g^T x style contractions. */
int gx_gemm_bf16(int M, int N, int K, const void* A, int a_mn, const void* B, int b_mn, void* C, int c_is_bf16,
                 int k_splits, void* stream);
/* Any M, N, K: the operand tensor maps carry the true extents (out-of-range elements read as zero); lda / ldb /
 * ldc are row strides in elements (0 = dense; operand strides multiples of 8).  C must be tile padded: at least
 * ceil(M/128)*128 rows of ldc >= ceil(N/128)*128 elements (whole 128-column tile rows are written, rows >= M never). */
int gx_gemm_bf16_ld(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                    int ldc, int c_is_bf16, int k_splits, void* stream);
int gx_transpose_bf16(int rows, int cols, const void* in, void* out, void* stream);
/* fp32-accurate contraction for fp32 callers (the reference's parity predicate is rtol 1e-4 / atol 1e-5,
 * core.py:17-20, which bf16 operand rounding cannot meet): gx_split3_bf16 writes x = x0 + x1 + x2 as three bf16
 * arrays (residual below 2^-24 |x|), gx_gemm_bf16x3_ld adds sum_{i+j<=2} A_i B_j^T into a caller-zeroed fp32 C
 * (A3 / B3: the three terms of each operand, same layout rules as gx_gemm_bf16_ld).  Six tensor-core products. */
int gx_split3_bf16(const void* in, void* out0, void* out1, void* out2, long long n, void* stream);
int gx_gemm_bf16x3_ld(int M, int N, int K, const void* const* A3, int lda, int a_mn, const void* const* B3, int ldb,
                      int b_mn, void* C, int ldc, int k_splits, void* stream);
/* Two independent contractions in ONE launch (index 0 / 1 of every array argument; each as in gx_gemm_bf16_ld): the
 * weight Jacobians of adjacent layers, each alone half a wave of split-K tiles.  Falls back to two launches when the
 * pair does not fit (a problem large enough for the persistent kernel). */
int gx_gemm_bf16_dual_ld(const int* M, const int* N, const int* K, const void* const* A, const int* lda, const int* a_mn,
                         const void* const* B, const int* ldb, const int* b_mn, void* const* C, const int* ldc,
                         const int* c_is_bf16, const int* k_splits, void* stream);
const char* gx_gemm_last_error(void);
/* Single-tile contractions as CTA pairs (tcgen05 cta_group::2: a cluster of two CTAs on the SMs of one TPC computes a
 * 256 x 128 tile, each CTA loading half of the B tile)?  Process-wide; default off (or GX_GEMM_PAIR=1); returns the
 * previous setting.  Needs an even number of 128-row tiles, otherwise the call runs one CTA per tile. */
int gx_gemm_set_pair(int on);
/* Diagnostics: the next `max_launches` single-tile contractions write 8 x uint64 phase stamps per CTA (%globaltimer at
 * entry, after the dependency wait, first operand stage landed, accumulator complete, epilogue done, exit) into
 * buf[launch][max_ctas][8] (device memory); NULL switches it off (tools/gemm_trace.py). */
int gx_gemm_set_trace(void* buf, int max_launches, int max_ctas);
/* The next single-tile contraction launched by the calling thread also clears `bytes` bytes at `ptr` (device memory,
 * 16-byte aligned, a multiple of 16; not an output of that contraction): the accumulator arena that split-K
 * contractions and fused column sums of LATER launches add into (the reference's zero-initialised fill-in,
 * core.py:555-562) - saves a memset launch and the dependency edge it brings.  One-shot. */
int gx_gemm_zero_with_next_launch(void* ptr, long long bytes);

/* Fused elementwise kernels of the tensor-valued path are emitted per plan (graphax_b200/dense.py) and
 * compiled at run time: the elementwise half of `_pure_broadcast_mul` / `_sparse_add` (sparse/tensor.py:
 * 826-866, 1267-1411) and of the elemental partials (primitives.py:150-205) on [B, H] tensors.
 * gx_colsum_f32 is the contraction of a dense cotangent with a broadcast (bias) edge: out[c] += sum_r in[r,c]. */
typedef struct gx_jit_kernel gx_jit_kernel;
int gx_jit_compile(const char* cuda_source, const char* kernel_name, int device, gx_jit_kernel** out);
void gx_jit_destroy(gx_jit_kernel* kernel);
int gx_jit_launch(gx_jit_kernel* kernel, unsigned grid_x, unsigned grid_y, unsigned block_x, void** args, void* stream);
int gx_colsum_f32(const void* in, void* out, int rows, int cols, void* stream);
/* The contraction with the elementwise consumers of its result fused into the epilogue: `kernel` is
 * csrc/gx_gemm_kernel.cuh compiled by gx_jit_compile behind a generated `GxEpi` / `gx_epi_chunk` (dense.py) -
 * the reference's `_mixed_mul` followed by `_pure_broadcast_mul` / elemental partials (sparse/tensor.py:1062-1231,
 * 826-866) without the intermediate edge ever reaching HBM.  `epi` points to the kernel's GxEpi block; `stages` is
 * the operand-ring depth the kernel was compiled with (GX_GEMM_STAGES: 3, or 6 for grids of at most one CTA per SM;
 * 4 / 8 name the CTA-pair variant of the same two - cta_group::2, clusters of two tile rows, M a multiple of 256 -);
 * M, N, K, lda, ldb as in gx_gemm_bf16_ld, `ldn` the row stride of every [M, N] epilogue operand and output. */
int gx_gemm_fused_launch(gx_jit_kernel* kernel, int M, int N, int K, const void* A, int lda, int a_mn, const void* B,
                         int ldb, int b_mn, int ldn, const void* epi, int stages, void* stream);

const char* gx_last_error(void);
int gx_abi_version(void);
/* Number of plan-specialised kernels compiled into this library and the key of the i-th one. */
int gx_aot_count(void);
uint64_t gx_aot_key(int index);
/* FNV-1a 64 over a byte range: the key function, exported so hosts can name kernels the same way. */
uint64_t gx_hash_bytes(const void* data, size_t nbytes);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHAX_B200_H */
