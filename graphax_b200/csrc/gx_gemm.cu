// gx_gemm.cu - host side of the dense x dense contraction (tcgen05 + TMEM + TMA); the kernel itself is in
// gx_gemm_kernel.cuh, which is also compiled at run time in front of plan-specific fused epilogues.
//
// The reference's `_pure_dot_product_mul` (sparse/tensor.py:999-1059) contracts two dense edge
// Jacobians with `lax.dot_general`.  For the per-sample graphs of configs 1-4 those are <= 3x3 and live
// in registers (gx_skeleton.cuh); for tensor-valued vertices (config 5: weight Jacobians of an MLP over
// a batch, [B,H]^T [B,H'] products) they are real GEMMs and go here.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <string>

#include "../../include/graphax_b200.h"
#include "gx_gemm_kernel.cuh"

namespace {
using namespace gxg;

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
// `rows` x `k` are the TRUE extents (the TMA engine zero-fills whatever a box reaches beyond them), `ld` the
// row stride in elements (0 = k)
int make_map(CUtensorMap* map, const void* ptr, int rows, int k, int box_rows, int ld = 0) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld ? ld : k) * 2};
  cuuint32_t box[2] = {(cuuint32_t)64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

int sm_count() {
  static const int v = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  return v;
}

int gemm_tile_n() {
  static const int v = [] {
    const char* e = getenv("GX_GEMM_TILE_N");
    return e ? atoi(e) : 256;
  }();
  return v;
}

bool gemm_use_persistent() {
  static const bool v = [] {
    const char* e = getenv("GX_GEMM_PERSISTENT");
    return !(e && e[0] == '0');
  }();
  return v;
}

// CTA pairs (cta_group::2: a cluster of two CTAs computes one 256 x 128 tile, each loading half of the B tile) for
// the single-tile kernels.  Off by default: on the MLP-sized problems of config 5 the pairs move 25 % fewer operand
// bytes but are no faster (7.4 vs 7.1 us for 4096 x 1024 x 512, B200 r02c - those launches are bound by their
// fill / epilogue latency, not by operand traffic).  GX_GEMM_PAIR=1 or gx_gemm_set_pair(1) turns them on.
int g_gemm_pair = [] {
  const char* e = getenv("GX_GEMM_PAIR");
  return (e && e[0] == '1') ? 1 : 0;
}();
bool gemm_use_pair() { return g_gemm_pair != 0; }

// phase stamps (GemmArgs::trace): a device buffer handed out launch by launch, `g_trace_ctas` CTAs x 8 words each
unsigned long long* g_trace = nullptr;
int g_trace_launches = 0, g_trace_ctas = 0, g_trace_next = 0;

// one-shot request of this host thread: the next single-tile launch clears [g_zero_ptr, +16 * g_zero_n)
thread_local float4* g_zero_ptr = nullptr;
thread_local unsigned g_zero_n = 0;

bool gemm_use_pdl() {
  static const bool v = [] {
    const char* e = getenv("GX_PDL");
    return !(e && e[0] == '0');
  }();
  return v;
}

}  // namespace

// tensor maps of both operands (shared with the run-time compiled kernels launched from gx_abi.cu)
// `pair`: the CTA-pair kernels load the B tile in two halves of 64 N-rows, one per CTA
int gxg_make_maps(CUtensorMap* map_a, CUtensorMap* map_b, int M, int N, int K, const void* A, int a_mn, const void* B,
                  int b_mn, int lda, int ldb, int pair) {
  const int ra = a_mn ? make_map(map_a, A, K, M, BK, lda) : make_map(map_a, A, M, K, BM, lda);
  const int rb = b_mn ? make_map(map_b, B, K, N, BK, ldb) : make_map(map_b, B, N, K, pair ? BN / 2 : BN, ldb);
  return (ra || rb) ? -1 : 0;
}

int gxg_pair_enabled(void) { return gemm_use_pair() ? 1 : 0; }

void gxg_take_zero(float4** ptr, unsigned* n16) {
  *ptr = g_zero_ptr;
  *n16 = g_zero_n;
  g_zero_ptr = nullptr;
  g_zero_n = 0;
}

unsigned long long* gxg_next_trace(int n_ctas) {
  if (!g_trace || g_trace_next >= g_trace_launches || n_ctas > g_trace_ctas) return nullptr;
  return g_trace + (size_t)(g_trace_next++) * g_trace_ctas * 8;
}

extern "C" {

const char* gx_gemm_last_error(void) { return g_gemm_error.c_str(); }

// Diagnostics: the next `max_launches` single-tile contractions (plain and fused) each write 8 x u64 phase stamps per
// CTA (%globaltimer: entry, dependency wait passed, first stage landed, accumulator complete, epilogue done, exit)
// into buf[launch][max_ctas][8]; launches with more CTAs than max_ctas are skipped.  NULL switches it off.
int gx_gemm_set_trace(void* buf, int max_launches, int max_ctas) {
  g_trace = static_cast<unsigned long long*>(buf);
  g_trace_launches = buf ? max_launches : 0;
  g_trace_ctas = max_ctas;
  g_trace_next = 0;
  return GX_OK;
}

// The next single-tile contraction launched by this thread (gx_gemm_bf16*, gx_gemm_fused_launch) also clears `bytes`
// bytes at `ptr` (16-byte aligned, a multiple of 16): its epilogue warps do it while the main loop runs, after the
// launch's dependency wait, so whatever the NEXT launch on the stream adds into the region finds it zeroed.  The
// region must not be an output of that contraction itself.  One-shot; a launch that cannot host it (persistent
// kernels) clears the region with a memset on its stream first.
int gx_gemm_zero_with_next_launch(void* ptr, long long bytes) {
  if (!ptr || bytes <= 0 || bytes % 16 || (reinterpret_cast<uintptr_t>(ptr) & 15) || bytes / 16 > 0xffffffffLL) {
    g_gemm_error = "gx_gemm_zero_with_next_launch: region must be 16-byte aligned and a multiple of 16 bytes";
    return GX_ERR_BAD_ARG;
  }
  g_zero_ptr = static_cast<float4*>(ptr);
  g_zero_n = (unsigned)(bytes / 16);
  return GX_OK;
}

// Use CTA pairs (cta_group::2) for single-tile contractions whose tile rows pair up?  Returns the previous setting.
int gx_gemm_set_pair(int on) {
  const int before = g_gemm_pair;
  g_gemm_pair = on ? 1 : 0;
  return before;
}

// C[M,N] (+)= A[M,K] * B[N,K]^T; A, B bf16 row-major device pointers; C fp32 (c_bf16 == 0) or bf16 row-major.
// k_splits > 1 accumulates with fp32 atomics into C, which the caller must have zeroed (fp32 only).
int gx_gemm_bf16_tn(int M, int N, int K, const void* A, const void* B, void* C, int c_is_bf16, int k_splits,
                    void* stream) {
  return gx_gemm_bf16(M, N, K, A, 0, B, 0, C, c_is_bf16, k_splits, stream);
}

// General form: a_mn != 0 means A is stored [K, M] row-major (M contiguous), likewise b_mn for B = [K, N].
int gx_gemm_bf16(int M, int N, int K, const void* A, int a_mn, const void* B, int b_mn, void* C, int c_is_bf16,
                 int k_splits, void* stream) {
  if (M > 0 && N > 0 && K > 0 && (M % BM || N % BN || K % BK || (k_splits >= 1 && (K / BK) % k_splits))) {
    g_gemm_error = "gx_gemm_bf16_tn: M, N must be multiples of 128, K of 64 x k_splits (gx_gemm_bf16_ld takes any size)";
    return GX_ERR_UNSUPPORTED;
  }
  return gx_gemm_bf16_ld(M, N, K, A, 0, a_mn, B, 0, b_mn, C, 0, c_is_bf16, k_splits, stream);
}

// True extents M, N, K of any size; lda / ldb / ldc are row strides in elements (0 = dense).  Operand rows must
// start 16-byte aligned (strides multiples of 8 elements); C must be allocated "tile padded": at least
// ceil(M/128)*128 rows of ldc >= ceil(N/128)*128 elements, because the epilogue writes whole 128-column tile rows
// (rows beyond M are never written).  Out-of-range operand elements read as zero (TMA fill).
static int gemm_launch(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                       int ldc, int c_is_bf16, int k_splits, bool accumulate, void* stream);

int gx_gemm_bf16_ld(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                    int ldc, int c_is_bf16, int k_splits, void* stream) {
  return gemm_launch(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, c_is_bf16, k_splits, false, stream);
}

// fp32-accurate contraction of fp32 matrices on the bf16 tensor cores.  Every operand is split into three bf16
// terms, x = x0 + x1 + x2 with x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1) (gx_split3_bf16: 8 + 8 + 8
// significant bits, the residual is below 2^-24 |x|), and C += sum over i + j <= 2 of A_i B_j^T: six tensor-core
// products whose bf16 x bf16 terms are exact in the fp32 accumulator; the dropped terms are below 2^-24 |a||b|.
// C is fp32, caller-zeroed (all six launches add with fp32 reductions); operands as in gx_gemm_bf16_ld.
int gx_gemm_bf16x3_ld(int M, int N, int K, const void* const* A3, int lda, int a_mn, const void* const* B3, int ldb,
                      int b_mn, void* C, int ldc, int k_splits, void* stream) {
  if (!A3 || !B3) {
    g_gemm_error = "gx_gemm_bf16x3_ld: bad argument";
    return GX_ERR_BAD_ARG;
  }
  static const int terms[6][2] = {{2, 0}, {1, 1}, {0, 2}, {1, 0}, {0, 1}, {0, 0}};   // small contributions first
  for (int t = 0; t < 6; ++t) {
    const int rc = gemm_launch(M, N, K, A3[terms[t][0]], lda, a_mn, B3[terms[t][1]], ldb, b_mn, C, ldc, 0,
                               k_splits < 1 ? 1 : k_splits, true, stream);
    if (rc != GX_OK) return rc;
  }
  return GX_OK;
}

namespace {
__global__ void gx_split3_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ o0,
                                      __nv_bfloat16* __restrict__ o1, __nv_bfloat16* __restrict__ o2, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = in[i];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h0);              // exact: both share the leading bits
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(h1);             // exact
    o0[i] = h0;
    o1[i] = h1;
    o2[i] = __float2bfloat16_rn(r2);
  }
}
}  // namespace

// x (fp32, n elements) -> three bf16 arrays with x0 + x1 + x2 = x up to 2^-24 |x| (see gx_gemm_bf16x3_ld)
int gx_split3_bf16(const void* in, void* out0, void* out1, void* out2, long long n, void* stream) {
  if (!in || !out0 || !out1 || !out2 || n <= 0) {
    g_gemm_error = "gx_split3_bf16: bad argument";
    return GX_ERR_BAD_ARG;
  }
  const int threads = 256;
  const long long want = (n + threads - 1) / threads;
  const int blocks = (int)(want < 148LL * 16 ? want : 148LL * 16);
  gx_split3_bf16_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(in), static_cast<__nv_bfloat16*>(out0), static_cast<__nv_bfloat16*>(out1),
      static_cast<__nv_bfloat16*>(out2), n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_gemm_error = std::string("gx_split3_bf16 launch: ") + cudaGetErrorString(e);
    return GX_ERR_CUDA;
  }
  return GX_OK;
}

static int gemm_launch(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                       int ldc, int c_is_bf16, int k_splits, bool accumulate, void* stream) {
  float4* zero_ptr = nullptr;     // a pending gx_gemm_zero_with_next_launch belongs to THIS launch, also if it fails
  unsigned zero_n = 0;
  gxg_take_zero(&zero_ptr, &zero_n);
  if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C) {
    g_gemm_error = "gx_gemm_bf16_tn: bad argument";
    return GX_ERR_BAD_ARG;
  }
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  if (ldc == 0) ldc = N;
  if (k_splits < 1 || ((k_splits > 1 || accumulate) && c_is_bf16) || ldc % 4 || ldc < (N % BN ? tiles_n * BN : N)) {
    g_gemm_error = "gx_gemm_bf16_tn: split-K needs an fp32 C; ldc must be a multiple of 4 and cover whole 128-column tiles";
    return GX_ERR_UNSUPPORTED;
  }
  const int k_blocks = (K + BK - 1) / BK;
  const int k_per_split = ((k_blocks + k_splits - 1) / k_splits) * BK;
  if ((long long)k_per_split * (k_splits - 1) >= K && k_splits > 1) {
    g_gemm_error = "gx_gemm_bf16_tn: more K splits than 64-wide K blocks";
    return GX_ERR_UNSUPPORTED;
  }
  // several waves of tiles: persistent CTAs (one per SM) with a double-buffered accumulator; otherwise CTA pairs
  // when they are switched on and the tile rows pair up
  const long long tiles = (long long)tiles_n * tiles_m;
  const bool persistent = k_splits == 1 && !accumulate && tiles > 2LL * sm_count() && gemm_use_persistent();
  const bool pair = !persistent && tiles_m % 2 == 0 && gemm_use_pair();
  CUtensorMap map_a, map_b;
  if (gxg_make_maps(&map_a, &map_b, M, N, K, A, a_mn, B, b_mn, lda, ldb, pair ? 1 : 0)) {
    g_gemm_error = "gx_gemm_bf16_tn: cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)";
    return GX_ERR_CUDA;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_bytes(kStagesShallow));
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel_deep, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_bytes(kStagesDeep));
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel_pair, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_bytes(kStagesPairShallow));
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel_pair_deep, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem_bytes(kStagesPairDeep));
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemPersistent);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel_persistent_n256, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemPersistent);
  });
  if (attr_err != cudaSuccess) {
    g_gemm_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(attr_err);
    return GX_ERR_CUDA;
  }
  GemmArgs args;
  args.c_f32 = c_is_bf16 ? nullptr : static_cast<float*>(C);
  args.c_bf16 = c_is_bf16 ? static_cast<uint16_t*>(C) : nullptr;
  args.M = M;
  args.N = N;
  args.K = K;
  args.ldc = ldc;
  args.epilogue = (k_splits > 1 || accumulate) ? EPI_ATOMIC_F32 : (c_is_bf16 ? EPI_STORE_BF16 : EPI_STORE_F32);
  args.k_per_split = k_per_split;
  args.a_mn = a_mn ? 1 : 0;
  args.b_mn = b_mn ? 1 : 0;
  args.trace = persistent ? nullptr : gxg_next_trace(tiles_n * tiles_m * k_splits);
  args.zero_ptr = zero_ptr;
  args.zero_n = zero_n;
  if (persistent && args.zero_n) {
    cudaMemsetAsync(args.zero_ptr, 0, (size_t)args.zero_n * 16, static_cast<cudaStream_t>(stream));
    args.zero_ptr = nullptr;
    args.zero_n = 0;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(tiles_n, tiles_m, k_splits);
  cfg.blockDim = dim3(kThreads);
  // a grid that fits the GPU once gets the deep ring (twice the bytes in flight per SM)
  const bool deep = (long long)tiles_n * tiles_m * k_splits <= sm_count();
  const int stages = pair ? (deep ? kStagesPairDeep : kStagesPairShallow) : (deep ? kStagesDeep : kStagesShallow);
  cfg.dynamicSmemBytes = smem_bytes(stages);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (gemm_use_pdl()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the predecessor's tail
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  if (pair) {   // two consecutive tile rows form a cluster = one 256 x 128 MMA tile; the pair must lie along x
    cfg.gridDim = dim3(tiles_m, tiles_n, k_splits);
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  CUtensorMap map_b_half = map_b;              // 128-row boxes of B: the half-width items of the persistent kernel
  void* params[] = {&map_a, &map_b, &args};
  void* params_persistent[] = {&map_a, &map_b, &map_b_half, &args};
  const void* fn = pair ? (deep ? (const void*)gx_gemm_tn_kernel_pair_deep : (const void*)gx_gemm_tn_kernel_pair)
                        : (deep ? (const void*)gx_gemm_tn_kernel_deep : (const void*)gx_gemm_tn_kernel);
  if (persistent) {
    cfg.gridDim = dim3(sm_count());
    cfg.dynamicSmemBytes = kSmemPersistent;
    fn = (const void*)gx_gemm_tn_kernel_persistent;
    if (N % 256 == 0 && tiles / 2 > 2LL * sm_count() && gemm_tile_n() == 256) {   // 128 x 256 tiles
      fn = (const void*)gx_gemm_tn_kernel_persistent_n256;
      if (!b_mn && make_map(&map_b, B, N, K, 256, ldb)) {      // K-major B: one 256-row box per stage
        g_gemm_error = "gx_gemm_bf16_tn: cuTensorMapEncodeTiled failed (256-row box)";
        return GX_ERR_CUDA;
      }
    }
  }
  cudaError_t e = cudaLaunchKernelExC(&cfg, fn, persistent ? params_persistent : params);
  if (e != cudaSuccess) {
    g_gemm_error = std::string("gx_gemm_tn_kernel launch: ") + cudaGetErrorString(e);
    return GX_ERR_CUDA;
  }
  return GX_OK;
}

// Two independent contractions in one launch (gx_gemm_tn_kernel_dual): both as in gx_gemm_bf16_ld (index 0 / 1 of
// every array argument).  Falls back to two launches when the pair does not fit the dual kernel (a problem large
// enough for the persistent kernel, more than 65535 CTAs).
int gx_gemm_bf16_dual_ld(const int* M, const int* N, const int* K, const void* const* A, const int* lda, const int* a_mn,
                         const void* const* B, const int* ldb, const int* b_mn, void* const* C, const int* ldc_in,
                         const int* c_is_bf16, const int* k_splits, void* stream) {
  if (!M || !N || !K || !A || !lda || !a_mn || !B || !ldb || !b_mn || !C || !ldc_in || !c_is_bf16 || !k_splits) {
    g_gemm_error = "gx_gemm_bf16_dual_ld: NULL argument";
    return GX_ERR_BAD_ARG;
  }
  CUtensorMap maps[4];
  GxDualArgs dual;
  memset(&dual, 0, sizeof dual);
  long long ctas[2];
  bool ok = true;
  for (int i = 0; i < 2 && ok; ++i) {
    const int tiles_m = (M[i] + BM - 1) / BM, tiles_n = (N[i] + BN - 1) / BN;
    const int ldc = ldc_in[i] ? ldc_in[i] : N[i];
    const int k_blocks = (K[i] + BK - 1) / BK;
    const int splits = k_splits[i];
    ok = M[i] > 0 && N[i] > 0 && K[i] > 0 && A[i] && B[i] && C[i] && splits >= 1 && !(splits > 1 && c_is_bf16[i]) &&
         ldc % 4 == 0 && ldc >= (N[i] % BN ? tiles_n * BN : N[i]);
    if (!ok) break;
    const int k_per_split = ((k_blocks + splits - 1) / splits) * BK;
    ok = !((long long)k_per_split * (splits - 1) >= K[i] && splits > 1) &&
         !(splits == 1 && (long long)tiles_m * tiles_n > 2LL * sm_count() && gemm_use_persistent());
    if (!ok) break;
    if (gxg_make_maps(&maps[2 * i], &maps[2 * i + 1], M[i], N[i], K[i], A[i], a_mn[i], B[i], b_mn[i], lda[i], ldb[i], 0)) {
      g_gemm_error = "gx_gemm_bf16_dual_ld: cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)";
      return GX_ERR_CUDA;
    }
    GemmArgs& a = dual.args[i];
    a.c_f32 = c_is_bf16[i] ? nullptr : static_cast<float*>(C[i]);
    a.c_bf16 = c_is_bf16[i] ? static_cast<uint16_t*>(C[i]) : nullptr;
    a.M = M[i], a.N = N[i], a.K = K[i], a.ldc = ldc;
    a.epilogue = splits > 1 ? EPI_ATOMIC_F32 : (c_is_bf16[i] ? EPI_STORE_BF16 : EPI_STORE_F32);
    a.k_per_split = k_per_split;
    a.a_mn = a_mn[i] ? 1 : 0, a.b_mn = b_mn[i] ? 1 : 0;
    dual.tiles_m[i] = tiles_m, dual.tiles_n[i] = tiles_n;
    ctas[i] = (long long)tiles_m * tiles_n * splits;
  }
  if (!ok || ctas[0] + ctas[1] > 65535 || g_zero_n) {   // not a pair for the dual kernel: one after the other
    for (int i = 0; i < 2; ++i) {
      const int rc = gx_gemm_bf16_ld(M[i], N[i], K[i], A[i], lda[i], a_mn[i], B[i], ldb[i], b_mn[i], C[i], ldc_in[i],
                                     c_is_bf16[i], k_splits[i], stream);
      if (rc != GX_OK) return rc;
    }
    return GX_OK;
  }
  dual.n_ctas0 = (int)ctas[0];
  dual.args[0].trace = gxg_next_trace((int)(ctas[0] + ctas[1]));
  dual.args[1].trace = dual.args[0].trace;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(gx_gemm_tn_kernel_dual, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_bytes(kStagesShallow));
  });
  if (attr_err != cudaSuccess) {
    g_gemm_error = std::string("cudaFuncSetAttribute(dual): ") + cudaGetErrorString(attr_err);
    return GX_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)(ctas[0] + ctas[1]));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes(kStagesShallow);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = gemm_use_pdl() ? 1 : 0;
  void* params[] = {&maps[0], &maps[1], &maps[2], &maps[3], &dual};
  cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)gx_gemm_tn_kernel_dual, params);
  if (e != cudaSuccess) {
    g_gemm_error = std::string("gx_gemm_tn_kernel_dual launch: ") + cudaGetErrorString(e);
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
