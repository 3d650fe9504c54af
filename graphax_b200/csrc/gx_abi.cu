// gx_abi.cu - C ABI of the graphax_b200 engine (include/graphax_b200.h).
//
//  * plan blob validation + upload (the flat program emitted by graphax_b200/plan.py)
//  * the generic persistent interpreter kernel: walks the plan with the batch on the grid, live edge
//    values in shared memory (always available, any elimination order)
//  * dispatch to plan-specialised kernels: ahead-of-time ones registered by generated TUs
//    (gx_skeleton.cuh) and run-time ones compiled with NVRTC and loaded through the driver API
//  * gx_jacve_host: chunked, double-buffered H2D -> kernel -> D2H pipeline for host buffers
//
// No CPU fallback: every compute entry point needs a CUDA device and says so when there is none.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/graphax_b200.h"
#include "gx_ops.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_error;

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}
#define GX_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(GX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__) +           \
                                   " (graphax_b200 has no CPU fallback; a CUDA device is required)"); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// AOT registry (filled by static registrars in generated translation units)
// ------------------------------------------------------------------------------------------------
struct GxAotKernel {
  unsigned long long key;
  int threads, smem_bytes, n_in, n_rows, n_cols, has_primal, samples_per_thread;
  const void* func;
};
std::vector<GxAotKernel>& aot_registry() {
  static std::vector<GxAotKernel> r;
  return r;
}
}  // namespace

extern "C" void gx_register_aot(const GxAotKernel* k) { aot_registry().push_back(*k); }

// launch-argument block shared with gx_skeleton.cuh (must stay layout-identical)
struct GxLaunchArgs {
  const float* in[GX_MAX_ARRAYS];
  float* jac;
  float* primal;
  long long batch;
  long long n_bulk_tiles;
};

// ------------------------------------------------------------------------------------------------
// generic interpreter kernel
// ------------------------------------------------------------------------------------------------
struct GxInterpArgs {
  const float* in[GX_MAX_ARRAYS];
  float* jac;
  float* primal;
  long long batch;
  const gx_op* ops;      // device
  const float* consts;   // device
  int n_ops, n_consts, n_slots, n_in, n_rows, n_cols, n_arr, has_primal;
  int in_size[GX_MAX_ARRAYS];
  int in_off[GX_MAX_ARRAYS];
  float* scratch;        // large-plan interpreter: live values, [grid][n_slots][T]
  const int* in_map;     // large-plan interpreter: input scalar -> (array, component)
};

namespace {

__host__ __device__ __forceinline__ bool gx_is_binary(unsigned op) {
  return op == GX_OP_ADD || op == GX_OP_SUB || op == GX_OP_MUL || op == GX_OP_DIV || op == GX_OP_FMA ||
         op == GX_OP_ATAN2 || op == GX_OP_POW || op == GX_OP_MAX || op == GX_OP_MIN || op == GX_OP_GT ||
         op == GX_OP_LT || op == GX_OP_EQ || op == GX_OP_SELECT;
}

__device__ __forceinline__ float gx_apply(unsigned op, float a, float b, float c) {
  switch (op) {
    case GX_OP_ADD: return __fadd_rn(a, b);
    case GX_OP_SUB: return __fsub_rn(a, b);
    case GX_OP_MUL: return __fmul_rn(a, b);
    case GX_OP_FMA: return __fmaf_rn(a, b, c);
    case GX_OP_DIV: return __fdiv_rn(a, b);
    case GX_OP_NEG: return -a;
    case GX_OP_ABS: return fabsf(a);
    case GX_OP_SQRT: return __fsqrt_rn(a);
    case GX_OP_SIN: return sinf(a);
    case GX_OP_COS: return cosf(a);
    case GX_OP_TAN: return tanf(a);
    case GX_OP_ATAN2: return atan2f(a, b);
    case GX_OP_EXP: return expf(a);
    case GX_OP_LOG: return logf(a);
    case GX_OP_LOG1P: return log1pf(a);
    case GX_OP_TANH: return tanhf(a);
    case GX_OP_SINH: return sinhf(a);
    case GX_OP_COSH: return coshf(a);
    case GX_OP_ASIN: return asinf(a);
    case GX_OP_ACOS: return acosf(a);
    case GX_OP_ATAN: return atanf(a);
    case GX_OP_ASINH: return asinhf(a);
    case GX_OP_ACOSH: return acoshf(a);
    case GX_OP_ATANH: return atanhf(a);
    case GX_OP_ERF: return erff(a);
    case GX_OP_POW: return powf(a, b);
    case GX_OP_LOGISTIC: return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-a)));
    case GX_OP_MAX: return fmaxf(a, b);
    case GX_OP_MIN: return fminf(a, b);
    case GX_OP_SELECT: return a != 0.0f ? b : c;
    case GX_OP_GT: return a > b ? 1.0f : 0.0f;
    case GX_OP_LT: return a < b ? 1.0f : 0.0f;
    case GX_OP_EQ: return a == b ? 1.0f : 0.0f;
    default: return 0.0f;
  }
}

// Shared memory per CTA of T samples:
//   ops[n_ops] (3 words each) | consts[n_consts] | in_base[n_in], in_stride[n_in] |
//   in_stage[n_in*T] | slots[n_slots*T] (slot-major: conflict free) | out[T*(tile+rows)]
template <int T>
__global__ void __launch_bounds__(T) gx_interp_kernel(const __grid_constant__ GxInterpArgs a) {
  extern __shared__ __align__(16) unsigned gx_ismem[];
  unsigned* s_ops = gx_ismem;
  float* s_consts = reinterpret_cast<float*>(s_ops + 3 * a.n_ops);
  int* s_in_base = reinterpret_cast<int*>(s_consts + a.n_consts);
  int* s_in_stride = s_in_base + a.n_in;
  float* s_in = reinterpret_cast<float*>(s_in_stride + a.n_in);
  float* s_slots = s_in + a.n_in * T;
  float* s_out = s_slots + a.n_slots * T;

  const int tid = threadIdx.x;
  const int tile_words = a.n_rows * a.n_cols;
  const int out_words = tile_words + (a.has_primal ? a.n_rows : 0);

  // one-time: plan into shared memory (every thread reads the same op -> broadcast LDS)
  {
    const unsigned* g = reinterpret_cast<const unsigned*>(a.ops);
    for (int i = tid; i < 3 * a.n_ops; i += T) s_ops[i] = g[i];
    for (int i = tid; i < a.n_consts; i += T) s_consts[i] = a.consts[i];
    if (tid == 0) {
      for (int arr = 0; arr < a.n_arr; ++arr)
        for (int c = 0; c < a.in_size[arr]; ++c) {
          s_in_base[a.in_off[arr] + c] = a.in_off[arr] * T + c;
          s_in_stride[a.in_off[arr] + c] = a.in_size[arr];
        }
    }
  }
  __syncthreads();

  const long long n_tiles = (a.batch + T - 1) / T;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * T;
    const int valid = (int)(a.batch - base < T ? a.batch - base : T);
    // coalesced stage-in of every input array's chunk
    for (int arr = 0; arr < a.n_arr; ++arr) {
      const int sz = a.in_size[arr];
      const float* src = a.in[arr] + base * sz;
      float* dst = s_in + a.in_off[arr] * T;
      for (int i = tid; i < T * sz; i += T) dst[i] = i < valid * sz ? __ldg(src + i) : 1.0f;
    }
    __syncthreads();

    auto fetch = [&](unsigned o) -> float {
      const unsigned kind = GX_OPERAND_KIND(o), idx = GX_OPERAND_INDEX(o);
      if (kind == GX_K_SLOT) return s_slots[idx * T + tid];
      if (kind == GX_K_CONST) return s_consts[idx];
      return s_in[s_in_base[idx] + tid * s_in_stride[idx]];
    };
    float* my_out = s_out + tid * out_words;
    for (int i = 0; i < a.n_ops; ++i) {
      const unsigned w0 = s_ops[3 * i], w1 = s_ops[3 * i + 1], w2 = s_ops[3 * i + 2];
      const unsigned op = w0 & 0xffffu, dst = w0 >> 16, oa = w1 & 0xffffu, ob = w1 >> 16, oc = w2 & 0xffffu;
      if (op == GX_OP_STORE_JAC) {
        my_out[dst] = fetch(oa);
      } else if (op == GX_OP_STORE_PRIMAL) {
        my_out[tile_words + dst] = fetch(oa);
      } else {
        const float va = fetch(oa);
        const float vb = gx_is_binary(op) ? fetch(ob) : 0.0f;
        const float vc = (op == GX_OP_FMA || op == GX_OP_SELECT) ? fetch(oc) : 0.0f;
        s_slots[dst * T + tid] = gx_apply(op, va, vb, vc);
      }
    }
    __syncthreads();
    // coalesced stage-out: the tile rows of `valid` consecutive samples are contiguous in HBM
    for (int i = tid; i < valid * tile_words; i += T) {
      const int smp = i / tile_words, w = i - smp * tile_words;
      a.jac[base * tile_words + i] = s_out[smp * out_words + w];
    }
    if (a.has_primal && a.primal != nullptr)
      for (int i = tid; i < valid * a.n_rows; i += T) {
        const int smp = i / a.n_rows, w = i - smp * a.n_rows;
        a.primal[base * a.n_rows + i] = s_out[smp * out_words + tile_words + w];
      }
    __syncthreads();
  }
}

// Plans whose program or live values do not fit in shared memory (forward mode over matrix-valued vertices:
// 10^5 operations, thousands of live values).  The program and the constants stay in global memory - every
// thread of the CTA reads the same word, one broadcast transaction served from L1 - the live values of the
// CTA's T samples sit in a private [n_slots][T] scratch area (slot-major: coalesced), and inputs and outputs
// are accessed in place.  Slow (it is the fallback of the fallback) but it runs every valid plan.
template <int T>
__global__ void __launch_bounds__(T) gx_interp_large_kernel(const __grid_constant__ GxInterpArgs a) {
  const int tid = threadIdx.x;
  const int tile_words = a.n_rows * a.n_cols;
  float* slots = a.scratch + (size_t)blockIdx.x * a.n_slots * T;
  const unsigned* ops = reinterpret_cast<const unsigned*>(a.ops);
  const long long n_tiles = (a.batch + T - 1) / T;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long smp = tile * T + tid;
    const bool valid = smp < a.batch;
    auto fetch = [&](unsigned o) -> float {
      const unsigned kind = GX_OPERAND_KIND(o), idx = GX_OPERAND_INDEX(o);
      if (kind == GX_K_SLOT) return slots[(size_t)idx * T + tid];
      if (kind == GX_K_CONST) return __ldg(a.consts + idx);
      const int arr = __ldg(a.in_map + 2 * idx), comp = __ldg(a.in_map + 2 * idx + 1);
      return valid ? __ldg(a.in[arr] + smp * a.in_size[arr] + comp) : 1.0f;
    };
    for (int i = 0; i < a.n_ops; ++i) {
      const unsigned w0 = __ldg(ops + 3 * i), w1 = __ldg(ops + 3 * i + 1), w2 = __ldg(ops + 3 * i + 2);
      const unsigned op = w0 & 0xffffu, dst = w0 >> 16, oa = w1 & 0xffffu, ob = w1 >> 16, oc = w2 & 0xffffu;
      if (op == GX_OP_STORE_JAC) {
        const float v = fetch(oa);
        if (valid) a.jac[smp * tile_words + dst] = v;
      } else if (op == GX_OP_STORE_PRIMAL) {
        const float v = fetch(oa);
        if (valid && a.has_primal && a.primal != nullptr) a.primal[smp * a.n_rows + dst] = v;
      } else {
        const float va = fetch(oa);
        const float vb = gx_is_binary(op) ? fetch(ob) : 0.0f;
        const float vc = (op == GX_OP_FMA || op == GX_OP_SELECT) ? fetch(oc) : 0.0f;
        slots[(size_t)dst * T + tid] = gx_apply(op, va, vb, vc);
      }
    }
  }
}
constexpr int GX_LARGE_T = 64;           // samples per CTA of the large-plan interpreter
constexpr int GX_LARGE_CTAS_PER_SM = 2;

size_t interp_smem_bytes(const gx_blob_header& h, int T) {
  const size_t tile = (size_t)h.n_rows * h.n_cols + ((h.flags & GX_FLAG_HAS_PRIMAL) ? h.n_rows : 0);
  return 4 * (3 * (size_t)h.n_ops + h.n_consts + 2 * (size_t)h.n_in_scalars + (size_t)h.n_in_scalars * T +
              (size_t)h.n_slots * T + tile * T) + 16;
}

uint64_t fnv1a64(const void* data, size_t n) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t h = 0xCBF29CE484222325ull;
  for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001B3ull;
  return h;
}

// ------------------------------------------------------------------------------------------------
// driver API through the runtime (no link-time dependency on libcuda) + NVRTC through dlopen
// ------------------------------------------------------------------------------------------------
struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

template <class F>
bool load_entry(const char* name, F*& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || p == nullptr) return false;
  fn = reinterpret_cast<F*>(p);
  return true;
}

DriverApi& driver() {
  static DriverApi d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.ok = load_entry("cuModuleLoadData", d.ModuleLoadData) && load_entry("cuModuleUnload", d.ModuleUnload) &&
           load_entry("cuModuleGetFunction", d.ModuleGetFunction) &&
           load_entry("cuFuncSetAttribute", d.FuncSetAttribute) &&
           load_entry("cuFuncGetAttribute", d.FuncGetAttribute) &&
           load_entry("cuOccupancyMaxActiveBlocksPerMultiprocessor", d.OccupancyMaxActiveBlocksPerMultiprocessor) &&
           load_entry("cuLaunchKernelEx", d.LaunchKernelEx) && load_entry("cuGetErrorString", d.GetErrorString);
  });
  return d;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString) driver().GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

struct Nvrtc {
  typedef struct _nvrtcProgram* Program;
  int (*CreateProgram)(Program*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(Program, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(Program, size_t*) = nullptr;
  int (*GetProgramLog)(Program, char*) = nullptr;
  int (*GetCUBINSize)(Program, size_t*) = nullptr;
  int (*GetCUBIN)(Program, char*) = nullptr;
  int (*DestroyProgram)(Program*) = nullptr;
  bool ok = false;
  std::string why;
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    void* h = nullptr;
    for (const char* nm : names)
      if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) {
      n.why = "libnvrtc.so.12 not found";
      return;
    }
#define GX_SYM(field, sym) *(void**)(&n.field) = dlsym(h, sym)
    GX_SYM(CreateProgram, "nvrtcCreateProgram");
    GX_SYM(CompileProgram, "nvrtcCompileProgram");
    GX_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    GX_SYM(GetProgramLog, "nvrtcGetProgramLog");
    GX_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
    GX_SYM(GetCUBIN, "nvrtcGetCUBIN");
    GX_SYM(DestroyProgram, "nvrtcDestroyProgram");
#undef GX_SYM
    n.ok = n.CreateProgram && n.CompileProgram && n.GetProgramLogSize && n.GetProgramLog && n.GetCUBINSize &&
           n.GetCUBIN && n.DestroyProgram;
    if (!n.ok) n.why = "libnvrtc is missing required symbols";
  });
  return n;
}

constexpr int kSpecThreads = 128;
constexpr int kSpecMinBlocks = 4;
constexpr int kSpecOutStages = 1;
// GX_PDL=0 switches programmatic dependent launch off (A/B measurements)
bool use_pdl() {
  static const bool v = [] {
    const char* e = getenv("GX_PDL");
    return !(e && e[0] == '0');
  }();
  return v;
}

// samples per pipeline stage of gx_jacve_host (GX_HOST_CHUNK overrides, for tuning)
int64_t host_chunk() {
  const char* e = getenv("GX_HOST_CHUNK");
  long long n = e ? atoll(e) : 0;
  return (int64_t)(n >= 1024 ? n : (1 << 17));
}

// small first chunks in gx_jacve_host (GX_HOST_RAMP=0: every chunk full sized)
bool host_ramp() {
  const char* e = getenv("GX_HOST_RAMP");
  return !(e && e[0] == '0');
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// plan object
// ------------------------------------------------------------------------------------------------
struct gx_plan {
  gx_blob_header h;
  std::vector<unsigned> in_sizes, out_sizes, in_cols;
  std::vector<float> consts;
  std::vector<gx_op> ops;
  uint64_t key = 0;
  int device = 0;
  int sm_count = 0;
  gx_op* d_ops = nullptr;
  float* d_consts = nullptr;

  // interpreter configuration
  int interp_threads = 0;
  size_t interp_smem = 0;
  int interp_ctas_per_sm = 1;
  bool interp_large = false;   // program / live values exceed shared memory: gx_interp_large_kernel
  float* d_scratch = nullptr;
  // the scratch area is shared by every launch of this plan: launches on different streams (the two slots of the
  // host pipeline, or a caller using several streams) are chained through this event
  std::mutex scratch_mu;
  cudaEvent_t scratch_free = nullptr;
  bool scratch_used = false;
  int* d_in_map = nullptr;
  // AOT kernel
  const GxAotKernel* aot = nullptr;
  int aot_ctas_per_sm = 0;
  int aot_regs = 0;
  // JIT kernel
  CUmodule jit_module = nullptr;
  CUfunction jit_func = nullptr;
  int jit_smem = 0;
  int jit_threads = 0;
  int jit_spt = 1;
  int jit_ctas_per_sm = 0;
  int jit_regs = 0;

  std::atomic<int> kernel_kind{GX_KERNEL_INTERPRETER};   // read by concurrent launches, set by select / specialize
  std::atomic<uint64_t> launches{0};

  // host pipeline resources (lazily created, guarded by host_mu)
  std::mutex host_mu;
  int64_t host_chunk = 0;
  struct HostSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    float* d_in = nullptr;    // all input arrays of one chunk, back to back
    float* d_jac = nullptr;
    float* d_primal = nullptr;
    float* h_in = nullptr;    // pinned staging (pageable callers only)
    float* h_jac = nullptr;
    float* h_primal = nullptr;
  } slots[2];
  bool host_ready = false;
  bool host_staging = false;
  int64_t host_dev_cap = 0;     // samples the device buffers of a slot hold
  int64_t host_stage_cap = 0;   // samples the pinned staging buffers of a slot hold
};

namespace {

int validate_blob(const void* blob, size_t nbytes, gx_plan* p) {
  if (blob == nullptr) return fail(GX_ERR_BAD_ARG, "gx_plan_create: blob is NULL");
  if (nbytes < sizeof(gx_blob_header)) return fail(GX_ERR_BAD_BLOB, "plan blob shorter than its header");
  memcpy(&p->h, blob, sizeof(gx_blob_header));
  const gx_blob_header& h = p->h;
  if (h.magic != GX_BLOB_MAGIC) return fail(GX_ERR_BAD_BLOB, "plan blob: bad magic (not a graphax_b200 plan)");
  if (h.version != GX_BLOB_VERSION) return fail(GX_ERR_BAD_BLOB, "plan blob: unsupported version");
  if (h.total_bytes != nbytes) return fail(GX_ERR_BAD_BLOB, "plan blob: size field does not match buffer length");
  if (h.n_in_arrays == 0 || h.n_in_arrays > GX_MAX_ARRAYS || h.n_out_arrays == 0 || h.n_out_arrays > GX_MAX_ARRAYS)
    return fail(GX_ERR_BAD_BLOB, "plan blob: between 1 and 32 (GX_MAX_ARRAYS) input and output arrays are supported");
  const size_t need = sizeof(gx_blob_header) + 4ull * (2 * h.n_in_arrays + h.n_out_arrays) + 4ull * h.n_consts +
                      sizeof(gx_op) * (size_t)h.n_ops;
  if (need != nbytes) return fail(GX_ERR_BAD_BLOB, "plan blob: section sizes do not add up to the buffer length");
  if (h.n_slots > 0x3fff || h.n_consts > 0x3fff || h.n_in_scalars > 0x3fff)
    return fail(GX_ERR_BAD_BLOB, "plan blob: index space exceeds the 14-bit operand encoding");
  if (h.n_rows == 0 || h.n_cols == 0 || (size_t)h.n_rows * h.n_cols > 0xffff)
    return fail(GX_ERR_BAD_BLOB, "plan blob: Jacobian tile must have between 1 and 65535 entries");

  const unsigned char* cur = static_cast<const unsigned char*>(blob) + sizeof(gx_blob_header);
  auto take_u32 = [&](std::vector<unsigned>& v, size_t n) {
    v.resize(n);
    memcpy(v.data(), cur, 4 * n);
    cur += 4 * n;
  };
  take_u32(p->in_sizes, h.n_in_arrays);
  take_u32(p->out_sizes, h.n_out_arrays);
  take_u32(p->in_cols, h.n_in_arrays);
  p->consts.resize(h.n_consts);
  memcpy(p->consts.data(), cur, 4ull * h.n_consts);
  cur += 4ull * h.n_consts;
  p->ops.resize(h.n_ops);
  memcpy(p->ops.data(), cur, sizeof(gx_op) * (size_t)h.n_ops);

  size_t sum_in = 0, sum_out = 0;
  for (unsigned s : p->in_sizes) {
    if (s == 0) return fail(GX_ERR_BAD_BLOB, "plan blob: empty input array");
    sum_in += s;
  }
  for (unsigned s : p->out_sizes) sum_out += s;
  if (sum_in != h.n_in_scalars) return fail(GX_ERR_BAD_BLOB, "plan blob: in_sizes do not sum to n_in_scalars");
  if (sum_out != h.n_rows) return fail(GX_ERR_BAD_BLOB, "plan blob: out_sizes do not sum to n_rows");

  const unsigned tile = h.n_rows * h.n_cols;
  std::vector<unsigned char> stored(tile, 0), stored_p(h.n_rows, 0), defined(h.n_slots, 0);
  auto check_operand = [&](unsigned o, size_t i) -> int {
    const unsigned kind = GX_OPERAND_KIND(o), idx = GX_OPERAND_INDEX(o);
    char buf[96];
    if (kind == GX_K_SLOT) {
      if (idx >= h.n_slots || !defined[idx]) {
        snprintf(buf, sizeof buf, "plan blob: op %zu reads slot %u before it is written", i, idx);
        return fail(GX_ERR_BAD_BLOB, buf);
      }
    } else if (kind == GX_K_CONST) {
      if (idx >= h.n_consts) return fail(GX_ERR_BAD_BLOB, "plan blob: constant index out of range");
    } else if (kind == GX_K_INPUT) {
      if (idx >= h.n_in_scalars) return fail(GX_ERR_BAD_BLOB, "plan blob: input index out of range");
    } else {
      return fail(GX_ERR_BAD_BLOB, "plan blob: unknown operand kind");
    }
    return GX_OK;
  };
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const gx_op& o = p->ops[i];
    if (o.op >= GX_OP_COUNT || o.op == GX_OP_INPUT || o.op == GX_OP_CONST)
      return fail(GX_ERR_BAD_BLOB, "plan blob: invalid opcode in op stream");
    int rc = check_operand(o.a, i);
    if (rc) return rc;
    const bool binary = gx_is_binary(o.op);
    if (binary && (rc = check_operand(o.b, i))) return rc;
    if ((o.op == GX_OP_FMA || o.op == GX_OP_SELECT) && (rc = check_operand(o.c, i))) return rc;
    if (o.op == GX_OP_STORE_JAC) {
      if (o.dst >= tile) return fail(GX_ERR_BAD_BLOB, "plan blob: store_jac outside the Jacobian tile");
      if (stored[o.dst]++) return fail(GX_ERR_BAD_BLOB, "plan blob: Jacobian tile entry stored twice");
    } else if (o.op == GX_OP_STORE_PRIMAL) {
      if (!(h.flags & GX_FLAG_HAS_PRIMAL) || o.dst >= h.n_rows)
        return fail(GX_ERR_BAD_BLOB, "plan blob: store_primal outside the output vector");
      if (stored_p[o.dst]++) return fail(GX_ERR_BAD_BLOB, "plan blob: primal output stored twice");
    } else {
      if (o.dst >= h.n_slots) return fail(GX_ERR_BAD_BLOB, "plan blob: destination slot out of range");
      defined[o.dst] = 1;
    }
  }
  for (unsigned i = 0; i < tile; ++i)
    if (!stored[i]) return fail(GX_ERR_BAD_BLOB, "plan blob: a Jacobian tile entry is never stored");
  if (h.flags & GX_FLAG_HAS_PRIMAL)
    for (unsigned i = 0; i < h.n_rows; ++i)
      if (!stored_p[i]) return fail(GX_ERR_BAD_BLOB, "plan blob: a primal output is never stored");
  p->key = fnv1a64(blob, nbytes);
  return GX_OK;
}

template <int T>
int config_interp(gx_plan* p, int max_smem_optin) {
  const size_t smem = interp_smem_bytes(p->h, T);
  if ((long long)smem > max_smem_optin) return 1;
  GX_CUDA(cudaFuncSetAttribute(gx_interp_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int n = 0;
  GX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gx_interp_kernel<T>, T, smem));
  if (n < 1) return 1;
  p->interp_threads = T;
  p->interp_smem = smem;
  p->interp_ctas_per_sm = n;
  return GX_OK;
}

void fill_args(const gx_plan* p, GxLaunchArgs& a, int64_t batch, const void* const* in_ptrs, void* jac, void* primal,
               int warp_tile) {
  memset(&a, 0, sizeof a);
  bool aligned = (reinterpret_cast<uintptr_t>(jac) & 15u) == 0 &&
                 (primal == nullptr || (reinterpret_cast<uintptr_t>(primal) & 15u) == 0);
  for (unsigned i = 0; i < p->h.n_in_arrays; ++i) {
    a.in[i] = static_cast<const float*>(in_ptrs[i]);
    aligned = aligned && (reinterpret_cast<uintptr_t>(in_ptrs[i]) & 15u) == 0;
  }
  a.jac = static_cast<float*>(jac);
  a.primal = static_cast<float*>(primal);
  a.batch = batch;
  a.n_bulk_tiles = aligned ? batch / warp_tile : 0;  // a warp tile is 32 x samples_per_thread samples
}

int launch_on(gx_plan* p, int64_t batch, const void* const* in_ptrs, void* jac, void* primal, cudaStream_t stream) {
  if (batch == 0) return GX_OK;
  const int kind = p->kernel_kind.load(std::memory_order_acquire);
  if (kind == GX_KERNEL_AOT && p->aot) {
    GxLaunchArgs a;
    const int wt = 32 * p->aot->samples_per_thread;
    fill_args(p, a, batch, in_ptrs, jac, primal, wt);
    const int wpc = p->aot->threads / 32;  // every warp is an independent pipeline over its own tiles
    const long long ctas = ((batch + wt - 1) / wt + wpc - 1) / wpc;
    const int grid = (int)std::min<long long>(ctas, (long long)p->sm_count * p->aot_ctas_per_sm);
    void* params[] = {&a};
    // programmatic dependent launch: the kernel's prologue may overlap the tail of its predecessor on the
    // stream; the kernel itself waits (griddepcontrol.wait) before it touches global memory
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(p->aot->threads);
    cfg.dynamicSmemBytes = p->aot->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    GX_CUDA(cudaLaunchKernelExC(&cfg, p->aot->func, params));
  } else if (kind == GX_KERNEL_JIT && p->jit_func) {
    GxLaunchArgs a;
    const int wt = 32 * p->jit_spt;
    fill_args(p, a, batch, in_ptrs, jac, primal, wt);
    const int wpc = p->jit_threads / 32;
    const long long ctas = ((batch + wt - 1) / wt + wpc - 1) / wpc;
    const int grid = (int)std::min<long long>(ctas, (long long)p->sm_count * p->jit_ctas_per_sm);
    void* params[] = {&a};
    CUlaunchConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDimX = grid;
    cfg.gridDimY = cfg.gridDimZ = 1;
    cfg.blockDimX = p->jit_threads;
    cfg.blockDimY = cfg.blockDimZ = 1;
    cfg.sharedMemBytes = p->jit_smem;
    cfg.hStream = (CUstream)stream;
    CUlaunchAttribute attr[1];
    attr[0].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    attr[0].value.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    CUresult r = driver().LaunchKernelEx(&cfg, p->jit_func, params, nullptr);
    if (r != CUDA_SUCCESS) return fail(GX_ERR_CUDA, "cuLaunchKernelEx(jit): " + cu_err(r));
  } else {
    GxInterpArgs a;
    memset(&a, 0, sizeof a);
    unsigned off = 0;
    for (unsigned i = 0; i < p->h.n_in_arrays; ++i) {
      a.in[i] = static_cast<const float*>(in_ptrs[i]);
      a.in_size[i] = (int)p->in_sizes[i];
      a.in_off[i] = (int)off;
      off += p->in_sizes[i];
    }
    a.jac = static_cast<float*>(jac);
    a.primal = static_cast<float*>(primal);
    a.batch = batch;
    a.ops = p->d_ops;
    a.consts = p->d_consts;
    a.n_ops = (int)p->h.n_ops;
    a.n_consts = (int)p->h.n_consts;
    a.n_slots = (int)p->h.n_slots;
    a.n_in = (int)p->h.n_in_scalars;
    a.n_rows = (int)p->h.n_rows;
    a.n_cols = (int)p->h.n_cols;
    a.n_arr = (int)p->h.n_in_arrays;
    a.has_primal = (p->h.flags & GX_FLAG_HAS_PRIMAL) ? 1 : 0;
    const int T = p->interp_threads;
    const long long tiles = (batch + T - 1) / T;
    const int grid = (int)std::min<long long>(tiles, (long long)p->sm_count * p->interp_ctas_per_sm);
    a.scratch = p->d_scratch;
    a.in_map = p->d_in_map;
    if (p->interp_large) {
      // live values of CTA b sit in scratch[b]: two launches of one plan must not overlap.  Outside stream capture
      // (a captured stream orders its own launches) every launch waits for the previous one's completion event.
      std::lock_guard<std::mutex> lock(p->scratch_mu);
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &cap);
      const bool chain = cap == cudaStreamCaptureStatusNone && p->scratch_free != nullptr;
      if (chain && p->scratch_used) GX_CUDA(cudaStreamWaitEvent(stream, p->scratch_free, 0));
      gx_interp_large_kernel<GX_LARGE_T><<<grid, GX_LARGE_T, 0, stream>>>(a);
      GX_CUDA(cudaGetLastError());
      if (chain) {
        GX_CUDA(cudaEventRecord(p->scratch_free, stream));
        p->scratch_used = true;
      }
    } else if (T == 128)
      gx_interp_kernel<128><<<grid, 128, p->interp_smem, stream>>>(a);
    else if (T == 64)
      gx_interp_kernel<64><<<grid, 64, p->interp_smem, stream>>>(a);
    else
      gx_interp_kernel<32><<<grid, 32, p->interp_smem, stream>>>(a);
    GX_CUDA(cudaGetLastError());
  }
  p->launches.fetch_add(1, std::memory_order_relaxed);
  return GX_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// exported API
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* gx_last_error(void) { return g_error.c_str(); }
int gx_abi_version(void) { return GX_ABI_VERSION; }
int gx_aot_count(void) { return (int)aot_registry().size(); }
uint64_t gx_aot_key(int index) {
  return index >= 0 && index < (int)aot_registry().size() ? aot_registry()[index].key : 0;
}
uint64_t gx_hash_bytes(const void* data, size_t nbytes) { return data ? fnv1a64(data, nbytes) : 0; }

int gx_plan_create(const void* blob, size_t nbytes, int device, gx_plan** out) {
  if (out == nullptr) return fail(GX_ERR_BAD_ARG, "gx_plan_create: out is NULL");
  *out = nullptr;
  gx_plan* p = new gx_plan();
  int rc = validate_blob(blob, nbytes, p);
  if (rc != GX_OK) {
    delete p;
    return rc;
  }
  p->device = device;
  auto cleanup = [&](int code) {
    gx_plan_destroy(p);
    return code;
  };
#define GX_CUDA_P(expr)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (expr);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return cleanup(fail(GX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__) +         \
                                           " (graphax_b200 has no CPU fallback; a CUDA device is required)")); \
  } while (0)
  GX_CUDA_P(cudaSetDevice(device));
  int max_optin = 0;
  GX_CUDA_P(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device));
  GX_CUDA_P(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  GX_CUDA_P(cudaMalloc(&p->d_ops, std::max<size_t>(sizeof(gx_op) * p->ops.size(), 16)));
  GX_CUDA_P(cudaMalloc(&p->d_consts, std::max<size_t>(4 * p->consts.size(), 16)));
  if (!p->ops.empty())
    GX_CUDA_P(cudaMemcpy(p->d_ops, p->ops.data(), sizeof(gx_op) * p->ops.size(), cudaMemcpyHostToDevice));
  if (!p->consts.empty())
    GX_CUDA_P(cudaMemcpy(p->d_consts, p->consts.data(), 4 * p->consts.size(), cudaMemcpyHostToDevice));

  // interpreter: the largest CTA whose live-value file fits, preferring two resident CTAs per SM
  rc = 1;
  if (interp_smem_bytes(p->h, 128) * 2 <= (size_t)max_optin) rc = config_interp<128>(p, max_optin);
  if (rc == 1 && interp_smem_bytes(p->h, 64) * 2 <= (size_t)max_optin) rc = config_interp<64>(p, max_optin);
  if (rc == 1) rc = config_interp<128>(p, max_optin);
  if (rc == 1) rc = config_interp<64>(p, max_optin);
  if (rc == 1) rc = config_interp<32>(p, max_optin);
  if (rc == 1) {
    // too long or too many live values for shared memory: global-memory interpreter with a private scratch area
    p->interp_large = true;
    p->interp_threads = GX_LARGE_T;
    p->interp_smem = 0;
    p->interp_ctas_per_sm = GX_LARGE_CTAS_PER_SM;
    const size_t ctas = (size_t)p->sm_count * GX_LARGE_CTAS_PER_SM;
    GX_CUDA_P(cudaMalloc(&p->d_scratch, std::max<size_t>(4 * ctas * p->h.n_slots * GX_LARGE_T, 16)));
    GX_CUDA_P(cudaEventCreateWithFlags(&p->scratch_free, cudaEventDisableTiming));
    std::vector<int> in_map(2 * (size_t)p->h.n_in_scalars);
    size_t k = 0;
    for (unsigned arr = 0; arr < p->h.n_in_arrays; ++arr)
      for (unsigned c = 0; c < p->in_sizes[arr]; ++c, ++k) {
        in_map[2 * k] = (int)arr;
        in_map[2 * k + 1] = (int)c;
      }
    GX_CUDA_P(cudaMalloc(&p->d_in_map, std::max<size_t>(4 * in_map.size(), 16)));
    GX_CUDA_P(cudaMemcpy(p->d_in_map, in_map.data(), 4 * in_map.size(), cudaMemcpyHostToDevice));
    rc = GX_OK;
  }
  if (rc != GX_OK) return cleanup(rc);

  // plan-specialised kernel compiled into the library for exactly this blob?
  for (const GxAotKernel& k : aot_registry()) {
    if (k.key != p->key) continue;
    if (k.n_in != (int)p->h.n_in_scalars || k.n_rows != (int)p->h.n_rows || k.n_cols != (int)p->h.n_cols)
      return cleanup(fail(GX_ERR_BAD_BLOB, "AOT kernel with the plan's key has a different shape (hash collision?)"));
    GX_CUDA_P(cudaFuncSetAttribute(k.func, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem_bytes));
    int n = 0;
    GX_CUDA_P(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k.func, k.threads, k.smem_bytes));
    if (n < 1) break;
    cudaFuncAttributes fa;
    GX_CUDA_P(cudaFuncGetAttributes(&fa, k.func));
    p->aot = &k;
    p->aot_ctas_per_sm = n;
    p->aot_regs = fa.numRegs;
    p->kernel_kind = GX_KERNEL_AOT;
    break;
  }
#undef GX_CUDA_P
  *out = p;
  return GX_OK;
}

void gx_plan_destroy(gx_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (auto& s : p->slots) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_jac) cudaFree(s.d_jac);
    if (s.d_primal) cudaFree(s.d_primal);
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.h_jac) cudaFreeHost(s.h_jac);
    if (s.h_primal) cudaFreeHost(s.h_primal);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (p->jit_module && driver().ok) driver().ModuleUnload(p->jit_module);
  if (p->d_ops) cudaFree(p->d_ops);
  if (p->d_consts) cudaFree(p->d_consts);
  if (p->d_scratch) cudaFree(p->d_scratch);
  if (p->scratch_free) cudaEventDestroy(p->scratch_free);
  if (p->d_in_map) cudaFree(p->d_in_map);
  delete p;
}

int gx_plan_info(const gx_plan* p, gx_info* out) {
  if (!p || !out) return fail(GX_ERR_BAD_ARG, "gx_plan_info: NULL argument");
  memset(out, 0, sizeof *out);
  out->abi_version = GX_ABI_VERSION;
  out->n_in_arrays = p->h.n_in_arrays;
  out->n_out_arrays = p->h.n_out_arrays;
  out->n_in_scalars = p->h.n_in_scalars;
  out->n_rows = p->h.n_rows;
  out->n_cols = p->h.n_cols;
  out->n_ops = p->h.n_ops;
  out->n_slots = p->h.n_slots;
  out->has_primal = (p->h.flags & GX_FLAG_HAS_PRIMAL) ? 1 : 0;
  const int kind_now = p->kernel_kind.load(std::memory_order_acquire);
  out->kernel_kind = (uint32_t)kind_now;
  if (kind_now == GX_KERNEL_AOT && p->aot) {
    out->threads = p->aot->threads;
    out->smem_bytes = p->aot->smem_bytes;
    out->ctas_per_sm = p->aot_ctas_per_sm;
    out->regs_per_thread = p->aot_regs;
  } else if (kind_now == GX_KERNEL_JIT && p->jit_func) {
    out->threads = p->jit_threads;
    out->smem_bytes = p->jit_smem;
    out->ctas_per_sm = p->jit_ctas_per_sm;
    out->regs_per_thread = p->jit_regs;
  } else {
    out->threads = p->interp_threads;
    out->smem_bytes = (uint32_t)p->interp_smem;
    out->ctas_per_sm = p->interp_ctas_per_sm;
  }
  for (unsigned i = 0; i < p->h.n_in_arrays; ++i) out->in_sizes[i] = p->in_sizes[i];
  for (unsigned i = 0; i < p->h.n_out_arrays; ++i) out->out_sizes[i] = p->out_sizes[i];
  out->plan_key = p->key;
  out->launches = p->launches.load();
  return GX_OK;
}

int gx_plan_select_kernel(gx_plan* p, int kind) {
  if (!p) return fail(GX_ERR_BAD_ARG, "gx_plan_select_kernel: plan is NULL");
  if (kind == GX_KERNEL_INTERPRETER || (kind == GX_KERNEL_AOT && p->aot) || (kind == GX_KERNEL_JIT && p->jit_func)) {
    p->kernel_kind = kind;
    return GX_OK;
  }
  return fail(GX_ERR_UNSUPPORTED, "requested kernel kind is not attached to this plan");
}

int gx_plan_specialize(gx_plan* p, const char* src, size_t nbytes, const char* arch) {
  return gx_plan_specialize_opts(p, src, nbytes, arch, kSpecThreads, kSpecMinBlocks, kSpecOutStages, 1);
}

int gx_plan_specialize_opts(gx_plan* p, const char* src, size_t nbytes, const char* arch, int threads,
                            int min_blocks, int out_stages, int spt) {
  if (!p || !src) return fail(GX_ERR_BAD_ARG, "gx_plan_specialize: NULL argument");
  if (threads < 32 || threads > 1024 || threads % 32 || min_blocks < 1 || min_blocks > 32 || out_stages < 1 ||
      out_stages > 2 || spt < 1 || spt > 2)
    return fail(GX_ERR_BAD_ARG,
                "gx_plan_specialize: threads must be a multiple of 32, out_stages and samples_per_thread 1 or 2");
  (void)nbytes;
  Nvrtc& rt = nvrtc();
  if (!rt.ok) return fail(GX_ERR_JIT, "NVRTC unavailable: " + rt.why);
  GX_CUDA(cudaSetDevice(p->device));
  GX_CUDA(cudaFree(0));
  if (!driver().ok) return fail(GX_ERR_JIT, "CUDA driver entry points unavailable");

  Nvrtc::Program prog = nullptr;
  if (rt.CreateProgram(&prog, src, "gx_plan.cu", 0, nullptr, nullptr) != 0)
    return fail(GX_ERR_JIT, "nvrtcCreateProgram failed");
  std::string archopt = std::string("--gpu-architecture=") + (arch && *arch ? arch : "sm_100a");
  // input stages per warp: same rule as codegen.input_stages
  const int tile = (int)(p->h.n_rows * p->h.n_cols + ((p->h.flags & GX_FLAG_HAS_PRIMAL) ? p->h.n_rows : 0));
  auto smem_for = [&](int in_stages) {
    return (threads / 32) * (in_stages * (int)p->h.n_in_scalars * 32 * spt + out_stages * tile * 32 * spt + 4) * 4;
  };
  // GX_INPUT_STAGED=0: one-sample-per-thread kernels load their inputs straight into registers (no stage)
  const char* staged_env = getenv("GX_INPUT_STAGED");
  const bool staged = spt != 1 || !(staged_env && staged_env[0] == '0');
  const int in_stages = !staged ? 0 : smem_for(2) <= 227 * 1024 ? 2 : 1;
  char d_threads[48], d_blocks[48], d_stages[48], d_spt[48], d_in[48];
  snprintf(d_in, sizeof d_in, "-DGX_IN_STAGES=%d", in_stages);
  snprintf(d_spt, sizeof d_spt, "-DGX_SPT=%d", spt);
  snprintf(d_threads, sizeof d_threads, "-DGX_THREADS=%d", threads);
  snprintf(d_blocks, sizeof d_blocks, "-DGX_MIN_BLOCKS=%d", min_blocks);
  snprintf(d_stages, sizeof d_stages, "-DGX_OUT_STAGES=%d", out_stages);
  const char* opts[] = {archopt.c_str(), "-std=c++17", "-lineinfo", d_threads, d_blocks, d_stages, d_spt, d_in};
  const int rc = rt.CompileProgram(prog, 8, opts);
  if (rc != 0) {
    size_t n = 0;
    rt.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) rt.GetProgramLog(prog, &log[0]);
    rt.DestroyProgram(&prog);
    return fail(GX_ERR_JIT, "nvrtcCompileProgram failed:\n" + log);
  }
  size_t n = 0;
  rt.GetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  rt.GetCUBIN(prog, cubin.data());
  rt.DestroyProgram(&prog);

  CUmodule mod = nullptr;
  CUresult r = driver().ModuleLoadData(&mod, cubin.data());
  if (r != CUDA_SUCCESS) return fail(GX_ERR_JIT, "cuModuleLoadData: " + cu_err(r));
  CUfunction fn = nullptr;
  r = driver().ModuleGetFunction(&fn, mod, "gx_jacve_kernel");
  if (r != CUDA_SUCCESS) {
    driver().ModuleUnload(mod);
    return fail(GX_ERR_JIT, "cuModuleGetFunction(gx_jacve_kernel): " + cu_err(r));
  }
  const int smem = smem_for(in_stages);
  r = driver().FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem);
  int occ = 0, regs = 0;
  if (r == CUDA_SUCCESS) r = driver().OccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem);
  if (r == CUDA_SUCCESS) r = driver().FuncGetAttribute(&regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fn);
  if (r != CUDA_SUCCESS || occ < 1) {
    driver().ModuleUnload(mod);
    return fail(GX_ERR_JIT, "specialised kernel does not fit on the device: " + cu_err(r));
  }
  if (p->jit_module) driver().ModuleUnload(p->jit_module);
  p->jit_module = mod;
  p->jit_func = fn;
  p->jit_smem = smem;
  p->jit_threads = threads;
  p->jit_spt = spt;
  p->jit_ctas_per_sm = occ;
  p->jit_regs = regs;
  p->kernel_kind = GX_KERNEL_JIT;
  return GX_OK;
}

int gx_jacve_launch(gx_plan* p, int64_t batch, const void* const* in_ptrs, void* jac_out, void* primal_out,
                    void* stream) {
  if (!p) return fail(GX_ERR_BAD_ARG, "gx_jacve_launch: plan is NULL");
  if (batch < 0) return fail(GX_ERR_BAD_ARG, "gx_jacve_launch: negative batch");
  if (batch == 0) return GX_OK;
  if (!in_ptrs || !jac_out) return fail(GX_ERR_BAD_ARG, "gx_jacve_launch: NULL buffer");
  for (unsigned i = 0; i < p->h.n_in_arrays; ++i)
    if (!in_ptrs[i]) return fail(GX_ERR_BAD_ARG, "gx_jacve_launch: NULL input array");
  if (primal_out && !(p->h.flags & GX_FLAG_HAS_PRIMAL))
    return fail(GX_ERR_BAD_ARG, "gx_jacve_launch: plan was built without primal outputs");
  GX_CUDA(cudaSetDevice(p->device));
  return launch_on(p, batch, in_ptrs, jac_out, primal_out, static_cast<cudaStream_t>(stream));
}

int gx_jacve_host(gx_plan* p, int64_t batch, const void* const* in_ptrs, void* jac_out, void* primal_out,
                  int pinned) {
  if (!p) return fail(GX_ERR_BAD_ARG, "gx_jacve_host: plan is NULL");
  if (batch < 0) return fail(GX_ERR_BAD_ARG, "gx_jacve_host: negative batch");
  if (batch == 0) return GX_OK;
  if (!in_ptrs || !jac_out) return fail(GX_ERR_BAD_ARG, "gx_jacve_host: NULL buffer");
  if (primal_out && !(p->h.flags & GX_FLAG_HAS_PRIMAL))
    return fail(GX_ERR_BAD_ARG, "gx_jacve_host: plan was built without primal outputs");
  std::lock_guard<std::mutex> lock(p->host_mu);
  GX_CUDA(cudaSetDevice(p->device));
  const size_t n_in = p->h.n_in_scalars, tile = (size_t)p->h.n_rows * p->h.n_cols, rows = p->h.n_rows;
  if (p->host_chunk == 0) p->host_chunk = host_chunk();   // fixed at the plan's first host call
  const int64_t base = p->host_chunk;
  const bool staging = !pinned;
  // Chunk schedule.  The result copy (D2H, 5x the input bytes for RoeFlux_3d) is what the call waits for, so the first
  // two chunks are half sized to start it early; everything after that moves `base` samples.  (Measured on a B200,
  // 2^20 RoeFlux_3d samples, base 2^17: 254.6 vs 254.1 M Jacobians/s with all chunks full sized, and the same with
  // chunks growing to 4 x base - neither the fill nor the number of copies is where the last 9 % to the raw-copy floor
  // go; a synchronous call cannot hide its first H2D copy and its last event wait the way back-to-back copies do.)
  const int64_t chunk_cap = base;
  std::vector<int64_t> starts, sizes;
  {
    int64_t b = 0, sz = host_ramp() ? std::max<int64_t>(base / 2, std::min<int64_t>(base, 1024)) : base;
    int n = 0;
    while (b < batch) {
      const int64_t nb = std::min<int64_t>(sz, batch - b);
      starts.push_back(b);
      sizes.push_back(nb);
      b += nb;
      if (++n >= 2) sz = std::min<int64_t>(chunk_cap, sz * 2);
    }
  }
  const int64_t need_dev = *std::max_element(sizes.begin(), sizes.end());
  if (!p->host_ready) {
    for (auto& s : p->slots) {
      GX_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      GX_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    p->host_ready = true;
  }
  if (need_dev > p->host_dev_cap) {   // first call, or a larger batch than any before: (re)allocate the device side
    const int64_t cap = std::min<int64_t>(chunk_cap, (need_dev + 1023) / 1024 * 1024);
    for (auto& s : p->slots) {
      GX_CUDA(cudaStreamSynchronize(s.stream));
      if (s.d_in) cudaFree(s.d_in);
      if (s.d_jac) cudaFree(s.d_jac);
      if (s.d_primal) cudaFree(s.d_primal);
      s.d_in = s.d_jac = s.d_primal = nullptr;
      GX_CUDA(cudaMalloc(&s.d_in, 4 * n_in * cap));
      GX_CUDA(cudaMalloc(&s.d_jac, 4 * tile * cap));
      GX_CUDA(cudaMalloc(&s.d_primal, 4 * rows * cap));
    }
    p->host_dev_cap = cap;
  }
  if (staging && !p->host_staging) {
    for (auto& s : p->slots) {
      GX_CUDA(cudaMallocHost(&s.h_in, 4 * n_in * base));
      GX_CUDA(cudaMallocHost(&s.h_jac, 4 * tile * base));
      GX_CUDA(cudaMallocHost(&s.h_primal, 4 * rows * base));
    }
    p->host_staging = true;
    p->host_stage_cap = base;
  }
  const int64_t dev_cap = p->host_dev_cap, stage_cap = p->host_stage_cap;
  const int64_t n_chunks = (int64_t)starts.size();
  auto drain = [&](int64_t c) -> int {  // staged results of chunk c -> caller memory
    auto& s = p->slots[c & 1];
    GX_CUDA(cudaEventSynchronize(s.done));
    if (staging) {
      const int64_t b0 = starts[c], nb = sizes[c];
      memcpy(static_cast<float*>(jac_out) + b0 * tile, s.h_jac, 4 * tile * nb);
      if (primal_out) memcpy(static_cast<float*>(primal_out) + b0 * rows, s.h_primal, 4 * rows * nb);
    }
    return GX_OK;
  };
  for (int64_t c = 0; c < n_chunks; ++c) {
    auto& s = p->slots[c & 1];
    const int64_t b0 = starts[c], nb = sizes[c];
    // Staged (pageable) buffers: the host copies chunk c-2 out of the slot before the slot is reused.  Pinned buffers
    // need no host in the loop at all: a slot's stream is in order (D2H of chunk c-2, then H2D of chunk c into the same
    // device buffers), so the whole call is enqueued ahead of the GPU and a descheduled host thread cannot stall it.
    if (c >= 2 && staging) {
      int rc = drain(c - 2);
      if (rc) return rc;
    }
    const void* d_ptrs[GX_MAX_ARRAYS];
    size_t off = 0;
    for (unsigned i = 0; i < p->h.n_in_arrays; ++i) {
      const size_t sz = p->in_sizes[i];
      const float* src = static_cast<const float*>(in_ptrs[i]) + b0 * sz;
      float* dst = s.d_in + off * dev_cap;
      if (staging) {
        float* st = s.h_in + off * stage_cap;
        memcpy(st, src, 4 * sz * nb);
        src = st;
      }
      GX_CUDA(cudaMemcpyAsync(dst, src, 4 * sz * nb, cudaMemcpyHostToDevice, s.stream));
      d_ptrs[i] = dst;
      off += sz;
    }
    int rc = launch_on(p, nb, d_ptrs, s.d_jac, primal_out ? s.d_primal : nullptr, s.stream);
    if (rc) return rc;
    float* hj = staging ? s.h_jac : static_cast<float*>(jac_out) + b0 * tile;
    GX_CUDA(cudaMemcpyAsync(hj, s.d_jac, 4 * tile * nb, cudaMemcpyDeviceToHost, s.stream));
    if (primal_out) {
      float* hp = staging ? s.h_primal : static_cast<float*>(primal_out) + b0 * rows;
      GX_CUDA(cudaMemcpyAsync(hp, s.d_primal, 4 * rows * nb, cudaMemcpyDeviceToHost, s.stream));
    }
    GX_CUDA(cudaEventRecord(s.done, s.stream));
  }
  for (int64_t c = std::max<int64_t>(0, n_chunks - 2); c < n_chunks; ++c) {
    int rc = drain(c);
    if (rc) return rc;
  }
  return GX_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// generic run-time compiled kernels (fused elementwise kernels of the dense / tensor-valued path)
// ------------------------------------------------------------------------------------------------
struct gx_jit_kernel {
  CUmodule module = nullptr;
  CUfunction func = nullptr;
  int device = 0;
  bool smem_set = false;
};

#include "gx_gemm_kernel_decl.h"

namespace {
// out[c] (+)= sum_r in[r, c]; one CTA owns a 32-column strip of up to rows_per_cta rows
__global__ void gx_colsum_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                     int rows_per_cta) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  float acc = 0.0f;
  if (c < cols)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) acc += in[(size_t)r * cols + c];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
    atomicAdd(out + c, s);
  }
}
}  // namespace

extern "C" {

int gx_jit_compile(const char* src, const char* kernel_name, int device, gx_jit_kernel** out) {
  if (!src || !kernel_name || !out) return fail(GX_ERR_BAD_ARG, "gx_jit_compile: NULL argument");
  *out = nullptr;
  Nvrtc& rt = nvrtc();
  if (!rt.ok) return fail(GX_ERR_JIT, "NVRTC unavailable: " + rt.why);
  GX_CUDA(cudaSetDevice(device));
  GX_CUDA(cudaFree(0));
  if (!driver().ok) return fail(GX_ERR_JIT, "CUDA driver entry points unavailable");
  Nvrtc::Program prog = nullptr;
  if (rt.CreateProgram(&prog, src, "gx_jit.cu", 0, nullptr, nullptr) != 0)
    return fail(GX_ERR_JIT, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
  if (rt.CompileProgram(prog, 3, opts) != 0) {
    size_t n = 0;
    rt.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) rt.GetProgramLog(prog, &log[0]);
    rt.DestroyProgram(&prog);
    return fail(GX_ERR_JIT, "nvrtcCompileProgram failed:\n" + log);
  }
  size_t n = 0;
  rt.GetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  rt.GetCUBIN(prog, cubin.data());
  rt.DestroyProgram(&prog);
  gx_jit_kernel* k = new gx_jit_kernel();
  k->device = device;
  CUresult r = driver().ModuleLoadData(&k->module, cubin.data());
  if (r == CUDA_SUCCESS) r = driver().ModuleGetFunction(&k->func, k->module, kernel_name);
  if (r != CUDA_SUCCESS) {
    if (k->module) driver().ModuleUnload(k->module);
    delete k;
    return fail(GX_ERR_JIT, "loading the compiled kernel failed: " + cu_err(r));
  }
  *out = k;
  return GX_OK;
}

void gx_jit_destroy(gx_jit_kernel* k) {
  if (!k) return;
  if (k->module && driver().ok) driver().ModuleUnload(k->module);
  delete k;
}

// args[i] points at the value of the i-th kernel parameter (the cuLaunchKernel convention)
int gx_jit_launch(gx_jit_kernel* k, unsigned grid_x, unsigned grid_y, unsigned block_x, void** args, void* stream) {
  if (!k || !args) return fail(GX_ERR_BAD_ARG, "gx_jit_launch: NULL argument");
  CUlaunchConfig cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDimX = grid_x;
  cfg.gridDimY = grid_y ? grid_y : 1;
  cfg.gridDimZ = 1;
  cfg.blockDimX = block_x;
  cfg.blockDimY = cfg.blockDimZ = 1;
  cfg.hStream = (CUstream)stream;
  CUresult r = driver().LaunchKernelEx(&cfg, k->func, args, nullptr);
  if (r != CUDA_SUCCESS) return fail(GX_ERR_CUDA, "cuLaunchKernelEx(jit kernel): " + cu_err(r));
  return GX_OK;
}

// A run-time compiled instance of gx_gemm_kernel.cuh with a generated epilogue (GX_GEMM_FUSED): same tiling and
// tensor maps as gx_gemm_bf16; `epi` is the kernel's GxEpi parameter block (pointers of the fused operands and
// outputs), copied by the launch.
int gx_gemm_fused_launch(gx_jit_kernel* k, int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb,
                         int b_mn, int ldn, const void* epi, int stages, void* stream) {
  float4* zero_ptr = nullptr;     // a pending gx_gemm_zero_with_next_launch belongs to THIS launch, also if it fails
  unsigned zero_n = 0;
  gxg_take_zero(&zero_ptr, &zero_n);
  if (!k || !A || !B || !epi) return fail(GX_ERR_BAD_ARG, "gx_gemm_fused_launch: NULL argument");
  if (M <= 0 || N <= 0 || K <= 0) return fail(GX_ERR_BAD_ARG, "gx_gemm_fused_launch: empty problem");
  if (ldn == 0) ldn = N;
  if (ldn % 4 || ldn < ((N + 127) / 128) * 128)
    return fail(GX_ERR_UNSUPPORTED, "gx_gemm_fused_launch: epilogue operands must be tile padded (ldn >= ceil(N/128)*128)");
  if (stages != 3 && stages != 6 && stages != 4 && stages != 8)
    return fail(GX_ERR_BAD_ARG, "gx_gemm_fused_launch: stages must be 3 / 6, or 4 / 8 for the CTA-pair kernels (GX_GEMM_STAGES)");
  const bool pair = gxg::stages_pair(stages);
  if (pair && ((M + 127) / 128) % 2)
    return fail(GX_ERR_UNSUPPORTED, "gx_gemm_fused_launch: the CTA-pair kernels need an even number of 128-row tiles");
  CUtensorMap map_a, map_b;
  if (gxg_make_maps(&map_a, &map_b, M, N, K, A, a_mn, B, b_mn, lda, ldb, pair ? 1 : 0))
    return fail(GX_ERR_CUDA, "gx_gemm_fused_launch: cuTensorMapEncodeTiled failed (pointers must be 16-byte aligned)");
  if (!k->smem_set) {
    CUresult r = driver().FuncSetAttribute(k->func, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)gxg::smem_bytes(stages));
    if (r != CUDA_SUCCESS) return fail(GX_ERR_CUDA, "cuFuncSetAttribute(fused gemm): " + cu_err(r));
    k->smem_set = true;
  }
  gxg::GemmArgs args;
  memset(&args, 0, sizeof args);
  args.M = M;
  args.N = N;
  args.K = K;
  args.ldc = ldn;
  args.k_per_split = ((K + 63) / 64) * 64;
  args.a_mn = a_mn ? 1 : 0;
  args.b_mn = b_mn ? 1 : 0;
  args.trace = gxg_next_trace(((N + 127) / 128) * ((M + 127) / 128));
  args.zero_ptr = zero_ptr;
  args.zero_n = zero_n;
  CUlaunchConfig cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDimX = (N + 127) / 128;
  cfg.gridDimY = (M + 127) / 128;
  cfg.gridDimZ = 1;
  cfg.blockDimX = gxg::kThreads;
  cfg.blockDimY = cfg.blockDimZ = 1;
  cfg.sharedMemBytes = gxg::smem_bytes(stages);
  cfg.hStream = (CUstream)stream;
  CUlaunchAttribute attr[2];
  unsigned n_attr = 0;
  if (use_pdl()) {
    attr[n_attr].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    attr[n_attr].value.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  if (pair) {   // two consecutive tile rows = one cluster = one 256 x 128 MMA tile, pairs along x (gx_gemm_kernel.cuh)
    cfg.gridDimX = (M + 127) / 128;
    cfg.gridDimY = (N + 127) / 128;
    attr[n_attr].id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
    attr[n_attr].value.clusterDim.x = 2;
    attr[n_attr].value.clusterDim.y = 1;
    attr[n_attr].value.clusterDim.z = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  void* params[] = {&map_a, &map_b, &args, const_cast<void*>(epi)};
  CUresult r = driver().LaunchKernelEx(&cfg, k->func, params, nullptr);
  if (r != CUDA_SUCCESS) return fail(GX_ERR_CUDA, "cuLaunchKernelEx(fused gemm): " + cu_err(r));
  return GX_OK;
}

// out[cols] += column sums of in[rows, cols] (fp32). The caller zeroes `out` first.
int gx_colsum_f32(const void* in, void* out, int rows, int cols, void* stream) {
  if (!in || !out || rows <= 0 || cols <= 0) return fail(GX_ERR_BAD_ARG, "gx_colsum_f32: bad argument");
  const int rows_per_cta = 256;
  dim3 grid((cols + 31) / 32, (rows + rows_per_cta - 1) / rows_per_cta), block(32, 8);
  gx_colsum_f32_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(in), static_cast<float*>(out), rows, cols, rows_per_cta);
  GX_CUDA(cudaGetLastError());
  return GX_OK;
}

}  // extern "C"
