// Host-visible constants of gx_gemm_kernel.cuh for translation units that launch run-time compiled instances of
// it without compiling the kernel themselves.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
namespace gxg {
constexpr int kThreads = 320;
// 3 / 6 stages: one CTA per 128 x 128 tile; 4 / 8 stages: the CTA-pair kernels (half of the B tile per CTA)
constexpr bool stages_pair(int stages) { return stages == 4 || stages == 8; }
constexpr uint32_t smem_bytes(int stages) {
  return stages * (128 * 64 * 2 + (stages_pair(stages) ? 64 : 128) * 64 * 2) + 1024 + 256;
}
struct GemmArgs {
  float* c_f32;
  uint16_t* c_bf16;
  int M, N, K, ldc, epilogue, k_per_split;
  int a_mn, b_mn;
  unsigned long long* trace;
  float4* zero_ptr;
  unsigned zero_n;
};
}  // namespace gxg
int gxg_make_maps(CUtensorMap* map_a, CUtensorMap* map_b, int M, int N, int K, const void* A, int a_mn, const void* B,
                  int b_mn, int lda, int ldb, int pair);
int gxg_pair_enabled(void);
// next slice of the trace buffer (gx_gemm_set_trace), or null
unsigned long long* gxg_next_trace(int n_ctas);
// region the next single-tile launch of this thread has been asked to clear (gx_gemm_zero_with_next_launch); one-shot
void gxg_take_zero(float4** ptr, unsigned* n16);
