// Host-visible constants of gx_gemm_kernel.cuh for translation units that launch run-time compiled instances of
// it without compiling the kernel themselves.
#pragma once
#include <cuda.h>
#include <stdint.h>
namespace gxg {
constexpr int kThreads = 320;
constexpr uint32_t smem_bytes(int stages) { return stages * (128 * 64 * 2 + 128 * 64 * 2) + 1024 + 256; }
struct GemmArgs {
  float* c_f32;
  uint16_t* c_bf16;
  int M, N, K, ldc, epilogue, k_per_split;
  int a_mn, b_mn;
};
}  // namespace gxg
int gxg_make_maps(CUtensorMap* map_a, CUtensorMap* map_b, int M, int N, int K, const void* A, int a_mn, const void* B,
                  int b_mn, int lda, int ldb);
