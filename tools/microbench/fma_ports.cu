// Dispatch cost of fp32 instructions by operand form on sm_100a: is a 3-register FFMA twice as expensive as a
// 2-register FMUL / FADD or an FFMA whose multiplier stays in the operand reuse cache?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o fma_ports fma_ports.cu && ./fma_ports
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define ACC 16   // independent accumulators per thread: no dependency stalls with >= 2 warps per scheduler

template <int MODE>
__global__ void __launch_bounds__(1024, 1) kern(float* out, const float* in, long long* cycles) {
  float a[ACC], b[ACC], c[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) {
    a[i] = in[threadIdx.x + 32 * i];
    b[i] = in[threadIdx.x + 32 * i + 1000];
    c[i] = in[threadIdx.x + 32 * i + 2000];
  }
  const float s = in[threadIdx.x + 5000], t = in[threadIdx.x + 6000];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
      if (MODE >= 8) break;
      if (MODE == 0) a[i] = __fmaf_rn(a[i], b[i], c[i]);                 // 3 distinct registers
      if (MODE == 1) a[i] = __fmaf_rn(s, a[i], c[i]);                    // shared multiplier (reuse cache)
      if (MODE == 2) a[i] = __fmul_rn(a[i], b[i]);                       // 2 registers
      if (MODE == 3) a[i] = __fadd_rn(a[i], b[i]);                       // 2 registers
      if (MODE == 4) a[i] = __fmaf_rn(a[i], 1.0009765625f, c[i]);        // immediate multiplier
      if (MODE == 5) a[i] = __fmaf_rn(a[i], b[i], a[i]);                 // 2 distinct registers
      if (MODE == 6) a[i] = __fmaf_rn(s, a[i], t);                       // shared multiplier and addend
      if (MODE == 7) { a[i] = __fmul_rn(a[i], b[i]); c[i] = __fadd_rn(c[i], b[i]); }   // 2 x 2-register
    }
    if (MODE >= 8) {      // packed f32x2: ACC / 2 instructions, each two operations
#pragma unroll
      for (int i = 0; i < ACC; i += 2) {
        float2 pa = make_float2(a[i], a[i + 1]), pb = make_float2(b[i], b[i + 1]), pc = make_float2(c[i], c[i + 1]), r;
        if (MODE == 8) r = __ffma2_rn(pa, pb, pc);                          // three register pairs
        if (MODE == 9) r = __ffma2_rn(make_float2(s, s), pa, pc);           // broadcast multiplier, two pairs
        if (MODE == 10) r = __fmul2_rn(pa, pb);                             // two pairs
        if (MODE == 11) r = __fmul2_rn(make_float2(s, s), pa);              // broadcast, one pair
        if (MODE == 12) r = __ffma2_rn(make_float2(s, s), pa, make_float2(t, t));
        a[i] = r.x; a[i + 1] = r.y;
      }
    }
  }
  const long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) r += a[i] + c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp, float* out, float* in, long long* cyc) {
  const int threads = 128 * warps_per_smsp;
  kern<MODE><<<148, threads>>>(out, in, cyc);
  cudaDeviceSynchronize();
  kern<MODE><<<148, threads>>>(out, in, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double inst = (double)ITERS * ACC * (MODE == 7 ? 2 : 1) * warps_per_smsp;   // scalar operations per scheduler
  // (packed modes: ACC / 2 instructions of two operations each; the figure stays "cycles per operation")
  printf("%-44s warps/scheduler %d: %.3f cycles per warp instruction\n", name, warps_per_smsp, avg / inst);
}

int main() {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&in, 16384 * 4);
  cudaMalloc(&cyc, 148 * 8);
  cudaMemset(in, 0, 16384 * 4);
  for (int w : {2, 4, 8}) {
    run<0>("FFMA a=a*b+c (3 registers)", w, out, in, cyc);
    run<1>("FFMA a=s*a+c (shared multiplier)", w, out, in, cyc);
    run<6>("FFMA a=s*a+t (shared multiplier, addend)", w, out, in, cyc);
    run<5>("FFMA a=a*b+a (2 registers)", w, out, in, cyc);
    run<4>("FFMA a=a*imm+c", w, out, in, cyc);
    run<2>("FMUL a=a*b", w, out, in, cyc);
    run<3>("FADD a=a+b", w, out, in, cyc);
    run<7>("FMUL + FADD pair", w, out, in, cyc);
    run<8>("FFMA2 three pairs            [per op]", w, out, in, cyc);
    run<9>("FFMA2 s.F32 * pair + pair    [per op]", w, out, in, cyc);
    run<12>("FFMA2 s.F32 * pair + t.F32   [per op]", w, out, in, cyc);
    run<10>("FMUL2 pair * pair            [per op]", w, out, in, cyc);
    run<11>("FMUL2 s.F32 * pair           [per op]", w, out, in, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
