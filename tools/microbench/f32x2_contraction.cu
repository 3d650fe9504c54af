// ptxas 12.9 (sm_100a, -O3) contracts mul.rn.f32x2 -> add.rn.f32x2 into FFMA2 although both instructions carry an
// explicit rounding modifier (the scalar mul.rn.f32 / add.rn.f32 pair is left alone) and although --fmad=false is
// given; it even duplicates the product when it has two consumers.  A packed product feeding a SCALAR sum, and
// scalar products feeding a packed sum, are not touched.  graphax_b200/slp.py therefore never lets a packed sum
// read the pair a packed product wrote.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o /tmp/c.cubin f32x2_contraction.cu
//   cuobjdump -sass /tmp/c.cubin | grep -E "Function|F(MUL|ADD|FMA)2?"
//     packed_packed : FFMA2                (wrong: one rounding instead of two)
//     packed_scalar : FMUL2, FADD, FADD
//     scalar_packed : FMUL, FMUL, FADD2
#include <cuda_runtime.h>

__global__ void packed_packed(const float2* a, const float2* b, const float2* c, float2* o) {
  int i = threadIdx.x;
  float2 p = __fmul2_rn(a[i], b[i]);
  o[i] = __fadd2_rn(p, c[i]);
}

__global__ void packed_scalar(const float2* a, const float2* b, const float2* c, float2* o) {
  int i = threadIdx.x;
  float2 p = __fmul2_rn(a[i], b[i]);
  float2 z = c[i];
  o[i] = make_float2(__fadd_rn(p.x, z.x), __fadd_rn(p.y, z.y));
}

__global__ void scalar_packed(const float2* a, const float2* b, const float2* c, float2* o) {
  int i = threadIdx.x;
  float2 x = a[i], y = b[i];
  float2 p = make_float2(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y));
  o[i] = __fadd2_rn(p, c[i]);
}
