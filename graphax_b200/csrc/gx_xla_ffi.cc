// gx_xla_ffi.cc - XLA FFI handler for the jacve launch (the `jax.ffi` boundary named by BASELINE.json).
//
// NOT built by build(): it needs xla/ffi/api/ffi.h, which ships with jaxlib (jax.ffi.include_dir()) and is not in
// this image.  graphax_b200/jax_ffi.py compiles it on first use on a machine that has JAX:
//   g++ -std=c++17 -O2 -fPIC -shared -I<jax.ffi.include_dir()> -I<cuda>/include -Iinclude \
//       graphax_b200/csrc/gx_xla_ffi.cc -Lgraphax_b200 -lgraphax_b200 -o graphax_b200/libgraphax_b200_xla.so
// EXPERIMENTAL - never compiled or run (no JAX, no XLA headers, in this image): written against the public FFI API
// only.  Do not rely on it before it has been exercised with JAX on a GPU.
//
// Operands: the function's arguments, each f32[batch, size_i] on the device.
// Result  : f32[..., n_rows, n_cols], the packed Jacobian tile (include/graphax_b200.h, gx_jacve_launch); every
//           leading axis is a batch axis (vmap_method="broadcast_all" prepends one per nested vmap, on the operands
//           as on the result), so the batch is the product of all but the last two dimensions.
// Attr    : plan = the gx_plan* of the elimination plan as int64 (immutable after creation, one per device).
#include <cuda_runtime_api.h>

#include <cstdint>

#include "graphax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error GxJacveImpl(cudaStream_t stream, int64_t plan_handle, ffi::RemainingArgs args,
                              ffi::Result<ffi::Buffer<ffi::F32>> jac) {
  auto* plan = reinterpret_cast<gx_plan*>(static_cast<intptr_t>(plan_handle));
  if (plan == nullptr) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "gx_jacve: null plan handle");
  gx_info info;
  if (gx_plan_info(plan, &info) != GX_OK) return ffi::Error(ffi::ErrorCode::kInternal, gx_last_error());
  // the operand list must be exactly the plan's input arrays: gx_jacve_launch reads in[0 .. n_in_arrays)
  if (args.size() != info.n_in_arrays)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "gx_jacve: operand count differs from the plan's input arrays");
  const auto dims = jac->dimensions();
  if (dims.size() < 2 || dims[dims.size() - 2] != (int64_t)info.n_rows || dims[dims.size() - 1] != (int64_t)info.n_cols)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "gx_jacve: result must be [..., n_rows, n_cols]");
  int64_t batch = 1;                       // all leading axes (one per enclosing vmap) fold into the batch
  for (size_t d = 0; d + 2 < dims.size(); ++d) batch *= dims[d];
  const void* in[GX_MAX_ARRAYS];
  for (size_t i = 0; i < args.size(); ++i) {
    auto buf = args.get<ffi::Buffer<ffi::F32>>(i);
    if (buf.has_error()) return buf.error();
    if ((int64_t)buf->element_count() != batch * (int64_t)info.in_sizes[i])
      return ffi::Error(ffi::ErrorCode::kInvalidArgument, "gx_jacve: operand size differs from batch x the plan's input size");
    in[i] = buf->untyped_data();
  }
  // asynchronous on XLA's stream; the plan never allocates or synchronises in launch
  if (gx_jacve_launch(plan, batch, in, jac->untyped_data(), /*primal_out=*/nullptr, stream) != GX_OK)
    return ffi::Error(ffi::ErrorCode::kInternal, gx_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(GxJacve, GxJacveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .RemainingArgs()
                                  .Ret<ffi::Buffer<ffi::F32>>());
