// XLA-FFI shim over the C ABI of include/nbm_b200.h: what a JAX-DIPS maintainer compiles to replace
// `value_and_grad(self.loss)` (jax_dips/solvers/poisson/trainer.py:786, 826, 893) by the CUDA step.
//
// jaxlib (and with it xla/ffi/api/ffi.h) is not installed in this repository's image: `__graft_entry__.build()` type-checks
// this file against a labelled STUB of that header (integration/stub/xla/ffi/api/ffi.h: handler signature vs binding, every
// nbm_* call vs include/nbm_b200.h).  Build where jax is:
//   g++ -shared -fPIC -std=c++17 -I$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))") \
//       -I../include -I/usr/local/cuda/include nbm_xla_ffi.cc -L../jax_dips_b200 -lnbm_b200 -o libnbm_xla_ffi.so
#include <cuda_runtime_api.h>
#include "xla/ffi/api/ffi.h"
#include "nbm_b200.h"
namespace ffi = xla::ffi;

// `plan` is the address of a caller-owned nbm_shared_step_t (built once per level through the calls of section 1).
// The descriptor is COPIED: the caller's plan is never written, so one plan may be used by concurrent calls
// (pmap runs one host thread per device; each device owns its plan and its output buffer).
static ffi::Error LossGradImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, int64_t plan,
                               ffi::ResultBuffer<ffi::F32> loss_grad) {
  nbm_shared_step_t s = *reinterpret_cast<const nbm_shared_step_t*>(plan);
  s.loss_grad = loss_grad->typed_data();
  if (nbm_upload_params(&s.net, params.typed_data(), reinterpret_cast<nbm_stream_t>(stream)))
    return ffi::Error(ffi::ErrorCode::kInternal, nbm_last_error());
  if (nbm_loss_grad_shared_f32(&s, reinterpret_cast<nbm_stream_t>(stream)))
    return ffi::Error(ffi::ErrorCode::kInternal, nbm_last_error());
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(NbmLossGrad, LossGradImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F32>>().Attr<int64_t>("plan").Ret<ffi::Buffer<ffi::F32>>());
