// XLA typed-FFI adaptor: exposes the hot path as a JAX custom call with no host round trip (north star: "the Python
// host keeps the repo's model/likelihood call signature and calls CUDA through a thin C-ABI exposed as a JAX FFI
// custom call").  Replaces the XLA:CPU executable numpyro builds around intensity_models.py:374-394,401 and its
// reverse pass: theta (F64[14] or F64[15] on the device) -> the flat result vector of include/bump.h
// (F64[BUMP_OUT_HEADER + nobs]: loglike, log_mu_sel, log_mu2, neff_sel, both gradients, neff[]).
//
// Compiled only where jaxlib's headers exist (`python -m bumpcosmology_b200._build --xla-ffi`, which adds
// -I$(python -c "import jaxlib, os; print(os.path.join(os.path.dirname(jaxlib.__file__), 'include'))")); this image
// has no JAX, so the adaptor is source-only here and bumpcosmology_b200/jax_ffi/__init__.py raises ImportError.
// The handler does nothing but forward to bump_eval_device: device pointers, XLA's stream, no allocation, no sync.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define BUMP_HAVE_XLA_FFI 1
#endif
#endif

#ifdef BUMP_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "../../include/bump.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error BumpLoglike(cudaStream_t stream, ffi::Buffer<ffi::F64> theta, int64_t ctx_handle,
                              ffi::ResultBuffer<ffi::F64> out) {
    bump_ctx* ctx = reinterpret_cast<bump_ctx*>(static_cast<intptr_t>(ctx_handle));
    if (!ctx) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bump_loglike: null context handle");
    const int64_t need = bump_out_len(ctx);
    if (static_cast<int64_t>(out->element_count()) < need || theta.element_count() < BUMP_NTHETA)
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "bump_loglike: theta needs 14 (15 with wa) entries and "
                                                            "the result BUMP_OUT_HEADER + nobs");
    if (bump_eval_device(ctx, theta.typed_data(), out->typed_data(), stream))
        return ffi::Error(ffi::ErrorCode::kInternal, bump_last_error());
    return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(BumpLoglikeFfi, BumpLoglike,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Attr<int64_t>("ctx")
                                  .Ret<ffi::Buffer<ffi::F64>>());
#else
// no XLA headers: keep the translation unit valid so that a plain compile of csrc/ does not fail
extern "C" int bump_xla_ffi_unavailable(void) { return 1; }
#endif
