// XLA FFI handlers over the C-ABI of include/bobe_b200.h, so that jit'd JAX callers of the reference
// (BOBE/samplers.py:112-115, BOBE/acquisition.py:390-394, BOBE/optim.py:118,211,309) can stay unchanged.
// NOT BUILT IN THIS IMAGE: neither JAX nor the XLA FFI headers are installed, so this file
// compiles to nothing unless "xla/ffi/api/ffi.h" is on the include path.  Build (where JAX exists):
//   g++ -O2 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -I../../include \
//       xla_ffi_shim.cc -L../lib -lbobe_b200 -lcudart -o ../lib/libbobe_xla_ffi.so
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include "bobe_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RI32 = ffi::ResultBuffer<ffi::S32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

static ffi::Error status(int32_t rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(bobe_last_error_string());
}

static ffi::Error KernelMatrixImpl(cudaStream_t s, F64 xa, F64 xb, F64 ls, int64_t kind, double kv, double noise,
                                   int64_t add_noise, RF64 out) {
  int64_t n1 = xa.dimensions()[0], d = xa.dimensions()[1], n2 = xb.dimensions()[0];
  return status(bobe_kernel_matrix(s, (int32_t)kind, xa.typed_data(), n1, xb.typed_data(), n2, d, ls.typed_data(), kv,
                                   noise, (int32_t)add_noise, out->typed_data(), n2));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeKernelMatrix, KernelMatrixImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Attr<int64_t>("kind").Attr<double>("kv").Attr<double>("noise")
                                  .Attr<int64_t>("add_noise").Ret<F64>());

static ffi::Error FactorizeImpl(cudaStream_t s, F64 X, F64 y, F64 ls, F64 kv, int64_t kind, double noise, RF64 L,
                                RF64 Linv, RF64 alpha, RF64 logdet, RF64 quad, RI32 info, RU8 ws) {
  int64_t n = X.dimensions()[0], d = X.dimensions()[1], batch = ls.dimensions()[0];
  return status(bobe_factorize(s, (int32_t)kind, X.typed_data(), y.typed_data(), n, d, ls.typed_data(), kv.typed_data(),
                               noise, batch, L->typed_data(), Linv->typed_data(), alpha->typed_data(),
                               logdet->typed_data(), quad->typed_data(), info->typed_data(), ws->untyped_data(),
                               (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeFactorize, FactorizeImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Attr<int64_t>("kind").Attr<double>("noise").Ret<F64>().Ret<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

static ffi::Error MllGradImpl(cudaStream_t s, F64 X, F64 y, F64 log_params, int64_t kind, int64_t has_kv,
                              double fixed_kv, double noise, RF64 val, RF64 grad, RI32 info, RU8 ws) {
  int64_t n = X.dimensions()[0], d = X.dimensions()[1], R = log_params.dimensions()[0], P = log_params.dimensions()[1];
  return status(bobe_mll_grad_batched(s, (int32_t)kind, X.typed_data(), y.typed_data(), n, d, log_params.typed_data(), R,
                                      P, (int32_t)has_kv, fixed_kv, noise, val->typed_data(), grad->typed_data(),
                                      info->typed_data(), ws->untyped_data(), (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeMllGrad, MllGradImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Attr<int64_t>("kind").Attr<int64_t>("has_kv").Attr<double>("fixed_kv")
                                  .Attr<double>("noise").Ret<F64>().Ret<F64>().Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

static ffi::Error PredictImpl(cudaStream_t s, F64 X, F64 ls, F64 Linv, F64 alpha, F64 Xq, int64_t kind, double kv,
                              double noise, double y_mean, double y_std, int64_t mode, RF64 mean, RF64 var, RU8 ws) {
  int64_t n = X.dimensions()[0], d = X.dimensions()[1], M = Xq.dimensions()[0];
  return status(bobe_predict(s, (int32_t)kind, X.typed_data(), n, d, ls.typed_data(), kv, noise, Linv.typed_data(),
                             alpha.typed_data(), Xq.typed_data(), M, y_mean, y_std, (int32_t)mode, mean->typed_data(),
                             var->typed_data(), ws->untyped_data(), (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobePredict, PredictImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Attr<int64_t>("kind").Attr<double>("kv").Attr<double>("noise")
                                  .Attr<double>("y_mean").Attr<double>("y_std").Attr<int64_t>("mode").Ret<F64>()
                                  .Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());

static ffi::Error FantasyImpl(cudaStream_t s, F64 X, F64 ls, F64 Linv, F64 Xmc, F64 Xcand, int64_t kind, double kv,
                              double noise, double y_std, int64_t reduce, RF64 out, RU8 ws) {
  int64_t n = X.dimensions()[0], d = X.dimensions()[1], n_mc = Xmc.dimensions()[0], C = Xcand.dimensions()[0];
  return status(bobe_fantasy_var(s, (int32_t)kind, X.typed_data(), n, d, ls.typed_data(), kv, noise, Linv.typed_data(),
                                 y_std, Xmc.typed_data(), n_mc, Xcand.typed_data(), C, (int32_t)reduce,
                                 out->typed_data(), ws->untyped_data(), (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeFantasyVar, FantasyImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Attr<int64_t>("kind").Attr<double>("kv").Attr<double>("noise")
                                  .Attr<double>("y_std").Attr<int64_t>("reduce").Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// WIPV / WIPStd value + gradient in the candidate: the custom_vjp rule of the n <= 500 polish (BOBE/acquisition.py:400-412)
static ffi::Error FantasyGradImpl(cudaStream_t s, F64 X, F64 ls, F64 Linv, F64 LinvT, F64 Xmc, F64 Xcand, int64_t kind,
                                  double kv, double noise, double y_std, int64_t reduce, RF64 out, RF64 dout, RU8 ws) {
  int64_t n = X.dimensions()[0], d = X.dimensions()[1], n_mc = Xmc.dimensions()[0], C = Xcand.dimensions()[0];
  return status(bobe_fantasy_var_grad(s, (int32_t)kind, X.typed_data(), n, d, ls.typed_data(), kv, noise,
                                      Linv.typed_data(), LinvT.typed_data(), y_std, Xmc.typed_data(), n_mc,
                                      Xcand.typed_data(), C, (int32_t)reduce, out->typed_data(), dout->typed_data(),
                                      ws->untyped_data(), (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeFantasyVarGrad, FantasyGradImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Attr<int64_t>("kind").Attr<double>("kv")
                                  .Attr<double>("noise").Attr<double>("y_std").Attr<int64_t>("reduce").Ret<F64>()
                                  .Ret<F64>().Ret<ffi::Buffer<ffi::U8>>());

// value + input gradients: the backward rule of the custom_vjp around bobe_predict (BOBE/samplers.py:268-285,
// BOBE/optim.py:118,309 differentiate predict_*_single with respect to x)
static ffi::Error PredictGradImpl(cudaStream_t s, F64 X, F64 ls, F64 Linv, F64 LinvT, F64 alpha, F64 Xq, int64_t kind,
                                  double kv, double noise, double y_mean, double y_std, int64_t mode, RF64 mean, RF64 var,
                                  RF64 dmean, RF64 dvar, RU8 ws) {
  int64_t n = X.dimensions()[0], d = X.dimensions()[1], M = Xq.dimensions()[0];
  return status(bobe_predict_grad(s, (int32_t)kind, X.typed_data(), n, d, ls.typed_data(), kv, noise, Linv.typed_data(),
                                  LinvT.typed_data(), alpha.typed_data(), Xq.typed_data(), M, y_mean, y_std,
                                  (int32_t)mode, mean->typed_data(), var->typed_data(), dmean->typed_data(),
                                  dvar->typed_data(), ws->untyped_data(), (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobePredictGrad, PredictGradImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Attr<int64_t>("kind").Attr<double>("kv")
                                  .Attr<double>("noise").Attr<double>("y_mean").Attr<double>("y_std")
                                  .Attr<int64_t>("mode").Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

static ffi::Error LinvTransposeImpl(cudaStream_t s, F64 Linv, int64_t n, RF64 LinvT) {
  return status(bobe_linv_transpose(s, Linv.typed_data(), n, LinvT->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeLinvTranspose, LinvTransposeImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Attr<int64_t>("n")
                                  .Ret<F64>());

// GP.update without the rebuild: L / Linv are donated (input_output_aliases on the Python side) and extended in place
static ffi::Error FactorAppendImpl(cudaStream_t s, F64 X, F64 y, F64 ls, F64 L_in, F64 Linv_in, int64_t kind,
                                   int64_t n_old, double kv, double noise, RF64 L, RF64 Linv, RF64 alpha, RI32 info,
                                   RU8 ws) {
  int64_t n = X.dimensions()[0], d = X.dimensions()[1];
  (void)L_in; (void)Linv_in;  // aliased to the results
  return status(bobe_factor_append(s, (int32_t)kind, X.typed_data(), y.typed_data(), n_old, n - n_old, d,
                                   ls.typed_data(), kv, noise, L->typed_data(), Linv->typed_data(), alpha->typed_data(),
                                   info->typed_data(), ws->untyped_data(), (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeFactorAppend, FactorAppendImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Attr<int64_t>("kind").Attr<int64_t>("n_old").Attr<double>("kv")
                                  .Attr<double>("noise").Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U8>>());

// GPwithClassifier mask (BOBE/clf_gp.py:173-205): mean / var are donated and masked in place
static ffi::Error SvmMaskImpl(cudaStream_t s, F64 sv, F64 dual, F64 Xq, F64 mean_in, F64 var_in, double intercept,
                              double gamma, double minus_inf, double var_fill, RF64 mean, RF64 var, RF64 decision,
                              RU8 ws) {
  int64_t n_sv = sv.dimensions()[0], d = sv.dimensions()[1], M = Xq.dimensions()[0];
  (void)mean_in; (void)var_in;
  return status(bobe_svm_mask(s, sv.typed_data(), n_sv, d, dual.typed_data(), intercept, gamma, Xq.typed_data(), M,
                              minus_inf, var_fill, mean->typed_data(), var->typed_data(), decision->typed_data(),
                              ws->untyped_data(), (int64_t)ws->size_bytes()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(BobeSvmMask, SvmMaskImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Attr<double>("intercept").Attr<double>("gamma")
                                  .Attr<double>("minus_inf").Attr<double>("var_fill").Ret<F64>().Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::U8>>());
#endif
#endif
