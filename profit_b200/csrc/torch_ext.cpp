// PyTorch extension over the C ABI (include/gpb.h): torch.ops.profit_b200.*  -- the "thin C-ABI CUDA library through a
// PyTorch extension" host layer of the north star.  No numerics live here: every operator validates its tensors, names
// torch's CURRENT CUDA stream to the library (gpb_set_stream, so device tensors produced on any torch stream are ordered
// correctly without a host synchronisation) and forwards raw pointers to libgpb.so.
//
// Operators (handle = the int64 value of a gpb_handle_t):
//   create(device) -> handle                      gpb_create
//   destroy(handle)                               gpb_destroy
//   set_train(handle, X, y)                       gpb_set_train        (gp_functions.optimize's resident (X, y))
//   nll(handle, kind, theta, want_grad)           gpb_nll              (negative_log_likelihood_cholesky, gp_functions.py:103-187)
//   factorize(handle, kind, theta)                gpb_factorize        (GPSurrogate.Ky / alpha, custom_surrogate.py:24-45)
//   predict(handle, Xc, add_data_variance)        gpb_predict          (GPSurrogate.predict, custom_surrogate.py:95-120)
//   kernel_matrix(handle, kind, Xa, Xb, ell, sf, sn, eval_gradient)   gpb_kernel_matrix (python_kernels.RBF, :8-57)
//   append_train(handle, Xnew, ynew)              gpb_append_train     (custom_surrogate.py:122-134)
//   launch_count(handle) -> int                   gpb_launch_count
// theta / ell are float64 CPU tensors (the library reads them on the host); X, y, Xc, Xa, Xb may live on the host or on
// the handle's GPU; outputs are allocated on the device of the corresponding input.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <string>
#include <tuple>

#include "gpb.h"

namespace {

gpb_handle_t as_handle(int64_t v) {
  TORCH_CHECK(v != 0, "profit_b200: null handle");
  return reinterpret_cast<gpb_handle_t>(static_cast<intptr_t>(v));
}

void check_status(int status, gpb_handle_t h, const char* what) {
  if (status == 0) return;
  const std::string msg = std::string("profit_b200::") + what + ": " + gpb_error_string(h);
  // > 0: LAPACK-style info of a non-positive pivot -> torch.linalg.LinAlgError (np.linalg.cholesky raises there, gp_functions.py:143)
  TORCH_CHECK_LINALG(status < 0, msg, " (info=", status, ")");
  TORCH_CHECK_VALUE(status != -1, msg);
  TORCH_CHECK(false, msg, " (status ", status, ")");
}

const double* f64_ptr(const at::Tensor& t, const char* name) {
  TORCH_CHECK_VALUE(t.scalar_type() == at::kDouble, "profit_b200: ", name, " must be float64");
  TORCH_CHECK_VALUE(t.is_contiguous(), "profit_b200: ", name, " must be contiguous");
  TORCH_CHECK_VALUE(t.is_cpu() || t.is_cuda(), "profit_b200: ", name, " must be a CPU or CUDA tensor");
  return t.data_ptr<double>();
}

const double* host_f64_ptr(const at::Tensor& t, const char* name) {
  TORCH_CHECK_VALUE(t.is_cpu(), "profit_b200: ", name, " must be a CPU tensor (hyper-parameters are read on the host)");
  return f64_ptr(t, name);
}

// Order the library's stream after torch's current stream on the device of the first CUDA tensor among the arguments.
void name_stream(gpb_handle_t h, std::initializer_list<const at::Tensor*> tensors) {
  for (const at::Tensor* t : tensors)
    if (t && t->defined() && t->is_cuda()) {
      const auto stream = c10::cuda::getCurrentCUDAStream(t->device().index());
      gpb_set_stream(h, reinterpret_cast<void*>(stream.stream()));
      return;
    }
}

int64_t op_create(int64_t device) {
  gpb_handle_t h = nullptr;
  const int st = gpb_create(static_cast<int>(device), &h);
  TORCH_CHECK(st == 0 && h, "profit_b200::create(device=", device, ") failed with status ", st,
              ": a CUDA device is required (there is no CPU fallback)");
  return static_cast<int64_t>(reinterpret_cast<intptr_t>(h));
}

void op_destroy(int64_t handle) { gpb_destroy(as_handle(handle)); }

void op_set_train(int64_t handle, const at::Tensor& X, const at::Tensor& y) {
  gpb_handle_t h = as_handle(handle);
  TORCH_CHECK_VALUE(X.dim() == 2, "profit_b200::set_train: X should have shape (n, d)");
  TORCH_CHECK_VALUE(y.dim() == 1 || y.dim() == 2, "profit_b200::set_train: y should have shape (n,) or (n, D)");
  TORCH_CHECK_VALUE(y.size(0) == X.size(0), "profit_b200::set_train: mismatched number of data points for X and y");
  name_stream(h, {&X, &y});
  const int n_out = y.dim() == 1 ? 1 : static_cast<int>(y.size(1));
  check_status(gpb_set_train(h, f64_ptr(X, "X"), f64_ptr(y, "y"), X.size(0), static_cast<int>(X.size(1)), n_out), h,
               "set_train");
}

std::tuple<at::Tensor, at::Tensor> op_nll(int64_t handle, int64_t kind, const at::Tensor& theta, bool want_grad) {
  gpb_handle_t h = as_handle(handle);
  TORCH_CHECK_VALUE(theta.dim() == 1 && theta.numel() >= 3, "profit_b200::nll: theta = [length_scale..., sigma_f, sigma_n]");
  const double* th = host_f64_ptr(theta, "theta");
  const auto opts = at::TensorOptions().dtype(at::kDouble);
  at::Tensor nll = at::empty({}, opts);
  at::Tensor grad = want_grad ? at::empty({theta.numel()}, opts) : at::empty({0}, opts);
  check_status(gpb_nll(h, static_cast<int>(kind), th, static_cast<int>(theta.numel()) - 2, want_grad ? 1 : 0,
                       nll.data_ptr<double>(), want_grad ? grad.data_ptr<double>() : nullptr),
               h, "nll");
  return {nll, grad};
}

void op_factorize(int64_t handle, int64_t kind, const at::Tensor& theta) {
  gpb_handle_t h = as_handle(handle);
  TORCH_CHECK_VALUE(theta.dim() == 1 && theta.numel() >= 3, "profit_b200::factorize: theta = [length_scale..., sigma_f, sigma_n]");
  check_status(gpb_factorize(h, static_cast<int>(kind), host_f64_ptr(theta, "theta"), static_cast<int>(theta.numel()) - 2), h,
               "factorize");
}

std::tuple<at::Tensor, at::Tensor> op_predict(int64_t handle, const at::Tensor& Xc, int64_t add_data_variance, int64_t n_out) {
  gpb_handle_t h = as_handle(handle);
  TORCH_CHECK_VALUE(Xc.dim() == 2, "profit_b200::predict: Xpred should have shape (n, d)");
  name_stream(h, {&Xc});
  at::Tensor mean = at::empty({Xc.size(0), n_out}, Xc.options());
  at::Tensor var = at::empty({Xc.size(0)}, Xc.options());
  check_status(gpb_predict(h, f64_ptr(Xc, "Xpred"), Xc.size(0), static_cast<int>(add_data_variance), mean.data_ptr<double>(),
                           var.data_ptr<double>()),
               h, "predict");
  return {mean, var};
}

std::tuple<at::Tensor, at::Tensor> op_kernel_matrix(int64_t handle, int64_t kind, const at::Tensor& Xa, const at::Tensor& Xb,
                                                    const at::Tensor& ell, double sigma_f, double sigma_n,
                                                    bool eval_gradient) {
  gpb_handle_t h = as_handle(handle);
  TORCH_CHECK_VALUE(Xa.dim() == 2 && Xb.dim() == 2 && Xa.size(1) == Xb.size(1),
                    "profit_b200::kernel_matrix: X and Y should have shapes (n, d) and (m, d)");
  TORCH_CHECK_VALUE(Xa.device() == Xb.device(), "profit_b200::kernel_matrix: X and Y must live on the same device");
  TORCH_CHECK_VALUE(ell.dim() == 1 && (ell.numel() == 1 || ell.numel() == Xa.size(1)),
                    "profit_b200::kernel_matrix: one length scale, or one per input dimension");
  name_stream(h, {&Xa, &Xb});
  const int64_t P = ell.numel() + 2;
  at::Tensor K = at::empty({Xa.size(0), Xb.size(0)}, Xa.options());
  at::Tensor dK = eval_gradient ? at::empty({Xa.size(0), Xb.size(0), P}, Xa.options()) : at::empty({0}, Xa.options());
  if (Xa.size(0) == 0 || Xb.size(0) == 0) return {K, dK};
  check_status(gpb_kernel_matrix(h, static_cast<int>(kind), f64_ptr(Xa, "X"), Xa.size(0), f64_ptr(Xb, "Y"), Xb.size(0),
                                 static_cast<int>(Xa.size(1)), host_f64_ptr(ell, "length_scale"), static_cast<int>(ell.numel()),
                                 sigma_f, sigma_n, K.data_ptr<double>(), eval_gradient ? dK.data_ptr<double>() : nullptr),
               h, "kernel_matrix");
  return {K, dK};
}

void op_append_train(int64_t handle, const at::Tensor& Xnew, const at::Tensor& ynew) {
  gpb_handle_t h = as_handle(handle);
  TORCH_CHECK_VALUE(Xnew.dim() == 2 && ynew.dim() == 2 && Xnew.size(0) == ynew.size(0),
                    "profit_b200::append_train: X (m, d) and y (m, D)");
  name_stream(h, {&Xnew, &ynew});
  check_status(gpb_append_train(h, f64_ptr(Xnew, "X"), f64_ptr(ynew, "y"), Xnew.size(0)), h, "append_train");
}

int64_t op_launch_count(int64_t handle) { return gpb_launch_count(as_handle(handle)); }

}  // namespace

TORCH_LIBRARY(profit_b200, m) {
  m.def("create(int device) -> int", &op_create);
  m.def("destroy(int handle) -> ()", &op_destroy);
  m.def("set_train(int handle, Tensor X, Tensor y) -> ()", &op_set_train);
  m.def("nll(int handle, int kind, Tensor theta, bool want_grad) -> (Tensor, Tensor)", &op_nll);
  m.def("factorize(int handle, int kind, Tensor theta) -> ()", &op_factorize);
  m.def("predict(int handle, Tensor Xpred, int add_data_variance, int n_out) -> (Tensor, Tensor)", &op_predict);
  m.def("kernel_matrix(int handle, int kind, Tensor X, Tensor Y, Tensor length_scale, float sigma_f, float sigma_n, "
        "bool eval_gradient) -> (Tensor, Tensor)",
        &op_kernel_matrix);
  m.def("append_train(int handle, Tensor X, Tensor y) -> ()", &op_append_train);
  m.def("launch_count(int handle) -> int", &op_launch_count);
}
