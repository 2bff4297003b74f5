// ref_hsb_harness.cpp -- drives the HSB / InvHSB / InvHSBGrad OpKernels registered in the stub TF registry
// with raw pointers: the reference's (unmodified) ones, or -- with -DHARNESS_PREFIX_SHIM -- this project's
// replacement shim (polee_b200/tf/hsb_ops_b200.cpp), exported as shim_*.  TEST INFRASTRUCTURE ONLY.  Linked together with
// /root/reference/src/tensorflow_ext/hsb_ops.cpp into oracle/_ref/libref_hsb_ops.so.
#include <cstring>
#include <memory>
#include "tf_stub_core.h"

using namespace tensorflow;

// device == "CPU": host tensors, outputs copied back.  device == "GPU" (the shim's DEVICE_GPU kernels): the data tensors
// and out_ptrs are device memory the caller owns, the kernel writes the outputs in place and enqueues on `stream`.
static int run_op(const char* name, std::vector<const Tensor*> inputs, std::vector<DataType> out_types,
                  std::vector<void*> out_ptrs, std::vector<size_t> out_bytes, int threads, const char* device = "CPU",
                  void* stream = nullptr) {
  auto& reg = OpRegistryStub::Global();
  const bool gpu = std::string(device) == "GPU";
  auto it = reg.kernels.find(std::string(name) + "/" + device);
  if (it == reg.kernels.end()) return -1;
  OpKernelConstruction c;
  std::unique_ptr<OpKernel> k(it->second(&c));
  OpKernelContext ctx;
  ctx.inputs = inputs;
  ctx.output_types = out_types;
  ctx.dev.w_.num_threads = threads;
  ctx.dev.w_.workers = nullptr;
  if (gpu) {
    ctx.output_prealloc = out_ptrs;
    ctx.gpu_stream = stream;
  }
  k->Compute(&ctx);
  if (!ctx.status.ok()) return -2;
  if (!gpu)
    for (size_t i = 0; i < out_ptrs.size(); ++i) std::memcpy(out_ptrs[i], ctx.outputs[i]->raw(), out_bytes[i]);
  return 0;
}

#ifdef HARNESS_PREFIX_SHIM
#define ref_hsb shim_hsb
#define ref_inv_hsb shim_inv_hsb
#define ref_inv_hsb_grad shim_inv_hsb_grad
#endif

extern "C" {

int ref_hsb(int64_t B, int64_t n, const float* y_logit, const int32_t* left, const int32_t* right,
            const int32_t* leaf, float* x, int threads) {
  int64_t N = 2 * n - 1;
  Tensor a(DT_FLOAT, TensorShape({B, n - 1}), (void*)y_logit), l(DT_INT32, TensorShape({B, N}), (void*)left),
      r(DT_INT32, TensorShape({B, N}), (void*)right), f(DT_INT32, TensorShape({B, N}), (void*)leaf);
  return run_op("HSB", {&a, &l, &r, &f}, {DT_FLOAT}, {x}, {(size_t)(B * n) * 4}, threads);
}

int ref_inv_hsb(int64_t B, int64_t n, const float* x, const int32_t* left, const int32_t* right,
                const int32_t* leaf, double* y, float* ladj, int threads) {
  int64_t N = 2 * n - 1;
  Tensor a(DT_FLOAT, TensorShape({B, n}), (void*)x), l(DT_INT32, TensorShape({B, N}), (void*)left),
      r(DT_INT32, TensorShape({B, N}), (void*)right), f(DT_INT32, TensorShape({B, N}), (void*)leaf);
  return run_op("InvHSB", {&a, &l, &r, &f}, {DT_DOUBLE, DT_FLOAT}, {y, ladj},
                {(size_t)(B * (n - 1)) * 8, (size_t)B * 4}, threads);
}

int ref_inv_hsb_grad(int64_t B, int64_t n, const double* y_grad, const float* ladj_grad, const double* y,
                     const float* ladj, const int32_t* left, const int32_t* right, const int32_t* leaf,
                     float* backprops, int threads) {
  int64_t N = 2 * n - 1;
  Tensor a(DT_DOUBLE, TensorShape({B, n - 1}), (void*)y_grad), b(DT_FLOAT, TensorShape({B, 1}), (void*)ladj_grad),
      c(DT_DOUBLE, TensorShape({B, n - 1}), (void*)y), d(DT_FLOAT, TensorShape({B, 1}), (void*)ladj),
      l(DT_INT32, TensorShape({B, N}), (void*)left), r(DT_INT32, TensorShape({B, N}), (void*)right),
      f(DT_INT32, TensorShape({B, N}), (void*)leaf);
  return run_op("InvHSBGrad", {&a, &b, &c, &d, &l, &r, &f}, {DT_FLOAT}, {backprops}, {(size_t)(B * n) * 4}, threads);
}

#ifdef HARNESS_PREFIX_SHIM
// the shim's DEVICE_GPU registrations: y_logit / x / y / ... are DEVICE pointers, the index tensors stay on the host
// (HostMemory), outputs are written in place, work is enqueued on `stream`
int shim_hsb_gpu(int64_t B, int64_t n, const float* d_y_logit, const int32_t* left, const int32_t* right,
                 const int32_t* leaf, float* d_x, void* stream) {
  int64_t N = 2 * n - 1;
  Tensor a(DT_FLOAT, TensorShape({B, n - 1}), (void*)d_y_logit), l(DT_INT32, TensorShape({B, N}), (void*)left),
      r(DT_INT32, TensorShape({B, N}), (void*)right), f(DT_INT32, TensorShape({B, N}), (void*)leaf);
  return run_op("HSB", {&a, &l, &r, &f}, {DT_FLOAT}, {d_x}, {0}, 1, "GPU", stream);
}

int shim_inv_hsb_gpu(int64_t B, int64_t n, const float* d_x, const int32_t* left, const int32_t* right,
                     const int32_t* leaf, double* d_y, float* d_ladj, void* stream) {
  int64_t N = 2 * n - 1;
  Tensor a(DT_FLOAT, TensorShape({B, n}), (void*)d_x), l(DT_INT32, TensorShape({B, N}), (void*)left),
      r(DT_INT32, TensorShape({B, N}), (void*)right), f(DT_INT32, TensorShape({B, N}), (void*)leaf);
  return run_op("InvHSB", {&a, &l, &r, &f}, {DT_DOUBLE, DT_FLOAT}, {d_y, d_ladj}, {0, 0}, 1, "GPU", stream);
}

int shim_inv_hsb_grad_gpu(int64_t B, int64_t n, const double* d_y_grad, const float* d_ladj_grad, const double* d_y,
                          const float* d_ladj, const int32_t* left, const int32_t* right, const int32_t* leaf,
                          float* d_backprops, void* stream) {
  int64_t N = 2 * n - 1;
  Tensor a(DT_DOUBLE, TensorShape({B, n - 1}), (void*)d_y_grad), b(DT_FLOAT, TensorShape({B, 1}), (void*)d_ladj_grad),
      c(DT_DOUBLE, TensorShape({B, n - 1}), (void*)d_y), d(DT_FLOAT, TensorShape({B, 1}), (void*)d_ladj),
      l(DT_INT32, TensorShape({B, N}), (void*)left), r(DT_INT32, TensorShape({B, N}), (void*)right),
      f(DT_INT32, TensorShape({B, N}), (void*)leaf);
  return run_op("InvHSBGrad", {&a, &b, &c, &d, &l, &r, &f}, {DT_FLOAT}, {d_backprops}, {0}, 1, "GPU", stream);
}
#endif
}
