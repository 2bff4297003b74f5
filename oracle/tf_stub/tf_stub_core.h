// tf_stub_core.h -- the whole stand-in TensorFlow op API used to build the reference's
// hsb_ops.cpp (and to compile-check this project's replacement shim) without TensorFlow.
// TEST INFRASTRUCTURE ONLY; see tensorflow/core/framework/op.h in this directory.
#ifndef POLEE_TF_STUB_CORE_H
#define POLEE_TF_STUB_CORE_H

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace tensorflow {

typedef std::int64_t int64;
typedef std::int32_t int32;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }

 private:
  bool ok_;
  std::string msg_;
};

#define TF_RETURN_IF_ERROR(expr)            \
  do {                                      \
    ::tensorflow::Status _s = (expr);       \
    if (!_s.ok()) return _s;                \
  } while (0)

class mutex {
 public:
  void lock() { mu_.lock(); }
  void unlock() { mu_.unlock(); }

 private:
  std::mutex mu_;
};

enum DataType { DT_FLOAT, DT_DOUBLE, DT_INT32 };
template <typename T> struct DataTypeOf;
template <> struct DataTypeOf<float> { static const DataType v = DT_FLOAT; };
template <> struct DataTypeOf<double> { static const DataType v = DT_DOUBLE; };
template <> struct DataTypeOf<int32> { static const DataType v = DT_INT32; };

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : dims_(d) {}
  int dims() const { return (int)dims_.size(); }
  int64 dim_size(int i) const { return dims_[i]; }
  int64 num_elements() const {
    int64 c = 1;
    for (auto d : dims_) c *= d;
    return c;
  }

 private:
  std::vector<int64> dims_;
};

// 2-D row-major view, as returned by Tensor::flat_inner_dims<T>() for rank-2 tensors.
template <typename T>
class Matrix2D {
 public:
  Matrix2D(T* p, int64 rows, int64 cols) : p_(p), rows_(rows), cols_(cols) {}
  T& operator()(int64 i, int64 j) const { return p_[i * cols_ + j]; }

 private:
  T* p_;
  int64 rows_, cols_;
};

class Tensor {
 public:
  Tensor() : data_(nullptr), owned_(false) {}
  // non-owning view over caller memory
  Tensor(DataType dt, const TensorShape& s, void* data) : dt_(dt), shape_(s), data_(data), owned_(false) {}
  // owning
  Tensor(DataType dt, const TensorShape& s) : dt_(dt), shape_(s), owned_(true) {
    size_t esz = dt == DT_DOUBLE ? 8 : 4;
    data_ = std::calloc((size_t)s.num_elements() + 1, esz);
  }
  ~Tensor() {
    if (owned_) std::free(data_);
  }
  Tensor(const Tensor&) = delete;
  Tensor& operator=(const Tensor&) = delete;
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  const TensorShape& shape() const { return shape_; }
  void* raw() const { return data_; }
  template <typename T>
  Matrix2D<T> flat_inner_dims() {
    return Matrix2D<T>((T*)data_, shape_.dim_size(0), shape_.dim_size(1));
  }
  template <typename T>
  Matrix2D<const T> flat_inner_dims() const {
    return Matrix2D<const T>((const T*)data_, shape_.dim_size(0), shape_.dim_size(1));
  }

 private:
  DataType dt_;
  TensorShape shape_;
  void* data_;
  bool owned_;
};

namespace thread {
class ThreadPool {};
}  // namespace thread

struct CpuWorkerThreads {
  int num_threads;
  thread::ThreadPool* workers;
};

class DeviceBase {
 public:
  const CpuWorkerThreads* tensorflow_cpu_worker_threads() const { return &w_; }
  CpuWorkerThreads w_;
};

class OpKernelConstruction {};

class OpKernelContext {
 public:
  std::vector<const Tensor*> inputs;
  std::vector<std::unique_ptr<Tensor>> outputs;
  std::vector<DataType> output_types;
  DeviceBase dev;
  Status status;
  // GPU-registered kernels (DEVICE_GPU): the harness supplies the output buffers (device memory) and the stream
  std::vector<void*> output_prealloc;
  void* gpu_stream = nullptr;
  const Tensor& input(int i) { return *inputs[i]; }
  Status allocate_output(int i, const TensorShape& s, Tensor** out) {
    if ((int)outputs.size() <= i) outputs.resize(i + 1);
    if (i < (int)output_prealloc.size() && output_prealloc[i])
      outputs[i].reset(new Tensor(output_types[i], s, output_prealloc[i]));
    else
      outputs[i].reset(new Tensor(output_types[i], s));
    *out = outputs[i].get();
    return Status::OK();
  }
  DeviceBase* device() { return &dev; }
  void SetStatus(const Status& s) { status = s; }
};

#define OP_REQUIRES_OK(CTX, ...)                 \
  do {                                           \
    ::tensorflow::Status _s(__VA_ARGS__);        \
    if (!_s.ok()) {                              \
      (CTX)->SetStatus(_s);                      \
      return;                                    \
    }                                            \
  } while (0)

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

// Shard(): split [0,total) into contiguous blocks, one std::thread per block.
inline void Shard(int max_parallelism, thread::ThreadPool*, int64 total, int64 /*cost_per_unit*/,
                  std::function<void(int64, int64)> work) {
  if (total <= 0) return;
  int nt = max_parallelism < 1 ? 1 : max_parallelism;
  if ((int64)nt > total) nt = (int)total;
  if (nt == 1) {
    work(0, total);
    return;
  }
  std::vector<std::thread> ts;
  int64 per = (total + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    int64 lo = t * per, hi = lo + per > total ? total : lo + per;
    if (lo >= hi) break;
    ts.emplace_back([=] { work(lo, hi); });
  }
  for (auto& t : ts) t.join();
}

// ---- shape inference (compiled, never executed by the harness) ----
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle { int64 v; };
struct DimensionOrConstant {
  DimensionOrConstant(DimensionHandle d) : v(d.v) {}  // NOLINT
  DimensionOrConstant(int64 c) : v(c) {}              // NOLINT
  DimensionOrConstant(int c) : v(c) {}                // NOLINT
  int64 v;
};
class InferenceContext {
 public:
  ShapeHandle input(int) { return ShapeHandle(); }
  Status WithRank(ShapeHandle, int, ShapeHandle*) { return Status::OK(); }
  DimensionHandle Dim(ShapeHandle, int) { return DimensionHandle{0}; }
  int64 Value(DimensionHandle d) { return d.v; }
  DimensionHandle MakeDim(int64 v) { return DimensionHandle{v}; }
  ShapeHandle MakeShape(std::initializer_list<DimensionOrConstant>) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
};
}  // namespace shape_inference

// ---- op / kernel registration ----
class OpDefBuilderWrapper {
 public:
  explicit OpDefBuilderWrapper(const char* name) : name_(name) {}
  OpDefBuilderWrapper& Input(const char* s) { inputs_.push_back(s); return *this; }
  OpDefBuilderWrapper& Output(const char* s) { outputs_.push_back(s); return *this; }
  OpDefBuilderWrapper& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)> f) {
    shape_fn_ = f;
    return *this;
  }
  std::string name_;
  std::vector<std::string> inputs_, outputs_;
  std::function<Status(shape_inference::InferenceContext*)> shape_fn_;
};

struct OpRegistryStub {
  std::map<std::string, OpDefBuilderWrapper> ops;
  std::map<std::string, std::function<OpKernel*(OpKernelConstruction*)>> kernels;
  static OpRegistryStub& Global() {
    static OpRegistryStub r;
    return r;
  }
};

struct OpDefBuilderReceiver {
  OpDefBuilderReceiver(const OpDefBuilderWrapper& w) {  // NOLINT (implicit on purpose)
    OpRegistryStub::Global().ops.insert(std::make_pair(w.name_, w));
  }
};

#define TF_STUB_CAT_(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT_(a, b)
#define REGISTER_OP(name)                                                    \
  static ::tensorflow::OpDefBuilderReceiver TF_STUB_CAT(register_op_, __COUNTER__) = \
      ::tensorflow::OpDefBuilderWrapper(name)

static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

class KernelDefBuilder {
 public:
  explicit KernelDefBuilder(const char* op) : op_(op) {}
  KernelDefBuilder& Device(const char* d) { device_ = d; return *this; }
  KernelDefBuilder& HostMemory(const char* arg) { host_args_.push_back(arg); return *this; }
  std::string op_, device_;
  std::vector<std::string> host_args_;
};
inline KernelDefBuilder Name(const char* op) { return KernelDefBuilder(op); }

struct KernelRegistrar {
  KernelRegistrar(const KernelDefBuilder& b, std::function<OpKernel*(OpKernelConstruction*)> f) {
    OpRegistryStub::Global().kernels[b.op_ + "/" + b.device_] = f;
  }
};

#define REGISTER_KERNEL_BUILDER(builder, ...)                                          \
  static ::tensorflow::KernelRegistrar TF_STUB_CAT(register_kernel_, __COUNTER__)(     \
      ::tensorflow::builder,                                                           \
      [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow
#endif
