// TEST INFRASTRUCTURE ONLY -- a ~150-line stand-in for the TensorFlow 1.0 kernel API, written for ONE purpose: to
// compile the reference's own lib/roi_pooling_layer/roi_pooling_op.cc UNMODIFIED (from where it lies under
// /root/reference) into oracle/_ref/libref_roi_pool.so, so that the C restatement of RoiPool / RoiPoolGrad in
// oracle_c.c is pinned against the reference's real loop bodies (roi_pooling_op.cc:123-182, :369-444).
// It implements exactly the accessors that file uses: Tensor::flat<T>() / dims() / dim_size() / shape(), TensorShape,
// TensorShapeUtils::MakeShape, OpKernelConstruction::GetAttr, OpKernelContext::input / allocate_output / device /
// status / eigen_device, OP_REQUIRES(_OK), errors::InvalidArgument, DeviceBase::CpuWorkerThreads, and
// REGISTER_KERNEL_BUILDER (a tiny factory registry keyed by "<op>/<device>").  Not TensorFlow, not a product path.
#pragma once
#include <cmath>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "tensorflow/core/lib/core/threadpool.h"
#include "tensorflow/core/platform/types.h"
#include "third_party/eigen3/unsupported/Eigen/CXX11/Tensor"

namespace tensorflow {

class Status {
 public:
    Status() : ok_(true) {}
    explicit Status(const std::string& m) : ok_(false), msg_(m) {}
    bool ok() const { return ok_; }
    const std::string& error_message() const { return msg_; }
    static Status OK() { return Status(); }
 private:
    bool ok_;
    std::string msg_;
};

namespace errors {
template <typename... Args>
Status InvalidArgument(const Args&... args) {
    std::ostringstream os;
    int dummy[] = {0, ((os << args), 0)...};
    (void)dummy;
    return Status(os.str());
}
}  // namespace errors

class TensorShape {
 public:
    TensorShape() {}
    explicit TensorShape(const std::vector<int64>& d) : dims_(d) {}
    int dims() const { return (int)dims_.size(); }
    int64 dim_size(int i) const { return dims_[i]; }
    int64 num_elements() const { int64 n = 1; for (int64 d : dims_) n *= d; return n; }
    std::vector<int64> dims_;
};

struct TensorShapeUtils {
    static Status MakeShape(const int* dims, int n, TensorShape* out) {
        out->dims_.assign(dims, dims + n);
        return Status::OK();
    }
};

template <typename T>
class Flat {
 public:
    Flat(T* p, int64 n) : p_(p), n_(n) {}
    T* data() const { return p_; }
    int64 size() const { return n_; }
    T& operator()(int64 i) const { return p_[i]; }
 private:
    T* p_;
    int64 n_;
};

// A tensor is a typed view of caller memory (inputs) or of a buffer it owns (outputs).
class Tensor {
 public:
    Tensor(const TensorShape& s, void* ext) : shape_(s), ext_(ext) {}
    Tensor(const TensorShape& s, size_t elem_bytes) : shape_(s), own_(s.num_elements() * elem_bytes + 16), ext_(nullptr) {}
    int dims() const { return shape_.dims(); }
    int64 dim_size(int i) const { return shape_.dim_size(i); }
    const TensorShape& shape() const { return shape_; }
    template <typename T>
    Flat<T> flat() const { return Flat<T>(reinterpret_cast<T*>(raw()), shape_.num_elements()); }
    void* raw() const { return ext_ ? ext_ : (void*)own_.data(); }
 private:
    TensorShape shape_;
    mutable std::vector<char> own_;
    void* ext_;
};

struct DeviceBase {
    struct CpuWorkerThreads {
        int num_threads = 1;
        thread::ThreadPool* workers = nullptr;
    };
    const CpuWorkerThreads* tensorflow_cpu_worker_threads() const { return &threads_; }
    CpuWorkerThreads threads_;
};

class OpKernelConstruction {
 public:
    std::map<std::string, double> attrs;
    Status GetAttr(const char* name, int* v) { *v = (int)attrs.at(name); return Status::OK(); }
    Status GetAttr(const char* name, float* v) { *v = (float)attrs.at(name); return Status::OK(); }
    void SetStatus(const Status& s) { status_ = s; }
    Status status_;
};

class OpKernelContext {
 public:
    std::vector<Tensor*> inputs;
    std::vector<std::unique_ptr<Tensor>> outputs;
    std::vector<void*> output_buffers;   // caller memory per output index (written in place)
    const Tensor& input(int i) const { return *inputs[i]; }
    Status allocate_output(int i, const TensorShape& shape, Tensor** out) {
        if ((int)outputs.size() <= i) outputs.resize(i + 1);
        outputs[i].reset(new Tensor(shape, output_buffers.at(i)));
        *out = outputs[i].get();
        return Status::OK();
    }
    DeviceBase* device() { return &device_; }
    const Status& status() const { return status_; }
    void SetStatus(const Status& s) { status_ = s; }
    template <typename D>
    const D& eigen_device() const { static D d; return d; }
    void CtxFailure(const Status& s) { status_ = s; }
    void CtxFailureWithWarning(const Status& s) { status_ = s; }
 private:
    DeviceBase device_;
    Status status_;
};

class OpKernel {
 public:
    explicit OpKernel(OpKernelConstruction*) {}
    virtual ~OpKernel() {}
    virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS)          \
    do {                                       \
        if (!(EXP)) { (CTX)->SetStatus(STATUS); return; } \
    } while (0)
#define OP_REQUIRES_OK(CTX, ...)               \
    do {                                       \
        ::tensorflow::Status _s(__VA_ARGS__);  \
        if (!_s.ok()) { (CTX)->SetStatus(_s); return; } \
    } while (0)

// ---- kernel registry ------------------------------------------------------------------------------------------
static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

namespace shim {
typedef std::function<OpKernel*(OpKernelConstruction*)> Factory;
inline std::map<std::string, Factory>& registry() {
    static std::map<std::string, Factory> r;
    return r;
}
struct KernelDef {
    std::string key;
    KernelDef& Device(const char* d) { key += std::string("/") + d; return *this; }
    template <typename T>
    KernelDef& TypeConstraint(const char*) { return *this; }
};
struct Registrar {
    Registrar(const KernelDef& d, Factory f) { registry()[d.key] = f; }
};
}  // namespace shim
inline shim::KernelDef Name(const char* op) { shim::KernelDef d; d.key = op; return d; }

#define REGISTER_KERNEL_BUILDER(DEF, ...)                                                        \
    static ::tensorflow::shim::Registrar TF_SHIM_CAT(tf_shim_kernel_, __COUNTER__)(              \
        DEF, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow
