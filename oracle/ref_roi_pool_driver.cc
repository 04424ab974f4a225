// TEST INFRASTRUCTURE ONLY.  C entry points over the reference's own RoiPoolOp<CPUDevice,float> /
// RoiPoolGradOp<CPUDevice,float> (lib/roi_pooling_layer/roi_pooling_op.cc:51-195, :295-457), compiled UNMODIFIED from
// /root/reference against the stand-in headers in oracle/tf_stub/ (recipe: oracle/build_ref_roi_pool.py -> oracle/_ref/
// libref_roi_pool.so).  Used by tests/test_oracle_vs_reference.py to pin oracle_c.c's orc_roi_pool_* restatement and to
// generate tests/golden/roi_pool.npz.
#include <cstdio>
#include <memory>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "work_sharder.h"

namespace tensorflow {
// work_sharder.h declares it; TensorFlow's implementation splits [0,total) over a thread pool -- any partition gives the
// same result (each unit writes its own output element), so the shim runs one shard.
void Shard(int, thread::ThreadPool*, int64 total, int64, std::function<void(int64, int64)> work) { work(0, total); }
}  // namespace tensorflow

// The reference declares the CUDA launchers next to the CPU kernels (roi_pooling_op.cc:197-202, :459-465); the GPU
// kernel classes are only registered under GOOGLE_CUDA, which this build does not define.

using namespace tensorflow;

static OpKernel* make(const char* key, int ph, int pw, float scale, OpKernelConstruction* c) {
    c->attrs["pooled_height"] = ph;
    c->attrs["pooled_width"] = pw;
    c->attrs["spatial_scale"] = scale;
    auto it = shim::registry().find(key);
    return it == shim::registry().end() ? nullptr : it->second(c);
}

extern "C" int ref_roi_pool_forward(const float* data, int B, int H, int W, int C, const float* rois, int R, int ph,
                                    int pw, float scale, float* top, int* argmax) {
    OpKernelConstruction c;
    std::unique_ptr<OpKernel> k(make("RoiPool/CPU", ph, pw, scale, &c));
    if (!k) return -1;
    Tensor t_data(TensorShape({B, H, W, C}), (void*)data), t_rois(TensorShape({R, 5}), (void*)rois);
    OpKernelContext ctx;
    ctx.inputs = {&t_data, &t_rois};
    ctx.output_buffers = {top, argmax};
    k->Compute(&ctx);
    if (!ctx.status().ok()) { fprintf(stderr, "RoiPool: %s\n", ctx.status().error_message().c_str()); return -2; }
    return 0;
}

extern "C" int ref_roi_pool_backward(const float* data, int B, int H, int W, int C, const float* rois, int R, int ph,
                                     int pw, float scale, const int* argmax, const float* grad, float* out) {
    OpKernelConstruction c;
    std::unique_ptr<OpKernel> k(make("RoiPoolGrad/CPU", ph, pw, scale, &c));
    if (!k) return -1;
    Tensor t_data(TensorShape({B, H, W, C}), (void*)data), t_rois(TensorShape({R, 5}), (void*)rois);
    Tensor t_arg(TensorShape({R, ph, pw, C}), (void*)argmax), t_grad(TensorShape({R, ph, pw, C}), (void*)grad);
    OpKernelContext ctx;
    ctx.inputs = {&t_data, &t_rois, &t_arg, &t_grad};
    ctx.output_buffers = {out};
    k->Compute(&ctx);
    if (!ctx.status().ok()) { fprintf(stderr, "RoiPoolGrad: %s\n", ctx.status().error_message().c_str()); return -2; }
    return 0;
}
