// TEST INFRASTRUCTURE ONLY -- stand-in for tensorflow/core/framework/op.h: REGISTER_OP(...) with its .Attr / .Input /
// .Output chain is accepted and ignored (the op's interface is fixed by the kernel classes the shim instantiates).
#pragma once
namespace tensorflow {
namespace shim {
struct OpDefBuilder {
    OpDefBuilder& Attr(const char*) { return *this; }
    OpDefBuilder& Input(const char*) { return *this; }
    OpDefBuilder& Output(const char*) { return *this; }
};
}  // namespace shim
}  // namespace tensorflow
#define TF_SHIM_CAT2(a, b) a##b
#define TF_SHIM_CAT(a, b) TF_SHIM_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::shim::OpDefBuilder TF_SHIM_CAT(tf_shim_op_, __COUNTER__) = ::tensorflow::shim::OpDefBuilder()
