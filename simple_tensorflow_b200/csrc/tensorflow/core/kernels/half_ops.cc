// DT_HALF variants of the hot-path ops for DEVICE_GPU on B200.
//
// The reference registers half on GPU for MatMul (core/kernels/matmul_op.cc:301-332), Conv2D and
// its gradients (conv_ops.cc:758-763, conv_grad_input_ops.cc:960-969, conv_grad_filter_ops.cc:
// 781-790), MaxPool (maxpooling_op.cc:646-651), BiasAdd(+Grad) (bias_op.cc:242-299), Relu(+Grad)
// (relu_op.cc GPU registrations) and the softmax family.  Here ONE adaptor kernel serves all of
// them: it owns the fp32 kernel of the same node (T = float), feeds it the half inputs widened to
// fp32 (exact), and rounds its outputs to half (nearest even, like Eigen::half).
//
// Why that is the fp16 tensor-core result and not an approximation of it: every fp16 value is
// exactly representable in TF32 (10 mantissa bits, wider exponent), so the tcgen05 kind::tf32
// MMAs of the fp32 MatMul / Conv kernels multiply the half operands exactly and accumulate in
// fp32 -- what a kind::f16 MMA with fp32 accumulation computes -- and the element-wise kernels
// compute in fp32 and round once, i.e. correctly rounded half arithmetic.  The cost is the two
// cast passes (3x the bytes of a native half kernel); half is not on the benchmarked path.
// Reference tolerance for half: 1e-3 (python/framework/test_util.py:515-523).
#include <memory>
#include <vector>

#include "tensorflow/core/kernels/gpu_kernel_util.h"

namespace tensorflow {
namespace {

class HalfViaFloatOp : public OpKernel {
 public:
  explicit HalfViaFloatOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    NodeDef nd = ctx->def();
    nd.name += "/_as_float";
    nd.attr["T"] = AttrValue::Type(DT_FLOAT);
    Status s = CreateOpKernel(ctx->device_type(), ctx->device(),
                              ctx->device()->GetAllocator(AllocatorAttributes()), nd, &inner_);
    if (!s.ok()) ctx->SetStatus(s);
  }

  void Compute(OpKernelContext* ctx) override {
    void* stream = GetCudaStream(ctx);
    const int n = ctx->num_inputs();
    std::vector<Tensor> held(n);
    std::vector<TensorValue> values(n);
    for (int i = 0; i < n; ++i) {
      OP_REQUIRES(ctx, !input_is_ref(i),
                  errors::Unimplemented("half adaptor: reference inputs are not supported"));
      const Tensor& in = ctx->input(i);
      if (input_type(i) == DT_HALF && inner_->input_type(i) == DT_FLOAT) {
        OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_FLOAT, in.shape(), &held[i]));
        if (in.NumElements() > 0)
          OP_REQUIRES_OK(ctx, FromAbi(b200_cast(B200_DT_HALF, B200_DT_FLOAT, in.raw_data(),
                                                held[i].raw_data(), in.NumElements(), stream),
                                      "half -> float"));
      } else {
        held[i] = in;  // host-memory shape vectors, indices ...: passed through
      }
      values[i] = TensorValue(&held[i]);
    }
    OpKernelContext::Params params;
    params.step_id = ctx->step_id();
    params.op_kernel = inner_.get();
    params.device = ctx->device();
    params.inputs = &values;
    params.op_device_context = ctx->op_device_context();
    OpKernelContext inner_ctx(&params);
    inner_->Compute(&inner_ctx);
    if (!inner_ctx.status().ok()) {
      ctx->SetStatus(inner_ctx.status());
      return;
    }
    for (int o = 0; o < num_outputs(); ++o) {
      TensorValue v = inner_ctx.release_output(o);
      std::unique_ptr<Tensor> result(v.tensor);
      OP_REQUIRES(ctx, result != nullptr,
                  errors::Internal("half adaptor: the fp32 kernel produced no output ", o));
      if (output_type(o) == DT_HALF && result->dtype() == DT_FLOAT) {
        Tensor* out = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(o, result->shape(), &out));
        if (result->NumElements() > 0)
          OP_REQUIRES_OK(ctx, FromAbi(b200_cast(B200_DT_FLOAT, B200_DT_HALF, result->raw_data(),
                                                out->raw_data(), result->NumElements(), stream),
                                      "float -> half"));
      } else {
        ctx->set_output(o, *result);
      }
    }
  }

 private:
  std::unique_ptr<OpKernel> inner_;
};

#define REGISTER_HALF(NAME) \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(DEVICE_GPU).TypeConstraint<half>("T"), HalfViaFloatOp)
REGISTER_HALF("MatMul");
REGISTER_HALF("BatchMatMul");
REGISTER_HALF("Conv2D");
REGISTER_HALF("BiasAdd");
REGISTER_HALF("BiasAddGrad");
REGISTER_HALF("Relu");
REGISTER_HALF("ReluGrad");
REGISTER_HALF("MaxPool");
REGISTER_HALF("MaxPoolGrad");
REGISTER_HALF("Softmax");
REGISTER_HALF("LogSoftmax");
REGISTER_HALF("SoftmaxCrossEntropyWithLogits");
REGISTER_HALF("AddN");
REGISTER_HALF("Add");
REGISTER_HALF("Mul");
#undef REGISTER_HALF
REGISTER_KERNEL_BUILDER(
    Name("Conv2DBackpropInput").Device(DEVICE_GPU).TypeConstraint<half>("T").HostMemory("input_sizes"),
    HalfViaFloatOp);
REGISTER_KERNEL_BUILDER(
    Name("Conv2DBackpropFilter").Device(DEVICE_GPU).TypeConstraint<half>("T").HostMemory("filter_sizes"),
    HalfViaFloatOp);
REGISTER_KERNEL_BUILDER(
    Name("Mean").Device(DEVICE_GPU).TypeConstraint<half>("T").HostMemory("reduction_indices"),
    HalfViaFloatOp);
REGISTER_KERNEL_BUILDER(
    Name("Sum").Device(DEVICE_GPU).TypeConstraint<half>("T").HostMemory("reduction_indices"),
    HalfViaFloatOp);
// Const of dtype half: the value travels as bits
}  // namespace
}  // namespace tensorflow
