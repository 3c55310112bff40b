// Cast and ArgMax for DEVICE_GPU on B200 -- the bit-exact ops of the hot path.
// Cast: CastOpBase (core/kernels/cast_op.cc:54-133): attrs SrcT/DstT; same-type aliases the
// input (:63-66); float->bfloat16 truncates (cast_op.h:119-141).
// ArgMax: ArgOp (core/kernels/argmax_op.cc:44-98): `dimension` is a host-memory scalar
// (:119-124), negative axes wrap (:60), output int64, ranks 1-5.
#include "tensorflow/core/kernels/gpu_kernel_util.h"

namespace tensorflow {

class GpuCastOp : public OpKernel {
 public:
  explicit GpuCastOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("SrcT", &src_dtype_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("DstT", &dst_dtype_));
    auto ok = [](DataType t) {
      return t == DT_FLOAT || t == DT_BFLOAT16 || t == DT_INT32 || t == DT_INT64;
    };
    const bool half_pair = (src_dtype_ == DT_HALF && (dst_dtype_ == DT_FLOAT || dst_dtype_ == DT_HALF)) ||
                           (dst_dtype_ == DT_HALF && src_dtype_ == DT_FLOAT);
    OP_REQUIRES(ctx, half_pair || (ok(src_dtype_) && ok(dst_dtype_)),
                errors::Unimplemented("Cast ", DataTypeString(src_dtype_), " to ",
                                      DataTypeString(dst_dtype_), " is not supported"));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& inp = ctx->input(0);
    if (src_dtype_ == dst_dtype_) {
      ctx->set_output(0, inp);
      return;
    }
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, inp.shape(), &out));
    OP_REQUIRES_OK(ctx, FromAbi(b200_cast(src_dtype_, dst_dtype_, inp.raw_data(), out->raw_data(),
                                          inp.NumElements(), GetCudaStream(ctx)),
                                "Cast"));
  }

 private:
  DataType src_dtype_;
  DataType dst_dtype_;
};
REGISTER_KERNEL_BUILDER(Name("Cast").Device(DEVICE_GPU), GpuCastOp);

template <typename T>
class ArgMaxOp : public OpKernel {
 public:
  explicit ArgMaxOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& input = ctx->input(0);
    const Tensor& dimension = ctx->input(1);  // host memory
    OP_REQUIRES(ctx, TensorShapeUtils::IsScalar(dimension.shape()),
                errors::InvalidArgument("dim must be a scalar, but received tensor of shape: ",
                                        dimension.shape().DebugString()));
    const int32 dim = dimension.dtype() == DT_INT64
                          ? static_cast<int32>(dimension.scalar<int64>())
                          : dimension.scalar<int32>();
    const int input_dims = input.dims();
    const int axis = dim < 0 ? dim + input_dims : dim;
    OP_REQUIRES(ctx, axis >= 0 && axis < input_dims,
                errors::InvalidArgument("Expected dimension in the range [", -input_dims, ", ",
                                        input_dims, "), but got ", dim));
    OP_REQUIRES(ctx, input.dim_size(axis) > 0,
                errors::InvalidArgument("Reduction axis ", dim, " is empty in shape ",
                                        input.shape().DebugString()));
    TensorShape output_shape;
    int64 outer = 1, inner = 1;
    for (int d = 0; d < input_dims; ++d) {
      if (d == axis) continue;
      output_shape.AddDim(input.dim_size(d));
      (d < axis ? outer : inner) *= input.dim_size(d);
    }
    Tensor* output = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, output_shape, &output));
    OP_REQUIRES_OK(ctx, FromAbi(b200_argmax(AbiType<T>::v, input.raw_data(),
                                            output->data<int64_t>(), outer, input.dim_size(axis),
                                            inner, GetCudaStream(ctx)),
                                "ArgMax"));
  }
};
#define REGISTER_ARGMAX(T)                                                  \
  REGISTER_KERNEL_BUILDER(Name("ArgMax")                                    \
                              .Device(DEVICE_GPU)                           \
                              .TypeConstraint<T>("T")                       \
                              .HostMemory("dimension"),                     \
                          ArgMaxOp<T>);
REGISTER_ARGMAX(float)
REGISTER_ARGMAX(int32)
REGISTER_ARGMAX(int64)
#undef REGISTER_ARGMAX

}  // namespace tensorflow
