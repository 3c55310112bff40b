// Conv2D / Conv2DBackpropInput / Conv2DBackpropFilter for DEVICE_GPU on B200 (NHWC-native, HWIO;
// data_format NCHW is served by transposing the activations in and out, gpu_kernel_util.h).
// Attr and shape validation follows Conv2DOp (core/kernels/conv_ops.cc:244-391),
// Conv2DSlowBackpropInputOp (conv_grad_input_ops.cc:533-917) and Conv2DSlowBackpropFilterOp
// (conv_grad_filter_ops.cc:361-738) with ConvBackpropComputeDimensions (conv_grad_ops.cc:37-126);
// the launches go to b200_conv2d* instead of cuDNN + layout shuffles.  `input_sizes` /
// `filter_sizes` arrive in host memory like in the reference registrations
// (conv_grad_input_ops.cc:960-969, conv_grad_filter_ops.cc:781-790).
#include <limits>

#include "tensorflow/core/kernels/gpu_kernel_util.h"
#include "tensorflow/core/util/padding.h"

namespace tensorflow {
namespace {

struct ConvAttrs {
  std::vector<int32> strides;  // always in NHWC order after Init()
  Padding padding;
  bool nchw = false;
  Status Init(OpKernelConstruction* context) {
    TF_RETURN_IF_ERROR(context->GetAttr("strides", &strides));
    std::string data_format;
    TF_RETURN_IF_ERROR(context->GetAttr("data_format", &data_format));
    TensorFormat fmt;
    if (!FormatFromString(data_format, &fmt)) return errors::InvalidArgument("Invalid data format");
    nchw = fmt == FORMAT_NCHW;
    if (strides.size() != 4)
      return errors::InvalidArgument("Sliding window strides field must specify 4 dimensions");
    if (nchw) strides = {strides[0], strides[2], strides[3], strides[1]};  // N C H W -> N H W C
    if (strides[0] != 1 || strides[3] != 1)
      return errors::InvalidArgument("Current implementation does not yet support strides in the "
                                     "batch and depth dimensions.");
    return context->GetAttr("padding", &padding);
  }
};

Status MakeGeometry(const char* label, const TensorShape& input, const TensorShape& filter,
                    const TensorShape* out_backprop, const ConvAttrs& a,
                    b200_conv2d_geometry* g, TensorShape* out_shape) {
  if (input.dims() != 4)
    return errors::InvalidArgument(label, ": input must be 4-dimensional", input.DebugString());
  if (filter.dims() != 4)
    return errors::InvalidArgument(label, ": filter must be 4-dimensional: ", filter.DebugString());
  for (int i = 0; i < 3; ++i)
    if (filter.dim_size(i) > std::numeric_limits<int>::max())
      return errors::InvalidArgument("filter too large");
  if (input.dim_size(3) != filter.dim_size(2))
    return errors::InvalidArgument(label, ": input and filter must have the same depth: ",
                                   input.dim_size(3), " vs ", filter.dim_size(2));
  for (int i = 0; i < 3; ++i)
    if (input.dim_size(i) > std::numeric_limits<int>::max())
      return errors::InvalidArgument(i == 0 ? "batch is too large"
                                            : (i == 1 ? "Input rows too large"
                                                      : "Input cols too large"));
  int64 out_rows = 0, out_cols = 0, pad_rows = 0, pad_cols = 0;
  TF_RETURN_IF_ERROR(GetWindowedOutputSize(input.dim_size(1), filter.dim_size(0), a.strides[1],
                                           a.padding, &out_rows, &pad_rows));
  TF_RETURN_IF_ERROR(GetWindowedOutputSize(input.dim_size(2), filter.dim_size(1), a.strides[2],
                                           a.padding, &out_cols, &pad_cols));
  *out_shape = TensorShape({input.dim_size(0), out_rows, out_cols, filter.dim_size(3)});
  if (out_backprop) {
    if (out_backprop->dims() != 4)
      return errors::InvalidArgument(label, ": out_backprop must be 4-dimensional");
    if (out_backprop->dim_size(0) != input.dim_size(0))
      return errors::InvalidArgument(label,
                                     ": input and out_backprop must have the same batch size");
    if (out_backprop->dim_size(3) != filter.dim_size(3))
      return errors::InvalidArgument(label,
                                     ": filter and out_backprop must have the same out_depth");
    if (out_backprop->dim_size(1) != out_rows || out_backprop->dim_size(2) != out_cols)
      return errors::InvalidArgument(label, ": Size of out_backprop doesn't match computed: ",
                                     "actual = ", out_backprop->DebugString(),
                                     ", computed = ", out_shape->DebugString());
  }
  g->batch = input.dim_size(0);
  g->in_h = input.dim_size(1);
  g->in_w = input.dim_size(2);
  g->in_c = input.dim_size(3);
  g->filter_h = filter.dim_size(0);
  g->filter_w = filter.dim_size(1);
  g->out_c = filter.dim_size(3);
  g->out_h = out_rows;
  g->out_w = out_cols;
  g->stride_h = a.strides[1];
  g->stride_w = a.strides[2];
  g->pad_top = static_cast<int32_t>(pad_rows);
  g->pad_left = static_cast<int32_t>(pad_cols);
  return Status::OK();
}

Status ShapeFromHostVector(const Tensor& t, const char* what, TensorShape* shape) {
  if (!TensorShapeUtils::IsVector(t.shape()))
    return errors::InvalidArgument("Conv2DBackprop: ", what, " input must be 1-dim, not ",
                                   t.dims());
  return TensorShapeUtils::MakeShape(t.data<int32>(), t.NumElements(), shape);
}

Status Scratch(OpKernelContext* ctx, size_t bytes, Tensor* t) {
  if (bytes == 0) return Status::OK();
  return ctx->allocate_temp(DT_UINT8, TensorShape({static_cast<int64>(bytes)}), t);
}

}  // namespace

template <typename T>
class Conv2DOp : public OpKernel {
 public:
  explicit Conv2DOp(OpKernelConstruction* context) : OpKernel(context) {
    OP_REQUIRES_OK(context, attrs_.Init(context));
  }
  void Compute(OpKernelContext* context) override {
    Tensor input = context->input(0);
    const Tensor& filter = context->input(1);
    OP_REQUIRES(context, input.dims() == 4,
                errors::InvalidArgument("input must be 4-dimensional", input.shape().DebugString()));
    const bool nchw = attrs_.nchw && input.NumElements() > 0;
    if (nchw) OP_REQUIRES_OK(context, NchwToNhwc<T>(context, context->input(0), &input));
    const TensorShape in_nhwc = attrs_.nchw && !nchw ? NchwToNhwcShape(input.shape()) : input.shape();
    b200_conv2d_geometry g;
    TensorShape out_shape;  // NHWC
    OP_REQUIRES_OK(context, MakeGeometry("Conv2D", in_nhwc, filter.shape(), nullptr, attrs_, &g,
                                         &out_shape));
    Tensor* output = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(
                                0, attrs_.nchw ? NhwcToNchwShape(out_shape) : out_shape, &output));
    if (out_shape.num_elements() == 0) return;  // conv_ops.cc:357-359
    Tensor out_nhwc;
    if (attrs_.nchw)
      OP_REQUIRES_OK(context, context->allocate_temp(output->dtype(), out_shape, &out_nhwc));
    const size_t ws = b200_conv2d_workspace_bytes(AbiType<T>::v, &g, 0);
    Tensor scratch;
    OP_REQUIRES_OK(context, Scratch(context, ws, &scratch));
    OP_REQUIRES_OK(context,
                   FromAbi(b200_conv2d(AbiType<T>::v, input.raw_data(), filter.raw_data(),
                                       attrs_.nchw ? out_nhwc.raw_data() : output->raw_data(), &g,
                                       ws ? scratch.raw_data() : nullptr, ws,
                                       GetCudaStream(context)),
                           "Conv2D"));
    if (attrs_.nchw) OP_REQUIRES_OK(context, NhwcToNchw<T>(context, out_nhwc, output));
  }

 private:
  ConvAttrs attrs_;
};

// Conv2D + BiasAdd (+ Relu) in one launch: created only by DirectSession::FuseConvChains for
// NHWC graphs; validation = Conv2DOp's plus BiasOp's (bias_op.cc:62-82).
template <typename T>
class FusedConv2DOp : public OpKernel {
 public:
  explicit FusedConv2DOp(OpKernelConstruction* context) : OpKernel(context) {
    OP_REQUIRES_OK(context, attrs_.Init(context));
    OP_REQUIRES(context, !attrs_.nchw,
                errors::InvalidArgument("_FusedConv2D is NHWC-only"));
    std::vector<std::string> fused;
    OP_REQUIRES_OK(context, context->GetAttr("fused_ops", &fused));
    if (fused == std::vector<std::string>{"BiasAdd"}) relu_ = false;
    else if (fused == std::vector<std::string>{"BiasAdd", "Relu"}) relu_ = true;
    else
      OP_REQUIRES(context, false, errors::InvalidArgument("Unsupported fused_ops for _FusedConv2D"));
    OP_REQUIRES(context, context->num_inputs() == 3,
                errors::InvalidArgument("_FusedConv2D expects exactly one extra argument"));
  }
  void Compute(OpKernelContext* context) override {
    const Tensor& input = context->input(0);
    const Tensor& filter = context->input(1);
    const Tensor& bias = context->input(2);
    OP_REQUIRES(context, input.dims() == 4,
                errors::InvalidArgument("input must be 4-dimensional", input.shape().DebugString()));
    b200_conv2d_geometry g;
    TensorShape out_shape;
    OP_REQUIRES_OK(context, MakeGeometry("Conv2D", input.shape(), filter.shape(), nullptr, attrs_,
                                         &g, &out_shape));
    OP_REQUIRES(context, TensorShapeUtils::IsVector(bias.shape()),
                errors::InvalidArgument("Biases must be 1D: ", bias.shape().DebugString()));
    OP_REQUIRES(context, bias.dim_size(0) == out_shape.dim_size(3),
                errors::InvalidArgument("Must provide as many biases as the last dimension of the "
                                        "input tensor: ", bias.shape().DebugString(), " vs. ",
                                        out_shape.DebugString()));
    Tensor* output = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(0, out_shape, &output));
    if (out_shape.num_elements() == 0) return;
    const size_t ws = b200_conv2d_workspace_bytes(AbiType<T>::v, &g, 0);
    Tensor scratch;
    OP_REQUIRES_OK(context, Scratch(context, ws, &scratch));
    OP_REQUIRES_OK(context,
                   FromAbi(b200_fused_conv2d(AbiType<T>::v, input.raw_data(), filter.raw_data(),
                                             bias.raw_data(), relu_ ? 1 : 0, output->raw_data(), &g,
                                             ws ? scratch.raw_data() : nullptr, ws,
                                             GetCudaStream(context)),
                           "_FusedConv2D"));
  }

 private:
  ConvAttrs attrs_;
  bool relu_ = false;
};

template <typename T>
class Conv2DBackpropInputOp : public OpKernel {
 public:
  explicit Conv2DBackpropInputOp(OpKernelConstruction* context) : OpKernel(context) {
    OP_REQUIRES_OK(context, attrs_.Init(context));
  }
  void Compute(OpKernelContext* context) override {
    const Tensor& input_sizes = context->input(0);  // host memory
    const Tensor& filter = context->input(1);
    Tensor out_backprop = context->input(2);
    TensorShape input_shape;  // as given: data_format order
    OP_REQUIRES_OK(context, ShapeFromHostVector(input_sizes, "input_sizes", &input_shape));
    OP_REQUIRES(context, input_shape.dims() == 4 && out_backprop.dims() == 4,
                errors::InvalidArgument("Conv2DBackpropInput: input_sizes and out_backprop must be "
                                        "4-dimensional"));
    const TensorShape in_nhwc = attrs_.nchw ? NchwToNhwcShape(input_shape) : input_shape;
    const bool nchw = attrs_.nchw && out_backprop.NumElements() > 0;
    if (nchw) OP_REQUIRES_OK(context, NchwToNhwc<T>(context, context->input(2), &out_backprop));
    const TensorShape dy_nhwc =
        attrs_.nchw && !nchw ? NchwToNhwcShape(out_backprop.shape()) : out_backprop.shape();
    b200_conv2d_geometry g;
    TensorShape out_shape;
    OP_REQUIRES_OK(context, MakeGeometry("Conv2DBackpropInput", in_nhwc, filter.shape(), &dy_nhwc,
                                         attrs_, &g, &out_shape));
    Tensor* in_backprop = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(0, input_shape, &in_backprop));
    if (input_shape.num_elements() == 0) return;
    Tensor dx_nhwc;
    if (attrs_.nchw)
      OP_REQUIRES_OK(context, context->allocate_temp(in_backprop->dtype(), in_nhwc, &dx_nhwc));
    const size_t ws = b200_conv2d_workspace_bytes(AbiType<T>::v, &g, 1);
    Tensor scratch;
    OP_REQUIRES_OK(context, Scratch(context, ws, &scratch));
    OP_REQUIRES_OK(context, FromAbi(b200_conv2d_backprop_input(
                                        AbiType<T>::v, filter.raw_data(), out_backprop.raw_data(),
                                        attrs_.nchw ? dx_nhwc.raw_data() : in_backprop->raw_data(),
                                        &g, ws ? scratch.raw_data() : nullptr, ws,
                                        GetCudaStream(context)),
                                    "Conv2DBackpropInput"));
    if (attrs_.nchw) OP_REQUIRES_OK(context, NhwcToNchw<T>(context, dx_nhwc, in_backprop));
  }

 private:
  ConvAttrs attrs_;
};

template <typename T>
class Conv2DBackpropFilterOp : public OpKernel {
 public:
  explicit Conv2DBackpropFilterOp(OpKernelConstruction* context) : OpKernel(context) {
    OP_REQUIRES_OK(context, attrs_.Init(context));
  }
  void Compute(OpKernelContext* context) override {
    Tensor input = context->input(0);
    const Tensor& filter_sizes = context->input(1);  // host memory
    Tensor out_backprop = context->input(2);
    TensorShape filter_shape;
    OP_REQUIRES_OK(context, ShapeFromHostVector(filter_sizes, "filter_sizes", &filter_shape));
    OP_REQUIRES(context, input.dims() == 4 && out_backprop.dims() == 4,
                errors::InvalidArgument("Conv2DBackpropFilter: input and out_backprop must be "
                                        "4-dimensional"));
    TensorShape x_nhwc = input.shape(), dy_nhwc = out_backprop.shape();
    if (attrs_.nchw) {
      x_nhwc = NchwToNhwcShape(input.shape());
      dy_nhwc = NchwToNhwcShape(out_backprop.shape());
      if (input.NumElements() > 0)
        OP_REQUIRES_OK(context, NchwToNhwc<T>(context, context->input(0), &input));
      if (out_backprop.NumElements() > 0)
        OP_REQUIRES_OK(context, NchwToNhwc<T>(context, context->input(2), &out_backprop));
    }
    b200_conv2d_geometry g;
    TensorShape out_shape;
    OP_REQUIRES_OK(context, MakeGeometry("Conv2DBackpropFilter", x_nhwc, filter_shape, &dy_nhwc,
                                         attrs_, &g, &out_shape));
    Tensor* filter_backprop = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(0, filter_shape, &filter_backprop));
    if (filter_shape.num_elements() == 0) return;
    const size_t ws = b200_conv2d_workspace_bytes(AbiType<T>::v, &g, 2);
    Tensor scratch;
    OP_REQUIRES_OK(context, Scratch(context, ws, &scratch));
    OP_REQUIRES_OK(context, FromAbi(b200_conv2d_backprop_filter(
                                        AbiType<T>::v, input.raw_data(), out_backprop.raw_data(),
                                        filter_backprop->raw_data(), &g,
                                        ws ? scratch.raw_data() : nullptr, ws,
                                        GetCudaStream(context)),
                                    "Conv2DBackpropFilter"));
  }

 private:
  ConvAttrs attrs_;
};

#define REGISTER_GPU(T)                                                                       \
  REGISTER_KERNEL_BUILDER(Name("_FusedConv2D").Device(DEVICE_GPU).TypeConstraint<T>("T"),     \
                          FusedConv2DOp<T>);                                                  \
  REGISTER_KERNEL_BUILDER(Name("Conv2D").Device(DEVICE_GPU).TypeConstraint<T>("T"),           \
                          Conv2DOp<T>);                                                       \
  REGISTER_KERNEL_BUILDER(Name("Conv2DBackpropInput")                                         \
                              .Device(DEVICE_GPU)                                             \
                              .TypeConstraint<T>("T")                                         \
                              .HostMemory("input_sizes"),                                     \
                          Conv2DBackpropInputOp<T>);                                          \
  REGISTER_KERNEL_BUILDER(Name("Conv2DBackpropFilter")                                        \
                              .Device(DEVICE_GPU)                                             \
                              .TypeConstraint<T>("T")                                         \
                              .HostMemory("filter_sizes"),                                    \
                          Conv2DBackpropFilterOp<T>);
REGISTER_B200_FLOAT_TYPES(REGISTER_GPU)
#undef REGISTER_GPU

}  // namespace tensorflow
