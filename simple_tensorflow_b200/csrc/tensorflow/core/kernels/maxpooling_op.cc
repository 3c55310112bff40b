// MaxPool / MaxPoolGrad for DEVICE_GPU on B200 (spatial pooling).  Both layouts are native: an NCHW
// tensor [N, C, H, W] (GPU-only in the reference, maxpooling_op.cc:341-404) IS the NHWC tensor
// [N * C, H, W, 1], so data_format NCHW runs the same kernels with batch = N * C and depth 1 --
// no layout change, loads coalesced along W.
// Parameter checks follow MaxPoolingOp / PoolParameters (core/kernels/pooling_ops_common.h:73-130,
// pooling_ops_common.cc:31-90) and MaxPoolingGradOp (core/kernels/maxpooling_op.cc:230-306).
#include "tensorflow/core/kernels/gpu_kernel_util.h"
#include "tensorflow/core/util/padding.h"

namespace tensorflow {
namespace {

struct PoolAttrs {
  std::vector<int32> ksize, stride;
  Padding padding;
  bool nchw = false;
  Status Init(OpKernelConstruction* context) {
    std::string data_format;
    if (context->GetAttr("data_format", &data_format).ok()) {
      TensorFormat fmt;
      if (!FormatFromString(data_format, &fmt)) return errors::InvalidArgument("Invalid data format");
      nchw = fmt == FORMAT_NCHW;
    }
    TF_RETURN_IF_ERROR(context->GetAttr("ksize", &ksize));
    if (ksize.size() != 4)
      return errors::InvalidArgument("Sliding window ksize field must specify 4 dimensions");
    TF_RETURN_IF_ERROR(context->GetAttr("strides", &stride));
    if (stride.size() != 4)
      return errors::InvalidArgument("Sliding window stride field must specify 4 dimensions");
    TF_RETURN_IF_ERROR(context->GetAttr("padding", &padding));
    if (nchw) {  // N C H W -> N H W C
      ksize = {ksize[0], ksize[2], ksize[3], ksize[1]};
      stride = {stride[0], stride[2], stride[3], stride[1]};
    }
    if (ksize[0] != 1 || stride[0] != 1)
      return errors::Unimplemented("Pooling is not yet supported on the batch dimension.");
    if (ksize[3] != 1 || stride[3] != 1)
      // same answer as the reference on a GPU device (pooling_ops_common.cc:80-86)
      return errors::Unimplemented("Depthwise max pooling is currently only implemented for CPU "
                                   "devices.");
    return Status::OK();
  }
};

struct PoolDims {
  int64 batch, rows, cols, depth, out_rows, out_cols, pad_rows, pad_cols;
};
Status ComputePoolDims(const TensorShape& in, const PoolAttrs& a, PoolDims* d) {
  if (in.dims() != 4) return errors::InvalidArgument("tensor_in must be 4-dimensional");
  d->batch = in.dim_size(0);
  d->rows = in.dim_size(1);
  d->cols = in.dim_size(2);
  d->depth = in.dim_size(3);
  TF_RETURN_IF_ERROR(GetWindowedOutputSize(d->rows, a.ksize[1], a.stride[1], a.padding,
                                           &d->out_rows, &d->pad_rows));
  return GetWindowedOutputSize(d->cols, a.ksize[2], a.stride[2], a.padding, &d->out_cols,
                               &d->pad_cols);
}

}  // namespace

template <typename T>
class MaxPoolingOp : public OpKernel {
 public:
  explicit MaxPoolingOp(OpKernelConstruction* context) : OpKernel(context) {
    OP_REQUIRES_OK(context, attrs_.Init(context));
  }
  void Compute(OpKernelContext* context) override {
    Tensor tensor_in = context->input(0);
    OP_REQUIRES(context, tensor_in.dims() == 4,
                errors::InvalidArgument("tensor_in must be 4-dimensional"));
    const TensorShape in_nhwc =
        attrs_.nchw ? NchwToNhwcShape(tensor_in.shape()) : tensor_in.shape();
    PoolDims d;
    OP_REQUIRES_OK(context, ComputePoolDims(in_nhwc, attrs_, &d));
    const TensorShape out_nhwc({d.batch, d.out_rows, d.out_cols, d.depth});
    Tensor* result = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(
                                0, attrs_.nchw ? NhwcToNchwShape(out_nhwc) : out_nhwc, &result));
    if (result->NumElements() == 0) return;
    const int64 batch = attrs_.nchw ? d.batch * d.depth : d.batch;
    const int64 depth = attrs_.nchw ? 1 : d.depth;
    OP_REQUIRES_OK(context,
                   FromAbi(b200_max_pool(AbiType<T>::v, tensor_in.raw_data(), result->raw_data(),
                                         batch, d.rows, d.cols, depth, d.out_rows, d.out_cols,
                                         attrs_.ksize[1], attrs_.ksize[2], attrs_.stride[1],
                                         attrs_.stride[2], (int)d.pad_rows, (int)d.pad_cols,
                                         GetCudaStream(context)),
                           "MaxPool"));
  }

 private:
  PoolAttrs attrs_;
};

template <typename T>
class MaxPoolingGradOp : public OpKernel {
 public:
  explicit MaxPoolingGradOp(OpKernelConstruction* context) : OpKernel(context) {
    OP_REQUIRES_OK(context, attrs_.Init(context));
  }
  void Compute(OpKernelContext* context) override {
    Tensor tensor_in = context->input(0);
    Tensor tensor_out = context->input(1);
    Tensor out_backprop = context->input(2);
    OP_REQUIRES(context, tensor_in.dims() == 4,
                errors::InvalidArgument("tensor_in must be 4-dimensional"));
    OP_REQUIRES(context, tensor_out.dims() == 4,
                errors::InvalidArgument("tensor_out must be 4-dimensional"));
    OP_REQUIRES(context, out_backprop.dims() == 4,
                errors::InvalidArgument("out_backprop must be 4-dimensional"));
    const TensorShape in_nhwc =
        attrs_.nchw ? NchwToNhwcShape(tensor_in.shape()) : tensor_in.shape();
    const TensorShape dy_nhwc =
        attrs_.nchw ? NchwToNhwcShape(out_backprop.shape()) : out_backprop.shape();
    PoolDims d;
    OP_REQUIRES_OK(context, ComputePoolDims(in_nhwc, attrs_, &d));
    const TensorShape expect({d.batch, d.out_rows, d.out_cols, d.depth});
    OP_REQUIRES(context, dy_nhwc == expect,
                errors::InvalidArgument("out_backprop shape ", out_backprop.shape().DebugString(),
                                        " does not match the pooled shape ", expect.DebugString()));
    Tensor* result = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(0, tensor_in.shape(), &result));
    if (result->NumElements() == 0) return;
    const int64 batch = attrs_.nchw ? d.batch * d.depth : d.batch;
    const int64 depth = attrs_.nchw ? 1 : d.depth;
    OP_REQUIRES_OK(context, FromAbi(b200_max_pool_grad(
                                        AbiType<T>::v, tensor_in.raw_data(), tensor_out.raw_data(),
                                        out_backprop.raw_data(), result->raw_data(), batch, d.rows,
                                        d.cols, depth, d.out_rows, d.out_cols, attrs_.ksize[1],
                                        attrs_.ksize[2], attrs_.stride[1], attrs_.stride[2],
                                        (int)d.pad_rows, (int)d.pad_cols, GetCudaStream(context)),
                                    "MaxPoolGrad"));
  }

 private:
  PoolAttrs attrs_;
};

// `_MaxPoolGradReluGradBiasAddGrad` (NHWC only; the rewrite checks the format): one kernel when the
// windows tile the input, else MaxPoolGrad into a temporary followed by the fused
// ReluGrad + BiasAddGrad -- the same arithmetic as the three graph nodes.
template <typename T>
class MaxPoolGradReluGradBiasAddGradOp : public OpKernel {
 public:
  explicit MaxPoolGradReluGradBiasAddGradOp(OpKernelConstruction* context) : OpKernel(context) {
    OP_REQUIRES_OK(context, attrs_.Init(context));
    OP_REQUIRES(context, !attrs_.nchw,
                errors::InvalidArgument("_MaxPoolGradReluGradBiasAddGrad is NHWC-only"));
  }
  void Compute(OpKernelContext* context) override {
    const Tensor& tensor_in = context->input(0);
    const Tensor& out_backprop = context->input(2);
    OP_REQUIRES(context, tensor_in.dims() == 4,
                errors::InvalidArgument("tensor_in must be 4-dimensional"));
    OP_REQUIRES(context, out_backprop.dims() == 4,
                errors::InvalidArgument("out_backprop must be 4-dimensional"));
    PoolDims d;
    OP_REQUIRES_OK(context, ComputePoolDims(tensor_in.shape(), attrs_, &d));
    const TensorShape expect({d.batch, d.out_rows, d.out_cols, d.depth});
    OP_REQUIRES(context, out_backprop.shape() == expect,
                errors::InvalidArgument("out_backprop shape ", out_backprop.shape().DebugString(),
                                        " does not match the pooled shape ", expect.DebugString()));
    Tensor* backprops = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(0, tensor_in.shape(), &backprops));
    Tensor* bias_grad = nullptr;
    OP_REQUIRES_OK(context, context->allocate_output(1, TensorShape({d.depth}), &bias_grad));
    if (d.depth == 0) return;
    void* stream = GetCudaStream(context);
    if (backprops->NumElements() == 0) {
      OP_REQUIRES_OK(context, FromAbi(b200_memset_async(bias_grad->raw_data(), 0,
                                                        bias_grad->TotalBytes(), stream),
                                      "BiasAddGrad"));
      return;
    }
    const int dt = AbiType<T>::v;
    const size_t ws = b200_max_pool_grad_relu_bias_grad_workspace_bytes(
        dt, d.batch, d.rows, d.cols, d.depth, d.out_rows, d.out_cols, attrs_.ksize[1],
        attrs_.ksize[2], attrs_.stride[1], attrs_.stride[2], (int)d.pad_rows, (int)d.pad_cols);
    if (ws > 0) {
      Tensor scratch;
      OP_REQUIRES_OK(context, context->allocate_temp(DT_UINT8, TensorShape({(int64)ws}), &scratch));
      const int rc = b200_max_pool_grad_relu_bias_grad(
          dt, tensor_in.raw_data(), out_backprop.raw_data(), backprops->raw_data(),
          bias_grad->raw_data(), d.batch, d.rows, d.cols, d.depth, d.out_rows, d.out_cols,
          attrs_.ksize[1], attrs_.ksize[2], attrs_.stride[1], attrs_.stride[2], (int)d.pad_rows,
          (int)d.pad_cols, scratch.raw_data(), ws, stream);
      if (rc != B200_UNIMPLEMENTED) {
        OP_REQUIRES_OK(context, FromAbi(rc, "_MaxPoolGradReluGradBiasAddGrad"));
        return;
      }
    }
    Tensor dx;
    OP_REQUIRES_OK(context, context->allocate_temp(backprops->dtype(), tensor_in.shape(), &dx));
    OP_REQUIRES_OK(context,
                   FromAbi(b200_max_pool_grad(dt, tensor_in.raw_data(), nullptr,
                                              out_backprop.raw_data(), dx.raw_data(), d.batch,
                                              d.rows, d.cols, d.depth, d.out_rows, d.out_cols,
                                              attrs_.ksize[1], attrs_.ksize[2], attrs_.stride[1],
                                              attrs_.stride[2], (int)d.pad_rows, (int)d.pad_cols,
                                              stream),
                           "MaxPoolGrad"));
    const int64 rows = dx.NumElements() / d.depth;
    const size_t ws2 = b200_relu_grad_bias_grad_workspace_bytes(dt, rows, d.depth);
    Tensor scratch2;
    if (ws2 > 0)
      OP_REQUIRES_OK(context, context->allocate_temp(DT_UINT8, TensorShape({(int64)ws2}), &scratch2));
    OP_REQUIRES_OK(context, FromAbi(b200_relu_grad_bias_grad(
                                        dt, dx.raw_data(), tensor_in.raw_data(),
                                        backprops->raw_data(), bias_grad->raw_data(), rows, d.depth,
                                        ws2 ? scratch2.raw_data() : nullptr, ws2, stream),
                                    "_ReluGradBiasAddGrad"));
  }

 private:
  PoolAttrs attrs_;
};

#define REGISTER_GPU(T)                                                                             \
  REGISTER_KERNEL_BUILDER(                                                                          \
      Name("_MaxPoolGradReluGradBiasAddGrad").Device(DEVICE_GPU).TypeConstraint<T>("T"),            \
      MaxPoolGradReluGradBiasAddGradOp<T>);                                                         \
  REGISTER_KERNEL_BUILDER(Name("MaxPool").Device(DEVICE_GPU).TypeConstraint<T>("T"),                \
                          MaxPoolingOp<T>);                                                         \
  REGISTER_KERNEL_BUILDER(Name("MaxPoolGrad").Device(DEVICE_GPU).TypeConstraint<T>("T"),            \
                          MaxPoolingGradOp<T>);
REGISTER_B200_FLOAT_TYPES(REGISTER_GPU)
#undef REGISTER_GPU

}  // namespace tensorflow
