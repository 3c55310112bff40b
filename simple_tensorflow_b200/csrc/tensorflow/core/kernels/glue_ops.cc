// Graph-glue kernels that keep a whole training step on the device:
// Const, Placeholder, Identity, Reshape, Shape, NoOp, VariableV2, Assign, AddN, Mul, Mean,
// ApplyGradientDescent, and the additive B200AllReduce.  Semantics follow the reference's
// core/kernels/{constant_op,identity_op,reshape_op,shape_ops,no_op,variable_ops,assign_op,
// aggregate_ops,cwise_op_mul_1,reduction_ops_mean,training_ops}.cc for the cases training
// graphs of the hot path produce.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <type_traits>

#include "tensorflow/core/common_runtime/device.h"
#include "tensorflow/core/common_runtime/gpu/gpu_device.h"
#include "tensorflow/core/kernels/gpu_kernel_util.h"

namespace tensorflow {

// ---------------------------------------------------------------- Const
// constant_op.cc: the tensor attr is materialised once, at kernel construction.  int32 outputs
// live in host memory like the reference's HostConstantOp registration (shape operands).
class ConstantOp : public OpKernel {
 public:
  explicit ConstantOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    Tensor host;
    OP_REQUIRES_OK(ctx, ctx->GetAttr("value", &host));
    DataType dtype;
    OP_REQUIRES_OK(ctx, ctx->GetAttr("dtype", &dtype));
    OP_REQUIRES(ctx, dtype == host.dtype(),
                errors::InvalidArgument("Type mismatch between value (", DataTypeString(host.dtype()),
                                        ") and dtype (", DataTypeString(dtype), ")"));
    if (ctx->output_memory_types()[0] == HOST_MEMORY) {
      tensor_ = host;
      return;
    }
    Device* device = dynamic_cast<Device*>(ctx->device());
    OP_REQUIRES(ctx, device != nullptr, errors::Internal("Const: no device"));
    OP_REQUIRES_OK(ctx, device->MakeTensorFromHost(host, &tensor_));
  }
  void Compute(OpKernelContext* ctx) override { ctx->set_output(0, tensor_); }

 private:
  Tensor tensor_;
};
REGISTER_KERNEL_BUILDER(Name("Const").Device(DEVICE_GPU).TypeConstraint<float>("dtype"), ConstantOp);
REGISTER_KERNEL_BUILDER(Name("Const").Device(DEVICE_GPU).TypeConstraint<bfloat16>("dtype"),
                        ConstantOp);
REGISTER_KERNEL_BUILDER(Name("Const").Device(DEVICE_GPU).TypeConstraint<half>("dtype"), ConstantOp);
REGISTER_KERNEL_BUILDER(Name("Const").Device(DEVICE_GPU).TypeConstraint<int64>("dtype"), ConstantOp);
REGISTER_KERNEL_BUILDER(
    Name("Const").Device(DEVICE_GPU).HostMemory("output").TypeConstraint<int32>("dtype"),
    ConstantOp);

// ---------------------------------------------------------------- Placeholder / NoOp / Identity
class PlaceholderOp : public OpKernel {
 public:
  explicit PlaceholderOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    // array_ops / constant_op.cc PlaceholderOp::Compute
    ctx->SetStatus(errors::InvalidArgument("You must feed a value for placeholder tensor '",
                                           name(), "' with dtype ",
                                           DataTypeString(output_type(0))));
  }
};
REGISTER_KERNEL_BUILDER(Name("Placeholder").Device(DEVICE_GPU), PlaceholderOp);

class NoOp : public OpKernel {
 public:
  explicit NoOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext*) override {}
};
REGISTER_KERNEL_BUILDER(Name("NoOp").Device(DEVICE_GPU), NoOp);

class IdentityOp : public OpKernel {
 public:
  explicit IdentityOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override { ctx->set_output(0, ctx->input(0)); }
};
REGISTER_KERNEL_BUILDER(Name("Identity").Device(DEVICE_GPU), IdentityOp);

// ---------------------------------------------------------------- Reshape / Shape
// reshape_op.h: `shape` is a host-memory vector, one -1 entry is inferred, the buffer is shared.
class ReshapeOp : public OpKernel {
 public:
  explicit ReshapeOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& input = ctx->input(0);
    const Tensor& sizes = ctx->input(1);
    OP_REQUIRES(ctx, TensorShapeUtils::IsVector(sizes.shape()),
                errors::InvalidArgument("sizes input must be 1-D, not shape ",
                                        sizes.shape().DebugString()));
    TensorShape shape;
    int64 product = 1;
    int unknown_index = -1;
    for (int d = 0; d < sizes.NumElements(); ++d) {
      const int64 size = sizes.dtype() == DT_INT64 ? sizes.data<int64>()[d]
                                                   : sizes.data<int32>()[d];
      if (size == -1) {
        OP_REQUIRES(ctx, unknown_index == -1,
                    errors::InvalidArgument("only one input size may be -1, not both ",
                                            unknown_index, " and ", d));
        unknown_index = d;
        shape.AddDim(1);
      } else {
        OP_REQUIRES(ctx, size >= 0, errors::InvalidArgument("size ", d, " must be non-negative, not ", size));
        shape.AddDim(size);
        product *= size;
      }
    }
    if (unknown_index != -1) {
      OP_REQUIRES(ctx, product > 0,
                  errors::InvalidArgument("Reshape cannot infer the missing input size for an "
                                          "empty tensor unless all specified input sizes are non-zero"));
      const int64 missing = input.NumElements() / product;
      OP_REQUIRES(ctx, product * missing == input.NumElements(),
                  errors::InvalidArgument("Input to reshape is a tensor with ", input.NumElements(),
                                          " values, but the requested shape requires a multiple of ",
                                          product));
      shape.set_dim(unknown_index, missing);
    }
    OP_REQUIRES(ctx, shape.num_elements() == input.NumElements(),
                errors::InvalidArgument("Input to reshape is a tensor with ", input.NumElements(),
                                        " values, but the requested shape has ",
                                        shape.num_elements()));
    Tensor output;
    OP_REQUIRES(ctx, output.CopyFrom(input, shape), errors::Internal("Reshape CopyFrom failed"));
    ctx->set_output(0, output);
  }
};
REGISTER_KERNEL_BUILDER(Name("Reshape").Device(DEVICE_GPU).HostMemory("shape"), ReshapeOp);

class ShapeOp : public OpKernel {
 public:
  explicit ShapeOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& inp = ctx->input(0);
    Tensor* out = nullptr;
    AllocatorAttributes host;
    host.set_on_host(true);
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({inp.dims()}), &out, host));
    for (int i = 0; i < inp.dims(); ++i) {
      if (output_type(0) == DT_INT64)
        out->data<int64>()[i] = inp.dim_size(i);
      else
        out->data<int32>()[i] = static_cast<int32>(inp.dim_size(i));
    }
  }
};
REGISTER_KERNEL_BUILDER(Name("Shape").Device(DEVICE_GPU).HostMemory("output"), ShapeOp);

// ---------------------------------------------------------------- VariableV2 / Assign
// variable_ops.h: the tensor lives with the (stateful) kernel, guarded by a mutex; the output is
// a reference to it.  It starts unallocated; Assign gives it storage (assign_op.h).
class VariableOp : public OpKernel {
 public:
  explicit VariableOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("shape", &shape_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("dtype", &dtype_));
    var_ = Tensor(dtype_, TensorShape({0}));
  }
  void Compute(OpKernelContext* ctx) override { ctx->set_output_ref(0, &mu_, &var_); }

 private:
  TensorShape shape_;
  DataType dtype_;
  std::mutex mu_;
  Tensor var_;
};
REGISTER_KERNEL_BUILDER(Name("VariableV2").Device(DEVICE_GPU), VariableOp);

class AssignOp : public OpKernel {
 public:
  explicit AssignOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("validate_shape", &validate_shape_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& rhs = ctx->input(1);
    ctx->forward_ref_input_to_ref_output(0, 0);
    std::mutex* mu = ctx->input_ref_mutex(0);
    std::lock_guard<std::mutex> l(*mu);
    Tensor old_lhs = ctx->mutable_input(0, /*lock_held=*/true);
    if (validate_shape_ && old_lhs.IsInitialized() && old_lhs.NumElements() > 0)
      OP_REQUIRES(ctx, old_lhs.shape().IsSameSize(rhs.shape()),
                  errors::InvalidArgument("Assign requires shapes of both tensors to match. "
                                          "lhs shape= ", old_lhs.shape().DebugString(),
                                          " rhs shape= ", rhs.shape().DebugString()));
    // assign_op.h: reuse the lhs buffer when the sizes agree, otherwise take a fresh copy.
    Tensor* lhs = (*params_inputs(ctx))[0].tensor;
    if (old_lhs.IsInitialized() && old_lhs.NumElements() == rhs.NumElements() &&
        old_lhs.NumElements() > 0) {
      lhs->CopyFrom(old_lhs, rhs.shape());
    } else {
      Tensor fresh(ctx->get_allocator(AllocatorAttributes()), rhs.dtype(), rhs.shape());
      OP_REQUIRES(ctx, fresh.IsInitialized(),
                  errors::ResourceExhausted("OOM when allocating variable ", name()));
      *lhs = fresh;
    }
    OP_REQUIRES_OK(ctx, FromAbi(b200_memcpy_d2d_async(lhs->raw_data(), rhs.raw_data(),
                                                      rhs.TotalBytes(), GetCudaStream(ctx)),
                                "Assign"));
  }

 private:
  // The ref input's Tensor object is the variable itself (TensorValue::tensor).
  static const std::vector<TensorValue>* params_inputs(OpKernelContext* ctx) {
    return ctx->ref_inputs_for_assign();
  }
  bool validate_shape_;
};
REGISTER_KERNEL_BUILDER(Name("Assign").Device(DEVICE_GPU), AssignOp);

// ---------------------------------------------------------------- AddN / Mul / Mean
template <typename T>
class AddNOp : public OpKernel {
 public:
  explicit AddNOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const int num = ctx->num_inputs();
    const Tensor& input0 = ctx->input(0);
    for (int i = 1; i < num; ++i)
      OP_REQUIRES(ctx, ctx->input(i).shape() == input0.shape(),
                  errors::InvalidArgument("Inputs to operation ", name(), " of type ",
                                          type_string(), " must have the same size and shape.  "
                                          "Input 0: ", input0.shape().DebugString(), " != input ",
                                          i, ": ", ctx->input(i).shape().DebugString()));
    Tensor* output = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, input0.shape(), &output));
    // One launch adds up to 8 operands; more inputs (aggregate_ops.cc:60-130 unrolls by 8 too) are
    // folded in passes of  out = out + next 7  -- same left-to-right summation order.
    const void* ptrs[8];
    int first = 0;
    while (first < num) {
      int k = 0;
      if (first > 0) ptrs[k++] = output->raw_data();
      while (k < 8 && first < num) ptrs[k++] = ctx->input(first++).raw_data();
      OP_REQUIRES_OK(ctx, FromAbi(b200_add_n(AbiType<T>::v, ptrs, k, output->raw_data(),
                                             input0.NumElements(), GetCudaStream(ctx)),
                                  "AddN"));
    }
  }
};

template <typename T>
class MulOp : public OpKernel {
 public:
  explicit MulOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor* x = &ctx->input(0);
    const Tensor* y = &ctx->input(1);
    if (x->NumElements() == 1 && y->NumElements() != 1) std::swap(x, y);  // commutative
    const bool scalar = y->NumElements() == 1 && x->NumElements() != 1;
    const bool both_single = x->NumElements() == 1 && y->NumElements() == 1;  // [] op [1,1] -> [1,1]
    OP_REQUIRES(ctx, scalar || both_single || x->shape() == y->shape(),
                errors::Unimplemented("Mul on B200 supports equal shapes or a scalar operand; got ",
                                      x->shape().DebugString(), " vs ", y->shape().DebugString()));
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(
                            0, both_single && y->dims() > x->dims() ? y->shape() : x->shape(), &out));
    OP_REQUIRES_OK(ctx, FromAbi(b200_mul(AbiType<T>::v, x->raw_data(), y->raw_data(),
                                         out->raw_data(), x->NumElements(), scalar,
                                         GetCudaStream(ctx)),
                                "Mul"));
  }
};

// Add with the same two shapes as Mul: equal shapes, or one operand a single element.
template <typename T>
class AddOp : public OpKernel {
 public:
  explicit AddOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor* x = &ctx->input(0);
    const Tensor* y = &ctx->input(1);
    if (x->NumElements() == 1 && y->NumElements() != 1) std::swap(x, y);  // commutative
    const bool scalar = y->NumElements() == 1 && x->NumElements() != 1;
    const bool both_single = x->NumElements() == 1 && y->NumElements() == 1;  // [] op [1,1] -> [1,1]
    OP_REQUIRES(ctx, scalar || both_single || x->shape() == y->shape(),
                errors::Unimplemented("Add on B200 supports equal shapes or a scalar operand; got ",
                                      x->shape().DebugString(), " vs ", y->shape().DebugString()));
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(
                            0, both_single && y->dims() > x->dims() ? y->shape() : x->shape(), &out));
    OP_REQUIRES_OK(ctx, FromAbi(b200_add(AbiType<T>::v, x->raw_data(), y->raw_data(),
                                         out->raw_data(), x->NumElements(), scalar,
                                         GetCudaStream(ctx)),
                                "Add"));
  }
};

// Sum / Mean (core/kernels/reduction_ops_common.h ReductionOp + ReductionHelper::Simplify):
// reduction_indices is a host-memory int32 / int64 vector, negative indices count from the back,
// duplicates are allowed, adjacent reduced / kept axes are collapsed.  The GPU kernel handles the
// patterns that collapse to [outer, REDUCE, inner] (one run of reduced axes: full reductions, row
// / column sums, NHWC channel means ...); alternating patterns return Unimplemented.
template <typename T, bool kMean>
class ReductionOp : public OpKernel {
 public:
  explicit ReductionOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("keep_dims", &keep_dims_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& data = ctx->input(0);
    const Tensor& axes = ctx->input(1);
    OP_REQUIRES(ctx, axes.dims() <= 1,
                errors::InvalidArgument("Expected scalar or vector as reduction_indices: ",
                                        axes.shape().DebugString()));
    const int rank = data.dims();
    std::vector<bool> reduced(rank, false);
    for (int64 i = 0; i < axes.NumElements(); ++i) {
      int64 a = axes.dtype() == DT_INT64 ? axes.flat<int64>()[i]
                                         : static_cast<int64>(axes.flat<int32>()[i]);
      OP_REQUIRES(ctx, a >= -rank && a < rank,
                  errors::InvalidArgument("Invalid reduction dimension (", a, " for input with ",
                                          rank, " dimension(s)"));
      reduced[(a + rank) % rank] = true;
    }
    TensorShape out_shape;
    for (int d = 0; d < rank; ++d) {
      if (!reduced[d])
        out_shape.AddDim(data.dim_size(d));
      else if (keep_dims_)
        out_shape.AddDim(1);
    }
    // collapse: runs of equal `reduced` flags (size-1 axes join either neighbour)
    std::vector<std::pair<bool, int64>> runs;
    for (int d = 0; d < rank; ++d) {
      if (data.dim_size(d) == 1) continue;
      if (!runs.empty() && runs.back().first == reduced[d])
        runs.back().second *= data.dim_size(d);
      else
        runs.push_back({reduced[d], data.dim_size(d)});
    }
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, out_shape, &out));
    if (out->NumElements() == 0) return;
    int64 outer = 1, reduce = 1, inner = 1;
    int n_reduced_runs = 0;
    for (const auto& r : runs) n_reduced_runs += r.first ? 1 : 0;
    OP_REQUIRES(ctx, n_reduced_runs <= 1 && runs.size() <= 3,
                errors::Unimplemented(type_string(), " on B200 reduces one contiguous run of axes "
                                      "(got input ", data.shape().DebugString(), ")"));
    bool seen = false;
    for (const auto& r : runs) {
      if (r.first) { reduce = r.second; seen = true; }
      else if (!seen) outer *= r.second;
      else inner *= r.second;
    }
    if (n_reduced_runs == 0 && !runs.empty()) {  // nothing (but size-1 axes) reduced: a copy
      outer = data.NumElements();
      inner = 1;
    }
    const float scale = kMean ? (reduce > 0 ? 1.0f / static_cast<float>(reduce) : 0.f) : 1.0f;
    OP_REQUIRES_OK(ctx, FromAbi(b200_reduce(AbiType<T>::v, data.raw_data(), out->raw_data(), outer,
                                            reduce, inner, scale, GetCudaStream(ctx)),
                                type_string().c_str()));
  }

 private:
  bool keep_dims_;
};
#define REGISTER_REDUCTIONS(T)                                                              \
  REGISTER_KERNEL_BUILDER(                                                                  \
      Name("Mean").Device(DEVICE_GPU).TypeConstraint<T>("T").HostMemory("reduction_indices"), \
      ReductionOp<T, true>);                                                                \
  REGISTER_KERNEL_BUILDER(                                                                  \
      Name("Sum").Device(DEVICE_GPU).TypeConstraint<T>("T").HostMemory("reduction_indices"),  \
      ReductionOp<T, false>);
REGISTER_B200_FLOAT_TYPES(REGISTER_REDUCTIONS)
#undef REGISTER_REDUCTIONS

// ---------------------------------------------------------------- ApplyGradientDescent
// training_ops.cc:369-412: var -= alpha * delta on the variable's own buffer; out = ref(var).
template <typename T>
class ApplyGradientDescentOp : public OpKernel {
 public:
  explicit ApplyGradientDescentOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    Tensor var = ctx->mutable_input(0, false);
    OP_REQUIRES(ctx, var.IsInitialized() && var.NumElements() > 0,
                errors::FailedPrecondition("Attempting to use uninitialized variables: ",
                                           def().input.empty() ? name() : def().input[0]));
    const Tensor& alpha = ctx->input(1);
    OP_REQUIRES(ctx, TensorShapeUtils::IsScalar(alpha.shape()),
                errors::InvalidArgument("alpha is not a scalar: ", alpha.shape().DebugString()));
    const Tensor& delta = ctx->input(2);
    OP_REQUIRES(ctx, var.shape().IsSameSize(delta.shape()),
                errors::InvalidArgument("var and delta do not have the same shape",
                                        var.shape().DebugString(), " ",
                                        delta.shape().DebugString()));
    OP_REQUIRES_OK(ctx, FromAbi(b200_apply_gradient_descent(AbiType<T>::v, var.raw_data(),
                                                            alpha.raw_data(), delta.raw_data(),
                                                            var.NumElements(), GetCudaStream(ctx)),
                                "ApplyGradientDescent"));
    ctx->forward_ref_input_to_ref_output(0, 0);
  }
};

// N ApplyGradientDescent nodes that sit next to each other in the plan, run as one launch
// (direct_session.cc FuseApplyGradientDescent).  Same checks and arithmetic per variable.
template <typename T>
class MultiApplyGradientDescentOp : public OpKernel {
 public:
  explicit MultiApplyGradientDescentOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("N", &n_));
  }
  void Compute(OpKernelContext* ctx) override {
    const int n = static_cast<int>(n_);
    std::vector<void*> vars(n);
    std::vector<const void*> alphas(n), deltas(n);
    std::vector<int64_t> counts(n);
    for (int i = 0; i < n; ++i) {
      Tensor var = ctx->mutable_input(i, false);
      OP_REQUIRES(ctx, var.IsInitialized() && var.NumElements() > 0,
                  errors::FailedPrecondition("Attempting to use uninitialized variables: ",
                                             def().input.size() > static_cast<size_t>(i)
                                                 ? def().input[i]
                                                 : name()));
      const Tensor& alpha = ctx->input(n + i);
      OP_REQUIRES(ctx, TensorShapeUtils::IsScalar(alpha.shape()),
                  errors::InvalidArgument("alpha is not a scalar: ", alpha.shape().DebugString()));
      const Tensor& delta = ctx->input(2 * n + i);
      OP_REQUIRES(ctx, var.shape().IsSameSize(delta.shape()),
                  errors::InvalidArgument("var and delta do not have the same shape",
                                          var.shape().DebugString(), " ",
                                          delta.shape().DebugString()));
      vars[i] = var.raw_data();
      alphas[i] = alpha.raw_data();
      deltas[i] = delta.raw_data();
      counts[i] = var.NumElements();
    }
    OP_REQUIRES_OK(ctx, FromAbi(b200_apply_gradient_descent_multi(AbiType<T>::v, n, vars.data(),
                                                                  alphas.data(), deltas.data(),
                                                                  counts.data(), GetCudaStream(ctx)),
                                "ApplyGradientDescent"));
    for (int i = 0; i < n; ++i) ctx->forward_ref_input_to_ref_output(i, i);
  }

 private:
  int64 n_;
};

// ---------------------------------------------------------------- B200AllReduce (additive)
// One ncclAllReduce(sum) on the compute stream over the referenced buffer, in place, followed by
// an optional scale (1/replicas for a gradient average).  No host synchronisation.
template <typename T>
class B200AllReduceOp : public OpKernel {
 public:
  explicit B200AllReduceOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("scale", &scale_));
  }
  void Compute(OpKernelContext* ctx) override {
    Tensor data = ctx->mutable_input(0, false);
    ctx->forward_ref_input_to_ref_output(0, 0);
    BaseGPUDevice* dev = dynamic_cast<BaseGPUDevice*>(ctx->device());
    OP_REQUIRES(ctx, dev != nullptr, errors::Internal("B200AllReduce needs a GPU device"));
    if (dev->num_replicas() <= 1 || dev->collective_comm() == nullptr) {
      if (scale_ != 1.0f)
        OP_REQUIRES_OK(ctx, FromAbi(b200_scale(AbiType<T>::v, data.raw_data(), scale_,
                                               data.raw_data(), data.NumElements(),
                                               GetCudaStream(ctx)),
                                    "B200AllReduce scale"));
      return;
    }
    OP_REQUIRES_OK(ctx, FromAbi(b200_nccl_all_reduce_sum(AbiType<T>::v, data.raw_data(),
                                                         data.raw_data(), data.NumElements(),
                                                         dev->collective_comm(),
                                                         GetCudaStream(ctx)),
                                "ncclAllReduce"));
    if (scale_ != 1.0f)
      OP_REQUIRES_OK(ctx, FromAbi(b200_scale(AbiType<T>::v, data.raw_data(), scale_,
                                             data.raw_data(), data.NumElements(),
                                             GetCudaStream(ctx)),
                                  "B200AllReduce scale"));
  }

 private:
  float scale_;
};

// N gradient tensors (one bucket) reduced across replicas by ONE collective: the tensors are
// gathered into a contiguous scratch arena, reduced by a single ncclAllReduce and scattered
// back (4 MB + 4 KB buckets: the two extra copies cost ~3 us each at HBM speed, a second
// collective costs its full NVLink latency).  B200TF_ALLREDUCE_PACK=0 keeps the tensors in place
// and issues one ncclAllReduce per tensor inside ncclGroupStart/End instead.  scale == 1/replicas
// maps to ncclAvg; any other scale runs as ncclSum followed by a scale kernel.  The kernel uses
// whatever stream its DeviceContext carries: the executor places it on the device's collective
// stream so the exchange overlaps the remaining backward kernels.  No host synchronisation.
template <typename T>
class B200AllReduceNOp : public OpKernel {
 public:
  explicit B200AllReduceNOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("scale", &scale_));
  }
  void Compute(OpKernelContext* ctx) override {
    BaseGPUDevice* dev = dynamic_cast<BaseGPUDevice*>(ctx->device());
    OP_REQUIRES(ctx, dev != nullptr, errors::Internal("B200AllReduceN needs a GPU device"));
    const int n = ctx->num_inputs();
    void* stream = GetCudaStream(ctx);
    const bool collective = dev->num_replicas() > 1 && dev->collective_comm() != nullptr;
    const bool average =
        collective && std::fabs(scale_ * dev->num_replicas() - 1.0f) < 1e-6f;
    std::vector<Tensor*> outs(n, nullptr);
    for (int i = 0; i < n; ++i)  // reduce in place when this op is the buffer's only user
      OP_REQUIRES_OK(ctx, ctx->forward_input_or_allocate_output({i}, i, ctx->input(i).shape(),
                                                                &outs[i]));
    // Inputs produced into one gradient arena (direct_session.cc PlanGradientArenas): consecutive
    // 256-byte-aligned windows of one root buffer -> ONE in-place collective, no copies.
    bool contiguous = collective && ctx->input(0).buffer() != nullptr;
    size_t span = 0;
    for (int i = 0; contiguous && i < n; ++i) {
      const Tensor& t = ctx->input(i);
      contiguous = t.buffer() != nullptr &&
                   t.buffer()->root_buffer() == ctx->input(0).buffer()->root_buffer() &&
                   static_cast<char*>(t.raw_data()) ==
                       static_cast<char*>(ctx->input(0).raw_data()) + span;
      span += (t.TotalBytes() + 255) / 256 * 256;
    }
    static const bool pack_enabled = [] {
      const char* v = getenv("B200TF_ALLREDUCE_PACK");
      return v == nullptr || std::strcmp(v, "0") != 0;
    }();
    if (contiguous) {
      void* base = ctx->input(0).raw_data();
      const int64 total = static_cast<int64>(span / sizeof(T));  // padding included: inside the arena
      const long long peer_offset = dev->PeerArenaOffset(base);
      // fp32 on either peer backend; bfloat16 when the arena lives in NVSwitch multicast memory
      const bool peer_dtype = std::is_same<T, float>::value ||
                              std::strcmp(b200_peer_arena_backend(), "nvls") == 0;
      if (peer_offset >= 0 && average && peer_dtype) {
        // the arena lives in NVLink peer memory: one kernel of peer loads instead of NCCL (its
        // CTAs fit beside resident GEMM CTAs, so it also runs well on the collective stream)
        // NVLS: 16 CTAs move 4 MB as fast as 128 do (the switch reduces; tools/arena_probe.py)
        // and leave the SMs to the backward pass they run beside.  0 = the kernel's own default.
        static const int ctas = [] {
          const char* v = getenv("B200TF_PEER_CTAS");
          if (v) return std::atoi(v);
          return std::strcmp(b200_peer_arena_backend(), "nvls") == 0 ? 16 : 0;
        }();
        OP_REQUIRES_OK(ctx, FromAbi(b200_peer_all_reduce(dev->peer_arena(), AbiType<T>::v,
                                                         static_cast<size_t>(peer_offset), total, 1,
                                                         ctas, stream),
                                    "b200_peer_all_reduce"));
      } else {
        OP_REQUIRES_OK(ctx, FromAbi(b200_nccl_all_reduce(AbiType<T>::v, base, base, total, average,
                                                         dev->collective_comm(), stream),
                                    "ncclAllReduce"));
      }
      for (int i = 0; i < n; ++i)  // an input that could not be forwarded gets its copy
        if (outs[i]->raw_data() != ctx->input(i).raw_data())
          OP_REQUIRES_OK(ctx, FromAbi(b200_memcpy_d2d_async(outs[i]->raw_data(),
                                                            ctx->input(i).raw_data(),
                                                            outs[i]->TotalBytes(), stream),
                                      "copy"));
    } else if (collective && n > 1 && pack_enabled) {
      int64 total = 0;
      std::vector<int64> offset(n);
      for (int i = 0; i < n; ++i) {  // 256-byte aligned slots (in elements)
        offset[i] = total;
        const int64 per = 256 / static_cast<int64>(sizeof(T));
        total += (ctx->input(i).NumElements() + per - 1) / per * per;
      }
      Tensor arena;
      OP_REQUIRES_OK(ctx, ctx->allocate_temp(DataTypeToEnum<T>::value, TensorShape({total}), &arena));
      char* base = static_cast<char*>(arena.raw_data());
      for (int i = 0; i < n; ++i)
        if (ctx->input(i).NumElements() > 0)
          OP_REQUIRES_OK(ctx, FromAbi(b200_memcpy_d2d_async(base + offset[i] * sizeof(T),
                                                            ctx->input(i).raw_data(),
                                                            ctx->input(i).TotalBytes(), stream),
                                      "gather"));
      OP_REQUIRES_OK(ctx, FromAbi(b200_nccl_all_reduce(AbiType<T>::v, base, base, total, average,
                                                       dev->collective_comm(), stream),
                                  "ncclAllReduce"));
      for (int i = 0; i < n; ++i)
        if (outs[i]->NumElements() > 0)
          OP_REQUIRES_OK(ctx, FromAbi(b200_memcpy_d2d_async(outs[i]->raw_data(),
                                                            base + offset[i] * sizeof(T),
                                                            outs[i]->TotalBytes(), stream),
                                      "scatter"));
    } else if (collective) {
      OP_REQUIRES_OK(ctx, FromAbi(b200_nccl_group_start(), "ncclGroupStart"));
      Status s;
      for (int i = 0; i < n; ++i)
        s.Update(FromAbi(b200_nccl_all_reduce(AbiType<T>::v, ctx->input(i).raw_data(),
                                              outs[i]->raw_data(), ctx->input(i).NumElements(),
                                              average, dev->collective_comm(), stream),
                         "ncclAllReduce"));
      Status e = FromAbi(b200_nccl_group_end(), "ncclGroupEnd");
      OP_REQUIRES_OK(ctx, s);
      OP_REQUIRES_OK(ctx, e);
    } else {
      for (int i = 0; i < n; ++i)
        if (outs[i]->raw_data() != ctx->input(i).raw_data())
          OP_REQUIRES_OK(ctx, FromAbi(b200_memcpy_d2d_async(outs[i]->raw_data(),
                                                            ctx->input(i).raw_data(),
                                                            outs[i]->TotalBytes(), stream),
                                      "copy"));
    }
    if (!average && scale_ != 1.0f)
      for (int i = 0; i < n; ++i)
        OP_REQUIRES_OK(ctx, FromAbi(b200_scale(AbiType<T>::v, outs[i]->raw_data(), scale_,
                                               outs[i]->raw_data(), outs[i]->NumElements(), stream),
                                    "scale"));
  }

 private:
  float scale_;
};

#define REGISTER_GPU(T)                                                                          \
  REGISTER_KERNEL_BUILDER(Name("B200AllReduceN").Device(DEVICE_GPU).TypeConstraint<T>("T"),      \
                          B200AllReduceNOp<T>);                                                  \
  REGISTER_KERNEL_BUILDER(Name("AddN").Device(DEVICE_GPU).TypeConstraint<T>("T"), AddNOp<T>);    \
  REGISTER_KERNEL_BUILDER(Name("Mul").Device(DEVICE_GPU).TypeConstraint<T>("T"), MulOp<T>);      \
  REGISTER_KERNEL_BUILDER(Name("Add").Device(DEVICE_GPU).TypeConstraint<T>("T"), AddOp<T>);      \
  REGISTER_KERNEL_BUILDER(Name("ApplyGradientDescent").Device(DEVICE_GPU).TypeConstraint<T>("T"),\
                          ApplyGradientDescentOp<T>);                                            \
  REGISTER_KERNEL_BUILDER(                                                                       \
      Name("_MultiApplyGradientDescent").Device(DEVICE_GPU).TypeConstraint<T>("T"),              \
      MultiApplyGradientDescentOp<T>);                                                           \
  REGISTER_KERNEL_BUILDER(Name("B200AllReduce").Device(DEVICE_GPU).TypeConstraint<T>("T"),       \
                          B200AllReduceOp<T>);
REGISTER_B200_FLOAT_TYPES(REGISTER_GPU)
#undef REGISTER_GPU

}  // namespace tensorflow
