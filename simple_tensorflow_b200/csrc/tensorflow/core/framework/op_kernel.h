// OpKernel / OpKernelConstruction / OpKernelContext / REGISTER_KERNEL_BUILDER -- the kernel half
// of the plugin surface, source-compatible with the members the hot-path kernels use in the
// reference's core/framework/op_kernel.h (OpKernel :71-167, OpKernelConstruction :216-353,
// OpKernelContext :458-1095, registration :1180-1240, OP_REQUIRES :1436-1449) and
// core/framework/kernel_def_builder.h:40-66.
#ifndef B200TF_CORE_FRAMEWORK_OP_KERNEL_H_
#define B200TF_CORE_FRAMEWORK_OP_KERNEL_H_

#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "tensorflow/core/framework/device_base.h"
#include "tensorflow/core/framework/node_def.h"
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/tensor.h"

namespace tensorflow {

class OpKernelConstruction;
class OpKernelContext;

// kernel_def.proto: {op, device_type, constraint[], host_memory_arg[], label}
struct KernelDef {
  struct AttrConstraint {
    std::string name;
    std::vector<DataType> allowed_values;
  };
  std::string op;
  std::string device_type;
  std::vector<AttrConstraint> constraint;
  std::vector<std::string> host_memory_arg;
  std::string label;
};

class KernelDefBuilder {
 public:
  explicit KernelDefBuilder(const char* op_name) { def_.op = op_name; }
  KernelDefBuilder& Device(const char* device_type) {
    def_.device_type = device_type;
    return *this;
  }
  template <typename T>
  KernelDefBuilder& TypeConstraint(const char* attr_name) {
    return TypeConstraint(attr_name, DataTypeToEnum<T>::v());
  }
  KernelDefBuilder& TypeConstraint(const char* attr_name, DataType allowed) {
    for (auto& c : def_.constraint)
      if (c.name == attr_name) {
        c.allowed_values.push_back(allowed);
        return *this;
      }
    def_.constraint.push_back({attr_name, {allowed}});
    return *this;
  }
  KernelDefBuilder& HostMemory(const char* arg_name) {
    def_.host_memory_arg.push_back(arg_name);
    return *this;
  }
  KernelDefBuilder& Label(const char* label) {
    def_.label = label;
    return *this;
  }
  const KernelDef* Build() { return new KernelDef(def_); }

 private:
  KernelDef def_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction* context);
  virtual ~OpKernel() {}
  // Must be thread-safe; on DEVICE_GPU it must only ENQUEUE work on the op's stream and
  // return (op_kernel.h:78-100, common_runtime/gpu/gpu_device.cc:337-399).
  virtual void Compute(OpKernelContext* context) = 0;
  virtual bool IsExpensive() { return expensive_; }

  const NodeDef& def() const { return def_; }
  const std::string& name() const { return def_.name; }
  const std::string& type_string() const { return def_.op; }
  int num_inputs() const { return static_cast<int>(input_types_.size()); }
  DataType input_type(int i) const { return input_types_[i]; }
  const DataTypeVector& input_types() const { return input_types_; }
  const MemoryTypeVector& input_memory_types() const { return input_memory_types_; }
  int num_outputs() const { return static_cast<int>(output_types_.size()); }
  DataType output_type(int o) const { return output_types_[o]; }
  const DataTypeVector& output_types() const { return output_types_; }
  const MemoryTypeVector& output_memory_types() const { return output_memory_types_; }
  bool input_is_ref(int i) const { return input_is_ref_[i]; }
  bool output_is_ref(int i) const { return output_is_ref_[i]; }

 private:
  const NodeDef def_;
  const DataTypeVector input_types_;
  const MemoryTypeVector input_memory_types_;
  const DataTypeVector output_types_;
  const MemoryTypeVector output_memory_types_;
  std::vector<bool> input_is_ref_, output_is_ref_;
  bool expensive_;
};

class OpKernelConstruction {
 public:
  OpKernelConstruction(DeviceType device_type, DeviceBase* device, Allocator* allocator,
                       const NodeDef* node_def, const OpDef* op_def,
                       const DataTypeVector& input_types,
                       const MemoryTypeVector& input_memory_types,
                       const DataTypeVector& output_types,
                       const MemoryTypeVector& output_memory_types, Status* status)
      : device_type_(device_type), device_(device), allocator_(allocator), def_(node_def),
        op_def_(op_def), input_types_(input_types), input_memory_types_(input_memory_types),
        output_types_(output_types), output_memory_types_(output_memory_types), status_(status) {}

  // op_kernel.h:308-309: attrs are read in the kernel constructor.
  template <class T>
  Status GetAttr(const std::string& attr_name, T* value) const {
    return GetNodeAttr(*def_, attr_name, value);
  }
  void SetStatus(const Status& s) { status_->Update(s); }
  const Status& status() const { return *status_; }
  void CtxFailure(Status s) { SetStatus(s); }
  void CtxFailureWithWarning(Status s) { SetStatus(s); }
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp);

  const NodeDef& def() const { return *def_; }
  const OpDef& op_def() const { return *op_def_; }
  const DeviceType& device_type() const { return device_type_; }
  DeviceBase* device() const { return device_; }
  int num_inputs() const { return static_cast<int>(input_types_.size()); }
  DataType input_type(int i) const { return input_types_[i]; }
  const DataTypeVector& input_types() const { return input_types_; }
  const MemoryTypeVector& input_memory_types() const { return input_memory_types_; }
  int num_outputs() const { return static_cast<int>(output_types_.size()); }
  DataType output_type(int i) const { return output_types_[i]; }
  const DataTypeVector& output_types() const { return output_types_; }
  const MemoryTypeVector& output_memory_types() const { return output_memory_types_; }

 private:
  const DeviceType device_type_;
  DeviceBase* const device_;
  Allocator* allocator_;
  const NodeDef* def_;
  const OpDef* op_def_;
  DataTypeVector input_types_;
  MemoryTypeVector input_memory_types_;
  DataTypeVector output_types_;
  MemoryTypeVector output_memory_types_;
  Status* status_;
};

// op_kernel.h:446-456: an input/output slot; `tensor` is owned by the executor (inputs), by the
// context (outputs until released) or by a variable kernel (ref).
struct TensorValue {
  TensorValue() : mutex_if_ref(nullptr), tensor(nullptr) {}
  TensorValue(Tensor* t) : mutex_if_ref(nullptr), tensor(t) {}  // NOLINT
  TensorValue(std::mutex* mu, Tensor* t) : mutex_if_ref(mu), tensor(t) {}
  Tensor* operator->() const { return tensor; }
  bool is_ref() const { return mutex_if_ref != nullptr; }
  std::mutex* mutex_if_ref;
  Tensor* tensor;
};

class OpKernelContext {
 public:
  // op_kernel.h:466-560 (subset).
  struct Params {
    int64 step_id = 0;
    OpKernel* op_kernel = nullptr;
    DeviceBase* device = nullptr;
    const std::vector<TensorValue>* inputs = nullptr;
    const std::vector<AllocatorAttributes>* input_alloc_attrs = nullptr;
    const AllocatorAttributes* output_attr_array = nullptr;
    DeviceContext* op_device_context = nullptr;
    // op_kernel.h:508: a kernel that runs on a stream other than the compute stream has the
    // tensors it touched recorded, so the executor can keep them alive until that stream is done.
    bool record_tensor_accesses = false;
    // Per output: a buffer the executor wants the kernel's allocate_output() to use (a slice of
    // a gradient arena, the job of later TensorFlow's ScopedAllocator).  Used when dtype and
    // element count match the request; otherwise allocation proceeds as usual.
    const std::vector<Tensor>* preallocated_outputs = nullptr;
  };
  explicit OpKernelContext(Params* params);
  ~OpKernelContext();

  int64 step_id() const { return params_->step_id; }
  const OpKernel& op_kernel() const { return *params_->op_kernel; }

  // ---- inputs (op_kernel.h:586-650)
  int num_inputs() const { return static_cast<int>(params_->inputs->size()); }
  DataType input_dtype(int index) const { return params_->op_kernel->input_type(index); }
  const Tensor& input(int index);
  bool has_input(int index) const { return (*params_->inputs)[index].tensor != nullptr; }
  // Ref inputs (variables): op_kernel.h:616-650.
  Tensor mutable_input(int index, bool lock_held);
  std::mutex* input_ref_mutex(int index) { return (*params_->inputs)[index].mutex_if_ref; }
  void forward_ref_input_to_ref_output(int input_index, int output_index);
  // The slot vector itself: AssignOp replaces the variable's Tensor through its ref slot.
  const std::vector<TensorValue>* ref_inputs_for_assign() const { return params_->inputs; }

  // ---- outputs (op_kernel.h:683-815)
  int num_outputs() const { return static_cast<int>(outputs_.size()); }
  DataType expected_output_dtype(int index) const { return params_->op_kernel->output_type(index); }
  Status allocate_output(int index, const TensorShape& shape, Tensor** tensor);
  Status allocate_output(int index, const TensorShape& shape, Tensor** tensor,
                         AllocatorAttributes attr);
  // In-place reuse: succeeds only when the input buffer's refcount is one and dtype / size /
  // memory type match (op_kernel.cc:402-439); otherwise allocates.
  Status forward_input_or_allocate_output(const std::vector<int>& candidate_input_indices,
                                          int output_index, const TensorShape& output_shape,
                                          Tensor** output);
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp) {
    return allocate_temp(type, shape, out_temp, AllocatorAttributes());
  }
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp,
                       AllocatorAttributes allocator_attr);
  // op_kernel.h:989: the tensors recorded under Params::record_tensor_accesses.
  void retrieve_accessed_tensors(std::vector<Tensor>* out_vector) {
    out_vector->swap(referenced_tensors_);
    referenced_tensors_.clear();
  }
  void set_output(int index, const Tensor& tensor);
  void set_output_ref(int index, std::mutex* mu, Tensor* tensor_for_ref);
  Tensor* mutable_output(int index) { return outputs_[index].tensor; }
  // Transfers ownership of the output Tensor object to the caller (op_kernel.h:1010,1319-1325).
  TensorValue release_output(int index);
  AllocatorAttributes output_alloc_attr(int index) const {
    return params_->output_attr_array ? params_->output_attr_array[index] : AllocatorAttributes();
  }

  // ---- device (op_kernel.h:875-882; device_base.h:67-72)
  DeviceBase* device() const { return params_->device; }
  template <typename T = DeviceContext>
  T* op_device_context() {
    return static_cast<T*>(params_->op_device_context);
  }
  Allocator* get_allocator(AllocatorAttributes attr) { return params_->device->GetAllocator(attr); }

  // ---- status (op_kernel.h:1030-1040)
  void SetStatus(const Status& status) { status_.Update(status); }
  const Status& status() const { return status_; }
  void CtxFailure(Status s) { SetStatus(s); }
  void CtxFailureWithWarning(Status s) { SetStatus(s); }

 private:
  Status allocate_tensor(DataType type, const TensorShape& shape, Tensor* out_tensor,
                         AllocatorAttributes attr);
  const Tensor* preallocated_output(int index) const {
    const std::vector<Tensor>* p = params_->preallocated_outputs;
    if (p == nullptr || index >= static_cast<int>(p->size())) return nullptr;
    return (*p)[index].buffer() != nullptr ? &(*p)[index] : nullptr;
  }
  Params* params_;
  Status status_;
  std::vector<TensorValue> outputs_;
  std::vector<bool> output_owned_;
  std::vector<Tensor> referenced_tensors_;  // only filled under record_tensor_accesses
};

// ------------------------------------------------------------------ registration
typedef OpKernel* (*OpKernelFactory)(OpKernelConstruction*);

namespace kernel_factory {
class OpKernelRegistrar {
 public:
  OpKernelRegistrar(const KernelDef* kernel_def, const std::string& kernel_class_name,
                    OpKernelFactory factory);
};
}  // namespace kernel_factory

// op_kernel.h:1180-1198
#define REGISTER_KERNEL_BUILDER(kernel_builder, ...) \
  REGISTER_KERNEL_BUILDER_UNIQ_HELPER(__COUNTER__, kernel_builder, __VA_ARGS__)
#define REGISTER_KERNEL_BUILDER_UNIQ_HELPER(ctr, kernel_builder, ...) \
  REGISTER_KERNEL_BUILDER_UNIQ(ctr, kernel_builder, __VA_ARGS__)
#define REGISTER_KERNEL_BUILDER_UNIQ(ctr, kernel_builder, ...)                              \
  static ::tensorflow::kernel_factory::OpKernelRegistrar registrar__body__##ctr##__object(  \
      ::tensorflow::register_kernel::kernel_builder.Build(), #__VA_ARGS__,                  \
      [](::tensorflow::OpKernelConstruction* context) -> ::tensorflow::OpKernel* {          \
        return new __VA_ARGS__(context);                                                    \
      });

namespace register_kernel {
class Name : public KernelDefBuilder {
 public:
  explicit Name(const char* op) : KernelDefBuilder(op) {}
};
}  // namespace register_kernel

// op_kernel.cc:998-1064: OpDef lookup -> ValidateNodeDef -> FindKernelRegistration ->
// MemoryTypesForNode -> factory.  Exactly one registration may match (:879-907).
Status CreateOpKernel(DeviceType device_type, DeviceBase* device, Allocator* allocator,
                      const NodeDef& node_def, std::unique_ptr<OpKernel>* kernel);
Status SupportedDeviceTypesForNode(const std::vector<DeviceType>& prioritized_types,
                                   const NodeDef& def, std::vector<DeviceType>* device_types);
// Lists "Op:DEVICE:label" keys of every registered kernel (for tests / debugging).
std::vector<std::string> RegisteredKernelKeys();

// op_kernel.h:1436-1449
#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->CtxFailure((STATUS));    \
      return;                         \
    }                                 \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)                   \
  do {                                             \
    ::tensorflow::Status _s(__VA_ARGS__);          \
    if (!_s.ok()) {                                \
      (CTX)->CtxFailureWithWarning(_s);            \
      return;                                      \
    }                                              \
  } while (0)

}  // namespace tensorflow
#endif
