// See op_kernel.h.  Registry semantics follow the reference's core/framework/op_kernel.cc:
// key "Op:DEVICE:label" (:794-798), single-match rule (:879-907), CreateOpKernel (:998-1064),
// first-error-wins status (:247-249), output ownership (:219-226, 576-591).
#include "tensorflow/core/framework/op_kernel.h"

#include <algorithm>
#include <map>

namespace tensorflow {

// ------------------------------------------------------------------ OpKernel
static std::vector<bool> RefFlags(const std::vector<OpDef::ArgDef>& args, const NodeDef& node) {
  std::vector<bool> out;
  for (const auto& a : args) {
    int64 count = 1;
    if (!a.number_attr.empty()) GetNodeAttr(node, a.number_attr, &count);
    for (int64 i = 0; i < count; ++i) out.push_back(a.is_ref);
  }
  return out;
}

OpKernel::OpKernel(OpKernelConstruction* context)
    : def_(context->def()),
      input_types_(context->input_types()),
      input_memory_types_(context->input_memory_types()),
      output_types_(context->output_types()),
      output_memory_types_(context->output_memory_types()),
      input_is_ref_(RefFlags(context->op_def().input_arg, context->def())),
      output_is_ref_(RefFlags(context->op_def().output_arg, context->def())) {
  // Kernels on GPU only enqueue work: always "inexpensive" so one host thread walks the
  // partition in ready order (op_kernel.cc:97-99).
  expensive_ = context->device_type() != DeviceType(DEVICE_GPU);
}

Status OpKernelConstruction::allocate_temp(DataType type, const TensorShape& shape,
                                           Tensor* out_temp) {
  Tensor t(allocator_, type, shape);
  if (!t.IsInitialized())
    return errors::ResourceExhausted("OOM when allocating temporary tensor with shape",
                                     shape.DebugString());
  *out_temp = std::move(t);
  return Status::OK();
}

// ------------------------------------------------------------------ OpKernelContext
OpKernelContext::OpKernelContext(Params* params)
    : params_(params),
      outputs_(params->op_kernel->num_outputs()),
      output_owned_(params->op_kernel->num_outputs(), false) {}

OpKernelContext::~OpKernelContext() {
  for (size_t i = 0; i < outputs_.size(); ++i)
    if (output_owned_[i] && outputs_[i].tensor) delete outputs_[i].tensor;
}

const Tensor& OpKernelContext::input(int index) { return *(*params_->inputs)[index].tensor; }

Tensor OpKernelContext::mutable_input(int index, bool lock_held) {
  const TensorValue& v = (*params_->inputs)[index];
  if (lock_held || !v.mutex_if_ref) return *v.tensor;
  std::lock_guard<std::mutex> l(*v.mutex_if_ref);
  return *v.tensor;
}

void OpKernelContext::forward_ref_input_to_ref_output(int input_index, int output_index) {
  const TensorValue& v = (*params_->inputs)[input_index];
  set_output_ref(output_index, v.mutex_if_ref, v.tensor);
}

Status OpKernelContext::allocate_tensor(DataType type, const TensorShape& shape,
                                        Tensor* out_tensor, AllocatorAttributes attr) {
  Allocator* a = get_allocator(attr);
  Tensor t(a, type, shape);
  if (!t.IsInitialized())
    return errors::ResourceExhausted("OOM when allocating tensor with shape", shape.DebugString(),
                                     " on allocator ", a->Name());
  *out_tensor = std::move(t);
  return Status::OK();
}

Status OpKernelContext::allocate_output(int index, const TensorShape& shape, Tensor** tensor) {
  return allocate_output(index, shape, tensor, output_alloc_attr(index));
}

Status OpKernelContext::allocate_output(int index, const TensorShape& shape, Tensor** output,
                                        AllocatorAttributes attr) {
  if (index < 0 || index >= num_outputs())
    return errors::Internal("allocate_output: bad output index ", index);
  Tensor* t = new Tensor();
  const Tensor* pre = preallocated_output(index);
  if (pre != nullptr && pre->dtype() == params_->op_kernel->output_type(index) &&
      pre->NumElements() == shape.num_elements() && shape.num_elements() > 0) {
    t->CopyFrom(*pre, shape);
  } else {
    Status s = allocate_tensor(params_->op_kernel->output_type(index), shape, t, attr);
    if (!s.ok()) {
      delete t;
      return s;
    }
  }
  if (output_owned_[index] && outputs_[index].tensor) delete outputs_[index].tensor;
  outputs_[index] = TensorValue(t);
  output_owned_[index] = true;
  *output = t;
  return Status::OK();
}

Status OpKernelContext::forward_input_or_allocate_output(
    const std::vector<int>& candidate_input_indices, int output_index,
    const TensorShape& output_shape, Tensor** output) {
  if (preallocated_output(output_index) != nullptr)  // the executor chose this output's home
    return allocate_output(output_index, output_shape, output);
  for (int input_index : candidate_input_indices) {
    const TensorValue& v = (*params_->inputs)[input_index];
    if (v.is_ref() || v.tensor == nullptr) continue;
    const Tensor& in = *v.tensor;
    const bool same_mem = params_->op_kernel->input_memory_types()[input_index] ==
                          params_->op_kernel->output_memory_types()[output_index];
    if (in.dtype() == expected_output_dtype(output_index) &&
        in.NumElements() == output_shape.num_elements() && same_mem && in.buffer() != nullptr &&
        in.buffer()->RefCountIsOne()) {
      Tensor* t = new Tensor();
      t->CopyFrom(in, output_shape);
      if (output_owned_[output_index] && outputs_[output_index].tensor)
        delete outputs_[output_index].tensor;
      outputs_[output_index] = TensorValue(t);
      output_owned_[output_index] = true;
      *output = t;
      return Status::OK();
    }
  }
  return allocate_output(output_index, output_shape, output);
}

Status OpKernelContext::allocate_temp(DataType type, const TensorShape& shape, Tensor* out_temp,
                                      AllocatorAttributes attr) {
  // On the compute stream the arena's stream-ordered reuse keeps the scratch valid until the
  // enqueued kernels have consumed it (gpu_device.cc:266-271).  A kernel running on another
  // stream has its scratch recorded so the executor can hold it until that stream has finished
  // (the reference's record_tensor_accesses / EventMgr::ThenDeleteTensors, gpu_device.cc:372-385).
  Status s = allocate_tensor(type, shape, out_temp, attr);
  if (s.ok() && params_->record_tensor_accesses) referenced_tensors_.push_back(*out_temp);
  return s;
}

void OpKernelContext::set_output(int index, const Tensor& tensor) {
  if (output_owned_[index] && outputs_[index].tensor) delete outputs_[index].tensor;
  outputs_[index] = TensorValue(new Tensor(tensor));
  output_owned_[index] = true;
}

void OpKernelContext::set_output_ref(int index, std::mutex* mu, Tensor* tensor_for_ref) {
  if (output_owned_[index] && outputs_[index].tensor) delete outputs_[index].tensor;
  outputs_[index] = TensorValue(mu, tensor_for_ref);
  output_owned_[index] = false;
}

TensorValue OpKernelContext::release_output(int index) {
  TensorValue v = outputs_[index];
  outputs_[index] = TensorValue();
  output_owned_[index] = false;
  return v;
}

// ------------------------------------------------------------------ registry
namespace {
struct KernelRegistration {
  KernelDef def;
  std::string kernel_class_name;
  OpKernelFactory factory;
};
typedef std::multimap<std::string, KernelRegistration> KernelRegistry;
KernelRegistry* GlobalKernelRegistry() {
  static KernelRegistry* r = new KernelRegistry;
  return r;
}
std::mutex* RegistryMutex() {
  static std::mutex* m = new std::mutex;
  return m;
}
std::string Key(const std::string& op, const std::string& device, const std::string& label) {
  return op + ":" + device + ":" + label;
}

Status AttrsMatch(const NodeDef& node, const KernelDef& kd, bool* match) {
  *match = false;
  for (const auto& c : kd.constraint) {
    auto it = node.attr.find(c.name);
    if (it == node.attr.end())
      return errors::InvalidArgument("OpKernel '", kd.op, "' has constraint on attr '", c.name,
                                     "' not in NodeDef '", SummarizeNodeDef(node), "'");
    if (it->second.kind != AttrValue::kType)
      return errors::Unimplemented("KernelDef constraint on non-type attr '", c.name, "'");
    if (std::find(c.allowed_values.begin(), c.allowed_values.end(), it->second.type) ==
        c.allowed_values.end())
      return Status::OK();
  }
  *match = true;
  return Status::OK();
}

Status FindKernelRegistration(const DeviceType& device_type, const NodeDef& node,
                              const KernelRegistration** reg) {
  *reg = nullptr;
  std::string label;
  auto it = node.attr.find("_kernel");
  if (it != node.attr.end() && it->second.kind == AttrValue::kS) label = it->second.s;
  const std::string key = Key(node.op, device_type.type(), label);
  std::lock_guard<std::mutex> l(*RegistryMutex());
  auto range = GlobalKernelRegistry()->equal_range(key);
  for (auto i = range.first; i != range.second; ++i) {
    bool match;
    TF_RETURN_IF_ERROR(AttrsMatch(node, i->second.def, &match));
    if (match) {
      if (*reg != nullptr)
        return errors::InvalidArgument("Multiple OpKernel registrations match NodeDef '",
                                       SummarizeNodeDef(node), "': '", (*reg)->kernel_class_name,
                                       "' and '", i->second.kernel_class_name, "'");
      *reg = &i->second;
    }
  }
  return Status::OK();
}

// memory_types.cc:105-133: args named by HostMemory() live in host memory.
void MemoryTypesForNode(const NodeDef& node, const OpDef& op_def, const KernelDef& kd,
                        MemoryTypeVector* in, MemoryTypeVector* out) {
  auto expand = [&](const std::vector<OpDef::ArgDef>& args, MemoryTypeVector* v) {
    for (const auto& a : args) {
      int64 count = 1;
      if (!a.number_attr.empty()) GetNodeAttr(node, a.number_attr, &count);
      const bool host = std::find(kd.host_memory_arg.begin(), kd.host_memory_arg.end(), a.name) !=
                        kd.host_memory_arg.end();
      for (int64 i = 0; i < count; ++i) v->push_back(host ? HOST_MEMORY : DEVICE_MEMORY);
    }
  };
  in->clear();
  out->clear();
  expand(op_def.input_arg, in);
  expand(op_def.output_arg, out);
}
}  // namespace

namespace kernel_factory {
OpKernelRegistrar::OpKernelRegistrar(const KernelDef* kernel_def,
                                     const std::string& kernel_class_name,
                                     OpKernelFactory factory) {
  std::lock_guard<std::mutex> l(*RegistryMutex());
  GlobalKernelRegistry()->insert(
      {Key(kernel_def->op, kernel_def->device_type, kernel_def->label),
       KernelRegistration{*kernel_def, kernel_class_name, factory}});
  delete kernel_def;
}
}  // namespace kernel_factory

std::vector<std::string> RegisteredKernelKeys() {
  std::lock_guard<std::mutex> l(*RegistryMutex());
  std::vector<std::string> out;
  for (const auto& kv : *GlobalKernelRegistry()) out.push_back(kv.first);
  return out;
}

Status SupportedDeviceTypesForNode(const std::vector<DeviceType>& prioritized_types,
                                   const NodeDef& def, std::vector<DeviceType>* device_types) {
  device_types->clear();
  for (const DeviceType& dt : prioritized_types) {
    const KernelRegistration* reg = nullptr;
    TF_RETURN_IF_ERROR(FindKernelRegistration(dt, def, &reg));
    if (reg) device_types->push_back(dt);
  }
  return Status::OK();
}

Status CreateOpKernel(DeviceType device_type, DeviceBase* device, Allocator* allocator,
                      const NodeDef& node_def_in, std::unique_ptr<OpKernel>* kernel) {
  const OpDef* op_def = OpRegistry::Global()->LookUp(node_def_in.op);
  if (op_def == nullptr)
    return errors::NotFound("Op type not registered '", node_def_in.op, "'");
  NodeDef node_def = node_def_in;
  TF_RETURN_IF_ERROR(ValidateNodeDef(&node_def, *op_def));
  const KernelRegistration* reg = nullptr;
  TF_RETURN_IF_ERROR(FindKernelRegistration(device_type, node_def, &reg));
  if (reg == nullptr)
    return errors::NotFound("No registered '", node_def.op, "' OpKernel for ", device_type.type(),
                            " devices compatible with node ", SummarizeNodeDef(node_def));
  DataTypeVector inputs, outputs;
  TF_RETURN_IF_ERROR(InOutTypesForNode(node_def, *op_def, &inputs, &outputs));
  MemoryTypeVector in_mem, out_mem;
  MemoryTypesForNode(node_def, *op_def, reg->def, &in_mem, &out_mem);
  Status s;
  OpKernelConstruction context(device_type, device, allocator, &node_def, op_def, inputs, in_mem,
                               outputs, out_mem, &s);
  OpKernel* k = reg->factory(&context);
  if (!s.ok()) {  // the framework deletes a kernel whose constructor failed (:1058-1063)
    delete k;
    return s;
  }
  kernel->reset(k);
  return Status::OK();
}

}  // namespace tensorflow
