#include "tensorflow/c/c_api.h"

#include "tensorflow/core/framework/graph_def_wire.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "b200_ops.h"
#include "tensorflow/core/common_runtime/direct_session.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/public/session.h"

using tensorflow::AttrValue;
using tensorflow::DataType;
using tensorflow::NodeDef;
using tensorflow::Status;
using tensorflow::Tensor;
using tensorflow::TensorShape;

struct TF_Status {
  Status status;
};
struct TF_Tensor {
  Tensor tensor;
};
struct TF_SessionOptions {
  tensorflow::SessionOptions options;
};
struct TF_Operation {
  NodeDef node;
  tensorflow::DataTypeVector input_types, output_types;
  bool opaque = false;  // imported node of an op type that is not registered here
  TF_Graph* graph = nullptr;
};
struct TF_ImportGraphDefOptions {
  std::string prefix;
};
struct TF_Graph {
  std::mutex mu;
  std::vector<std::unique_ptr<TF_Operation>> operations;
  std::map<std::string, TF_Operation*> by_name;
  int num_sessions = 0;
};
struct TF_OperationDescription {
  TF_Graph* graph;
  NodeDef node;
  Status deferred;  // first error seen while describing the op; reported by TF_FinishOperation
};
struct TF_Session {
  TF_Graph* graph;
  tensorflow::Session* session;
  size_t num_nodes_sent = 0;
};
struct TF_Library {
  void* handle;
};

namespace {
// Pinned allocator shared by TF_AllocateTensor when a GPU is present.
tensorflow::Allocator* HostTensorAllocator() {
  static tensorflow::Allocator* a = [] {
    if (b200_device_count() > 0)
      return static_cast<tensorflow::Allocator*>(tensorflow::GPUHostAllocator::Process());
    return tensorflow::cpu_allocator();
  }();
  return a;
}
// A TensorBuffer view over caller-owned memory with a deallocator (TF_NewTensor).
class ClientAllocator : public tensorflow::Allocator {
 public:
  ClientAllocator(void* data, size_t len, void (*d)(void*, size_t, void*), void* arg)
      : data_(data), len_(len), dealloc_(d), arg_(arg) {}
  std::string Name() override { return "client"; }
  void* AllocateRaw(size_t, size_t) override { return data_; }
  void DeallocateRaw(void*) override {
    if (dealloc_) dealloc_(data_, len_, arg_);
    delete this;
  }
 private:
  void* data_;
  size_t len_;
  void (*dealloc_)(void*, size_t, void*);
  void* arg_;
};
}  // namespace

extern "C" {

const char* TF_Version(void) { return "1.0.0-b200"; }
size_t TF_DataTypeSize(TF_DataType dt) {
  return tensorflow::DataTypeSize(static_cast<DataType>(dt));
}

TF_Status* TF_NewStatus(void) { return new TF_Status; }
void TF_DeleteStatus(TF_Status* s) { delete s; }
void TF_SetStatus(TF_Status* s, TF_Code code, const char* msg) {
  s->status = code == TF_OK ? Status::OK()
                            : Status(static_cast<tensorflow::error::Code>(code), msg);
}
TF_Code TF_GetCode(const TF_Status* s) { return static_cast<TF_Code>(s->status.code()); }
const char* TF_Message(const TF_Status* s) { return s->status.error_message().c_str(); }

TF_Tensor* TF_NewTensor(TF_DataType dtype, const int64_t* dims, int num_dims, void* data,
                        size_t len, void (*deallocator)(void*, size_t, void*),
                        void* deallocator_arg) {
  TensorShape shape;
  for (int i = 0; i < num_dims; ++i) shape.AddDim(dims[i]);
  if (!shape.IsValid(tensorflow::DataTypeSize(static_cast<DataType>(dtype)))) return nullptr;
  const size_t need =
      static_cast<size_t>(shape.num_elements()) * tensorflow::DataTypeSize(static_cast<DataType>(dtype));
  if (need > len || need == 0) {
    if (need == 0) {
      if (deallocator) deallocator(data, len, deallocator_arg);
      return new TF_Tensor{Tensor(static_cast<DataType>(dtype), shape)};
    }
    return nullptr;
  }
  auto* a = new ClientAllocator(data, len, deallocator, deallocator_arg);
  return new TF_Tensor{Tensor(a, static_cast<DataType>(dtype), shape)};
}
TF_Tensor* TF_AllocateTensor(TF_DataType dtype, const int64_t* dims, int num_dims, size_t len) {
  TensorShape shape;
  for (int i = 0; i < num_dims; ++i) shape.AddDim(dims[i]);
  (void)len;
  if (!shape.IsValid(tensorflow::DataTypeSize(static_cast<DataType>(dtype)))) return nullptr;
  Tensor t(HostTensorAllocator(), static_cast<DataType>(dtype), shape);
  if (!t.IsInitialized()) return nullptr;
  return new TF_Tensor{std::move(t)};
}
void TF_DeleteTensor(TF_Tensor* t) { delete t; }
TF_DataType TF_TensorType(const TF_Tensor* t) { return static_cast<TF_DataType>(t->tensor.dtype()); }
int TF_NumDims(const TF_Tensor* t) { return t->tensor.dims(); }
int64_t TF_Dim(const TF_Tensor* t, int i) { return t->tensor.dim_size(i); }
size_t TF_TensorByteSize(const TF_Tensor* t) { return t->tensor.TotalBytes(); }
void* TF_TensorData(const TF_Tensor* t) { return t->tensor.raw_data(); }

TF_SessionOptions* TF_NewSessionOptions(void) { return new TF_SessionOptions; }
void TF_SetTarget(TF_SessionOptions* o, const char* target) { o->options.target = target; }
void TF_DeleteSessionOptions(TF_SessionOptions* o) { delete o; }
void B200TF_SetGpuDevice(TF_SessionOptions* o, int gpu_id) { o->options.gpu_device_id = gpu_id; }
void B200TF_SetGpuMemoryLimit(TF_SessionOptions* o, size_t bytes) {
  o->options.gpu_memory_limit_bytes = bytes;
}
void B200TF_SetCollective(TF_SessionOptions* o, void* comm, int num_replicas) {
  o->options.collective_comm = comm;
  o->options.num_replicas = num_replicas;
}

TF_Graph* TF_NewGraph(void) { return new TF_Graph; }
void TF_DeleteGraph(TF_Graph* g) { delete g; }

TF_OperationDescription* TF_NewOperation(TF_Graph* graph, const char* op_type,
                                         const char* oper_name) {
  auto* d = new TF_OperationDescription;
  d->graph = graph;
  d->node.op = op_type;
  d->node.name = oper_name;
  return d;
}
void TF_SetDevice(TF_OperationDescription* d, const char* device) { d->node.device = device; }
void TF_AddInput(TF_OperationDescription* d, TF_Output input) {
  if (input.oper == nullptr) {
    d->deferred.Update(tensorflow::errors::InvalidArgument("TF_AddInput: null operation"));
    return;
  }
  d->node.input.push_back(input.index == 0 ? input.oper->node.name
                                           : input.oper->node.name + ":" + std::to_string(input.index));
}
void TF_AddInputList(TF_OperationDescription* d, const TF_Output* inputs, int num_inputs) {
  for (int i = 0; i < num_inputs; ++i) TF_AddInput(d, inputs[i]);
}
void TF_AddControlInput(TF_OperationDescription* d, TF_Operation* input) {
  d->node.input.push_back("^" + input->node.name);
}
void TF_SetAttrString(TF_OperationDescription* d, const char* name, const void* value,
                      size_t length) {
  d->node.attr[name] = AttrValue::S(std::string(static_cast<const char*>(value), length));
}
void TF_SetAttrInt(TF_OperationDescription* d, const char* name, int64_t value) {
  d->node.attr[name] = AttrValue::I(value);
}
void TF_SetAttrIntList(TF_OperationDescription* d, const char* name, const int64_t* values,
                       int num_values) {
  d->node.attr[name] = AttrValue::ListI(std::vector<tensorflow::int64>(values, values + num_values));
}
void TF_SetAttrFloat(TF_OperationDescription* d, const char* name, float value) {
  d->node.attr[name] = AttrValue::F(value);
}
void TF_SetAttrBool(TF_OperationDescription* d, const char* name, unsigned char value) {
  d->node.attr[name] = AttrValue::B(value != 0);
}
void TF_SetAttrType(TF_OperationDescription* d, const char* name, TF_DataType value) {
  d->node.attr[name] = AttrValue::Type(static_cast<DataType>(value));
}
void TF_SetAttrShape(TF_OperationDescription* d, const char* name, const int64_t* dims,
                     int num_dims) {
  TensorShape s;
  for (int i = 0; i < num_dims; ++i) s.AddDim(dims[i]);
  d->node.attr[name] = AttrValue::Shape(s);
}
void TF_SetAttrTensor(TF_OperationDescription* d, const char* name, TF_Tensor* value,
                      TF_Status* status) {
  // Deep copy: the attr must outlive the caller's tensor.
  Tensor copy(value->tensor.dtype(), value->tensor.shape());
  if (value->tensor.TotalBytes() > 0)
    memcpy(copy.raw_data(), value->tensor.raw_data(), value->tensor.TotalBytes());
  d->node.attr[name] = AttrValue::TensorV(copy);
  status->status = Status::OK();
}

TF_Operation* TF_FinishOperation(TF_OperationDescription* d, TF_Status* status) {
  std::unique_ptr<TF_OperationDescription> owner(d);
  status->status = d->deferred;
  if (!status->status.ok()) return nullptr;
  std::lock_guard<std::mutex> l(d->graph->mu);
  if (d->graph->by_name.count(d->node.name)) {
    status->status = tensorflow::errors::InvalidArgument("Duplicate node name in graph: '",
                                                         d->node.name, "'");
    return nullptr;
  }
  const tensorflow::OpDef* op_def = tensorflow::OpRegistry::Global()->LookUp(d->node.op);
  if (op_def == nullptr) {
    status->status = tensorflow::errors::NotFound("Op type not registered '", d->node.op, "'");
    return nullptr;
  }
  std::unique_ptr<TF_Operation> op(new TF_Operation);
  op->node = d->node;
  status->status = tensorflow::ValidateNodeDef(&op->node, *op_def);
  if (!status->status.ok()) return nullptr;
  status->status = tensorflow::InOutTypesForNode(op->node, *op_def, &op->input_types,
                                                 &op->output_types);
  if (!status->status.ok()) return nullptr;
  size_t data_inputs = 0;
  for (const auto& in : op->node.input) data_inputs += (in.empty() || in[0] != '^');
  if (data_inputs != op->input_types.size()) {
    status->status = tensorflow::errors::InvalidArgument(
        "Node '", op->node.name, "' of type ", op->node.op, " expects ", op->input_types.size(),
        " inputs but ", data_inputs, " were added");
    return nullptr;
  }
  TF_Operation* raw = op.get();
  raw->graph = d->graph;
  d->graph->by_name[raw->node.name] = raw;
  d->graph->operations.push_back(std::move(op));
  return raw;
}

const char* TF_OperationName(TF_Operation* oper) { return oper->node.name.c_str(); }
const char* TF_OperationOpType(TF_Operation* oper) { return oper->node.op.c_str(); }
int TF_OperationNumOutputs(TF_Operation* oper) { return static_cast<int>(oper->output_types.size()); }
TF_DataType TF_OperationOutputType(TF_Output o) {
  if (o.index < 0 || o.index >= static_cast<int>(o.oper->output_types.size()))
    return static_cast<TF_DataType>(0);  // opaque imported node: types unknown
  return static_cast<TF_DataType>(o.oper->output_types[o.index]);
}
int TF_OperationNumInputs(TF_Operation* oper) {
  if (!oper->opaque) return static_cast<int>(oper->input_types.size());
  int n = 0;  // opaque imported node: the types are unknown, the edges are not
  for (const std::string& name : oper->node.input) n += name.empty() || name[0] != '^';
  return n;
}
TF_Operation* TF_GraphOperationByName(TF_Graph* graph, const char* oper_name) {
  std::lock_guard<std::mutex> l(graph->mu);
  auto it = graph->by_name.find(oper_name);
  return it == graph->by_name.end() ? nullptr : it->second;
}

// ---------------------------------------------------------------- operation introspection
static const AttrValue* FindAttr(TF_Operation* oper, const char* name, TF_Status* status) {
  auto it = oper->node.attr.find(name);
  if (it == oper->node.attr.end()) {
    status->status = tensorflow::errors::InvalidArgument("Operation '", oper->node.name,
                                                         "' has no attr named '", name, "'.");
    return nullptr;
  }
  status->status = Status::OK();
  return &it->second;
}
static bool RequireKind(const AttrValue* a, AttrValue::Kind kind, const char* name,
                        TF_Status* status) {
  if (a == nullptr) return false;
  if (a->kind != kind) {
    status->status = tensorflow::errors::InvalidArgument("Attribute '", name,
                                                         "' does not have the requested type");
    return false;
  }
  return true;
}

TF_Output TF_OperationInput(TF_Input in) {
  TF_Output none{nullptr, 0};
  if (in.oper == nullptr || in.oper->graph == nullptr || in.index < 0) return none;
  int seen = 0;
  for (const std::string& name : in.oper->node.input) {
    if (!name.empty() && name[0] == '^') continue;
    if (seen++ != in.index) continue;
    std::string src = name;
    int slot = 0;
    const size_t colon = src.rfind(':');
    if (colon != std::string::npos) {
      slot = atoi(src.c_str() + colon + 1);
      src = src.substr(0, colon);
    }
    std::lock_guard<std::mutex> l(in.oper->graph->mu);
    auto it = in.oper->graph->by_name.find(src);
    if (it == in.oper->graph->by_name.end()) return none;
    return TF_Output{it->second, slot};
  }
  return none;
}
int TF_OperationNumControlInputs(TF_Operation* oper) {
  int n = 0;
  for (const std::string& name : oper->node.input) n += !name.empty() && name[0] == '^';
  return n;
}
int TF_OperationGetControlInputs(TF_Operation* oper, TF_Operation** control_inputs, int max) {
  int n = 0;
  if (oper->graph == nullptr) return 0;
  std::lock_guard<std::mutex> l(oper->graph->mu);
  for (const std::string& name : oper->node.input) {
    if (name.empty() || name[0] != '^' || n >= max) continue;
    auto it = oper->graph->by_name.find(name.substr(1));
    if (it != oper->graph->by_name.end()) control_inputs[n++] = it->second;
  }
  return n;
}

TF_AttrMetadata TF_OperationGetAttrMetadata(TF_Operation* oper, const char* name,
                                            TF_Status* status) {
  TF_AttrMetadata m{0, -1, TF_ATTR_PLACEHOLDER, -1};
  const AttrValue* a = FindAttr(oper, name, status);
  if (a == nullptr) return m;
  switch (a->kind) {
    case AttrValue::kS: m.type = TF_ATTR_STRING; m.total_size = (int64_t)a->s.size(); break;
    case AttrValue::kI: m.type = TF_ATTR_INT; break;
    case AttrValue::kF: m.type = TF_ATTR_FLOAT; break;
    case AttrValue::kB: m.type = TF_ATTR_BOOL; break;
    case AttrValue::kType: m.type = TF_ATTR_TYPE; break;
    case AttrValue::kShape: m.type = TF_ATTR_SHAPE; m.total_size = a->shape.dims(); break;
    case AttrValue::kTensor: m.type = TF_ATTR_TENSOR; break;
    case AttrValue::kListI:
      m.is_list = 1; m.type = TF_ATTR_INT; m.list_size = (int64_t)a->list_i.size(); break;
    case AttrValue::kListS:
      m.is_list = 1; m.type = TF_ATTR_STRING; m.list_size = (int64_t)a->list_s.size();
      m.total_size = 0;
      for (const auto& s : a->list_s) m.total_size += (int64_t)s.size();
      break;
    case AttrValue::kListType:
      m.is_list = 1; m.type = TF_ATTR_TYPE; m.list_size = (int64_t)a->list_type.size(); break;
    case AttrValue::kRaw: m.type = TF_ATTR_PLACEHOLDER; m.total_size = (int64_t)a->raw.size(); break;
    case AttrValue::kNone: break;
  }
  return m;
}
void TF_OperationGetAttrString(TF_Operation* oper, const char* name, void* value, size_t max_length,
                               TF_Status* status) {
  const AttrValue* a = FindAttr(oper, name, status);
  if (!RequireKind(a, AttrValue::kS, name, status)) return;
  memcpy(value, a->s.data(), std::min(max_length, a->s.size()));
}
void TF_OperationGetAttrInt(TF_Operation* oper, const char* name, int64_t* value,
                            TF_Status* status) {
  const AttrValue* a = FindAttr(oper, name, status);
  if (RequireKind(a, AttrValue::kI, name, status)) *value = a->i;
}
void TF_OperationGetAttrIntList(TF_Operation* oper, const char* name, int64_t* values,
                                int max_values, TF_Status* status) {
  const AttrValue* a = FindAttr(oper, name, status);
  if (!RequireKind(a, AttrValue::kListI, name, status)) return;
  for (int i = 0; i < max_values && i < (int)a->list_i.size(); ++i) values[i] = a->list_i[i];
}
void TF_OperationGetAttrFloat(TF_Operation* oper, const char* name, float* value,
                              TF_Status* status) {
  const AttrValue* a = FindAttr(oper, name, status);
  if (RequireKind(a, AttrValue::kF, name, status)) *value = a->f;
}
void TF_OperationGetAttrBool(TF_Operation* oper, const char* name, unsigned char* value,
                             TF_Status* status) {
  const AttrValue* a = FindAttr(oper, name, status);
  if (RequireKind(a, AttrValue::kB, name, status)) *value = a->b ? 1 : 0;
}
void TF_OperationGetAttrType(TF_Operation* oper, const char* name, TF_DataType* value,
                             TF_Status* status) {
  const AttrValue* a = FindAttr(oper, name, status);
  if (RequireKind(a, AttrValue::kType, name, status)) *value = static_cast<TF_DataType>(a->type);
}
void TF_OperationGetAttrShape(TF_Operation* oper, const char* name, int64_t* value, int num_dims,
                              TF_Status* status) {
  const AttrValue* a = FindAttr(oper, name, status);
  if (!RequireKind(a, AttrValue::kShape, name, status)) return;
  for (int i = 0; i < num_dims && i < a->shape.dims(); ++i) value[i] = a->shape.dim_size(i);
}
void TF_OperationGetAttrTensor(TF_Operation* oper, const char* name, TF_Tensor** value,
                               TF_Status* status) {
  *value = nullptr;
  const AttrValue* a = FindAttr(oper, name, status);
  if (!RequireKind(a, AttrValue::kTensor, name, status)) return;
  Tensor copy(HostTensorAllocator(), a->tensor.dtype(), a->tensor.shape());
  if (a->tensor.TotalBytes() > 0) {
    if (!copy.IsInitialized()) {
      status->status = tensorflow::errors::ResourceExhausted("OOM copying attr '", name, "'");
      return;
    }
    memcpy(copy.raw_data(), a->tensor.raw_data(), a->tensor.TotalBytes());
  }
  *value = new TF_Tensor{copy};
}
char* B200TF_OperationAttrNames(TF_Operation* oper) {
  std::string s;
  for (const auto& kv : oper->node.attr) s += kv.first + "\n";
  char* out = static_cast<char*>(malloc(s.size() + 1));
  memcpy(out, s.c_str(), s.size() + 1);
  return out;
}

TF_Operation* TF_GraphNextOperation(TF_Graph* graph, size_t* pos) {
  std::lock_guard<std::mutex> l(graph->mu);
  if (*pos >= graph->operations.size()) return nullptr;
  return graph->operations[(*pos)++].get();
}

// ---------------------------------------------------------------- GraphDef import / export
TF_Buffer* TF_NewBuffer(void) { return new TF_Buffer{nullptr, 0, nullptr}; }
TF_Buffer* TF_NewBufferFromString(const void* proto, size_t proto_len) {
  void* copy = malloc(proto_len ? proto_len : 1);
  if (proto_len) memcpy(copy, proto, proto_len);
  return new TF_Buffer{copy, proto_len, [](void* data, size_t) { free(data); }};
}
void TF_DeleteBuffer(TF_Buffer* b) {
  if (b == nullptr) return;
  if (b->data_deallocator) b->data_deallocator(const_cast<void*>(b->data), b->length);
  delete b;
}
TF_Buffer TF_GetBuffer(TF_Buffer* b) { return *b; }

void TF_GraphToGraphDef(TF_Graph* graph, TF_Buffer* out, TF_Status* status) {
  tensorflow::GraphDef def;
  {
    std::lock_guard<std::mutex> l(graph->mu);
    for (const auto& op : graph->operations) def.node.push_back(op->node);
  }
  def.versions_raw = std::string("\x08\x15", 2);  // VersionDef{ producer = 21 }
  std::string bytes;
  tensorflow::SerializeGraphDef(def, &bytes);
  if (out->data_deallocator) out->data_deallocator(const_cast<void*>(out->data), out->length);
  void* copy = malloc(bytes.size() ? bytes.size() : 1);
  memcpy(copy, bytes.data(), bytes.size());
  out->data = copy;
  out->length = bytes.size();
  out->data_deallocator = [](void* data, size_t) { free(data); };
  status->status = Status::OK();
}

TF_ImportGraphDefOptions* TF_NewImportGraphDefOptions(void) { return new TF_ImportGraphDefOptions; }
void TF_DeleteImportGraphDefOptions(TF_ImportGraphDefOptions* o) { delete o; }
void TF_ImportGraphDefOptionsSetPrefix(TF_ImportGraphDefOptions* o, const char* prefix) {
  o->prefix = prefix ? prefix : "";
}

void TF_GraphImportGraphDef(TF_Graph* graph, const TF_Buffer* graph_def,
                            const TF_ImportGraphDefOptions* options, TF_Status* status) {
  tensorflow::GraphDef def;
  status->status = tensorflow::ParseGraphDef(graph_def->data, graph_def->length, &def);
  if (!status->status.ok()) return;
  std::string prefix = options ? options->prefix : "";
  if (!prefix.empty() && prefix.back() != '/') prefix += "/";
  std::vector<std::unique_ptr<TF_Operation>> added;
  std::lock_guard<std::mutex> l(graph->mu);
  std::map<std::string, TF_Operation*> names;
  for (NodeDef& nd : def.node) {
    std::unique_ptr<TF_Operation> op(new TF_Operation);
    op->node = nd;
    op->node.name = prefix + nd.name;
    for (std::string& in : op->node.input)  // "node", "node:k", "^node"
      in = (!in.empty() && in[0] == '^') ? "^" + prefix + in.substr(1) : prefix + in;
    if (graph->by_name.count(op->node.name) || names.count(op->node.name)) {
      status->status = tensorflow::errors::InvalidArgument("Duplicate node name in graph: '",
                                                           op->node.name, "'");
      return;
    }
    const tensorflow::OpDef* op_def = tensorflow::OpRegistry::Global()->LookUp(op->node.op);
    if (op_def == nullptr) {
      op->opaque = true;
    } else {
      status->status = tensorflow::ValidateNodeDef(&op->node, *op_def);
      if (status->status.ok())
        status->status = tensorflow::InOutTypesForNode(op->node, *op_def, &op->input_types,
                                                       &op->output_types);
      if (!status->status.ok()) return;
      size_t data_inputs = 0;
      for (const auto& in : op->node.input) data_inputs += (in.empty() || in[0] != '^');
      if (data_inputs != op->input_types.size()) {
        status->status = tensorflow::errors::InvalidArgument(
            "Node '", op->node.name, "' of type ", op->node.op, " expects ",
            op->input_types.size(), " inputs but has ", data_inputs);
        return;
      }
    }
    names[op->node.name] = op.get();
    added.push_back(std::move(op));
  }
  // every input must name a node of the graph (graph_constructor.cc's edge check)
  for (const auto& op : added)
    for (const std::string& in : op->node.input) {
      std::string src = (!in.empty() && in[0] == '^') ? in.substr(1) : in;
      const size_t colon = src.rfind(':');
      if (colon != std::string::npos) src = src.substr(0, colon);
      if (!names.count(src) && !graph->by_name.count(src)) {
        status->status = tensorflow::errors::InvalidArgument(
            "Node '", op->node.name, "': Unknown input node '", in, "'");
        return;
      }
    }
  for (auto& op : added) {
    op->graph = graph;
    graph->by_name[op->node.name] = op.get();
    graph->operations.push_back(std::move(op));
  }
  status->status = Status::OK();
}

char* B200TF_GraphDefToText(const void* proto, size_t proto_len, TF_Status* status) {
  tensorflow::GraphDef def;
  status->status = tensorflow::ParseGraphDef(proto, proto_len, &def);
  if (!status->status.ok()) return nullptr;
  const std::string text = tensorflow::GraphDefDebugString(def);
  char* out = static_cast<char*>(malloc(text.size() + 1));
  memcpy(out, text.c_str(), text.size() + 1);
  return out;
}

TF_Session* TF_NewSession(TF_Graph* graph, const TF_SessionOptions* opts, TF_Status* status) {
  tensorflow::Session* session = nullptr;
  status->status = tensorflow::NewSession(
      opts ? opts->options : tensorflow::SessionOptions(), &session);
  if (!status->status.ok()) return nullptr;
  auto* s = new TF_Session;
  s->graph = graph;
  s->session = session;
  return s;
}
void TF_CloseSession(TF_Session* s, TF_Status* status) { status->status = s->session->Close(); }
void TF_DeleteSession(TF_Session* s, TF_Status* status) {
  status->status = Status::OK();
  delete s->session;
  delete s;
}

// c_api.cc ExtendSessionGraphHelper: ship the nodes added to the TF_Graph since the last Run.
static Status ExtendSession(TF_Session* s) {
  tensorflow::GraphDef delta;
  {
    std::lock_guard<std::mutex> l(s->graph->mu);
    for (size_t i = s->num_nodes_sent; i < s->graph->operations.size(); ++i)
      delta.node.push_back(s->graph->operations[i]->node);
    if (delta.node.empty()) return Status::OK();
    const bool first = s->num_nodes_sent == 0;
    // last_num_graph_nodes advances only on success (c_api.cc ExtendSessionGraphHelper): a failed
    // extend leaves the session unchanged and the same delta is offered again next time.
    Status st = first ? s->session->Create(delta) : s->session->Extend(delta);
    if (st.ok()) s->num_nodes_sent += delta.node.size();
    return st;
  }
}

static std::string OutputName(const TF_Output& o) {
  return o.oper->node.name + ":" + std::to_string(o.index);
}

void TF_SessionRun(TF_Session* session, const void* run_options, const TF_Output* inputs,
                   TF_Tensor* const* input_values, int ninputs, const TF_Output* outputs,
                   TF_Tensor** output_values, int noutputs,
                   const TF_Operation* const* target_opers, int ntargets, void* run_metadata,
                   TF_Status* status) {
  for (int i = 0; i < noutputs; ++i) output_values[i] = nullptr;
  if (run_options != nullptr || run_metadata != nullptr) {
    status->status = tensorflow::errors::Unimplemented(
        "RunOptions / RunMetadata are protobuf buffers and are not supported; pass NULL");
    return;
  }
  status->status = ExtendSession(session);
  if (!status->status.ok()) return;
  std::vector<std::pair<std::string, Tensor>> feed;
  for (int i = 0; i < ninputs; ++i)
    feed.emplace_back(OutputName(inputs[i]), input_values[i]->tensor);
  std::vector<std::string> fetch, targets;
  for (int i = 0; i < noutputs; ++i) fetch.push_back(OutputName(outputs[i]));
  for (int i = 0; i < ntargets; ++i) targets.push_back(target_opers[i]->node.name);
  std::vector<Tensor> out;
  status->status = session->session->Run(feed, fetch, targets, &out);
  if (!status->status.ok()) return;
  for (int i = 0; i < noutputs; ++i) output_values[i] = new TF_Tensor{out[i]};
}

void B200TF_SessionLastRunStats(TF_Session* s, B200TF_RunStats* out) {
  const tensorflow::RunStats& r = s->session->last_run_stats();
  out->nodes_executed = r.nodes_executed;
  out->kernels_launched = r.kernels_launched;
  out->h2d_bytes = r.h2d_bytes;
  out->d2h_bytes = r.d2h_bytes;
  out->host_enqueue_us = r.host_enqueue_us;
  out->host_total_us = r.host_total_us;
}

TF_Tensor* B200TF_SessionStageTensor(TF_Session* s, TF_Tensor* host, TF_Status* status) {
  tensorflow::Tensor staged;
  status->status = s->session->StageFeed(host->tensor, &staged);
  if (!status->status.ok()) return nullptr;
  return new TF_Tensor{staged};
}

void* B200TF_SessionStream(TF_Session* s) {
  auto* ds = dynamic_cast<tensorflow::DirectSession*>(s->session);
  return ds ? ds->device()->compute_stream()->cuda_stream() : nullptr;
}

static char* JoinLines(const std::vector<std::string>& v) {
  std::string s;
  for (const auto& e : v) s += e + "\n";
  char* out = static_cast<char*>(malloc(s.size() + 1));
  memcpy(out, s.c_str(), s.size() + 1);
  return out;
}
char* B200TF_ListRegisteredOps(void) {
  return JoinLines(tensorflow::OpRegistry::Global()->ListOps());
}
char* B200TF_ListRegisteredKernels(void) { return JoinLines(tensorflow::RegisteredKernelKeys()); }

TF_Library* TF_LoadLibrary(const char* library_filename, TF_Status* status) {
  void* h = dlopen(library_filename, RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    status->status = tensorflow::errors::NotFound(dlerror());
    return nullptr;
  }
  status->status = Status::OK();
  return new TF_Library{h};
}
void TF_DeleteLibraryHandle(TF_Library* lib) { delete lib; }

}  // extern "C"
