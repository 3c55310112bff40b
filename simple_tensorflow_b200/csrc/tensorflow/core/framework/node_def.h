// NodeDef / AttrValue / GraphDef as plain C++ structs.
// The reference defines them as protobufs (core/framework/{node_def,attr_value,graph}.proto);
// protoc is not available here, so the same fields are carried by value types:
//   NodeDef{name=1, op=2, input=3, device=4, attr=5}
//   AttrValue{list=1, s=2, i=3, f=4, b=5, type=6, shape=7, tensor=8}
#ifndef B200TF_CORE_FRAMEWORK_NODE_DEF_H_
#define B200TF_CORE_FRAMEWORK_NODE_DEF_H_

#include <map>
#include <string>
#include <vector>

#include "tensorflow/core/framework/tensor.h"

namespace tensorflow {

struct AttrValue {
  enum Kind { kNone, kS, kI, kF, kB, kType, kShape, kTensor, kListI, kListS, kListType, kRaw };
  Kind kind = kNone;
  std::string s;
  int64 i = 0;
  float f = 0.f;
  bool b = false;
  DataType type = DT_INVALID;
  TensorShape shape;
  Tensor tensor;  // host tensor (Const)
  std::vector<int64> list_i;
  std::vector<std::string> list_s;
  std::vector<DataType> list_type;
  // kRaw: a serialized AttrValue this runtime does not model (list(shape), list(float), func,
  // string tensors, ...), carried verbatim so that graphs round-trip losslessly.
  std::string raw;

  static AttrValue S(const std::string& v) { AttrValue a; a.kind = kS; a.s = v; return a; }
  static AttrValue I(int64 v) { AttrValue a; a.kind = kI; a.i = v; return a; }
  static AttrValue F(float v) { AttrValue a; a.kind = kF; a.f = v; return a; }
  static AttrValue B(bool v) { AttrValue a; a.kind = kB; a.b = v; return a; }
  static AttrValue Type(DataType v) { AttrValue a; a.kind = kType; a.type = v; return a; }
  static AttrValue Shape(const TensorShape& v) { AttrValue a; a.kind = kShape; a.shape = v; return a; }
  static AttrValue TensorV(const Tensor& v) { AttrValue a; a.kind = kTensor; a.tensor = v; return a; }
  static AttrValue ListI(const std::vector<int64>& v) { AttrValue a; a.kind = kListI; a.list_i = v; return a; }
  static AttrValue ListS(const std::vector<std::string>& v) { AttrValue a; a.kind = kListS; a.list_s = v; return a; }
};

struct NodeDef {
  std::string name;
  std::string op;
  std::vector<std::string> input;  // "node", "node:k", or "^node" (control dependency)
  std::string device;
  std::map<std::string, AttrValue> attr;
};

struct GraphDef {
  std::vector<NodeDef> node;
  // graph.proto fields 4 (VersionDef) and 2 (FunctionDefLibrary), serialized, carried verbatim.
  std::string versions_raw, library_raw;
};

// node_def_util.h GetNodeAttr overloads.
Status GetNodeAttr(const NodeDef& n, const std::string& name, std::string* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, int64* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, int32* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, float* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, bool* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, DataType* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, TensorShape* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, Tensor* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, std::vector<int32>* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, std::vector<int64>* v);
Status GetNodeAttr(const NodeDef& n, const std::string& name, std::vector<std::string>* v);
std::string SummarizeNodeDef(const NodeDef& n);

}  // namespace tensorflow
#endif
