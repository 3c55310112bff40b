// OpDef / OpRegistry / REGISTER_OP -- the op half of the plugin surface.
// Mirrors core/framework/op.h:288-296 (REGISTER_OP -> OpRegistry::Global()->Register) and the
// builder spec-string grammar of core/framework/op_def_builder.h:
//   .Input("a: T")  .Input("inputs: N * T")  .Output("out: Ref(T)")
//   .Attr("transpose_a: bool = false")  .Attr("T: {half, float, double}")
//   .Attr("padding: {'SAME', 'VALID'}")  .Attr("strides: list(int)")  .Attr("N: int >= 1")
#ifndef B200TF_CORE_FRAMEWORK_OP_H_
#define B200TF_CORE_FRAMEWORK_OP_H_

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tensorflow/core/framework/node_def.h"

namespace tensorflow {

struct OpDef {
  struct ArgDef {
    std::string name;
    DataType type = DT_INVALID;  // fixed type, or
    std::string type_attr;       // name of a "type" attr
    std::string number_attr;     // "N * T": name of the int attr giving the count
    bool is_ref = false;
  };
  struct AttrDef {
    std::string name;
    std::string type;  // "string","int","float","bool","type","shape","tensor","list(int)",...
    bool has_default = false;
    AttrValue default_value;
    std::vector<DataType> allowed_types;      // for type attrs with {...}
    std::vector<std::string> allowed_strings; // for string attrs with {'a','b'}
    bool has_minimum = false;
    int64 minimum = 0;
  };
  std::string name;
  std::vector<ArgDef> input_arg;
  std::vector<ArgDef> output_arg;
  std::vector<AttrDef> attr;
  bool is_stateful = false;
  const AttrDef* FindAttr(const std::string& n) const {
    for (const auto& a : attr)
      if (a.name == n) return &a;
    return nullptr;
  }
};

class OpDefBuilder {
 public:
  explicit OpDefBuilder(const std::string& name) { def_.name = name; }
  OpDefBuilder& Input(const std::string& spec) { inputs_.push_back(spec); return *this; }
  OpDefBuilder& Output(const std::string& spec) { outputs_.push_back(spec); return *this; }
  OpDefBuilder& Attr(const std::string& spec) { attrs_.push_back(spec); return *this; }
  OpDefBuilder& SetIsStateful() { def_.is_stateful = true; return *this; }
  OpDefBuilder& SetIsCommutative() { return *this; }
  OpDefBuilder& SetAllowsUninitializedInput() { return *this; }
  OpDefBuilder& Doc(const std::string&) { return *this; }
  template <typename F> OpDefBuilder& SetShapeFn(F) { return *this; }  // shape inference is host-side only
  Status Finalize(OpDef* out) const;

 private:
  OpDef def_;
  std::vector<std::string> inputs_, outputs_, attrs_;
};

class OpRegistry {
 public:
  static OpRegistry* Global();
  Status Register(const OpDef& def);
  const OpDef* LookUp(const std::string& op_type_name) const;
  std::vector<std::string> ListOps() const;

 private:
  mutable std::mutex mu_;
  std::map<std::string, OpDef> registry_;
};

namespace register_op {
struct OpDefBuilderReceiver {
  OpDefBuilderReceiver(const OpDefBuilder& b);  // NOLINT: REGISTER_OP assigns the builder here
};
}  // namespace register_op

#define REGISTER_OP(name) REGISTER_OP_UNIQ_HELPER(__COUNTER__, name)
#define REGISTER_OP_UNIQ_HELPER(ctr, name) REGISTER_OP_UNIQ(ctr, name)
#define REGISTER_OP_UNIQ(ctr, name)                                                      \
  static ::tensorflow::register_op::OpDefBuilderReceiver register_op##ctr __attribute__( \
      (unused)) = ::tensorflow::OpDefBuilder(name)

// Applies attr defaults of the OpDef to the node, checks attr types / allowed values
// (node_def_util.cc ValidateNodeDef + AddDefaultsToNodeDef).
Status ValidateNodeDef(NodeDef* node, const OpDef& op_def);
// Resolves the dtypes of a node's inputs/outputs from its attrs (InOutTypesForNode).
Status InOutTypesForNode(const NodeDef& node, const OpDef& op_def, DataTypeVector* inputs,
                         DataTypeVector* outputs);

}  // namespace tensorflow
#endif
