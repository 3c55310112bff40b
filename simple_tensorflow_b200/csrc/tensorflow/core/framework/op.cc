// OpDef spec-string parsing, OpRegistry, NodeDef validation.  See op.h.
#include "tensorflow/core/framework/op.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>

namespace tensorflow {
namespace {

std::string Trim(const std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && isspace(static_cast<unsigned char>(s[b]))) ++b;
  while (e > b && isspace(static_cast<unsigned char>(s[e - 1]))) --e;
  return s.substr(b, e - b);
}
std::vector<std::string> SplitTop(const std::string& s, char sep) {
  std::vector<std::string> out;
  std::string cur;
  bool in_quote = false;
  for (char c : s) {
    if (c == '\'') in_quote = !in_quote;
    if (c == sep && !in_quote) {
      out.push_back(Trim(cur));
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  if (!Trim(cur).empty()) out.push_back(Trim(cur));
  return out;
}

Status ParseArg(const std::string& spec, OpDef::ArgDef* arg) {
  const size_t colon = spec.find(':');
  if (colon == std::string::npos)
    return errors::InvalidArgument("Trouble parsing '<name>:' from arg spec '", spec, "'");
  arg->name = Trim(spec.substr(0, colon));
  std::string t = Trim(spec.substr(colon + 1));
  const size_t star = t.find('*');
  if (star != std::string::npos) {
    arg->number_attr = Trim(t.substr(0, star));
    t = Trim(t.substr(star + 1));
  }
  if (t.compare(0, 4, "Ref(") == 0 && t.back() == ')') {
    arg->is_ref = true;
    t = Trim(t.substr(4, t.size() - 5));
  }
  DataType dt;
  if (DataTypeFromString(t, &dt))
    arg->type = dt;
  else
    arg->type_attr = t;
  return Status::OK();
}

Status ParseDefault(const std::string& type, const std::string& text, AttrValue* v) {
  if (type == "bool") {
    if (text != "true" && text != "false")
      return errors::InvalidArgument("bad bool default '", text, "'");
    *v = AttrValue::B(text == "true");
  } else if (type == "int") {
    *v = AttrValue::I(strtoll(text.c_str(), nullptr, 10));
  } else if (type == "float") {
    *v = AttrValue::F(strtof(text.c_str(), nullptr));
  } else if (type == "string") {
    std::string s = text;
    if (s.size() >= 2 && (s.front() == '\'' || s.front() == '"')) s = s.substr(1, s.size() - 2);
    *v = AttrValue::S(s);
  } else if (type == "type") {
    std::string s = text;
    if (s.compare(0, 3, "DT_") == 0) {
      s = s.substr(3);
      std::transform(s.begin(), s.end(), s.begin(), ::tolower);
    }
    DataType dt;
    if (!DataTypeFromString(s, &dt)) return errors::InvalidArgument("bad type default '", text, "'");
    *v = AttrValue::Type(dt);
  } else if (type == "shape") {
    *v = AttrValue::Shape(TensorShape());
  } else if (type == "list(int)") {
    std::string s = text;
    if (!s.empty() && s.front() == '[') s = s.substr(1, s.size() - 2);
    std::vector<int64> vals;
    for (const auto& p : SplitTop(s, ',')) vals.push_back(strtoll(p.c_str(), nullptr, 10));
    *v = AttrValue::ListI(vals);
  } else {
    return errors::Unimplemented("default values for attr type '", type, "'");
  }
  return Status::OK();
}

Status ParseAttr(const std::string& spec, OpDef::AttrDef* attr) {
  const size_t colon = spec.find(':');
  if (colon == std::string::npos)
    return errors::InvalidArgument("Trouble parsing '<name>:' from Attr spec '", spec, "'");
  attr->name = Trim(spec.substr(0, colon));
  std::string rest = Trim(spec.substr(colon + 1));
  std::string def;
  // split off "= default" (not inside braces / quotes)
  int depth = 0;
  bool in_quote = false;
  for (size_t i = 0; i < rest.size(); ++i) {
    const char c = rest[i];
    if (c == '\'') in_quote = !in_quote;
    if (in_quote) continue;
    if (c == '{' || c == '(') ++depth;
    if (c == '}' || c == ')') --depth;
    if (c == '=' && depth == 0 && (i == 0 || rest[i - 1] != '>')) {
      def = Trim(rest.substr(i + 1));
      rest = Trim(rest.substr(0, i));
      break;
    }
  }
  if (!rest.empty() && rest.front() == '{') {
    const size_t close = rest.find('}');
    const std::string inner = rest.substr(1, close - 1);
    const auto parts = SplitTop(inner, ',');
    if (!parts.empty() && !parts[0].empty() && parts[0].front() == '\'') {
      attr->type = "string";
      for (const auto& p : parts) attr->allowed_strings.push_back(p.substr(1, p.size() - 2));
    } else {
      attr->type = "type";
      for (const auto& p : parts) {
        if (p == "numbertype") {
          for (DataType t : {DT_FLOAT, DT_DOUBLE, DT_INT64, DT_INT32, DT_UINT8, DT_INT16, DT_INT8,
                             DT_HALF, DT_BFLOAT16})
            attr->allowed_types.push_back(t);
          continue;
        }
        DataType dt;
        if (!DataTypeFromString(p, &dt))
          return errors::InvalidArgument("Unrecognized type '", p, "' in Attr spec '", spec, "'");
        attr->allowed_types.push_back(dt);
      }
    }
  } else if (rest == "numbertype" || rest == "realnumbertype") {
    attr->type = "type";
    for (DataType t : {DT_FLOAT, DT_DOUBLE, DT_INT64, DT_INT32, DT_UINT8, DT_INT16, DT_INT8,
                       DT_HALF, DT_BFLOAT16})
      attr->allowed_types.push_back(t);
  } else {
    const size_t ge = rest.find(">=");
    if (ge != std::string::npos) {
      attr->has_minimum = true;
      attr->minimum = strtoll(rest.c_str() + ge + 2, nullptr, 10);
      rest = Trim(rest.substr(0, ge));
    }
    attr->type = rest;
  }
  if (!def.empty()) {
    attr->has_default = true;
    TF_RETURN_IF_ERROR(ParseDefault(attr->type, def, &attr->default_value));
  }
  return Status::OK();
}

bool KindMatches(const std::string& type, const AttrValue& v) {
  switch (v.kind) {
    case AttrValue::kS: return type == "string";
    case AttrValue::kI: return type == "int";
    case AttrValue::kF: return type == "float";
    case AttrValue::kB: return type == "bool";
    case AttrValue::kType: return type == "type";
    case AttrValue::kShape: return type == "shape";
    case AttrValue::kTensor: return type == "tensor";
    case AttrValue::kListI: return type == "list(int)";
    case AttrValue::kListS: return type == "list(string)";
    case AttrValue::kListType: return type == "list(type)";
    case AttrValue::kRaw: return true;  // imported, not modelled here (checked by whoever uses it)
    default: return false;
  }
}

}  // namespace

Status OpDefBuilder::Finalize(OpDef* out) const {
  *out = def_;
  for (const auto& s : attrs_) {
    OpDef::AttrDef a;
    TF_RETURN_IF_ERROR(ParseAttr(s, &a));
    out->attr.push_back(a);
  }
  for (const auto& s : inputs_) {
    OpDef::ArgDef a;
    TF_RETURN_IF_ERROR(ParseArg(s, &a));
    out->input_arg.push_back(a);
  }
  for (const auto& s : outputs_) {
    OpDef::ArgDef a;
    TF_RETURN_IF_ERROR(ParseArg(s, &a));
    out->output_arg.push_back(a);
  }
  return Status::OK();
}

OpRegistry* OpRegistry::Global() {
  static OpRegistry* r = new OpRegistry;
  return r;
}
Status OpRegistry::Register(const OpDef& def) {
  std::lock_guard<std::mutex> l(mu_);
  if (registry_.count(def.name))
    return errors::AlreadyExists("Op with name ", def.name);
  registry_[def.name] = def;
  return Status::OK();
}
const OpDef* OpRegistry::LookUp(const std::string& name) const {
  std::lock_guard<std::mutex> l(mu_);
  auto it = registry_.find(name);
  return it == registry_.end() ? nullptr : &it->second;
}
std::vector<std::string> OpRegistry::ListOps() const {
  std::lock_guard<std::mutex> l(mu_);
  std::vector<std::string> out;
  for (const auto& kv : registry_) out.push_back(kv.first);
  return out;
}

namespace register_op {
OpDefBuilderReceiver::OpDefBuilderReceiver(const OpDefBuilder& b) {
  OpDef def;
  Status s = b.Finalize(&def);
  if (s.ok()) s = OpRegistry::Global()->Register(def);
  if (!s.ok()) {
    fprintf(stderr, "REGISTER_OP failed: %s\n", s.ToString().c_str());
    abort();  // the reference LOG(FATAL)s on a bad registration as well (op.cc)
  }
}
}  // namespace register_op

Status ValidateNodeDef(NodeDef* node, const OpDef& op_def) {
  for (const auto& a : op_def.attr) {
    auto it = node->attr.find(a.name);
    if (it == node->attr.end()) {
      if (!a.has_default)
        return errors::InvalidArgument("NodeDef missing attr '", a.name, "' from Op<name=",
                                       op_def.name, ">; NodeDef: ", SummarizeNodeDef(*node));
      node->attr[a.name] = a.default_value;
      continue;
    }
    const AttrValue& v = it->second;
    if (!KindMatches(a.type, v))
      return errors::InvalidArgument("AttrValue for attr '", a.name, "' of node '", node->name,
                                     "' does not have the declared type ", a.type);
    if (a.type == "type" && !a.allowed_types.empty() &&
        std::find(a.allowed_types.begin(), a.allowed_types.end(), v.type) ==
            a.allowed_types.end())
      return errors::InvalidArgument("Value for attr '", a.name, "' of ", DataTypeString(v.type),
                                     " is not in the list of allowed values for Op ", op_def.name);
    if (a.type == "string" && !a.allowed_strings.empty() &&
        std::find(a.allowed_strings.begin(), a.allowed_strings.end(), v.s) ==
            a.allowed_strings.end())
      return errors::InvalidArgument("Value for attr '", a.name, "' of \"", v.s,
                                     "\" is not in the list of allowed values for Op ",
                                     op_def.name);
    if (a.type == "list(int)" && a.has_minimum && (int64)v.list_i.size() < a.minimum)
      return errors::InvalidArgument("Length for attr '", a.name, "' of ", v.list_i.size(),
                                     " must be at least minimum ", a.minimum);
    if (a.type == "int" && a.has_minimum && v.i < a.minimum)
      return errors::InvalidArgument("Value for attr '", a.name, "' of ", v.i,
                                     " must be at least minimum ", a.minimum);
  }
  for (const auto& kv : node->attr) {
    if (!kv.first.empty() && kv.first[0] == '_') continue;  // "_kernel", "_class", ...
    if (!op_def.FindAttr(kv.first))
      return errors::InvalidArgument("NodeDef mentions attr '", kv.first, "' not in Op<name=",
                                     op_def.name, ">; NodeDef: ", SummarizeNodeDef(*node));
  }
  return Status::OK();
}

static Status ArgTypes(const NodeDef& node, const OpDef::ArgDef& arg, DataTypeVector* out) {
  DataType dt = arg.type;
  if (dt == DT_INVALID) {
    TF_RETURN_IF_ERROR(GetNodeAttr(node, arg.type_attr, &dt));
  }
  int64 count = 1;
  if (!arg.number_attr.empty()) TF_RETURN_IF_ERROR(GetNodeAttr(node, arg.number_attr, &count));
  for (int64 i = 0; i < count; ++i) out->push_back(dt);
  return Status::OK();
}

Status InOutTypesForNode(const NodeDef& node, const OpDef& op_def, DataTypeVector* inputs,
                         DataTypeVector* outputs) {
  inputs->clear();
  outputs->clear();
  for (const auto& a : op_def.input_arg) TF_RETURN_IF_ERROR(ArgTypes(node, a, inputs));
  for (const auto& a : op_def.output_arg) TF_RETURN_IF_ERROR(ArgTypes(node, a, outputs));
  return Status::OK();
}

// ------------------------------------------------------------------ node_def helpers
static Status Find(const NodeDef& n, const std::string& name, AttrValue::Kind kind,
                   const char* kind_name, const AttrValue** out) {
  auto it = n.attr.find(name);
  if (it == n.attr.end())
    return errors::NotFound("No attr named '", name, "' in NodeDef: ", SummarizeNodeDef(n));
  if (it->second.kind != kind)
    return errors::InvalidArgument("Attr '", name, "' of node '", n.name, "' is not of type ",
                                   kind_name);
  *out = &it->second;
  return Status::OK();
}
#define B200TF_GET_ATTR(TYPE, KIND, KNAME, EXPR)                                   \
  Status GetNodeAttr(const NodeDef& n, const std::string& name, TYPE* v) {         \
    const AttrValue* a = nullptr;                                                  \
    TF_RETURN_IF_ERROR(Find(n, name, AttrValue::KIND, KNAME, &a));                 \
    *v = EXPR;                                                                     \
    return Status::OK();                                                           \
  }
B200TF_GET_ATTR(std::string, kS, "string", a->s)
B200TF_GET_ATTR(int64, kI, "int", a->i)
B200TF_GET_ATTR(int32, kI, "int", static_cast<int32>(a->i))
B200TF_GET_ATTR(float, kF, "float", a->f)
B200TF_GET_ATTR(bool, kB, "bool", a->b)
B200TF_GET_ATTR(DataType, kType, "type", a->type)
B200TF_GET_ATTR(TensorShape, kShape, "shape", a->shape)
B200TF_GET_ATTR(Tensor, kTensor, "tensor", a->tensor)
B200TF_GET_ATTR(std::vector<int64>, kListI, "list(int)", a->list_i)
B200TF_GET_ATTR(std::vector<std::string>, kListS, "list(string)", a->list_s)
#undef B200TF_GET_ATTR
Status GetNodeAttr(const NodeDef& n, const std::string& name, std::vector<int32>* v) {
  std::vector<int64> tmp;
  TF_RETURN_IF_ERROR(GetNodeAttr(n, name, &tmp));
  v->assign(tmp.begin(), tmp.end());
  return Status::OK();
}

std::string SummarizeNodeDef(const NodeDef& n) {
  std::string s = n.name + " = " + n.op + "[";
  bool first = true;
  for (const auto& kv : n.attr) {
    if (!first) s += ", ";
    first = false;
    s += kv.first;
  }
  s += "](";
  for (size_t i = 0; i < n.input.size(); ++i) {
    if (i) s += ", ";
    s += n.input[i];
  }
  return s + ")";
}

}  // namespace tensorflow
