#include "tensorflow/core/framework/graph_def_wire.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace tensorflow {
namespace {

// ------------------------------------------------------------------ wire primitives
// https://protobuf.dev/programming-guides/encoding: tag = (field << 3) | wire type;
// 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32.
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  Reader(const void* d, size_t n)
      : p(static_cast<const uint8_t*>(d)), end(static_cast<const uint8_t*>(d) + n) {}
  bool done() const { return p >= end || !ok; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) break;
      const uint8_t b = *p++;
      v |= static_cast<uint64_t>(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  uint32_t fixed32() {
    if (end - p < 4) {
      ok = false;
      return 0;
    }
    uint32_t v;
    std::memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  uint64_t fixed64() {
    if (end - p < 8) {
      ok = false;
      return 0;
    }
    uint64_t v;
    std::memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  // length-delimited payload as a sub-reader
  Reader bytes() {
    const uint64_t n = varint();
    if (!ok || n > static_cast<uint64_t>(end - p)) {
      ok = false;
      return Reader(p, 0);
    }
    Reader r(p, static_cast<size_t>(n));
    p += n;
    return r;
  }
  std::string str() {
    Reader r = bytes();
    return std::string(reinterpret_cast<const char*>(r.p), r.end - r.p);
  }
  void skip(int wire_type) {
    switch (wire_type) {
      case 0: varint(); break;
      case 1: fixed64(); break;
      case 2: bytes(); break;
      case 5: fixed32(); break;
      default: ok = false;  // groups (3, 4) do not occur in these protos
    }
  }
};

struct Writer {
  std::string* out;
  explicit Writer(std::string* o) : out(o) {}
  void varint(uint64_t v) {
    while (v >= 0x80) {
      out->push_back(static_cast<char>((v & 0x7F) | 0x80));
      v >>= 7;
    }
    out->push_back(static_cast<char>(v));
  }
  void tag(int field, int wire_type) { varint(static_cast<uint64_t>(field) << 3 | wire_type); }
  void bytes(int field, const std::string& s) {
    tag(field, 2);
    varint(s.size());
    out->append(s);
  }
  void bytes(int field, const void* d, size_t n) {
    tag(field, 2);
    varint(n);
    out->append(static_cast<const char*>(d), n);
  }
  void varint_field(int field, uint64_t v) {
    tag(field, 0);
    varint(v);
  }
  void fixed32_field(int field, uint32_t v) {
    tag(field, 5);
    out->append(reinterpret_cast<const char*>(&v), 4);
  }
};

// ------------------------------------------------------------------ tensor_shape.proto
// TensorShapeProto{ repeated Dim dim = 2 { int64 size = 1; string name = 2 }; bool unknown_rank = 3 }
// Returns false for shapes TensorShape cannot carry (unknown rank / unknown dims / named dims).
bool ParseShape(Reader r, TensorShape* shape) {
  shape->Clear();
  bool representable = true;
  while (!r.done()) {
    const uint64_t t = r.varint();
    if ((t >> 3) == 2 && (t & 7) == 2) {
      Reader d = r.bytes();
      int64 size = 0;
      while (!d.done()) {
        const uint64_t dt = d.varint();
        if ((dt >> 3) == 1 && (dt & 7) == 0) {
          size = static_cast<int64>(d.varint());
        } else {
          if ((dt >> 3) == 2) representable = false;  // named dimension
          d.skip(dt & 7);
        }
      }
      if (!d.ok) return false;
      if (size < 0) representable = false;
      shape->AddDim(size);
      if (shape->dims() > 254) return false;  // TensorShape::MaxDimensions()
    } else {
      if ((t >> 3) == 3 && (t & 7) == 0) {
        if (r.varint()) representable = false;
      } else {
        r.skip(t & 7);
      }
    }
  }
  return r.ok && representable;
}
void WriteShape(const TensorShape& shape, std::string* out) {
  Writer w(out);
  for (int i = 0; i < shape.dims(); ++i) {
    std::string dim;
    Writer dw(&dim);
    // proto3 omits zero scalars, but TensorFlow writes `size` for every Dim it adds
    if (shape.dim_size(i) != 0) dw.varint_field(1, static_cast<uint64_t>(shape.dim_size(i)));
    w.bytes(2, dim);
  }
}

// ------------------------------------------------------------------ tensor.proto
// TensorProto{ dtype=1, tensor_shape=2, version_number=3, tensor_content=4, float_val=5,
//   double_val=6, int_val=7, string_val=8, scomplex_val=9, int64_val=10, bool_val=11,
//   dcomplex_val=12, half_val=13 }.  Typed *_val fields may be packed or not; fewer values than
// elements means "repeat the last one" (tensor.cc FromProtoField).
template <typename T>
void FillTail(T* data, int64 have, int64 n) {
  if (n <= 0) return;
  const T last = have > 0 ? data[have - 1] : T();
  for (int64 i = have; i < n; ++i) data[i] = last;
}
bool ParseTensor(Reader r, Tensor* out) {
  DataType dtype = DT_INVALID;
  TensorShape shape;
  Reader content(nullptr, 0);
  bool has_content = false;
  std::vector<uint32_t> f32;      // float_val bit patterns
  std::vector<uint64_t> f64;      // double_val bit patterns
  std::vector<int64> ints;        // int_val / int64_val / bool_val / half_val
  bool unsupported = false;
  while (!r.done()) {
    const uint64_t t = r.varint();
    const int field = static_cast<int>(t >> 3), wt = static_cast<int>(t & 7);
    switch (field) {
      case 1:
        if (wt != 0) return false;  // dtype is a varint enum: anything else is a malformed proto
        dtype = static_cast<DataType>(r.varint());
        break;
      case 2:
        if (wt != 2 || !ParseShape(r.bytes(), &shape)) return false;
        break;
      case 4:
        if (wt != 2) return false;
        content = r.bytes();
        has_content = true;
        break;
      case 5:
        if (wt == 2) {
          Reader v = r.bytes();
          while (!v.done()) f32.push_back(v.fixed32());
        } else {
          f32.push_back(r.fixed32());
        }
        break;
      case 6:
        if (wt == 2) {
          Reader v = r.bytes();
          while (!v.done()) f64.push_back(v.fixed64());
        } else {
          f64.push_back(r.fixed64());
        }
        break;
      case 7: case 10: case 11: case 13:
        if (wt == 2) {
          Reader v = r.bytes();
          while (!v.done()) ints.push_back(static_cast<int64>(v.varint()));
        } else {
          ints.push_back(static_cast<int64>(r.varint()));
        }
        break;
      case 8: case 9: case 12: case 14:
        unsupported = true;
        r.skip(wt);
        break;
      default: r.skip(wt);
    }
  }
  const size_t esize = DataTypeSize(dtype);
  if (!r.ok || unsupported || esize == 0 || dtype == DT_COMPLEX64 || dtype == DT_COMPLEX128)
    return false;
  // Untrusted dimensions: reject negative / overflowing products before any allocation
  // (TensorShape::IsValid in the reference, tensor.cc Tensor::FromProto), and never allocate
  // more than the proto itself can describe -- a typed *_val list repeats its last value, so a
  // tiny proto may legally describe a large tensor, but not an absurd one.
  if (!shape.IsValid(esize)) return false;
  if (static_cast<unsigned long long>(shape.num_elements()) * esize > (1ull << 34)) return false;
  Tensor t(dtype, shape);
  const int64 n = t.NumElements();
  if (n > 0 && !t.IsInitialized()) return false;
  if (n == 0) {
    *out = t;
    return true;
  }
  char* dst = static_cast<char*>(t.raw_data());
  if (has_content) {
    if (static_cast<size_t>(content.end - content.p) != static_cast<size_t>(n) * esize) return false;
    std::memcpy(dst, content.p, static_cast<size_t>(n) * esize);
  } else if (dtype == DT_FLOAT) {
    const int64 have = std::min<int64>(n, f32.size());
    std::memcpy(dst, f32.data(), have * 4);
    FillTail(reinterpret_cast<uint32_t*>(dst), have, n);
  } else if (dtype == DT_DOUBLE) {
    const int64 have = std::min<int64>(n, f64.size());
    std::memcpy(dst, f64.data(), have * 8);
    FillTail(reinterpret_cast<uint64_t*>(dst), have, n);
  } else {
    const int64 have = std::min<int64>(n, ints.size());
    for (int64 i = 0; i < n; ++i) {
      const int64 v = have == 0 ? 0 : ints[i < have ? i : have - 1];
      switch (esize) {
        case 1: reinterpret_cast<int8_t*>(dst)[i] = static_cast<int8_t>(v); break;
        case 2: reinterpret_cast<int16_t*>(dst)[i] = static_cast<int16_t>(v); break;
        case 4: reinterpret_cast<int32_t*>(dst)[i] = static_cast<int32_t>(v); break;
        default: reinterpret_cast<int64*>(dst)[i] = v;
      }
    }
  }
  *out = t;
  return true;
}
void WriteTensor(const Tensor& t, std::string* out) {
  Writer w(out);
  w.varint_field(1, static_cast<uint64_t>(t.dtype()));
  std::string shape;
  WriteShape(t.shape(), &shape);
  w.bytes(2, shape);
  if (t.NumElements() > 0 && t.raw_data() != nullptr) w.bytes(4, t.raw_data(), t.TotalBytes());
}

// ------------------------------------------------------------------ attr_value.proto
// AttrValue{ list=1, s=2, i=3, f=4, b=5, type=6, shape=7, tensor=8, placeholder=9, func=10 }
// ListValue{ s=2, i=3, f=4, b=5, type=6, shape=7, tensor=8, func=9 } (scalars packed or not)
bool ParseList(Reader r, AttrValue* a) {
  std::vector<int64> li;
  std::vector<std::string> ls;
  std::vector<DataType> lt;
  bool other = false;
  while (!r.done()) {
    const uint64_t t = r.varint();
    const int field = static_cast<int>(t >> 3), wt = static_cast<int>(t & 7);
    if (field == 2 && wt == 2) {
      ls.push_back(r.str());
    } else if (field == 3 || field == 6) {
      std::vector<int64> vals;
      if (wt == 2) {
        Reader v = r.bytes();
        while (!v.done()) vals.push_back(static_cast<int64>(v.varint()));
      } else {
        vals.push_back(static_cast<int64>(r.varint()));
      }
      for (int64 v : vals) {
        if (field == 3)
          li.push_back(v);
        else
          lt.push_back(static_cast<DataType>(v));
      }
    } else {
      other = true;
      r.skip(wt);
    }
  }
  if (!r.ok) return false;
  const int kinds = !li.empty() + !ls.empty() + !lt.empty();
  if (other || kinds != 1) return false;  // empty or mixed lists stay raw
  if (!li.empty()) *a = AttrValue::ListI(li);
  if (!ls.empty()) *a = AttrValue::ListS(ls);
  if (!lt.empty()) {
    a->kind = AttrValue::kListType;
    a->list_type = lt;
  }
  return true;
}
Status ParseAttrValue(Reader r, AttrValue* a) {
  const uint8_t* begin = r.p;
  const size_t size = r.end - r.p;
  bool modelled = true;
  *a = AttrValue();
  while (!r.done()) {
    const uint64_t t = r.varint();
    const int field = static_cast<int>(t >> 3), wt = static_cast<int>(t & 7);
    if (field == 2 && wt == 2) {
      *a = AttrValue::S(r.str());
    } else if (field == 3 && wt == 0) {
      *a = AttrValue::I(static_cast<int64>(r.varint()));
    } else if (field == 4 && wt == 5) {
      const uint32_t bits = r.fixed32();
      float f;
      std::memcpy(&f, &bits, 4);
      *a = AttrValue::F(f);
    } else if (field == 5 && wt == 0) {
      *a = AttrValue::B(r.varint() != 0);
    } else if (field == 6 && wt == 0) {
      *a = AttrValue::Type(static_cast<DataType>(r.varint()));
    } else if (field == 7 && wt == 2) {
      TensorShape shape;
      if (ParseShape(r.bytes(), &shape))
        *a = AttrValue::Shape(shape);
      else
        modelled = false;
    } else if (field == 8 && wt == 2) {
      Tensor tensor;
      if (ParseTensor(r.bytes(), &tensor))
        *a = AttrValue::TensorV(tensor);
      else
        modelled = false;
    } else if (field == 1 && wt == 2) {
      if (!ParseList(r.bytes(), a)) modelled = false;
    } else {
      modelled = false;
      r.skip(wt);
    }
  }
  if (!r.ok) return errors::InvalidArgument("malformed AttrValue");
  if (!modelled) {
    *a = AttrValue();
    a->kind = AttrValue::kRaw;
    a->raw.assign(reinterpret_cast<const char*>(begin), size);
  }
  return Status::OK();
}
void WriteAttrValue(const AttrValue& a, std::string* out) {
  Writer w(out);
  switch (a.kind) {
    case AttrValue::kNone: break;
    case AttrValue::kRaw: out->append(a.raw); break;
    case AttrValue::kS: w.bytes(2, a.s); break;
    case AttrValue::kI: w.varint_field(3, static_cast<uint64_t>(a.i)); break;
    case AttrValue::kF: {
      uint32_t bits;
      std::memcpy(&bits, &a.f, 4);
      w.fixed32_field(4, bits);
      break;
    }
    case AttrValue::kB: w.varint_field(5, a.b ? 1 : 0); break;
    case AttrValue::kType: w.varint_field(6, static_cast<uint64_t>(a.type)); break;
    case AttrValue::kShape: {
      std::string s;
      WriteShape(a.shape, &s);
      w.bytes(7, s);
      break;
    }
    case AttrValue::kTensor: {
      std::string s;
      WriteTensor(a.tensor, &s);
      w.bytes(8, s);
      break;
    }
    case AttrValue::kListI: case AttrValue::kListS: case AttrValue::kListType: {
      std::string list;
      Writer lw(&list);
      if (a.kind == AttrValue::kListS) {
        for (const std::string& s : a.list_s) lw.bytes(2, s);
      } else {
        std::string packed;
        Writer pw(&packed);
        if (a.kind == AttrValue::kListI)
          for (int64 v : a.list_i) pw.varint(static_cast<uint64_t>(v));
        else
          for (DataType v : a.list_type) pw.varint(static_cast<uint64_t>(v));
        if (!packed.empty()) lw.bytes(a.kind == AttrValue::kListI ? 3 : 6, packed);
      }
      w.bytes(1, list);
      break;
    }
  }
}

// ------------------------------------------------------------------ node_def.proto
// NodeDef{ name=1, op=2, repeated input=3, device=4, map<string, AttrValue> attr=5 }
// (a map field is a repeated message { key=1, value=2 }).
Status ParseNodeDef(Reader r, NodeDef* n) {
  while (!r.done()) {
    const uint64_t t = r.varint();
    const int field = static_cast<int>(t >> 3), wt = static_cast<int>(t & 7);
    if (wt != 2) {
      r.skip(wt);
      continue;
    }
    switch (field) {
      case 1: n->name = r.str(); break;
      case 2: n->op = r.str(); break;
      case 3: n->input.push_back(r.str()); break;
      case 4: n->device = r.str(); break;
      case 5: {
        Reader e = r.bytes();
        std::string key;
        AttrValue value;
        while (!e.done()) {
          const uint64_t et = e.varint();
          if ((et >> 3) == 1 && (et & 7) == 2) {
            key = e.str();
          } else if ((et >> 3) == 2 && (et & 7) == 2) {
            TF_RETURN_IF_ERROR(ParseAttrValue(e.bytes(), &value));
          } else {
            e.skip(et & 7);
          }
        }
        if (!e.ok) return errors::InvalidArgument("malformed attr entry in node '", n->name, "'");
        n->attr[key] = value;
        break;
      }
      default: r.bytes();
    }
  }
  if (!r.ok) return errors::InvalidArgument("malformed NodeDef", n->name.empty() ? "" : " '" + n->name + "'");
  return Status::OK();
}
void WriteNodeDef(const NodeDef& n, std::string* out) {
  Writer w(out);
  if (!n.name.empty()) w.bytes(1, n.name);
  if (!n.op.empty()) w.bytes(2, n.op);
  for (const std::string& in : n.input) w.bytes(3, in);
  if (!n.device.empty()) w.bytes(4, n.device);
  for (const auto& kv : n.attr) {
    std::string entry, value;
    Writer ew(&entry);
    ew.bytes(1, kv.first);
    WriteAttrValue(kv.second, &value);
    ew.bytes(2, value);
    w.bytes(5, entry);
  }
}

std::string Quote(const std::string& s) {
  std::string out = "\"";
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') {
      out += '\\';
      out += static_cast<char>(c);
    } else if (c < 32 || c > 126) {
      char buf[8];
      snprintf(buf, sizeof(buf), "\\%03o", c);
      out += buf;
    } else {
      out += static_cast<char>(c);
    }
  }
  return out + "\"";
}

}  // namespace

// graph.proto: GraphDef{ repeated NodeDef node = 1; FunctionDefLibrary library = 2;
//                        int32 version = 3 (deprecated); VersionDef versions = 4 }
Status ParseGraphDef(const void* data, size_t size, GraphDef* out) {
  *out = GraphDef();
  Reader r(data, size);
  while (!r.done()) {
    const uint64_t t = r.varint();
    const int field = static_cast<int>(t >> 3), wt = static_cast<int>(t & 7);
    if (field == 1 && wt == 2) {
      out->node.emplace_back();
      TF_RETURN_IF_ERROR(ParseNodeDef(r.bytes(), &out->node.back()));
    } else if (field == 2 && wt == 2) {
      out->library_raw = r.str();
    } else if (field == 4 && wt == 2) {
      out->versions_raw = r.str();
    } else {
      r.skip(wt);
    }
  }
  if (!r.ok) return errors::InvalidArgument("Invalid GraphDef: truncated or malformed protobuf");
  return Status::OK();
}

void SerializeGraphDef(const GraphDef& graph, std::string* out) {
  out->clear();
  Writer w(out);
  for (const NodeDef& n : graph.node) {
    std::string node;
    WriteNodeDef(n, &node);
    w.bytes(1, node);
  }
  if (!graph.library_raw.empty()) w.bytes(2, graph.library_raw);
  if (!graph.versions_raw.empty()) w.bytes(4, graph.versions_raw);
}

std::string GraphDefDebugString(const GraphDef& graph) {
  std::string out;
  for (const NodeDef& n : graph.node) {
    out += "node\t" + n.name + "\t" + n.op + "\t" + n.device + "\t";
    for (size_t i = 0; i < n.input.size(); ++i) out += (i ? "," : "") + n.input[i];
    out += "\t";
    bool first = true;
    for (const auto& kv : n.attr) {
      if (!first) out += ";";
      first = false;
      const AttrValue& a = kv.second;
      out += kv.first + "=";
      switch (a.kind) {
        case AttrValue::kNone: out += "none"; break;
        case AttrValue::kRaw: out += "raw:" + std::to_string(a.raw.size()); break;
        case AttrValue::kS: out += "s:" + Quote(a.s); break;
        case AttrValue::kI: out += "i:" + std::to_string(a.i); break;
        case AttrValue::kF: {
          char buf[32];
          snprintf(buf, sizeof(buf), "f:%.9g", a.f);
          out += buf;
          break;
        }
        case AttrValue::kB: out += a.b ? "b:true" : "b:false"; break;
        case AttrValue::kType: out += "type:" + std::to_string(static_cast<int>(a.type)); break;
        case AttrValue::kShape: out += "shape:" + a.shape.DebugString(); break;
        case AttrValue::kTensor: {
          out += "tensor:" + std::to_string(static_cast<int>(a.tensor.dtype())) + ":" +
                 a.tensor.shape().DebugString() + ":";
          const int64 n_show = std::min<int64>(a.tensor.NumElements(), 8);
          for (int64 i = 0; i < n_show; ++i) {
            char buf[40];
            switch (a.tensor.dtype()) {
              case DT_FLOAT: snprintf(buf, sizeof(buf), "%.9g", a.tensor.data<float>()[i]); break;
              case DT_DOUBLE: snprintf(buf, sizeof(buf), "%.17g", a.tensor.data<double>()[i]); break;
              case DT_INT32: snprintf(buf, sizeof(buf), "%d", a.tensor.data<int32>()[i]); break;
              case DT_INT64:
                snprintf(buf, sizeof(buf), "%lld", static_cast<long long>(a.tensor.data<int64>()[i]));
                break;
              default: snprintf(buf, sizeof(buf), "?");
            }
            out += (i ? "," : "") + std::string(buf);
          }
          break;
        }
        case AttrValue::kListI: {
          out += "list_i:[";
          for (size_t i = 0; i < a.list_i.size(); ++i)
            out += (i ? "," : "") + std::to_string(a.list_i[i]);
          out += "]";
          break;
        }
        case AttrValue::kListS: {
          out += "list_s:[";
          for (size_t i = 0; i < a.list_s.size(); ++i) out += (i ? "," : "") + Quote(a.list_s[i]);
          out += "]";
          break;
        }
        case AttrValue::kListType: {
          out += "list_type:[";
          for (size_t i = 0; i < a.list_type.size(); ++i)
            out += (i ? "," : "") + std::to_string(static_cast<int>(a.list_type[i]));
          out += "]";
          break;
        }
      }
    }
    out += "\n";
  }
  return out;
}

}  // namespace tensorflow
