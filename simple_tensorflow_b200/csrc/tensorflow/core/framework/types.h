// DataType, bfloat16, DeviceType, MemoryType -- subset of the reference's
// core/framework/types.{h,proto} and core/framework/numeric_types.h.
#ifndef B200TF_CORE_FRAMEWORK_TYPES_H_
#define B200TF_CORE_FRAMEWORK_TYPES_H_

#include <cstring>
#include <string>
#include <vector>

#include "tensorflow/core/lib/core/status.h"

namespace tensorflow {

// framework/types.proto:13-40 numbering.
enum DataType {
  DT_INVALID = 0,
  DT_FLOAT = 1,
  DT_DOUBLE = 2,
  DT_INT32 = 3,
  DT_UINT8 = 4,
  DT_INT16 = 5,
  DT_INT8 = 6,
  DT_STRING = 7,
  DT_COMPLEX64 = 8,
  DT_INT64 = 9,
  DT_BOOL = 10,
  DT_BFLOAT16 = 14,
  DT_COMPLEX128 = 18,
  DT_HALF = 19,
  DT_FLOAT_REF = 101,
};
typedef std::vector<DataType> DataTypeVector;

// framework/numeric_types.h:45-50: a 16-bit storage struct; conversion from float TRUNCATES
// (framework/bfloat16.cc:20-31).
struct bfloat16 {
  bfloat16() : value(0) {}
  explicit bfloat16(float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    value = static_cast<uint16_t>(u >> 16);
  }
  explicit operator float() const {
    uint32_t u = static_cast<uint32_t>(value) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  }
  uint16_t value;
};

// IEEE binary16 storage (the reference's Eigen::half, framework/numeric_types.h): the GPU ops take
// it through fp32 (core/kernels/half_ops.cc); the struct only carries the bits.
struct half {
  half() : value(0) {}
  uint16_t value;
};

inline size_t DataTypeSize(DataType dt) {
  switch (dt) {
    case DT_FLOAT: case DT_INT32: return 4;
    case DT_DOUBLE: case DT_INT64: case DT_COMPLEX64: return 8;
    case DT_BFLOAT16: case DT_HALF: case DT_INT16: return 2;
    case DT_UINT8: case DT_INT8: case DT_BOOL: return 1;
    default: return 0;
  }
}
inline std::string DataTypeString(DataType dt) {
  switch (dt) {
    case DT_FLOAT: return "float";
    case DT_DOUBLE: return "double";
    case DT_INT32: return "int32";
    case DT_INT64: return "int64";
    case DT_BFLOAT16: return "bfloat16";
    case DT_HALF: return "half";
    case DT_BOOL: return "bool";
    case DT_UINT8: return "uint8";
    case DT_INT8: return "int8";
    case DT_INT16: return "int16";
    case DT_STRING: return "string";
    case DT_INVALID: return "INVALID";
    default: return strings::StrCat("dtype(", static_cast<int>(dt), ")");
  }
}
inline bool DataTypeFromString(const std::string& s, DataType* dt) {
  static const struct { const char* n; DataType t; } kTable[] = {
      {"float", DT_FLOAT}, {"float32", DT_FLOAT}, {"double", DT_DOUBLE}, {"int32", DT_INT32},
      {"int64", DT_INT64}, {"bfloat16", DT_BFLOAT16}, {"half", DT_HALF}, {"bool", DT_BOOL},
      {"uint8", DT_UINT8}, {"int8", DT_INT8}, {"int16", DT_INT16}, {"string", DT_STRING},
      {"complex64", DT_COMPLEX64}, {"complex128", DT_COMPLEX128}};
  for (const auto& e : kTable)
    if (s == e.n) { *dt = e.t; return true; }
  return false;
}

template <typename T> struct DataTypeToEnum;
template <DataType V> struct EnumToDataType;
#define B200TF_MATCH_TYPE(TYPE, ENUM)                                  \
  template <> struct DataTypeToEnum<TYPE> {                            \
    static DataType v() { return ENUM; }                               \
    static constexpr DataType value = ENUM;                            \
  };                                                                   \
  template <> struct EnumToDataType<ENUM> { typedef TYPE Type; }
B200TF_MATCH_TYPE(float, DT_FLOAT);
B200TF_MATCH_TYPE(double, DT_DOUBLE);
B200TF_MATCH_TYPE(int32, DT_INT32);
B200TF_MATCH_TYPE(int64, DT_INT64);
B200TF_MATCH_TYPE(bfloat16, DT_BFLOAT16);
B200TF_MATCH_TYPE(half, DT_HALF);
B200TF_MATCH_TYPE(bool, DT_BOOL);
B200TF_MATCH_TYPE(uint8, DT_UINT8);
#undef B200TF_MATCH_TYPE

// framework/types.h: DeviceType is a string wrapper; DEVICE_CPU / DEVICE_GPU constants.
class DeviceType {
 public:
  DeviceType(const char* type) : type_(type) {}  // NOLINT
  explicit DeviceType(const std::string& type) : type_(type) {}
  const std::string& type() const { return type_; }
  bool operator==(const DeviceType& o) const { return type_ == o.type_; }
  bool operator!=(const DeviceType& o) const { return type_ != o.type_; }
 private:
  std::string type_;
};
static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

enum MemoryType { DEVICE_MEMORY = 0, HOST_MEMORY = 1 };
typedef std::vector<MemoryType> MemoryTypeVector;

}  // namespace tensorflow
#endif
