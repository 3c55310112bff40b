// tensorflow::Status / tensorflow::errors -- source-compatible subset of the reference's
// core/lib/core/status.h and core/lib/core/errors.h (codes: core/lib/core/error_codes.proto).
#ifndef B200TF_CORE_LIB_CORE_STATUS_H_
#define B200TF_CORE_LIB_CORE_STATUS_H_

#include <cstdint>
#include <sstream>
#include <string>
#include <utility>

namespace tensorflow {

using int32 = int32_t;
using int64 = long long;
using uint8 = uint8_t;
using uint16 = uint16_t;
using uint32 = uint32_t;
using uint64 = unsigned long long;
using string = std::string;
typedef const std::string& StringPiece;

namespace error {
enum Code {
  OK = 0,
  CANCELLED = 1,
  UNKNOWN = 2,
  INVALID_ARGUMENT = 3,
  DEADLINE_EXCEEDED = 4,
  NOT_FOUND = 5,
  ALREADY_EXISTS = 6,
  PERMISSION_DENIED = 7,
  RESOURCE_EXHAUSTED = 8,
  FAILED_PRECONDITION = 9,
  ABORTED = 10,
  OUT_OF_RANGE = 11,
  UNIMPLEMENTED = 12,
  INTERNAL = 13,
  UNAVAILABLE = 14,
  DATA_LOSS = 15,
  UNAUTHENTICATED = 16
};
}  // namespace error

class Status {
 public:
  Status() : code_(error::OK) {}
  Status(error::Code code, const std::string& msg) : code_(code), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return code_ == error::OK; }
  error::Code code() const { return code_; }
  const std::string& error_message() const { return msg_; }
  // First error wins (framework/op_kernel.cc:247-249 relies on this).
  void Update(const Status& s) {
    if (ok()) *this = s;
  }
  std::string ToString() const;
  bool operator==(const Status& o) const { return code_ == o.code_ && msg_ == o.msg_; }
  bool operator!=(const Status& o) const { return !(*this == o); }

 private:
  error::Code code_;
  std::string msg_;
};

const char* error_code_name(error::Code c);
inline std::string Status::ToString() const {
  if (ok()) return "OK";
  return std::string(error_code_name(code_)) + ": " + msg_;
}
inline const char* error_code_name(error::Code c) {
  switch (c) {
    case error::OK: return "OK";
    case error::CANCELLED: return "Cancelled";
    case error::UNKNOWN: return "Unknown";
    case error::INVALID_ARGUMENT: return "Invalid argument";
    case error::DEADLINE_EXCEEDED: return "Deadline exceeded";
    case error::NOT_FOUND: return "Not found";
    case error::ALREADY_EXISTS: return "Already exists";
    case error::PERMISSION_DENIED: return "Permission denied";
    case error::RESOURCE_EXHAUSTED: return "Resource exhausted";
    case error::FAILED_PRECONDITION: return "Failed precondition";
    case error::ABORTED: return "Aborted";
    case error::OUT_OF_RANGE: return "Out of range";
    case error::UNIMPLEMENTED: return "Unimplemented";
    case error::INTERNAL: return "Internal";
    case error::UNAVAILABLE: return "Unavailable";
    case error::DATA_LOSS: return "Data loss";
    case error::UNAUTHENTICATED: return "Unauthenticated";
  }
  return "Unknown code";
}

namespace strings {
inline void StrAppendTo(std::ostringstream&) {}
template <typename T, typename... Rest>
void StrAppendTo(std::ostringstream& os, const T& v, const Rest&... rest) {
  os << v;
  StrAppendTo(os, rest...);
}
template <typename... Args>
std::string StrCat(const Args&... args) {
  std::ostringstream os;
  StrAppendTo(os, args...);
  return os.str();
}
}  // namespace strings

namespace errors {
#define B200TF_DECLARE_ERROR(FUNC, CONST)                          \
  template <typename... Args>                                      \
  Status FUNC(const Args&... args) {                               \
    return Status(error::CONST, strings::StrCat(args...));         \
  }                                                                \
  inline bool Is##FUNC(const Status& s) { return s.code() == error::CONST; }
B200TF_DECLARE_ERROR(Cancelled, CANCELLED)
B200TF_DECLARE_ERROR(InvalidArgument, INVALID_ARGUMENT)
B200TF_DECLARE_ERROR(NotFound, NOT_FOUND)
B200TF_DECLARE_ERROR(AlreadyExists, ALREADY_EXISTS)
B200TF_DECLARE_ERROR(ResourceExhausted, RESOURCE_EXHAUSTED)
B200TF_DECLARE_ERROR(Unavailable, UNAVAILABLE)
B200TF_DECLARE_ERROR(FailedPrecondition, FAILED_PRECONDITION)
B200TF_DECLARE_ERROR(OutOfRange, OUT_OF_RANGE)
B200TF_DECLARE_ERROR(Unimplemented, UNIMPLEMENTED)
B200TF_DECLARE_ERROR(Internal, INTERNAL)
B200TF_DECLARE_ERROR(Aborted, ABORTED)
B200TF_DECLARE_ERROR(DeadlineExceeded, DEADLINE_EXCEEDED)
B200TF_DECLARE_ERROR(DataLoss, DATA_LOSS)
B200TF_DECLARE_ERROR(Unknown, UNKNOWN)
#undef B200TF_DECLARE_ERROR
}  // namespace errors

#define TF_RETURN_IF_ERROR(...)                          \
  do {                                                   \
    const ::tensorflow::Status _status = (__VA_ARGS__);  \
    if (!_status.ok()) return _status;                   \
  } while (0)

}  // namespace tensorflow
#endif
