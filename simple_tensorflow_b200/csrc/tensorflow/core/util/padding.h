// Padding / TensorFormat attrs and the SAME/VALID arithmetic -- core/util/padding.h,
// core/util/tensor_format.h and core/framework/common_shape_fns.cc:19-56.
#ifndef B200TF_CORE_UTIL_PADDING_H_
#define B200TF_CORE_UTIL_PADDING_H_

#include <algorithm>
#include <string>

#include "tensorflow/core/framework/node_def.h"

namespace tensorflow {

enum Padding { VALID = 1, SAME = 2 };
enum TensorFormat { FORMAT_NHWC = 0, FORMAT_NCHW = 1 };

inline Status GetPaddingFromString(const std::string& s, Padding* p) {
  if (s == "SAME") *p = SAME;
  else if (s == "VALID") *p = VALID;
  else return errors::NotFound(s, " is not an allowed padding type");
  return Status::OK();
}
inline bool FormatFromString(const std::string& s, TensorFormat* f) {
  if (s == "NHWC") { *f = FORMAT_NHWC; return true; }
  if (s == "NCHW") { *f = FORMAT_NCHW; return true; }
  return false;
}
// GetNodeAttr specialisation used by kernels: context->GetAttr("padding", &padding_).
inline Status GetNodeAttr(const NodeDef& n, const std::string& name, Padding* v) {
  std::string s;
  TF_RETURN_IF_ERROR(GetNodeAttr(n, name, &s));
  return GetPaddingFromString(s, v);
}

// common_shape_fns.cc:19-47
inline Status GetWindowedOutputSizeVerbose(int64 input_size, int64 filter_size, int64 stride,
                                           Padding padding_type, int64* output_size,
                                           int64* padding_before, int64* padding_after) {
  if (stride <= 0) return errors::InvalidArgument("Stride must be > 0, but got ", stride);
  switch (padding_type) {
    case VALID:
      *output_size = (input_size - filter_size + stride) / stride;
      *padding_before = *padding_after = 0;
      break;
    case SAME: {
      *output_size = (input_size + stride - 1) / stride;
      const int64 padding_needed =
          std::max<int64>(0, (*output_size - 1) * stride + filter_size - input_size);
      // odd total padding: the extra cell goes on the bottom / right
      *padding_before = padding_needed / 2;
      *padding_after = padding_needed - *padding_before;
      break;
    }
  }
  if (*output_size < 0) return errors::InvalidArgument("computed output size would be negative");
  return Status::OK();
}
// common_shape_fns.cc:49-56
inline Status GetWindowedOutputSize(int64 input_size, int64 filter_size, int64 stride,
                                    Padding padding_type, int64* output_size, int64* padding) {
  int64 padding_after_unused;
  return GetWindowedOutputSizeVerbose(input_size, filter_size, stride, padding_type, output_size,
                                      padding, &padding_after_unused);
}

}  // namespace tensorflow
#endif
