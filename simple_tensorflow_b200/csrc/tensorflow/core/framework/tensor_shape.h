// TensorShape / TensorShapeUtils -- subset of the reference's core/framework/tensor_shape.h.
#ifndef B200TF_CORE_FRAMEWORK_TENSOR_SHAPE_H_
#define B200TF_CORE_FRAMEWORK_TENSOR_SHAPE_H_

#include <initializer_list>
#include <string>
#include <vector>

#include "tensorflow/core/framework/types.h"

namespace tensorflow {

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> dims) : dims_(dims) {}
  explicit TensorShape(const std::vector<int64>& dims) : dims_(dims) {}
  int dims() const { return static_cast<int>(dims_.size()); }
  int64 dim_size(int d) const { return dims_[d]; }
  const std::vector<int64>& dim_sizes() const { return dims_; }
  void AddDim(int64 size) { dims_.push_back(size); }
  void set_dim(int d, int64 size) { dims_[d] = size; }
  void Clear() { dims_.clear(); }
  int64 num_elements() const {
    int64 n = 1;
    for (int64 d : dims_) n *= d;
    return n;
  }
  // TensorShape::IsValid (tensor_shape.cc): every dimension >= 0 and the element count (and,
  // with `element_bytes`, the byte size) representable.  Untrusted shapes (GraphDef import,
  // TF_NewTensor / TF_AllocateTensor) must pass this before num_elements() is trusted.
  static constexpr int64 kMaxElements = int64(1) << 40;  // 1 Ti elements: far above 180 GB tensors
  bool IsValid(size_t element_bytes = 1) const {
    unsigned long long n = 1;
    bool empty = false;
    for (int64 d : dims_) {
      if (d < 0) return false;
      if (d == 0) empty = true;
    }
    if (empty) return true;  // zero elements: nothing can overflow
    for (int64 d : dims_) {
      if (n > static_cast<unsigned long long>(kMaxElements) / static_cast<unsigned long long>(d))
        return false;
      n *= static_cast<unsigned long long>(d);
    }
    if (element_bytes > 1 && n > (~0ull >> 1) / element_bytes) return false;
    return true;
  }
  bool IsSameSize(const TensorShape& b) const { return dims_ == b.dims_; }
  bool operator==(const TensorShape& b) const { return dims_ == b.dims_; }
  bool operator!=(const TensorShape& b) const { return dims_ != b.dims_; }
  std::string DebugString() const {
    std::string s = "[";
    for (size_t i = 0; i < dims_.size(); ++i) {
      if (i) s += ",";
      s += std::to_string(dims_[i]);
    }
    return s + "]";
  }

 private:
  std::vector<int64> dims_;
};

struct TensorShapeUtils {
  static bool IsScalar(const TensorShape& s) { return s.dims() == 0; }
  static bool IsVector(const TensorShape& s) { return s.dims() == 1; }
  static bool IsVectorOrHigher(const TensorShape& s) { return s.dims() >= 1; }
  static bool IsMatrix(const TensorShape& s) { return s.dims() == 2; }
  static bool IsMatrixOrHigher(const TensorShape& s) { return s.dims() >= 2; }
  static Status MakeShape(const int32* dims, int64 n, TensorShape* out) {
    out->Clear();
    for (int64 i = 0; i < n; ++i) {
      if (dims[i] < 0)
        return errors::InvalidArgument("Dimension ", dims[i], " must be >= 0");
      out->AddDim(dims[i]);
    }
    return Status::OK();
  }
};

}  // namespace tensorflow
#endif
