// Tensor / TensorBuffer -- subset of the reference's core/framework/tensor.h:43-500.
// Dense row-major storage (framework/tensor_types.h:25-28), ref-counted buffers from an
// Allocator (tensor.h:480-500).  Typed accessors return raw pointers instead of Eigen maps:
// the B200 kernels take pointers through the C ABI.
#ifndef B200TF_CORE_FRAMEWORK_TENSOR_H_
#define B200TF_CORE_FRAMEWORK_TENSOR_H_

#include <atomic>
#include <cstring>
#include <string>

#include "tensorflow/core/framework/allocator.h"
#include "tensorflow/core/framework/tensor_shape.h"
#include "tensorflow/core/framework/types.h"

namespace tensorflow {

class TensorBuffer {
 public:
  TensorBuffer(Allocator* a, size_t bytes)
      : alloc_(a), data_(a->AllocateRaw(Allocator::kAllocatorAlignment, bytes)), size_(bytes) {}
  // Wraps memory owned elsewhere (e.g. TF_NewTensor with a deallocator handled by the caller).
  TensorBuffer(void* data, size_t bytes) : alloc_(nullptr), data_(data), size_(bytes) {}
  // A window [offset, offset + bytes) of `root` (tensor.cc SubBuffer): keeps the root alive and
  // never frees memory itself.  Used for slices of a gradient arena (direct_session.cc).
  TensorBuffer(TensorBuffer* root, size_t offset, size_t bytes)
      : alloc_(nullptr), data_(static_cast<char*>(root->data()) + offset), size_(bytes),
        root_(root->root_buffer()) {
    root_->Ref();
  }
  void* data() const { return data_; }
  size_t size() const { return size_; }
  // The allocator the memory came from (a window reports its root's).
  Allocator* allocator() const { return root_ ? root_->alloc_ : alloc_; }
  TensorBuffer* root_buffer() { return root_ ? root_ : this; }
  void Ref() { ref_.fetch_add(1, std::memory_order_relaxed); }
  bool Unref() {
    if (ref_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      delete this;
      return true;
    }
    return false;
  }
  bool RefCountIsOne() const { return ref_.load(std::memory_order_acquire) == 1; }

 private:
  ~TensorBuffer() {
    if (root_)
      root_->Unref();
    else if (alloc_ && data_)
      alloc_->DeallocateRaw(data_);
  }
  Allocator* alloc_;
  void* data_;
  size_t size_;
  TensorBuffer* root_ = nullptr;
  std::atomic<int> ref_{1};
};

class Tensor {
 public:
  Tensor() : dtype_(DT_FLOAT), shape_({0}), buf_(nullptr) {}
  Tensor(DataType type, const TensorShape& shape) : Tensor(cpu_allocator(), type, shape) {}
  Tensor(Allocator* a, DataType type, const TensorShape& shape)
      : dtype_(type), shape_(shape), buf_(nullptr) {
    const size_t bytes = static_cast<size_t>(shape.num_elements()) * DataTypeSize(type);
    if (bytes > 0) {
      buf_ = new TensorBuffer(a, bytes);
      if (buf_->data() == nullptr) {  // allocation failure: IsInitialized() turns false
        buf_->Unref();
        buf_ = nullptr;
      }
    }
  }
  Tensor(DataType type, const TensorShape& shape, TensorBuffer* buf)
      : dtype_(type), shape_(shape), buf_(buf) {
    if (buf_) buf_->Ref();
  }
  Tensor(const Tensor& o) : dtype_(o.dtype_), shape_(o.shape_), buf_(o.buf_) {
    if (buf_) buf_->Ref();
  }
  Tensor(Tensor&& o) : dtype_(o.dtype_), shape_(std::move(o.shape_)), buf_(o.buf_) {
    o.buf_ = nullptr;
  }
  ~Tensor() {
    if (buf_) buf_->Unref();
  }
  Tensor& operator=(const Tensor& o) {
    if (this != &o) {
      if (o.buf_) o.buf_->Ref();
      if (buf_) buf_->Unref();
      dtype_ = o.dtype_;
      shape_ = o.shape_;
      buf_ = o.buf_;
    }
    return *this;
  }
  Tensor& operator=(Tensor&& o) {
    if (this != &o) {
      if (buf_) buf_->Unref();
      dtype_ = o.dtype_;
      shape_ = std::move(o.shape_);
      buf_ = o.buf_;
      o.buf_ = nullptr;
    }
    return *this;
  }

  DataType dtype() const { return dtype_; }
  const TensorShape& shape() const { return shape_; }
  int dims() const { return shape_.dims(); }
  int64 dim_size(int d) const { return shape_.dim_size(d); }
  int64 NumElements() const { return shape_.num_elements(); }
  bool IsSameSize(const Tensor& b) const { return shape_.IsSameSize(b.shape_); }
  bool SharesBufferWith(const Tensor& b) const { return buf_ != nullptr && buf_ == b.buf_; }
  bool IsInitialized() const { return buf_ != nullptr || NumElements() == 0; }
  size_t TotalBytes() const { return static_cast<size_t>(NumElements()) * DataTypeSize(dtype_); }
  size_t AllocatedBytes() const { return buf_ ? buf_->size() : 0; }

  // Same buffer, new shape with the same element count (tensor.h CopyFrom).
  bool CopyFrom(const Tensor& other, const TensorShape& shape) {
    if (other.NumElements() != shape.num_elements()) return false;
    *this = other;
    shape_ = shape;
    return true;
  }

  void* raw_data() const { return buf_ ? buf_->data() : nullptr; }
  template <typename T> T* data() const { return static_cast<T*>(raw_data()); }
  // flat<T>() returns the base pointer; use NumElements() for the extent.
  template <typename T> T* flat() const { return data<T>(); }
  template <typename T> const T& scalar() const { return *data<T>(); }
  TensorBuffer* buffer() const { return buf_; }
  std::string DebugString() const {
    return strings::StrCat("Tensor<type: ", DataTypeString(dtype_), " shape: ",
                           shape_.DebugString(), ">");
  }

 private:
  DataType dtype_;
  TensorShape shape_;
  TensorBuffer* buf_;
};

}  // namespace tensorflow
#endif
