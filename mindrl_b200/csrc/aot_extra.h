// aot_extra.h -- the attribute interface MindSpore hands to an "aot" custom operator.
//
// The reference mounts native operators only through MindSpore's ops.Custom(..., "aot"): the operator's
// Init function and the operator itself receive an `AotExtra *` through which they read the attributes
// the Python side registered (CustomRegOp.attr) and keep per-operator data between Init and the calls
// (reference: mindspore_rl/utils/mcts/custom_aot_extra.h:24-105 is the copy of that interface the
// reference compiles against; its users are utils/mcts/cpu/cpu_mcts_ops.cc:38-53).
//
// This header restates that BINARY interface so that libmindrl_b200.so can be mounted the same way: what
// matters is the order of the virtual functions (it fixes the vtable slots MindSpore's concrete subclass
// fills) and the two data members behind the vptr.  Slots, in order:
//   [0,1] virtual destructor   [2] HasAttr
//   [3] bool  [4] int64  [5] float  [6] string  [7] vector<int64>  [8] vector<float>
//   [9] vector<vector<int64>>  [10] vector<vector<float>>          (the typed getters, private)
// Members: std::vector<size_t> workspace sizes, AotKernelData * owned by the operator.
//
// Hosts without MindSpore (our tests, other frameworks) get a concrete implementation backed by a
// std::map through the mrl_aot_extra_* C functions of aot_shim.cu.
#ifndef MINDRL_B200_AOT_EXTRA_H_
#define MINDRL_B200_AOT_EXTRA_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

// base of whatever an operator parks between its Init and its calls
class AotKernelData {
 public:
  AotKernelData() = default;
  virtual ~AotKernelData() = default;
};

class AotExtra {
 public:
  AotExtra() = default;
  virtual ~AotExtra() = default;
  virtual bool HasAttr(std::string name) = 0;

  // typed read of a registered attribute; the specialisations below route to the virtual getters
  template <typename T>
  inline T Attr(std::string name) {
    return T();
  }

  void SetWorkSpace(const std::vector<size_t> &sizes) { workspace_ = sizes; }
  const std::vector<size_t> &WorkSpace() const { return workspace_; }
  void SetKernelData(AotKernelData *data) { kernel_data_ = data; }
  AotKernelData *KernelData() const { return kernel_data_; }
  void DestructKernelData() {
    delete kernel_data_;
    kernel_data_ = nullptr;
  }

 private:
  // (declaration order = vtable order: do not reorder)
  virtual bool GetAttrBool(std::string name) = 0;
  virtual int64_t GetAttrInt(std::string name) = 0;
  virtual float GetAttrFloat(std::string name) = 0;
  virtual std::string GetAttrStr(std::string name) = 0;
  virtual std::vector<int64_t> GetAttrIntVec(std::string name) = 0;
  virtual std::vector<float> GetAttrFloatVec(std::string name) = 0;
  virtual std::vector<std::vector<int64_t>> GetAttrInt2DVec(std::string name) = 0;
  virtual std::vector<std::vector<float>> GetAttrFloat2DVec(std::string name) = 0;

  std::vector<size_t> workspace_;
  AotKernelData *kernel_data_{nullptr};
};

template <>
inline bool AotExtra::Attr<bool>(std::string name) {
  return GetAttrBool(name);
}
template <>
inline int64_t AotExtra::Attr<int64_t>(std::string name) {
  return GetAttrInt(name);
}
template <>
inline float AotExtra::Attr<float>(std::string name) {
  return GetAttrFloat(name);
}
template <>
inline std::string AotExtra::Attr<std::string>(std::string name) {
  return GetAttrStr(name);
}
template <>
inline std::vector<int64_t> AotExtra::Attr<std::vector<int64_t>>(std::string name) {
  return GetAttrIntVec(name);
}
template <>
inline std::vector<float> AotExtra::Attr<std::vector<float>>(std::string name) {
  return GetAttrFloatVec(name);
}
template <>
inline std::vector<std::vector<int64_t>> AotExtra::Attr<std::vector<std::vector<int64_t>>>(std::string name) {
  return GetAttrInt2DVec(name);
}
template <>
inline std::vector<std::vector<float>> AotExtra::Attr<std::vector<std::vector<float>>>(std::string name) {
  return GetAttrFloat2DVec(name);
}

#endif  // MINDRL_B200_AOT_EXTRA_H_
