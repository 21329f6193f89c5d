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
#include <type_traits>
#include <vector>

// base of whatever an operator parks between its Init and its calls (deleted through this base)
struct AotKernelData {
  AotKernelData() = default;
  virtual ~AotKernelData() = default;
};

class AotExtra {
 public:
  using IntRow = std::vector<int64_t>;
  using FloatRow = std::vector<float>;

  AotExtra() = default;
  virtual ~AotExtra() = default;                 // vtable slots 0, 1
  virtual bool HasAttr(std::string key) = 0;     // slot 2

  // typed read of a registered attribute: routed to the virtual getter of that type (slots 3..10)
  template <typename T>
  T Attr(std::string key) {
    if constexpr (std::is_same<T, bool>::value) return GetAttrBool(key);
    else if constexpr (std::is_same<T, int64_t>::value) return GetAttrInt(key);
    else if constexpr (std::is_same<T, float>::value) return GetAttrFloat(key);
    else if constexpr (std::is_same<T, std::string>::value) return GetAttrStr(key);
    else if constexpr (std::is_same<T, IntRow>::value) return GetAttrIntVec(key);
    else if constexpr (std::is_same<T, FloatRow>::value) return GetAttrFloatVec(key);
    else if constexpr (std::is_same<T, std::vector<IntRow>>::value) return GetAttrInt2DVec(key);
    else if constexpr (std::is_same<T, std::vector<FloatRow>>::value) return GetAttrFloat2DVec(key);
    else return T();
  }

  // per-operator data and workspace sizes travel with the object between Init and the calls
  AotKernelData *KernelData() const { return data_; }
  void SetKernelData(AotKernelData *d) { data_ = d; }
  void DestructKernelData() {
    if (data_) delete data_;
    data_ = nullptr;
  }
  const std::vector<size_t> &WorkSpace() const { return scratch_sizes_; }
  void SetWorkSpace(const std::vector<size_t> &sizes) { scratch_sizes_ = sizes; }

 private:
  // one getter per attribute type -- declaration order is vtable order, it must not change
  virtual bool GetAttrBool(std::string key) = 0;                                // 3
  virtual int64_t GetAttrInt(std::string key) = 0;                              // 4
  virtual float GetAttrFloat(std::string key) = 0;                              // 5
  virtual std::string GetAttrStr(std::string key) = 0;                          // 6
  virtual IntRow GetAttrIntVec(std::string key) = 0;                            // 7
  virtual FloatRow GetAttrFloatVec(std::string key) = 0;                        // 8
  virtual std::vector<IntRow> GetAttrInt2DVec(std::string key) = 0;             // 9
  virtual std::vector<FloatRow> GetAttrFloat2DVec(std::string key) = 0;         // 10

  // data members behind the vptr, in this order: workspace sizes (24 bytes), then the operator's data (8)
  std::vector<size_t> scratch_sizes_;
  AotKernelData *data_ = nullptr;
};

#endif  // MINDRL_B200_AOT_EXTRA_H_
