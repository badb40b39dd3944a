// teq.hpp — graph IR of the host side: shapes, tensors, travelers, the evaluator entry
// points and the reverse-mode graph builder.
//
// Mirrors the *interfaces* of the reference's internal/teq so that code written against
// it (eteq, layr, trainers, the pybind module) is a drop-in with only the device
// changing: Shape (internal/teq/shape.hpp:60), iTensor / iDeviceRef / iMetadata
// (itensor.hpp:22-90), iLeaf (ileaf.hpp:26), iFunctor (ifunctor.hpp:28), iTraveler
// (itraveler.hpp:13), iDevice / iEvaluator (ievaluator.hpp:10-24), Evaluator /
// TravEvaluator (evaluator.hpp:11-62), derive / partial_derive (derive.hpp:53-66).
// The implementation is new and compact; it is built with DimT = uint32_t, i.e. the
// reference's own -DSDIM_BYTES=4 switch (shape.hpp:27-33), which BASELINE's batch 65536
// needs.
#ifndef TCR_HOST_TEQ_HPP
#define TCR_HOST_TEQ_HPP

#include <algorithm>
#include <array>
#include <cstdarg>
#include <cstdint>
#include <functional>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

// ---------------------------------------------------------------- global:: error convention
namespace global {

// The reference reports programmer errors with global::fatal[f] and user (shape) errors
// with global::throw_err[f] (internal/global/logs.hpp:35-38,70-81); both surface as C++
// exceptions here (and as Python RuntimeError through pybind11).
struct FatalError : public std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline std::string vformat(const char* fmt, va_list ap) {
  char buf[2048];
  vsnprintf(buf, sizeof(buf), fmt, ap);
  return buf;
}
[[noreturn]] inline void fatal(const std::string& msg) { throw FatalError(msg); }
[[noreturn]] inline void fatalf(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); std::string s = vformat(fmt, ap); va_end(ap); throw FatalError(s);
}
[[noreturn]] inline void throw_err(const std::string& msg) { throw std::runtime_error(msg); }
[[noreturn]] inline void throw_errf(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); std::string s = vformat(fmt, ap); va_end(ap); throw std::runtime_error(s);
}

}  // namespace global

namespace fmts {
template <typename It>
std::string to_string(It begin, It end) {  // cppkg fmts: "[a\b\c]"
  std::stringstream ss;
  ss << "[";
  for (It it = begin; it != end; ++it) {
    if (it != begin) ss << "\\";
    ss << (long long)*it;
  }
  ss << "]";
  return ss.str();
}
}  // namespace fmts

// ---------------------------------------------------------------- marsh:: attributes
namespace teq { struct iTensor; using TensptrT = std::shared_ptr<iTensor>; }

namespace marsh {

// Attribute objects (reference: internal/marsh/objs.hpp:10-432 + internal/teq/objs.hpp).
// Only the value kinds the hot path packs are modelled: integer arrays, pair arrays,
// scalars, strings and tensor references (incl. the "layer" marker).
struct iObject {
  virtual ~iObject() = default;
  virtual iObject* clone() const = 0;
  virtual std::string to_string() const = 0;
  virtual bool equals(const iObject& other) const { return to_string() == other.to_string(); }
};
using ObjptrT = std::unique_ptr<iObject>;

struct IntArray final : public iObject {
  explicit IntArray(std::vector<int64_t> v) : vals_(std::move(v)) {}
  iObject* clone() const override { return new IntArray(vals_); }
  std::string to_string() const override { return fmts::to_string(vals_.begin(), vals_.end()); }
  std::vector<int64_t> vals_;
};

struct PairArray final : public iObject {
  explicit PairArray(std::vector<std::pair<int64_t, int64_t>> v) : vals_(std::move(v)) {}
  iObject* clone() const override { return new PairArray(vals_); }
  std::string to_string() const override {
    std::stringstream ss;
    ss << "[";
    for (size_t i = 0; i < vals_.size(); ++i) ss << (i ? "\\" : "") << "[" << vals_[i].first << ":" << vals_[i].second << "]";
    ss << "]";
    return ss.str();
  }
  std::vector<std::pair<int64_t, int64_t>> vals_;
};

struct Integer final : public iObject {
  explicit Integer(int64_t v) : val_(v) {}
  iObject* clone() const override { return new Integer(val_); }
  std::string to_string() const override { return std::to_string(val_); }
  int64_t val_;
};

struct Float final : public iObject {
  explicit Float(double v) : val_(v) {}
  iObject* clone() const override { return new Float(val_); }
  std::string to_string() const override { return std::to_string(val_); }
  double val_;
};

struct String final : public iObject {
  explicit String(std::string v) : val_(std::move(v)) {}
  iObject* clone() const override { return new String(val_); }
  std::string to_string() const override { return val_; }
  std::string val_;
};

struct iAttributed {
  virtual ~iAttributed() = default;
  virtual std::vector<std::string> ls_attrs() const = 0;
  virtual const iObject* get_attr(const std::string& name) const = 0;
  virtual iObject* get_attr(const std::string& name) = 0;
  virtual void add_attr(const std::string& name, ObjptrT&& attr) = 0;
  virtual void rm_attr(const std::string& name) = 0;
  virtual size_t size() const = 0;
};

struct Maps final : public iAttributed {
  Maps() = default;
  Maps(const Maps& o) { for (auto& kv : o.contents_) contents_.emplace(kv.first, ObjptrT(kv.second->clone())); }
  Maps(Maps&&) = default;
  Maps& operator=(Maps&&) = default;
  std::vector<std::string> ls_attrs() const override {
    std::vector<std::string> out;
    for (auto& kv : contents_) out.push_back(kv.first);
    return out;
  }
  const iObject* get_attr(const std::string& name) const override {
    auto it = contents_.find(name);
    return it == contents_.end() ? nullptr : it->second.get();
  }
  iObject* get_attr(const std::string& name) override {
    auto it = contents_.find(name);
    return it == contents_.end() ? nullptr : it->second.get();
  }
  void add_attr(const std::string& name, ObjptrT&& attr) override { contents_[name] = std::move(attr); }
  void rm_attr(const std::string& name) override { contents_.erase(name); }
  size_t size() const override { return contents_.size(); }
  std::map<std::string, ObjptrT> contents_;  // ordered: deterministic traversal
};

inline void get_attrs(Maps& out, const iAttributed& attributed) {
  for (auto& name : attributed.ls_attrs()) out.add_attr(name, ObjptrT(attributed.get_attr(name)->clone()));
}

}  // namespace marsh

// ---------------------------------------------------------------- teq::
namespace teq {

using RankT = uint8_t;
using DimT = uint32_t;  // reference built with -DSDIM_BYTES=4 (internal/teq/shape.hpp:27-33)
using NElemT = uint64_t;
using RanksT = std::vector<RankT>;
using DimsT = std::vector<DimT>;
const RankT rank_cap = 8;  // shape.hpp:45
using ShapeT = std::array<DimT, rank_cap>;

struct Shape final {
  Shape() { dims_.fill(1); }
  Shape(std::vector<DimT> dims) { vector_assign(dims); }
  Shape(std::initializer_list<DimT> dims) { vector_assign(std::vector<DimT>(dims)); }
  std::string to_string() const { return fmts::to_string(dims_.begin(), dims_.end()); }
  DimT at(RankT idx) const {
    if (rank_cap <= idx) global::throw_errf("cannot access out of bounds index %d", (int)idx);
    return dims_[idx];
  }
  NElemT n_elems() const { return std::accumulate(dims_.begin(), dims_.end(), (NElemT)1, std::multiplies<NElemT>()); }
  bool compatible_before(const Shape& other, RankT idx) const {
    auto it = dims_.begin();
    return std::equal(it, it + std::min(idx, rank_cap), other.begin(), [](DimT a, DimT b) { return a == 0 || b == 0 || a == b; });
  }
  bool compatible_after(const Shape& other, RankT idx) const {
    return idx < rank_cap && std::equal(dims_.begin() + idx, dims_.end(), other.begin() + idx,
                                        [](DimT a, DimT b) { return a == 0 || b == 0 || a == b; });
  }
  bool operator==(const Shape& o) const { return dims_ == o.dims_; }
  ShapeT::iterator begin() { return dims_.begin(); }
  ShapeT::iterator end() { return dims_.end(); }
  ShapeT::const_iterator begin() const { return dims_.begin(); }
  ShapeT::const_iterator end() const { return dims_.end(); }

 private:
  void vector_assign(const std::vector<DimT>& dims) {
    if (std::any_of(dims.begin(), dims.end(), [](DimT d) { return d == 0; }))
      global::throw_errf("cannot create shape with vector containing zero: %s", fmts::to_string(dims.begin(), dims.end()).c_str());
    RankT rank = (RankT)std::min((size_t)rank_cap, dims.size());
    std::copy(dims.begin(), dims.begin() + rank, dims_.begin());
    std::fill(dims_.begin() + rank, dims_.end(), 1);
  }
  ShapeT dims_;
};
using ShapesT = std::vector<Shape>;

/// list of shape dimensions with trailing ones trimmed (shape.hpp:198)
inline DimsT narrow_shape(const Shape& s) {
  auto it = s.begin(), et = s.end();
  while (it != et && *(et - 1) == 1) --et;
  return DimsT(it, et);
}

// -- Once: a value plus a "consumer is done" callback (internal/teq/once.hpp:20-92)
template <typename T>
struct Once final {
  Once(T obj, std::function<void(void)> killsig = {}) : obj_(obj), term_(std::move(killsig)) {}
  Once(const Once&) = delete;
  Once(Once&& o) : obj_(o.obj_), term_(std::move(o.term_)) { o.term_ = nullptr; }
  Once& operator=(const Once&) = delete;
  Once& operator=(Once&& o) {
    if (this != &o) { if (term_) term_(); obj_ = o.obj_; term_ = std::move(o.term_); o.term_ = nullptr; }
    return *this;
  }
  ~Once() { if (term_) term_(); }
  T get() const { return obj_; }
  operator T() const { return obj_; }

 private:
  T obj_;
  std::function<void(void)> term_;
};

struct iLeaf;
struct iFunctor;

struct iTraveler {
  virtual ~iTraveler() = default;
  virtual void visit(iLeaf& leaf) = 0;
  virtual void visit(iFunctor& func) = 0;
};

/// Device reference of a node (itensor.hpp:22-35). `data()` keeps the reference's
/// contract — a HOST-readable pointer, nullptr before the first assign — and is a lazy,
/// version-tracked mirror of the device buffer (a D2H sync point, never used inside a
/// step). `device_data()` is the resident HBM buffer the kernels read and write.
struct iDeviceRef {
  virtual ~iDeviceRef() = default;
  virtual void* data() = 0;
  virtual const void* data() const = 0;
  virtual Once<void*> odata() = 0;              // device pointer + consumer-done tick
  virtual Once<const void*> odata() const = 0;
  virtual void* device_data() = 0;
  virtual const void* device_data() const = 0;
};

struct iMetadata {
  virtual ~iMetadata() = default;
  virtual size_t type_code() const = 0;
  virtual std::string type_label() const = 0;
  virtual size_t type_size() const = 0;
  virtual size_t state_version() const = 0;
};

struct iTensor : public std::enable_shared_from_this<iTensor> {
  virtual ~iTensor() = default;
  iTensor* clone() const { return this->clone_impl(); }
  virtual void accept(iTraveler& visiter) = 0;
  virtual iDeviceRef& device() = 0;
  virtual const iDeviceRef& device() const = 0;
  virtual const iMetadata& get_meta() const = 0;
  virtual Shape shape() const = 0;
  virtual std::string to_string() const = 0;

 protected:
  virtual iTensor* clone_impl() const = 0;
};

using TensptrT = std::shared_ptr<iTensor>;
using TensrefT = std::weak_ptr<iTensor>;
using TensptrsT = std::vector<TensptrT>;
using CTensT = std::vector<const iTensor*>;
using TensSetT = std::unordered_set<iTensor*>;
using TensptrSetT = std::unordered_set<TensptrT>;
template <typename V> using TensMapT = std::unordered_map<iTensor*, V>;
using OwnMapT = TensMapT<TensptrT>;
using RefMapT = TensMapT<TensrefT>;

enum Usage { UNKNOWN_USAGE = 0, IMMUTABLE, VARUSAGE, PLACEHOLDER };  // ileaf.hpp:13-19

struct iLeaf : public iTensor {
  iLeaf* clone() const { return static_cast<iLeaf*>(this->clone_impl()); }
  void accept(iTraveler& visiter) override { visiter.visit(*this); }
  virtual Usage get_usage() const = 0;
};
using LeafptrT = std::shared_ptr<iLeaf>;
using LeafsT = std::vector<iLeaf*>;

struct Opcode final {
  std::string name_;
  size_t code_;
};

struct iFunctor : public iTensor, public marsh::iAttributed {
  iFunctor* clone() const { return static_cast<iFunctor*>(this->clone_impl()); }
  void accept(iTraveler& visiter) override { visiter.visit(*this); }
  virtual Opcode get_opcode() const = 0;
  virtual TensptrsT get_args() const = 0;
  /// args without the vector copy (the reference copies a vector<shared_ptr> per visit,
  /// internal/teq/evaluator.hpp:40; the hot traversal here does not)
  virtual const TensptrsT& args_ref() const = 0;
  virtual void update_child(TensptrT arg, size_t index) = 0;
};
using FuncptrT = std::shared_ptr<iFunctor>;

// tensor-valued attributes (internal/teq/objs.hpp:37-150)
const std::string layer_attr = "layer";

struct TensorRef : public marsh::iObject {
  virtual TensptrT& get_tensor() = 0;
  virtual const TensptrT& get_tensor() const = 0;
  virtual TensorRef* copynreplace(TensptrT) const = 0;
};

struct TensorObj final : public TensorRef {
  explicit TensorObj(TensptrT tens) : tens_(std::move(tens)) {}
  marsh::iObject* clone() const override { return new TensorObj(tens_); }
  std::string to_string() const override { return tens_->to_string(); }
  bool equals(const marsh::iObject& o) const override {
    auto p = dynamic_cast<const TensorObj*>(&o);
    return p && p->tens_ == tens_;
  }
  TensptrT& get_tensor() override { return tens_; }
  const TensptrT& get_tensor() const override { return tens_; }
  TensorRef* copynreplace(TensptrT t) const override { return new TensorObj(t); }
  TensptrT tens_;
};

struct LayerObj final : public TensorRef {
  LayerObj(const std::string& opname, TensptrT input) : opname_(opname), input_(std::move(input)) {
    if (nullptr == input_) global::fatalf("cannot `%s` with null input", opname.c_str());
  }
  marsh::iObject* clone() const override { return new LayerObj(opname_, input_); }
  std::string to_string() const override { return opname_; }
  bool equals(const marsh::iObject& o) const override {
    auto p = dynamic_cast<const LayerObj*>(&o);
    return p && p->opname_ == opname_ && p->input_ == input_;
  }
  TensptrT& get_tensor() override { return input_; }
  const TensptrT& get_tensor() const override { return input_; }
  TensorRef* copynreplace(TensptrT t) const override { return new LayerObj(opname_, t); }
  std::string get_opname() const { return opname_; }
  std::string opname_;
  TensptrT input_;
};

template <typename TS>
void multi_visit(iTraveler& visiter, const TS& tensors) {
  for (auto& t : tensors) t->accept(visiter);
}

/// visits every node once (traveler.hpp:152-190)
struct iOnceTraveler : public iTraveler {
  void visit(iLeaf& leaf) override;
  void visit(iFunctor& func) override;
  virtual void clear() { visited_.clear(); }
  TensSetT visited_;

 protected:
  virtual void visit_leaf(iLeaf& leaf) = 0;
  virtual void visit_func(iFunctor& func) = 0;
};

/// attribute tensors of a functor (teq::FindTensAttr)
TensptrsT attr_tensors(const iFunctor& func);

/// longest distance from the leaves (GraphStat::graphsize_[x].upper_, traveler.hpp:51-110)
struct GraphStat final : public iTraveler {
  void visit(iLeaf& leaf) override { height_.emplace((iTensor*)&leaf, 0); }
  void visit(iFunctor& func) override;
  TensMapT<size_t> height_;
};

/// post-order index (traveler.hpp:112-148)
struct GraphIndex final : public iTraveler {
  void visit(iLeaf& leaf) override { indices_.emplace((iTensor*)&leaf, indices_.size()); }
  void visit(iFunctor& func) override;
  TensMapT<size_t> indices_;
};

struct PathDirection {
  std::vector<size_t> args_;
  std::vector<std::string> attrs_;
};

/// marks the functors that lead to a target and through which children (traveler.hpp:200-300)
struct PathFinder final : public iOnceTraveler {
  explicit PathFinder(TensSetT targets, bool follow_attrs = true) : targets_(std::move(targets)), follow_attrs_(follow_attrs) {}
  TensMapT<PathDirection> roadmap_;

 private:
  void visit_leaf(iLeaf&) override {}
  void visit_func(iFunctor& func) override;
  TensSetT targets_;
  bool follow_attrs_;
};

/// deep copy of a subgraph except `ignores` (traveler.hpp:378-432)
struct Copier final : public iOnceTraveler {
  explicit Copier(TensSetT ignores = {}) : ignores_(std::move(ignores)) {}
  OwnMapT clones_;
  TensSetT ignores_;

 private:
  void visit_leaf(iLeaf& leaf) override;
  void visit_func(iFunctor& func) override;
};

OwnMapT track_ownptrs(const TensptrsT& roots);

// ---- evaluation entry points (ievaluator.hpp:10-24, evaluator.hpp:11-62)
struct iDevice {
  virtual ~iDevice() = default;
  virtual void calc(iTensor& tens, size_t cache_ttl) = 0;
};

struct iEvaluator {
  virtual ~iEvaluator() = default;
  virtual void evaluate(iDevice& device, const TensSetT& targets, const TensSetT& ignored = {}) = 0;
};
using iEvalptrT = std::shared_ptr<iEvaluator>;

/// post-order DFS from the targets, once per node, honouring `ignored`
struct TravEvaluator final : public iOnceTraveler {
  TravEvaluator(iDevice& device, const TensSetT& targets, const TensSetT& ignored);
  TensSetT ignored_;

 private:
  void visit_leaf(iLeaf&) override {}
  void visit_func(iFunctor& func) override;
  iDevice* device_;
  TensSetT targets_;
};

struct Evaluator final : public iEvaluator {
  void evaluate(iDevice& device, const TensSetT& targets, const TensSetT& ignored = {}) override {
    TravEvaluator eval(device, targets, ignored);
    for (auto t : targets) t->accept(eval);
  }
};

/// context-pluggable evaluator slot (internal/teq/src/evaluator.cpp:11-30)
void set_eval(iEvalptrT eval);
iEvaluator& get_eval();

// ---- reverse-mode graph builder (derive.hpp:19-66, src/derive.cpp:13-161)
struct iDerivativeFuncs {
  virtual ~iDerivativeFuncs() = default;
  virtual TensptrT lderive(FuncptrT op, TensptrT supgrad, size_t i) const = 0;
  virtual TensptrT get_const_one(iTensor& reference) const = 0;
  virtual TensptrT get_const_zero(iTensor& reference) const = 0;
  virtual TensptrT add(TensptrsT elems) const = 0;
};
using GradMapT = TensMapT<TensptrsT>;

TensptrsT derive(TensptrT root, const TensptrsT& targets, const iDerivativeFuncs& funcs);
void partial_derive(GradMapT& grads, const TensptrSetT& parents, const TensSetT& targets, const iDerivativeFuncs& funcs);

}  // namespace teq

#endif  // TCR_HOST_TEQ_HPP
