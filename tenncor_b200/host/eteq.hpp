// eteq.hpp — data holders, device, functor / variable / constant nodes and the gradient
// rules of the host side.
//
// Interfaces mirror the reference so the layers above are a drop-in:
//   eigen::iEigen + holders   internal/eigen/device.hpp:12-21,201-544
//   eigen::iRuntimeMemory     internal/eigen/memory.hpp:14-139 (Expirable TTL semantics kept)
//   eigen::Observable         internal/eigen/observable.hpp:18-144
//   cuda::Device              internal/eigen/device.hpp:546-576 (eigen::Device::calc)
//   cuda::typed_exec          generated egen::typed_exec<T> (tools/egen/plugins/opcodes.py:36-46)
//   eteq::Functor/Variable/Constant, make_functor, DerivativeFuncs
//                             tenncor/eteq/{functor,variable,constant,make,backprop}.hpp
// What changed is the device: holders own HBM buffers from the library arena and their
// assign() enqueues hand-written sm_100a kernels through the C-ABI (include/tcr_b200.h).
// The element type is a run-time dtype code instead of a template parameter.
#ifndef TCR_HOST_ETEQ_HPP
#define TCR_HOST_ETEQ_HPP

#include "egen.hpp"
#include "tcr_b200.h"

namespace eigen {

// ---- runtime memory (device arena) with the reference's TTL contract
struct iRuntimeMemory {
  virtual ~iRuntimeMemory() = default;
  virtual void* allocate(size_t size) = 0;
  virtual void deallocate(void* ptr, size_t size) = 0;
};
using RTMemptrT = std::shared_ptr<iRuntimeMemory>;

/// tcr_alloc / tcr_free: stream-ordered, size-bucketed HBM arena
struct DeviceRuntimeMemory final : public iRuntimeMemory {
  void* allocate(size_t size) override;
  void deallocate(void* ptr, size_t size) override;
};

void set_runtime(RTMemptrT mem);
RTMemptrT get_runtime();

/// buffer whose lifetime is counted in consumer reads (memory.hpp:54-139)
struct Expirable final {
  ~Expirable() { expire(); }
  void expire();
  void tick() { if (!pinned_ && ttl_ > 0 && 0 == (--ttl_)) expire(); }
  /// keep the buffer for the holder's lifetime once it exists (the reference's Functor::cache_init swaps in a permanent holder,
  /// tenncor/eteq/functor.hpp:231-243, internal/eigen/device.hpp:41-101): consumer reads no longer count it down
  void pin() { pinned_ = true; }
  bool is_expired() const { return 0 == ttl_; }
  size_t get_ttl() const { return ttl_; }
  void* get() { return ptr_; }
  size_t bytes() const { return size_; }
  void borrow(RTMemptrT& memory, size_t bytes, size_t ttl);
  void extend_life(size_t ttl);

 private:
  size_t ttl_ = 0;
  bool pinned_ = false;
  void* ptr_ = nullptr;
  size_t size_ = 0;
  RTMemptrT allocator_ = nullptr;
};

struct iEigen : public teq::iDeviceRef {
  /// compute now; the result must survive `ttl` consumer reads
  virtual void assign(size_t ttl, RTMemptrT& runtime) = 0;
  virtual bool valid_for(size_t /*desired_ttl*/) const { return true; }
  virtual void extend_life(size_t /*ttl*/) {}
  /// the device buffer was written by a kernel: the host mirror is stale
  virtual void mark_device_dirty() {}
};
using EigenptrT = std::shared_ptr<iEigen>;

struct iMutableLeaf : public teq::iLeaf {
  virtual void upversion(size_t version) = 0;
};

struct EMetadata final : public teq::iMetadata {
  explicit EMetadata(egen::_GENERATED_DTYPE dtype, size_t version = 0) : dtype_(dtype), version_(version) {}
  size_t type_code() const override { return dtype_; }
  std::string type_label() const override { return egen::name_type(dtype_); }
  size_t type_size() const override { return egen::type_size(dtype_); }
  size_t state_version() const override { return version_; }
  egen::_GENERATED_DTYPE dtype_;
  size_t version_;
};

/// functor base: subscriber set (= consumer count for the TTL) + attributes
struct Observable : public teq::iFunctor {
  explicit Observable(const teq::TensptrsT& args);
  Observable(const teq::TensptrsT& args, marsh::Maps&& attrs);
  Observable(const Observable& other) : attrs_(other.attrs_) {}
  virtual ~Observable() = default;
  void subscribe(Observable* sub) { subs_.emplace(sub); }
  void unsubscribe(Observable* sub) { subs_.erase(sub); }
  size_t nsubs() const { return subs_.size(); }
  virtual bool has_data() const = 0;
  virtual void uninitialize() = 0;
  virtual bool initialize() = 0;
  virtual void must_initialize() = 0;
  virtual bool prop_version(size_t max_version) = 0;
  // marsh::iAttributed
  std::vector<std::string> ls_attrs() const override { return attrs_.ls_attrs(); }
  const marsh::iObject* get_attr(const std::string& name) const override { return attrs_.get_attr(name); }
  marsh::iObject* get_attr(const std::string& name) override { return attrs_.get_attr(name); }
  void add_attr(const std::string& name, marsh::ObjptrT&& attr) override { attrs_.add_attr(name, std::move(attr)); }
  void rm_attr(const std::string& name) override { attrs_.rm_attr(name); }
  size_t size() const override { return attrs_.size(); }

 protected:
  std::unordered_set<Observable*> subs_;
  marsh::Maps attrs_;
};

}  // namespace eigen

namespace cuda {

/// check a C-ABI status; failure -> global::fatal(tcr_last_error()) (no status codes above the boundary)
void check(int rc, const char* what);
/// tcr_init on first use; fatal when there is no device (no CPU fallback)
void ensure_device();
void sync();

int gemm_precision();          // TCR_GEMM_* used for FLOAT contractions
void set_gemm_precision(int p);

/// lazily synchronised host copy of a device buffer (keeps `iDeviceRef::data()` host-readable)
struct HostMirror {
  void* sync_from(const void* dev, size_t bytes);
  void invalidate() { valid_ = false; }
  std::vector<char> host_;
  bool valid_ = false;
};

/// leaf storage (replaces eigen::SrcRef<T>, device.hpp:201-247): owns an HBM buffer
struct DevSrc final : public eigen::iEigen {
  DevSrc(const void* host_data, egen::_GENERATED_DTYPE dtype, teq::Shape shape, bool keep_host);
  ~DevSrc();
  void* data() override;
  const void* data() const override { return const_cast<DevSrc*>(this)->data(); }
  teq::Once<void*> odata() override { return teq::Once<void*>(device_data()); }
  teq::Once<const void*> odata() const override { return teq::Once<const void*>(device_data()); }
  /// uploads the staged host copy on first use (so graphs can be built without a GPU)
  void* device_data() override;
  const void* device_data() const override { return const_cast<DevSrc*>(this)->device_data(); }
  void assign(size_t, eigen::RTMemptrT&) override {}
  void mark_device_dirty() override { mirror_.invalidate(); }
  bool resident() const { return dev_ != nullptr; }
  /// host -> HBM (async on the library stream)
  void assign_host(const void* host_data);
  /// HBM -> HBM
  void assign_device(const void* dev_data);
  /// pinned host -> device staging buffer on the copy stream (overlaps the running step); `commit_prefetch`
  /// makes the staged batch this leaf's data (stream-ordered HBM -> HBM copy)
  void prefetch_host(const void* host_data);
  void commit_prefetch();
  size_t bytes() const { return bytes_; }

 private:
  void* dev_ = nullptr;
  void* staging_ = nullptr;
  bool staged_ = false;
  size_t bytes_;
  bool keep_host_;  // constants keep their host copy (is_scalar / const folding read it)
  mutable HostMirror mirror_;
};

/// temporary output of an op (replaces TensOp / MatOp, device.hpp:295-381)
using LaunchF = std::function<void(void* out, const std::vector<const void*>& in)>;

struct DevOp final : public eigen::iEigen {
  DevOp(size_t out_bytes, teq::CTensT args, LaunchF launch) : bytes_(out_bytes), args_(std::move(args)), launch_(std::move(launch)) {}
  void* data() override { return data_.get() ? mirror_.sync_from(data_.get(), bytes_) : nullptr; }
  const void* data() const override { return const_cast<DevOp*>(this)->data(); }
  teq::Once<void*> odata() override { return teq::Once<void*>(data_.get(), [this] { data_.tick(); }); }
  teq::Once<const void*> odata() const override { return teq::Once<const void*>(data_.get(), [this] { data_.tick(); }); }
  void* device_data() override { return data_.get(); }
  const void* device_data() const override { return data_.get(); }
  void assign(size_t ttl, eigen::RTMemptrT& runtime) override;
  bool valid_for(size_t desired_ttl) const override { return desired_ttl <= data_.get_ttl(); }
  void extend_life(size_t ttl) override { data_.extend_life(ttl); }
  void mark_device_dirty() override { mirror_.invalidate(); }
  /// bind an externally produced result (used by the plan executor for target nodes)
  void* ensure_buffer(size_t ttl, eigen::RTMemptrT& runtime);
  /// Functor::cache_init: the result outlives its consumers' reads (evaluations that name this node in their `ignored` set read it as data)
  void pin() { data_.pin(); }
  /// run this op's kernel(s) on explicit buffers (plan executor: buffers come from the plan)
  void launch_with(void* out, const std::vector<const void*>& in) const { launch_(out, in); }
  size_t out_bytes() const { return bytes_; }
  /// test support: swap the kernel launcher (the role of the op lambda handed to TensOp in internal/eigen/test/test_device.cpp)
  void set_launch(LaunchF launch) { launch_ = std::move(launch); }
  /// MATMUL / CONTRACT lowered to one tcr_gemm: the descriptor, so a launcher may add an epilogue
  void set_gemm(const tcr_gemm_desc& d) { gemm_ = std::make_shared<tcr_gemm_desc>(d); }
  const tcr_gemm_desc* gemm() const { return gemm_.get(); }

 private:
  size_t bytes_;
  teq::CTensT args_;
  LaunchF launch_;
  std::shared_ptr<tcr_gemm_desc> gemm_;
  mutable eigen::Expirable data_;
  mutable HostMirror mirror_;
};

/// alias of another node's buffer, optionally at an element offset
/// (replaces TensRef / UnsafeTensRef, device.hpp:383-505)
struct DevRef final : public eigen::iEigen {
  DevRef(teq::iTensor& ref, size_t byte_offset = 0) : ref_(&ref), offset_(byte_offset) {}
  void* data() override { auto p = (char*)ref_->device().data(); return p ? p + offset_ : nullptr; }
  const void* data() const override { return const_cast<DevRef*>(this)->data(); }
  teq::Once<void*> odata() override { return teq::Once<void*>(device_data(), [this] { tick(); }); }
  teq::Once<const void*> odata() const override { return teq::Once<const void*>(device_data(), [this] { tick(); }); }
  void* device_data() override { auto p = (char*)ref_->device().device_data(); return p ? p + offset_ : nullptr; }
  const void* device_data() const override { return const_cast<DevRef*>(this)->device_data(); }
  void assign(size_t ttl, eigen::RTMemptrT&) override { extend_life(ttl); }
  bool valid_for(size_t desired_ttl) const override { return desired_ttl <= ref_ttl_; }
  void extend_life(size_t ttl) override { if (ref_ttl_ < ttl) ref_ttl_ = ttl; }
  teq::iTensor* referent() const { return ref_; }
  size_t byte_offset() const { return offset_; }

 private:
  void tick() const {  // the ref's own ttl forwards one tick to the referent when exhausted (device.hpp:389-413)
    if (ref_ttl_ > 0 && 0 == (--ref_ttl_)) ref_->device().odata();
  }
  teq::iTensor* ref_;
  size_t offset_;
  mutable size_t ref_ttl_ = 0;
};

/// in-place update of variable storage (replaces TensAssign<T>, device.hpp:507-544)
struct DevAssign final : public eigen::iEigen {
  DevAssign(egen::_GENERATED_OPCODE op, teq::iTensor& target, const teq::iTensor& arg) : op_(op), ref_(&target), arg_(&arg) {}
  void* data() override { return ref_->device().data(); }
  const void* data() const override { return ref_->device().data(); }
  teq::Once<void*> odata() override { return teq::Once<void*>(device_data(), [this] { tick(); }); }
  teq::Once<const void*> odata() const override { return teq::Once<const void*>(device_data(), [this] { tick(); }); }
  void* device_data() override { return ref_->device().device_data(); }
  const void* device_data() const override { return ref_->device().device_data(); }
  void assign(size_t ttl, eigen::RTMemptrT&) override;
  bool valid_for(size_t desired_ttl) const override { return desired_ttl <= ref_ttl_; }
  void extend_life(size_t ttl) override { if (ref_ttl_ < ttl) ref_ttl_ = ttl; }
  egen::_GENERATED_OPCODE opcode() const { return op_; }
  teq::iTensor* target() const { return ref_; }
  const teq::iTensor* source() const { return arg_; }

 private:
  void tick() const {
    if (ref_ttl_ > 0 && 0 == (--ref_ttl_)) ref_->device().odata();
  }
  egen::_GENERATED_OPCODE op_;
  teq::iTensor* ref_;
  const teq::iTensor* arg_;
  mutable size_t ref_ttl_ = 0;
};

/// the generated switch, re-pointed at the device:
/// out = holder whose assign() launches the kernel(s) of `opcode`
void typed_exec(egen::_GENERATED_OPCODE opcode, egen::_GENERATED_DTYPE dtype, eigen::EigenptrT& out,
                teq::Shape outshape, const teq::TensptrsT& in, const marsh::iAttributed& attrib);

/// eigen::Device (device.hpp:546-576): TTL = nsubs + is_target; recompute iff the version
/// propagates or there is no data, else extend life
struct Device final : public teq::iDevice {
  explicit Device(size_t max_version = std::numeric_limits<size_t>::max()) : max_version_(max_version), memory_(eigen::get_runtime()) {}
  Device(eigen::RTMemptrT memory, size_t max_version = std::numeric_limits<size_t>::max()) : max_version_(max_version), memory_(std::move(memory)) {}
  void calc(teq::iTensor& tens, size_t cache_ttl) override;
  eigen::RTMemptrT& memory() { return memory_; }
  size_t max_version_;

 private:
  eigen::RTMemptrT memory_;
};

/// describes how CONTRACT maps onto the strided GEMM of the C-ABI (false: needs tcr_contract)
bool contract_as_gemm(const teq::Shape& ashape, const teq::Shape& bshape, const eigen::PairVecT<teq::RankT>& pairs, tcr_gemm_desc& d);
/// MATMUL (operator.hpp:1108-1139) as a (batched) GEMM
void matmul_as_gemm(const teq::Shape& ashape, const teq::Shape& bshape, tcr_gemm_desc& d);

}  // namespace cuda

namespace eteq {

/// highest version handed out so far (stands in for the registry scan of
/// tenncor/eteq/variable.hpp:18-26)
size_t get_lastvers();
void note_version(size_t v);

/// RNG state of RAND_UNIF: Philox key + running counter (replaces global::Randomizer,
/// internal/global/random.hpp:78-146)
void seed(uint64_t s);
uint64_t rng_seed();
uint64_t rng_advance(uint64_t n);  // returns the offset to use, then advances by n
void rng_flush();                  // push a pending seed() to the device-resident generator

struct Variable final : public eigen::iMutableLeaf {
  static Variable* get(const void* host_data, egen::_GENERATED_DTYPE dtype, teq::Shape shape, std::string label = "",
                       teq::Usage usage = teq::VARUSAGE);
  Variable* clone() const { return static_cast<Variable*>(clone_impl()); }
  /// host pointer of `dtype` elements -> device storage; bumps the version past every other
  void assign(const void* input, egen::_GENERATED_DTYPE dtype, teq::Shape shape);
  /// device pointer of this variable's own dtype (stays on HBM)
  void assign_device(const void* dev_input);
  /// double-buffered input: `prefetch` starts the host -> HBM copy of the NEXT batch (pinned memory, this variable's
  /// dtype) on the copy stream while the current step computes; `commit` makes it the variable's data and bumps the
  /// version exactly like `assign`. Extension of Variable<T>::assign (tenncor/eteq/variable.hpp:55-90).
  void prefetch(const void* input, egen::_GENERATED_DTYPE dtype, teq::Shape shape);
  void commit();
  teq::Shape shape() const override { return shape_; }
  teq::iDeviceRef& device() override { return *ref_; }
  const teq::iDeviceRef& device() const override { return *ref_; }
  const teq::iMetadata& get_meta() const override { return meta_; }
  std::string to_string() const override { return label_; }
  teq::Usage get_usage() const override { return usage_; }
  void upversion(size_t version) override;

 private:
  Variable(const void* host_data, egen::_GENERATED_DTYPE dtype, teq::Shape shape, std::string label, teq::Usage usage);
  Variable(const Variable& other);
  teq::iTensor* clone_impl() const override { return new Variable(*this); }
  std::shared_ptr<cuda::DevSrc> ref_;
  teq::Shape shape_;
  std::string label_;
  eigen::EMetadata meta_;
  teq::Usage usage_;
};
using VarptrT = std::shared_ptr<Variable>;
using VarptrsT = std::vector<VarptrT>;

struct Constant final : public teq::iLeaf {
  static Constant* get(const void* host_data, egen::_GENERATED_DTYPE dtype, teq::Shape shape);
  teq::Shape shape() const override { return shape_; }
  teq::iDeviceRef& device() override { return *ref_; }
  const teq::iDeviceRef& device() const override { return *ref_; }
  const teq::iMetadata& get_meta() const override { return meta_; }
  std::string to_string() const override;
  teq::Usage get_usage() const override { return teq::IMMUTABLE; }
  /// all elements equal (constant.hpp `is_scalar`): the planner folds such operands into immediates
  bool is_scalar() const { return scalar_; }
  double scalar_value() const { return scalar_value_; }

 private:
  Constant(const void* host_data, egen::_GENERATED_DTYPE dtype, teq::Shape shape);
  Constant(const Constant& other) = default;
  teq::iTensor* clone_impl() const override { return new Constant(*this); }
  std::shared_ptr<cuda::DevSrc> ref_;
  teq::Shape shape_;
  eigen::EMetadata meta_;
  bool scalar_ = false;
  double scalar_value_ = 0;
};

struct Functor final : public eigen::Observable {
  static Functor* get(egen::_GENERATED_OPCODE opcode, egen::_GENERATED_DTYPE dtype, teq::TensptrsT children, marsh::Maps&& attrs);
  ~Functor();
  Functor* clone() const { return static_cast<Functor*>(clone_impl()); }
  teq::Shape shape() const override { return shape_; }
  std::string to_string() const override { return opcode_.name_; }
  teq::Opcode get_opcode() const override { return opcode_; }
  teq::TensptrsT get_args() const override { return args_; }
  const teq::TensptrsT& args_ref() const override { return args_; }
  void update_child(teq::TensptrT arg, size_t index) override;
  teq::iDeviceRef& device() override;
  const teq::iDeviceRef& device() const override;
  const teq::iMetadata& get_meta() const override { return meta_; }
  bool has_data() const override { return nullptr != ref_; }
  void uninitialize() override;
  bool initialize() override;
  void must_initialize() override;
  bool prop_version(size_t max_version) override;
  eigen::EigenptrT& holder() { return ref_; }

 private:
  Functor(egen::_GENERATED_OPCODE opcode, egen::_GENERATED_DTYPE dtype, teq::Shape shape, teq::TensptrsT args, marsh::Maps&& attrs);
  Functor(const Functor& other);
  teq::iTensor* clone_impl() const override { return new Functor(*this); }
  eigen::EigenptrT ref_ = nullptr;
  teq::Opcode opcode_;
  teq::Shape shape_;
  teq::TensptrsT args_;
  eigen::EMetadata meta_;
};

// ---- node factories (tenncor/eteq/make.hpp:14-260)
teq::TensptrT make_funcattr(egen::_GENERATED_OPCODE opcode, teq::TensptrsT children, marsh::Maps& attrs);
teq::TensptrT make_tfuncattr(egen::_GENERATED_DTYPE dtype, egen::_GENERATED_OPCODE opcode, teq::TensptrsT children, marsh::Maps& attrs);

template <typename... ARGS>
teq::TensptrT make_functor(egen::_GENERATED_OPCODE opcode, const teq::TensptrsT& children, ARGS... vargs) {
  marsh::Maps attrs;
  eigen::pack_attr(attrs, vargs...);
  return make_funcattr(opcode, children, attrs);
}

VarptrT make_variable_scalar(double scalar, teq::Shape shape, std::string label = "", egen::_GENERATED_DTYPE dtype = egen::default_dtype);
VarptrT make_variable(const void* data, egen::_GENERATED_DTYPE dtype, teq::Shape shape, std::string label = "");
teq::TensptrT make_constant_tensor(const void* data, egen::_GENERATED_DTYPE dtype, teq::Shape shape);
teq::TensptrT make_constant_scalar(double scalar, teq::Shape shape, egen::_GENERATED_DTYPE dtype);
/// scalar constant EXTENDed to `like` (how every scalar operand reaches the graph, make.hpp:222-254)
teq::TensptrT make_constant_like(double scalar, teq::TensptrT like);

/// per-opcode gradient rules (tenncor/eteq/backprop.hpp:58-576)
struct DerivativeFuncs final : public teq::iDerivativeFuncs {
  teq::TensptrT lderive(teq::FuncptrT op, teq::TensptrT supgrad, size_t arg_idx) const override;
  teq::TensptrT get_const_one(teq::iTensor& reference) const override;
  teq::TensptrT get_const_zero(teq::iTensor& reference) const override;
  teq::TensptrT add(teq::TensptrsT elems) const override;
};

/// tcr::derive (tenncor/src/eteq.cpp:40-76), local path
teq::TensptrsT derive(teq::TensptrT root, const teq::TensptrsT& targets);

/// evaluate `targets` through the context evaluator on the device (ETensor::calc,
/// tenncor/eteq/etens.hpp:151-162 / eteq::run, src/etens.cpp:46-62)
void run(const teq::TensptrsT& targets, const teq::TensSetT& ignored = {}, size_t max_version = std::numeric_limits<size_t>::max());

}  // namespace eteq

#endif  // TCR_HOST_ETEQ_HPP
