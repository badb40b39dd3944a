// pymod.cpp — pybind11 module `_tenncor`: the python surface of the reference's `tenncor`
// module (tenncor/python/eteq_ext.cpp:20-522, layr_ext.cpp, generated pyapi_tenncor.cpp)
// over the B200 back end. numpy shapes are REVERSED into teq shapes exactly like the
// reference (tenncor/pyutils/src/convert.cpp:8-37).
#include <fstream>
#include <iterator>

#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "api.hpp"
#include "dbg.hpp"
#include "dp.hpp"
#include "hone.hpp"
#include "onnx.hpp"
#include "planner.hpp"

namespace py = pybind11;
using namespace teq;
using layr::ETensor;
using layr::ETensorsT;

static DimsT c2pshape(const Shape& cshape) {
  DimsT fwd = narrow_shape(cshape);
  return DimsT(fwd.rbegin(), fwd.rend());
}

static Shape p2cshape(const std::vector<size_t>& pyshape) {
  DimsT slist(pyshape.rbegin(), pyshape.rend());
  return Shape(slist);
}

static egen::_GENERATED_DTYPE np2dtype(const py::dtype& dt) {
  if (dt.is(py::dtype::of<double>())) return egen::DOUBLE;
  if (dt.is(py::dtype::of<float>())) return egen::FLOAT;
  if (dt.is(py::dtype::of<int8_t>())) return egen::INT8;
  if (dt.is(py::dtype::of<uint8_t>())) return egen::UINT8;
  if (dt.is(py::dtype::of<int16_t>())) return egen::INT16;
  if (dt.is(py::dtype::of<uint16_t>())) return egen::UINT16;
  if (dt.is(py::dtype::of<int32_t>())) return egen::INT32;
  if (dt.is(py::dtype::of<uint32_t>())) return egen::UINT32;
  if (dt.is(py::dtype::of<int64_t>())) return egen::INT64;
  if (dt.is(py::dtype::of<uint64_t>())) return egen::UINT64;
  return egen::BAD_TYPE;
}

static py::dtype dtype2np(egen::_GENERATED_DTYPE t) {
  switch (t) {
    case egen::DOUBLE: return py::dtype::of<double>();
    case egen::FLOAT: return py::dtype::of<float>();
    case egen::INT8: return py::dtype::of<int8_t>();
    case egen::UINT8: return py::dtype::of<uint8_t>();
    case egen::INT16: return py::dtype::of<int16_t>();
    case egen::UINT16: return py::dtype::of<uint16_t>();
    case egen::INT32: return py::dtype::of<int32_t>();
    case egen::UINT32: return py::dtype::of<uint32_t>();
    case egen::INT64: return py::dtype::of<int64_t>();
    case egen::UINT64: return py::dtype::of<uint64_t>();
    default: global::fatal("bad dtype");
  }
}

static egen::_GENERATED_DTYPE parse_dtype(const py::object& o) {
  if (o.is_none()) return egen::default_dtype;
  if (py::isinstance<py::str>(o)) {
    std::string s = o.cast<std::string>();
    auto t = egen::get_type(s);
    if (t != egen::BAD_TYPE) return t;
  }
  auto t = np2dtype(py::dtype::from_args(o));
  if (t == egen::BAD_TYPE) global::fatal("unsupported dtype");
  return t;
}

/// numpy array -> (contiguous array of a supported dtype, teq shape); unsupported dtypes become FLOAT (PybindT)
static py::array normalise(py::array data, Shape& shape, egen::_GENERATED_DTYPE& dtype) {
  dtype = np2dtype(data.dtype());
  if (dtype == egen::BAD_TYPE) {
    data = py::array_t<float, py::array::c_style | py::array::forcecast>(data);
    dtype = egen::FLOAT;
  }
  data = py::array::ensure(data, py::array::c_style);
  std::vector<size_t> ps(data.shape(), data.shape() + data.ndim());
  shape = p2cshape(ps);
  return data;
}

static py::array to_array(iTensor& tens) {
  auto dtype = (egen::_GENERATED_DTYPE)tens.get_meta().type_code();
  const void* host = tens.device().data();
  if (nullptr == host) global::fatalf("%s has no data: evaluate it first", tens.to_string().c_str());
  DimsT ps = c2pshape(tens.shape());
  std::vector<py::ssize_t> pshape(ps.begin(), ps.end());
  py::array out(dtype2np(dtype), pshape);
  std::memcpy(out.mutable_data(), host, tens.shape().n_elems() * egen::type_size(dtype));
  return out;
}

// A result on its way to the host: ETensor.get_later() queues the device -> pinned-host copy behind the evaluation and returns at once,
// so the host can launch the next step while this one computes; result() waits for the copy (not for anything queued after it).
// The reference's session loop reads the error of every iteration synchronously (tenncor/pyutils: sess.update_target + get);
// on a device that read would idle the GPU for one host round trip per step.
struct PendingRead {
  void* host = nullptr;
  void* event = nullptr;
  size_t bytes = 0;
  egen::_GENERATED_DTYPE dtype = egen::BAD_TYPE;
  std::vector<py::ssize_t> shape;
  bool waited = false;
  PendingRead() = default;
  PendingRead(const PendingRead&) = delete;
  PendingRead& operator=(const PendingRead&) = delete;
  ~PendingRead() {
    if (event) { tcr_event_sync(event); tcr_event_destroy(event); }
    if (host) pool().emplace_back(bytes, host);
  }
  // pinned staging buffers are kept for reuse (cudaHostAlloc costs more than a training step)
  static std::vector<std::pair<size_t, void*>>& pool() {
    static auto* p = new std::vector<std::pair<size_t, void*>>();
    return *p;
  }
  static void* take(size_t bytes) {
    auto& p = pool();
    for (size_t i = 0; i < p.size(); ++i)
      if (p[i].first == bytes) { void* h = p[i].second; p.erase(p.begin() + i); return h; }
    void* h = nullptr;
    cuda::check(tcr_host_alloc(&h, bytes ? bytes : 1), "tcr_host_alloc");
    return h;
  }
  py::array result() {
    if (!waited) { cuda::check(tcr_event_sync(event), "tcr_event_sync"); waited = true; }
    py::array out(dtype2np(dtype), shape);
    std::memcpy(out.mutable_data(), host, bytes);
    return out;
  }
};

static std::unique_ptr<PendingRead> read_later(iTensor& tens) {
  void* dev = tens.device().device_data();
  if (nullptr == dev) global::fatalf("%s has no device data: evaluate it first", tens.to_string().c_str());
  std::unique_ptr<PendingRead> r(new PendingRead());
  r->dtype = (egen::_GENERATED_DTYPE)tens.get_meta().type_code();
  r->bytes = tens.shape().n_elems() * egen::type_size(r->dtype);
  DimsT ps = c2pshape(tens.shape());
  r->shape.assign(ps.begin(), ps.end());
  r->host = PendingRead::take(r->bytes);
  cuda::check(tcr_event_create(&r->event), "tcr_event_create");
  cuda::check(tcr_d2h(r->host, dev, r->bytes), "tcr_d2h");
  cuda::check(tcr_event_record(r->event), "tcr_event_record");
  return r;
}

static TensSetT to_set(const ETensorsT& ts) {
  TensSetT out;
  for (auto& t : ts) out.emplace(t.get());
  return out;
}

static py::object attr_to_py(const marsh::iObject* obj) {
  if (auto a = dynamic_cast<const marsh::IntArray*>(obj)) return py::cast(a->vals_);
  if (auto a = dynamic_cast<const marsh::PairArray*>(obj)) return py::cast(a->vals_);
  if (auto a = dynamic_cast<const marsh::Integer*>(obj)) return py::cast(a->val_);
  if (auto a = dynamic_cast<const marsh::Float*>(obj)) return py::cast(a->val_);
  if (auto a = dynamic_cast<const marsh::String*>(obj)) return py::cast(a->val_);
  return py::none();
}

/// post-order description of the graph under `targets`, in the exact order the reference
/// evaluator visits it; leaves carry a copy of their current data. Consumed by the CPU
/// oracle (tests only) and by debugging tools — the analogue of the reference's dbg print.
static py::list dump_graph(const ETensorsT& targets) {
  py::list out;
  std::unordered_map<iTensor*, size_t> ids;
  std::function<void(const TensptrT&)> visit = [&](const TensptrT& t) {
    if (ids.count(t.get())) return;
    py::dict node;
    if (auto f = dynamic_cast<iFunctor*>(t.get())) {
      py::list args;
      for (auto& a : f->args_ref()) {
        visit(a);
        args.append(ids.at(a.get()));
      }
      node["kind"] = "func";
      node["op"] = f->get_opcode().name_;
      node["args"] = args;
      py::dict attrs;
      for (auto& name : f->ls_attrs()) {
        auto obj = f->get_attr(name);
        if (auto ref = dynamic_cast<const TensorRef*>(obj)) {
          if (name == eigen::tensor_key) {
            Shape s = ref->get_tensor()->shape();
            attrs["tensor_shape"] = std::vector<int64_t>(s.begin(), s.end());
          }
          continue;
        }
        attrs[py::str(name)] = attr_to_py(obj);
      }
      node["attrs"] = attrs;
    } else {
      auto leaf = static_cast<iLeaf*>(t.get());
      node["kind"] = "leaf";
      node["usage"] = (int)leaf->get_usage();
      node["label"] = t->to_string();
      node["data"] = to_array(*t).attr("reshape")(-1);
    }
    Shape s = t->shape();
    node["id"] = ids.size();
    node["shape"] = std::vector<int64_t>(s.begin(), s.end());
    node["dtype"] = (int)t->get_meta().type_code();
    ids.emplace(t.get(), ids.size());
    out.append(node);
  };
  for (auto& t : targets) visit(t);
  return out;
}

struct InitHolder {
  layr::InitF f;
};

static eteq::VarptrT as_var(const ETensor& t) {
  auto v = std::dynamic_pointer_cast<eteq::Variable>(t);
  if (nullptr == v) global::fatalf("%s is not a variable", t->to_string().c_str());
  return v;
}

// a teq::iDevice that computes nothing and records the calc() calls an evaluator makes (the role of MockDevice in
// internal/teq/test/test_evaluator.cpp): the evaluators' traversal order, target flags and `ignored` handling become testable
// without a GPU
struct RecordingDevice final : public teq::iDevice {
  void calc(iTensor& tens, size_t cache_ttl) override { calls_.push_back({tens.shared_from_this(), cache_ttl}); }
  std::vector<std::pair<TensptrT, size_t>> calls_;
};

// MockRuntimeMemory of internal/eigen/mock/memory.hpp as a recorder: host blocks, every call logged
struct CountingMemory final : public eigen::iRuntimeMemory {
  void* allocate(size_t size) override {
    void* p = std::malloc(size ? size : 1);
    log_.push_back({"allocate", size, (uintptr_t)p});
    return p;
  }
  void deallocate(void* ptr, size_t size) override {
    log_.push_back({"deallocate", size, (uintptr_t)ptr});
    std::free(ptr);
  }
  std::vector<std::tuple<std::string, size_t, uintptr_t>> log_;
};

struct LaunchLog {
  std::vector<std::pair<uintptr_t, std::vector<uintptr_t>>> calls_;
};

// an api.init.* initializer object or a python callable (numpy_shape, label) -> EVariable
static layr::InitF to_initf(py::object f) {
  if (f.is_none()) return layr::InitF();
  if (py::isinstance<InitHolder>(f)) return f.cast<InitHolder&>().f;
  py::function pf = f.cast<py::function>();
  return [pf](Shape shape, std::string label) {
    DimsT ps = c2pshape(shape);
    return as_var(pf(std::vector<size_t>(ps.begin(), ps.end()), label).cast<ETensor>());
  };
}

static std::string& log_level() {
  static std::string level = "info";
  return level;
}

#define UN(NAME, OP) api.def(NAME, [](const ETensor& x) { return tenncor::unary(egen::OP, x); }, py::arg("input"))
#define BIN(NAME, OP)                                                                                               \
  api.def(NAME, [](const ETensor& a, const ETensor& b) { return tenncor::binary(egen::OP, a, b); });                \
  api.def(NAME, [](const ETensor& a, double b) { return tenncor::binary(egen::OP, a, b); });                        \
  api.def(NAME, [](double a, const ETensor& b) { return tenncor::binary(egen::OP, a, b); })

PYBIND11_MODULE(_tenncor, m) {
  m.doc() = "tenncor_b200 host module: TEQ functor graphs evaluated on B200 (sm_100a)";
  py::register_exception<global::FatalError>(m, "FatalError", PyExc_RuntimeError);

  py::class_<PendingRead>(m, "PendingRead")
      .def("result", &PendingRead::result, "Wait for the queued device -> host copy and return the value as a numpy array")
      .def("done", [](PendingRead& self) { return self.waited; });

  py::class_<iTensor, TensptrT> etens(m, "ETensor");
  etens.def("__str__", [](const iTensor& self) { return self.to_string(); })
      .def("__hash__", [](const iTensor& self) { return (size_t)&self; })
      .def("__eq__", [](const iTensor& self, py::object other) {
        return py::isinstance<iTensor>(other) && other.cast<iTensor*>() == &self;
      })
      .def("shape", [](const iTensor& self) {
        DimsT ps = c2pshape(self.shape());
        return std::vector<size_t>(ps.begin(), ps.end());
      }, "Return this instance's (numpy-ordered) shape")
      .def("teq_shape", [](const iTensor& self) { Shape s = self.shape(); return std::vector<size_t>(s.begin(), s.end()); })
      .def("dtype", [](const iTensor& self) { return dtype2np((egen::_GENERATED_DTYPE)self.get_meta().type_code()); })
      .def("data", [](iTensor& self) { return to_array(self); }, "Host copy of the current data (D2H sync point)")
      .def("get", [](TensptrT self, ETensorsT ignored, size_t max_version) {
        eteq::run({self}, to_set(ignored), max_version);
        return to_array(*self);
      }, py::arg("ignored") = ETensorsT{}, py::arg("max_version") = std::numeric_limits<size_t>::max(),
      "Evaluate on the device and return the result as a numpy array")
      .def("get_later", [](TensptrT self, ETensorsT ignored, size_t max_version) {
        eteq::run({self}, to_set(ignored), max_version);
        return read_later(*self);
      }, py::arg("ignored") = ETensorsT{}, py::arg("max_version") = std::numeric_limits<size_t>::max(),
      "Evaluate on the device and queue the copy of the result to pinned host memory; returns a PendingRead whose result() waits for "
      "that copy only — the host is free to launch the next step meanwhile")
      .def("calc", [](TensptrT self, ETensorsT ignored, size_t max_version) {
        eteq::run({self}, to_set(ignored), max_version);
      }, py::arg("ignored") = ETensorsT{}, py::arg("max_version") = std::numeric_limits<size_t>::max(),
      "Evaluate on the device; the result stays in HBM (no host copy)")
      .def("release_data", [](iTensor& self) {  // get_releasedata (eteq_ext.cpp:8-18): read through odata(), i.e. this read counts as one consumer
        py::array out = to_array(self);
        if (nullptr != dynamic_cast<iFunctor*>(&self)) { auto once = self.device().odata(); (void)once; }  // leaf storage has no lifetime to spend
        return out;
      }, "Host copy of the current data that spends one of the result's planned reads (a temporary is released after its last one)")
      .def("release_get", [](TensptrT self, ETensorsT ignored, size_t max_version) {  // ETensor::calc_release (etens.hpp:165-176)
        eteq::run({self}, to_set(ignored), max_version);
        py::array out = to_array(*self);
        { auto once = self->device().odata(); (void)once; }
        return out;
      }, py::arg("ignored") = ETensorsT{}, py::arg("max_version") = std::numeric_limits<size_t>::max(),
      "Evaluate, return the result and release the target's buffer back to the arena")
      .def("cache", [](iTensor& self) {  // Observable::cache_init (eteq_ext.cpp:163-170, functor.hpp:229-243)
        if (nullptr == dynamic_cast<eigen::Observable*>(&self)) return;
        if (auto op = dynamic_cast<cuda::DevOp*>(&self.device())) op->pin();
      }, "Keep this functor's result across its consumers' reads (evaluations that name it in `ignored` read it as data)")
      .def("tag", [](iTensor& self, const std::string& key, const std::string& val) {  // eteq_ext.cpp:193-201
        if (auto f = dynamic_cast<iFunctor*>(&self)) f->add_attr(key, std::make_unique<marsh::String>(val));
      }, py::arg("key"), py::arg("val"), "Attach a string attribute to a functor")
      .def("label", [](const iTensor& self) { return self.to_string(); }, "iTensor::to_string: a leaf's label / constant value, a functor's opcode name")
      .def("is_leaf", [](const iTensor& self) { return nullptr == dynamic_cast<const iFunctor*>(&self); })
      .def("usage", [](const iTensor& self) -> std::string {  // teq::get_usage_name (internal/teq/src/ileaf.cpp:12-35); "" for functors
        auto leaf = dynamic_cast<const teq::iLeaf*>(&self);
        if (nullptr == leaf) return "";
        switch (leaf->get_usage()) {
          case teq::IMMUTABLE: return "constant";
          case teq::VARUSAGE: return "variable";
          case teq::PLACEHOLDER: return "placeholder";
          default: return "unknown";
        }
      })
      .def("device_ptr", [](iTensor& self) { return (uintptr_t)self.device().device_data(); })
      .def_property_readonly("__cuda_array_interface__", [](iTensor& self) {
        // zero-copy view of the resident HBM buffer for torch.as_tensor / cupy.asarray (CUDA Array Interface v3);
        // `stream` names the library stream so that consumers order themselves after the producing kernels
        void* ptr = self.device().device_data();
        if (nullptr == ptr) global::fatalf("%s has no device data: evaluate it first (calc / get)", self.to_string().c_str());
        DimsT ps = c2pshape(self.shape());
        py::tuple shape(ps.size());
        for (size_t i = 0; i < ps.size(); ++i) shape[i] = (size_t)ps[i];
        py::dict d;
        d["shape"] = shape;
        d["typestr"] = dtype2np((egen::_GENERATED_DTYPE)self.get_meta().type_code()).attr("str");
        d["data"] = py::make_tuple((uintptr_t)ptr, false);
        d["strides"] = py::none();
        d["version"] = 3;
        uintptr_t stream = (uintptr_t)tcr_stream();
        d["stream"] = stream ? py::cast(stream) : py::cast(1);
        return d;
      })
      .def("get_version", [](const iTensor& self) { return self.get_meta().state_version(); })
      .def("type_label", [](const iTensor& self) { return self.get_meta().type_label(); })
      .def("type_size", [](const iTensor& self) { return self.get_meta().type_size(); })
      // eigen::Observable (internal/eigen/observable.hpp) / eteq::Functor (tenncor/eteq/functor.hpp:106-269) as the reference's
      // tests drive them (tenncor/eteq/test/test_functor.cpp)
      .def("has_data", [](const iTensor& self) { auto f = dynamic_cast<const eigen::Observable*>(&self); return f ? f->has_data() : true; })
      .def("must_initialize", [](iTensor& self) { if (auto f = dynamic_cast<eigen::Observable*>(&self)) f->must_initialize(); })
      .def("uninitialize", [](iTensor& self) { if (auto f = dynamic_cast<eigen::Observable*>(&self)) f->uninitialize(); })
      .def("prop_version", [](iTensor& self, size_t max_version) {
        auto f = dynamic_cast<eigen::Observable*>(&self);
        if (nullptr == f) global::fatalf("%s is not a functor", self.to_string().c_str());
        return f->prop_version(max_version);
      }, py::arg("max_version") = std::numeric_limits<size_t>::max())
      .def("update_child", [](iTensor& self, TensptrT arg, size_t index) {
        auto f = dynamic_cast<iFunctor*>(&self);
        if (nullptr == f) global::fatalf("%s is not a functor", self.to_string().c_str());
        f->update_child(arg, index);
      }, py::arg("arg"), py::arg("index"))
      .def("nsubs", [](const iTensor& self) { auto f = dynamic_cast<const eigen::Observable*>(&self); return f ? f->nsubs() : (size_t)0; })
      .def("clone", [](const iTensor& self) { return TensptrT(self.clone()); })
      .def("opname", [](const iTensor& self) {
        auto f = dynamic_cast<const iFunctor*>(&self);
        return f ? f->get_opcode().name_ : std::string();
      })
      .def("args", [](const iTensor& self) {
        auto f = dynamic_cast<const iFunctor*>(&self);
        return f ? f->get_args() : TensptrsT{};
      })
      .def("get_input", [](TensptrT self) { return layr::get_input(self); })
      .def("connect", [](TensptrT self, const ETensor& input) { return layr::connect(self, input); })
      .def("deep_clone", [](TensptrT self) { return layr::deep_clone(self); })
      .def("get_storage", [](TensptrT self) {
        auto vars = layr::get_storage(self);
        return ETensorsT(vars.begin(), vars.end());
      })
      .def("__neg__", [](TensptrT a) { return tenncor::neg(a); })
      .def("__add__", [](TensptrT a, const ETensor& b) { return tenncor::add(a, b); })
      .def("__add__", [](TensptrT a, double b) { return tenncor::add(a, b); })
      .def("__radd__", [](TensptrT a, double b) { return tenncor::add(b, a); })
      .def("__sub__", [](TensptrT a, const ETensor& b) { return tenncor::sub(a, b); })
      .def("__sub__", [](TensptrT a, double b) { return tenncor::sub(a, b); })
      .def("__rsub__", [](TensptrT a, double b) { return tenncor::sub(b, a); })
      .def("__mul__", [](TensptrT a, const ETensor& b) { return tenncor::mul(a, b); })
      .def("__mul__", [](TensptrT a, double b) { return tenncor::mul(a, b); })
      .def("__rmul__", [](TensptrT a, double b) { return tenncor::mul(b, a); })
      .def("__truediv__", [](TensptrT a, const ETensor& b) { return tenncor::div(a, b); })
      .def("__truediv__", [](TensptrT a, double b) { return tenncor::div(a, b); })
      .def("__rtruediv__", [](TensptrT a, double b) { return tenncor::div(b, a); })
      .def("__lt__", [](TensptrT a, const ETensor& b) { return tenncor::lt(a, b); })
      .def("__gt__", [](TensptrT a, const ETensor& b) { return tenncor::gt(a, b); });

  py::class_<eteq::Variable, iTensor, eteq::VarptrT>(m, "EVariable")
      .def(py::init([](std::vector<size_t> slist, double scalar, const std::string& label, py::object dtype) {
        return eteq::make_variable_scalar(scalar, p2cshape(slist), label, parse_dtype(dtype));
      }), py::arg("shape"), py::arg("scalar") = 0, py::arg("label") = "", py::arg("dtype") = py::none())
      .def("assign", [](eteq::Variable& self, py::object obj) {
        if (py::hasattr(obj, "__cuda_array_interface__")) {
          // a device array (torch / cupy / another tensor of this module): HBM -> HBM, no host round trip. The producer's
          // stream must already be synchronised with the library stream (tc.sync() / torch.cuda.synchronize()).
          py::dict cai = obj.attr("__cuda_array_interface__");
          if (!cai["strides"].is_none()) global::fatal("assign from a device array needs a C-contiguous array");
          std::vector<size_t> ps = cai["shape"].cast<std::vector<size_t>>();
          Shape shape = p2cshape(ps);
          if (false == shape.compatible_after(self.shape(), 0))
            global::fatalf("assigning data shaped %s to tensor %s", shape.to_string().c_str(), self.shape().to_string().c_str());
          py::dtype want = dtype2np((egen::_GENERATED_DTYPE)self.get_meta().type_code());
          if (cai["typestr"].cast<std::string>() != want.attr("str").cast<std::string>())
            global::fatalf("assign from a device array needs dtype %s (got %s): conversions happen on the host path only",
                           want.attr("str").cast<std::string>().c_str(), cai["typestr"].cast<std::string>().c_str());
          self.assign_device((const void*)cai["data"].cast<py::tuple>()[0].cast<uintptr_t>());
          return;
        }
        py::array data = py::array::ensure(obj);
        if (!data) global::fatal("assign needs a numpy-convertible array or a device array");
        Shape shape;
        egen::_GENERATED_DTYPE dtype;
        py::array arr = normalise(data, shape, dtype);
        self.assign(arr.data(), dtype, shape);
      }, py::arg("data"), "Assign a numpy array (host -> HBM, asynchronous for pinned arrays) or any object exposing __cuda_array_interface__ (HBM -> HBM)")
      .def("prefetch", [](eteq::Variable& self, py::array data) {
        Shape shape;
        egen::_GENERATED_DTYPE dtype;
        py::array arr = normalise(data, shape, dtype);
        if (arr.data() != data.data()) global::fatal("prefetch needs a C-contiguous array that stays alive until commit (pinned for overlap)");
        self.prefetch(arr.data(), dtype, shape);
      }, py::arg("data"), "Start copying the NEXT batch host -> HBM on the copy stream; overlaps the step being evaluated")
      .def("commit", [](eteq::Variable& self) { self.commit(); }, "Make the prefetched batch this variable's data (version bump like assign)")
      .def("touch", [](eteq::Variable& self) { self.upversion(eteq::get_lastvers() + 1); },
           "Bump the version as if new data had been assigned (the data already in HBM is kept)")
      .def("assign_device", [](eteq::Variable& self, uintptr_t dev_ptr) { self.assign_device((const void*)dev_ptr); },
           "Assign from a device pointer holding this variable's dtype and element count");

  // ---- evaluation / creation
  m.def("run", [](ETensorsT targets, ETensorsT ignored, size_t max_version) {
    eteq::run(targets, to_set(ignored), max_version);
    std::vector<py::array> out;
    for (auto& t : targets) out.push_back(to_array(*t));
    return out;
  }, py::arg("targets"), py::arg("ignored") = ETensorsT{}, py::arg("max_version") = std::numeric_limits<size_t>::max());
  m.def("scalar_constant", [](double scalar, std::vector<size_t> slist, py::object dtype) {
    return eteq::make_constant_scalar(scalar, p2cshape(slist), parse_dtype(dtype));
  }, py::arg("scalar"), py::arg("slist"), py::arg("dtype") = py::none(), "Return scalar constant etens");
  m.def("constant", [](py::array data) {
    Shape shape;
    egen::_GENERATED_DTYPE dtype;
    py::array arr = normalise(data, shape, dtype);
    return eteq::make_constant_tensor(arr.data(), dtype, shape);
  }, "Return constant etens with data");
  m.def("scalar_variable", [](double scalar, std::vector<size_t> slist, const std::string& label, py::object dtype) {
    return eteq::make_variable_scalar(scalar, p2cshape(slist), label, parse_dtype(dtype));
  }, py::arg("scalar"), py::arg("slist"), py::arg("label") = "", py::arg("dtype") = py::none());
  m.def("variable_like", [](double scalar, ETensor like, const std::string& label) {
    return eteq::make_variable_scalar(scalar, like->shape(), label, (egen::_GENERATED_DTYPE)like->get_meta().type_code());
  }, py::arg("scalar"), py::arg("like"), py::arg("label") = "");
  m.def("variable", [](py::array data, const std::string& label) {
    Shape shape;
    egen::_GENERATED_DTYPE dtype;
    py::array arr = normalise(data, shape, dtype);
    return eteq::make_variable(arr.data(), dtype, shape, label);
  }, py::arg("data"), py::arg("label") = "");
  m.def("placeholder", [](py::array data, const std::string& label) {
    Shape shape;
    egen::_GENERATED_DTYPE dtype;
    py::array arr = normalise(data, shape, dtype);
    return eteq::VarptrT(eteq::Variable::get(arr.data(), dtype, shape, label, teq::PLACEHOLDER));  // Variable<T>::get(..., teq::PLACEHOLDER)
  }, py::arg("data"), py::arg("label") = "", "A variable with PLACEHOLDER usage: saved as a graph input instead of an initializer");
  m.def("to_variable", [](const ETensor& tens) { return as_var(tens); });
  m.def("derive", [](const ETensor& root, const ETensorsT& targets) { return tenncor::derive(root, targets); },
        "Return derivative of first tensor with respect to each target");
  m.def("trail", [](const ETensor& root, const std::vector<std::pair<ETensor, ETensor>>& inps) {
    OwnMapT inputs;
    for (auto& p : inps) inputs.emplace(p.first.get(), p.second);
    return layr::trail(root, inputs);
  });
  m.def("seed", &tenncor::seed, "Seed internal RNG");
  m.def("dump_graph", &dump_graph, "Post-order node list (evaluation order) of the graph under the targets");
  m.def("dump_ids", [](const ETensorsT& targets, py::object) {
    // ids of `targets` in the numbering dump_graph uses
    std::unordered_map<iTensor*, size_t> ids;
    std::function<void(const TensptrT&)> visit = [&](const TensptrT& t) {
      if (ids.count(t.get())) return;
      if (auto f = dynamic_cast<iFunctor*>(t.get()))
        for (auto& a : f->args_ref()) visit(a);
      ids.emplace(t.get(), ids.size());
    };
    py::dict out;
    for (auto& t : targets) {
      visit(t);
      out[py::cast(t)] = ids.at(t.get());
    }
    return out;
  });
  m.def("apply_update", [](const ETensorsT& models, std::function<layr::VarErrsT(const ETensor&, const ETensorsT&)> update, layr::ErrorF err) {
    layr::ApproxF approx = [update](const ETensor& e, const eteq::VarptrsT& vars) { return update(e, ETensorsT(vars.begin(), vars.end())); };
    return trainer::apply_update(models, approx, err);
  }, py::arg("models"), py::arg("update"), py::arg("err_func"));

  // ---- the generated egen surface, for the host-logic tests that mirror internal/eigen/test/test_shaper.cpp / test_funcopt.cpp
  auto eg = m.def_submodule("egen", "ShapeParser / TypeParser / FuncOpt of the opcode table (cfg/ops.yml)");
  auto to_maps = [](const py::dict& d) {
    // values are packed the way eigen::Packer packs them (internal/eigen/packattr.hpp): the key decides the attribute kind
    marsh::Maps attrs;
    for (auto kv : d) {
      const std::string key = kv.first.cast<std::string>();
      py::handle v = kv.second;
      if (key == eigen::dimpairs_key) eigen::pack_attr(attrs, v.cast<eigen::PairVecT<DimT>>());
      else if (key == eigen::rankpairs_key) eigen::pack_attr(attrs, v.cast<eigen::PairVecT<RankT>>());
      else if (key == eigen::dims_key) eigen::pack_attr(attrs, v.cast<DimsT>());
      else if (key == eigen::ranks_key) eigen::pack_attr(attrs, v.cast<teq::RanksT>());
      else if (key == eigen::rankset_key) eigen::pack_attr(attrs, v.cast<std::set<RankT>>());
      else if (key == eigen::rank_key) eigen::pack_attr(attrs, v.cast<RankT>());
      else if (key == eigen::shape_key) eigen::pack_attr(attrs, Shape(v.cast<DimsT>()));
      else if (key == eigen::dtype_key) eigen::pack_attr(attrs, egen::get_type(v.cast<std::string>()));
      else if (key == eigen::tensor_key) eigen::pack_attr(attrs, v.cast<ETensor>());
      else global::fatalf("unknown attribute key `%s`", key.c_str());
    }
    return attrs;
  };
  eg.def("shape_parse", [to_maps](const std::string& opname, const py::dict& attrs, const std::vector<DimsT>& shapes) {
    teq::ShapesT ss;
    for (auto& sh : shapes) ss.push_back(Shape(sh));
    marsh::Maps m = to_maps(attrs);
    Shape out = eigen::shape_parse(egen::get_op(opname), m, ss);
    return std::vector<size_t>(out.begin(), out.end());
  }, py::arg("opname"), py::arg("attrs"), py::arg("shapes"), "teq shape (rank 0 first, padded to 8) the opcode's shape rule gives");
  eg.def("type_parse", [to_maps](const std::string& opname, const py::dict& attrs, const std::vector<std::string>& dtypes) {
    eigen::DTypesT ds;
    for (auto& d : dtypes) ds.push_back(egen::get_type(d));
    marsh::Maps m = to_maps(attrs);
    return egen::name_type(eigen::type_parse(egen::get_op(opname), m, ds));
  }, py::arg("opname"), py::arg("attrs"), py::arg("dtypes"));
  eg.def("func_opt", [to_maps](const std::string& opname, const py::dict& attrs, const ETensorsT& args, const std::string& out_dtype) {
    marsh::Maps m = to_maps(attrs);
    return eigen::func_opt(egen::get_op(opname), egen::get_type(out_dtype), m, args);
  }, py::arg("opname"), py::arg("attrs"), py::arg("args"), py::arg("out_dtype") = "DOUBLE", "true when the functor would be redundant (make_funcattr returns its first argument)");
  eg.def("make_functor", [to_maps](const std::string& opname, const ETensorsT& args, const py::dict& attrs) {
    marsh::Maps m = to_maps(attrs);
    return eteq::make_funcattr(egen::get_op(opname), args, m);  // eteq::make_functor with the attributes eigen::Packer would pack
  }, py::arg("opname"), py::arg("args"), py::arg("attrs") = py::dict());
  eg.def("make_tfunctor", [to_maps](const std::string& dtype, const std::string& opname, ETensorsT args, const py::dict& attrs) {
    marsh::Maps m = to_maps(attrs);
    return eteq::make_tfuncattr(egen::get_type(dtype), egen::get_op(opname), std::move(args), m);  // make_tfuncattr<T>, make.hpp:66-84
  }, py::arg("dtype"), py::arg("opname"), py::arg("args"), py::arg("attrs") = py::dict(),
  "eteq::make_tfunctor<T>: the functor in an explicit element type; TypeCaster (caster.hpp:10-44) wraps arguments of another type in CAST");
  eg.def("lderive", [](const ETensor& op, const ETensor& supgrad, size_t arg_idx) {
    auto f = std::dynamic_pointer_cast<teq::iFunctor>(op);
    if (nullptr == f) global::fatalf("%s is not a functor", op->to_string().c_str());
    return eteq::DerivativeFuncs().lderive(f, supgrad, arg_idx);
  }, py::arg("op"), py::arg("supgrad"), py::arg("arg_idx"), "DerivativeFuncs::lderive (tenncor/eteq/backprop.hpp:58-545): the local gradient rule of one functor");
  eg.def("const_zero", [](ETensor like) { return eteq::DerivativeFuncs().get_const_zero(*like); });
  eg.def("const_one", [](ETensor like) { return eteq::DerivativeFuncs().get_const_one(*like); });
  eg.def("grad_add", [](const ETensorsT& elems) { return eteq::DerivativeFuncs().add(elems); });
  eg.def("is_commutative", [](const std::string& opname) { return egen::is_commutative(egen::get_op(opname)); });
  eg.def("is_idempotent", [](const std::string& opname) { return egen::is_idempotent(egen::get_op(opname)); });
  eg.def("dtypes", [] {
    // (name, bytes per element, conversion precision rank) in enum order, as tools/egen/plugins/dtypes.py generates them from the type file
    std::vector<std::tuple<std::string, size_t, size_t>> out;
    for (int t = 1; t < egen::_N_GENERATED_DTYPES; ++t) {
      auto dt = (egen::_GENERATED_DTYPE)t;
      out.push_back({egen::name_type(dt), (size_t)egen::type_size(dt), (size_t)egen::type_precision(dt)});
    }
    return out;
  });
  eg.def("default_dtype", [] { return egen::name_type(egen::default_dtype); });
  eg.def("opcodes", [] {
    std::vector<std::string> out;
    for (int op = 1; op < egen::_N_GENERATED_OPCODES; ++op) out.push_back(egen::name_op((egen::_GENERATED_OPCODE)op));
    return out;
  });

  // ---- teq: shapes and the graph travelers the derivative builder and the evaluators stand on (internal/teq/shape.hpp, traveler.hpp)
  auto tq = m.def_submodule("teq", "teq::Shape and travelers, as internal/teq/test/test_shape.cpp / test_traveler.cpp drive them");
  py::class_<Shape>(tq, "Shape")
      .def(py::init<>())
      .def(py::init([](const DimsT& dims) { return Shape(dims); }))
      .def("at", [](const Shape& self, size_t idx) { return (size_t)self.at((RankT)std::min<size_t>(idx, 255)); })
      .def("n_elems", [](const Shape& self) { return (uint64_t)self.n_elems(); })
      .def("compatible_before", [](const Shape& self, const Shape& other, size_t idx) { return self.compatible_before(other, (RankT)idx); })
      .def("compatible_after", [](const Shape& self, const Shape& other, size_t idx) { return self.compatible_after(other, (RankT)idx); })
      .def("to_list", [](const Shape& self) { return std::vector<size_t>(self.begin(), self.end()); })
      .def("narrow", [](const Shape& self) { DimsT d = teq::narrow_shape(self); return std::vector<size_t>(d.begin(), d.end()); })
      .def("__len__", [](const Shape&) { return (size_t)rank_cap; })
      .def("__eq__", [](const Shape& a, const Shape& b) { return a == b; })
      .def("__str__", [](const Shape& self) { return self.to_string(); });
  tq.attr("rank_cap") = (size_t)rank_cap;
  tq.def("graph_stat", [](const ETensor& root) {
    teq::GraphStat stat;
    root->accept(stat);
    std::vector<std::pair<ETensor, size_t>> out;
    for (auto& kv : stat.height_) out.push_back({kv.first->shared_from_this(), kv.second});
    return out;
  }, "GraphStat (traveler.hpp:51-110): (tensor, longest distance to a leaf) for every node under root");
  tq.def("graph_index", [](const ETensor& root) {
    teq::GraphIndex index;
    root->accept(index);
    std::vector<std::pair<ETensor, size_t>> out;
    for (auto& kv : index.indices_) out.push_back({kv.first->shared_from_this(), kv.second});
    return out;
  }, "GraphIndex (traveler.hpp:112-148): post-order index of every node under root");
  tq.def("path_finder", [](const ETensor& root, const ETensorsT& targets, bool follow_attrs) {
    teq::PathFinder finder(to_set(targets), follow_attrs);
    root->accept(finder);
    std::vector<std::tuple<ETensor, std::vector<size_t>, std::vector<std::string>>> out;
    for (auto& kv : finder.roadmap_) out.push_back({kv.first->shared_from_this(), kv.second.args_, kv.second.attrs_});
    return out;
  }, py::arg("root"), py::arg("targets"), py::arg("follow_attrs") = true,
  "PathFinder (traveler.hpp:200-300): the functors under root that lead to a target, with the argument indices and attribute names to follow");
  tq.def("copy_graph", [](const ETensor& root, const ETensorsT& ignores) {
    teq::Copier copier(to_set(ignores));
    root->accept(copier);
    std::vector<std::pair<ETensor, ETensor>> out;
    for (auto& kv : copier.clones_) out.push_back({kv.first->shared_from_this(), kv.second});
    return out;
  }, py::arg("root"), py::arg("ignores") = ETensorsT{}, "Copier (traveler.hpp:378-432): (original, clone) pairs; ignored nodes are shared, not cloned");
  tq.def("attr_tensors", [](const ETensor& func) {
    auto f = dynamic_cast<iFunctor*>(func.get());
    return f ? teq::attr_tensors(*f) : TensptrsT{};
  }, "FindTensAttr (objs.hpp): tensors referenced by a functor's attributes");

  // ---- host random generators and logging level (eteq_ext.cpp:383-405)
  m.def("unif_gen", [](double lower, double upper) {
    return py::cpp_function([lower, upper]() { return std::uniform_real_distribution<double>(lower, upper)(tenncor::host_rng()); });
  }, py::arg("lower") = 0, py::arg("upper") = 1, "Return a generator function drawing U[lower, upper) from the seeded host generator");
  m.def("norm_gen", [](double mean, double stdev) {
    return py::cpp_function([mean, stdev]() { return std::normal_distribution<double>(mean, stdev)(tenncor::host_rng()); });
  }, py::arg("mean") = 0, py::arg("stdev") = 1);
  m.def("set_log_level", [](const std::string& level) { log_level() = level; }, py::arg("level"), "Set log level (recorded; the host reports errors as exceptions)");
  m.def("get_log_level", [] { return log_level(); });
  m.def("variable_from_init", [](py::object init, std::vector<size_t> slist, const std::string& label) {
    return to_initf(init)(p2cshape(slist), label);
  }, py::arg("init"), py::arg("slist"), py::arg("label") = "", "Return labelled variable containing data created from initializer");

  // ---- serialization (tenncor/python/eteq_ext.cpp:408-487)
  m.def("load_from_file", [](const std::string& filename, const std::unordered_map<std::string, size_t>& key_prec) {
    return onnx::load_from_file(filename, key_prec);
  }, py::arg("filename"), py::arg("key_prec") = std::unordered_map<std::string, size_t>{});
  m.def("onnx_describe", [](const py::bytes& data) {
    // the parsed ModelProto as plain python objects: lets a test compare two files message by message (the role of
    // google::protobuf::util::MessageDifferencer in tenncor/test/test_serialize.cpp SaveGraph)
    onnx::ModelProto pb;
    onnx::parse(pb, std::string(data));
    std::function<py::dict(const onnx::TensorProto&)> tensor = [](const onnx::TensorProto& t) {
      py::dict d;
      d["dims"] = t.dims; d["data_type"] = t.data_type; d["name"] = t.name; d["float_data"] = t.float_data; d["int32_data"] = t.int32_data;
      d["int64_data"] = t.int64_data; d["double_data"] = t.double_data; d["uint64_data"] = t.uint64_data; d["raw_data"] = py::bytes(t.raw_data);
      return d;
    };
    std::function<py::dict(const onnx::GraphProto&)> graph = [&](const onnx::GraphProto& g) {
      py::list nodes, inits, ins, outs, annos;
      for (auto& n : g.node) {
        py::list attrs;
        for (auto& a : n.attribute) {
          py::dict d;
          d["name"] = a.name; d["type"] = a.type; d["f"] = a.f; d["i"] = a.i; d["s"] = py::bytes(a.s); d["floats"] = a.floats; d["ints"] = a.ints;
          py::list strs, tens;
          for (auto& st : a.strings) strs.append(py::bytes(st));
          for (auto& t : a.tensors) tens.append(tensor(t));
          d["strings"] = strs; d["tensors"] = tens; d["t"] = tensor(a.t);
          d["g"] = a.g ? py::object(graph(*a.g)) : py::object(py::none());
          attrs.append(d);
        }
        py::dict d;
        d["input"] = n.input; d["output"] = n.output; d["name"] = n.name; d["op_type"] = n.op_type; d["attribute"] = attrs;
        nodes.append(d);
      }
      for (auto& t : g.initializer) inits.append(tensor(t));
      auto vinfo = [](const onnx::ValueInfoProto& v) { py::dict d; d["name"] = v.name; d["elem_type"] = v.elem_type; d["dims"] = v.dims; return d; };
      for (auto& v : g.input) ins.append(vinfo(v));
      for (auto& v : g.output) outs.append(vinfo(v));
      for (auto& a : g.quantization_annotation) { py::dict d; d["tensor_name"] = a.tensor_name; d["params"] = a.quant_parameter_tensor_names; annos.append(d); }
      py::dict d;
      d["node"] = nodes; d["name"] = g.name; d["initializer"] = inits; d["input"] = ins; d["output"] = outs; d["quantization_annotation"] = annos;
      return d;
    };
    py::dict d;
    d["ir_version"] = pb.ir_version; d["model_version"] = pb.model_version; d["producer_name"] = pb.producer_name;
    d["producer_version"] = pb.producer_version; d["domain"] = pb.domain; d["graph"] = graph(pb.graph);
    return d;
  }, py::arg("data"));
  m.def("load_model_ids", [](const std::string& filename) {
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) global::fatalf("failed to read file `%s`", filename.c_str());
    std::string bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    onnx::ModelProto pb;
    onnx::parse(pb, bytes);
    onnx::TensIds ids;
    ETensorsT roots = onnx::load_model(ids, pb);  // tcr::load_model(ids, model), tenncor/src/serial.cpp:53-85
    std::map<std::string, ETensor> named(ids.by_id.begin(), ids.by_id.end());
    return py::make_tuple(roots, named);
  }, py::arg("filename"), "tcr::load_model with its id map: (graph outputs in file order, {id: tensor})");
  m.def("save_to_file", [](const std::string& filename, const ETensorsT& models, const std::map<std::string, ETensor>& keys) {
    std::vector<std::pair<std::string, ETensor>> k(keys.begin(), keys.end());
    return onnx::save_to_file(filename, models, k);
  }, py::arg("filename"), py::arg("models"), py::arg("keys") = std::map<std::string, ETensor>{});

  // ---- evaluator plugins (dbg/python/peval.cpp:9-80) + per-opcode profiler
  py::class_<teq::iEvaluator, std::shared_ptr<teq::iEvaluator>>(m, "iEvaluator");
  py::class_<dbg::iPlugin, std::shared_ptr<dbg::iPlugin>>(m, "Plugin");
  py::class_<dbg::PlugableEvaluator, teq::iEvaluator, std::shared_ptr<dbg::PlugableEvaluator>>(m, "PlugableEvaluator")
      .def(py::init<>())
      .def("add_plugin", &dbg::PlugableEvaluator::add_plugin);
  py::class_<dbg::Inspector, dbg::iPlugin, std::shared_ptr<dbg::Inspector>>(m, "Inspector")
      .def(py::init<>())
      .def("add", &dbg::Inspector::add, py::arg("target"), py::arg("label") = "")
      .def("last", [](dbg::Inspector& self) { return self.last_; }, "label -> (min, max) seen by the latest evaluation");
  py::class_<dbg::OpProfiler, teq::iEvaluator, std::shared_ptr<dbg::OpProfiler>>(m, "OpProfiler")
      .def(py::init<>())
      .def("reset", &dbg::OpProfiler::reset)
      .def("report", [](dbg::OpProfiler& self) {
        py::list out;
        double total = 0;
        for (auto& kv : self.stats_) total += kv.second.ms;
        for (auto& kv : self.stats_) {
          py::dict d;
          d["opcode"] = kv.first; d["calls"] = kv.second.calls; d["ms"] = kv.second.ms; d["bytes"] = kv.second.bytes;
          d["share"] = total > 0 ? kv.second.ms / total : 0.0;
          d["GBps"] = kv.second.ms > 0 ? kv.second.bytes / kv.second.ms / 1e6 : 0.0;
          out.append(d);
        }
        return out;
      }, "Per-opcode calls, device ms (CUDA events), algorithmic bytes, time share and GB/s since the last reset");
  py::class_<teq::Evaluator, teq::iEvaluator, std::shared_ptr<teq::Evaluator>>(m, "Evaluator").def(py::init<>());  // eteq_ext.cpp:205
  py::class_<cuda::PlanEvaluator, teq::iEvaluator, std::shared_ptr<cuda::PlanEvaluator>>(m, "PlanEvaluator").def(py::init<>());
  m.def("evaluate", [](std::shared_ptr<teq::iEvaluator> self, ETensorsT targeted, size_t max_version, ETensorsT ignored) {
    cuda::Device device(max_version);  // iEvaluator.evaluate (eteq_ext.cpp:208-225), as a function taking the evaluator
    self->evaluate(device, to_set(targeted), to_set(ignored));
  }, py::arg("evaluator"), py::arg("targeted"), py::arg("max_version") = std::numeric_limits<size_t>::max(), py::arg("ignored") = ETensorsT{});
  py::class_<RecordingDevice>(m, "RecordingDevice")
      .def(py::init<>())
      .def("calls", [](RecordingDevice& self) { return self.calls_; }, "(tensor, cache_ttl) of every calc() so far, in order")
      .def("clear", [](RecordingDevice& self) { self.calls_.clear(); });
  m.def("evaluate_on", [](std::shared_ptr<teq::iEvaluator> self, RecordingDevice& device, ETensorsT targeted, ETensorsT ignored) {
    self->evaluate(device, to_set(targeted), to_set(ignored));
  }, py::arg("evaluator"), py::arg("device"), py::arg("targeted"), py::arg("ignored") = ETensorsT{},
  "iEvaluator::evaluate on a recording device (no computation): which functors are visited, in what order");
  // test support only (the counterpart of the reference tests' MockDeviceRef): hands a functor's holder a small HOST block so that
  // device_data() is non-null and the evaluators' "cannot ignore tensor without existing data" precondition can be met on a
  // machine without a GPU. Nothing ever reads the block; never call this on a graph that will be evaluated on the device.
  auto testing = m.def_submodule("testing", "hooks for the CPU-side mirrors of the reference's unit tests");
  testing.def("mock_data", [](ETensor t, size_t ttl) {
    struct HostMemory final : public eigen::iRuntimeMemory {
      void* allocate(size_t size) override { return std::malloc(size ? size : 1); }
      void deallocate(void* ptr, size_t) override { std::free(ptr); }
    };
    static eigen::RTMemptrT host = std::make_shared<HostMemory>();
    auto op = dynamic_cast<cuda::DevOp*>(&t->device());
    if (nullptr == op) global::fatalf("%s has no temporary holder to mock", t->to_string().c_str());
    op->ensure_buffer(ttl, host);
  }, py::arg("tensor"), py::arg("ttl") = 1);
  // The holder contract of SURVEY §8 rows a2-a7 / a16 (iEigen: data / odata / assign / valid_for / extend_life, Expirable TTL,
  // iRuntimeMemory) exercised the way internal/eigen/test/test_device.cpp and tenncor/eteq/test/test_functor.cpp do: a recording
  // allocator stands in for MockRuntimeMemory and a recording launcher for the op lambda, so no kernel runs.
  py::class_<CountingMemory, std::shared_ptr<CountingMemory>>(testing, "CountingMemory")
      .def(py::init<>())
      .def("log", [](CountingMemory& self) { return self.log_; }, "('allocate' | 'deallocate', bytes, pointer) in call order")
      .def("clear", [](CountingMemory& self) { self.log_.clear(); });
  py::class_<LaunchLog, std::shared_ptr<LaunchLog>>(testing, "LaunchLog")
      .def("calls", [](LaunchLog& self) { return self.calls_; }, "(out pointer, [argument pointers]) of every launch so far");
  testing.def("stub_launch", [](ETensor t) {
    auto op = dynamic_cast<cuda::DevOp*>(&t->device());
    if (nullptr == op) global::fatalf("%s has no temporary holder to stub", t->to_string().c_str());
    auto log = std::make_shared<LaunchLog>();
    op->set_launch([log](void* out, const std::vector<const void*>& in) {
      std::vector<uintptr_t> args;
      for (auto p : in) args.push_back((uintptr_t)p);
      log->calls_.push_back({(uintptr_t)out, args});
    });
    return log;
  }, py::arg("tensor"), "Replace the functor's kernel launcher by a recorder (never call on a graph that will be evaluated on the device)");
  testing.def("holder_kind", [](ETensor t) -> std::string {
    auto& dev = t->device();
    if (dynamic_cast<cuda::DevOp*>(&dev)) return "DevOp";
    if (dynamic_cast<cuda::DevRef*>(&dev)) return "DevRef";
    if (dynamic_cast<cuda::DevSrc*>(&dev)) return "DevSrc";
    if (dynamic_cast<cuda::DevAssign*>(&dev)) return "DevAssign";
    return "other";
  });
  testing.def("cache_init", [](ETensor t) {  // Functor::cache_init (functor.hpp:229-243): CacheEigen keeps the result across reads
    auto op = dynamic_cast<cuda::DevOp*>(&t->device());
    if (nullptr == op) global::fatalf("%s has no temporary holder to cache", t->to_string().c_str());
    op->pin();
  });
  testing.def("holder_ptr", [](ETensor t) { return (uintptr_t)t->device().device_data(); }, "device_data() of the holder: 0 before the first assign");
  testing.def("holder_assign", [](ETensor t, size_t ttl, std::shared_ptr<CountingMemory> memory) {
    eigen::RTMemptrT mem = memory;
    static_cast<eigen::iEigen&>(t->device()).assign(ttl, mem);
  }, py::arg("tensor"), py::arg("ttl"), py::arg("memory"));
  testing.def("holder_read", [](ETensor t) {
    uintptr_t p;
    { auto once = t->device().odata(); p = (uintptr_t)once.get(); }  // the Once's destructor is "one consumer done"
    return p;
  }, py::arg("tensor"), "Take odata() and release it: one consumer read (ticks the TTL)");
  testing.def("holder_valid_for", [](ETensor t, size_t ttl) { return static_cast<eigen::iEigen&>(t->device()).valid_for(ttl); });
  testing.def("holder_extend_life", [](ETensor t, size_t ttl) { static_cast<eigen::iEigen&>(t->device()).extend_life(ttl); });
  testing.def("device_calc", [](ETensor t, size_t cache_ttl, std::shared_ptr<CountingMemory> memory, size_t max_version) {
    cuda::Device dev(memory, max_version);  // eigen::Device::calc, internal/eigen/device.hpp:555-570
    dev.calc(*t, cache_ttl);
  }, py::arg("tensor"), py::arg("cache_ttl"), py::arg("memory"), py::arg("max_version") = std::numeric_limits<size_t>::max());
  m.def("set_eval", [](std::shared_ptr<teq::iEvaluator> eval) { teq::set_eval(std::move(eval)); },
        "Install an evaluator object in the context slot (teq::set_eval, internal/teq/evaluator.hpp:65)");

  // ---- hone: pre-evaluation rewrites (tenncor/hone/src/optimize.cpp)
  m.def("optimize", [](ETensorsT roots, bool fold_constants) {
    hone::Stats st;
    ETensorsT out = hone::optimize(std::move(roots), &st, fold_constants);
    py::dict d;
    d["functors_before"] = st.functors_before; d["functors_after"] = st.functors_after; d["merged"] = st.merged; d["folded"] = st.folded; d["rounds"] = st.rounds;
    return py::make_tuple(out, d);
  }, py::arg("roots"), py::arg("fold_constants") = true,
  "Merge structurally equal sub-graphs and fold constant functors (evaluated on the device); returns (new roots, stats)");
  m.def("fold_candidates", [](const ETensorsT& roots) { return hone::fold_candidates(roots); }, py::arg("roots"),
        "The functors constant folding would evaluate and replace by constants (top-most constant functors; IDENTITY never folds)");
  m.def("merge_dups", [](ETensorsT roots) {
    size_t n = 0;
    ETensorsT out = hone::merge_dups(std::move(roots), &n);
    return py::make_tuple(out, n);
  }, py::arg("roots"));

  // ---- back-end controls
  m.def("sync", [] { cuda::sync(); }, "Wait for all queued device work");
  m.def("shutdown", [] {
    if (tcr_stream() == nullptr) return;  // never initialised
    cuda::drop_all_plans();
    tcr_sync();
    tcr_prefetch_sync();
  }, "Destroy every cached plan (and its CUDA graph) and drain the device: call before the interpreter exits");
  m.def("launch_count", [] { return tcr_launch_count(); });
  m.def("sync_prefetch", [] { cuda::check(tcr_prefetch_sync(), "tcr_prefetch_sync"); }, "Wait for input copies started by EVariable.prefetch");
  m.def("set_matmul_precision", [](const std::string& p) {
    cuda::set_gemm_precision(p == "exact" ? TCR_GEMM_EXACT : p == "tf32" ? TCR_GEMM_TF32 : p == "3xtf32" ? TCR_GEMM_3XTF32 : -1);
  }, "fp32 MATMUL/CONTRACT precision: 'exact' (SIMT FMA), 'tf32' or '3xtf32' (tcgen05)");
  m.def("set_evaluator", [](const std::string& kind) {
    if (kind == "plan") teq::set_eval(std::make_shared<cuda::PlanEvaluator>());
    else if (kind == "node") teq::set_eval(std::make_shared<teq::Evaluator>());
    else global::fatalf("unknown evaluator `%s` (plan | node)", kind.c_str());
  }, "'plan': fused launch plan replayed as a CUDA graph (default); 'node': one kernel per functor like the reference");
  m.def("plan_stats", [] {
    auto s = cuda::last_plan_stats();
    py::dict d;
    d["nodes"] = s.nodes; d["steps"] = s.steps; d["launches_per_run"] = s.launches; d["graph"] = s.graph; d["plans_cached"] = s.cached;
    d["steps_run"] = s.steps_run; d["partial_runs"] = s.partial_runs;
    return d;
  });
  m.def("describe_plan", [](const ETensorsT& targets) {
    teq::TensSetT set;
    for (auto& t : targets) set.insert(t.get());
    return cuda::describe_plan(set);
  }, "Launch steps the planned evaluator lowers `targets` to (host-side lowering only; needs no device)");
  m.def("profile_plan", [](int repeats) {
    py::list out;
    for (auto& t : cuda::profile_last_plan(repeats)) {
      py::dict d;
      d["what"] = t.what; d["shape"] = t.shape; d["ms"] = t.ms; d["bytes"] = t.bytes;
      out.append(d);
    }
    return out;
  }, py::arg("repeats") = 5, "Per-step CUDA-event timings of the last evaluated plan");
  m.def("arena_stats", [] {
    size_t a = 0, b = 0, c = 0;
    tcr_arena_stats(&a, &b, &c);
    py::dict d;
    d["bytes_in_use"] = a; d["bytes_reserved"] = b; d["device_mallocs"] = c;
    return d;
  });

  // ---- api
  auto api = m.def_submodule("api", "TenncorAPI (cfg/tenncor/core.yml)");
  UN("abs", ABS); UN("neg", NEG); UN("sin", SIN); UN("cos", COS); UN("tan", TAN); UN("exp", EXP); UN("log", LOG);
  UN("sqrt", SQRT); UN("round", ROUND); UN("sigmoid", SIGMOID); UN("tanh", TANH); UN("square", SQUARE); UN("cube", CUBE);
  BIN("pow", POW); BIN("add", ADD); BIN("sub", SUB); BIN("mul", MUL); BIN("div", DIV); BIN("eq", EQ); BIN("neq", NEQ);
  BIN("lt", LT); BIN("gt", GT); BIN("min", MIN); BIN("max", MAX);
  api.def("cast", [](const ETensor& x, py::object dtype) { return tenncor::cast(x, parse_dtype(dtype)); });
  api.def("assign", [](const ETensor& t, const ETensor& s) { return tenncor::assign(as_var(t), s); });
  api.def("assign_add", [](const ETensor& t, const ETensor& s) { return tenncor::assign_add(as_var(t), s); });
  api.def("assign_sub", [](const ETensor& t, const ETensor& s) { return tenncor::assign_sub(as_var(t), s); });
  api.def("assign_mul", [](const ETensor& t, const ETensor& s) { return tenncor::assign_mul(as_var(t), s); });
  api.def("assign_div", [](const ETensor& t, const ETensor& s) { return tenncor::assign_div(as_var(t), s); });
  api.def("identity", &tenncor::identity, py::arg("input"), py::arg("execute_in_parallel") = ETensorsT{});
  api.def("if_then_else", &tenncor::if_then_else);
  api.def("reverse", &tenncor::reverse);
  api.def("permute", &tenncor::permute);
  api.def("extend", [](const ETensor& a, const DimsT& bcast) { return tenncor::extend(a, bcast); });
  api.def("extend", [](const ETensor& a, RankT offset, const DimsT& xlist) { return tenncor::extend(a, offset, xlist); });
  api.def("extend_like", &tenncor::extend_like);
  api.def("concat", [](const ETensor& l, const ETensor& r, RankT axis) { return tenncor::concat(l, r, axis); });
  api.def("concat", [](const ETensorsT& args, RankT axis) { return tenncor::concat(args, axis); });
  api.def("reshape", [](const ETensor& a, std::vector<size_t> pyshape) { return tenncor::reshape(a, p2cshape(pyshape)); });
#define RED(NAME, OP)                                                                                                          \
  api.def(NAME, [](const ETensor& t, std::set<RankT> dims) { return tenncor::reduce(egen::OP, t, dims); });                   \
  api.def(NAME, [](const ETensor& t, RankT offset, RankT ndims) { return tenncor::reduce(egen::OP, t, offset, ndims); },      \
          py::arg("tens"), py::arg("offset") = 0, py::arg("ndims") = rank_cap);                                                \
  api.def(NAME "_1d", [](const ETensor& t, RankT d) { return tenncor::reduce_1d(egen::OP, t, d); })
  RED("reduce_sum", REDUCE_SUM); RED("reduce_prod", REDUCE_PROD); RED("reduce_min", REDUCE_MIN); RED("reduce_max", REDUCE_MAX);
  api.def("argmax", &tenncor::argmax, py::arg("tens"), py::arg("return_dim") = 8);
  api.def("n_elems", &tenncor::n_elems);
  api.def("n_dims", &tenncor::n_dims);
  api.def("slice", [](const ETensor& a, eigen::PairVecT<DimT> extents) { return tenncor::slice(a, extents); });
  api.def("slice", [](const ETensor& a, DimT offset, DimT extent, RankT dim) { return tenncor::slice(a, offset, extent, dim); });
  api.def("pad", [](const ETensor& a, eigen::PairVecT<DimT> p) { return tenncor::pad(a, p); });
  api.def("pad", [](const ETensor& a, tenncor::DimPairsT p, RankT dim) { return tenncor::pad(a, p, dim); });
  api.def("stride", &tenncor::stride);
  api.def("scatter", [](const ETensor& a, std::vector<size_t> pyshape, const DimsT& incrs) { return tenncor::scatter(a, p2cshape(pyshape), incrs); });
  api.def("contract", &tenncor::contract, py::arg("a"), py::arg("b"), py::arg("dims") = eigen::PairVecT<RankT>{{0, 1}});
  api.def("matmul", &tenncor::matmul);
  api.def("convolution", &tenncor::convolution);
  api.def("transpose", &tenncor::transpose);
  api.def("reduce_mean", &tenncor::reduce_mean);
  api.def("reduce_mean_1d", &tenncor::reduce_mean_1d);
  api.def("reduce_variance", &tenncor::reduce_variance);
  api.def("reduce_variance_1d", &tenncor::reduce_variance_1d);
  api.def("reduce_l2norm", &tenncor::reduce_l2norm, py::arg("arg"), py::arg("offset") = 0, py::arg("ndims") = rank_cap);
  api.def("reduce_l2norm_1d", &tenncor::reduce_l2norm_1d);
  api.def("clip_by_range", &tenncor::clip_by_range);
  api.def("clip_by_l2norm", &tenncor::clip_by_l2norm);
  api.def("sum", &tenncor::sum);
  api.def("prod", &tenncor::prod);
  api.def("softmax", &tenncor::softmax, py::arg("arg"), py::arg("offset") = 0, py::arg("ndims") = rank_cap);
  api.def("relu", &tenncor::relu);
  api.def("softplus", &tenncor::softplus);
  api.def("sign", &tenncor::sign);

  auto rnd = api.def_submodule("random");
  rnd.def("rand_unif", &tenncor::random::rand_unif);
  rnd.def("rand_binom_one", &tenncor::random::rand_binom_one);

  py::class_<InitHolder>(m, "Initializer")
      .def("__call__", [](InitHolder& self, std::vector<size_t> pyshape, const std::string& label) { return self.f(p2cshape(pyshape), label); },
           py::arg("shape"), py::arg("label") = "");
  auto ini = api.def_submodule("init");
  ini.def("random_normal", [](double mean, double stddev) { return InitHolder{tenncor::init::random_normal(mean, stddev)}; }, py::arg("mean") = 0, py::arg("stddev") = 1);
  ini.def("random_uniform", [](double lo, double hi) { return InitHolder{tenncor::init::random_uniform(lo, hi)}; }, py::arg("minval") = -0.05, py::arg("maxval") = 0.05);
  ini.def("zeros", [] { return InitHolder{tenncor::init::zeros()}; });
  ini.def("ones", [] { return InitHolder{tenncor::init::ones()}; });
  ini.def("constants", [](double v) { return InitHolder{tenncor::init::constants(v)}; });
  ini.def("xavier_uniform", [](double f) { return InitHolder{tenncor::init::xavier_uniform(f)}; }, py::arg("factor") = 1);
  ini.def("xavier_normal", [](double f) { return InitHolder{tenncor::init::xavier_normal(f)}; }, py::arg("factor") = 1);
  ini.def("glorot_uniform", [](double f) { return InitHolder{tenncor::init::xavier_uniform(f)}; }, py::arg("factor") = 1);
  ini.def("truncated_normal", [](double mean, double stddev) { return InitHolder{tenncor::init::truncated_normal(mean, stddev)}; }, py::arg("mean") = 0, py::arg("stddev") = 1);
  ini.def("identity", [](double gain) { return InitHolder{tenncor::init::identity(gain)}; }, py::arg("gain") = 1);
  ini.def("variance_scaling", [](double factor, py::object sfactor) {
    std::function<double(Shape)> f;
    if (!sfactor.is_none()) {
      py::function pf = sfactor.cast<py::function>();
      f = [pf](Shape shape) { DimsT ps = c2pshape(shape); return pf(std::vector<size_t>(ps.begin(), ps.end())).cast<double>(); };
    }
    return InitHolder{tenncor::init::variance_scaling(factor, f)};
  }, py::arg("factor"), py::arg("shape_factor") = py::none());
  ini.def("glorot_normal", [](double f) { return InitHolder{tenncor::init::xavier_normal(f)}; }, py::arg("factor") = 1);

  auto nn = api.def_submodule("nn");
  nn.def("fully_connect", &tenncor::nn::fully_connect, py::arg("lefts"), py::arg("rights"), py::arg("bias") = ETensor(),
         py::arg("dims") = eigen::PairVecT<RankT>{{0, 1}});
  nn.def("conv2d", &tenncor::nn::conv2d, py::arg("image"), py::arg("kernel"), py::arg("bias") = ETensor(),
         py::arg("zero_paddings") = std::pair<tenncor::DimPairsT, tenncor::DimPairsT>{{0, 0}, {0, 0}});
  nn.def("dropout", &tenncor::nn::dropout, py::arg("input"), py::arg("drop_rate"));
  nn.def("dropout", [](const ETensor& input, double drop_rate) {  // nn.yml:99-109: the rate becomes a scalar variable of the input's type
    return tenncor::nn::dropout(input, eteq::make_variable_scalar(drop_rate, Shape(), "drop_rate", (egen::_GENERATED_DTYPE)input->get_meta().type_code()));
  }, py::arg("input"), py::arg("drop_rate"));
  nn.def("batch_normalization", [](const ETensor& input, py::object offset, py::object scale, py::object eps, py::object get_mean, py::object get_variance) {
    auto unary = [](py::object f) { return f.is_none() ? layr::UnaryF() : f.cast<layr::UnaryF>(); };
    if (py::isinstance<py::float_>(offset) || py::isinstance<py::int_>(offset))
      return tenncor::nn::batch_normalization(input, offset.cast<double>(), scale.cast<double>(), eps.is_none() ? -1. : eps.cast<double>(), unary(get_mean), unary(get_variance));
    ETensor e = eps.is_none() ? eteq::make_constant_like(1e-7, input) : eps.cast<ETensor>();
    return tenncor::nn::batch_normalization(input, offset.cast<ETensor>(), scale.cast<ETensor>(), e, unary(get_mean), unary(get_variance));
  }, py::arg("input"), py::arg("offset") = 0., py::arg("scale") = 1., py::arg("eps") = py::none(), py::arg("get_mean") = py::none(), py::arg("get_variance") = py::none());
  nn.def("mean_pool2d", &tenncor::nn::mean_pool2d, py::arg("arg"), py::arg("dims") = std::pair<RankT, RankT>{0, 1});
  nn.def("max_pool2d", &tenncor::nn::max_pool2d, py::arg("arg"), py::arg("dims") = std::pair<RankT, RankT>{0, 1});

  auto wrap_init = [](py::object f) -> layr::InitF {
    if (f.is_none()) return layr::InitF();
    if (py::isinstance<InitHolder>(f)) return f.cast<InitHolder&>().f;
    // a python callable (numpy_shape, label) -> EVariable
    py::function pf = f.cast<py::function>();
    return [pf](Shape shape, std::string label) {
      DimsT ps = c2pshape(shape);
      return as_var(pf(std::vector<size_t>(ps.begin(), ps.end()), label).cast<ETensor>());
    };
  };
  auto lay = api.def_submodule("layer");
  lay.def("bind", [](layr::UnaryF unary, std::vector<size_t> inshape) { return tenncor::layer::bind(unary, p2cshape(inshape)); },
          py::arg("unary"), py::arg("inshape") = std::vector<size_t>{});
  lay.def("link", &tenncor::layer::link, py::arg("layers"), py::arg("input") = ETensor());
  lay.def("dense", [wrap_init](std::vector<size_t> inshape, std::vector<size_t> hidden_dims, py::object kinit, py::object binit, bool with_bias, py::object dtype) {
    DimsT hd(hidden_dims.rbegin(), hidden_dims.rend());
    return tenncor::layer::dense(p2cshape(inshape), hd, wrap_init(kinit), wrap_init(binit), with_bias, {{0, 1}}, parse_dtype(dtype));
  }, py::arg("inshape"), py::arg("hidden_dims"), py::arg("kernel_init") = py::none(), py::arg("bias_init") = py::none(), py::arg("with_bias") = true,
          py::arg("dtype") = py::none());
  lay.def("dense_on", [](const ETensor& input, const ETensor& kernel, const ETensor& bias) { return tenncor::layer::dense(input, kernel, bias); },
          py::arg("input"), py::arg("kernel"), py::arg("bias") = ETensor());
  using ZeroPadT = std::pair<tenncor::DimPairsT, tenncor::DimPairsT>;
  lay.def("conv2d", [wrap_init](tenncor::DimPairsT kernel_hw, DimT in_ncol, DimT out_ncol, py::object kinit, py::object binit, ZeroPadT zero_padding,
                                bool with_bias, py::object dtype) {
    return tenncor::layer::conv2d(kernel_hw, in_ncol, out_ncol, wrap_init(kinit), wrap_init(binit), zero_padding, with_bias, parse_dtype(dtype));
  }, py::arg("kernel_hw"), py::arg("in_ncol"), py::arg("out_ncol"), py::arg("kernel_init") = py::none(), py::arg("bias_init") = py::none(),
          py::arg("zero_padding") = ZeroPadT{{0, 0}, {0, 0}}, py::arg("with_bias") = true, py::arg("dtype") = py::none());
  lay.def("conv2d", [wrap_init](const ETensor& input, DimT out_ncol, tenncor::DimPairsT kernel_hw, py::object kinit, py::object binit, py::object padding,
                                bool with_bias) {  // layer.yml:159-251: on an existing image, padding = "valid" | "same" | ((x0, x1), (y0, y1))
    if (py::isinstance<py::str>(padding))
      return tenncor::layer::conv2d(input, out_ncol, kernel_hw, wrap_init(kinit), wrap_init(binit), padding.cast<std::string>(), with_bias);
    return tenncor::layer::conv2d(input, out_ncol, kernel_hw, wrap_init(kinit), wrap_init(binit), padding.cast<ZeroPadT>(), with_bias);
  }, py::arg("input"), py::arg("out_ncol"), py::arg("kernel_hw"), py::arg("kernel_init") = py::none(), py::arg("bias_init") = py::none(),
          py::arg("padding") = "valid", py::arg("with_bias") = true);
  lay.def("rnn", [wrap_init](DimT indim, DimT hidden_dim, layr::UnaryF activation, DimT nseq, py::object kinit, py::object binit, RankT seq_dim, py::object dtype) {
    return tenncor::layer::rnn(indim, hidden_dim, activation, nseq, wrap_init(kinit), wrap_init(binit), seq_dim, true, parse_dtype(dtype));
  }, py::arg("indim"), py::arg("hidden_dim"), py::arg("activation"), py::arg("nseq"), py::arg("kernel_init") = py::none(),
          py::arg("bias_init") = py::none(), py::arg("seq_dim") = 1, py::arg("dtype") = py::none());
  lay.def("lstm", [wrap_init](std::vector<size_t> inshape, DimT hidden_dim, DimT nseq, py::object kinit, py::object binit, RankT seq_dim) {
    return tenncor::layer::lstm(p2cshape(inshape), hidden_dim, nseq, wrap_init(kinit), wrap_init(binit), seq_dim);
  }, py::arg("inshape"), py::arg("hidden_dim"), py::arg("nseq"), py::arg("kernel_init") = py::none(), py::arg("bias_init") = py::none(),
          py::arg("seq_dim") = 1);
  lay.def("gru", [wrap_init](std::vector<size_t> inshape, DimT hidden_dim, DimT nseq, py::object kinit, py::object binit, RankT seq_dim) {
    return tenncor::layer::gru(p2cshape(inshape), hidden_dim, nseq, wrap_init(kinit), wrap_init(binit), seq_dim);
  }, py::arg("inshape"), py::arg("hidden_dim"), py::arg("nseq"), py::arg("kernel_init") = py::none(), py::arg("bias_init") = py::none(),
          py::arg("seq_dim") = 1);
  lay.def("rnn_on", [](const ETensor& input, const ETensor& init_state, const ETensor& cell, layr::UnaryF activation, RankT seq_dim) {
    return tenncor::layer::rnn(input, init_state, cell, activation, seq_dim);
  }, py::arg("input"), py::arg("init_state"), py::arg("cell"), py::arg("activation"), py::arg("seq_dim") = 1);
  py::class_<layr::RBMLayer>(m, "RBMLayer")
      .def(py::init([](ETensor fwd, ETensor bwd) { return layr::RBMLayer{fwd, bwd}; }), py::arg("fwd"), py::arg("bwd"))
      .def("connect", &layr::RBMLayer::connect)
      .def("backward_connect", &layr::RBMLayer::backward_connect)
      .def("deep_clone", &layr::RBMLayer::deep_clone)
      .def("fwd", [](layr::RBMLayer& self) { return self.fwd_; })
      .def("bwd", [](layr::RBMLayer& self) { return self.bwd_; });
  lay.def("dropout", [](const ETensor& input, py::object drop_rate, ETensor training) {
    ETensor rate = (py::isinstance<py::float_>(drop_rate) || py::isinstance<py::int_>(drop_rate))
                       ? ETensor(eteq::make_variable_scalar(drop_rate.cast<double>(), Shape(), "drop_rate", (egen::_GENERATED_DTYPE)input->get_meta().type_code()))
                       : drop_rate.cast<ETensor>();
    return tenncor::layer::dropout(input, rate, training);
  }, py::arg("input"), py::arg("drop_rate"), py::arg("training") = ETensor());
  lay.def("batch_normalization", [wrap_init](ETensor input, py::object offset, py::object scale, py::object eps, ETensor training, py::object momentum,
                                            py::object moving_mean_init, py::object moving_var_init, RankT axis) {
    auto as_tensor = [&input](py::object o, double dflt) -> ETensor {
      if (o.is_none()) return eteq::make_constant_like(dflt, input);
      if (py::isinstance<py::float_>(o) || py::isinstance<py::int_>(o)) return eteq::make_constant_like(o.cast<double>(), input);
      return o.cast<ETensor>();
    };
    const bool is_double = (egen::_GENERATED_DTYPE)input->get_meta().type_code() == egen::DOUBLE;
    ETensor m = momentum.is_none() ? ETensor() : as_tensor(momentum, 0.99);
    return tenncor::layer::batch_normalization(input, as_tensor(offset, 0), as_tensor(scale, 1),
                                               as_tensor(eps, is_double ? std::numeric_limits<double>::epsilon() : std::numeric_limits<float>::epsilon()),
                                               training, m, wrap_init(moving_mean_init), wrap_init(moving_var_init), axis);
  }, py::arg("input"), py::arg("offset") = 0., py::arg("scale") = 1., py::arg("eps") = py::none(), py::arg("training") = ETensor(),
          py::arg("momentum") = py::none(), py::arg("moving_mean_init") = py::none(), py::arg("moving_var_init") = py::none(), py::arg("axis") = teq::rank_cap);
  lay.def("rbm", [wrap_init](DimT nvisible, DimT nhidden, py::object kinit, py::object binit, bool with_bias) {
    return tenncor::layer::rbm(nvisible, nhidden, wrap_init(kinit), wrap_init(binit), with_bias);
  }, py::arg("nvisible"), py::arg("nhidden"), py::arg("kernel_init") = py::none(), py::arg("bias_init") = py::none(), py::arg("with_bias") = true);

  auto los = api.def_submodule("loss");
  los.def("sqr_diff", &tenncor::loss::sqr_diff);
  los.def("mean_squared", &tenncor::loss::mean_squared, py::arg("target"), py::arg("input"), py::arg("axis") = rank_cap);
  los.def("cross_entropy", &tenncor::loss::cross_entropy, py::arg("target"), py::arg("input"), py::arg("eps") = std::numeric_limits<float>::epsilon());

  auto to_vars = [](const ETensorsT& ts) {
    eteq::VarptrsT vars;
    for (auto& t : ts) vars.push_back(as_var(t));
    return vars;
  };
  auto apx = api.def_submodule("approx");
  const double feps = std::numeric_limits<float>::epsilon();
  apx.def("sgd", [to_vars](const ETensor& e, const ETensorsT& v, double lr) { return tenncor::approx::sgd(e, to_vars(v), lr); },
          py::arg("error"), py::arg("variables"), py::arg("learning_rate") = 0.5);
  apx.def("adagrad", [to_vars](const ETensor& e, const ETensorsT& v, double lr, double eps) { return tenncor::approx::adagrad(e, to_vars(v), lr, eps); },
          py::arg("error"), py::arg("variables"), py::arg("learning_rate") = 0.5, py::arg("epsilon") = feps);
  apx.def("adam", [to_vars](const ETensor& e, const ETensorsT& v, double sr, double d1, double d2, double eps) {
    return tenncor::approx::adam(e, to_vars(v), sr, d1, d2, eps);
  }, py::arg("error"), py::arg("variables"), py::arg("step_rate") = 0.001, py::arg("decay1") = 0.9, py::arg("decay2") = 0.999, py::arg("epsilon") = feps);
  apx.def("adadelta", [to_vars](const ETensor& e, const ETensorsT& v, double sr, double decay, double offset, double eps) {
    return tenncor::approx::adadelta(e, to_vars(v), sr, decay, offset, eps);
  }, py::arg("error"), py::arg("variables"), py::arg("step_rate") = 1, py::arg("decay") = 0.9, py::arg("offset") = 0.0001, py::arg("epsilon") = feps);
  apx.def("rms_momentum", [to_vars](const ETensor& e, const ETensorsT& v, double lr, double discount, double eps, py::object apply) {
    layr::UnaryF f;
    if (!apply.is_none()) f = apply.cast<layr::UnaryF>();
    return tenncor::approx::rms_momentum(e, to_vars(v), lr, discount, eps, f);
  }, py::arg("error"), py::arg("variables"), py::arg("learning_rate") = 0.5, py::arg("discount_factor") = 0.99, py::arg("epsilon") = feps,
          py::arg("apply") = py::none());

  // trainer::DBNTrainer (tenncor/python/layr_ext.cpp:30-70)
  py::class_<trainer::DBNTrainer>(m, "DBNTrainer")
      .def(py::init<const std::vector<layr::RBMLayer>&, ETensor, RankT, DimT, double, double, size_t, double, double>(), py::arg("rbms"), py::arg("dense"),
           py::arg("softmax_dim"), py::arg("batch_size"), py::arg("pretrain_lr") = 0.1, py::arg("train_lr") = 0.1, py::arg("cdk") = 10,
           py::arg("l2_reg") = 0., py::arg("lr_scaling") = 0.95)
      .def("pretrain", [](trainer::DBNTrainer& self, py::array x, size_t nepochs, py::object logger) {
        Shape shape;
        egen::_GENERATED_DTYPE dtype;
        py::array arr = normalise(x, shape, dtype);
        if (shape.n_elems() != self.trainx_->shape().n_elems()) global::fatalf("pretrain input has %d elements, the trainer expects %d", (int)shape.n_elems(), (int)self.trainx_->shape().n_elems());
        std::function<void(size_t, size_t)> log;
        if (!logger.is_none()) log = [logger](size_t epoch, size_t layer) { logger(epoch, layer); };
        self.pretrain(arr.data(), dtype, nepochs, log);
      }, py::arg("x"), py::arg("nepochs") = 100, py::arg("logger") = py::none())
      .def("finetune", [](trainer::DBNTrainer& self, py::array x, py::array y, size_t nepochs, py::object logger) {
        Shape xs, ys;
        egen::_GENERATED_DTYPE xd, yd;
        py::array xa = normalise(x, xs, xd), ya = normalise(y, ys, yd);
        if (xd != yd) global::fatal("finetune input and labels need the same dtype");
        if (xs.n_elems() != self.trainx_->shape().n_elems() || ys.n_elems() != self.trainy_->shape().n_elems()) global::fatal("finetune data does not match the trainer's batch shape");
        std::function<void(size_t)> log;
        if (!logger.is_none()) log = [logger](size_t epoch) { logger(epoch); };
        self.finetune(xa.data(), ya.data(), xd, nepochs, log);
      }, py::arg("x"), py::arg("y"), py::arg("nepochs") = 100, py::arg("logger") = py::none())
      .def("reconstruction_cost", &trainer::DBNTrainer::reconstruction_cost)
      .def("training_cost", &trainer::DBNTrainer::training_cost)
      .def("sample_pipes", [](trainer::DBNTrainer& self) { return self.sample_pipes_; })
      .def("update_graphs", [](trainer::DBNTrainer& self) {
        ETensorsT out;
        for (auto& layer : self.rupdates_) out.insert(out.end(), layer.begin(), layer.end());
        out.push_back(self.tupdate_);
        return out;
      }, "every per-layer CD update followed by the logistic-layer update (for inspection / describe_plan)");

  m.def("rbm_train", [](const layr::RBMLayer& model, ETensor visible, double lr, double discount, size_t cdk) {
    return trainer::rbm(model, visible, lr, discount, {}, cdk);
  }, py::arg("model"), py::arg("visible"), py::arg("learning_rate"), py::arg("discount_factor"), py::arg("cdk") = 1);

  // ---- data parallel
  auto dpm = m.def_submodule("dp", "batch-sharded data parallelism over NCCL (dp.hpp)");
  dpm.def("unique_id", [] { return py::bytes(dp::unique_id()); });
  dpm.def("init", [](int rank, int nranks, py::bytes id, bool mean_reduce) { dp::init(rank, nranks, std::string(id), mean_reduce); },
          py::arg("rank"), py::arg("nranks"), py::arg("id") = py::bytes(""), py::arg("mean_reduce") = true);
  dpm.def("set_mean_reduce", &dp::set_mean_reduce, py::arg("mean_reduce"));
  dpm.def("shutdown", &dp::shutdown);
  dpm.def("rank", &dp::rank);
  dpm.def("size", &dp::size);
  dpm.def("shard", &dp::shard);
}
