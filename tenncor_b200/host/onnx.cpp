// onnx.cpp — see onnx.hpp. Reference: internal/onnx/{save,load,marshal}.hpp, src/{load,marshal}.cpp,
// tenncor/serial/serialize.hpp, tenncor/serial/src/serialize.cpp, tenncor/src/serial.cpp.
#include "onnx.hpp"

#include <cstring>
#include <fstream>
#include <random>
#include <sstream>

#include "api.hpp"

namespace onnx {

using namespace teq;

// ================================================================ protobuf wire format
namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  Reader(const std::string& s) : p((const uint8_t*)s.data()), end((const uint8_t*)s.data() + s.size()) {}
  Reader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t out = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      if (p >= end) global::throw_err("onnx: truncated varint");
      const uint8_t c = *p++;
      out |= (uint64_t)(c & 0x7f) << shift;
      if (!(c & 0x80)) return out;
    }
    global::throw_err("onnx: varint too long");
    return 0;
  }
  Reader sub() {
    const uint64_t n = varint();
    if ((uint64_t)(end - p) < n) global::throw_err("onnx: truncated length-delimited field");
    Reader r(p, p + n);
    p += n;
    return r;
  }
  std::string bytes() {
    Reader r = sub();
    return std::string((const char*)r.p, (const char*)r.end);
  }
  template <typename T>
  T fixed() {
    if ((size_t)(end - p) < sizeof(T)) global::throw_err("onnx: truncated fixed-width field");
    T v;
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  void skip(int wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: fixed<uint64_t>(); break;
      case 2: sub(); break;
      case 5: fixed<uint32_t>(); break;
      default: global::throw_errf("onnx: unsupported wire type %d", wire);
    }
  }
};

// repeated scalar: packed (wire 2) or one element per tag
template <typename T, typename F>
void read_repeated(Reader& r, int wire, std::vector<T>& out, int elem_wire, F one) {
  if (wire == 2 && elem_wire != 2) {
    Reader s = r.sub();
    while (!s.done()) out.push_back(one(s));
  } else {
    out.push_back(one(r));
  }
}

struct Writer {
  std::string out;
  void varint(uint64_t v) {
    while (v >= 0x80) {
      out.push_back((char)((v & 0x7f) | 0x80));
      v >>= 7;
    }
    out.push_back((char)v);
  }
  void tag(int field, int wire) { varint(((uint64_t)field << 3) | (uint64_t)wire); }
  void str(int field, const std::string& s) {
    tag(field, 2);
    varint(s.size());
    out += s;
  }
  void str_if(int field, const std::string& s) {
    if (!s.empty()) str(field, s);
  }
  void i64_if(int field, int64_t v) {
    if (v != 0) {
      tag(field, 0);
      varint((uint64_t)v);
    }
  }
  template <typename T>
  void packed_varint(int field, const std::vector<T>& v) {
    if (v.empty()) return;
    Writer w;
    for (T e : v) w.varint((uint64_t)(int64_t)e);
    str(field, w.out);
  }
  template <typename T>
  void packed_fixed(int field, const std::vector<T>& v) {
    if (v.empty()) return;
    tag(field, 2);
    varint(v.size() * sizeof(T));
    out.append((const char*)v.data(), v.size() * sizeof(T));
  }
};

void parse_graph(GraphProto& g, Reader r);

// TensorProto: dims=1 data_type=2 float_data=4 int32_data=5 int64_data=7 name=8 raw_data=9 double_data=10 uint64_data=11 (onnx.proto:353-440)
void parse_tensor(TensorProto& t, Reader r) {
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    switch (field) {
      case 1: read_repeated(r, wire, t.dims, 0, [](Reader& s) { return (int64_t)s.varint(); }); break;
      case 2: t.data_type = (int32_t)r.varint(); break;
      case 4: read_repeated(r, wire, t.float_data, 5, [](Reader& s) { return s.fixed<float>(); }); break;
      case 5: read_repeated(r, wire, t.int32_data, 0, [](Reader& s) { return (int32_t)s.varint(); }); break;
      case 7: read_repeated(r, wire, t.int64_data, 0, [](Reader& s) { return (int64_t)s.varint(); }); break;
      case 8: t.name = r.bytes(); break;
      case 9: t.raw_data = r.bytes(); break;
      case 10: read_repeated(r, wire, t.double_data, 1, [](Reader& s) { return s.fixed<double>(); }); break;
      case 11: read_repeated(r, wire, t.uint64_data, 0, [](Reader& s) { return (uint64_t)s.varint(); }); break;
      default: r.skip(wire);
    }
  }
}

void write_tensor(Writer& w, const TensorProto& t) {
  w.packed_varint(1, t.dims);
  w.i64_if(2, t.data_type);
  w.packed_fixed(4, t.float_data);
  w.packed_varint(5, t.int32_data);
  w.packed_varint(7, t.int64_data);
  w.str_if(8, t.name);
  w.str_if(9, t.raw_data);
  w.packed_fixed(10, t.double_data);
  w.packed_varint(11, t.uint64_data);
}

// AttributeProto: name=1 f=2 i=3 s=4 t=5 g=6 floats=7 ints=8 strings=9 tensors=10 type=20 (onnx.proto:124-158)
void parse_attr(AttributeProto& a, Reader r) {
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    switch (field) {
      case 1: a.name = r.bytes(); break;
      case 2: a.f = r.fixed<float>(); break;
      case 3: a.i = (int64_t)r.varint(); break;
      case 4: a.s = r.bytes(); break;
      case 5: parse_tensor(a.t, r.sub()); break;
      case 6: a.g = std::make_shared<GraphProto>(); parse_graph(*a.g, r.sub()); break;
      case 7: read_repeated(r, wire, a.floats, 5, [](Reader& s) { return s.fixed<float>(); }); break;
      case 8: read_repeated(r, wire, a.ints, 0, [](Reader& s) { return (int64_t)s.varint(); }); break;
      case 9: a.strings.push_back(r.bytes()); break;
      case 10: a.tensors.emplace_back(); parse_tensor(a.tensors.back(), r.sub()); break;
      case 20: a.type = (int32_t)r.varint(); break;
      default: r.skip(wire);
    }
  }
}

std::string graph_bytes(const GraphProto& g);

void write_attr(Writer& w, const AttributeProto& a) {
  w.str_if(1, a.name);
  if (a.type == ATTR_FLOAT) {
    w.tag(2, 5);
    w.out.append((const char*)&a.f, 4);
  }
  if (a.type == ATTR_INT) w.i64_if(3, a.i);
  if (a.type == ATTR_STRING) w.str_if(4, a.s);
  if (a.type == ATTR_TENSOR) {
    Writer t;
    write_tensor(t, a.t);
    w.str(5, t.out);
  }
  if (a.type == ATTR_GRAPH && a.g) w.str(6, graph_bytes(*a.g));
  w.packed_fixed(7, a.floats);
  w.packed_varint(8, a.ints);
  for (auto& s : a.strings) w.str(9, s);
  for (auto& t : a.tensors) {
    Writer tw;
    write_tensor(tw, t);
    w.str(10, tw.out);
  }
  w.i64_if(20, a.type);
}

// NodeProto: input=1 output=2 name=3 op_type=4 attribute=5 (onnx.proto:181-194)
void parse_node(NodeProto& n, Reader r) {
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    switch (field) {
      case 1: n.input.push_back(r.bytes()); break;
      case 2: n.output.push_back(r.bytes()); break;
      case 3: n.name = r.bytes(); break;
      case 4: n.op_type = r.bytes(); break;
      case 5: n.attribute.emplace_back(); parse_attr(n.attribute.back(), r.sub()); break;
      default: r.skip(wire);
    }
  }
}

// ValueInfoProto: name=1 type=2 -> TypeProto.tensor_type=1 -> {elem_type=1, shape=2 -> dim=1 -> dim_value=1} (onnx.proto:165-168,504-553)
void parse_value_info(ValueInfoProto& v, Reader r) {
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 1) v.name = r.bytes();
    else if (field == 2) {
      Reader type = r.sub();
      while (!type.done()) {
        const uint64_t k2 = type.varint();
        if ((k2 >> 3) != 1) { type.skip((int)(k2 & 7)); continue; }
        Reader tt = type.sub();
        while (!tt.done()) {
          const uint64_t k3 = tt.varint();
          if ((k3 >> 3) == 1) v.elem_type = (int32_t)tt.varint();
          else if ((k3 >> 3) == 2) {
            Reader shape = tt.sub();
            while (!shape.done()) {
              const uint64_t k4 = shape.varint();
              if ((k4 >> 3) != 1) { shape.skip((int)(k4 & 7)); continue; }
              Reader dim = shape.sub();
              int64_t value = 0;
              while (!dim.done()) {
                const uint64_t k5 = dim.varint();
                if ((k5 >> 3) == 1) value = (int64_t)dim.varint();
                else dim.skip((int)(k5 & 7));
              }
              v.dims.push_back(value);
            }
          } else tt.skip((int)(k3 & 7));
        }
      }
    } else r.skip(wire);
  }
}

void write_value_info(Writer& w, const ValueInfoProto& v) {
  w.str_if(1, v.name);
  Writer shape;
  for (int64_t d : v.dims) {
    Writer dim;
    dim.i64_if(1, d);
    shape.str(1, dim.out);
  }
  Writer tt;
  tt.i64_if(1, v.elem_type);
  if (!v.dims.empty()) tt.str(2, shape.out);
  Writer type;
  type.str(1, tt.out);
  w.str(2, type.out);
}

// TensorAnnotation: tensor_name=1 quant_parameter_tensor_names=2 -> StringStringEntryProto{key=1,value=2} (onnx.proto:254-265)
void parse_annotation(TensorAnnotation& a, Reader r) {
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 1) a.tensor_name = r.bytes();
    else if (field == 2) {
      Reader e = r.sub();
      std::pair<std::string, std::string> kv;
      while (!e.done()) {
        const uint64_t k2 = e.varint();
        if ((k2 >> 3) == 1) kv.first = e.bytes();
        else if ((k2 >> 3) == 2) kv.second = e.bytes();
        else e.skip((int)(k2 & 7));
      }
      a.quant_parameter_tensor_names.push_back(kv);
    } else r.skip(wire);
  }
}

// GraphProto: node=1 name=2 initializer=5 input=11 output=12 quantization_annotation=14 (onnx.proto:278-306)
void parse_graph(GraphProto& g, Reader r) {
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    switch (field) {
      case 1: g.node.emplace_back(); parse_node(g.node.back(), r.sub()); break;
      case 2: g.name = r.bytes(); break;
      case 5: g.initializer.emplace_back(); parse_tensor(g.initializer.back(), r.sub()); break;
      case 11: g.input.emplace_back(); parse_value_info(g.input.back(), r.sub()); break;
      case 12: g.output.emplace_back(); parse_value_info(g.output.back(), r.sub()); break;
      case 14: g.quantization_annotation.emplace_back(); parse_annotation(g.quantization_annotation.back(), r.sub()); break;
      default: r.skip(wire);
    }
  }
}

std::string graph_bytes(const GraphProto& g) {
  Writer w;
  for (auto& n : g.node) {
    Writer nw;
    for (auto& s : n.input) nw.str(1, s);
    for (auto& s : n.output) nw.str(2, s);
    nw.str_if(3, n.name);
    nw.str_if(4, n.op_type);
    for (auto& a : n.attribute) {
      Writer aw;
      write_attr(aw, a);
      nw.str(5, aw.out);
    }
    w.str(1, nw.out);
  }
  w.str_if(2, g.name);
  for (auto& t : g.initializer) {
    Writer tw;
    write_tensor(tw, t);
    w.str(5, tw.out);
  }
  for (auto& v : g.input) {
    Writer vw;
    write_value_info(vw, v);
    w.str(11, vw.out);
  }
  for (auto& v : g.output) {
    Writer vw;
    write_value_info(vw, v);
    w.str(12, vw.out);
  }
  for (auto& a : g.quantization_annotation) {
    Writer aw;
    aw.str_if(1, a.tensor_name);
    for (auto& kv : a.quant_parameter_tensor_names) {
      Writer e;
      e.str_if(1, kv.first);
      e.str_if(2, kv.second);
      aw.str(2, e.out);
    }
    w.str(14, aw.out);
  }
  return w.out;
}

}  // namespace

// ModelProto: ir_version=1 producer_name=2 producer_version=3 domain=4 model_version=5 graph=7 (onnx.proto:209-246)
void parse(ModelProto& m, const std::string& bytes) {
  Reader r(bytes);
  while (!r.done()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    switch (field) {
      case 1: m.ir_version = (int64_t)r.varint(); break;
      case 2: m.producer_name = r.bytes(); break;
      case 3: m.producer_version = r.bytes(); break;
      case 4: m.domain = r.bytes(); break;
      case 5: m.model_version = (int64_t)r.varint(); break;
      case 7: parse_graph(m.graph, r.sub()); break;
      default: r.skip(wire);
    }
  }
}

std::string serialize(const ModelProto& m) {
  Writer w;
  w.i64_if(1, m.ir_version);
  w.str_if(2, m.producer_name);
  w.str_if(3, m.producer_version);
  w.str_if(4, m.domain);
  w.i64_if(5, m.model_version);
  w.str(7, graph_bytes(m.graph));
  return w.out;
}

// ================================================================ save (onnx/save.hpp)
namespace {

const std::unordered_map<std::string, int32_t>& name2onnxtype() {  // serialize.hpp:32-44
  static const std::unordered_map<std::string, int32_t> m = {
      {"DOUBLE", DOUBLE}, {"FLOAT", FLOAT}, {"UINT8", UINT8}, {"INT8", INT8}, {"UINT16", UINT16},
      {"INT16", INT16}, {"UINT32", UINT32}, {"INT32", INT32}, {"UINT64", UINT64}, {"INT64", INT64}};
  return m;
}

int32_t onnx_typecode(const iTensor& tens) {  // MarshFuncs::get_typecode, serialize.hpp:48-53
  return name2onnxtype().at(egen::name_type((egen::_GENERATED_DTYPE)tens.get_meta().type_code()));
}

egen::_GENERATED_DTYPE egen_type(int32_t onnx_type) {
  for (auto& kv : name2onnxtype())
    if (kv.second == onnx_type) return egen::get_type(kv.first);
  global::fatalf("unknown onnx type %d", (int)onnx_type);
  return egen::BAD_TYPE;
}

template <typename CAST, typename T>
void pack(const void* data, size_t n, std::vector<T>& out) {  // serialize.hpp:20-29
  const CAST* ptr = (const CAST*)data;
  out.reserve(n);
  for (size_t i = 0; i < n; ++i) out.push_back((T)ptr[i]);
}

void marsh_leaf(TensorProto& out, const iLeaf& leaf) {  // MarshFuncs::marsh_leaf, serialize.hpp:55-110
  const void* data = leaf.device().data();
  if (nullptr == data) global::fatalf("cannot save leaf %s without data", leaf.to_string().c_str());
  const size_t n = leaf.shape().n_elems();
  out.data_type = onnx_typecode(leaf);
  switch (out.data_type) {
    case DOUBLE: pack<double>(data, n, out.double_data); break;
    case FLOAT: pack<float>(data, n, out.float_data); break;
    case INT32: pack<int32_t>(data, n, out.int32_data); break;
    case UINT8: pack<uint8_t>(data, n, out.int32_data); break;
    case INT8: pack<int8_t>(data, n, out.int32_data); break;
    case UINT16: pack<uint16_t>(data, n, out.int32_data); break;
    case INT16: pack<int16_t>(data, n, out.int32_data); break;
    case UINT32: pack<uint32_t>(data, n, out.uint64_data); break;
    case UINT64: pack<uint64_t>(data, n, out.uint64_data); break;
    case INT64: pack<int64_t>(data, n, out.int64_data); break;
    default: global::fatalf("unknown onnx type %d", (int)out.data_type);
  }
}

std::string usage_name(Usage u) {  // internal/teq/src/ileaf.cpp:12-35
  switch (u) {
    case IMMUTABLE: return "constant";
    case VARUSAGE: return "variable";
    case PLACEHOLDER: return "placeholder";
    default: return "";
  }
}

Usage named_usage(const std::string& name) {
  if (name == "constant") return IMMUTABLE;
  if (name == "variable") return VARUSAGE;
  if (name == "placeholder") return PLACEHOLDER;
  return UNKNOWN_USAGE;
}

std::string new_id() {  // the reference draws boost uuids (global::get_generator()->get_str())
  static std::mt19937_64 rng{std::random_device{}()};
  const uint64_t a = rng(), b = rng();
  char buf[40];
  std::snprintf(buf, sizeof(buf), "%08x-%04x-4%03x-%04x-%012llx", (unsigned)(a >> 32), (unsigned)((a >> 16) & 0xffff), (unsigned)(a & 0xfff),
                (unsigned)(0x8000 | ((b >> 48) & 0x3fff)), (unsigned long long)(b & 0xffffffffffffull));
  return buf;
}

// OnnxAttrMarshaler + marshal_attrs (marshal.hpp:28-151, src/marshal.cpp:8-25): every packed array is INTS (pairs are
// stored flattened, packattr.hpp encode_pair), scalars INT / FLOAT, strings STRING, tensor references TENSOR by id
void marshal_attrs(std::vector<AttributeProto>& out, const marsh::iAttributed& attrib, const std::unordered_map<const iTensor*, std::string>& tensid) {
  for (const std::string& key : attrib.ls_attrs()) {
    if (key == layer_attr) continue;  // the marshaler resolves the layer's graph itself
    const marsh::iObject* obj = attrib.get_attr(key);
    AttributeProto pb;
    pb.name = key;
    if (auto s = dynamic_cast<const marsh::String*>(obj)) {
      pb.type = ATTR_STRING;
      pb.s = s->val_;
    } else if (auto i = dynamic_cast<const marsh::Integer*>(obj)) {
      pb.type = ATTR_INT;
      pb.i = i->val_;
    } else if (auto f = dynamic_cast<const marsh::Float*>(obj)) {
      pb.type = ATTR_FLOAT;
      pb.f = (float)f->val_;
    } else if (auto arr = dynamic_cast<const marsh::IntArray*>(obj)) {
      pb.type = ATTR_INTS;
      pb.ints = arr->vals_;
    } else if (auto pairs = dynamic_cast<const marsh::PairArray*>(obj)) {
      pb.type = ATTR_INTS;
      for (auto& p : pairs->vals_) {
        pb.ints.push_back(p.first);
        pb.ints.push_back(p.second);
      }
    } else if (auto t = dynamic_cast<const TensorObj*>(obj)) {
      auto it = tensid.find(t->get_tensor().get());
      if (it == tensid.end()) global::fatalf("cannot find %s", t->get_tensor()->to_string().c_str());
      pb.type = ATTR_TENSOR;
      pb.t.name = it->second;
    } else {
      global::fatalf("onnx does not support attribute `%s` (%s)", key.c_str(), obj->to_string().c_str());
    }
    out.push_back(std::move(pb));
  }
}

void value_info(ValueInfoProto& out, const std::string& id, int32_t elem_type, const Shape& shape) {
  out.name = id;
  out.elem_type = elem_type;
  out.dims.assign(shape.begin(), shape.end());
}

struct OnnxMarshaler final : public iTraveler {  // save.hpp:31-281
  OnnxMarshaler(GraphProto& graph, const TensIds& identified, TensSetT stops) : pb_graph_(graph), identified_(identified), stops_(std::move(stops)) {
    for (auto& n : graph.node) preexisting_.insert(n.name);
    for (auto& t : graph.initializer) preexisting_.insert(t.name);
    for (auto& v : graph.input) preexisting_.insert(v.name);
  }

  void visit(iLeaf& leaf) override {
    if (tens_.count(&leaf)) return;
    const std::string id = get_id(leaf);
    roots_.insert(&leaf);
    tens_.emplace(&leaf, id);
    if (stops_.count(&leaf)) {
      add_input(id, leaf);
      return;
    }
    TensorAnnotation ann;  // marshal_annotation, src/marshal.cpp:44-53
    ann.tensor_name = id;
    ann.quant_parameter_tensor_names.push_back({leafname_key, leaf.to_string()});
    ann.quant_parameter_tensor_names.push_back({leafusage_key, usage_name(leaf.get_usage())});
    pb_graph_.quantization_annotation.push_back(std::move(ann));
    if (PLACEHOLDER == leaf.get_usage()) {
      pb_graph_.input.emplace_back();
      value_info(pb_graph_.input.back(), id, onnx_typecode(leaf), leaf.shape());
    } else {  // constant or variable
      pb_graph_.initializer.emplace_back();
      TensorProto& pb = pb_graph_.initializer.back();
      pb.name = id;
      const Shape shape = leaf.shape();
      pb.dims.assign(shape.begin(), shape.end());
      marsh_leaf(pb, leaf);
    }
  }

  void visit(iFunctor& func) override {
    if (tens_.count(&func)) return;
    if (stops_.count(&func)) {
      const std::string id = get_id(func);
      roots_.insert(&func);
      tens_.emplace(&func, id);
      add_input(id, func);
      return;
    }
    if (auto lattr = func.get_attr(layer_attr)) marshal_layer(func, static_cast<const LayerObj*>(lattr));
    else marshal_func(func);
  }

  void marshal_func(iFunctor& func) {  // save.hpp:155-204
    roots_.insert(&func);
    TensptrsT deps = func.get_args();
    for (auto& key : func.ls_attrs())
      if (auto ref = dynamic_cast<const TensorRef*>(func.get_attr(key))) deps.push_back(ref->get_tensor());
    multi_visit(*this, deps);
    for (auto& key : func.ls_attrs())
      if (auto tattr = dynamic_cast<const TensorObj*>(func.get_attr(key))) roots_.erase(tattr->get_tensor().get());
    const std::string id = get_id(func);
    NodeProto node;
    node.name = id;
    node.output.push_back(id);
    node.op_type = func.get_opcode().name_;
    for (auto& child : func.args_ref()) {
      auto it = tens_.find(child.get());
      if (it == tens_.end()) global::fatalf("cannot find child traversed %s", child->to_string().c_str());
      node.input.push_back(it->second);
      roots_.erase(child.get());
    }
    marshal_attrs(node.attribute, func, tens_);
    pb_graph_.node.push_back(std::move(node));
    tens_.emplace(&func, id);
  }

  void marshal_layer(iFunctor& func, const LayerObj* layer) {  // save.hpp:206-257
    TensptrT input = layer->get_tensor();
    input->accept(*this);  // the layer's input is marshalled in the enclosing graph first
    roots_.erase(input.get());
    NodeProto node;
    node.op_type = layer->get_opname();
    AttributeProto inner;
    inner.name = layer_attr;
    inner.type = ATTR_GRAPH;
    inner.g = std::make_shared<GraphProto>();
    const std::string subid = tens_.at(input.get());
    node.input.push_back(subid);
    inner.g->input.emplace_back();
    value_info(inner.g->input.back(), subid, onnx_typecode(*input), input->shape());
    TensSetT substops = stops_;
    substops.insert(input.get());
    OnnxMarshaler sub(*inner.g, identified_, substops);
    sub.roots_ = roots_;
    sub.tens_ = tens_;
    sub.marshal_func(func);
    roots_ = sub.roots_;
    tens_ = sub.tens_;
    const std::string id = tens_.at(&func);
    inner.g->output.emplace_back();
    value_info(inner.g->output.back(), id, UNDEFINED, func.shape());  // marshal_io writes the shape only
    node.name = id;
    node.output.push_back(id);
    node.attribute.push_back(std::move(inner));
    marshal_attrs(node.attribute, func, tens_);
    pb_graph_.node.push_back(std::move(node));
  }

  std::unordered_set<const iTensor*> roots_;
  std::unordered_map<const iTensor*, std::string> tens_;

 private:
  void add_input(const std::string& id, iTensor& tens) {
    if (preexisting_.count(id)) return;
    pb_graph_.input.emplace_back();
    value_info(pb_graph_.input.back(), id, onnx_typecode(tens), tens.shape());
  }

  std::string get_id(iTensor& tens) const {
    auto it = identified_.by_tens.find(&tens);
    if (it != identified_.by_tens.end()) return it->second;
    std::string out = new_id();
    while (preexisting_.count(out)) out = new_id();
    return out;
  }

  GraphProto& pb_graph_;
  const TensIds& identified_;
  TensSetT stops_;
  std::unordered_set<std::string> preexisting_;
};

}  // namespace

void save_graph(GraphProto& pb_graph, const TensptrsT& roots, const TensIds& identified, const TensSetT& stops) {
  OnnxMarshaler marshal(pb_graph, identified, stops);
  multi_visit(marshal, roots);
  // the reference iterates an unordered_set here (ORDERED_SAVE sorts by id); keep the callers' root order
  // first so that load_from_file returns models in the order they were saved
  std::vector<const iTensor*> rtens;
  for (auto& r : roots)
    if (marshal.roots_.count(r.get()) && std::find(rtens.begin(), rtens.end(), r.get()) == rtens.end()) rtens.push_back(r.get());
  std::vector<const iTensor*> rest;
  for (auto r : marshal.roots_)
    if (std::find(rtens.begin(), rtens.end(), r) == rtens.end()) rest.push_back(r);
  std::sort(rest.begin(), rest.end(), [&](const iTensor* a, const iTensor* b) { return marshal.tens_.at(a) < marshal.tens_.at(b); });
  rtens.insert(rtens.end(), rest.begin(), rest.end());
  for (const iTensor* root : rtens) {
    pb_graph.output.emplace_back();
    value_info(pb_graph.output.back(), marshal.tens_.at(root), UNDEFINED, root->shape());
  }
}

// ================================================================ load (onnx/src/load.cpp, serial/src/serialize.cpp)
namespace {

template <typename CAST, typename T>
TensptrT unpack(Usage usage, egen::_GENERATED_DTYPE dtype, Shape shape, const std::string& label, const std::vector<T>& data) {  // serialize.cpp:8-33
  const size_t n = shape.n_elems();
  std::vector<CAST> cdata(data.begin(), data.end());
  switch (usage) {
    case IMMUTABLE:
      if (cdata.size() != n) global::fatalf("leaf %s holds %d values for shape %s", label.c_str(), (int)cdata.size(), shape.to_string().c_str());
      return TensptrT(eteq::Constant::get(cdata.data(), dtype, shape));
    case VARUSAGE:
      if (cdata.size() != n) global::fatalf("leaf %s holds %d values for shape %s", label.c_str(), (int)cdata.size(), shape.to_string().c_str());
      return TensptrT(eteq::Variable::get(cdata.data(), dtype, shape, label, usage));
    case PLACEHOLDER: {
      std::vector<CAST> z(n, 0);
      return TensptrT(eteq::Variable::get(z.data(), dtype, shape, label, usage));
    }
    default:
      global::fatal("cannot unpack leaf of unknown usage");
  }
  return nullptr;
}

TensptrT unmarsh_leaf(const TensorProto& pb, Usage usage, const std::string& label) {  // serialize.cpp:37-92
  std::vector<DimT> slist(pb.dims.begin(), pb.dims.end());
  Shape shape(slist);
  switch (pb.data_type) {
    case DOUBLE: return unpack<double>(usage, egen::DOUBLE, shape, label, pb.double_data);
    case FLOAT: return unpack<float>(usage, egen::FLOAT, shape, label, pb.float_data);
    case INT32: return unpack<int32_t>(usage, egen::INT32, shape, label, pb.int32_data);
    case UINT8: return unpack<uint8_t>(usage, egen::UINT8, shape, label, pb.int32_data);
    case INT8: return unpack<int8_t>(usage, egen::INT8, shape, label, pb.int32_data);
    case UINT16: return unpack<uint16_t>(usage, egen::UINT16, shape, label, pb.int32_data);
    case INT16: return unpack<int16_t>(usage, egen::INT16, shape, label, pb.int32_data);
    case UINT32: return unpack<uint32_t>(usage, egen::UINT32, shape, label, pb.uint64_data);
    case UINT64: return unpack<uint64_t>(usage, egen::UINT64, shape, label, pb.uint64_data);
    case INT64: return unpack<int64_t>(usage, egen::INT64, shape, label, pb.int64_data);
    default: global::fatalf("unknown onnx type %d", (int)pb.data_type);
  }
  return nullptr;
}

// unmarshal_attrs (src/marshal.cpp:57-150). The reference keeps every integer list as NumArray<int64>; this host's
// packers (egen.cpp) distinguish pair lists, so the two pair-valued keys are re-paired here.
const GraphProto* unmarshal_attrs(marsh::Maps& out, const std::vector<AttributeProto>& pb_attrs, const TensIds& identified) {
  const GraphProto* subgraph = nullptr;
  for (const auto& pb : pb_attrs) {
    marsh::iObject* val = nullptr;
    switch (pb.type) {
      case ATTR_STRING: val = new marsh::String(pb.s); break;
      case ATTR_INT: val = new marsh::Integer(pb.i); break;
      case ATTR_FLOAT: val = new marsh::Float(pb.f); break;
      case ATTR_INTS:
        if (pb.name == eigen::dimpairs_key || pb.name == eigen::rankpairs_key) {
          if (pb.ints.size() % 2) global::fatalf("cannot decode odd vector %s into vec of pairs", fmts::to_string(pb.ints.begin(), pb.ints.end()).c_str());
          std::vector<std::pair<int64_t, int64_t>> pairs;
          for (size_t i = 0; i + 1 < pb.ints.size(); i += 2) pairs.push_back({pb.ints[i], pb.ints[i + 1]});
          val = new marsh::PairArray(pairs);
        } else {
          val = new marsh::IntArray(pb.ints);
        }
        break;
      case ATTR_TENSOR: {
        auto it = identified.by_id.find(pb.t.name);
        if (it == identified.by_id.end()) global::fatalf("cannot find tensor id %s", pb.t.name.c_str());
        val = new TensorObj(it->second);
      } break;
      case ATTR_GRAPH:
        if (pb.name == layer_attr) subgraph = pb.g.get();
        else global::fatalf("unknown graph attribute `%s`", pb.name.c_str());
        continue;
      default:
        global::fatalf("unknown onnx attribute type of `%s`", pb.name.c_str());
    }
    out.add_attr(pb.name, marsh::ObjptrT(val));
  }
  return subgraph;
}

}  // namespace

TensptrsT load_graph(TensIds& identified, const GraphProto& pb_graph) {
  std::unordered_map<std::string, std::unordered_map<std::string, std::string>> annotations;  // unmarshal_annotation
  for (auto& a : pb_graph.quantization_annotation)
    for (auto& kv : a.quant_parameter_tensor_names) annotations[a.tensor_name].emplace(kv.first, kv.second);
  auto annotated = [&](const std::string& id, const std::string& key) {
    auto it = annotations.find(id);
    if (it == annotations.end()) return std::string();
    auto jt = it->second.find(key);
    return jt == it->second.end() ? std::string() : jt->second;
  };
  for (const ValueInfoProto& pb_input : pb_graph.input) {
    const std::string& id = pb_input.name;
    if (identified.by_id.count(id)) continue;  // allow previously defined ids
    TensorProto pb_ten;
    pb_ten.dims = pb_input.dims;
    pb_ten.data_type = pb_input.elem_type;
    identified.insert(unmarsh_leaf(pb_ten, PLACEHOLDER, annotated(id, leafname_key)), id);
  }
  for (const TensorProto& pb_ten : pb_graph.initializer) {
    const std::string& id = pb_ten.name;
    if (identified.by_id.count(id)) continue;
    identified.insert(unmarsh_leaf(pb_ten, named_usage(annotated(id, leafusage_key)), annotated(id, leafname_key)), id);
  }
  for (const NodeProto& pb_node : pb_graph.node) {
    if (pb_node.op_type.empty()) global::fatal("onnx node without op_type");
    TensptrsT args;
    for (const std::string& input : pb_node.input) {
      auto it = identified.by_id.find(input);
      if (it == identified.by_id.end()) global::fatalf("failed to find input %s", input.c_str());
      args.push_back(it->second);
    }
    marsh::Maps attrs;
    TensptrT tens;
    if (const GraphProto* sub = unmarshal_attrs(attrs, pb_node.attribute, identified)) {
      TensptrsT roots = load_graph(identified, *sub);
      if (roots.empty() || args.empty()) global::fatalf("layer %s has no root or no input", pb_node.op_type.c_str());
      tens = layr::make_layer(roots.front(), pb_node.op_type, args.front());  // UnmarshFuncs::unmarsh_layr
    } else {
      const std::string& id = pb_node.name;
      if (identified.by_id.count(id)) global::fatalf("duplicate id %s", id.c_str());
      if (args.empty()) global::fatalf("cannot generate func %s without args", pb_node.op_type.c_str());
      tens = eteq::make_funcattr(egen::get_op(pb_node.op_type), args, attrs);  // UnmarshFuncs::unmarsh_func
      identified.insert(tens, id);
    }
  }
  TensptrsT roots;
  for (const ValueInfoProto& pb_output : pb_graph.output) {
    auto it = identified.by_id.find(pb_output.name);
    if (it == identified.by_id.end()) global::fatalf("failed to find output %s", pb_output.name.c_str());
    roots.push_back(it->second);
  }
  return roots;
}

// ================================================================ model / file level
void save_model(ModelProto& pb_model, const TensptrsT& roots, const TensIds& identified) {  // serial.cpp:12-46
  pb_model.ir_version = IR_VERSION;
  pb_model.producer_name = "tenncor";
  pb_model.producer_version = "1.0.0";
  pb_model.domain = "com.mingkaic.tenncor";
  pb_model.model_version = IR_VERSION;
  if (roots.empty()) return;
  save_graph(pb_model.graph, roots, identified);
}

TensptrsT load_model(TensIds& identified, const ModelProto& pb_model) { return load_graph(identified, pb_model.graph); }

bool save_to_file(const std::string& filename, const TensptrsT& models, const std::vector<std::pair<std::string, TensptrT>>& keys) {  // eteq_ext.cpp:461-487
  if (models.empty()) {
    std::fprintf(stderr, "[warn] attempting to save to file `%s` without specifying models\n", filename.c_str());
    return false;
  }
  std::ofstream output(filename, std::ios::binary);
  if (!output.is_open()) global::throw_errf("file %s not found", filename.c_str());
  ModelProto pb_model;
  TensIds identified;
  for (auto& kv : keys) identified.insert(kv.second, kv.first);
  save_model(pb_model, models, identified);
  const std::string bytes = serialize(pb_model);
  output.write(bytes.data(), (std::streamsize)bytes.size());
  return output.good();
}

TensptrsT load_from_file(const std::string& filename, const std::unordered_map<std::string, size_t>& key_prec) {  // eteq_ext.cpp:408-460
  std::ifstream input(filename, std::ios::binary);
  if (!input.is_open()) global::throw_errf("file %s not found", filename.c_str());
  std::stringstream ss;
  ss << input.rdbuf();
  ModelProto pb_model;
  try {
    parse(pb_model, ss.str());
  } catch (...) {
    global::throw_errf("failed to parse onnx from %s", filename.c_str());
  }
  TensIds ids;
  TensptrsT roots = load_model(ids, pb_model);
  // roots named in key_prec come first, ordered by their precedence; the rest follow in file order
  std::vector<std::string> precids, root_ids;
  for (auto& root : roots) {
    const std::string& id = ids.by_tens.at(root.get());
    (key_prec.count(id) ? precids : root_ids).push_back(id);
  }
  std::sort(precids.begin(), precids.end(), [&](const std::string& a, const std::string& b) { return key_prec.at(a) < key_prec.at(b); });
  TensptrsT out;
  for (auto& id : precids) out.push_back(ids.by_id.at(id));
  for (auto& id : root_ids) out.push_back(ids.by_id.at(id));
  return out;
}

}  // namespace onnx
