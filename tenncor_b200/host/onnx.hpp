// onnx.hpp — the reference's ONNX-dialect model files.
//
// Mirrors internal/onnx/{marshal,save,load}.hpp + src/*.cpp, tenncor/serial/serialize.hpp and
// tenncor/src/serial.cpp:12-140: functors become NodeProtos named by an id, leaves become
// initializers (variables / constants) or graph inputs (placeholders) annotated with their label
// and usage, and every layer root (IDENTITY carrying the "layer" attribute) becomes ONE node whose
// "layer" attribute is a nested GraphProto holding the layer's sub-graph. Files written by the
// reference (models/gd.onnx, rbm.onnx, ...) load here; files written here follow the same layout.
//
// protobuf is not available to the C++ host in this image, so the wire format is read and written
// by a small codec restricted to the messages and fields of internal/onnx/onnx.proto that the
// reference's save / load touch (field numbers cited in onnx.cpp).
#ifndef TCR_HOST_ONNX_HPP
#define TCR_HOST_ONNX_HPP

#include <memory>
#include <string>
#include <vector>

#include "layr.hpp"

namespace onnx {

// TensorProto.DataType (onnx.proto:321-346)
enum DataType { UNDEFINED = 0, FLOAT = 1, UINT8 = 2, INT8 = 3, UINT16 = 4, INT16 = 5, INT32 = 6, INT64 = 7, STRING = 8, BOOL = 9,
                FLOAT16 = 10, DOUBLE = 11, UINT32 = 12, UINT64 = 13 };
// AttributeProto.AttributeType (onnx.proto:106-121)
enum AttributeType { ATTR_UNDEFINED = 0, ATTR_FLOAT = 1, ATTR_INT = 2, ATTR_STRING = 3, ATTR_TENSOR = 4, ATTR_GRAPH = 5, ATTR_FLOATS = 6,
                     ATTR_INTS = 7, ATTR_STRINGS = 8, ATTR_TENSORS = 9, ATTR_GRAPHS = 10 };
const int64_t IR_VERSION = 6;  // onnx.proto:93

struct TensorProto {
  std::vector<int64_t> dims;
  int32_t data_type = UNDEFINED;
  std::string name;
  std::vector<float> float_data;
  std::vector<int32_t> int32_data;
  std::vector<int64_t> int64_data;
  std::vector<double> double_data;
  std::vector<uint64_t> uint64_data;
  std::string raw_data;
};

struct GraphProto;

struct AttributeProto {
  std::string name;
  int32_t type = ATTR_UNDEFINED;
  float f = 0;
  int64_t i = 0;
  std::string s;
  TensorProto t;
  std::shared_ptr<GraphProto> g;
  std::vector<float> floats;
  std::vector<int64_t> ints;
  std::vector<std::string> strings;
  std::vector<TensorProto> tensors;
};

struct NodeProto {
  std::vector<std::string> input, output;
  std::string name, op_type;
  std::vector<AttributeProto> attribute;
};

struct ValueInfoProto {  // name + TypeProto.Tensor {elem_type, shape}
  std::string name;
  int32_t elem_type = UNDEFINED;
  std::vector<int64_t> dims;
};

struct TensorAnnotation {
  std::string tensor_name;
  std::vector<std::pair<std::string, std::string>> quant_parameter_tensor_names;
};

struct GraphProto {
  std::vector<NodeProto> node;
  std::string name;
  std::vector<TensorProto> initializer;
  std::vector<ValueInfoProto> input, output;
  std::vector<TensorAnnotation> quantization_annotation;
};

struct ModelProto {
  int64_t ir_version = 0, model_version = 0;
  std::string producer_name, producer_version, domain;
  GraphProto graph;
};

/// wire format; `parse` throws (global::throw_err) on malformed input
void parse(ModelProto& out, const std::string& bytes);
std::string serialize(const ModelProto& model);

const std::string leafname_key = "TENSOR_NAME";  // marshal.hpp:24-26
const std::string leafusage_key = "LEAF_USAGE";

/// tensor <-> id (the reference's boost::bimap TensptrIdT / TensIdT)
struct TensIds {
  void insert(const teq::TensptrT& tens, const std::string& id) {
    if (by_tens.count(tens.get()) || by_id.count(id)) return;  // bimap insert semantics: first mapping wins
    by_tens.emplace(tens.get(), id);
    by_id.emplace(id, tens);
  }
  std::unordered_map<teq::iTensor*, std::string> by_tens;
  std::unordered_map<std::string, teq::TensptrT> by_id;
};

/// serial::save_graph (serialize.hpp:113-120 over onnx/save.hpp:283-307)
void save_graph(GraphProto& pb_graph, const teq::TensptrsT& roots, const TensIds& identified = {}, const teq::TensSetT& stops = {});
/// serial::load_graph (serial/src/serialize.cpp:117-123 over onnx/src/load.cpp:8-107): graph outputs in file order
teq::TensptrsT load_graph(TensIds& identified, const GraphProto& pb_graph);

/// tcr::save_model / tcr::load_model (tenncor/src/serial.cpp:12-85) without the distributed manager
void save_model(ModelProto& pb_model, const teq::TensptrsT& roots, const TensIds& identified = {});
teq::TensptrsT load_model(TensIds& identified, const ModelProto& pb_model);

/// tc.save_to_file / tc.load_from_file (tenncor/python/eteq_ext.cpp:408-487)
bool save_to_file(const std::string& filename, const teq::TensptrsT& models, const std::vector<std::pair<std::string, teq::TensptrT>>& keys = {});
teq::TensptrsT load_from_file(const std::string& filename, const std::unordered_map<std::string, size_t>& key_prec = {});

}  // namespace onnx

#endif  // TCR_HOST_ONNX_HPP
