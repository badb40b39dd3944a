// egen.hpp — the generated OPCODE / DTYPE surface and per-opcode shape / type / elision
// rules, restated by hand.
//
// The reference generates these from cfg/ops.yml + cfg/fulltype.yml with tools/egen
// (plugins/opcodes.py:9-141, plugins/dtypes.py:8-145): `_GENERATED_OPCODE` (IDENTITY = 1
// ... CAST = 50), `_GENERATED_DTYPE`, `name_op/get_op`, `is_idempotent/is_commutative`,
// `ShapeParser<OP>`, `TypeParser<OP>`, `FuncOpt<OP>` and `typed_exec<T>`. The numeric
// values are kept identical (they also travel through the C-ABI, include/tcr_b200.h).
// Attribute keys are the reference's Packer keys (internal/eigen/src/packattr.cpp:8-24).
#ifndef TCR_HOST_EGEN_HPP
#define TCR_HOST_EGEN_HPP

#include "teq.hpp"

namespace egen {

enum _GENERATED_OPCODE {
  BAD_OP = 0,
  IDENTITY, ABS, NEG, SIN, COS, TAN, EXP, LOG, SQRT, ROUND, SIGMOID, TANH, SQUARE, CUBE,
  RAND_UNIF, REVERSE, REDUCE_SUM, REDUCE_PROD, REDUCE_MIN, REDUCE_MAX, ARGMAX, PERMUTE,
  EXTEND, RESHAPE, SLICE, PAD, STRIDE, SCATTER, POW, ADD, SUB, MUL, DIV, MIN, MAX, EQ, NEQ,
  LT, GT, MATMUL, CONTRACT, CONV, SELECT, CONCAT, ASSIGN, ASSIGN_ADD, ASSIGN_SUB, ASSIGN_MUL,
  ASSIGN_DIV, CAST,
  _N_GENERATED_OPCODES,
};

enum _GENERATED_DTYPE {
  BAD_TYPE = 0,
  DOUBLE, FLOAT, INT8, UINT8, INT16, UINT16, INT32, UINT32, INT64, UINT64,
  _N_GENERATED_DTYPES,
};

const _GENERATED_DTYPE default_dtype = FLOAT;  // cfg/fulltype.yml:3

std::string name_op(_GENERATED_OPCODE code);
_GENERATED_OPCODE get_op(const std::string& name);
bool is_commutative(_GENERATED_OPCODE code);  // ADD MUL MIN MAX EQ NEQ (cfg/ops.yml "commutative")
bool is_idempotent(_GENERATED_OPCODE code);   // false: RAND_UNIF ASSIGN_ADD/SUB/MUL/DIV CAST (cfg/ops.yml "idempotent")

std::string name_type(_GENERATED_DTYPE type);
_GENERATED_DTYPE get_type(const std::string& name);
uint8_t type_size(_GENERATED_DTYPE type);
size_t type_precision(_GENERATED_DTYPE type);

template <typename T> _GENERATED_DTYPE get_type() { return BAD_TYPE; }
template <> inline _GENERATED_DTYPE get_type<double>() { return DOUBLE; }
template <> inline _GENERATED_DTYPE get_type<float>() { return FLOAT; }
template <> inline _GENERATED_DTYPE get_type<int8_t>() { return INT8; }
template <> inline _GENERATED_DTYPE get_type<uint8_t>() { return UINT8; }
template <> inline _GENERATED_DTYPE get_type<int16_t>() { return INT16; }
template <> inline _GENERATED_DTYPE get_type<uint16_t>() { return UINT16; }
template <> inline _GENERATED_DTYPE get_type<int32_t>() { return INT32; }
template <> inline _GENERATED_DTYPE get_type<uint32_t>() { return UINT32; }
template <> inline _GENERATED_DTYPE get_type<int64_t>() { return INT64; }
template <> inline _GENERATED_DTYPE get_type<uint64_t>() { return UINT64; }

/// host-side element conversion (dtypes.py `type_convert`): out[i] = OUT(in[i])
void type_convert(void* out, _GENERATED_DTYPE outtype, const void* input, _GENERATED_DTYPE intype, size_t nelems);

#define TCR_TYPE_LOOKUP(DTYPE, T, ...)                                  \
  switch (DTYPE) {                                                      \
    case egen::DOUBLE: { using T = double; __VA_ARGS__; } break;        \
    case egen::FLOAT: { using T = float; __VA_ARGS__; } break;          \
    case egen::INT8: { using T = int8_t; __VA_ARGS__; } break;          \
    case egen::UINT8: { using T = uint8_t; __VA_ARGS__; } break;        \
    case egen::INT16: { using T = int16_t; __VA_ARGS__; } break;        \
    case egen::UINT16: { using T = uint16_t; __VA_ARGS__; } break;      \
    case egen::INT32: { using T = int32_t; __VA_ARGS__; } break;        \
    case egen::UINT32: { using T = uint32_t; __VA_ARGS__; } break;      \
    case egen::INT64: { using T = int64_t; __VA_ARGS__; } break;        \
    case egen::UINT64: { using T = uint64_t; __VA_ARGS__; } break;      \
    default: global::fatal("executing bad type");                       \
  }

}  // namespace egen

namespace eigen {

template <typename T> using PairVecT = std::vector<std::pair<T, T>>;
using DTypesT = std::vector<egen::_GENERATED_DTYPE>;
using OptDimsT = std::pair<bool, teq::DimsT>;  // (present, dims)

const std::string no_argument_err = "cannot operate without inputs";

// attribute keys (internal/eigen/src/packattr.cpp:8-24)
const std::string dtype_key = "dtype";
const std::string dimpairs_key = "dimension_pairs";
const std::string rankpairs_key = "rank_pairs";
const std::string dims_key = "dimensions";
const std::string ranks_key = "ranks";
const std::string rankset_key = "rank_set";
const std::string rank_key = "rank";
const std::string shape_key = "shape";
const std::string tensor_key = "tensor";

// pack (eigen::pack_attr overloads, packattr.hpp:356-361)
void pack_attr(marsh::iAttributed&);
void pack_attr(marsh::iAttributed& a, egen::_GENERATED_DTYPE dtype);
void pack_attr(marsh::iAttributed& a, const PairVecT<teq::DimT>& dimpairs);
void pack_attr(marsh::iAttributed& a, const PairVecT<teq::RankT>& rankpairs);
void pack_attr(marsh::iAttributed& a, const teq::DimsT& dims);
void pack_attr(marsh::iAttributed& a, const teq::RanksT& ranks);
void pack_attr(marsh::iAttributed& a, const std::set<teq::RankT>& rankset);
void pack_attr(marsh::iAttributed& a, teq::RankT rank);
void pack_attr(marsh::iAttributed& a, const teq::Shape& shape);
void pack_attr(marsh::iAttributed& a, const teq::TensptrT& tens);
template <typename A, typename B, typename... R>
void pack_attr(marsh::iAttributed& a, const A& x, const B& y, const R&... rest) {
  pack_attr(a, x);
  pack_attr(a, y, rest...);
}

// unpack: fatal "cannot find `key` attribute" when missing (packattr.hpp:78-90)
egen::_GENERATED_DTYPE unpack_dtype(const marsh::iAttributed& a);
PairVecT<teq::DimT> unpack_dimpairs(const marsh::iAttributed& a);
PairVecT<teq::RankT> unpack_rankpairs(const marsh::iAttributed& a);
teq::DimsT unpack_dims(const marsh::iAttributed& a);
teq::RanksT unpack_ranks(const marsh::iAttributed& a);
std::set<teq::RankT> unpack_rankset(const marsh::iAttributed& a);
teq::RankT unpack_rank(const marsh::iAttributed& a);
teq::Shape unpack_shape(const marsh::iAttributed& a);
teq::TensptrT unpack_tensor(const marsh::iAttributed& a);

/// broadcast list from "dimensions", or derived from a "tensor" attribute's shape
/// (internal/eigen/src/packattr.cpp:28-60)
OptDimsT unpack_extend(teq::Shape inshape, const marsh::iAttributed& attrib);

/// ShapeParser<OP> (cfg/ops.yml per_op + opcalls)
teq::Shape shape_parse(egen::_GENERATED_OPCODE op, const marsh::iAttributed& attrs, const teq::ShapesT& shapes);
/// TypeParser<OP>: max precision; ASSIGN* take the target's; CAST takes the "dtype" attr
egen::_GENERATED_DTYPE type_parse(egen::_GENERATED_OPCODE op, const marsh::iAttributed& attrs, const DTypesT& dtypes);
/// FuncOpt<OP>: true when the functor is redundant and the first child is returned instead
bool func_opt(egen::_GENERATED_OPCODE op, egen::_GENERATED_DTYPE outtype, const marsh::iAttributed& attrs, const teq::TensptrsT& args);

inline bool is_2d(const teq::Shape& shape) {
  return std::all_of(shape.begin() + 2, shape.end(), [](teq::DimT d) { return 1 == d; });
}

}  // namespace eigen

#endif  // TCR_HOST_EGEN_HPP
