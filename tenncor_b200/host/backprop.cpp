// backprop.cpp — per-opcode gradient graph rules.
//
// Restates tenncor/eteq/backprop.hpp:58-576: every rule emits functors from the same
// 50-opcode set, so the derivative graphs need no kernels of their own. The shapes of the
// emitted sub-graphs are kept identical to the reference's (they decide which opcodes and
// operand orders the kernels see, and the reference's golden gradients are matched to
// EXPECT_DOUBLE_EQ precision by tests/test_equation_golden.py).
#include <numeric>

#include "eteq.hpp"

namespace eteq {

using namespace teq;
using namespace egen;

/// gradient of a reduction: broadcast the upstream gradient back over the reduced ranks
/// (backprop.hpp:18-31)
static TensptrT reduce_grad(Shape shape, TensptrT bwd, FuncptrT fwd) {
  DimsT bcast(rank_cap, 1);
  for (RankT d : eigen::unpack_rankset(*fwd))
    if (d < rank_cap) bcast[d] = shape.at(d);
  return make_functor(EXTEND, {bwd}, bcast);
}

/// inverse of a (completed) permutation order (backprop.hpp:33-55)
static RanksT reorder_permute(RanksT order) {
  std::array<bool, rank_cap> visited;
  visited.fill(false);
  for (RankT i = 0, n = order.size(); i < n; ++i) visited[order[i]] = true;
  for (RankT i = 0; i < rank_cap; ++i)
    if (!visited[i]) order.push_back(i);
  RanksT reorder(rank_cap);
  for (size_t i = 0; i < rank_cap; ++i) reorder[order[i]] = i;
  return reorder;
}

static TensptrT constant_like(float scalar, TensptrT like) { return make_constant_like(scalar, like); }

TensptrT DerivativeFuncs::lderive(FuncptrT op, TensptrT supgrad, size_t arg_idx) const {
  auto args = op->get_args();
  Opcode opcode = op->get_opcode();
  TensptrT out;
  switch (opcode.code_) {
    case IDENTITY: case CAST: case ROUND: case ADD:
      out = supgrad;
      break;
    case NEG:
      out = make_functor(NEG, {supgrad});
      break;
    case TAN:
      out = make_functor(DIV, {supgrad, make_functor(SQUARE, {make_functor(COS, {args.front()})})});
      break;
    case LOG:
      out = make_functor(DIV, {supgrad, args.front()});
      break;
    case SQRT:
      out = make_functor(DIV, {supgrad, make_functor(MUL, {constant_like(2.f, op), op})});
      break;
    case ABS: case SIN: case COS: case EXP: case SQUARE: case CUBE: case SIGMOID: case TANH: case POW: case MUL:
    case MAX: case MIN: {
      TensptrT local_der;
      switch (opcode.code_) {
        case ABS: local_der = make_functor(DIV, {args.front(), op}); break;
        case SIN: local_der = make_functor(COS, {args.front()}); break;
        case COS: local_der = make_functor(NEG, {make_functor(SIN, {args.front()})}); break;
        case EXP: local_der = op; break;
        case SQUARE: local_der = make_functor(MUL, {constant_like(2.f, args.front()), args.front()}); break;
        case CUBE: local_der = make_functor(MUL, {constant_like(3.f, args.front()), make_functor(SQUARE, {args.front()})}); break;
        case SIGMOID: local_der = make_functor(MUL, {op, make_functor(SUB, {constant_like(1.f, op), op})}); break;
        case TANH: local_der = make_functor(SUB, {constant_like(1.f, op), make_functor(SQUARE, {op})}); break;
        case POW:
          local_der = arg_idx == 0
                          ? make_functor(MUL, {args[1], make_functor(POW, {args[0], make_functor(SUB, {args[1], constant_like(1.f, args[1])})})})
                          : make_functor(MUL, {make_functor(LOG, {args.front()}), op});
          break;
        case MUL: {
          TensptrsT nodes;
          for (size_t i = 0, n = args.size(); i < n; ++i)
            if (i != arg_idx) nodes.push_back(args[i]);
          local_der = make_functor(MUL, nodes);
        } break;
        case MAX: case MIN: local_der = make_functor(EQ, {op, args.at(arg_idx)}); break;
      }
      out = make_functor(MUL, {local_der, supgrad});
    } break;
    case SUB:
      out = arg_idx == 0 ? supgrad : make_functor(NEG, {supgrad});
      break;
    case DIV:
      out = arg_idx == 0 ? make_functor(DIV, {supgrad, args[1]})
                         : make_functor(DIV, {make_functor(DIV, {make_functor(MUL, {make_functor(NEG, {supgrad}), args[0]}), args[1]}), args[1]});
      break;
    case REDUCE_SUM:
      out = reduce_grad(args.front()->shape(), supgrad, op);
      break;
    case REDUCE_PROD:
      out = make_functor(MUL, {reduce_grad(args.front()->shape(), supgrad, op),
                               make_functor(DIV, {reduce_grad(args.front()->shape(), op, op), args.front()})});
      break;
    case REDUCE_MAX: case REDUCE_MIN:
      out = make_functor(EQ, {reduce_grad(args.front()->shape(), op, op),
                              make_functor(MUL, {args.front(), reduce_grad(args.front()->shape(), supgrad, op)})});
      break;
    case EXTEND: {
      DimsT bcast = eigen::unpack_extend(args.front()->shape(), *op).second;
      std::set<RankT> dims;
      for (size_t i = 0, n = std::min((size_t)rank_cap, bcast.size()); i < n; ++i)
        if (bcast[i] > 1) dims.emplace(i);
      out = make_functor(REDUCE_SUM, {supgrad}, dims);
    } break;
    case PERMUTE:
      out = make_functor(PERMUTE, {supgrad}, reorder_permute(eigen::unpack_ranks(*op)));
      break;
    case RESHAPE:
      out = make_functor(RESHAPE, {supgrad}, args.front()->shape());
      break;
    case MATMUL:
      if (arg_idx == 0) out = make_functor(MATMUL, {supgrad, make_functor(PERMUTE, {args[1]}, RanksT{1, 0})});
      else out = make_functor(MATMUL, {make_functor(PERMUTE, {args[0]}, RanksT{1, 0}), supgrad});
      break;
    case CONTRACT: {
      // contract(A, B, u) = C with ranks <b-free, a-free>; the gradient w.r.t. one operand
      // contracts the upstream gradient with the other operand over that operand's free
      // ranks and permutes the result back into the operand's rank order (backprop.hpp:269-359)
      auto dims = eigen::unpack_rankpairs(*op);
      std::array<bool, rank_cap> lvisit, rvisit;
      lvisit.fill(false);
      rvisit.fill(false);
      RanksT lucom_ranks, rucom_ranks, lcom_ranks, rcom_ranks;
      for (auto coms : dims) {
        lvisit[coms.first] = true;
        rvisit[coms.second] = true;
        lcom_ranks.push_back(coms.first);
        rcom_ranks.push_back(coms.second);
      }
      for (RankT i = 0, n = narrow_shape(args[0]->shape()).size(); i < n; ++i)
        if (!lvisit[i]) lucom_ranks.push_back(i);
      for (RankT i = 0, n = narrow_shape(args[1]->shape()).size(); i < n; ++i)
        if (!rvisit[i]) rucom_ranks.push_back(i);
      TensptrT right;
      RanksT order;
      eigen::PairVecT<RankT> grad_dims;
      if (arg_idx == 0) {
        right = args[1];
        for (RankT i = 0, n = rucom_ranks.size(); i < n; ++i) grad_dims.push_back({i, rucom_ranks[i]});
        order = lcom_ranks;  // contract output has ranks <lucom, lcom>
        order.insert(order.end(), lucom_ranks.begin(), lucom_ranks.end());
        order = reorder_permute(order);
      } else {
        right = args[0];
        for (RankT i = 0, n = lucom_ranks.size(); i < n; ++i) grad_dims.push_back({(RankT)(rucom_ranks.size() + i), lucom_ranks[i]});
        order = rcom_ranks;  // contract output has ranks <rcom, rucom>
        order.insert(order.end(), rucom_ranks.begin(), rucom_ranks.end());
        order = reorder_permute(order);
      }
      if (grad_dims.empty())
        grad_dims.push_back({(RankT)narrow_shape(supgrad->shape()).size(), (RankT)narrow_shape(right->shape()).size()});
      out = make_functor(PERMUTE, {make_functor(CONTRACT, {supgrad, right}, grad_dims)}, order);
    } break;
    case CONV: {
      RanksT order = eigen::unpack_ranks(*op);
      RanksT dims;
      for (size_t i = 0, n = std::min((size_t)rank_cap, order.size()); i < n && order[i] < rank_cap; ++i) dims.push_back(order[i]);
      if (arg_idx == 0) {
        // convolve(pad(C_grad_sup, Y.shape[dims]-1), reverse(Y))
        size_t ndims = dims.size();
        Shape kernshape = args[1]->shape();
        eigen::PairVecT<DimT> paddings(rank_cap, {0, 0});
        for (size_t i = 0; i < ndims; ++i) {
          DimT kpad = kernshape.at(i) - 1;
          paddings[dims[i]] = {kpad, kpad};
        }
        RanksT revdims(ndims);
        std::iota(revdims.begin(), revdims.end(), 0);
        out = make_functor(CONV, {make_functor(PAD, {supgrad}, paddings),
                                  make_functor(REVERSE, {args[1]}, std::set<RankT>(revdims.begin(), revdims.end()))}, dims);
      } else {
        // convolve(X, C_grad_sup)
        RanksT indices(rank_cap);
        std::iota(indices.begin(), indices.end(), 0);
        out = make_functor(PERMUTE, {make_functor(CONV, {args[0], supgrad}, indices)}, dims);
      }
    } break;
    case SLICE: {
      auto extents = eigen::unpack_dimpairs(*op);
      Shape cshape = args.front()->shape();
      eigen::PairVecT<DimT> paddings;
      for (size_t i = 0, n = std::min(extents.size(), (size_t)rank_cap); i < n; ++i) {
        DimT offset = std::min(extents[i].first, (DimT)(cshape.at(i) - 1));
        DimT extent = std::min(extents[i].second, (DimT)(cshape.at(i) - offset));
        paddings.push_back({offset, (DimT)(cshape.at(i) - (offset + extent))});
      }
      out = make_functor(PAD, {supgrad}, paddings);
    } break;
    case PAD: {
      auto paddings = eigen::unpack_dimpairs(*op);
      Shape oshape = op->shape();
      eigen::PairVecT<DimT> extents;
      for (size_t i = 0; i < std::min(paddings.size(), (size_t)rank_cap); ++i) {
        DimT offset = paddings[i].first;
        extents.push_back({offset, (DimT)(oshape.at(i) - paddings[i].second - offset)});
      }
      out = make_functor(SLICE, {supgrad}, extents);
    } break;
    case CONCAT: {
      Shape cshape = args[arg_idx]->shape();
      RankT axis = eigen::unpack_rank(*op);
      eigen::PairVecT<DimT> extents(std::max(rank_cap, axis), {0, std::numeric_limits<DimT>::max()});
      if (args.size() > 2) {
        extents[axis] = {(DimT)arg_idx, 1};
      } else {
        DimT offset = arg_idx ? args[0]->shape().at(axis) : 0;
        extents[axis] = {offset, cshape.at(axis)};
      }
      out = make_functor(SLICE, {supgrad}, extents);
    } break;
    case STRIDE:
      out = make_functor(SCATTER, {supgrad}, args[0]->shape(), eigen::unpack_dims(*op));
      break;
    case SCATTER: {
      DimsT c = eigen::unpack_dims(*op);
      DimsT strides(c.begin(), c.begin() + std::min((size_t)rank_cap, c.size()));
      out = make_functor(STRIDE, {supgrad}, strides);
    } break;
    case REVERSE:
      out = make_functor(REVERSE, {supgrad}, eigen::unpack_rankset(*op));
      break;
    case SELECT: {
      if (0 == arg_idx) {
        out = constant_like(0.f, args.front());
        break;
      }
      TensptrT condition = args[0], then, otherwise;
      if (arg_idx == 1) { then = supgrad; otherwise = constant_like(0.f, op); }
      else { then = constant_like(0.f, op); otherwise = supgrad; }
      out = make_functor(SELECT, {condition, then, otherwise});
    } break;
    case RAND_UNIF: case EQ: case NEQ: case GT: case LT:
      out = constant_like(0.f, args.front());
      break;
    case ASSIGN: case ASSIGN_ADD: case ASSIGN_SUB: case ASSIGN_MUL: case ASSIGN_DIV: case ARGMAX:
      global::fatalf("cannot derive %s", opcode.name_.c_str());
    default:
      global::fatalf("Unknown op %s", opcode.name_.c_str());
  }
  return out;
}

TensptrT DerivativeFuncs::get_const_one(iTensor& reference) const {
  return make_constant_scalar(1, reference.shape(), (_GENERATED_DTYPE)reference.get_meta().type_code());
}

TensptrT DerivativeFuncs::get_const_zero(iTensor& reference) const {
  return make_constant_scalar(0, reference.shape(), (_GENERATED_DTYPE)reference.get_meta().type_code());
}

TensptrT DerivativeFuncs::add(TensptrsT elems) const {
  if (elems.empty()) global::fatal("cannot add without gradients");
  return make_functor(ADD, elems);
}

}  // namespace eteq
