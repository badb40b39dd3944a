// backprop.cpp — gradient-graph rules as a table: opcode -> builder.
//
// What each rule must EMIT is fixed by the reference (tenncor/eteq/backprop.hpp:58-576): the derivative graphs are part
// of the contract — they decide which opcodes and operand orders the kernels see, tenncor/eteq/test/test_backprop.cpp
// pins them as printed graphs (tests/test_backprop_golden.py reproduces all 35 verbatim) and test_equation.cpp pins
// their values (tests/test_equation_golden.py). How the rules are organised is ours: one small builder per rule family,
// registered in a dense table indexed by opcode; `lderive` is a table lookup. Families:
//   pass-through      d/dx = upstream                      IDENTITY CAST ROUND ADD, SUB (arg 0)
//   chain rule        d/dx = local(x) * upstream           ABS SIN COS EXP SQUARE CUBE SIGMOID TANH POW MUL MAX MIN
//   quotient forms    TAN LOG SQRT DIV
//   reductions        upstream broadcast back over the reduced ranks (REDUCE_*), and its adjoint (EXTEND)
//   layout adjoints   PERMUTE RESHAPE SLICE<->PAD CONCAT STRIDE<->SCATTER REVERSE
//   products          MATMUL CONTRACT CONV
//   no gradient       comparisons, RAND_UNIF (zero); ASSIGN*, ARGMAX (fatal)
#include <array>
#include <numeric>

#include "eteq.hpp"

namespace eteq {

using namespace teq;
using namespace egen;

namespace {

/// one differentiation site: functor `f` with operands `in`, differentiated w.r.t. operand `at`, upstream gradient `up`
struct Site {
  const FuncptrT& f;
  const TensptrsT& in;
  const TensptrT& up;
  size_t at;
};

using Rule = TensptrT (*)(const Site&);

TensptrT like(float value, const TensptrT& shape_of) { return make_constant_like(value, shape_of); }

/// ranks listed in `order` followed by the ranks it omits, then inverted: where each rank of the input ended up
RanksT inverse_order(const RanksT& order) {
  uint32_t seen = 0;
  RanksT full(order.begin(), order.end());
  for (RankT r : order) seen |= 1u << r;
  for (RankT r = 0; r < rank_cap; ++r)
    if (!(seen >> r & 1u)) full.push_back(r);
  RanksT inv(rank_cap);
  for (size_t pos = 0; pos < rank_cap; ++pos) inv[full[pos]] = (RankT)pos;
  return inv;
}

/// `what` broadcast back to `wide` along the ranks the reduction `red` removed
TensptrT unreduce(const Shape& wide, const TensptrT& what, const FuncptrT& red) {
  DimsT bcast(rank_cap, 1);
  for (RankT r : eigen::unpack_rankset(*red))
    if (r < rank_cap) bcast[r] = wide.at(r);
  return make_functor(EXTEND, {what}, bcast);
}

// ---------------------------------------------------------------- pass-through / sign
TensptrT rule_pass(const Site& s) { return s.up; }
TensptrT rule_neg(const Site& s) { return make_functor(NEG, {s.up}); }
TensptrT rule_sub(const Site& s) { return s.at == 0 ? s.up : make_functor(NEG, {s.up}); }

// ---------------------------------------------------------------- chain rule: local derivative times upstream
template <TensptrT (*Local)(const Site&)>
TensptrT chain(const Site& s) {
  return make_functor(MUL, {Local(s), s.up});
}

TensptrT d_abs(const Site& s) { return make_functor(DIV, {s.in[0], s.f}); }
TensptrT d_sin(const Site& s) { return make_functor(COS, {s.in[0]}); }
TensptrT d_cos(const Site& s) { return make_functor(NEG, {make_functor(SIN, {s.in[0]})}); }
TensptrT d_exp(const Site& s) { return s.f; }
TensptrT d_square(const Site& s) { return make_functor(MUL, {like(2.f, s.in[0]), s.in[0]}); }
TensptrT d_cube(const Site& s) { return make_functor(MUL, {like(3.f, s.in[0]), make_functor(SQUARE, {s.in[0]})}); }
TensptrT d_sigmoid(const Site& s) { return make_functor(MUL, {s.f, make_functor(SUB, {like(1.f, s.f), s.f})}); }
TensptrT d_tanh(const Site& s) { return make_functor(SUB, {like(1.f, s.f), make_functor(SQUARE, {s.f})}); }
TensptrT d_pow(const Site& s) {
  if (s.at == 0) {
    auto exponent_less_one = make_functor(SUB, {s.in[1], like(1.f, s.in[1])});
    return make_functor(MUL, {s.in[1], make_functor(POW, {s.in[0], exponent_less_one})});
  }
  return make_functor(MUL, {make_functor(LOG, {s.in[0]}), s.f});
}
TensptrT d_mul(const Site& s) {  // product of the other operands
  TensptrsT others;
  for (size_t k = 0; k < s.in.size(); ++k)
    if (k != s.at) others.push_back(s.in[k]);
  return make_functor(MUL, others);
}
TensptrT d_extremum(const Site& s) { return make_functor(EQ, {s.f, s.in.at(s.at)}); }

// ---------------------------------------------------------------- quotient forms
TensptrT rule_tan(const Site& s) { return make_functor(DIV, {s.up, make_functor(SQUARE, {make_functor(COS, {s.in[0]})})}); }
TensptrT rule_log(const Site& s) { return make_functor(DIV, {s.up, s.in[0]}); }
TensptrT rule_sqrt(const Site& s) { return make_functor(DIV, {s.up, make_functor(MUL, {like(2.f, s.f), s.f})}); }
TensptrT rule_div(const Site& s) {
  if (s.at == 0) return make_functor(DIV, {s.up, s.in[1]});
  auto numer = make_functor(MUL, {make_functor(NEG, {s.up}), s.in[0]});
  return make_functor(DIV, {make_functor(DIV, {numer, s.in[1]}), s.in[1]});
}

// ---------------------------------------------------------------- reductions and their adjoint
TensptrT rule_reduce_sum(const Site& s) { return unreduce(s.in[0]->shape(), s.up, s.f); }
TensptrT rule_reduce_prod(const Site& s) {
  const Shape wide = s.in[0]->shape();
  return make_functor(MUL, {unreduce(wide, s.up, s.f), make_functor(DIV, {unreduce(wide, s.f, s.f), s.in[0]})});
}
TensptrT rule_reduce_extremum(const Site& s) {  // reproduced as is, incl. the comparison against arg * upstream
  const Shape wide = s.in[0]->shape();
  return make_functor(EQ, {unreduce(wide, s.f, s.f), make_functor(MUL, {s.in[0], unreduce(wide, s.up, s.f)})});
}
TensptrT rule_extend(const Site& s) {
  const DimsT bcast = eigen::unpack_extend(s.in[0]->shape(), *s.f).second;
  std::set<RankT> widened;
  for (size_t r = 0; r < bcast.size() && r < rank_cap; ++r)
    if (bcast[r] > 1) widened.insert((RankT)r);
  return make_functor(REDUCE_SUM, {s.up}, widened);
}

// ---------------------------------------------------------------- layout adjoints
TensptrT rule_permute(const Site& s) { return make_functor(PERMUTE, {s.up}, inverse_order(eigen::unpack_ranks(*s.f))); }
TensptrT rule_reshape(const Site& s) { return make_functor(RESHAPE, {s.up}, s.in[0]->shape()); }
TensptrT rule_reverse(const Site& s) { return make_functor(REVERSE, {s.up}, eigen::unpack_rankset(*s.f)); }
TensptrT rule_stride(const Site& s) { return make_functor(SCATTER, {s.up}, s.in[0]->shape(), eigen::unpack_dims(*s.f)); }
TensptrT rule_scatter(const Site& s) {
  DimsT incrs = eigen::unpack_dims(*s.f);
  if (incrs.size() > rank_cap) incrs.resize(rank_cap);
  return make_functor(STRIDE, {s.up}, incrs);
}
TensptrT rule_slice(const Site& s) {  // zero-pad back to the sliced tensor's extents
  const Shape whole = s.in[0]->shape();
  eigen::PairVecT<DimT> pads;
  const auto cuts = eigen::unpack_dimpairs(*s.f);
  for (size_t r = 0; r < cuts.size() && r < rank_cap; ++r) {
    const DimT extent_r = whole.at(r);
    const DimT lo = std::min(cuts[r].first, (DimT)(extent_r - 1));
    const DimT len = std::min(cuts[r].second, (DimT)(extent_r - lo));
    pads.push_back({lo, (DimT)(extent_r - lo - len)});
  }
  return make_functor(PAD, {s.up}, pads);
}
TensptrT rule_pad(const Site& s) {  // cut the padding off again
  const Shape padded = s.f->shape();
  eigen::PairVecT<DimT> cuts;
  const auto pads = eigen::unpack_dimpairs(*s.f);
  for (size_t r = 0; r < pads.size() && r < rank_cap; ++r)
    cuts.push_back({pads[r].first, (DimT)(padded.at(r) - pads[r].first - pads[r].second)});
  return make_functor(SLICE, {s.up}, cuts);
}
TensptrT rule_concat(const Site& s) {  // the operand's own block of the upstream gradient
  const RankT axis = eigen::unpack_rank(*s.f);
  eigen::PairVecT<DimT> cuts(std::max(rank_cap, axis), {0, std::numeric_limits<DimT>::max()});
  if (s.in.size() > 2) cuts[axis] = {(DimT)s.at, 1};  // n-ary form: every operand has extent 1 along the axis
  else cuts[axis] = {s.at ? s.in[0]->shape().at(axis) : (DimT)0, s.in[s.at]->shape().at(axis)};
  return make_functor(SLICE, {s.up}, cuts);
}

// ---------------------------------------------------------------- products
TensptrT rule_matmul(const Site& s) {
  auto flipped = make_functor(PERMUTE, {s.in[1 - s.at]}, RanksT{1, 0});
  return s.at == 0 ? make_functor(MATMUL, {s.up, flipped}) : make_functor(MATMUL, {flipped, s.up});
}

/// contract(A, B, pairs) lays its result out as <free ranks of B, free ranks of A>. The gradient w.r.t. one operand
/// contracts the upstream gradient with the OTHER operand over that operand's free ranks — which sit at a known
/// position inside the upstream gradient — and then permutes <paired ranks, free ranks> of the differentiated
/// operand back into its own rank order.
TensptrT rule_contract(const Site& s) {
  struct Side {
    RanksT paired, free;
  } side[2];
  uint32_t used[2] = {0, 0};
  for (auto pr : eigen::unpack_rankpairs(*s.f)) {
    side[0].paired.push_back(pr.first);
    side[1].paired.push_back(pr.second);
    used[0] |= 1u << pr.first;
    used[1] |= 1u << pr.second;
  }
  for (int k = 0; k < 2; ++k)
    for (RankT r = 0, n = (RankT)narrow_shape(s.in[k]->shape()).size(); r < n; ++r)
      if (!(used[k] >> r & 1u)) side[k].free.push_back(r);
  const int me = (int)s.at, other = 1 - me;
  // where the other operand's free ranks start inside the upstream gradient: B-free first, then A-free
  const RankT base = other == 1 ? 0 : (RankT)side[1].free.size();
  eigen::PairVecT<RankT> pairs;
  for (RankT k = 0, n = (RankT)side[other].free.size(); k < n; ++k) pairs.push_back({(RankT)(base + k), side[other].free[k]});
  if (pairs.empty())  // outer product: contract over the first singular rank of both
    pairs.push_back({(RankT)narrow_shape(s.up->shape()).size(), (RankT)narrow_shape(s.in[other]->shape()).size()});
  RanksT landed = side[me].paired;  // the new product comes out as <my paired ranks, my free ranks>
  landed.insert(landed.end(), side[me].free.begin(), side[me].free.end());
  return make_functor(PERMUTE, {make_functor(CONTRACT, {s.up, s.in[other]}, pairs)}, inverse_order(landed));
}

TensptrT rule_conv(const Site& s) {
  RanksT slid;  // image rank each kernel rank slides along
  for (RankT r : eigen::unpack_ranks(*s.f)) {
    if (slid.size() >= rank_cap || r >= rank_cap) break;
    slid.push_back(r);
  }
  if (s.at == 1) {  // kernel: correlate the image with the upstream gradient, then name the ranks like the kernel does
    RanksT identity(rank_cap);
    std::iota(identity.begin(), identity.end(), 0);
    return make_functor(PERMUTE, {make_functor(CONV, {s.in[0], s.up}, identity)}, slid);
  }
  // image: full correlation = pad the upstream gradient by (extent - 1) on every slid rank, reverse the kernel
  const Shape kernel = s.in[1]->shape();
  eigen::PairVecT<DimT> pads(rank_cap, {0, 0});
  std::set<RankT> every;
  for (size_t q = 0; q < slid.size(); ++q) {
    const DimT reach = kernel.at(q) - 1;
    pads[slid[q]] = {reach, reach};
    every.insert((RankT)q);
  }
  return make_functor(CONV, {make_functor(PAD, {s.up}, pads), make_functor(REVERSE, {s.in[1]}, every)}, slid);
}

// ---------------------------------------------------------------- no gradient
TensptrT rule_select(const Site& s) {
  if (s.at == 0) return like(0.f, s.in[0]);
  auto zero = like(0.f, s.f);
  return s.at == 1 ? make_functor(SELECT, {s.in[0], s.up, zero}) : make_functor(SELECT, {s.in[0], zero, s.up});
}
TensptrT rule_zero(const Site& s) { return like(0.f, s.in[0]); }

struct RuleTable {
  std::array<Rule, _N_GENERATED_OPCODES> rule{};  // nullptr: not differentiable
  RuleTable() {
    auto set = [this](std::initializer_list<int> ops, Rule r) {
      for (int op : ops) rule[op] = r;
    };
    set({IDENTITY, CAST, ROUND, ADD}, rule_pass);
    set({NEG}, rule_neg);
    set({SUB}, rule_sub);
    set({ABS}, chain<d_abs>);
    set({SIN}, chain<d_sin>);
    set({COS}, chain<d_cos>);
    set({EXP}, chain<d_exp>);
    set({SQUARE}, chain<d_square>);
    set({CUBE}, chain<d_cube>);
    set({SIGMOID}, chain<d_sigmoid>);
    set({TANH}, chain<d_tanh>);
    set({POW}, chain<d_pow>);
    set({MUL}, chain<d_mul>);
    set({MAX, MIN}, chain<d_extremum>);
    set({TAN}, rule_tan);
    set({LOG}, rule_log);
    set({SQRT}, rule_sqrt);
    set({DIV}, rule_div);
    set({REDUCE_SUM}, rule_reduce_sum);
    set({REDUCE_PROD}, rule_reduce_prod);
    set({REDUCE_MAX, REDUCE_MIN}, rule_reduce_extremum);
    set({EXTEND}, rule_extend);
    set({PERMUTE}, rule_permute);
    set({RESHAPE}, rule_reshape);
    set({REVERSE}, rule_reverse);
    set({STRIDE}, rule_stride);
    set({SCATTER}, rule_scatter);
    set({SLICE}, rule_slice);
    set({PAD}, rule_pad);
    set({CONCAT}, rule_concat);
    set({MATMUL}, rule_matmul);
    set({CONTRACT}, rule_contract);
    set({CONV}, rule_conv);
    set({SELECT}, rule_select);
    set({RAND_UNIF, EQ, NEQ, GT, LT}, rule_zero);
  }
};

}  // namespace

TensptrT DerivativeFuncs::lderive(FuncptrT op, TensptrT supgrad, size_t arg_idx) const {
  static const RuleTable table;
  const Opcode opcode = op->get_opcode();
  const size_t code = opcode.code_;
  if (code >= table.rule.size() || code == BAD_OP) global::fatalf("Unknown op %s", opcode.name_.c_str());
  const Rule rule = table.rule[code];
  if (rule == nullptr) global::fatalf("cannot derive %s", opcode.name_.c_str());  // ASSIGN*, ARGMAX
  const TensptrsT operands = op->get_args();
  return rule(Site{op, operands, supgrad, arg_idx});
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
