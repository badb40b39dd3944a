// hone.cpp — see hone.hpp. Reference: tenncor/hone/duplicates.hpp, src/duplicates.cpp, cstrules.hpp, src/optimize.cpp.
#include "hone.hpp"

#include <algorithm>
#include <map>
#include <unordered_set>

namespace hone {

using namespace teq;

namespace {

// post-order over args (and tensor-valued attributes, which the reference's hash reads too)
struct Collector {
  std::vector<iTensor*> order;
  std::unordered_set<iTensor*> seen;
  std::unordered_map<iTensor*, TensptrT> owners;
  std::unordered_map<iTensor*, size_t> height;

  void visit(const TensptrT& t) {
    if (!seen.insert(t.get()).second) return;
    owners.emplace(t.get(), t);
    size_t h = 0;
    if (auto f = dynamic_cast<iFunctor*>(t.get())) {
      for (auto& a : f->args_ref()) {
        visit(a);
        h = std::max(h, height[a.get()] + 1);
      }
      for (auto& key : f->ls_attrs())
        if (auto ref = dynamic_cast<const TensorRef*>(f->get_attr(key))) visit(ref->get_tensor());
    }
    height[t.get()] = h;
    order.push_back(t.get());
  }
};

struct Hasher {
  TensMapT<std::string> hashes;
  std::unordered_map<std::string, std::string> ids;  // label -> short id
  size_t next = 0;

  std::string fresh() { return "#" + std::to_string(next++); }

  const std::string& encode(iTensor* t, const std::string& label) {
    auto it = ids.find(label);
    if (it == ids.end()) it = ids.emplace(label, fresh()).first;
    return hashes.emplace(t, it->second).first->second;
  }

  void leaf(iLeaf& l) {
    if (IMMUTABLE == l.get_usage()) {  // constants are equal when shape, type and bytes are
      auto& meta = l.get_meta();
      const char* data = (const char*)l.device().data();
      std::string label = l.shape().to_string() + "|" + meta.type_label();
      label.append(data, data + l.shape().n_elems() * meta.type_size());
      encode(&l, label);
    } else {
      hashes.emplace(&l, fresh());
    }
  }

  void func(iFunctor& f) {
    const auto opcode = (egen::_GENERATED_OPCODE)f.get_opcode().code_;
    if (!egen::is_idempotent(opcode)) {  // see the deviation note in hone.hpp
      hashes.emplace(&f, fresh());
      return;
    }
    std::vector<std::string> hs;
    for (auto& a : f.args_ref()) hs.push_back(hashes.at(a.get()));
    if (egen::is_commutative(opcode)) std::sort(hs.begin(), hs.end());
    std::map<std::string, std::string> attrs;
    for (auto& key : f.ls_attrs()) {
      const marsh::iObject* value = f.get_attr(key);
      if (auto ref = dynamic_cast<const TensorRef*>(value)) attrs.emplace(key, hashes.at(ref->get_tensor().get()) + (dynamic_cast<const LayerObj*>(value) ? "@" + value->to_string() : ""));
      else attrs.emplace(key, value->to_string());
    }
    std::string label = f.shape().to_string() + "|" + f.get_opcode().name_ + "|" + f.get_meta().type_label() + "\\";
    for (auto& kv : attrs) label += kv.first + ":" + kv.second + ";";
    label += "\\";
    for (auto& h : hs) label += h + ",";
    encode(&f, label);
  }
};

// point every parent at the canonical owner of its children; returns the new roots
TensptrsT apply(const Collector& graph, const OwnMapT& converts, TensptrsT roots) {
  if (converts.empty()) return roots;
  auto canon = [&](const TensptrT& t) {
    auto it = converts.find(t.get());
    return it == converts.end() ? t : it->second;
  };
  for (iTensor* t : graph.order) {
    if (converts.count(t)) continue;  // dropped node: nobody will reach it
    auto f = dynamic_cast<iFunctor*>(t);
    if (!f) continue;
    const TensptrsT args = f->get_args();
    for (size_t i = 0; i < args.size(); ++i) {
      TensptrT c = canon(args[i]);
      if (c != args[i]) f->update_child(c, i);
    }
  }
  for (auto& r : roots) r = canon(r);
  return roots;
}

}  // namespace

TensMapT<std::string> hash_graph(const TensptrsT& roots) {
  Collector graph;
  for (auto& r : roots) graph.visit(r);
  Hasher hasher;
  for (iTensor* t : graph.order) {
    if (auto f = dynamic_cast<iFunctor*>(t)) hasher.func(*f);
    else hasher.leaf(*static_cast<iLeaf*>(t));
  }
  return hasher.hashes;
}

TensptrsT merge_dups(TensptrsT roots, size_t* merged) {
  Collector graph;
  for (auto& r : roots) graph.visit(r);
  Hasher hasher;
  // children first (post-order), so a parent's hash already sees its children's canonical identity
  std::unordered_map<std::string, TensptrT> first;
  OwnMapT converts;
  for (iTensor* t : graph.order) {
    if (auto f = dynamic_cast<iFunctor*>(t)) hasher.func(*f);
    else hasher.leaf(*static_cast<iLeaf*>(t));
    const bool mergeable = dynamic_cast<iFunctor*>(t) != nullptr || IMMUTABLE == static_cast<iLeaf*>(t)->get_usage();
    if (!mergeable) continue;
    auto ins = first.emplace(hasher.hashes.at(t), graph.owners.at(t));
    if (!ins.second) converts.emplace(t, ins.first->second);
  }
  if (merged) *merged = converts.size();
  return apply(graph, converts, std::move(roots));
}

// Which functors constant folding evaluates: post-order, a functor is constant when every argument is a constant leaf or was
// itself found constant; only the top-most constant functor of each chain is evaluated (its constant children are evaluated
// on the way, on the device). The reference reaches the same fixed point one level per round (generate_cstrules,
// tenncor/hone/src/cstrules.cpp:8-37: one source pattern "op over constant leaves" per opcode and branching factor) and, like it,
// never folds an IDENTITY: an identity is structure (a layer root, a dependency carrier), so nothing above one folds either.
static std::vector<iTensor*> fold_tops(const Collector& graph) {
  std::unordered_set<iTensor*> constant;
  std::vector<iTensor*> tops;
  for (iTensor* t : graph.order) {
    auto f = dynamic_cast<iFunctor*>(t);
    if (!f) {
      if (IMMUTABLE == static_cast<iLeaf*>(t)->get_usage()) constant.insert(t);
      continue;
    }
    if (!egen::is_idempotent((egen::_GENERATED_OPCODE)f->get_opcode().code_)) continue;
    if (egen::IDENTITY == (egen::_GENERATED_OPCODE)f->get_opcode().code_) continue;
    bool all = !f->args_ref().empty();
    for (auto& a : f->args_ref()) all &= constant.count(a.get()) > 0;
    if (all) constant.insert(t);
  }
  // top-most: constant functors with a non-constant parent (or that are roots)
  std::unordered_set<iTensor*> has_const_parent;
  for (iTensor* t : graph.order)
    if (auto f = dynamic_cast<iFunctor*>(t))
      if (constant.count(t))
        for (auto& a : f->args_ref()) has_const_parent.insert(a.get());
  for (iTensor* t : graph.order)
    if (dynamic_cast<iFunctor*>(t) && constant.count(t) && !has_const_parent.count(t)) tops.push_back(t);
  return tops;
}

TensptrsT fold_candidates(const TensptrsT& roots) {
  Collector graph;
  for (auto& r : roots) graph.visit(r);
  TensptrsT out;
  for (iTensor* t : fold_tops(graph)) out.push_back(graph.owners.at(t));
  return out;
}

TensptrsT fold_constants(TensptrsT roots, size_t* folded) {
  Collector graph;
  for (auto& r : roots) graph.visit(r);
  OwnMapT converts;
  std::vector<iTensor*> tops = fold_tops(graph);
  if (!tops.empty()) {
    TensptrsT targets;
    for (iTensor* t : tops) targets.push_back(graph.owners.at(t));
    eteq::run(targets);  // ConstantTarget::convert: evaluate, then read the value back (cstrules.hpp:48-58)
    for (auto& t : targets) {
      const void* data = t->device().data();
      if (nullptr == data) global::fatalf("constant folding: %s produced no data", t->to_string().c_str());
      converts.emplace(t.get(), eteq::make_constant_tensor(data, (egen::_GENERATED_DTYPE)t->get_meta().type_code(), t->shape()));
    }
  }
  if (folded) *folded = converts.size();
  return apply(graph, converts, std::move(roots));
}

TensptrsT optimize(TensptrsT roots, Stats* stats, bool fold) {
  Stats s;
  {
    Collector g;
    for (auto& r : roots) g.visit(r);
    for (auto t : g.order) s.functors_before += dynamic_cast<iFunctor*>(t) != nullptr;
  }
  size_t n = 0;
  roots = merge_dups(std::move(roots), &n);  // remove duplicates first to reduce the search space
  s.merged += n;
  for (size_t round = 0; fold && round < 50; ++round) {  // convert_round_limit
    size_t f = 0, m = 0;
    roots = fold_constants(std::move(roots), &f);
    if (f == 0) break;
    roots = merge_dups(std::move(roots), &m);
    s.folded += f;
    s.merged += m;
    s.rounds = round + 1;
  }
  {
    Collector g;
    for (auto& r : roots) g.visit(r);
    for (auto t : g.order) s.functors_after += dynamic_cast<iFunctor*>(t) != nullptr;
  }
  if (stats) *stats = s;
  return roots;
}

}  // namespace hone
