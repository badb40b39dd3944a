// layr.cpp — make_layer / get_input / trail / connect / deep_clone / get_storage.
#include "layr.hpp"

namespace layr {

using namespace teq;

Shape gen_rshape(DimsT runcoms, Shape left, eigen::PairVecT<RankT> lrdims) {
  // right ranks named in lrdims take the matching left extent, the rest are filled in order
  std::array<bool, rank_cap> unvisited;
  unvisited.fill(true);
  DimsT slist(rank_cap, 1);
  for (auto& lr : lrdims) {
    slist[lr.second] = left.at(lr.first);
    unvisited[lr.second] = false;
  }
  for (size_t i = 0, j = 0, n = runcoms.size(); i < rank_cap && j < n; ++i)
    if (unvisited[i]) slist[i] = runcoms[j++];
  return Shape(slist);
}

ETensor make_layer(ETensor root, const std::string& layername, ETensor input) {
  auto f = dynamic_cast<iFunctor*>(root.get());
  if (nullptr == f) global::fatalf("cannot make a layer out of non-functor %s", root->to_string().c_str());
  if (nullptr != f->get_attr(layer_attr))
    global::fatalf("attempting to attach layer attribute to node %s with an existing layer attribute", root->to_string().c_str());
  f->add_attr(layer_attr, std::make_unique<LayerObj>(layername, input));
  return root;
}

ETensor get_input(const ETensor& root) {
  if (nullptr == root) global::fatal("cannot get layer attr with null root");
  auto froot = dynamic_cast<iFunctor*>(root.get());
  if (nullptr == froot) global::fatalf("%s is not a layer", root->to_string().c_str());
  auto layerattr = dynamic_cast<LayerObj*>(froot->get_attr(layer_attr));
  if (nullptr == layerattr) global::fatalf("%s has no layer attribute", root->to_string().c_str());
  return layerattr->get_tensor();
}

namespace {

/// re-instantiates every functor on a path to one of the inputs (layer.hpp:63-120)
struct Trailer final : public iOnceTraveler {
  explicit Trailer(const OwnMapT& inputs) : trailed_(inputs), pfinder_(keys(inputs), /*follow_attrs=*/true) {}
  OwnMapT trailed_;

 private:
  static TensSetT keys(const OwnMapT& m) {
    TensSetT out;
    for (auto& kv : m) out.emplace(kv.first);
    return out;
  }
  void visit_leaf(iLeaf&) override {}
  void visit_func(iFunctor& func) override {
    if (trailed_.count(&func)) return;
    func.accept(pfinder_);
    auto it = pfinder_.roadmap_.find(&func);
    if (it == pfinder_.roadmap_.end()) return;
    PathDirection& dir = it->second;
    marsh::Maps dup_attrs;
    marsh::get_attrs(dup_attrs, func);
    for (const std::string& attr : dir.attrs_) {
      auto ref = static_cast<const TensorRef*>(func.get_attr(attr));
      auto ctens = ref->get_tensor();
      ctens->accept(*this);
      auto tit = trailed_.find(ctens.get());
      if (tit != trailed_.end()) {
        dup_attrs.rm_attr(attr);
        dup_attrs.add_attr(attr, marsh::ObjptrT(ref->copynreplace(tit->second)));
      }
    }
    TensptrsT children = func.get_args();
    for (size_t i : dir.args_) {
      auto child = children[i];
      child->accept(*this);
      children[i] = trailed_.at(child.get());
    }
    auto opcode = (egen::_GENERATED_OPCODE)func.get_opcode().code_;
    trailed_.emplace(&func, eteq::make_funcattr(opcode, children, dup_attrs));
  }
  PathFinder pfinder_;
};

struct VarExtract final : public iOnceTraveler {
  explicit VarExtract(TensSetT term) : term_(std::move(term)) {}
  LeafsT variables_;

 private:
  void visit_leaf(iLeaf& leaf) override {
    if (term_.count(&leaf)) return;
    if (IMMUTABLE != leaf.get_usage()) variables_.push_back(&leaf);
  }
  void visit_func(iFunctor& func) override {
    if (term_.count(&func)) return;
    multi_visit(*this, func.args_ref());
  }
  TensSetT term_;
};

}  // namespace

ETensor trail(const ETensor& root, const OwnMapT& inputs) {
  Trailer trailer(inputs);
  root->accept(trailer);
  auto it = trailer.trailed_.find(root.get());
  return it == trailer.trailed_.end() ? nullptr : it->second;
}

ETensor connect(const ETensor& root, const ETensor& input) { return trail(root, OwnMapT{{get_input(root).get(), input}}); }

ETensor deep_clone(const ETensor& root) {
  Copier kamino({get_input(root).get()});
  root->accept(kamino);
  return kamino.clones_.at(root.get());
}

eteq::VarptrsT get_storage(const ETensor& root) {
  OwnMapT owner = track_ownptrs(TensptrsT{root});
  auto intens = get_input(root).get();
  VarExtract extra({intens});
  root->accept(extra);
  eteq::VarptrsT vars;
  for (auto leaf : extra.variables_)
    if (auto var = std::dynamic_pointer_cast<eteq::Variable>(owner.at(leaf))) vars.push_back(var);
  return vars;
}

}  // namespace layr
