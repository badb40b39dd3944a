// teq.cpp — travelers, evaluator traversal and the reverse-mode graph builder.
#include "teq.hpp"

#include <cassert>

namespace teq {

void iOnceTraveler::visit(iLeaf& leaf) {
  if (visited_.emplace((iTensor*)&leaf).second) visit_leaf(leaf);
}

void iOnceTraveler::visit(iFunctor& func) {
  if (visited_.emplace((iTensor*)&func).second) visit_func(func);
}

TensptrsT attr_tensors(const iFunctor& func) {
  TensptrsT out;
  for (auto& name : func.ls_attrs())
    if (auto ref = dynamic_cast<const TensorRef*>(func.get_attr(name))) out.push_back(ref->get_tensor());
  return out;
}

void GraphStat::visit(iFunctor& func) {
  if (height_.count(&func)) return;
  size_t h = 0;
  for (auto& child : func.args_ref()) {
    child->accept(*this);
    h = std::max(h, height_.at(child.get()));
  }
  height_.emplace(&func, h + 1);
}

void GraphIndex::visit(iFunctor& func) {
  if (indices_.count(&func)) return;
  for (auto& child : func.args_ref()) child->accept(*this);
  indices_.emplace(&func, indices_.size());
}

void PathFinder::visit_func(iFunctor& func) {
  if (targets_.count(&func)) return;
  auto& args = func.args_ref();
  multi_visit(*this, args);
  TensptrsT atens;
  if (func.size() > 0) atens = attr_tensors(func);
  if (follow_attrs_) multi_visit(*this, atens);
  PathDirection next;
  for (size_t i = 0, n = args.size(); i < n; ++i) {
    iTensor* t = args[i].get();
    if (targets_.count(t) || roadmap_.count(t)) next.args_.push_back(i);
  }
  for (auto& name : func.ls_attrs()) {
    if (auto ref = dynamic_cast<const TensorRef*>(func.get_attr(name))) {
      iTensor* t = ref->get_tensor().get();
      if (targets_.count(t) || roadmap_.count(t)) next.attrs_.push_back(name);
    }
  }
  if (!next.args_.empty() || !next.attrs_.empty()) roadmap_.emplace(&func, std::move(next));
}

void Copier::visit_leaf(iLeaf& leaf) {
  if (ignores_.count(&leaf)) return;
  clones_.emplace(&leaf, TensptrT(leaf.clone()));
}

void Copier::visit_func(iFunctor& func) {
  if (ignores_.count(&func)) return;
  auto deps = func.get_args();
  auto fcpy = func.clone();
  multi_visit(*this, deps);
  for (size_t i = 0, n = deps.size(); i < n; ++i) {
    auto it = clones_.find(deps[i].get());
    if (it != clones_.end()) fcpy->update_child(it->second, i);
  }
  for (auto& attr : fcpy->ls_attrs()) {
    if (auto ref = dynamic_cast<const TensorRef*>(fcpy->get_attr(attr))) {
      auto reftens = ref->get_tensor();
      reftens->accept(*this);
      auto it = clones_.find(reftens.get());
      if (it != clones_.end()) {
        auto alt = ref->copynreplace(it->second);
        fcpy->rm_attr(attr);
        fcpy->add_attr(attr, marsh::ObjptrT(alt));
      }
    }
  }
  clones_.emplace(&func, TensptrT(fcpy));
}

namespace {
struct OwnerTracker final : public iOnceTraveler {
  OwnMapT owners_;

 private:
  void visit_leaf(iLeaf&) override {}
  void visit_func(iFunctor& func) override {
    auto& deps = func.args_ref();
    multi_visit(*this, deps);
    for (auto& dep : deps) owners_.emplace(dep.get(), dep);
  }
};
}  // namespace

OwnMapT track_ownptrs(const TensptrsT& roots) {
  OwnerTracker tracker;
  multi_visit(tracker, roots);
  for (auto& root : roots) tracker.owners_.emplace(root.get(), root);
  return tracker.owners_;
}

// ---------------------------------------------------------------- evaluation
TravEvaluator::TravEvaluator(iDevice& device, const TensSetT& targets, const TensSetT& ignored)
    : ignored_(ignored), device_(&device), targets_(targets) {
  for (auto ig : ignored)
    if (nullptr != ig && nullptr == ig->device().device_data())
      global::throw_errf("cannot ignore tensor %s without existing data", ig->to_string().c_str());
}

void TravEvaluator::visit_func(iFunctor& func) {
  if (ignored_.count(&func)) return;
  multi_visit(*this, func.args_ref());
  device_->calc(func, (size_t)targets_.count(&func));
}

static iEvalptrT& eval_slot() {
  static iEvalptrT slot = std::make_shared<Evaluator>();
  return slot;
}

void set_eval(iEvalptrT eval) { eval_slot() = eval ? eval : std::make_shared<Evaluator>(); }
iEvaluator& get_eval() { return *eval_slot(); }

// ---------------------------------------------------------------- derive
TensptrsT derive(TensptrT root, const TensptrsT& targets, const iDerivativeFuncs& funcs) {
  TensptrsT out;
  out.reserve(targets.size());
  if (root == nullptr) {
    for (auto& target : targets) out.push_back(funcs.get_const_zero(*target));
    return out;
  }
  GradMapT grads = {{root.get(), {funcs.get_const_one(*root)}}};
  TensSetT targset;
  for (auto& target : targets) targset.emplace(target.get());
  partial_derive(grads, {root}, targset, funcs);
  for (auto& target : targets) {
    TensptrT tens;
    if (nullptr == target) {
      tens = funcs.get_const_zero(*root);
    } else {
      auto it = grads.find(target.get());
      if (it != grads.end() && !it->second.empty())
        tens = it->second.size() == 1 ? it->second.front() : funcs.add(it->second);
      else
        tens = funcs.get_const_zero(*target);
    }
    out.push_back(tens);
  }
  return out;
}

void partial_derive(GradMapT& grads, const TensptrSetT& parents, const TensSetT& targets, const iDerivativeFuncs& funcs) {
  if (targets.empty()) return;
  TensSetT parset;
  for (auto& p : parents) parset.emplace(p.get());
  TensSetT tids;
  for (auto target : targets) {
    if (nullptr == target) continue;
    if (parset.count(target)) assert(grads.count(target));
    else tids.emplace(target);
  }
  if (tids.empty()) return;

  PathFinder pfinder(tids, /*follow_attrs=*/false);
  TensptrsT plist(parents.begin(), parents.end());
  multi_visit(pfinder, plist);
  if (pfinder.roadmap_.empty()) return;

  OwnMapT owners = track_ownptrs(plist);
  GraphStat stat;
  GraphIndex indexer;
  multi_visit(stat, plist);
  multi_visit(indexer, plist);

  // parents before children: max height descending; name / post-order index break ties
  // (internal/teq/src/derive.cpp:110-134)
  std::vector<iFunctor*> tovisits;
  tovisits.reserve(pfinder.roadmap_.size());
  for (auto& kv : pfinder.roadmap_) tovisits.push_back(static_cast<iFunctor*>(kv.first));
  std::sort(tovisits.begin(), tovisits.end(), [&](iFunctor* a, iFunctor* b) {
    size_t ah = stat.height_.at(a), bh = stat.height_.at(b);
    if (ah == bh) {
      std::string as = a->to_string(), bs = b->to_string();
      if (as == bs) return indexer.indices_.at(a) > indexer.indices_.at(b);
      return as > bs;
    }
    return ah > bh;
  });

  for (iFunctor* tens : tovisits) {
    auto git = grads.find(tens);
    if (git == grads.end() || git->second.empty())
      global::fatalf("failed to find existing grads for %s", tens->to_string().c_str());
    TensptrsT prevs = git->second;
    TensptrT bwd = prevs.size() > 1 ? funcs.add(prevs) : prevs.front();
    auto& nexts = pfinder.roadmap_.at(tens).args_;
    auto visitable = std::static_pointer_cast<iFunctor>(owners.at(tens));
    TensptrsT children = tens->get_args();
    for (size_t i : nexts)
      if (i < children.size()) grads[children[i].get()].push_back(funcs.lderive(visitable, bwd, i));
  }
}

}  // namespace teq
