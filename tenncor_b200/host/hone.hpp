// hone.hpp — pre-evaluation graph rewrites: duplicate merging and constant folding.
//
// Mirrors tenncor/hone: `Hasher` + `merge_dups` (duplicates.hpp:16-103, src/duplicates.cpp:8-107) give every
// structurally equal sub-graph one node; `generate_cstrules` + `ConstantTarget` (cstrules.hpp:41-86,
// src/cstrules.cpp) replace a functor whose arguments are all constants by the constant it evaluates to —
// evaluated through the installed evaluator, i.e. on the device — and `optimize` (src/optimize.cpp:12-32) runs
// them to a fixed point (at most 50 rounds). The json rule file the reference feeds to the same entry point
// (cfg/optimizations.json) is a missing blob, so only these code-defined passes exist here; the generic
// opt/query rule engine is out of scope.
//
// Deliberate deviation: non-idempotent opcodes (RAND_UNIF, ASSIGN_*, CAST per cfg/ops.yml) are neither merged nor
// folded — the reference's Hasher would give two `rand_unif(lo, hi)` nodes of equal shape one identity and make
// every "independent" sample the same tensor.
#ifndef TCR_HOST_HONE_HPP
#define TCR_HOST_HONE_HPP

#include "eteq.hpp"

namespace hone {

struct Stats {
  size_t functors_before = 0, functors_after = 0, merged = 0, folded = 0, rounds = 0;
};

/// structural identity of every node under `roots`: equal strings <=> interchangeable nodes (duplicates.hpp Hasher)
teq::TensMapT<std::string> hash_graph(const teq::TensptrsT& roots);

/// give duplicates one owner; returns the (possibly replaced) roots and the number of nodes removed
teq::TensptrsT merge_dups(teq::TensptrsT roots, size_t* merged = nullptr);

/// functors over constants only -> constants (needs a device: the value is computed by the evaluator)
teq::TensptrsT fold_constants(teq::TensptrsT roots, size_t* folded = nullptr);
/// the functors `fold_constants` would evaluate and replace by constant leaves (host-side decision only; needs no device)
teq::TensptrsT fold_candidates(const teq::TensptrsT& roots);

/// hone::optimize without a rule file: merge_dups, then constant folding + merging to a fixed point
teq::TensptrsT optimize(teq::TensptrsT roots, Stats* stats = nullptr, bool fold = true);

}  // namespace hone

#endif  // TCR_HOST_HONE_HPP
