// planner.hpp — compiled evaluation: the functor DAG under a target set is lowered once
// into a linear launch plan (views resolved, elementwise / EXTEND / constant chains fused
// into register-machine programs, bias + activation folded into GEMM epilogues, trailing
// PERMUTE absorbed into GEMM output strides), buffers are assigned from the arena with
// liveness-based reuse, and the whole step is captured into a CUDA graph and replayed.
//
// It implements the same teq::iEvaluator entry point as the reference's Evaluator
// (internal/teq/evaluator.hpp:49-62) and is installed through the same slot
// (teq::set_eval, evaluator.hpp:65): callers do not change. What it removes is the
// per-node host work of the reference's hot loop — hash-set lookups, vector<shared_ptr>
// copies, std::function indirection and malloc/free per intermediate per step
// (internal/teq/evaluator.hpp:34-43, internal/eigen/device.hpp:304-328).
#ifndef TCR_HOST_PLANNER_HPP
#define TCR_HOST_PLANNER_HPP

#include "eteq.hpp"

namespace cuda {

struct PlanStats {
  size_t nodes = 0;     // functors covered by the plan
  size_t steps = 0;     // launch steps after fusion
  size_t launches = 0;  // kernel launches of one run
  bool graph = false;   // replayed as a CUDA graph
  size_t cached = 0;    // plans alive in the cache
  size_t steps_run = 0;      // steps launched by the last evaluate() (0: nothing was stale)
  size_t partial_runs = 0;   // evaluations of this process that re-ran only the stale part of a plan
};

PlanStats last_plan_stats();

/// per-step device timings of the most recently evaluated plan (steps launched one by one with
/// CUDA events on the library stream; runs the step sequence once more, eagerly)
struct StepTiming {
  std::string what;   // opcode of the node the step produces (+ "fused(n)" for register-machine programs)
  std::string shape;  // output shape
  double ms = 0;
  size_t bytes = 0;   // algorithmic bytes (inputs un-broadcast + output) for HBM-bound steps
};
std::vector<StepTiming> profile_last_plan(int repeats = 5);

/// lowering only (no device needed): the launch steps a plan for `targets` would consist of, in order
std::vector<std::string> describe_plan(const teq::TensSetT& targets);

struct PlanCache;

struct PlanEvaluator final : public teq::iEvaluator {
  PlanEvaluator();
  ~PlanEvaluator();
  void evaluate(teq::iDevice& device, const teq::TensSetT& targets, const teq::TensSetT& ignored = {}) override;
  void drop_plans();

 private:
  std::unique_ptr<PlanCache> cache_;
};

/// destroy every cached plan (and its captured CUDA graph) of every live PlanEvaluator.  Must
/// precede tcr_comm_destroy: NCCL waits for graphs holding captured collectives to be destroyed.
void drop_all_plans();

}  // namespace cuda

#endif  // TCR_HOST_PLANNER_HPP
