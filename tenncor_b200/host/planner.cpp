// planner.cpp — (stage 1) delegates to the node-by-node evaluator; the fused plan lands next.
#include "planner.hpp"

namespace cuda {

static PlanStats g_stats;
PlanStats last_plan_stats() { return g_stats; }

struct PlanCache {};

PlanEvaluator::PlanEvaluator() : cache_(new PlanCache()) {}
PlanEvaluator::~PlanEvaluator() = default;

void PlanEvaluator::evaluate(teq::iDevice& device, const teq::TensSetT& targets, const teq::TensSetT& ignored) {
  teq::Evaluator fallback;
  fallback.evaluate(device, targets, ignored);
}

}  // namespace cuda
