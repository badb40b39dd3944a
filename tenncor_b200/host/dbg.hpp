// dbg.hpp — evaluator plugins and per-opcode profiling.
//
// `PlugableEvaluator` / `iPlugin` mirror dbg/peval/plugin_eval.hpp:9-46: the reference traversal (one device.calc per
// functor) followed by every plugin's `process(targets, visited)`; `Inspector` mirrors dbg/peval/stats/inspect.hpp:33-56
// (min / max of chosen functors). `OpProfiler` is the B200 addition the survey asks for: the same plugin slot
// (teq::set_eval) but every functor's kernels are bracketed by CUDA events on the library stream, accumulated per
// opcode with the algorithmic bytes moved, so `report()` reads as time share and GB/s per opcode.
#ifndef TCR_HOST_DBG_HPP
#define TCR_HOST_DBG_HPP

#include <functional>
#include <map>

#include "eteq.hpp"

namespace dbg {

struct iPlugin {
  virtual ~iPlugin() = default;
  virtual void process(const teq::TensSetT& targets, const teq::TensSetT& visited) = 0;
};

struct PlugableEvaluator final : public teq::iEvaluator {
  void evaluate(teq::iDevice& device, const teq::TensSetT& targets, const teq::TensSetT& ignored = {}) override;
  void add_plugin(std::shared_ptr<iPlugin> plugin) { plugins_.push_back(std::move(plugin)); }
  std::vector<std::shared_ptr<iPlugin>> plugins_;
};

struct Inspector final : public iPlugin {
  void process(const teq::TensSetT& targets, const teq::TensSetT& visited) override;
  void add(const teq::TensptrT& target, const std::string& label);
  std::unordered_map<teq::iFunctor*, std::string> insps_;
  std::map<std::string, std::pair<double, double>> last_;  // label -> (min, max) of the latest evaluation
};

struct OpStat {
  size_t calls = 0;
  double ms = 0;
  size_t bytes = 0;  // output + arguments, elements x element size (views count their referent once)
};

struct OpProfiler final : public teq::iEvaluator {
  void evaluate(teq::iDevice& device, const teq::TensSetT& targets, const teq::TensSetT& ignored = {}) override;
  void reset() { stats_.clear(); }
  std::map<std::string, OpStat> stats_;  // by opcode name
};

}  // namespace dbg

#endif  // TCR_HOST_DBG_HPP
