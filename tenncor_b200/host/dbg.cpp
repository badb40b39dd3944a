// dbg.cpp — see dbg.hpp.
#include "dbg.hpp"

#include <algorithm>

namespace dbg {

using namespace teq;

void PlugableEvaluator::evaluate(iDevice& device, const TensSetT& targets, const TensSetT& ignored) {
  TravEvaluator eval(device, targets, ignored);
  for (auto t : targets) t->accept(eval);
  for (auto& plugin : plugins_) plugin->process(targets, eval.visited_);
}

void Inspector::add(const TensptrT& target, const std::string& label) {
  if (auto f = dynamic_cast<iFunctor*>(target.get())) insps_.emplace(f, label);
}

void Inspector::process(const TensSetT&, const TensSetT& visited) {
  for (auto vis : visited) {
    auto func = dynamic_cast<iFunctor*>(vis);
    if (!func) continue;
    auto it = insps_.find(func);
    if (it == insps_.end()) continue;
    const void* data = func->device().data();  // host mirror: a D2H sync point, like the reference's host read
    if (nullptr == data) {
      std::fprintf(stderr, "[error] cannot inspect null data of shape %s\n", func->shape().to_string().c_str());
      continue;
    }
    const size_t n = func->shape().n_elems();
    std::vector<double> d(n);
    egen::type_convert(d.data(), egen::DOUBLE, data, (egen::_GENERATED_DTYPE)func->get_meta().type_code(), n);
    auto mm = std::minmax_element(d.begin(), d.end());
    last_[it->second] = {*mm.first, *mm.second};
    std::fprintf(stderr, "[info] (%s) => min: %g, max: %g\n", it->second.c_str(), *mm.first, *mm.second);
  }
}

namespace {

struct TimedDevice final : public iDevice {
  explicit TimedDevice(iDevice& inner) : inner_(&inner) {}
  ~TimedDevice() {
    for (void* e : events_) tcr_event_destroy(e);
  }
  void calc(iTensor& tens, size_t cache_ttl) override {
    void *e0 = nullptr, *e1 = nullptr;
    cuda::check(tcr_event_create(&e0), "tcr_event_create");
    cuda::check(tcr_event_create(&e1), "tcr_event_create");
    events_.push_back(e0);
    events_.push_back(e1);
    cuda::check(tcr_event_record(e0), "tcr_event_record");
    inner_->calc(tens, cache_ttl);
    cuda::check(tcr_event_record(e1), "tcr_event_record");
    nodes_.push_back(&tens);
  }
  iDevice* inner_;
  std::vector<void*> events_;
  std::vector<iTensor*> nodes_;
};

}  // namespace

void OpProfiler::evaluate(iDevice& device, const TensSetT& targets, const TensSetT& ignored) {
  cuda::ensure_device();
  TimedDevice timed(device);
  TravEvaluator eval(timed, targets, ignored);
  for (auto t : targets) t->accept(eval);
  cuda::sync();
  for (size_t i = 0; i < timed.nodes_.size(); ++i) {
    auto f = dynamic_cast<iFunctor*>(timed.nodes_[i]);
    if (!f) continue;
    float ms = 0;
    cuda::check(tcr_event_elapsed_ms(timed.events_[2 * i], timed.events_[2 * i + 1], &ms), "tcr_event_elapsed_ms");
    OpStat& s = stats_[f->get_opcode().name_];
    s.calls += 1;
    s.ms += ms;
    s.bytes += f->shape().n_elems() * f->get_meta().type_size();
    for (auto& a : f->args_ref()) s.bytes += a->shape().n_elems() * a->get_meta().type_size();
  }
}

}  // namespace dbg
