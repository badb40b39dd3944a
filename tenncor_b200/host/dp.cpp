// dp.cpp — data-parallel group state and gradient marking.
#include "dp.hpp"
#include "planner.hpp"

namespace dp {

static int g_rank = 0, g_size = 1;
static bool g_mean = true;

std::string unique_id() {
  char id[TCR_COMM_ID_BYTES];
  cuda::check(tcr_comm_unique_id(id), "tcr_comm_unique_id");
  return std::string(id, id + TCR_COMM_ID_BYTES);
}

void init(int rank, int nranks, const std::string& id, bool mean_reduce) {
  if (nranks < 1 || rank < 0 || rank >= nranks) global::fatalf("bad data-parallel rank %d of %d", rank, nranks);
  if (nranks > 1) {
    if (id.size() != TCR_COMM_ID_BYTES) global::fatalf("NCCL id must be %d bytes, got %d", TCR_COMM_ID_BYTES, (int)id.size());
    cuda::ensure_device();
    cuda::check(tcr_comm_init(rank, nranks, id.data()), "tcr_comm_init");
  }
  g_rank = rank;
  g_size = nranks;
  g_mean = mean_reduce;
}

void shutdown() {
  cuda::drop_all_plans();  // captured graphs hold references on the communicator
  tcr_sync();
  tcr_comm_destroy();
  g_rank = 0;
  g_size = 1;
}

int rank() { return g_rank; }
int size() { return g_size; }
bool active() { return g_size > 1; }
double scale() { return g_mean ? 1.0 / g_size : 1.0; }
void set_mean_reduce(bool mean_reduce) { g_mean = mean_reduce; }

layr::ETensorsT wrap_gradients(const layr::ETensorsT& grads) {
  if (!active()) return grads;
  layr::ETensorsT out;
  for (auto& g : grads) {
    auto wrapped = eteq::make_functor(egen::IDENTITY, {g});
    auto f = static_cast<teq::iFunctor*>(wrapped.get());
    f->add_attr(allreduce_attr, std::make_unique<marsh::Float>(scale()));
    // the holder was created as an alias before the mark existed: rebuild it
    static_cast<eigen::Observable*>(f)->uninitialize();
    static_cast<eigen::Observable*>(f)->must_initialize();
    out.push_back(wrapped);
  }
  return out;
}

double allreduce_scale(const teq::iFunctor& func) {
  auto attr = dynamic_cast<const marsh::Float*>(func.get_attr(allreduce_attr));
  return attr ? attr->val_ : 0.0;
}

std::pair<size_t, size_t> shard(size_t total, int rank, int nranks) {
  size_t base = total / nranks, rem = total % nranks;
  size_t count = base + ((size_t)rank < rem ? 1 : 0);
  size_t offset = base * rank + std::min<size_t>(rank, rem);
  return {offset, count};
}

}  // namespace dp
