// dp.hpp — batch-sharded data parallelism: one process per GPU, identical graphs with
// local-batch shapes, SUM all-reduce of the gradients between tcr::derive and ASSIGN_*.
//
// This replaces, for that one case, the reference's only distributed back end — graph-node
// sharing over gRPC + Consul (tenncor/distr/p2p.hpp:39-138; DistrOpService::evaluate,
// tenncor/eteq/opsvc/service.hpp:110-160) — with NCCL over NVLink.
// A gradient is marked by wrapping it in an IDENTITY functor that carries the
// "dp_allreduce" attribute (value = post-reduction scale: 1/nranks for batch-mean losses,
// 1 for summed losses); the holder / planner all-reduces when it evaluates that node.
#ifndef TCR_HOST_DP_HPP
#define TCR_HOST_DP_HPP

#include "layr.hpp"

namespace dp {

const std::string allreduce_attr = "dp_allreduce";

/// NCCL unique id (rank 0 creates it; the launcher ships the bytes to the other ranks)
std::string unique_id();
/// join the communicator; mean_reduce: scale reduced gradients by 1/nranks
void init(int rank, int nranks, const std::string& id, bool mean_reduce);
void shutdown();
int rank();
int size();
bool active();
double scale();
/// gradients derived from now on are scaled by 1/nranks after the exchange (batch-mean losses) or left as sums
void set_mean_reduce(bool mean_reduce);

/// identity when no group is active; otherwise each gradient is wrapped in a marked IDENTITY
layr::ETensorsT wrap_gradients(const layr::ETensorsT& grads);
/// scale of a marked node, or 0 when `func` is not marked
double allreduce_scale(const teq::iFunctor& func);

/// [offset, offset + count) of `total` samples owned by `rank` (even split, remainder to the low ranks)
std::pair<size_t, size_t> shard(size_t total, int rank, int nranks);

}  // namespace dp

#endif  // TCR_HOST_DP_HPP
