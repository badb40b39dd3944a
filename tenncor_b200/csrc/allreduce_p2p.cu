// allreduce_p2p.cu — gradient all-reduce over NVLink / NVSwitch peer memory (one process per GPU, one node).
//
// What it replaces: for batch-sharded training the reference's only distributed path ships tensors between processes over
// gRPC (tenncor/distr, tenncor/eteq/opsvc/service.hpp:110-160); round 1 replaced that exchange with one ncclAllReduce over
// the flat gradient bucket. For the bucket sizes of the demo models (C3: 3.26 MB, C4: 19.4 MB, C1 / C5: < 1 KB) that
// call is latency-bound — 73 us at 8 ranks for 3.26 MB, the whole loss of weak-scaling efficiency. Here every rank maps
// every peer's bucket (CUDA IPC) and ONE kernel does the exchange with plain loads and stores through NVSwitch:
//   one-shot (<= 512 KB) : barrier; every rank reads all peers' buckets and sums them in rank order; barrier
//   two-shot             : barrier; rank r pulls slice r of every bucket, sums it and pushes the result into every bucket
//                          (reduce-scatter by loads, all-gather by stores); barrier
// Sums run in rank order on every rank, so all ranks hold bit-identical results (replicated optimiser state stays consistent)
// and repeated runs are deterministic. The post-reduction scale (1 / ranks for batch-mean losses) is applied in the same pass.
// Barriers are per-block flag exchanges in the peers' memory (st / ld at system scope), epochs live in device memory so a
// captured CUDA graph replays correctly. Buffers must come from the symmetric region (tcr_comm_symm_alloc): every rank
// allocates in the same order, so an offset names the same gradient everywhere.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace tcr {

namespace {

constexpr int MAX_RANKS = 8;
constexpr int MAX_BLOCKS = 128;
constexpr int AR_THREADS = 512;
constexpr size_t SIGNAL_BYTES = 64 * 1024;  // head of the symmetric region (sizeof(Signal) = 12.5 KB)

struct Signal {
  uint32_t phase[3][MAX_BLOCKS][MAX_RANKS];  // phase[q][b][r]: rank r has passed barrier q of its block b (epoch value)
  uint32_t epoch[MAX_BLOCKS];                // last epoch block b of THIS rank used
};

struct P2PState {
  bool ready = false;
  int rank = 0, size = 1;
  char* local = nullptr;
  size_t bytes = 0, used = 0;
  char* peer[MAX_RANKS] = {nullptr};
};
P2PState g_p2p;

struct ArParams {
  float* data[MAX_RANKS];   // this buffer in every rank's memory (data[rank] is local)
  Signal* sig[MAX_RANKS];
  int rank, size;
  int64_t n;                // floats
  float scale;
};

__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// every rank's block b meets here: thread t < size tells peer t and waits for peer t
__device__ __forceinline__ void p2p_barrier(const ArParams& p, int q, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < p.size) {
    __threadfence_system();  // my writes (this kernel's and earlier ones') before the signal
    st_flag(&p.sig[threadIdx.x]->phase[q][blockIdx.x][p.rank], epoch);
    const uint32_t* mine = &p.sig[p.rank]->phase[q][blockIdx.x][threadIdx.x];
    uint32_t spins = 0;
    while (ld_flag(mine) != epoch)
      if (++spins > (1u << 30)) __trap();  // a peer died: fail instead of hanging the GPU
  }
  __syncthreads();
}

__device__ __forceinline__ float4 ld_peer(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

template <bool TWO_SHOT>
__global__ void __launch_bounds__(AR_THREADS) allreduce_p2p_kernel(const __grid_constant__ ArParams p) {
  TCR_PDL_ENTER();
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = p.sig[p.rank]->epoch[blockIdx.x] + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int64_t nvec = p.n / 4;  // n % 4 == 0 (the bucket is padded)
  const int64_t tid = (int64_t)blockIdx.x * AR_THREADS + threadIdx.x, nthr = (int64_t)gridDim.x * AR_THREADS;
  p2p_barrier(p, 0, epoch);  // every rank's gradients are in its bucket
  if (!TWO_SHOT) {
    // at most one vector per thread (the host sizes the grid): read everybody, wait until everybody has read, store
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < nvec) {
      acc = ld_peer(p.data[0] + 4 * tid);
      for (int r = 1; r < p.size; ++r) {
        const float4 v = ld_peer(p.data[r] + 4 * tid);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
    }
    p2p_barrier(p, 1, epoch);  // no peer reads my bucket any more
    if (tid < nvec) *reinterpret_cast<float4*>(p.data[p.rank] + 4 * tid) = acc;
  } else {
    // rank r owns slice r: it PULLS that slice from every bucket (all ranks' loads of a vector are in flight together: one trip
    // through NVSwitch, ~1-2 us, per U vectors instead of one per rank), sums in rank order, and PUSHES the result into every
    // bucket — stores need no round trip, so the all-gather costs no second latency and no third barrier: an element of slice r in
    // a peer's bucket is read and then overwritten by the same thread of rank r, nobody else touches it.
    const int64_t per = (nvec + p.size - 1) / p.size;  // vectors per slice
    const int64_t lo = per * p.rank, hi = lo + per < nvec ? lo + per : nvec;
    constexpr int U = 2;
    for (int64_t i0 = lo + tid; i0 < hi; i0 += nthr * U) {
      float4 v[MAX_RANKS][U];
#pragma unroll
      for (int r = 0; r < MAX_RANKS; ++r) {
        if (r < p.size) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * nthr;
            v[r][u] = i < hi ? ld_peer(p.data[r] + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 acc = v[0][u];
#pragma unroll
        for (int r = 1; r < MAX_RANKS; ++r)
          if (r < p.size) { acc.x += v[r][u].x; acc.y += v[r][u].y; acc.z += v[r][u].z; acc.w += v[r][u].w; }
        acc = make_float4(acc.x * p.scale, acc.y * p.scale, acc.z * p.scale, acc.w * p.scale);
        const int64_t i = i0 + u * nthr;
        if (i < hi) {
#pragma unroll
          for (int r = 0; r < MAX_RANKS; ++r)
            if (r < p.size) *reinterpret_cast<float4*>(p.data[r] + 4 * i) = acc;
        }
      }
    }
    p2p_barrier(p, 1, epoch);  // every slice has arrived in every bucket
  }
  if (threadIdx.x == 0) p.sig[p.rank]->epoch[blockIdx.x] = epoch;
}

}  // namespace

bool pdl_enabled();

// gathers `bytes` from every rank into out[rank * bytes ..] over the NCCL communicator (runtime.cu)
int nccl_allgather_bytes(const void* mine, void* all, size_t bytes);

int p2p_setup(int rank, int nranks) {
  P2PState& s = g_p2p;
  if (s.ready || nranks < 2 || nranks > MAX_RANKS) return TCR_OK;
  static const int enabled = std::getenv("TCR_P2P_ALLREDUCE") ? std::atoi(std::getenv("TCR_P2P_ALLREDUCE")) : 1;
  if (!enabled) return TCR_OK;
  const size_t region = (size_t)(std::getenv("TCR_P2P_REGION_MB") ? std::atoll(std::getenv("TCR_P2P_REGION_MB")) : 256) << 20;
  // every rank takes part in every exchange below, whatever happened locally: nobody may be left waiting in a collective
  auto all_agree = [&](bool mine_ok, int& rc) {
    int32_t flag = mine_ok ? 1 : 0;
    std::vector<int32_t> flags(nranks);
    rc = nccl_allgather_bytes(&flag, flags.data(), sizeof(flag));
    bool ok = rc == TCR_OK;
    for (int r = 0; r < nranks && ok; ++r) ok = flags[r] == 1;
    return ok;
  };
  void* base = nullptr;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  bool local_ok = cudaMalloc(&base, region) == cudaSuccess;
  if (local_ok) local_ok = cudaMemset(base, 0, region) == cudaSuccess && cudaIpcGetMemHandle(&mine, base) == cudaSuccess;
  cudaGetLastError();
  std::vector<cudaIpcMemHandle_t> all(nranks);
  int rc = nccl_allgather_bytes(&mine, all.data(), sizeof(mine));
  int rc2 = TCR_OK;
  bool ok = all_agree(local_ok && rc == TCR_OK, rc2);
  if (ok) {
    bool opened = true;
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) { s.peer[r] = (char*)base; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = false; break; }
      s.peer[r] = (char*)p;
    }
    ok = all_agree(opened, rc2);
  }
  if (!ok) {  // no peer path (no IPC, no peer access, ...): NCCL stays
    for (int r = 0; r < nranks; ++r)
      if (r != rank && s.peer[r]) cudaIpcCloseMemHandle(s.peer[r]);
    if (base) cudaFree(base);
    cudaGetLastError();
    memset(s.peer, 0, sizeof(s.peer));
    return TCR_OK;
  }
  s.local = (char*)base;
  s.bytes = region;
  s.used = SIGNAL_BYTES;
  s.rank = rank;
  s.size = nranks;
  s.ready = true;
  return TCR_OK;
}

void p2p_teardown() {
  P2PState& s = g_p2p;
  if (!s.ready) return;
  cudaDeviceSynchronize();
  for (int r = 0; r < s.size; ++r)
    if (r != s.rank && s.peer[r]) cudaIpcCloseMemHandle(s.peer[r]);
  cudaFree(s.local);
  s = P2PState();
}

// returns true when the peer kernel took the exchange
bool p2p_allreduce(void* buf, int64_t n, int dtype, double scale, int* rc) {
  P2PState& s = g_p2p;
  *rc = TCR_OK;
  if (!s.ready || dtype != TCR_FLOAT) return false;
  char* b = (char*)buf;
  if (b < s.local + SIGNAL_BYTES || b + n * 4 > s.local + s.bytes || (((uintptr_t)b) & 15) != 0) return false;
  const int64_t n4 = (n + 3) / 4 * 4;  // symmetric allocations are padded to 16 bytes
  ArParams p;
  memset(&p, 0, sizeof(p));
  const size_t off = (size_t)(b - s.local);
  for (int r = 0; r < s.size; ++r) {
    p.data[r] = reinterpret_cast<float*>(s.peer[r] + off);
    p.sig[r] = reinterpret_cast<Signal*>(s.peer[r]);
  }
  p.rank = s.rank;
  p.size = s.size;
  p.n = n4;
  p.scale = (float)scale;
  const int64_t nvec = n4 / 4;
  const bool one_shot = nvec <= (int64_t)MAX_BLOCKS * AR_THREADS && n4 * 4 <= 512 * 1024;
  int blocks;
  if (one_shot) {
    blocks = (int)ceil_div(nvec, AR_THREADS);  // at most one vector per thread: the mid barrier sits inside the loop
    if (blocks < 1) blocks = 1;
    TCR_LAUNCH((allreduce_p2p_kernel<false>), blocks, AR_THREADS, 0, p);
  } else {
    const int64_t per = ceil_div(nvec, s.size);
    blocks = (int)ceil_div(per, (int64_t)AR_THREADS * 2);
    if (blocks > MAX_BLOCKS) blocks = MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    TCR_LAUNCH((allreduce_p2p_kernel<true>), blocks, AR_THREADS, 0, p);
  }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) *rc = fail_cuda(e, "allreduce_p2p launch", __FILE__, __LINE__);
  return true;
}

}  // namespace tcr

using namespace tcr;

extern "C" {

int tcr_comm_symm_alloc(void** out, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(out != nullptr, "tcr_comm_symm_alloc: null out");
  *out = nullptr;
  P2PState& s = g_p2p;
  if (!s.ready) return TCR_ERR_UNSUPPORTED;
  const size_t need = (bytes + 255) / 256 * 256;
  if (s.used + need > s.bytes) return TCR_ERR_UNSUPPORTED;  // every rank asks for the same sizes in the same order: all fall back together
  *out = s.local + s.used;
  s.used += need;
  return TCR_OK;
}

int tcr_comm_symm_reset(void) {
  if (g_p2p.ready) g_p2p.used = SIGNAL_BYTES;
  return TCR_OK;
}

int tcr_comm_p2p_ready(void) { return g_p2p.ready ? 1 : 0; }

}  // extern "C"
