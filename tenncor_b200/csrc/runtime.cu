// runtime.cu — device selection, stream, arena allocator, copies, events, CUDA graphs
// and the NCCL communicator behind the C-ABI of include/tcr_b200.h.
//
// The arena replaces eigen::RuntimeMemory (malloc/free per intermediate per step,
// internal/eigen/memory.hpp:26-37,101-115): blocks are size-bucketed and recycled in
// stream order, so steady-state training performs zero cudaMalloc calls and captured
// CUDA graphs see stable addresses.
#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <cmath>

#include "common.cuh"

namespace tcr {

static thread_local std::string g_err;

State& state() {
  static State* s = new State();  // leaked on purpose: holders may outlive static destruction at exit
  return *s;
}

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return TCR_ERR_CUDA;
}

bool pdl_enabled() {
  // opt-in (TCR_PDL=1): measured gain is ~1 us per dependent launch in a captured chain (C4: 8.91 -> 8.83 ms per step), and
  // Nsight Compute 2025.2 crashes the application (SIGSEGV) when kernels carry the programmatic-serialization attribute
  static const bool on = std::getenv("TCR_PDL") && std::atoi(std::getenv("TCR_PDL")) != 0;
  return on;
}

// ---------------------------------------------------------------- ticket counters
// Self-resetting per-tile arrival counters for kernels that finish a split reduction in their last
// CTA. A launch takes a fresh range of a zero-initialised ring; ranges baked into captured graphs stay
// valid because every user leaves its counters at zero. Two launches share a range only after the
// ring wraps (2^20 counters), and launches of different plans are ordered on the library stream.
static int* g_counter_ring = nullptr;
static size_t g_counter_pos = 0;
static std::mutex g_counter_mu;
constexpr size_t COUNTER_RING = 1u << 20;

int* counter_ring_take(int n) {
  if (g_counter_ring == nullptr) return nullptr;
  std::lock_guard<std::mutex> lk(g_counter_mu);
  if (n < 1) n = 1;
  if (g_counter_pos + (size_t)n > COUNTER_RING) g_counter_pos = 0;
  int* p = g_counter_ring + g_counter_pos;
  g_counter_pos += (size_t)n;
  return p;
}

// ---------------------------------------------------------------- arena
struct Arena {
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;        // bucket size -> block
  std::unordered_map<void*, size_t> live;          // block -> bucket size
  size_t in_use = 0, reserved = 0, n_mallocs = 0;
  // While a CUDA graph is being captured, freed blocks stay reserved for that graph
  // (its kernels keep their addresses): they are recycled only inside the capture and
  // return to the global pool when the graph is destroyed.
  bool capturing = false;
  // one pool per capture lane: two lanes may run concurrently inside the graph, so a scratch
  // block freed on one lane must not be handed to a kernel on another
  std::multimap<size_t, void*> capture_pools[TCR_GRAPH_LANES];
  int lane = 0;
  std::multimap<size_t, void*>& capture_pool() { return capture_pools[lane]; }
  std::unordered_map<void*, std::vector<std::pair<size_t, void*>>> graph_blocks;
  std::unordered_map<void*, uint64_t> graph_kernels;  // kernel nodes per instantiated graph

  static size_t bucket(size_t bytes) {
    if (bytes < 512) return 512;
    if (bytes <= (1u << 20)) {  // next power of two up to 1 MiB
      size_t b = 512;
      while (b < bytes) b <<= 1;
      return b;
    }
    const size_t g = 2u << 20;  // 2 MiB granules above (HBM page size)
    return (bytes + g - 1) / g * g;
  }
};

static Arena& arena() {
  static Arena* a = new Arena();  // leaked on purpose (see state())
  return *a;
}

}  // namespace tcr

using namespace tcr;

extern "C" {

const char* tcr_last_error(void) { return g_err.c_str(); }

int tcr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int tcr_init(int device) {
  State& s = state();
  if (s.ready && s.device == device) return TCR_OK;
  // the stream, the counter ring, the prefetch stream and every arena block belong to the device of the first init
  TCR_ARG(!s.ready, "tcr_init: already initialised on device %d; call tcr_shutdown() before switching to device %d", s.device, device);
  int n = tcr_device_count();
  if (n <= 0) {
    set_error("tcr_init: no CUDA device visible (this back end has no CPU fallback)");
    return TCR_ERR_NODEVICE;
  }
  TCR_ARG(device >= 0 && device < n, "tcr_init: device %d out of range [0,%d)", device, n);
  TCR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TCR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("tcr_init: device %d is sm_%d%d; libtcr_b200 is built for sm_100a only", device,
              prop.major, prop.minor);
    return TCR_ERR_NODEVICE;
  }
  if (s.stream == nullptr) TCR_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  if (g_counter_ring == nullptr) {
    TCR_CUDA(cudaMalloc(&g_counter_ring, COUNTER_RING * sizeof(int)));
    TCR_CUDA(cudaMemset(g_counter_ring, 0, COUNTER_RING * sizeof(int)));
  }
  s.device = device;
  s.sm_count = prop.multiProcessorCount;
  s.ready = true;
  return TCR_OK;
}

// input prefetch state (tcr_h2d_prefetch / tcr_prefetch_commit below)
static cudaStream_t g_copy_stream = nullptr;
static cudaEvent_t g_copy_done = nullptr, g_commit_done = nullptr;
static bool g_copy_pending = false, g_commit_recorded = false;

int tcr_shutdown(void) {
  State& s = state();
  if (!s.ready) return TCR_OK;
  tcr_comm_destroy();
  cudaStreamSynchronize(s.stream);
  if (g_copy_stream) {
    cudaStreamSynchronize(g_copy_stream);
    cudaStreamDestroy(g_copy_stream);
    cudaEventDestroy(g_copy_done);
    cudaEventDestroy(g_commit_done);
    g_copy_stream = nullptr;
    g_copy_pending = g_commit_recorded = false;
  }
  tcr_arena_trim();
  if (g_counter_ring) {
    cudaFree(g_counter_ring);
    g_counter_ring = nullptr;
    g_counter_pos = 0;
  }
  cudaStreamDestroy(s.stream);
  s.stream = nullptr;
  s.ready = false;
  return TCR_OK;
}

int tcr_sm_count(void) { return state().sm_count; }
void* tcr_stream(void) { return (void*)state().stream; }
uint64_t tcr_launch_count(void) { return state().launches.load(); }

int tcr_sync(void) {
  TCR_REQUIRE_DEVICE();
  TCR_CUDA(cudaStreamSynchronize(state().stream));
  return TCR_OK;
}

int tcr_alloc(void** out, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(out != nullptr, "tcr_alloc: null out");
  Arena& a = arena();
  size_t b = Arena::bucket(bytes);
  std::lock_guard<std::mutex> lk(a.mu);
  void* p = nullptr;
  auto& cpool = a.capture_pool();
  auto cit = a.capturing ? cpool.find(b) : cpool.end();
  auto it = a.free_blocks.find(b);
  if (cit != cpool.end()) {
    p = cit->second;
    cpool.erase(cit);
  } else if (it != a.free_blocks.end()) {
    p = it->second;
    a.free_blocks.erase(it);
  } else {
    cudaError_t e = cudaMalloc(&p, b);
    if (e != cudaSuccess) {
      // return cached blocks to the driver and retry once
      cudaGetLastError();
      if (!a.capturing) cudaStreamSynchronize(state().stream);
      for (auto& kv : a.free_blocks) {
        cudaFree(kv.second);
        a.reserved -= kv.first;
      }
      a.free_blocks.clear();
      e = cudaMalloc(&p, b);
      if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc", __FILE__, __LINE__);
    }
    a.reserved += b;
    a.n_mallocs += 1;
  }
  a.live[p] = b;
  a.in_use += b;
  *out = p;
  return TCR_OK;
}

int tcr_free(void* ptr) {
  if (ptr == nullptr) return TCR_OK;
  Arena& a = arena();
  std::lock_guard<std::mutex> lk(a.mu);
  auto it = a.live.find(ptr);
  TCR_ARG(it != a.live.end(), "tcr_free: pointer %p was not allocated by tcr_alloc", ptr);
  if (a.capturing) a.capture_pool().emplace(it->second, ptr);
  else a.free_blocks.emplace(it->second, ptr);
  a.in_use -= it->second;
  a.live.erase(it);
  return TCR_OK;
}

int tcr_arena_stats(size_t* bytes_in_use, size_t* bytes_reserved, size_t* n_device_mallocs) {
  Arena& a = arena();
  std::lock_guard<std::mutex> lk(a.mu);
  if (bytes_in_use) *bytes_in_use = a.in_use;
  if (bytes_reserved) *bytes_reserved = a.reserved;
  if (n_device_mallocs) *n_device_mallocs = a.n_mallocs;
  return TCR_OK;
}

int tcr_arena_trim(void) {
  Arena& a = arena();
  std::lock_guard<std::mutex> lk(a.mu);
  if (state().stream) cudaStreamSynchronize(state().stream);
  for (auto& kv : a.free_blocks) {
    cudaFree(kv.second);
    a.reserved -= kv.first;
  }
  a.free_blocks.clear();
  return TCR_OK;
}

int tcr_host_alloc(void** out, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  TCR_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return TCR_OK;
}

int tcr_host_free(void* ptr) {
  if (ptr) TCR_CUDA(cudaFreeHost(ptr));
  return TCR_OK;
}

int tcr_h2d(void* dst, const void* host_src, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  if (bytes == 0) return TCR_OK;
  TCR_CUDA(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, state().stream));
  return TCR_OK;
}

// ---- input prefetch: the next batch crosses PCIe on a copy stream while the current step computes

int tcr_h2d_prefetch(void* staging, const void* host_src, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(staging && host_src, "tcr_h2d_prefetch: null argument");
  if (g_copy_stream == nullptr) {
    TCR_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
    TCR_CUDA(cudaEventCreateWithFlags(&g_copy_done, cudaEventDisableTiming));
    TCR_CUDA(cudaEventCreateWithFlags(&g_commit_done, cudaEventDisableTiming));
  }
  // the staging buffer may still be the source of the previous commit's device copy
  if (g_commit_recorded) TCR_CUDA(cudaStreamWaitEvent(g_copy_stream, g_commit_done, 0));
  if (bytes) TCR_CUDA(cudaMemcpyAsync(staging, host_src, bytes, cudaMemcpyHostToDevice, g_copy_stream));
  TCR_CUDA(cudaEventRecord(g_copy_done, g_copy_stream));
  g_copy_pending = true;
  return TCR_OK;
}

int tcr_prefetch_sync(void) {
  TCR_REQUIRE_DEVICE();
  if (g_copy_stream) TCR_CUDA(cudaStreamSynchronize(g_copy_stream));
  return TCR_OK;
}

int tcr_prefetch_commit(void* dst, const void* staging, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(dst && staging, "tcr_prefetch_commit: null argument");
  TCR_ARG(g_copy_pending, "tcr_prefetch_commit: nothing was prefetched");
  TCR_CUDA(cudaStreamWaitEvent(state().stream, g_copy_done, 0));
  if (bytes) TCR_CUDA(cudaMemcpyAsync(dst, staging, bytes, cudaMemcpyDeviceToDevice, state().stream));
  TCR_CUDA(cudaEventRecord(g_commit_done, state().stream));
  g_commit_recorded = true;
  return TCR_OK;
}

int tcr_d2h(void* host_dst, const void* src, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  if (bytes == 0) return TCR_OK;
  TCR_CUDA(cudaMemcpyAsync(host_dst, src, bytes, cudaMemcpyDeviceToHost, state().stream));
  return TCR_OK;
}

int tcr_d2d(void* dst, const void* src, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  if (bytes == 0 || dst == src) return TCR_OK;
  TCR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, state().stream));
  return TCR_OK;
}

int tcr_memset(void* dst, int byte, size_t bytes) {
  TCR_REQUIRE_DEVICE();
  if (bytes == 0) return TCR_OK;
  TCR_CUDA(cudaMemsetAsync(dst, byte, bytes, state().stream));
  return TCR_OK;
}

int tcr_event_create(void** out) {
  TCR_REQUIRE_DEVICE();
  cudaEvent_t ev;
  TCR_CUDA(cudaEventCreate(&ev));
  *out = (void*)ev;
  return TCR_OK;
}

int tcr_event_destroy(void* ev) {
  if (ev) TCR_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return TCR_OK;
}

int tcr_event_record(void* ev) {
  TCR_REQUIRE_DEVICE();
  TCR_CUDA(cudaEventRecord((cudaEvent_t)ev, state().stream));
  return TCR_OK;
}

int tcr_event_sync(void* ev) {
  TCR_REQUIRE_DEVICE();
  TCR_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return TCR_OK;
}

int tcr_event_elapsed_ms(void* start, void* stop, float* ms) {
  TCR_REQUIRE_DEVICE();
  TCR_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  TCR_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return TCR_OK;
}

// ---------------------------------------------------------------- graphs
struct Lanes {
  cudaStream_t origin = nullptr;
  cudaStream_t side[TCR_GRAPH_LANES] = {};
  bool forked[TCR_GRAPH_LANES] = {};
  cudaEvent_t root = nullptr;
  std::vector<cudaEvent_t> events;
};
static Lanes& lanes() {
  static Lanes* l = new Lanes();
  return *l;
}

int tcr_graph_begin(void) {
  TCR_REQUIRE_DEVICE();
  Arena& a = arena();
  TCR_ARG(!a.capturing, "tcr_graph_begin: a capture is already in progress");
  Lanes& l = lanes();
  l.origin = state().stream;
  TCR_CUDA(cudaStreamBeginCapture(l.origin, cudaStreamCaptureModeRelaxed));
  for (int i = 0; i < TCR_GRAPH_LANES; ++i) l.forked[i] = false;
  l.forked[0] = true;
  l.root = nullptr;
  {  // side lanes join the capture by waiting on this mark of the origin stream
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) {
      l.events.push_back(ev);
      if (cudaEventRecord(ev, l.origin) == cudaSuccess) l.root = ev;
    }
  }
  std::lock_guard<std::mutex> lk(a.mu);
  a.capturing = true;
  a.lane = 0;
  return TCR_OK;
}

// ---- capture lanes: independent steps of one plan are captured on side streams so the
// instantiated graph has parallel branches (kernels that fill a fraction of the SMs overlap)
int tcr_graph_lane(int lane) {
  TCR_REQUIRE_DEVICE();
  Arena& a = arena();
  Lanes& l = lanes();
  TCR_ARG(a.capturing, "tcr_graph_lane: no capture in progress");
  TCR_ARG(lane >= 0 && lane < TCR_GRAPH_LANES, "tcr_graph_lane: lane %d out of range [0,%d)", lane, TCR_GRAPH_LANES);
  if (lane > 0) {
    if (l.side[lane] == nullptr) TCR_CUDA(cudaStreamCreateWithFlags(&l.side[lane], cudaStreamNonBlocking));
    if (!l.forked[lane]) {
      TCR_ARG(l.root != nullptr, "tcr_graph_lane: the capture has no root mark");
      TCR_CUDA(cudaStreamWaitEvent(l.side[lane], l.root, 0));
      l.forked[lane] = true;
    }
  }
  state().stream = lane == 0 ? l.origin : l.side[lane];
  std::lock_guard<std::mutex> lk(a.mu);
  a.lane = lane;
  return TCR_OK;
}

int tcr_graph_record(int* event_id) {
  TCR_REQUIRE_DEVICE();
  Lanes& l = lanes();
  TCR_ARG(arena().capturing && event_id != nullptr, "tcr_graph_record: no capture in progress");
  cudaEvent_t ev;
  TCR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  l.events.push_back(ev);
  TCR_CUDA(cudaEventRecord(ev, state().stream));
  *event_id = (int)l.events.size() - 1;
  return TCR_OK;
}

int tcr_graph_wait(int event_id) {
  TCR_REQUIRE_DEVICE();
  Lanes& l = lanes();
  TCR_ARG(arena().capturing, "tcr_graph_wait: no capture in progress");
  TCR_ARG(event_id >= 0 && event_id < (int)l.events.size(), "tcr_graph_wait: unknown mark %d", event_id);
  TCR_CUDA(cudaStreamWaitEvent(state().stream, l.events[event_id], 0));
  return TCR_OK;
}

int tcr_graph_end(void** out_exec) {
  TCR_REQUIRE_DEVICE();
  Arena& a = arena();
  std::vector<std::pair<size_t, void*>> held;
  Lanes& l = lanes();
  cudaError_t join_err = cudaSuccess;
  if (l.origin != nullptr) {
    // every side lane rejoins the origin stream before the capture ends
    for (int i = 1; i < TCR_GRAPH_LANES; ++i) {
      if (!l.forked[i]) continue;
      cudaEvent_t ev;
      cudaError_t je = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (je == cudaSuccess) {
        l.events.push_back(ev);
        je = cudaEventRecord(ev, l.side[i]);
        if (je == cudaSuccess) je = cudaStreamWaitEvent(l.origin, ev, 0);
      }
      if (je != cudaSuccess) join_err = je;
      l.forked[i] = false;
    }
    state().stream = l.origin;
  }
  {
    std::lock_guard<std::mutex> lk(a.mu);
    a.capturing = false;
    a.lane = 0;
    for (auto& pool : a.capture_pools) {
      for (auto& kv : pool) held.push_back({kv.first, kv.second});
      pool.clear();
    }
  }
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(state().stream, &graph);
  for (auto ev : l.events) cudaEventDestroy(ev);
  l.events.clear();
  l.root = nullptr;
  if (e == cudaSuccess && join_err != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    graph = nullptr;
    e = join_err;
  }
  cudaGraphExec_t exec = nullptr;
  uint64_t n_kernels = 0;
  if (e == cudaSuccess) {
    size_t n_nodes = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n_nodes) == cudaSuccess && n_nodes > 0) {
      std::vector<cudaGraphNode_t> gn(n_nodes);
      if (cudaGraphGetNodes(graph, gn.data(), &n_nodes) == cudaSuccess)
        for (auto node : gn) {
          cudaGraphNodeType ty;
          if (cudaGraphNodeGetType(node, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) ++n_kernels;
        }
    }
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
  }
  std::lock_guard<std::mutex> lk(a.mu);
  if (e != cudaSuccess) {
    for (auto& kv : held) a.free_blocks.emplace(kv.first, kv.second);
    return fail_cuda(e, "cudaStreamEndCapture/cudaGraphInstantiate", __FILE__, __LINE__);
  }
  a.graph_blocks[(void*)exec] = std::move(held);
  a.graph_kernels[(void*)exec] = n_kernels;
  *out_exec = (void*)exec;
  return TCR_OK;
}

int tcr_graph_launch(void* exec) {
  TCR_REQUIRE_DEVICE();
  TCR_CUDA(cudaGraphLaunch((cudaGraphExec_t)exec, state().stream));
  {
    Arena& a = arena();
    std::lock_guard<std::mutex> lk(a.mu);
    auto it = a.graph_kernels.find(exec);
    if (it != a.graph_kernels.end()) state().launches.fetch_add(it->second, std::memory_order_relaxed);
  }
  return TCR_OK;
}

int tcr_graph_destroy(void* exec) {
  if (!exec) return TCR_OK;
  if (state().stream) cudaStreamSynchronize(state().stream);
  TCR_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)exec));
  Arena& a = arena();
  std::lock_guard<std::mutex> lk(a.mu);
  auto it = a.graph_blocks.find(exec);
  if (it != a.graph_blocks.end()) {
    for (auto& kv : it->second) a.free_blocks.emplace(kv.first, kv.second);
    a.graph_blocks.erase(it);
  }
  a.graph_kernels.erase(exec);
  return TCR_OK;
}

// ---------------------------------------------------------------- NCCL (dlopen'd)
// NCCL is resolved at run time so that a process which already loaded torch's bundled
// libnccl.so.2 shares that copy, and a CPU-only process can still load this library.
typedef struct { char internal[TCR_COMM_ID_BYTES]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_comm_init)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_comm_destroy)(nccl_comm);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
typedef const char* (*fn_errstr)(int);

static struct {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_comm_init comm_init = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_allgather allgather = nullptr;
  fn_errstr errstr = nullptr;
  nccl_comm comm = nullptr;
  int rank = 0, size = 1;
} g_nccl;

static int nccl_load() {
  if (g_nccl.lib) return TCR_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    set_error("NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror());
    return TCR_ERR_NCCL;
  }
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.comm_init = (fn_comm_init)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.comm_init || !g_nccl.comm_destroy || !g_nccl.allreduce) {
    set_error("NCCL: missing symbols in libnccl");
    return TCR_ERR_NCCL;
  }
  return TCR_OK;
}

static int nccl_fail(int rc, const char* what) {
  set_error("NCCL error %d (%s) in %s", rc, g_nccl.errstr ? g_nccl.errstr(rc) : "?", what);
  return TCR_ERR_NCCL;
}

}  // extern "C" (reopened below)

namespace tcr {
int p2p_setup(int rank, int nranks);  // allreduce_p2p.cu
void p2p_teardown();
bool p2p_allreduce(void* buf, int64_t n, int dtype, double scale, int* rc);

// set-up helper: every rank contributes `bytes`, everybody receives all of them in rank order (host buffers)
int nccl_allgather_bytes(const void* mine, void* all, size_t bytes) {
  if (!g_nccl.comm || !g_nccl.allgather) { set_error("NCCL all-gather unavailable"); return TCR_ERR_NCCL; }
  const size_t n = (size_t)g_nccl.size;
  char *dsend = nullptr, *drecv = nullptr;
  TCR_CUDA(cudaMalloc(&dsend, bytes));
  TCR_CUDA(cudaMalloc(&drecv, bytes * n));
  TCR_CUDA(cudaMemcpy(dsend, mine, bytes, cudaMemcpyHostToDevice));
  int e = g_nccl.allgather(dsend, drecv, bytes, /*ncclInt8*/ 0, g_nccl.comm, state().stream);
  if (e) { cudaFree(dsend); cudaFree(drecv); return nccl_fail(e, "ncclAllGather"); }
  TCR_CUDA(cudaStreamSynchronize(state().stream));
  TCR_CUDA(cudaMemcpy(all, drecv, bytes * n, cudaMemcpyDeviceToHost));
  cudaFree(dsend);
  cudaFree(drecv);
  return TCR_OK;
}
}  // namespace tcr

extern "C" {

int tcr_comm_unique_id(char id[TCR_COMM_ID_BYTES]) {
  int rc = nccl_load();
  if (rc) return rc;
  nccl_uid uid;
  memset(&uid, 0, sizeof(uid));
  int e = g_nccl.get_uid(&uid);
  if (e) return nccl_fail(e, "ncclGetUniqueId");
  memcpy(id, uid.internal, TCR_COMM_ID_BYTES);
  return TCR_OK;
}

int tcr_comm_init(int rank, int nranks, const char id[TCR_COMM_ID_BYTES]) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "tcr_comm_init: bad rank %d / %d", rank, nranks);
  int rc = nccl_load();
  if (rc) return rc;
  if (g_nccl.comm) tcr_comm_destroy();
  nccl_uid uid;
  memcpy(uid.internal, id, TCR_COMM_ID_BYTES);
  int e = g_nccl.comm_init(&g_nccl.comm, nranks, uid, rank);
  if (e) return nccl_fail(e, "ncclCommInitRank");
  g_nccl.rank = rank;
  g_nccl.size = nranks;
  return p2p_setup(rank, nranks);  // peer-memory exchange path (allreduce_p2p.cu); NCCL remains the fallback
}

int tcr_comm_destroy(void) {
  if (g_nccl.comm) {
    if (state().stream) cudaStreamSynchronize(state().stream);
    p2p_teardown();
    g_nccl.comm_destroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  g_nccl.rank = 0;
  g_nccl.size = 1;
  return TCR_OK;
}

int tcr_comm_rank(void) { return g_nccl.rank; }
int tcr_comm_size(void) { return g_nccl.size; }

int tcr_scale_inplace(void* buf, int64_t n, int dtype, double scale);  // elementwise.cu

int tcr_allreduce_sum(void* buf, int64_t n, int dtype, double scale) {
  TCR_REQUIRE_DEVICE();
  if (n <= 0) return TCR_OK;
  if (g_nccl.comm != nullptr && g_nccl.size > 1) {
    int prc = TCR_OK;
    if (p2p_allreduce(buf, n, dtype, scale, &prc)) return prc;  // NVLink peer-memory kernel, scale included
    int nccl_type;
    switch (dtype) {  // ncclDataType_t
      case TCR_FLOAT: nccl_type = 7; break;
      case TCR_DOUBLE: nccl_type = 8; break;
      case TCR_INT32: nccl_type = 2; break;
      case TCR_INT64: nccl_type = 4; break;
      default: set_error("tcr_allreduce_sum: unsupported dtype %d", dtype); return TCR_ERR_DTYPE;
    }
    // mean over ranks (batch-normalised losses): NCCL's own averaging op, no second kernel
    const bool avg = (dtype == TCR_FLOAT || dtype == TCR_DOUBLE) && fabs(scale * g_nccl.size - 1.0) < 1e-12;
    int e = g_nccl.allreduce(buf, buf, (size_t)n, nccl_type, avg ? /*ncclAvg*/ 4 : /*ncclSum*/ 0, g_nccl.comm, state().stream);
    if (e) return nccl_fail(e, "ncclAllReduce");
    if (avg) return TCR_OK;
  }
  if (scale != 1.0) return tcr_scale_inplace(buf, n, dtype, scale);
  return TCR_OK;
}

}  // extern "C"
