// common.cuh — shared state and helpers for the libtcr_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <utility>
#include <atomic>
#include <string>

#include "tcr_b200.h"

namespace tcr {

struct State {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  std::atomic<uint64_t> launches{0};
};

State& state();

void set_error(const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* what, const char* file, int line);

#define TCR_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return ::tcr::fail_cuda(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define TCR_REQUIRE_DEVICE()                                                   \
  do {                                                                        \
    if (!::tcr::state().ready) {                                              \
      ::tcr::set_error("tcr_b200: no CUDA device initialised (call tcr_init; there is no CPU fallback)"); \
      return TCR_ERR_NODEVICE;                                                \
    }                                                                         \
  } while (0)

#define TCR_ARG(cond, ...)                                                    \
  do {                                                                        \
    if (!(cond)) {                                                            \
      ::tcr::set_error(__VA_ARGS__);                                          \
      return TCR_ERR_ARG;                                                     \
    }                                                                         \
  } while (0)

// Programmatic dependent launch (PDL). Every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and begins with TCR_PDL_ENTER(): `griddepcontrol.wait` blocks until the
// preceding kernel of the stream (in a captured graph: the programmatic edge) has completed and its writes are visible, then
// `griddepcontrol.launch_dependents` lets the NEXT kernel's CTAs be scheduled while this one runs. Semantics are those of
// plain stream order; what is gained is the launch latency between dependent kernels (~2 us each, and the recurrent
// workloads chain hundreds of small launches per step). Kernels with a costly prologue that touches no memory (barrier
// set-up, TMEM allocation, tensor-map prefetch) place the wait after it. TCR_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define TCR_PDL_ENTER()              \
  do {                               \
    ::tcr::pdl_wait();               \
    ::tcr::pdl_launch_dependents();  \
  } while (0)

bool pdl_enabled();  // runtime.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = state().stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// every kernel launch goes through this so bench.py can report gpu_launches
#define TCR_LAUNCH(kernel, grid, block, smem, ...)                            \
  do {                                                                        \
    ::tcr::launch_kernel(kernel, (grid), (block), (smem), __VA_ARGS__);       \
    ::tcr::state().launches.fetch_add(1, std::memory_order_relaxed);          \
  } while (0)

#define TCR_CHECK_LAUNCH()                                                    \
  do {                                                                        \
    cudaError_t _e = cudaPeekAtLastError();                                   \
    if (_e != cudaSuccess) return ::tcr::fail_cuda(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

inline size_t dtype_size(int dtype) {
  switch (dtype) {
    case TCR_DOUBLE: case TCR_INT64: case TCR_UINT64: return 8;
    case TCR_FLOAT: case TCR_INT32: case TCR_UINT32: return 4;
    case TCR_INT16: case TCR_UINT16: return 2;
    case TCR_INT8: case TCR_UINT8: return 1;
    default: return 0;
  }
}

template <typename T> struct DTypeOf;
template <> struct DTypeOf<double> { static constexpr int value = TCR_DOUBLE; };
template <> struct DTypeOf<float> { static constexpr int value = TCR_FLOAT; };
template <> struct DTypeOf<int8_t> { static constexpr int value = TCR_INT8; };
template <> struct DTypeOf<uint8_t> { static constexpr int value = TCR_UINT8; };
template <> struct DTypeOf<int16_t> { static constexpr int value = TCR_INT16; };
template <> struct DTypeOf<uint16_t> { static constexpr int value = TCR_UINT16; };
template <> struct DTypeOf<int32_t> { static constexpr int value = TCR_INT32; };
template <> struct DTypeOf<uint32_t> { static constexpr int value = TCR_UINT32; };
template <> struct DTypeOf<int64_t> { static constexpr int value = TCR_INT64; };
template <> struct DTypeOf<uint64_t> { static constexpr int value = TCR_UINT64; };

// compute types with kernels: the reference's min type set (cfg/mintype.yml) plus int64
#define TCR_DISPATCH_COMPUTE(dtype, T, ...)                                    \
  switch (dtype) {                                                            \
    case TCR_FLOAT: { using T = float; __VA_ARGS__; } break;                  \
    case TCR_DOUBLE: { using T = double; __VA_ARGS__; } break;                \
    case TCR_INT32: { using T = int32_t; __VA_ARGS__; } break;                \
    case TCR_INT64: { using T = int64_t; __VA_ARGS__; } break;                \
    default:                                                                  \
      ::tcr::set_error("dtype %d has no compute kernels (supported: DOUBLE, FLOAT, INT32, INT64)", (int)(dtype)); \
      return TCR_ERR_DTYPE;                                                   \
  }

#define TCR_DISPATCH_ALL(dtype, T, ...)                                        \
  switch (dtype) {                                                            \
    case TCR_FLOAT: { using T = float; __VA_ARGS__; } break;                  \
    case TCR_DOUBLE: { using T = double; __VA_ARGS__; } break;                \
    case TCR_INT8: { using T = int8_t; __VA_ARGS__; } break;                  \
    case TCR_UINT8: { using T = uint8_t; __VA_ARGS__; } break;                \
    case TCR_INT16: { using T = int16_t; __VA_ARGS__; } break;                \
    case TCR_UINT16: { using T = uint16_t; __VA_ARGS__; } break;              \
    case TCR_INT32: { using T = int32_t; __VA_ARGS__; } break;                \
    case TCR_UINT32: { using T = uint32_t; __VA_ARGS__; } break;              \
    case TCR_INT64: { using T = int64_t; __VA_ARGS__; } break;                \
    case TCR_UINT64: { using T = uint64_t; __VA_ARGS__; } break;              \
    default:                                                                  \
      ::tcr::set_error("bad dtype %d", (int)(dtype));                         \
      return TCR_ERR_DTYPE;                                                   \
  }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid sized in whole waves of the SM count for grid-stride kernels
inline int wave_grid(int64_t work_items, int per_block, int blocks_per_sm) {
  int64_t need = ceil_div(work_items, per_block);
  int64_t cap = (int64_t)state().sm_count * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace tcr
