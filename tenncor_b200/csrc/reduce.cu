// reduce.cu — REDUCE_SUM/PROD/MIN/MAX and ARGMAX (HBM-bound).
//
// Replaces internal/eigen/operator.hpp:54-155. The rank-8 shape and the reduced-rank
// mask are collapsed on the host into [Kin, R, Kout] (kept-inner, reduced, kept-outer):
//   * Kin == 1  -> "row" kernel: contiguous runs of R elements, warp-shuffle + smem tree,
//                  16-byte loads, split across blocks (two-pass) when rows are few;
//   * Kin  > 1  -> "column" kernel: lanes walk the kept-inner rank (coalesced), the block's
//                  y-threads and grid.y split R, smem tree over y, second pass over splits;
//   * anything that does not collapse to three segments takes a generic gather kernel.
#include <float.h>

#include <vector>

#include "common.cuh"

namespace tcr {

template <typename T> struct Lim;
template <> struct Lim<float> { static __device__ float lo() { return -INFINITY; } static __device__ float hi() { return INFINITY; } };
template <> struct Lim<double> { static __device__ double lo() { return -INFINITY; } static __device__ double hi() { return INFINITY; } };
template <> struct Lim<int32_t> { static __device__ int32_t lo() { return INT32_MIN; } static __device__ int32_t hi() { return INT32_MAX; } };
template <> struct Lim<int64_t> { static __device__ int64_t lo() { return INT64_MIN; } static __device__ int64_t hi() { return INT64_MAX; } };

enum { R_SUM = 0, R_PROD = 1, R_MIN = 2, R_MAX = 3 };

template <typename T, int OP> struct Red {
  static __device__ __forceinline__ T init() {
    if (OP == R_SUM) return T(0);
    if (OP == R_PROD) return T(1);
    if (OP == R_MIN) return Lim<T>::hi();
    return Lim<T>::lo();
  }
  static __device__ __forceinline__ T op(T a, T b) {
    if (OP == R_SUM) return a + b;
    if (OP == R_PROD) return a * b;
    if (OP == R_MIN) return b < a ? b : a;
    return a < b ? b : a;
  }
};

template <typename T> __device__ __forceinline__ T shfl_down(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <> __device__ __forceinline__ int64_t shfl_down(int64_t v, int d) { return (int64_t)__shfl_down_sync(0xffffffffu, (long long)v, d); }

template <typename T, int OP>
__device__ __forceinline__ T warp_reduce(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = Red<T, OP>::op(v, shfl_down(v, d));
  return v;
}

// block-wide reduction; result valid in thread 0
template <typename T, int OP>
__device__ __forceinline__ T block_reduce(T v) {
  __shared__ T smem[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warp_reduce<T, OP>(v);
  __syncthreads();  // smem reuse across calls
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nwarps ? smem[lane] : Red<T, OP>::init();
    v = warp_reduce<T, OP>(v);
  }
  return v;
}

// ---- rows: out[ko*S + s] = reduce in[ko*R + chunk_s]; grid = (S, Kout)
template <typename T, int OP>
__global__ void __launch_bounds__(256) reduce_rows_block(const T* __restrict__ in, T* __restrict__ out, int64_t R,
                                                         int64_t chunk, int S, int64_t Kout) {
  TCR_PDL_ENTER();
  const int s = blockIdx.x;
  // grid.y walks the kept rows (any count: 65536 rows of 4096 is a SURVEY §8d case)
  for (int64_t ko = blockIdx.y; ko < Kout; ko += gridDim.y) {
  const T* row = in + ko * R;
  int64_t lo = (int64_t)s * chunk, hi = lo + chunk < R ? lo + chunk : R;
  T acc = Red<T, OP>::init();
  constexpr int N = 16 / sizeof(T);
  // align to 16 bytes
  int64_t i = lo + threadIdx.x;
  const uintptr_t addr = (uintptr_t)(row + lo);
  int64_t head = ((16 - (addr & 15)) & 15) / sizeof(T);
  if (head > hi - lo) head = hi - lo;
  if ((int64_t)threadIdx.x < head) acc = Red<T, OP>::op(acc, row[i]);
  const int64_t vlo = lo + head;
  const int64_t nvec = (hi - vlo) / N;
  const uint4* vp = reinterpret_cast<const uint4*>(row + vlo);
  T acc2 = Red<T, OP>::init();
  int64_t v = threadIdx.x;
  for (; v + blockDim.x < nvec; v += 2 * blockDim.x) {
    uint4 q0 = __ldg(vp + v), q1 = __ldg(vp + v + blockDim.x);
    const T* e0 = reinterpret_cast<const T*>(&q0);
    const T* e1 = reinterpret_cast<const T*>(&q1);
#pragma unroll
    for (int k = 0; k < N; ++k) { acc = Red<T, OP>::op(acc, e0[k]); acc2 = Red<T, OP>::op(acc2, e1[k]); }
  }
  for (; v < nvec; v += blockDim.x) {
    uint4 q0 = __ldg(vp + v);
    const T* e0 = reinterpret_cast<const T*>(&q0);
#pragma unroll
    for (int k = 0; k < N; ++k) acc = Red<T, OP>::op(acc, e0[k]);
  }
  acc = Red<T, OP>::op(acc, acc2);
  for (int64_t t = vlo + nvec * N + threadIdx.x; t < hi; t += blockDim.x) acc = Red<T, OP>::op(acc, row[t]);
  acc = block_reduce<T, OP>(acc);
  if (threadIdx.x == 0) out[ko * S + s] = acc;
  __syncthreads();  // block_reduce's scratch is reused by the next row
  }
}

// ---- rows, one warp per row (short rows, many rows)
template <typename T, int OP>
__global__ void __launch_bounds__(256) reduce_rows_warp(const T* __restrict__ in, T* __restrict__ out, int64_t R,
                                                        int64_t Kout) {
  TCR_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t ko = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ko < Kout; ko += warps_per_grid) {
    const T* row = in + ko * R;
    T acc = Red<T, OP>::init();
    for (int64_t r = lane; r < R; r += 32) acc = Red<T, OP>::op(acc, __ldg(row + r));
    acc = warp_reduce<T, OP>(acc);
    if (lane == 0) out[ko] = acc;
  }
}

// ---- columns: out[ki + Kin*(s + S*ko)] = reduce_{r in chunk s} in[ki + Kin*(r + R*ko)]
// block = (32, 8); grid = (ceil(Kin/32), S, Kout)
template <typename T, int OP>
__global__ void __launch_bounds__(256) reduce_cols(const T* __restrict__ in, T* __restrict__ out, int64_t Kin,
                                                   int64_t R, int64_t chunk, int S) {
  TCR_PDL_ENTER();
  __shared__ T tile[8][33];
  const int64_t ki = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int s = blockIdx.y;
  const int64_t ko = blockIdx.z;
  int64_t lo = (int64_t)s * chunk, hi = lo + chunk < R ? lo + chunk : R;
  T acc = Red<T, OP>::init();
  if (ki < Kin) {
    const T* base = in + ki + Kin * (R * ko);
    int64_t r = lo + threadIdx.y;
    T a1 = Red<T, OP>::init(), a2 = Red<T, OP>::init(), a3 = Red<T, OP>::init();
    for (; r + 24 < hi; r += 32) {
      T x0 = __ldg(base + Kin * r), x1 = __ldg(base + Kin * (r + 8)), x2 = __ldg(base + Kin * (r + 16)),
        x3 = __ldg(base + Kin * (r + 24));
      acc = Red<T, OP>::op(acc, x0); a1 = Red<T, OP>::op(a1, x1); a2 = Red<T, OP>::op(a2, x2); a3 = Red<T, OP>::op(a3, x3);
    }
    for (; r < hi; r += 8) acc = Red<T, OP>::op(acc, __ldg(base + Kin * r));
    acc = Red<T, OP>::op(Red<T, OP>::op(acc, a1), Red<T, OP>::op(a2, a3));
  }
  tile[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && ki < Kin) {
#pragma unroll
    for (int y = 1; y < 8; ++y) acc = Red<T, OP>::op(acc, tile[y][threadIdx.x]);
    out[ki + Kin * ((int64_t)s + (int64_t)S * ko)] = acc;
  }
}

// 16-byte variant of reduce_cols for 4-byte elements: a thread owns FOUR adjacent kept columns, a warp reads 512 contiguous
// bytes of a row (128-byte reads at a 16 KB row stride ran at 50 % of the copy bandwidth on [4096, 65536]).
// block = (32, 8); grid = (ceil(Kin/128), S, Kout)
template <typename T, int OP>
__global__ void __launch_bounds__(256) reduce_cols_v4(const T* __restrict__ in, T* __restrict__ out, int64_t Kin, int64_t R, int64_t chunk, int S) {
  static_assert(sizeof(T) == 4, "16-byte loads of four elements");
  TCR_PDL_ENTER();
  struct alignas(16) Q { T v[4]; };
  __shared__ Q tile[8][33];
  const int64_t ki = ((int64_t)blockIdx.x * 32 + threadIdx.x) * 4;
  const int s = blockIdx.y;
  const int64_t ko = blockIdx.z;
  const int64_t lo = (int64_t)s * chunk, hi = lo + chunk < R ? lo + chunk : R;
  Q acc, a1;
#pragma unroll
  for (int v = 0; v < 4; ++v) acc.v[v] = a1.v[v] = Red<T, OP>::init();
  if (ki < Kin) {
    const T* base = in + ki + Kin * (R * ko);
    int64_t r = lo + threadIdx.y;
    for (; r + 24 < hi; r += 32) {
      const Q x0 = *reinterpret_cast<const Q*>(base + Kin * r), x1 = *reinterpret_cast<const Q*>(base + Kin * (r + 8)),
              x2 = *reinterpret_cast<const Q*>(base + Kin * (r + 16)), x3 = *reinterpret_cast<const Q*>(base + Kin * (r + 24));
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        acc.v[v] = Red<T, OP>::op(acc.v[v], Red<T, OP>::op(x0.v[v], x1.v[v]));
        a1.v[v] = Red<T, OP>::op(a1.v[v], Red<T, OP>::op(x2.v[v], x3.v[v]));
      }
    }
    for (; r < hi; r += 8) {
      const Q x = *reinterpret_cast<const Q*>(base + Kin * r);
#pragma unroll
      for (int v = 0; v < 4; ++v) acc.v[v] = Red<T, OP>::op(acc.v[v], x.v[v]);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) acc.v[v] = Red<T, OP>::op(acc.v[v], a1.v[v]);
  }
  tile[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && ki < Kin) {
#pragma unroll
    for (int y = 1; y < 8; ++y)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc.v[v] = Red<T, OP>::op(acc.v[v], tile[y][threadIdx.x].v[v]);
    *reinterpret_cast<Q*>(out + ki + Kin * ((int64_t)s + (int64_t)S * ko)) = acc;
  }
}

template <typename T, int OP, bool FOUR = sizeof(T) == 4>
struct ColsV4 {
  static void launch(dim3 grid, const T* in, T* out, int64_t Kin, int64_t R, int64_t chunk, int S) { TCR_LAUNCH((reduce_cols_v4<T, OP>), grid, dim3(32, 8), 0, in, out, Kin, R, chunk, S); }
};
template <typename T, int OP>
struct ColsV4<T, OP, false> {
  static void launch(dim3, const T*, T*, int64_t, int64_t, int64_t, int) {}
};

// ---- generic: arbitrary mask, one thread per output element
struct GenericDesc {
  int64_t shape[8];
  int64_t in_stride[8];
  int nk, nr;         // number of kept / reduced ranks
  int kdims[8], rdims[8];
  int64_t n_out, n_red;
};

template <typename T, int OP>
__global__ void __launch_bounds__(256) reduce_generic(const T* __restrict__ in, T* __restrict__ out, GenericDesc d) {
  TCR_PDL_ENTER();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < d.n_out; o += stride) {
    int64_t base = 0, t = o;
    for (int k = 0; k < d.nk; ++k) {
      int64_t c = t % d.shape[d.kdims[k]];
      t /= d.shape[d.kdims[k]];
      base += c * d.in_stride[d.kdims[k]];
    }
    T acc = Red<T, OP>::init();
    for (int64_t q = 0; q < d.n_red; ++q) {
      int64_t off = base, u = q;
      for (int k = 0; k < d.nr; ++k) {
        int64_t c = u % d.shape[d.rdims[k]];
        u /= d.shape[d.rdims[k]];
        off += c * d.in_stride[d.rdims[k]];
      }
      acc = Red<T, OP>::op(acc, in[off]);
    }
    out[o] = acc;
  }
}

// ------------------------------------------------------------------ host drivers
template <typename T, int OP>
static int run_rows(const T* in, T* out, int64_t R, int64_t Kout) {
  State& st = state();
  if (R <= 1024 && Kout >= 64) {
    int grid = wave_grid(Kout, 8, 8);
    TCR_LAUNCH((reduce_rows_warp<T, OP>), grid, 256, 0, in, out, R, Kout);
    TCR_CHECK_LAUNCH();
    return TCR_OK;
  }
  // split rows so that the grid fills the machine (~4 blocks per SM)
  int64_t want = (int64_t)st.sm_count * 4;
  int64_t S = 1;
  if (Kout < want) {
    S = ceil_div(want, Kout);
    int64_t maxS = ceil_div(R, 4096);  // at least 16 elements per thread
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
  }
  const unsigned gy = (unsigned)(Kout < 65535 ? Kout : 65535);
  int64_t chunk = ceil_div(R, S);
  chunk = (chunk + 3) / 4 * 4;
  S = ceil_div(R, chunk);
  if (S == 1) {
    TCR_LAUNCH((reduce_rows_block<T, OP>), dim3(1, gy), 256, 0, in, out, R, chunk, 1, Kout);
    TCR_CHECK_LAUNCH();
    return TCR_OK;
  }
  void* part = nullptr;
  int rc = tcr_alloc(&part, sizeof(T) * (size_t)(Kout * S));
  if (rc) return rc;
  TCR_LAUNCH((reduce_rows_block<T, OP>), dim3((unsigned)S, gy), 256, 0, in, (T*)part, R, chunk, (int)S, Kout);
  TCR_CHECK_LAUNCH();
  rc = run_rows<T, OP>((const T*)part, out, S, Kout);
  tcr_free(part);
  return rc;
}

template <typename T, int OP>
static int run_cols(const T* in, T* out, int64_t Kin, int64_t R, int64_t Kout) {
  State& st = state();
  TCR_ARG(Kout <= 65535, "tcr_reduce: kept-outer extent %lld exceeds grid.z", (long long)Kout);
  // four columns per thread when the rows are 16-byte aligned runs of 4-byte elements
  const bool v4 = sizeof(T) == 4 && (Kin % 4) == 0 && Kin >= 128 && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
  int64_t bx = ceil_div(Kin, v4 ? 128 : 32);
  int64_t base_blocks = bx * Kout;
  int64_t want = (int64_t)st.sm_count * 8;
  int64_t S = 1;
  if (base_blocks < want) {
    S = ceil_div(want, base_blocks);
    int64_t maxS = ceil_div(R, 64);
    if (S > maxS) S = maxS;
    if (S < 1) S = 1;
    if (S > 65535) S = 65535;
  }
  int64_t chunk = ceil_div(R, S);
  S = ceil_div(R, chunk);
  if (S == 1) {
    if (v4) ColsV4<T, OP>::launch(dim3((unsigned)bx, 1, (unsigned)Kout), in, out, Kin, R, chunk, 1);
    else TCR_LAUNCH((reduce_cols<T, OP>), dim3((unsigned)bx, 1, (unsigned)Kout), dim3(32, 8), 0, in, out, Kin, R, chunk, 1);
    TCR_CHECK_LAUNCH();
    return TCR_OK;
  }
  void* part = nullptr;
  int rc = tcr_alloc(&part, sizeof(T) * (size_t)(Kin * S * Kout));
  if (rc) return rc;
  if (v4) ColsV4<T, OP>::launch(dim3((unsigned)bx, (unsigned)S, (unsigned)Kout), in, (T*)part, Kin, R, chunk, (int)S);
  else TCR_LAUNCH((reduce_cols<T, OP>), dim3((unsigned)bx, (unsigned)S, (unsigned)Kout), dim3(32, 8), 0, in, (T*)part, Kin, R, chunk, (int)S);
  TCR_CHECK_LAUNCH();
  rc = run_cols<T, OP>((const T*)part, out, Kin, S, Kout);
  tcr_free(part);
  return rc;
}

template <typename T, int OP>
static int run_reduce(const void* in, void* out, const int64_t shape[8], uint32_t mask) {
  // collapse: drop extent-1 ranks, merge neighbours with the same reduced flag
  std::vector<std::pair<int64_t, bool>> seg;
  int64_t n = 1;
  for (int r = 0; r < 8; ++r) {
    n *= shape[r];
    if (shape[r] == 1) continue;
    bool red = (mask >> r) & 1u;
    if (!seg.empty() && seg.back().second == red) seg.back().first *= shape[r];
    else seg.push_back({shape[r], red});
  }
  if (n == 0) return TCR_OK;
  bool any_red = false;
  for (auto& s : seg) any_red |= s.second;
  if (!any_red) return tcr_d2d(out, in, sizeof(T) * (size_t)n);
  int64_t Kin = 1, R = 1, Kout = 1;
  bool simple = true;
  if (seg.size() == 1) { R = seg[0].first; }
  else if (seg.size() == 2 && seg[0].second) { R = seg[0].first; Kout = seg[1].first; }
  else if (seg.size() == 2) { Kin = seg[0].first; R = seg[1].first; }
  else if (seg.size() == 3 && !seg[0].second) { Kin = seg[0].first; R = seg[1].first; Kout = seg[2].first; }
  else simple = false;
  if (simple && Kin == 1) return run_rows<T, OP>((const T*)in, (T*)out, R, Kout);
  if (simple && Kout <= 65535) return run_cols<T, OP>((const T*)in, (T*)out, Kin, R, Kout);
  GenericDesc d;
  memset(&d, 0, sizeof(d));
  int64_t stride = 1;
  d.n_out = 1; d.n_red = 1;
  for (int r = 0; r < 8; ++r) {
    d.shape[r] = shape[r];
    d.in_stride[r] = stride;
    stride *= shape[r];
    if (shape[r] == 1) continue;
    if ((mask >> r) & 1u) { d.rdims[d.nr++] = r; d.n_red *= shape[r]; }
    else { d.kdims[d.nk++] = r; d.n_out *= shape[r]; }
  }
  int grid = wave_grid(d.n_out, 256, 8);
  TCR_LAUNCH((reduce_generic<T, OP>), grid, 256, 0, (const T*)in, (T*)out, d);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

// ------------------------------------------------------------------ argmax
template <typename T> struct ValIdx { T v; int64_t i; };

template <typename T>
__device__ __forceinline__ ValIdx<T> vi_better(ValIdx<T> a, ValIdx<T> b) {
  // strict > keeps the first (lowest-index) maximum; NaN never wins
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

template <typename T>
__device__ __forceinline__ ValIdx<T> vi_warp(ValIdx<T> x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    ValIdx<T> y;
    y.v = shfl_down(x.v, d);
    y.i = (int64_t)__shfl_down_sync(0xffffffffu, (long long)x.i, d);
    x = vi_better(x, y);
  }
  return x;
}

// rows of length R with element stride `es` and row offsets given by (ki, ko):
// element(r) = in[ki + Kin*(r + R*ko)]. Kin == 1: one warp per row; else one thread per (ki, ko).
template <typename T>
__global__ void __launch_bounds__(256) argmax_warp(const T* __restrict__ in, T* __restrict__ out, int64_t R, int64_t Kout) {
  TCR_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int64_t wpg = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t ko = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ko < Kout; ko += wpg) {
    const T* row = in + ko * R;
    ValIdx<T> best{Lim<T>::lo(), INT64_MAX};
    for (int64_t r = lane; r < R; r += 32) best = vi_better(best, ValIdx<T>{row[r], r});
    best = vi_warp(best);
    if (lane == 0) out[ko] = (T)(best.i == INT64_MAX ? 0 : best.i);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) argmax_cols(const T* __restrict__ in, T* __restrict__ out, int64_t Kin, int64_t R,
                                                   int64_t Kout) {
  TCR_PDL_ENTER();
  const int64_t total = Kin * Kout, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
    int64_t ki = o % Kin, ko = o / Kin;
    const T* base = in + ki + Kin * R * ko;
    T bv = base[0];
    int64_t bi = 0;
    for (int64_t r = 1; r < R; ++r) {
      T x = base[Kin * r];
      if (x > bv) { bv = x; bi = r; }
    }
    out[o] = (T)bi;
  }
}

// columns with few outputs and a long reduced rank: block = (32 lanes along ki, 8 along r),
// grid = (ceil(Kin/32), S, Kout). Pass 1 (pidx == nullptr) scans `in` and emits per-split winners
// (value, original r); pass 2 merges the S winners. Ties resolve to the lowest r in both passes.
template <typename T>
__global__ void __launch_bounds__(256) argmax_cols_split(const T* __restrict__ in, const int64_t* __restrict__ pidx, T* __restrict__ out_val,
                                                         int64_t* __restrict__ out_idx, T* __restrict__ out_final, int64_t Kin, int64_t R,
                                                         int64_t chunk, int S) {
  TCR_PDL_ENTER();
  __shared__ T sv[8][33];
  __shared__ int64_t si[8][33];
  const int64_t ki = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int s = blockIdx.y;
  const int64_t ko = blockIdx.z;
  const int64_t lo = (int64_t)s * chunk, hi = lo + chunk < R ? lo + chunk : R;
  ValIdx<T> best{Lim<T>::lo(), INT64_MAX};
  if (ki < Kin) {
    const int64_t base = ki + Kin * (R * ko);
    for (int64_t r = lo + threadIdx.y; r < hi; r += 8) {
      const int64_t at = base + Kin * r;
      best = vi_better(best, ValIdx<T>{in[at], pidx ? pidx[at] : r});
    }
  }
  sv[threadIdx.y][threadIdx.x] = best.v;
  si[threadIdx.y][threadIdx.x] = best.i;
  __syncthreads();
  if (threadIdx.y == 0 && ki < Kin) {
#pragma unroll
    for (int y = 1; y < 8; ++y) best = vi_better(best, ValIdx<T>{sv[y][threadIdx.x], si[y][threadIdx.x]});
    const int64_t o = ki + Kin * ((int64_t)s + (int64_t)S * ko);
    if (out_final) out_final[ki + Kin * ko] = (T)(best.i == INT64_MAX ? 0 : best.i);
    else { out_val[o] = best.v; out_idx[o] = best.i; }
  }
}

// flat: pass 1 -> per-block (val, idx); pass 2 (single block) -> out
template <typename T>
__global__ void __launch_bounds__(256) argmax_flat1(const T* __restrict__ in, int64_t n, T* __restrict__ pv, int64_t* __restrict__ pi) {
  TCR_PDL_ENTER();
  __shared__ T sv[8];
  __shared__ int64_t si[8];
  ValIdx<T> best{Lim<T>::lo(), INT64_MAX};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t done = 0;
  if (sizeof(T) == 4 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    // 16-byte loads, four vectors in flight per thread. A thread meets its elements in increasing index order, so a strict `>`
    // keeps the lowest index of equal maxima without comparing indices; the `== ... && none yet` arm lets a tensor made of the
    // lowest representable value (or of -inf) still name its first element (NaN never wins, as in vi_better).
    struct alignas(16) Q { T v[4]; };
    const Q* in4 = reinterpret_cast<const Q*>(in);
    const int64_t nvec = n / 4;
    int64_t i = tid;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
      const Q q0 = in4[i], q1 = in4[i + stride], q2 = in4[i + 2 * stride], q3 = in4[i + 3 * stride];
#pragma unroll
      for (int v = 0; v < 4; ++v) if (q0.v[v] > best.v || (q0.v[v] == best.v && best.i == INT64_MAX)) { best.v = q0.v[v]; best.i = 4 * i + v; }
#pragma unroll
      for (int v = 0; v < 4; ++v) if (q1.v[v] > best.v || (q1.v[v] == best.v && best.i == INT64_MAX)) { best.v = q1.v[v]; best.i = 4 * (i + stride) + v; }
#pragma unroll
      for (int v = 0; v < 4; ++v) if (q2.v[v] > best.v || (q2.v[v] == best.v && best.i == INT64_MAX)) { best.v = q2.v[v]; best.i = 4 * (i + 2 * stride) + v; }
#pragma unroll
      for (int v = 0; v < 4; ++v) if (q3.v[v] > best.v || (q3.v[v] == best.v && best.i == INT64_MAX)) { best.v = q3.v[v]; best.i = 4 * (i + 3 * stride) + v; }
    }
    for (; i < nvec; i += stride) {
      const Q q = in4[i];
#pragma unroll
      for (int v = 0; v < 4; ++v) if (q.v[v] > best.v || (q.v[v] == best.v && best.i == INT64_MAX)) { best.v = q.v[v]; best.i = 4 * i + v; }
    }
    done = nvec * 4;
  }
  for (int64_t i = done + tid; i < n; i += stride) best = vi_better(best, ValIdx<T>{in[i], i});
  best = vi_warp(best);
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best.v; si[threadIdx.x >> 5] = best.i; }
  __syncthreads();
  if (threadIdx.x < 32) {
    ValIdx<T> x{Lim<T>::lo(), INT64_MAX};
    if (threadIdx.x < (blockDim.x >> 5)) { x.v = sv[threadIdx.x]; x.i = si[threadIdx.x]; }
    x = vi_warp(x);
    if (threadIdx.x == 0) { pv[blockIdx.x] = x.v; pi[blockIdx.x] = x.i; }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) argmax_flat2(const T* __restrict__ pv, const int64_t* __restrict__ pi, int nparts, T* __restrict__ out) {
  TCR_PDL_ENTER();
  __shared__ T sv[8];
  __shared__ int64_t si[8];
  ValIdx<T> best{Lim<T>::lo(), INT64_MAX};
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) best = vi_better(best, ValIdx<T>{pv[i], pi[i]});
  best = vi_warp(best);
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best.v; si[threadIdx.x >> 5] = best.i; }
  __syncthreads();
  if (threadIdx.x < 32) {
    ValIdx<T> x{Lim<T>::lo(), INT64_MAX};
    if (threadIdx.x < (blockDim.x >> 5)) { x.v = sv[threadIdx.x]; x.i = si[threadIdx.x]; }
    x = vi_warp(x);
    if (threadIdx.x == 0) out[0] = (T)(x.i == INT64_MAX ? 0 : x.i);
  }
}

template <typename T>
static int run_argmax(const void* in, void* out, const int64_t shape[8], int return_dim) {
  int64_t n = 1;
  for (int r = 0; r < 8; ++r) n *= shape[r];
  if (n == 0) return TCR_OK;
  if (return_dim >= 8) {
    int grid = wave_grid(n, 256 * 8, 4);
    void* pv = nullptr; void* pi = nullptr;
    int rc = tcr_alloc(&pv, sizeof(T) * grid);
    if (rc) return rc;
    rc = tcr_alloc(&pi, sizeof(int64_t) * grid);
    if (rc) { tcr_free(pv); return rc; }
    TCR_LAUNCH((argmax_flat1<T>), grid, 256, 0, (const T*)in, n, (T*)pv, (int64_t*)pi);
    TCR_LAUNCH((argmax_flat2<T>), 1, 256, 0, (const T*)pv, (const int64_t*)pi, grid, (T*)out);
    TCR_CHECK_LAUNCH();
    tcr_free(pv); tcr_free(pi);
    return TCR_OK;
  }
  int64_t Kin = 1, Kout = 1, R = shape[return_dim];
  for (int r = 0; r < return_dim; ++r) Kin *= shape[r];
  for (int r = return_dim + 1; r < 8; ++r) Kout *= shape[r];
  if (Kin == 1) {
    int grid = wave_grid(Kout, 8, 8);
    TCR_LAUNCH((argmax_warp<T>), grid, 256, 0, (const T*)in, (T*)out, R, Kout);
  } else if (Kin * Kout >= (int64_t)state().sm_count * 512 || R < 64 || Kout > 65535) {
    int grid = wave_grid(Kin * Kout, 256, 8);
    TCR_LAUNCH((argmax_cols<T>), grid, 256, 0, (const T*)in, (T*)out, Kin, R, Kout);
  } else {
    const int64_t bx = ceil_div(Kin, 32);
    int64_t S = ceil_div((int64_t)state().sm_count * 8, bx * Kout);
    if (S > ceil_div(R, 64)) S = ceil_div(R, 64);
    if (S < 1) S = 1;
    if (S > 65535) S = 65535;
    int64_t chunk = ceil_div(R, S);
    S = ceil_div(R, chunk);
    if (S == 1) {
      TCR_LAUNCH((argmax_cols_split<T>), dim3((unsigned)bx, 1, (unsigned)Kout), dim3(32, 8), 0, (const T*)in, (const int64_t*)nullptr,
                 (T*)nullptr, (int64_t*)nullptr, (T*)out, Kin, R, chunk, 1);
    } else {
      void* pv = nullptr; void* pi = nullptr;
      int rc = tcr_alloc(&pv, sizeof(T) * (size_t)(Kin * S * Kout));
      if (rc) return rc;
      rc = tcr_alloc(&pi, sizeof(int64_t) * (size_t)(Kin * S * Kout));
      if (rc) { tcr_free(pv); return rc; }
      TCR_LAUNCH((argmax_cols_split<T>), dim3((unsigned)bx, (unsigned)S, (unsigned)Kout), dim3(32, 8), 0, (const T*)in,
                 (const int64_t*)nullptr, (T*)pv, (int64_t*)pi, (T*)nullptr, Kin, R, chunk, (int)S);
      TCR_LAUNCH((argmax_cols_split<T>), dim3((unsigned)bx, 1, (unsigned)Kout), dim3(32, 8), 0, (const T*)pv, (const int64_t*)pi,
                 (T*)nullptr, (int64_t*)nullptr, (T*)out, Kin, S, S, 1);
      tcr_free(pv); tcr_free(pi);
    }
  }
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

}  // namespace tcr

using namespace tcr;

extern "C" {

int tcr_reduce(int opcode, const void* in, void* out, const int64_t shape[TCR_RANK_CAP], uint32_t reduce_mask, int dtype) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(in && out && shape, "tcr_reduce: null argument");
  for (int r = 0; r < 8; ++r) TCR_ARG(shape[r] >= 0, "tcr_reduce: negative extent at rank %d", r);
  TCR_DISPATCH_COMPUTE(dtype, T, {
    switch (opcode) {
      case TCR_OP_REDUCE_SUM: return run_reduce<T, R_SUM>(in, out, shape, reduce_mask);
      case TCR_OP_REDUCE_PROD: return run_reduce<T, R_PROD>(in, out, shape, reduce_mask);
      case TCR_OP_REDUCE_MIN: return run_reduce<T, R_MIN>(in, out, shape, reduce_mask);
      case TCR_OP_REDUCE_MAX: return run_reduce<T, R_MAX>(in, out, shape, reduce_mask);
      default: set_error("tcr_reduce: opcode %d is not a reduction", opcode); return TCR_ERR_ARG;
    }
  });
  return TCR_OK;
}

int tcr_argmax(const void* in, void* out, const int64_t shape[TCR_RANK_CAP], int return_dim, int dtype) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(in && out && shape, "tcr_argmax: null argument");
  TCR_ARG(return_dim >= 0, "tcr_argmax: negative return_dim");
  TCR_DISPATCH_COMPUTE(dtype, T, return run_argmax<T>(in, out, shape, return_dim));
  return TCR_OK;
}

}  // extern "C"
