// gemm_skinny.cu — memory-bound GEMM shapes: one extent of the product is tiny.
//
// The output layers of the reference's demo models are skinny (MNIST-shaped MLP: 1024 -> 10,
// gd_demo: 9 -> 5): their forward (N = 10), weight-gradient (M = 10, K = batch) and
// input-gradient (K = 10) contractions (tenncor/eteq/backprop.hpp:269-359) stream one large
// operand once and are bound by HBM, not by the tensor pipe. Tiling them 64x64 or 128x128
// leaves most of the machine idle (16 CTAs for dW of a 1024x10 layer), so they get their own
// coalesced streaming kernels:
//   skinny_rk  reduce along the contiguous rank of the big operand  (one warp per row)
//   skinny_rs  reduce along the strided rank of the big operand     (thread per column, split-K,
//              deterministic two-pass reduction over the splits)
//   small_k    K <= 16: every output is a short dot product         (4 outputs per thread)
#include <cstdlib>

#include "common.cuh"

namespace tcr {

template <typename T> __device__ __forceinline__ T sk_act(int act, T x) { return x; }
template <> __device__ __forceinline__ float sk_act(int act, float x) {
  if (act == TCR_EW_SIGMOID) return 1.0f / (1.0f + expf(-x));
  if (act == TCR_EW_TANH) return tanhf(x);
  return x;
}
template <> __device__ __forceinline__ double sk_act(int act, double x) {
  if (act == TCR_EW_SIGMOID) return 1.0 / (1.0 + exp(-x));
  if (act == TCR_EW_TANH) return tanh(x);
  return x;
}

// normalised problem: C(l, s) = sum_k X(l, k) * Y(s, k), l < L (large), s < S (<= 16)
struct SkinnyDesc {
  int64_t L, S, K;
  int64_t x_sl, x_sk, y_ss, y_sk, c_sl, c_ss;
  int bias_on_l, has_bias, activation, accumulate;
  const void* bias;
};

constexpr int SK_MAX = 16;

template <typename T>
__device__ __forceinline__ void sk_store(const SkinnyDesc& d, T* C, int64_t l, int s, T v) {
  T* dst = C + l * d.c_sl + s * d.c_ss;
  if (d.accumulate) v += *dst;
  if (d.has_bias) v += ((const T*)d.bias)[d.bias_on_l ? l : s];
  if (d.activation) v = sk_act<T>(d.activation, v);
  *dst = v;
}

template <typename T> __device__ __forceinline__ T sk_shfl_down(T v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ int64_t sk_shfl_down(int64_t v, int o) { return (int64_t)__shfl_down_sync(0xffffffffu, (long long)v, o); }

// X is K-major (x_sk == 1): a block owns 32 rows (4 per warp); the small operand is staged in shared
// memory 128 k at a time ([s][k]: lanes read consecutive k, conflict-free), lanes stride over k
template <typename T>
__global__ void __launch_bounds__(256) skinny_rk_kernel(const T* __restrict__ X, const T* __restrict__ Y, T* __restrict__ C,
                                                        const __grid_constant__ SkinnyDesc d) {
  TCR_PDL_ENTER();
  constexpr int KC = 128, RPW = 4;
  __shared__ T ysm[SK_MAX][KC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t row_base = (int64_t)blockIdx.x * 32; row_base < d.L; row_base += (int64_t)gridDim.x * 32) {
    T acc[RPW][SK_MAX];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int s = 0; s < SK_MAX; ++s) acc[r][s] = T(0);
    for (int64_t kb = 0; kb < d.K; kb += KC) {
      for (int e = threadIdx.x; e < SK_MAX * KC; e += 256) {
        const int s = e / KC, kk = e % KC;
        ysm[s][kk] = (s < d.S && kb + kk < d.K) ? Y[s * d.y_ss + (kb + kk) * d.y_sk] : T(0);
      }
      __syncthreads();
      T x[RPW][KC / 32];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int64_t l = row_base + warp * RPW + r;
#pragma unroll
        for (int i = 0; i < KC / 32; ++i) {
          const int64_t k = kb + lane + 32 * i;
          x[r][i] = (l < d.L && k < d.K) ? X[l * d.x_sl + k] : T(0);
        }
      }
#pragma unroll
      for (int i = 0; i < KC / 32; ++i)
#pragma unroll
        for (int s = 0; s < SK_MAX; ++s) {
          const T y = ysm[s][lane + 32 * i];
#pragma unroll
          for (int r = 0; r < RPW; ++r) acc[r][s] += x[r][i] * y;
        }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int64_t l = row_base + warp * RPW + r;
#pragma unroll
      for (int s = 0; s < SK_MAX; ++s) {
        T v = acc[r][s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += sk_shfl_down(v, o);
        if (lane == 0 && l < d.L && s < d.S) sk_store<T>(d, C, l, s, v);
      }
    }
  }
}

// X is L-major (x_sl == 1): thread per l, block.y chunks of K; partial sums to ws[z][s][l] when split
template <typename T>
__global__ void __launch_bounds__(256) skinny_rs_kernel(const T* __restrict__ X, const T* __restrict__ Y, T* __restrict__ C,
                                                        T* __restrict__ ws, int64_t kchunk, const __grid_constant__ SkinnyDesc d) {
  TCR_PDL_ENTER();
  const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k0 = (int64_t)blockIdx.y * kchunk, k1 = k0 + kchunk < d.K ? k0 + kchunk : d.K;
  __shared__ T ysm[SK_MAX][64];
  T acc[SK_MAX];
#pragma unroll
  for (int s = 0; s < SK_MAX; ++s) acc[s] = T(0);
  for (int64_t kb = k0; kb < k1; kb += 64) {
    // stage the small operand for 64 k
    for (int e = threadIdx.x; e < SK_MAX * 64; e += blockDim.x) {
      int s = e / 64, kk = e % 64;
      ysm[s][kk] = (s < d.S && kb + kk < k1) ? Y[s * d.y_ss + (kb + kk) * d.y_sk] : T(0);
    }
    __syncthreads();
    if (l < d.L) {
      const int64_t kend = kb + 64 < k1 ? kb + 64 : k1;
      for (int64_t k = kb; k < kend; ++k) {
        const T x = X[l + k * d.x_sk];
#pragma unroll
        for (int s = 0; s < SK_MAX; ++s) acc[s] += x * ysm[s][k - kb];
      }
    }
    __syncthreads();
  }
  if (l >= d.L) return;
  if (gridDim.y == 1) {
#pragma unroll
    for (int s = 0; s < SK_MAX; ++s)
      if (s < d.S) sk_store<T>(d, C, l, s, acc[s]);
  } else {
#pragma unroll
    for (int s = 0; s < SK_MAX; ++s)
      if (s < d.S) ws[((int64_t)blockIdx.y * d.S + s) * d.L + l] = acc[s];
  }
}

// second pass of the split-K variants: out(l, s) = epilogue(sum_z ws[z][s][l]), z ascending
// (deterministic). A block owns 32 consecutive l of one s; its 8 warps split the z range so the
// loads of one output are spread over 8 threads, then combine through shared memory in warp order.
template <typename T>
__global__ void __launch_bounds__(256) skinny_rs_reduce_kernel(const T* __restrict__ ws, T* __restrict__ C, int splits,
                                                               const __grid_constant__ SkinnyDesc d) {
  TCR_PDL_ENTER();
  __shared__ T part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t lblocks = (d.L + 31) / 32;
  for (int64_t blk = blockIdx.x; blk < lblocks * d.S; blk += gridDim.x) {
    const int s = (int)(blk / lblocks);
    const int64_t l = (blk % lblocks) * 32 + lane;
    const int per = (splits + 7) / 8;
    const int z0 = warp * per, z1 = z0 + per < splits ? z0 + per : splits;
    T v = T(0);
    if (l < d.L) {
#pragma unroll 8
      for (int z = z0; z < z1; ++z) v += ws[((int64_t)z * d.S + s) * d.L + l];
    }
    part[warp][lane] = v;
    __syncthreads();
    if (warp == 0 && l < d.L) {
      T t = part[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += part[w][lane];
      sk_store<T>(d, C, l, s, t);
    }
    __syncthreads();
  }
}

// K <= 16: every output is a short dot product. A thread keeps its column B(:, n) in registers,
// a block stages 64 rows of A in shared memory and streams them; consecutive threads walk n
// (coalesced stores when c_sn == 1)
template <typename T>
__global__ void __launch_bounds__(256) small_k_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C,
                                                      const __grid_constant__ tcr_gemm_desc d) {
  TCR_PDL_ENTER();
  __shared__ T a_sm[64][SK_MAX];
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool n_ok = n < d.n;
  const T* bias = (const T*)d.bias;
  T b[SK_MAX];
#pragma unroll
  for (int k = 0; k < SK_MAX; ++k) b[k] = (n_ok && k < d.k) ? B[k * d.b_sk + n * d.b_sn] : T(0);
  for (int64_t m0 = (int64_t)blockIdx.y * 64; m0 < d.m; m0 += (int64_t)gridDim.y * 64) {
    for (int e = threadIdx.x; e < 64 * SK_MAX; e += 256) {
      const int mm = e / SK_MAX, k = e % SK_MAX;
      a_sm[mm][k] = (m0 + mm < d.m && k < d.k) ? A[(m0 + mm) * d.a_sm + k * d.a_sk] : T(0);
    }
    __syncthreads();
    if (n_ok) {
      const int rows = d.m - m0 < 64 ? (int)(d.m - m0) : 64;
      for (int mm = 0; mm < rows; ++mm) {
        T v = T(0);
#pragma unroll
        for (int k = 0; k < SK_MAX; ++k) v += a_sm[mm][k] * b[k];
        const int64_t m = m0 + mm;
        T* dst = C + m * d.c_sm + n * d.c_sn;
        if (d.accumulate) v += *dst;
        if (d.epilogue == TCR_EPI_BIAS_N) v += bias[n];
        else if (d.epilogue == TCR_EPI_BIAS_M) v += bias[m];
        if (d.activation) v = sk_act<T>(d.activation, v);
        *dst = v;
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------- vectorised streaming variants
// The kernels above are the any-stride fallbacks. When the big operand's contiguous rank is
// 16-byte aligned the variants below move it with 128-bit loads, keep several of them in flight
// per thread (HBM latency x bandwidth needs ~64 KB in flight per SM) and trim the padded small
// extent to a multiple of 4 (template SP) so a 10-wide layer does 12, not 16, FMAs per element.
template <typename T> struct alignas(16) SkVec { T v[16 / sizeof(T)]; };

// X K-major, rows 16-byte aligned, K % V == 0. A warp owns 4 rows; the small operand is staged in
// shared memory as [SP][kc] (whole K when it fits: no barrier inside the row loop).
template <typename T, int SP>
__global__ void __launch_bounds__(256, (SP * sizeof(T) <= 48 ? 2 : 1)) skinny_rk2_kernel(const T* __restrict__ X, const T* __restrict__ Y, T* __restrict__ C,
                                                         int kc, const __grid_constant__ SkinnyDesc d) {
  TCR_PDL_ENTER();
  constexpr int V = 16 / sizeof(T), RPW = 4, KSTEP = 32 * V;
  extern __shared__ __align__(16) unsigned char sk_smem[];
  T* ysm = reinterpret_cast<T*>(sk_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool single = d.K <= kc;
  auto stage = [&](int64_t kb) {
    const int64_t kw = d.K - kb < kc ? d.K - kb : kc;
    // (no div/mod by the run-time chunk length: this loop runs SP*kc/256 times per thread)
#pragma unroll 1
    for (int s = 0; s < SP; ++s) {
      const T* yrow = Y + s * d.y_ss + kb * d.y_sk;
      for (int kk = threadIdx.x; kk < kc; kk += 256) ysm[s * kc + kk] = (s < d.S && kk < kw) ? yrow[kk * d.y_sk] : T(0);
    }
  };
  if (single) {
    stage(0);
    __syncthreads();
  }
  for (int64_t row_base = (int64_t)blockIdx.x * 32; row_base < d.L; row_base += (int64_t)gridDim.x * 32) {
    T acc[RPW][SP];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int s = 0; s < SP; ++s) acc[r][s] = T(0);
    const int64_t l0 = row_base + warp * RPW;
    for (int64_t kb = 0; kb < d.K; kb += kc) {
      if (!single) {
        __syncthreads();
        stage(kb);
        __syncthreads();
      }
      const int kw = (int)(d.K - kb < kc ? d.K - kb : kc);
#pragma unroll 2
      for (int k0 = 0; k0 < kw; k0 += KSTEP) {
        const int k = k0 + lane * V;
        SkVec<T> x[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          if (k < kw && l0 + r < d.L) x[r] = *reinterpret_cast<const SkVec<T>*>(X + (l0 + r) * d.x_sl + kb + k);
          else
#pragma unroll
            for (int e = 0; e < V; ++e) x[r].v[e] = T(0);
        }
#pragma unroll
        for (int s = 0; s < SP; ++s) {
          const SkVec<T> y = *reinterpret_cast<const SkVec<T>*>(ysm + s * kc + k);
#pragma unroll
          for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[r][s] += x[r].v[e] * y.v[e];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        T v = acc[r][s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += sk_shfl_down(v, o);
        if (lane == 0 && l0 + r < d.L && s < d.S) sk_store<T>(d, C, l0 + r, s, v);
      }
    }
  }
}

// X L-major (x_sl == 1), 16-byte aligned rows of L, L % V == 0: a thread owns V consecutive l,
// blockIdx.y owns a chunk of K; the small operand is staged as [kk][SP] so one k costs SP/V
// broadcast 128-bit shared loads; 8 big-operand loads are kept in flight per thread.
template <typename T, int SP>
__global__ void __launch_bounds__(256) skinny_rs2_kernel(const T* __restrict__ X, const T* __restrict__ Y, T* __restrict__ C,
                                                         T* __restrict__ ws, int64_t kchunk, const __grid_constant__ SkinnyDesc d) {
  TCR_PDL_ENTER();
  constexpr int V = 16 / sizeof(T), KS = 128, U = 8;
  __shared__ __align__(16) T ysm[KS * SP];
  const int64_t l = ((int64_t)blockIdx.x * 256 + threadIdx.x) * V;
  const bool l_ok = l < d.L;
  const int64_t k0 = (int64_t)blockIdx.y * kchunk, k1 = k0 + kchunk < d.K ? k0 + kchunk : d.K;
  T acc[V][SP];
#pragma unroll
  for (int e = 0; e < V; ++e)
#pragma unroll
    for (int s = 0; s < SP; ++s) acc[e][s] = T(0);
  for (int64_t kb = k0; kb < k1; kb += KS) {
    const int kw = (int)(k1 - kb < KS ? k1 - kb : KS);
    __syncthreads();
    for (int e = threadIdx.x; e < KS * SP; e += 256) {
      const int kk = e / SP, sidx = e % SP;
      ysm[e] = (sidx < d.S && kk < kw) ? Y[sidx * d.y_ss + (kb + kk) * d.y_sk] : T(0);
    }
    __syncthreads();
    if (l_ok) {
      const T* xp = X + kb * d.x_sk + l;
      for (int kk = 0; kk < kw; kk += U) {
        SkVec<T> x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (kk + u < kw) x[u] = *reinterpret_cast<const SkVec<T>*>(xp + (int64_t)(kk + u) * d.x_sk);
          else
#pragma unroll
            for (int e = 0; e < V; ++e) x[u].v[e] = T(0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int krow = kk + u < KS ? kk + u : KS - 1;  // x is zero past kw, any staged row will do
#pragma unroll
          for (int s0 = 0; s0 < SP; s0 += V) {
            const SkVec<T> y = *reinterpret_cast<const SkVec<T>*>(ysm + krow * SP + s0);
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int e = 0; e < V; ++e) acc[e][s0 + j] += x[u].v[e] * y.v[j];
          }
        }
      }
    }
  }
  if (!l_ok) return;
#pragma unroll
  for (int e = 0; e < V; ++e)
#pragma unroll
    for (int s = 0; s < SP; ++s)
      if (s < d.S) {
        if (gridDim.y == 1) sk_store<T>(d, C, l + e, s, acc[e][s]);
        else ws[((int64_t)blockIdx.y * d.S + s) * d.L + l + e] = acc[e][s];
      }
}

// K <= 16, vectorised: A rows staged as [64][KP] and read back with 128-bit broadcast loads
template <typename T, int KP>
__global__ void __launch_bounds__(256) small_k2_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C,
                                                       const __grid_constant__ tcr_gemm_desc d) {
  TCR_PDL_ENTER();
  constexpr int V = 16 / sizeof(T);
  __shared__ __align__(16) T a_sm[64 * KP];
  const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const bool n_ok = n < d.n;
  const T* bias = (const T*)d.bias;
  T b[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) b[k] = (n_ok && k < d.k) ? B[k * d.b_sk + n * d.b_sn] : T(0);
  for (int64_t m0 = (int64_t)blockIdx.y * 64; m0 < d.m; m0 += (int64_t)gridDim.y * 64) {
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * KP; e += 256) {
      const int mm = e / KP, k = e % KP;
      a_sm[e] = (m0 + mm < d.m && k < d.k) ? A[(m0 + mm) * d.a_sm + k * d.a_sk] : T(0);
    }
    __syncthreads();
    if (n_ok) {
      const int rows = d.m - m0 < 64 ? (int)(d.m - m0) : 64;
#pragma unroll 4
      for (int mm = 0; mm < rows; ++mm) {
        T v = T(0);
#pragma unroll
        for (int k = 0; k < KP; k += V) {
          const SkVec<T> a = *reinterpret_cast<const SkVec<T>*>(a_sm + mm * KP + k);
#pragma unroll
          for (int e = 0; e < V; ++e) v += a.v[e] * b[k + e];
        }
        const int64_t m = m0 + mm;
        T* dst = C + m * d.c_sm + n * d.c_sn;
        if (d.accumulate) v += *dst;
        if (d.epilogue == TCR_EPI_BIAS_N) v += bias[n];
        else if (d.epilogue == TCR_EPI_BIAS_M) v += bias[m];
        if (d.activation) v = sk_act<T>(d.activation, v);
        *dst = v;
      }
    }
  }
}


// K <= 16 and a narrow output (n <= 16, e.g. the 10-9-9 DQN layers at replay batch 4096): a thread
// owns one row m; B (k x n) sits in shared memory and is read by broadcast
template <typename T, int KP>
__global__ void __launch_bounds__(256) small_kn_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C,
                                                       const __grid_constant__ tcr_gemm_desc d) {
  TCR_PDL_ENTER();
  __shared__ T b_sm[KP][SK_MAX];
  for (int e = threadIdx.x; e < KP * SK_MAX; e += 256) {
    const int k = e / SK_MAX, n = e % SK_MAX;
    b_sm[k][n] = (k < d.k && n < d.n) ? B[k * d.b_sk + n * d.b_sn] : T(0);
  }
  __syncthreads();
  const T* bias = (const T*)d.bias;
  for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < d.m; m += (int64_t)gridDim.x * 256) {
    T a[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) a[k] = k < d.k ? A[m * d.a_sm + k * d.a_sk] : T(0);
#pragma unroll
    for (int n = 0; n < SK_MAX; ++n) {
      if (n < d.n) {
        T v = T(0);
#pragma unroll
        for (int k = 0; k < KP; ++k) v += a[k] * b_sm[k][n];
        T* dst = C + m * d.c_sm + n * d.c_sn;
        if (d.accumulate) v += *dst;
        if (d.epilogue == TCR_EPI_BIAS_N) v += bias[n];
        else if (d.epilogue == TCR_EPI_BIAS_M) v += bias[m];
        if (d.activation) v = sk_act<T>(d.activation, v);
        *dst = v;
      }
    }
  }
}

template <typename T, int SP>
static int launch_rk2(const T* X, const T* Y, T* C, const SkinnyDesc& d) {
  constexpr int V = 16 / sizeof(T), KSTEP = 32 * V;
  const int64_t budget = 96 * 1024 / (SP * (int64_t)sizeof(T)) / KSTEP * KSTEP;
  int64_t kc = ceil_div(d.K, KSTEP) * KSTEP;
  if (kc > budget) kc = budget;
  const size_t smem = (size_t)SP * kc * sizeof(T);
  static bool configured = false;
  if (!configured) {
    TCR_CUDA(cudaFuncSetAttribute(skinny_rk2_kernel<T, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  int grid = wave_grid(d.L, 32, 2);
  TCR_LAUNCH((skinny_rk2_kernel<T, SP>), grid, 256, smem, X, Y, C, (int)kc, d);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

template <typename T, int SP>
static int launch_rs2(const T* X, const T* Y, T* C, const SkinnyDesc& d) {
  constexpr int V = 16 / sizeof(T);
  const int64_t bx = ceil_div(d.L, 256 * V);
  const int64_t want = (int64_t)state().sm_count * 2;
  int64_t splits = bx >= want ? 1 : ceil_div(want, bx);
  int64_t kchunk = ceil_div(ceil_div(d.K, splits), 32) * 32;
  if (kchunk < 32) kchunk = 32;
  splits = ceil_div(d.K, kchunk);
  if (splits > 65535) {
    kchunk = ceil_div(ceil_div(d.K, 65535), 32) * 32;
    splits = ceil_div(d.K, kchunk);
  }
  void* ws = nullptr;
  if (splits > 1) {
    int rc = tcr_alloc(&ws, sizeof(T) * (size_t)(splits * d.S * d.L));
    if (rc) return rc;
  }
  TCR_LAUNCH((skinny_rs2_kernel<T, SP>), dim3((unsigned)bx, (unsigned)splits), 256, 0, X, Y, C, (T*)ws, kchunk, d);
  if (splits > 1) {
    int grid = wave_grid(ceil_div(d.L, (int64_t)32) * d.S, 1, 8);
    TCR_LAUNCH((skinny_rs_reduce_kernel<T>), grid, 256, 0, (const T*)ws, C, (int)splits, d);
  }
  TCR_CHECK_LAUNCH();
  if (ws) tcr_free(ws);
  return TCR_OK;
}

#define TCR_SK_SP(S, CALL)                        \
  do {                                            \
    if ((S) <= 4) { constexpr int SP = 4; CALL; } \
    else if ((S) <= 8) { constexpr int SP = 8; CALL; } \
    else if ((S) <= 12) { constexpr int SP = 12; CALL; } \
    else { constexpr int SP = 16; CALL; }         \
  } while (0)

template <typename T>
static bool aligned16(const void* p, int64_t pitch_elems) {
  constexpr int V = 16 / sizeof(T);
  return (((uintptr_t)p) & 15) == 0 && (pitch_elems % V) == 0;
}

template <typename T>
static int run_skinny(const T* X, const T* Y, T* C, const SkinnyDesc& d) {
  constexpr int V = 16 / sizeof(T);
  static const int vec = std::getenv("TCR_SKINNY_VEC") ? std::atoi(std::getenv("TCR_SKINNY_VEC")) : 1;
  if (vec && d.x_sk == 1 && d.K >= 64 && (d.K % V) == 0 && aligned16<T>(X, d.x_sl)) {
    int rc = TCR_OK;
    TCR_SK_SP(d.S, rc = (launch_rk2<T, SP>(X, Y, C, d)));
    return rc;
  }
  if (vec && d.x_sl == 1 && d.K >= 32 && (d.L % V) == 0 && aligned16<T>(X, d.x_sk)) {
    int rc = TCR_OK;
    TCR_SK_SP(d.S, rc = (launch_rs2<T, SP>(X, Y, C, d)));
    return rc;
  }
  if (d.x_sk == 1 || d.K == 1) {
    int grid = wave_grid(d.L, 32, 4);
    TCR_LAUNCH((skinny_rk_kernel<T>), grid, 256, 0, X, Y, C, d);
    TCR_CHECK_LAUNCH();
    return TCR_OK;
  }
  // x_sl == 1
  const int64_t bx = ceil_div(d.L, 256);
  int64_t want = (int64_t)state().sm_count * 4;
  int64_t splits = bx >= want ? 1 : ceil_div(want, bx);
  if (splits > ceil_div(d.K, 64)) splits = ceil_div(d.K, 64);
  if (splits > 65535) splits = 65535;
  int64_t kchunk = ceil_div(ceil_div(d.K, splits), 64) * 64;
  splits = ceil_div(d.K, kchunk);
  void* ws = nullptr;
  if (splits > 1) {
    int rc = tcr_alloc(&ws, sizeof(T) * (size_t)(splits * d.S * d.L));
    if (rc) return rc;
  }
  TCR_LAUNCH((skinny_rs_kernel<T>), dim3((unsigned)bx, (unsigned)splits), 256, 0, X, Y, C, (T*)ws, kchunk, d);
  if (splits > 1) {
    int grid = wave_grid(ceil_div(d.L, (int64_t)32) * d.S, 1, 8);
    TCR_LAUNCH((skinny_rs_reduce_kernel<T>), grid, 256, 0, (const T*)ws, C, (int)splits, d);
  }
  TCR_CHECK_LAUNCH();
  if (ws) tcr_free(ws);
  return TCR_OK;
}

// returns handled = true when one of the streaming kernels took the problem
int gemm_skinny_dispatch(const void* a, const void* b, void* c, const tcr_gemm_desc* d, bool* handled) {
  *handled = false;
  if (d->batch != 1) return TCR_OK;
  if (d->dtype != TCR_FLOAT && d->dtype != TCR_DOUBLE && d->dtype != TCR_INT32 && d->dtype != TCR_INT64) return TCR_OK;
  const bool big = d->m * d->n * d->k >= (1ll << 16);
  if (!big) return TCR_OK;
  if (d->k <= SK_MAX && d->m * d->n >= 4096) {
    int64_t gy = ceil_div(d->m, 64);
    if (gy > 4096) gy = 4096;
    dim3 grid((unsigned)ceil_div(d->n, 256), (unsigned)gy);
    static const int vec = std::getenv("TCR_SKINNY_VEC") ? std::atoi(std::getenv("TCR_SKINNY_VEC")) : 1;
    if (vec && d->n <= SK_MAX && d->m >= 256) {
      int g1 = wave_grid(d->m, 256, 4);
      TCR_DISPATCH_COMPUTE(d->dtype, T, {
        if (d->k <= 4) TCR_LAUNCH((small_kn_kernel<T, 4>), g1, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
        else if (d->k <= 8) TCR_LAUNCH((small_kn_kernel<T, 8>), g1, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
        else if (d->k <= 12) TCR_LAUNCH((small_kn_kernel<T, 12>), g1, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
        else TCR_LAUNCH((small_kn_kernel<T, 16>), g1, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
      });
    } else if (vec) {
      TCR_DISPATCH_COMPUTE(d->dtype, T, {
        if (d->k <= 4) TCR_LAUNCH((small_k2_kernel<T, 4>), grid, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
        else if (d->k <= 8) TCR_LAUNCH((small_k2_kernel<T, 8>), grid, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
        else if (d->k <= 12) TCR_LAUNCH((small_k2_kernel<T, 12>), grid, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
        else TCR_LAUNCH((small_k2_kernel<T, 16>), grid, 256, 0, (const T*)a, (const T*)b, (T*)c, *d);
      });
    } else
    TCR_DISPATCH_COMPUTE(d->dtype, T, TCR_LAUNCH((small_k_kernel<T>), grid, 256, 0, (const T*)a, (const T*)b, (T*)c, *d));
    TCR_CHECK_LAUNCH();
    *handled = true;
    return TCR_OK;
  }
  SkinnyDesc s;
  memset(&s, 0, sizeof(s));
  s.K = d->k;
  s.activation = d->activation;
  s.accumulate = d->accumulate;
  s.has_bias = d->epilogue != TCR_EPI_NONE;
  s.bias = d->bias;
  const void *X = nullptr, *Y = nullptr;
  // both extents tiny, K long (DQN / gd_demo weight gradients: K = batch): the split-K kernel over
  // the L-major operand gives the parallelism a single 64x64 SIMT tile cannot
  const bool tiny_a = d->m < 64 && d->n <= SK_MAX && d->k >= 512 && d->a_sm == 1 && d->m >= d->n;
  const bool tiny_b = d->n < 64 && d->m <= SK_MAX && d->k >= 512 && d->b_sn == 1 && d->n > d->m;
  if (d->n <= SK_MAX && (d->m >= 64 || tiny_a) && d->k >= 32 && (d->a_sk == 1 || d->a_sm == 1)) {
    // big operand A(m,k), small operand B(k,n)
    s.L = d->m; s.S = d->n;
    s.x_sl = d->a_sm; s.x_sk = d->a_sk; s.y_ss = d->b_sn; s.y_sk = d->b_sk;
    s.c_sl = d->c_sm; s.c_ss = d->c_sn;
    s.bias_on_l = d->epilogue == TCR_EPI_BIAS_M;
    X = a; Y = b;
  } else if (d->m <= SK_MAX && (d->n >= 64 || tiny_b) && d->k >= 32 && (d->b_sk == 1 || d->b_sn == 1)) {
    // transposed view: big operand B(k,n) as X(l = n, k), small operand A(m,k) as Y(s = m, k)
    s.L = d->n; s.S = d->m;
    s.x_sl = d->b_sn; s.x_sk = d->b_sk; s.y_ss = d->a_sm; s.y_sk = d->a_sk;
    s.c_sl = d->c_sn; s.c_ss = d->c_sm;
    s.bias_on_l = d->epilogue == TCR_EPI_BIAS_N;
    X = b; Y = a;
  } else {
    return TCR_OK;
  }
  TCR_DISPATCH_COMPUTE(d->dtype, T, {
    int rc = run_skinny<T>((const T*)X, (const T*)Y, (T*)c, s);
    if (rc) return rc;
  });
  *handled = true;
  return TCR_OK;
}

}  // namespace tcr
