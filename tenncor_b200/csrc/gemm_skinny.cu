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

template <typename T>
__global__ void __launch_bounds__(256) skinny_rs_reduce_kernel(const T* __restrict__ ws, T* __restrict__ C, int splits,
                                                               const __grid_constant__ SkinnyDesc d) {
  const int64_t total = d.L * d.S, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t l = i % d.L;
    const int s = (int)(i / d.L);
    T v = T(0);
    for (int z = 0; z < splits; ++z) v += ws[((int64_t)z * d.S + s) * d.L + l];
    sk_store<T>(d, C, l, s, v);
  }
}

// K <= 16: every output is a short dot product. A thread keeps its column B(:, n) in registers,
// a block stages 64 rows of A in shared memory and streams them; consecutive threads walk n
// (coalesced stores when c_sn == 1)
template <typename T>
__global__ void __launch_bounds__(256) small_k_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C,
                                                      const __grid_constant__ tcr_gemm_desc d) {
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

template <typename T>
static int run_skinny(const T* X, const T* Y, T* C, const SkinnyDesc& d) {
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
    int grid = wave_grid(d.L * d.S, 256, 8);
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
  if (d->n <= SK_MAX && d->m >= 64 && d->k >= 32 && (d->a_sk == 1 || d->a_sm == 1)) {
    // big operand A(m,k), small operand B(k,n)
    s.L = d->m; s.S = d->n;
    s.x_sl = d->a_sm; s.x_sk = d->a_sk; s.y_ss = d->b_sn; s.y_sk = d->b_sk;
    s.c_sl = d->c_sm; s.c_ss = d->c_sn;
    s.bias_on_l = d->epilogue == TCR_EPI_BIAS_M;
    X = a; Y = b;
  } else if (d->m <= SK_MAX && d->n >= 64 && d->k >= 32 && (d->b_sk == 1 || d->b_sn == 1)) {
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
