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

#include <type_traits>

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
    if (d.y_ss == 1 && d.y_sk == d.S) {
      // the small operand is one contiguous [K][S] block (a dense layer's weights, S = its output width)
      // a thread takes whole k rows: S consecutive elements each (independent loads, no division), written down the columns
      const T* src = Y + kb * d.y_sk;
      const int S = (int)d.S;
      for (int k = threadIdx.x; k < kc; k += 256) {
        T v[SP];
#pragma unroll
        for (int s2 = 0; s2 < SP; ++s2) v[s2] = (s2 < S && k < kw) ? src[(int64_t)k * S + s2] : T(0);
#pragma unroll
        for (int s2 = 0; s2 < SP; ++s2) ysm[s2 * kc + k] = v[s2];
      }
      return;
    }
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
  // every block owns an equal share of the 32-row groups' worth of work in units of 4 rows (one warp pass): a grid of whole
  // waves with 32-row blocks left 40 % of the SMs with half the work of the others
  const int64_t quads = (d.L + RPW - 1) / RPW;
  const int64_t q_begin = quads * blockIdx.x / gridDim.x, q_end = quads * (blockIdx.x + 1) / gridDim.x;
  for (int64_t row_base = q_begin * RPW; row_base < q_end * RPW; row_base += 32) {
    T acc[RPW][SP];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int s = 0; s < SP; ++s) acc[r][s] = T(0);
    const int64_t l0 = row_base + warp * RPW;
    const int64_t l_end = q_end * RPW < d.L ? q_end * RPW : d.L;  // rows past the block's share belong to the next block
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
          if (k < kw && l0 + r < l_end) x[r] = *reinterpret_cast<const SkVec<T>*>(X + (l0 + r) * d.x_sl + kb + k);
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
        if (lane == 0 && l0 + r < l_end && s < d.S) sk_store<T>(d, C, l0 + r, s, v);
      }
    }
  }
}

// fp32, X K-major with 16-byte aligned rows, K % 4 == 0 and the whole small operand in shared memory (SP * K * 4 <= 96 KB).
// ncu on the 8192 x 10 x 1024 output layer (profiles/r2_ncu_skinny.md) showed skinny_rk2 latency-bound, not bandwidth-bound:
// 33.5 MB in 28 us, 1.2-1.5 IPC at 21 % occupancy, every warp walking the same sequence of dependent phases (stage the small
// operand -> barrier -> four rounds of load 4 rows / wait / FMA -> shuffle -> store) with ~one 4-row group per warp, so no
// phase overlapped another. Here a warp owns TWO rows and issues all of their loads (up to 8 k-steps = 16 x 16 bytes per
// lane) BEFORE the block stages the small operand and synchronises: one exposed memory latency instead of five.
template <int SP>
__global__ void __launch_bounds__(256, 2) skinny_rk3_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ C,
                                                            int kc, const __grid_constant__ SkinnyDesc d) {
  TCR_PDL_ENTER();
  constexpr int RPW = 2, PF = 8;  // rows per warp pass, k-steps prefetched per row
  extern __shared__ __align__(16) unsigned char sk_smem[];
  float* ysm = reinterpret_cast<float*>(sk_smem);  // [SP][kc], zero padded in s and k
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = (int)d.K, S = (int)d.S;
  const int steps = (K + 127) / 128;  // lanes beyond K in the last step read zeros from ysm: clamp their x address instead of predicating
  const int64_t pairs = (d.L + RPW - 1) / RPW;
  const int64_t p_begin = pairs * blockIdx.x / gridDim.x, p_end = pairs * (blockIdx.x + 1) / gridDim.x;
  bool staged = false;
  for (int64_t pr = p_begin + warp; pr < p_end || !staged; pr += 8) {
    const bool active = pr < p_end;
    const float* xr[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      int64_t row = pr * RPW + r;
      if (row > d.L - 1) row = d.L - 1;  // clamped: computed twice, stored once
      xr[r] = X + row * d.x_sl;
    }
    float acc[RPW][SP];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int s2 = 0; s2 < SP; ++s2) acc[r][s2] = 0.f;
    for (int k0 = 0; k0 < steps; k0 += PF) {
      float4 x[PF][RPW];
      if (active) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          int koff = (k0 + u) * 128 + lane * 4;
          if (koff > K - 4) koff = K - 4;  // past the end (last ragged step or u beyond `steps`): any valid address, multiplied by zeros
#pragma unroll
          for (int r = 0; r < RPW; ++r) x[u][r] = __ldg(reinterpret_cast<const float4*>(xr[r] + koff));
        }
      }
      if (!staged) {  // first pass: the small operand is staged while the big operand's loads are in flight
        if (d.y_ss == 1 && d.y_sk == d.S) {  // contiguous [K][S] weights: whole rows per thread, independent loads
          for (int k = threadIdx.x; k < kc; k += 256) {
            float v[SP];
#pragma unroll
            for (int s2 = 0; s2 < SP; ++s2) v[s2] = (s2 < S && k < K) ? Y[(int64_t)k * S + s2] : 0.f;
#pragma unroll
            for (int s2 = 0; s2 < SP; ++s2) ysm[s2 * kc + k] = v[s2];
          }
        } else {
          for (int e = threadIdx.x; e < SP * kc; e += 256) {
            const int s2 = e / kc, k = e - s2 * kc;
            ysm[e] = (s2 < S && k < K) ? Y[s2 * d.y_ss + k * d.y_sk] : 0.f;
          }
        }
        __syncthreads();
        staged = true;
      }
      if (active) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          if (k0 + u < steps) {
            const float* yp = ysm + (k0 + u) * 128 + lane * 4;  // zero beyond K
#pragma unroll
            for (int s2 = 0; s2 < SP; ++s2) {
              const float4 y = *reinterpret_cast<const float4*>(yp + s2 * kc);
#pragma unroll
              for (int r = 0; r < RPW; ++r) {
                float a = acc[r][s2];
                a = fmaf(x[u][r].x, y.x, a);
                a = fmaf(x[u][r].y, y.y, a);
                a = fmaf(x[u][r].z, y.z, a);
                a = fmaf(x[u][r].w, y.w, a);
                acc[r][s2] = a;
              }
            }
          }
        }
      }
    }
    if (!active) continue;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
#pragma unroll
      for (int s2 = 0; s2 < SP; ++s2) {
        float v = acc[r][s2];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][s2] = v;
      }
    }
    // every lane holds every sum: lane (r * SP + s2) stores element (row r, column s2)
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int s2 = 0; s2 < SP; ++s2)
        if (lane == r * SP + s2) mine = acc[r][s2];
    if (lane < RPW * SP) {
      const int r = lane / SP, s2 = lane % SP;
      const int64_t row = pr * RPW + r;
      if (row < d.L && s2 < S) sk_store<float>(d, C, row, s2, mine);
    }
  }
}

template <int SP>
static int launch_rk3(const float* X, const float* Y, float* C, const SkinnyDesc& d) {
  const int kc = (int)(ceil_div(d.K, 128) * 128);
  const size_t smem = (size_t)SP * kc * sizeof(float);
  static bool configured = false;
  if (!configured) {
    TCR_CUDA(cudaFuncSetAttribute(skinny_rk3_kernel<SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  const int64_t pairs = ceil_div(d.L, 2);
  int64_t want = pairs / 8 < 1 ? 1 : pairs / 8;  // at least one pass of the 8 warps per block
  const int64_t cap = (int64_t)state().sm_count * 2;
  const int grid = (int)(want < cap ? want : cap);
  TCR_LAUNCH((skinny_rk3_kernel<SP>), grid, 256, smem, X, Y, C, kc, d);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
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


// K <= 16, wide output with unit column stride, fp32: a thread owns FOUR consecutive columns (B(:, n..n+3) in registers) and
// walks 32 rows staged in shared memory; every store is 16 bytes (the 4-byte-store form above issues four times the
// instructions per byte and reached 2.2 TB/s writing the 33.5 MB input gradient of the 1024 -> 10 layer). Optional post-op:
// multiply by the activation's derivative read from `aux` — the dX product of a dense layer and the SIGMOID / TANH gradient
// rule behind it in ONE pass over memory instead of write + read + read + write.
template <int KP>
__global__ void __launch_bounds__(256) small_k4_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                       const __grid_constant__ tcr_gemm_desc d) {
  TCR_PDL_ENTER();
  constexpr int RB = 32;
  __shared__ __align__(16) float a_sm[RB * KP];
  const int64_t n = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  const bool n_ok = n < d.n;  // n % 4 == 0: all four columns or none
  const float* bias = (const float*)d.bias;
  const float* aux = (const float*)d.aux;
  float b[KP][4];
#pragma unroll
  for (int k = 0; k < KP; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) b[k][e] = (n_ok && k < d.k) ? B[k * d.b_sk + (n + e) * d.b_sn] : 0.f;
  float bn[4] = {0.f, 0.f, 0.f, 0.f};
  if (n_ok && d.epilogue == TCR_EPI_BIAS_N)
#pragma unroll
    for (int e = 0; e < 4; ++e) bn[e] = bias[n + e];
  for (int64_t m0 = (int64_t)blockIdx.y * RB; m0 < d.m; m0 += (int64_t)gridDim.y * RB) {
    __syncthreads();
    for (int e = threadIdx.x; e < RB * KP; e += 256) {
      const int mm = e / KP, k = e % KP;
      a_sm[e] = (m0 + mm < d.m && k < d.k) ? A[(m0 + mm) * d.a_sm + k * d.a_sk] : 0.f;
    }
    __syncthreads();
    if (!n_ok) continue;
    const int rows = d.m - m0 < RB ? (int)(d.m - m0) : RB;
    for (int mb = 0; mb < rows; mb += 8) {
    float4 xs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)  // eight 16-byte loads in flight per thread before any arithmetic
      xs[u] = (d.post_op && mb + u < rows) ? __ldg(reinterpret_cast<const float4*>(aux + (m0 + mb + u) * d.c_sm + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int mm = mb + u;
      if (mm >= rows) break;
      const int64_t m = m0 + mm;
      const float4 x = xs[u];
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < KP; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(a_sm + mm * KP + k);
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] += a.x * b[k][e] + a.y * b[k + 1][e] + a.z * b[k + 2][e] + a.w * b[k + 3][e];
      }
      float* dst = C + m * d.c_sm + n;
      if (d.accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(dst);
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
      }
      const float bm = d.epilogue == TCR_EPI_BIAS_M ? bias[m] : 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[e] += bn[e] + bm;
        if (d.activation) v[e] = sk_act<float>(d.activation, v[e]);
      }
      if (d.post_op == TCR_POST_MUL_DSIGMOID) {
        v[0] *= x.x * (1.f - x.x); v[1] *= x.y * (1.f - x.y); v[2] *= x.z * (1.f - x.z); v[3] *= x.w * (1.f - x.w);
      } else if (d.post_op == TCR_POST_MUL_DTANH) {
        v[0] *= 1.f - x.x * x.x; v[1] *= 1.f - x.y * x.y; v[2] *= 1.f - x.z * x.z; v[3] *= 1.f - x.w * x.w;
      }
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
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
  // two blocks per SM, each with an equal share of the 4-row groups (a block stages the small operand once, so few large
  // blocks; at least four row groups each so that the staging is amortised)
  const int64_t quads = ceil_div(d.L, 4);
  int64_t want = quads / 4 < 1 ? 1 : quads / 4;
  const int64_t cap = (int64_t)state().sm_count * 2;
  int grid = (int)(want < cap ? want : cap);
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
    if constexpr (std::is_same<T, float>::value) {
      static const int lean = std::getenv("TCR_SKINNY_RK3") ? std::atoi(std::getenv("TCR_SKINNY_RK3")) : 1;
      const int64_t sp = d.S <= 4 ? 4 : d.S <= 8 ? 8 : d.S <= 12 ? 12 : 16;
      if (lean && sp * ceil_div(d.K, 128) * 128 * 4 <= 96 * 1024 && d.K < (1 << 24)) {
        TCR_SK_SP(d.S, rc = (launch_rk3<SP>(X, Y, C, d)));
        return rc;
      }
    }
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
    const bool k4 = d->dtype == TCR_FLOAT && d->c_sn == 1 && (d->n % 4) == 0 && (d->c_sm % 4) == 0 && d->n >= 64 && ((((uintptr_t)c) & 15) == 0) &&
                    (!d->post_op || ((((uintptr_t)d->aux) & 15) == 0));
    if (d->post_op && !k4) {
      set_error("tcr_gemm: post_op needs an fp32 product with k <= 16, unit column stride, n %% 4 == 0 and 16-byte aligned C / aux");
      return TCR_ERR_UNSUPPORTED;
    }
    if (vec && k4) {
      int64_t gy4 = ceil_div(d->m, 32);
      if (gy4 > 65535) gy4 = 65535;
      dim3 grid4((unsigned)ceil_div(d->n, 1024), (unsigned)gy4);
      const float *fa = (const float*)a, *fb = (const float*)b;
      float* fc = (float*)c;
      if (d->k <= 4) TCR_LAUNCH((small_k4_kernel<4>), grid4, 256, 0, fa, fb, fc, *d);
      else if (d->k <= 8) TCR_LAUNCH((small_k4_kernel<8>), grid4, 256, 0, fa, fb, fc, *d);
      else if (d->k <= 12) TCR_LAUNCH((small_k4_kernel<12>), grid4, 256, 0, fa, fb, fc, *d);
      else TCR_LAUNCH((small_k4_kernel<16>), grid4, 256, 0, fa, fb, fc, *d);
    } else if (vec && d->n <= SK_MAX && d->m >= 256) {
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
