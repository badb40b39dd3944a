// gemm.cu — exact (SIMT FMA) GEMM with fully strided operands, generic tensor
// contraction and N-d valid correlation.
//
// Replaces internal/eigen/operator.hpp:1069-1187 for every dtype. fp32 tensor-core
// modes (TF32 / 3xTF32 on tcgen05 with TMEM accumulators) live in gemm_tc.cu and are
// selected by tcr_gemm_desc.precision; this file is the bit-faithful path used for
// double / int32 / int64 (the reference's golden tests are double and int32) and for
// shapes the tensor-core kernel does not take.
#include "common.cuh"

namespace tcr {

int gemm_tc_dispatch(const void* a, const void* b, void* c, const tcr_gemm_desc* d, bool* handled);      // gemm_tc.cu
int gemm_skinny_dispatch(const void* a, const void* b, void* c, const tcr_gemm_desc* d, bool* handled);  // gemm_skinny.cu

template <typename T> __device__ __forceinline__ T act_apply(int act, T x) { return x; }
template <> __device__ __forceinline__ float act_apply(int act, float x) {
  if (act == TCR_EW_SIGMOID) return 1.0f / (1.0f + expf(-x));
  if (act == TCR_EW_TANH) return tanhf(x);
  return x;
}
template <> __device__ __forceinline__ double act_apply(int act, double x) {
  if (act == TCR_EW_SIGMOID) return 1.0 / (1.0 + exp(-x));
  if (act == TCR_EW_TANH) return tanh(x);
  return x;
}

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C,
                                                        const __grid_constant__ tcr_gemm_desc d) {
  TCR_PDL_ENTER();
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t bz = blockIdx.z;
  A += bz * d.a_sb;
  B += bz * d.b_sb;
  C += bz * d.c_sb;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  const bool a_kfast = d.a_sk == 1;  // walk the unit-stride rank with consecutive threads
  const bool b_nfast = d.b_sn == 1;
  for (int64_t k0 = 0; k0 < d.k; k0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {
      int mm, kk;
      if (a_kfast) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < d.m && k < d.k) ? A[m * d.a_sm + k * d.a_sk] : T(0);
    }
#pragma unroll
    for (int e = tid; e < BN * BK; e += 256) {
      int nn, kk;
      if (b_nfast) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < d.n && k < d.k) ? B[k * d.b_sk + n * d.b_sn] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  const T* bias = (const T*)d.bias;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= d.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= d.n) continue;
      T v = acc[i][j];
      T* dst = C + m * d.c_sm + n * d.c_sn;
      if (d.accumulate) v += *dst;
      if (d.epilogue == TCR_EPI_BIAS_N) v += bias[n];
      else if (d.epilogue == TCR_EPI_BIAS_M) v += bias[m];
      if (d.activation) v = act_apply<T>(d.activation, v);
      *dst = v;
    }
  }
}

// ---- generic contraction: out dims = b-free (in order) then a-free
struct ContractDesc {
  int n_bfree, n_afree, n_pairs;
  int64_t bfree_ext[8], bfree_stride[8];
  int64_t afree_ext[8], afree_stride[8];
  int64_t pair_ext[8], pair_astride[8], pair_bstride[8];
  int64_t n_out, n_red;
};

template <typename T>
__global__ void __launch_bounds__(256) contract_generic_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                               T* __restrict__ out, const __grid_constant__ ContractDesc d) {
  TCR_PDL_ENTER();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < d.n_out; o += stride) {
    int64_t t = o, aoff = 0, boff = 0;
    for (int k = 0; k < d.n_bfree; ++k) { int64_t c = t % d.bfree_ext[k]; t /= d.bfree_ext[k]; boff += c * d.bfree_stride[k]; }
    for (int k = 0; k < d.n_afree; ++k) { int64_t c = t % d.afree_ext[k]; t /= d.afree_ext[k]; aoff += c * d.afree_stride[k]; }
    T acc = T(0);
    for (int64_t q = 0; q < d.n_red; ++q) {
      int64_t u = q, ao = aoff, bo = boff;
      for (int k = 0; k < d.n_pairs; ++k) { int64_t c = u % d.pair_ext[k]; u /= d.pair_ext[k]; ao += c * d.pair_astride[k]; bo += c * d.pair_bstride[k]; }
      acc += a[ao] * b[bo];
    }
    out[o] = acc;
  }
}

// ---- N-d valid correlation
struct ConvDesc {
  int64_t out_shape[8], img_stride[8];
  int64_t kern_shape[8];
  int32_t order[8];  // kernel rank i slides along image rank order[i]
  int64_t n_out, n_kern;
};

template <typename T>
__global__ void __launch_bounds__(256) conv_generic_kernel(const T* __restrict__ img, const T* __restrict__ kern,
                                                           T* __restrict__ out, const __grid_constant__ ConvDesc d) {
  TCR_PDL_ENTER();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < d.n_out; o += stride) {
    int64_t t = o, base = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { int64_t c = t % d.out_shape[k]; t /= d.out_shape[k]; base += c * d.img_stride[k]; }
    T acc = T(0);
    for (int64_t q = 0; q < d.n_kern; ++q) {
      int64_t u = q, off = base;
#pragma unroll
      for (int k = 0; k < 8; ++k) { int64_t c = u % d.kern_shape[k]; u /= d.kern_shape[k]; off += c * d.img_stride[d.order[k]]; }
      acc += img[off] * kern[q];
    }
    out[o] = acc;
  }
}

}  // namespace tcr

using namespace tcr;

extern "C" {

int tcr_gemm(const void* a, const void* b, void* c, const tcr_gemm_desc* desc) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(a && b && c && desc, "tcr_gemm: null argument");
  TCR_ARG(desc->m >= 0 && desc->n >= 0 && desc->k >= 0 && desc->batch >= 0, "tcr_gemm: negative extent");
  TCR_ARG(desc->epilogue == TCR_EPI_NONE || desc->bias != nullptr, "tcr_gemm: bias epilogue without bias pointer");
  if (desc->m == 0 || desc->n == 0 || desc->batch == 0) return TCR_OK;
  TCR_ARG(desc->batch <= 65535, "tcr_gemm: batch %lld exceeds grid.z", (long long)desc->batch);
  {
    // skinny products (an extent <= 16) are HBM-bound: streaming kernels, exact FMA in the element type
    bool handled = false;
    int rc = gemm_skinny_dispatch(a, b, c, desc, &handled);
    if (rc) return rc;
    if (handled) return TCR_OK;
  }
  if (desc->post_op != TCR_POST_NONE) {
    set_error("tcr_gemm: post_op is only implemented for products with k <= 16 (the small-K streaming kernel)");
    return TCR_ERR_UNSUPPORTED;
  }
  if (desc->precision != TCR_GEMM_EXACT) {
    TCR_ARG(desc->dtype == TCR_FLOAT, "tcr_gemm: tensor-core precisions are fp32 only");
    bool handled = false;
    int rc = gemm_tc_dispatch(a, b, c, desc, &handled);
    if (rc) return rc;
    if (handled) return TCR_OK;
    // shapes the tcgen05 kernel does not take fall through to the exact SIMT kernel (still on device)
  }
  dim3 grid((unsigned)ceil_div(desc->n, BN), (unsigned)ceil_div(desc->m, BM), (unsigned)desc->batch);
  TCR_ARG(grid.y <= 65535, "tcr_gemm: m too large for the SIMT grid");
  TCR_DISPATCH_COMPUTE(desc->dtype, T, TCR_LAUNCH((gemm_simt_kernel<T>), grid, 256, 0, (const T*)a, (const T*)b, (T*)c, *desc));
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_contract(const void* a, const void* b, void* out, const int64_t a_shape[8], const int64_t b_shape[8],
                 const int32_t* pairs, int npairs, int dtype) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(a && b && out && a_shape && b_shape, "tcr_contract: null argument");
  TCR_ARG(npairs >= 0 && npairs <= 8 && (npairs == 0 || pairs), "tcr_contract: bad pairs");
  ContractDesc d;
  memset(&d, 0, sizeof(d));
  int64_t astride[8], bstride[8], sa = 1, sb = 1;
  bool acom[8] = {false}, bcom[8] = {false};
  for (int k = 0; k < 8; ++k) { astride[k] = sa; sa *= a_shape[k]; bstride[k] = sb; sb *= b_shape[k]; }
  d.n_red = 1;
  for (int p = 0; p < npairs; ++p) {
    int ra = pairs[2 * p], rb = pairs[2 * p + 1];
    // a pair naming a rank >= 8 is the reference's "no common rank" marker
    // (tenncor/eteq/backprop.hpp:347-352): an outer product
    if (ra >= 8 || rb >= 8) continue;
    TCR_ARG(ra >= 0 && rb >= 0, "tcr_contract: negative rank");
    TCR_ARG(a_shape[ra] == b_shape[rb], "tcr_contract: common dimensions do not match (a rank %d = %lld, b rank %d = %lld)", ra, (long long)a_shape[ra], rb, (long long)b_shape[rb]);
    TCR_ARG(!acom[ra] && !bcom[rb], "tcr_contract: contraction dimensions must be unique for each side");
    acom[ra] = bcom[rb] = true;
    d.pair_ext[d.n_pairs] = a_shape[ra];
    d.pair_astride[d.n_pairs] = astride[ra];
    d.pair_bstride[d.n_pairs] = bstride[rb];
    d.n_red *= a_shape[ra];
    d.n_pairs++;
  }
  d.n_out = 1;
  for (int k = 0; k < 8; ++k)
    if (!bcom[k] && b_shape[k] != 1) { d.bfree_ext[d.n_bfree] = b_shape[k]; d.bfree_stride[d.n_bfree] = bstride[k]; d.n_bfree++; d.n_out *= b_shape[k]; }
  for (int k = 0; k < 8; ++k)
    if (!acom[k] && a_shape[k] != 1) { d.afree_ext[d.n_afree] = a_shape[k]; d.afree_stride[d.n_afree] = astride[k]; d.n_afree++; d.n_out *= a_shape[k]; }
  if (d.n_out == 0) return TCR_OK;
  int grid = wave_grid(d.n_out, 256, 8);
  TCR_DISPATCH_COMPUTE(dtype, T, TCR_LAUNCH((contract_generic_kernel<T>), grid, 256, 0, (const T*)a, (const T*)b, (T*)out, d));
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_conv(const void* image, const void* kernel, void* out, const int64_t img_shape[8], const int64_t kern_shape[8],
             const int32_t order[8], int dtype) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(image && kernel && out && img_shape && kern_shape && order, "tcr_conv: null argument");
  ConvDesc d;
  memset(&d, 0, sizeof(d));
  bool seen[8] = {false};
  int64_t s = 1;
  for (int k = 0; k < 8; ++k) { d.img_stride[k] = s; s *= img_shape[k]; d.out_shape[k] = img_shape[k]; }
  d.n_kern = 1;
  for (int k = 0; k < 8; ++k) {
    TCR_ARG(order[k] >= 0 && order[k] < 8 && !seen[order[k]], "tcr_conv: convolution does not support repeated kernel dimensions");
    seen[order[k]] = true;
    d.order[k] = order[k];
    d.kern_shape[k] = kern_shape[k];
    d.n_kern *= kern_shape[k];
    TCR_ARG(kern_shape[k] <= img_shape[order[k]], "tcr_conv: kernel larger than image at kernel rank %d", k);
    d.out_shape[order[k]] = img_shape[order[k]] - kern_shape[k] + 1;
  }
  d.n_out = 1;
  for (int k = 0; k < 8; ++k) d.n_out *= d.out_shape[k];
  if (d.n_out == 0) return TCR_OK;
  int grid = wave_grid(d.n_out, 256, 8);
  TCR_DISPATCH_COMPUTE(dtype, T, TCR_LAUNCH((conv_generic_kernel<T>), grid, 256, 0, (const T*)image, (const T*)kernel, (T*)out, d));
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

}  // extern "C"
