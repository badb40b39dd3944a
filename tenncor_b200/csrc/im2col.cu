// im2col.cu — patch gather that turns the reference's conv2d composite into a GEMM operand.
//
// The reference has no multi-filter convolution opcode: nn.conv2d (cfg/tenncor/nn.yml:48-98)
// zero-pads the image along a fresh rank by (out-1, out-1) and runs the single-kernel N-d valid
// correlation CONV (internal/eigen/operator.hpp:1143-1187) with the reversed kernel sliding along
// that rank, so every output channel is one position of the slide and (2*out-1)/1 of the products
// multiply a zero. The planner recognises the composite (planner.cpp, fuse_convs) and computes
//     out[(x,y,b), o] = sum_{c,i,j} img[c, x+i, y+j, b] * k[o, c, i, j]
// as cols[(x,y,b), (c,i,j)] . k[(c,i,j), o] on the tcgen05 GEMM. This file produces `cols`.
//
// HBM-bound: algorithmic bytes = 4 * (rows * pitch written + image read once).
#include <cstring>

#include "common.cuh"

namespace tcr {

struct Im2colDesc {
  int n_pos, n_win;
  int64_t rows, k, pitch4;       // pitch4 = row pitch in 16-byte groups
  int64_t pos_ext[8], pos_stride[8];
  int64_t win_ext[8], win_stride[8];
};

// generic form: one thread per 16-byte group of a row, 64-bit coordinate arithmetic (any size)
__global__ void __launch_bounds__(256) im2col_kernel(const uint32_t* __restrict__ img, uint4* __restrict__ cols, Im2colDesc d) {
  TCR_PDL_ENTER();
  const int64_t total = d.rows * d.pitch4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = t / d.pitch4, g = t - row * d.pitch4;
    int64_t base = 0, r = row;
#pragma unroll 1
    for (int q = 0; q < d.n_pos; ++q) {
      const int64_t c = r % d.pos_ext[q];
      r /= d.pos_ext[q];
      base += c * d.pos_stride[q];
    }
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int64_t kk = g * 4 + e;
      v[e] = 0u;
      if (kk < d.k) {
        int64_t off = base;
#pragma unroll 1
        for (int q = 0; q < d.n_win; ++q) {
          const int64_t c = kk % d.win_ext[q];
          kk /= d.win_ext[q];
          off += c * d.win_stride[q];
        }
        v[e] = __ldg(img + off);
      }
    }
    cols[t] = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

// Tiled form (image < 2^31 elements, window offsets fit shared memory). A block owns IM_ROWS consecutive rows
// per iteration: the window offsets of a row are the same for every row (computed once per block into shared
// memory), the row bases once per tile, so the copy loop is one shared-memory lookup, the loads and one 16-byte
// store per thread with no division chains. VEC: every run of the window is a multiple of 4 elements and
// 16-byte aligned in the image, so the four elements of a group are one 16-byte load.
constexpr int IM_ROWS = 32;

template <bool VEC>
__global__ void __launch_bounds__(256) im2col_tiled_kernel(const uint32_t* __restrict__ img, uint4* __restrict__ cols, Im2colDesc d) {
  TCR_PDL_ENTER();
  extern __shared__ int32_t woff[];  // [pitch4 * 4] window offsets (-1: zero fill), then IM_ROWS row bases
  const int pitch4 = (int)d.pitch4;
  int32_t* base = woff + pitch4 * 4;
  for (int kk = threadIdx.x; kk < pitch4 * 4; kk += blockDim.x) {
    int32_t off = -1;
    if (kk < (int)d.k) {
      int rem = kk;
      off = 0;
      for (int q = 0; q < d.n_win; ++q) {
        const int e = (int)d.win_ext[q];
        off += (rem % e) * (int32_t)d.win_stride[q];
        rem /= e;
      }
    }
    woff[kk] = off;
  }
  const int64_t n_tiles = (d.rows + IM_ROWS - 1) / IM_ROWS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();  // woff ready (first pass) / previous tile done with `base`
    const int64_t row0 = tile * IM_ROWS;
    if (threadIdx.x < IM_ROWS) {
      int64_t r = row0 + threadIdx.x;
      int32_t b = 0;
      if (r < d.rows) {
        for (int q = 0; q < d.n_pos; ++q) {
          const int64_t e = d.pos_ext[q];
          b += (int32_t)(r % e) * (int32_t)d.pos_stride[q];
          r /= e;
        }
      }
      base[threadIdx.x] = b;
    }
    __syncthreads();
    const int rows_here = (int)(d.rows - row0 < IM_ROWS ? d.rows - row0 : IM_ROWS);
    const int work = rows_here * pitch4;
    uint4* out = cols + row0 * pitch4;
    for (int t = threadIdx.x; t < work; t += blockDim.x) {
      const int rr = t / pitch4, g = t - rr * pitch4;
      const int32_t b = base[rr];
      const int4 o = *reinterpret_cast<const int4*>(woff + 4 * g);
      uint4 v;
      if (VEC) {
        v = o.x >= 0 ? __ldg(reinterpret_cast<const uint4*>(img + b + o.x)) : make_uint4(0u, 0u, 0u, 0u);
      } else {
        v.x = o.x >= 0 ? __ldg(img + b + o.x) : 0u;
        v.y = o.y >= 0 ? __ldg(img + b + o.y) : 0u;
        v.z = o.z >= 0 ? __ldg(img + b + o.z) : 0u;
        v.w = o.w >= 0 ? __ldg(img + b + o.w) : 0u;
      }
      out[t] = v;
    }
  }
}

// col2im: the adjoint of the patch gather. image[u] = sum over window coordinates w with a valid position u - w of
// cols[position(u - w)][w]. One thread per image element; the terms are added in a fixed order (deterministic, no
// atomics). Ranks whose window covers the whole rank have one position (w = u), ranks with a unit window have w = 0:
// only the sliding ranks (kernel width / height) are looped over. The window offsets of the sliding ranks are the
// same for every element: a per-block table in shared memory (w per sliding rank + column offset), so the inner loop
// is a range test, one multiply-add per sliding rank and the load. 32-bit coordinates (image and rows < 2^31).
struct Col2imDesc {
  int nd;                       // non-singular image ranks, ascending
  int n_slide;
  int64_t n_img, pitch;
  int32_t n_combo;
  int32_t ext[8], pos[8], win[8];
  int32_t row_stride[8], col_stride[8];
  int32_t slide[8];             // indices (into the nd ranks) of the sliding ranks
};

constexpr int C2I_MAX_SLIDE = 4;

// VEC: the fastest image rank is spanned by the window (one position), so 4 consecutive image elements are 4
// consecutive columns of the same rows: one 16-byte load per term and a quarter of the instructions per byte.
template <bool VEC>
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ cols, float* __restrict__ img, Col2imDesc d) {
  TCR_PDL_ENTER();
  extern __shared__ int32_t table[];  // n_combo x (C2I_MAX_SLIDE window coordinates + column offset)
  constexpr int TW = C2I_MAX_SLIDE + 1;
  for (int c = threadIdx.x; c < d.n_combo; c += blockDim.x) {
    int rem = c;
    int32_t col = 0;
    for (int k = 0; k < C2I_MAX_SLIDE; ++k) {
      int32_t w = 0;
      if (k < d.n_slide) {
        const int q = d.slide[k];
        w = rem % d.win[q];
        rem /= d.win[q];
        col += w * d.col_stride[q];
      }
      table[c * TW + k] = w;
    }
    table[c * TW + C2I_MAX_SLIDE] = col;
  }
  __syncthreads();
  int32_t s_pos[C2I_MAX_SLIDE], s_rs[C2I_MAX_SLIDE];
  int s_q[C2I_MAX_SLIDE];
#pragma unroll
  for (int k = 0; k < C2I_MAX_SLIDE; ++k) {
    s_q[k] = k < d.n_slide ? d.slide[k] : 0;
    s_pos[k] = k < d.n_slide ? d.pos[s_q[k]] : 1;
    s_rs[k] = k < d.n_slide ? d.row_stride[s_q[k]] : 0;
  }
  const uint32_t n_items = (uint32_t)(VEC ? d.n_img / 4 : d.n_img);
  for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
    const uint32_t t = VEC ? item * 4 : item;
    uint32_t rem = t;
    int32_t row0 = 0, col0 = 0, us[C2I_MAX_SLIDE] = {0, 0, 0, 0};
#pragma unroll 1
    for (int q = 0; q < d.nd; ++q) {
      const uint32_t e = (uint32_t)d.ext[q];
      const int32_t u = (int32_t)(rem % e);
      rem /= e;
      if (d.pos[q] == 1) col0 += u * d.col_stride[q];        // the window spans the rank: w = u
      else if (d.win[q] == 1) row0 += u * d.row_stride[q];   // unit window: w = 0
      else {
#pragma unroll
        for (int k = 0; k < C2I_MAX_SLIDE; ++k)
          if (k < d.n_slide && s_q[k] == q) us[k] = u;
      }
    }
    const float* base = cols + col0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int c = 0; c < d.n_combo; ++c) {
      const int32_t* e = table + c * TW;
      int32_t row = row0;
      bool ok = true;
#pragma unroll
      for (int k = 0; k < C2I_MAX_SLIDE; ++k) {
        const int32_t p = us[k] - e[k];
        ok = ok && (uint32_t)p < (uint32_t)s_pos[k];
        row += p * s_rs[k];
      }
      if (ok) {
        const float* src = base + (int64_t)row * d.pitch + e[C2I_MAX_SLIDE];
        if (VEC) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        } else {
          acc.x += __ldg(src);
        }
      }
    }
    if (VEC) *reinterpret_cast<float4*>(img + t) = acc;
    else img[t] = acc.x;
  }
}

}  // namespace tcr

using namespace tcr;

extern "C" {

int tcr_im2col(const void* image, void* cols, const int64_t img_shape[8], const int64_t win_shape[8], int64_t row_pitch, int elem_size) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(image && cols && img_shape && win_shape, "tcr_im2col: null argument");
  TCR_ARG(elem_size == 4, "tcr_im2col: 4-byte elements only (got %d)", elem_size);
  TCR_ARG((((uintptr_t)cols) & 15) == 0, "tcr_im2col: cols must be 16-byte aligned");
  Im2colDesc d;
  memset(&d, 0, sizeof(d));
  d.rows = 1;
  d.k = 1;
  int64_t stride = 1;
  bool vec = (((uintptr_t)image) & 15) == 0;
  for (int r = 0; r < 8; ++r) {
    TCR_ARG(img_shape[r] >= 1 && win_shape[r] >= 1 && win_shape[r] <= img_shape[r], "tcr_im2col: window %lld does not fit image extent %lld at rank %d",
            (long long)win_shape[r], (long long)img_shape[r], r);
    const int64_t pos = img_shape[r] - win_shape[r] + 1;
    if (pos > 1) {
      d.pos_ext[d.n_pos] = pos; d.pos_stride[d.n_pos] = stride; d.n_pos++; d.rows *= pos;
      if (stride % 4 != 0) vec = false;
    }
    if (win_shape[r] > 1) {
      // a window that covers a whole rank continues the run of the rank below it
      if (d.n_win > 0 && d.win_stride[d.n_win - 1] * d.win_ext[d.n_win - 1] == stride) d.win_ext[d.n_win - 1] *= win_shape[r];
      else { d.win_ext[d.n_win] = win_shape[r]; d.win_stride[d.n_win] = stride; d.n_win++; }
      d.k *= win_shape[r];
    }
    stride *= img_shape[r];
  }
  TCR_ARG(row_pitch >= d.k && row_pitch % 4 == 0, "tcr_im2col: row pitch %lld must be a multiple of 4 and at least %lld", (long long)row_pitch, (long long)d.k);
  d.pitch4 = row_pitch / 4;
  // 16-byte loads: the fastest run starts at stride 1, is a multiple of 4 long, and every other run starts 16-byte aligned
  if (d.n_win == 0 || d.win_stride[0] != 1 || d.win_ext[0] % 4 != 0) vec = false;
  for (int q = 1; q < d.n_win; ++q)
    if (d.win_stride[q] % 4 != 0) vec = false;
  const size_t smem = (size_t)(row_pitch + IM_ROWS) * sizeof(int32_t);
  if (stride < (1ll << 31) && smem <= 96 * 1024) {
    const int64_t n_tiles = ceil_div(d.rows, IM_ROWS);
    const int64_t cap = (int64_t)state().sm_count * 8;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    if (smem > 48 * 1024) {
      static bool raised = false;
      if (!raised) {
        TCR_CUDA(cudaFuncSetAttribute(im2col_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        TCR_CUDA(cudaFuncSetAttribute(im2col_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        raised = true;
      }
    }
    if (vec) TCR_LAUNCH((im2col_tiled_kernel<true>), grid, 256, smem, (const uint32_t*)image, (uint4*)cols, d);
    else TCR_LAUNCH((im2col_tiled_kernel<false>), grid, 256, smem, (const uint32_t*)image, (uint4*)cols, d);
  } else {
    int grid = wave_grid(d.rows * d.pitch4, 256, 8);
    TCR_LAUNCH(im2col_kernel, grid, 256, 0, (const uint32_t*)image, (uint4*)cols, d);
  }
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_col2im(const void* cols, void* image, const int64_t img_shape[8], const int64_t win_shape[8], int64_t row_pitch, int dtype) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(cols && image && img_shape && win_shape, "tcr_col2im: null argument");
  TCR_ARG(dtype == TCR_FLOAT, "tcr_col2im: FLOAT only (got dtype %d)", dtype);
  Col2imDesc d;
  memset(&d, 0, sizeof(d));
  d.n_img = 1;
  d.pitch = row_pitch;
  int64_t rows = 1, k = 1, n_combo = 1;
  for (int r = 0; r < 8; ++r) {
    TCR_ARG(img_shape[r] >= 1 && win_shape[r] >= 1 && win_shape[r] <= img_shape[r], "tcr_col2im: window %lld does not fit image extent %lld at rank %d",
            (long long)win_shape[r], (long long)img_shape[r], r);
    const int64_t pos = img_shape[r] - win_shape[r] + 1;
    if (img_shape[r] > 1) {
      const int q = d.nd++;
      TCR_ARG(rows < (1ll << 31) && k < (1ll << 31), "tcr_col2im: patch matrix too large for 32-bit coordinates");
      d.ext[q] = (int32_t)img_shape[r]; d.pos[q] = (int32_t)pos; d.win[q] = (int32_t)win_shape[r];
      d.row_stride[q] = (int32_t)rows; d.col_stride[q] = (int32_t)k;
      if (pos > 1 && win_shape[r] > 1) {
        TCR_ARG(d.n_slide < C2I_MAX_SLIDE, "tcr_col2im: more than %d sliding ranks", C2I_MAX_SLIDE);
        d.slide[d.n_slide++] = q;
        n_combo *= win_shape[r];
      }
    }
    rows *= pos;
    k *= win_shape[r];
    d.n_img *= img_shape[r];
  }
  TCR_ARG(row_pitch >= k, "tcr_col2im: row pitch %lld is smaller than the window (%lld elements)", (long long)row_pitch, (long long)k);
  TCR_ARG(d.n_img < (1ll << 31) && rows < (1ll << 31), "tcr_col2im: image or patch matrix too large for 32-bit coordinates");
  TCR_ARG(n_combo <= 2048, "tcr_col2im: %lld window positions per element exceed the shared-memory table", (long long)n_combo);
  d.n_combo = (int32_t)n_combo;
  const size_t smem = (size_t)n_combo * (C2I_MAX_SLIDE + 1) * sizeof(int32_t);
  // 16-byte path: the fastest non-singular rank is window-spanned with unit column stride and a multiple of 4 long
  const bool vec = d.nd > 0 && d.pos[0] == 1 && d.col_stride[0] == 1 && d.ext[0] % 4 == 0 && row_pitch % 4 == 0 && img_shape[0] > 1 &&
                   ((((uintptr_t)cols) | ((uintptr_t)image)) & 15) == 0;
  int grid = wave_grid(vec ? d.n_img / 4 : d.n_img, 256, 8);
  if (vec) TCR_LAUNCH((col2im_kernel<true>), grid, 256, smem, (const float*)cols, (float*)image, d);
  else TCR_LAUNCH((col2im_kernel<false>), grid, 256, smem, (const float*)cols, (float*)image, d);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

}  // extern "C"
