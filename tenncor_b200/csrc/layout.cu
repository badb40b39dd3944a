// layout.cu — EXTEND / PERMUTE / SLICE / PAD / STRIDE / SCATTER / REVERSE / CONCAT.
//
// Replaces internal/eigen/operator.hpp:159-368. These are pure data movement and are
// type-agnostic: kernels are instantiated per element size (1/2/4/8 bytes), results are
// bit-exact. One coordinate-mapped copy kernel covers every op after the host collapses
// the rank-8 description to its effective ranks; transposes of the two fastest ranks
// go through a 32x32 shared-memory tile so that both the read and the write coalesce.
#include <vector>

#include <algorithm>

#include "common.cuh"

namespace tcr {

template <int S> struct Elem;
template <> struct Elem<1> { using type = uint8_t; };
template <> struct Elem<2> { using type = uint16_t; };
template <> struct Elem<4> { using type = uint32_t; };
template <> struct Elem<8> { using type = uint64_t; };

struct MapDim {
  int64_t ext;        // output extent
  int64_t mul, add, div;
  int64_t in_dim;     // input extent of the mapped rank
  int64_t in_stride;  // input stride (elements) of the mapped rank
};

struct MapPlan {
  int nd;
  int check;          // any rank needs a validity test
  int64_t n_out;
  int64_t base;       // constant input offset from collapsed extent-1 ranks
  MapDim d[8];
};

template <typename E, typename I, bool CHECK>
__global__ void __launch_bounds__(256) map_copy_kernel(const E* __restrict__ in, E* __restrict__ out, const __grid_constant__ MapPlan p) {
  TCR_PDL_ENTER();
  const I n = (I)p.n_out;
  const I stride = (I)gridDim.x * blockDim.x;
  for (I o = (I)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
    I t = o;
    int64_t off = p.base;
    bool valid = true;
#pragma unroll 1
    for (int k = 0; k < p.nd; ++k) {
      const MapDim& d = p.d[k];
      I c;
      if (k == p.nd - 1) c = t;
      else { c = t % (I)d.ext; t = t / (I)d.ext; }
      int64_t q = (int64_t)c * d.mul + d.add;
      if (CHECK) {
        if (d.div != 1) {
          if (q % d.div != 0) valid = false;
          q /= d.div;
        }
        if (q < 0 || q >= d.in_dim) valid = false;
      }
      off += q * d.in_stride;
    }
    E v = E();  // zero fill (PAD / SCATTER holes)
    if (!CHECK || valid) v = in[off];
    out[o] = v;
  }
}

// out[j + B*(i + A*c)] = in[i + A*(j + B*c)]  (in is [A, B, C], out is [B, A, C])
template <typename E>
__global__ void __launch_bounds__(256) transpose_kernel(const E* __restrict__ in, E* __restrict__ out, int64_t A, int64_t B) {
  TCR_PDL_ENTER();
  __shared__ E tile[32][33];
  const int64_t c = blockIdx.z;
  const E* src = in + c * A * B;
  E* dst = out + c * A * B;
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int y = threadIdx.y; y < 32; y += 8) {
    int64_t i = i0 + threadIdx.x, j = j0 + y;
    if (i < A && j < B) tile[y][threadIdx.x] = src[i + A * j];
  }
  __syncthreads();
#pragma unroll
  for (int y = threadIdx.y; y < 32; y += 8) {
    int64_t j = j0 + threadIdx.x, i = i0 + y;
    if (i < A && j < B) dst[j + B * i] = tile[threadIdx.x][y];
  }
}

// Vectorised transpose for tile-aligned extents: a block moves a [64 x 16 vectors] tile, every
// global access is a 16-byte vector, each thread keeps four of them in flight (64 KB+ per SM;
// ncu on the 32x32 kernel: 43 % of HBM, long_scoreboard-bound with 16 bytes in flight per thread).
// in is [A (fast), B, C]: tile = 16*V elements along A x 64 along B; out is [B (fast), A, C].
template <typename E>
__global__ void __launch_bounds__(256) transpose_vec_kernel(const E* __restrict__ in, E* __restrict__ out, int64_t A, int64_t B) {
  TCR_PDL_ENTER();
  constexpr int V = 16 / sizeof(E), TA = 16 * V, TB = 64;
  __shared__ E tile[TB][TA + 1];
  struct alignas(16) Vec { E v[V]; };
  const int64_t c = blockIdx.z;
  const E* src = in + c * A * B;
  E* dst = out + c * A * B;
  const int64_t a0 = (int64_t)blockIdx.x * TA, b0 = (int64_t)blockIdx.y * TB;
  {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 vectors along A, 16 rows of B per pass
    Vec x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const Vec*>(src + a0 + tx * V + A * (b0 + ty + 16 * i));
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < V; ++e) tile[ty + 16 * i][tx * V + e] = x[i].v[e];
  }
  __syncthreads();
  {
    constexpr int VPR = TB / V;         // vectors per output row (64 elements along B)
    constexpr int RPP = 256 / VPR;      // output rows (positions along A) per pass
    const int tx = threadIdx.x % VPR, ty = threadIdx.x / VPR;
#pragma unroll
    for (int i = 0; i < TA / RPP; ++i) {
      const int a = ty + RPP * i;
      Vec y;
#pragma unroll
      for (int e = 0; e < V; ++e) y.v[e] = tile[tx * V + e][a];
      *reinterpret_cast<Vec*>(dst + b0 + tx * V + B * (a0 + a)) = y;
    }
  }
}

// concat: copy one argument (shape [inner, ext, outer]) into out at axis offset `off`
// Many strided 2-D copies in one launch (blockIdx.y = item): the [x_t | h_{t-1}] operands of every time step of an unrolled
// recurrent layer (cfg/tenncor/layer.yml:716-813 builds one CONCAT per step) are laid down after the forward pass in one go.
constexpr int COPY2D_MAX = 512;
struct Copy2dItemDev {
  char* dst;
  const char* src;
  uint32_t row_bytes, rows;
  int64_t dst_pitch, src_pitch;
};
struct Copy2dBatch {
  int32_t count;
  Copy2dItemDev it[COPY2D_MAX];
};
template <typename V>
__global__ void __launch_bounds__(256) copy2d_batched_kernel(const __grid_constant__ Copy2dBatch b) {
  TCR_PDL_ENTER();
  const Copy2dItemDev& it = b.it[blockIdx.y];
  const uint32_t per_row = it.row_bytes / (uint32_t)sizeof(V);
  const uint32_t total = per_row * it.rows;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
    const uint32_t r = i / per_row, c = i - r * per_row;
    *reinterpret_cast<V*>(it.dst + (int64_t)r * it.dst_pitch + (int64_t)c * sizeof(V)) =
        *reinterpret_cast<const V*>(it.src + (int64_t)r * it.src_pitch + (int64_t)c * sizeof(V));
  }
}

// SLICE / PAD along ONE rank (the recurrent layers' per-step slices, conv2d's zero border along a fresh rank): the tensor is
// `outer` rows; an output row is [zeros lo | data | zeros hi] of a window of the input row. 16-byte vectors, one division per vector
// (tcr_map_copy spends a div/mod chain over eight ranks per element: 49-52 % of the copy bandwidth on [1024,128,64]).
template <typename I>
__global__ void __launch_bounds__(256) row_window_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t out_row, uint32_t lo, uint32_t data,
                                                          int64_t in_row, int64_t in_off, int64_t total_) {
  TCR_PDL_ENTER();
  // I = uint32_t whenever the vector count allows it (a 64-bit division is ~10x the instructions of a 32-bit one, and a 16 MB
  // slice is only a few vectors per thread); four vectors in flight per thread
  constexpr int U = 4;
  const I total = (I)total_, stride = (I)gridDim.x * 256;
  for (I i0 = (I)blockIdx.x * 256 + threadIdx.x; i0 < total; i0 += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const I i = i0 + (I)u * stride;
      v[u] = make_uint4(0, 0, 0, 0);
      if (i < total) {
        const I r = i / out_row;
        const uint32_t c = (uint32_t)(i - r * out_row);
        if (c >= lo && c - lo < data) v[u] = in[(int64_t)r * in_row + in_off + (c - lo)];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const I i = i0 + (I)u * stride;
      if (i < total) out[i] = v[u];
    }
  }
}

// one rank differs between the two shapes -> (outer rows, bytes per row below and along that rank); false otherwise
static bool single_rank(const int64_t a[8], const int64_t b[8], int* rank) {
  int found = -1;
  for (int k = 0; k < 8; ++k)
    if (a[k] != b[k]) { if (found >= 0) return false; found = k; }
  *rank = found;
  return found >= 0;
}
static bool launch_row_window(const void* in, void* out, const int64_t in_shape[8], int rank, int64_t window_off, int64_t window, int64_t lo, int64_t hi, int elem_size) {
  int64_t inner = elem_size, outer = 1;
  for (int k = 0; k < rank; ++k) inner *= in_shape[k];
  for (int k = rank + 1; k < 8; ++k) outer *= in_shape[k];
  const int64_t in_row = inner * in_shape[rank], off = inner * window_off, data = inner * window, lo_b = inner * lo, out_row = inner * (lo + window + hi);
  if (((in_row | off | data | lo_b | out_row) & 15) != 0 || (((uintptr_t)in | (uintptr_t)out) & 15) != 0) return false;
  if (out_row / 16 >= (1ll << 32) || outer * out_row == 0) return false;
  const int64_t total = outer * out_row / 16;
  const int grid = wave_grid(ceil_div(total, 4), 256, 8);
  // i0 + 3 * stride must not wrap: stride <= 148 * 8 * 256
  if (total < (1ll << 32) - (1ll << 22))
    TCR_LAUNCH(row_window_kernel<uint32_t>, grid, 256, 0, (const uint4*)in, (uint4*)out, (uint32_t)(out_row / 16), (uint32_t)(lo_b / 16), (uint32_t)(data / 16), in_row / 16, off / 16, total);
  else
    TCR_LAUNCH(row_window_kernel<int64_t>, grid, 256, 0, (const uint4*)in, (uint4*)out, (uint32_t)(out_row / 16), (uint32_t)(lo_b / 16), (uint32_t)(data / 16), in_row / 16, off / 16, total);
  return true;
}

struct ConcatArgs {
  const void* ptr[32];
  int n;
};

template <typename E>
__global__ void __launch_bounds__(256) concat_one_kernel(const E* __restrict__ in, E* __restrict__ out, int64_t inner,
                                                         int64_t ext, int64_t outer, int64_t off, int64_t out_ext) {
  TCR_PDL_ENTER();
  const int64_t n = inner * ext * outer, stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t row = inner * ext;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = i % row, o = i / row;
    out[r + inner * off + inner * out_ext * o] = in[i];
  }
}

// binary concat in one launch: rows of the output are [inner*ext0 from a | inner*ext1 from b]
template <typename E>
__global__ void __launch_bounds__(256) concat_pair_kernel(const E* __restrict__ a, const E* __restrict__ b, E* __restrict__ out,
                                                          int64_t row0, int64_t row1, int64_t outer) {
  TCR_PDL_ENTER();
  const int64_t row = row0 + row1, n = row * outer, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i % row, o = i / row;
    out[i] = r < row0 ? a[o * row0 + r] : b[o * row1 + (r - row0)];
  }
}

// n-ary concat: every argument has extent 1 along the axis; args k0..k0+n-1 of `total`
template <typename E>
__global__ void __launch_bounds__(256) concat_many_kernel(const __grid_constant__ ConcatArgs a, E* __restrict__ out,
                                                          int64_t inner, int64_t outer, int k0, int total) {
  TCR_PDL_ENTER();
  const int64_t per = inner * outer, n = per * a.n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t ii = i % inner, t = i / inner;
    int k = (int)(t % a.n);
    int64_t o = t / a.n;
    const E* src = (const E*)a.ptr[k];
    out[ii + inner * ((k0 + k) + (int64_t)total * o)] = src[ii + inner * o];
  }
}

template <typename E>
static int launch_map(const void* in, void* out, const MapPlan& p, int64_t n_in) {
  int grid = wave_grid(p.n_out, 256, 8);
  bool small = p.n_out < (1ll << 31) && n_in < (1ll << 31);
  if (small) {
    if (p.check) TCR_LAUNCH((map_copy_kernel<E, uint32_t, true>), grid, 256, 0, (const E*)in, (E*)out, p);
    else TCR_LAUNCH((map_copy_kernel<E, uint32_t, false>), grid, 256, 0, (const E*)in, (E*)out, p);
  } else {
    if (p.check) TCR_LAUNCH((map_copy_kernel<E, int64_t, true>), grid, 256, 0, (const E*)in, (E*)out, p);
    else TCR_LAUNCH((map_copy_kernel<E, int64_t, false>), grid, 256, 0, (const E*)in, (E*)out, p);
  }
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

#define TCR_DISPATCH_ELEM(es, E, ...)                                           \
  switch (es) {                                                                \
    case 1: { using E = Elem<1>::type; __VA_ARGS__; } break;                   \
    case 2: { using E = Elem<2>::type; __VA_ARGS__; } break;                   \
    case 4: { using E = Elem<4>::type; __VA_ARGS__; } break;                   \
    case 8: { using E = Elem<8>::type; __VA_ARGS__; } break;                   \
    default: set_error("layout op: unsupported element size %d", (int)(es)); return TCR_ERR_ARG; \
  }

static void identity_desc(tcr_map_desc* d, const int64_t in_shape[8], const int64_t out_shape[8]) {
  for (int k = 0; k < 8; ++k) {
    d->in_shape[k] = in_shape[k];
    d->out_shape[k] = out_shape[k];
    d->perm[k] = k;
    d->mul[k] = 1;
    d->add[k] = 0;
    d->div[k] = 1;
  }
}

}  // namespace tcr

using namespace tcr;

extern "C" {

int tcr_map_copy(const void* in, void* out, const tcr_map_desc* desc, int elem_size) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(in && out && desc, "tcr_map_copy: null argument");
  int64_t in_stride[8], n_in = 1, n_out = 1;
  bool seen[8] = {false};
  for (int k = 0; k < 8; ++k) {
    TCR_ARG(desc->in_shape[k] >= 0 && desc->out_shape[k] >= 0, "tcr_map_copy: negative extent");
    TCR_ARG(desc->perm[k] >= 0 && desc->perm[k] < 8 && !seen[desc->perm[k]], "tcr_map_copy: perm is not a permutation");
    TCR_ARG(desc->div[k] >= 1, "tcr_map_copy: div must be >= 1");
    seen[desc->perm[k]] = true;
    in_stride[k] = n_in;
    n_in *= desc->in_shape[k];
    n_out *= desc->out_shape[k];
  }
  if (n_out == 0) return TCR_OK;
  if (n_in == 0) return tcr_memset(out, 0, (size_t)n_out * elem_size);

  // collapse to effective ranks
  MapPlan p;
  memset(&p, 0, sizeof(p));
  p.n_out = n_out;
  bool dead = false;  // a collapsed extent-1 rank maps outside the input -> all zeros
  for (int k = 0; k < 8; ++k) {
    int r = desc->perm[k];
    MapDim d{desc->out_shape[k], desc->mul[k], desc->add[k], desc->div[k], desc->in_shape[r], in_stride[r]};
    if (d.ext == 1) {
      int64_t q = d.add;
      if (q % d.div != 0) { dead = true; continue; }
      q /= d.div;
      if (q < 0 || q >= d.in_dim) { dead = true; continue; }
      p.base += q * d.in_stride;
      continue;
    }
    bool need_check = d.div != 1 || d.add < 0 || ((d.ext - 1) * d.mul + d.add) / d.div >= d.in_dim ||
                      (d.mul < 0 && (d.ext - 1) * d.mul + d.add < 0);
    if (need_check) p.check = 1;
    if (p.nd > 0) {
      MapDim& prev = p.d[p.nd - 1];
      // merge (prev, d) when both walk contiguous full input ranks in order
      bool plain_prev = prev.mul == 1 && prev.add == 0 && prev.div == 1 && prev.in_dim == prev.ext;
      bool plain_d = d.mul == 1 && d.add == 0 && d.div == 1;
      if (plain_prev && plain_d && d.in_stride == prev.in_stride * prev.in_dim && !need_check) {
        prev.ext *= d.ext;
        prev.in_dim *= d.in_dim;
        continue;
      }
      // merge broadcast ranks (mul == 0)
      if (prev.mul == 0 && d.mul == 0 && prev.div == 1 && d.div == 1 && !need_check) {
        p.base += 0;
        int64_t q = d.add;  // constant coordinate
        if (q < 0 || q >= d.in_dim) { dead = true; continue; }
        p.base += q * d.in_stride;
        prev.ext *= d.ext;
        continue;
      }
    }
    p.d[p.nd++] = d;
  }
  if (dead) return tcr_memset(out, 0, (size_t)n_out * elem_size);
  if (p.nd == 0) {  // single element
    return tcr_d2d(out, (const char*)in + p.base * elem_size, (size_t)elem_size);
  }
  // plain contiguous copy
  if (p.nd == 1 && !p.check && p.d[0].mul == 1 && p.d[0].in_stride == 1)
    return tcr_d2d(out, (const char*)in + (p.base + p.d[0].add) * elem_size, (size_t)n_out * elem_size);
  // transpose of the two fastest effective ranks (optionally batched)
  if (!p.check && (p.nd == 2 || p.nd == 3) && p.base == 0) {
    const MapDim& d0 = p.d[0];
    const MapDim& d1 = p.d[1];
    bool t01 = d0.mul == 1 && d0.add == 0 && d1.mul == 1 && d1.add == 0 && d1.in_stride == 1 &&
               d1.in_dim == d1.ext && d0.in_stride == d1.ext && d0.in_dim == d0.ext;
    bool batch_ok = p.nd == 2 || (p.d[2].mul == 1 && p.d[2].add == 0 && p.d[2].in_stride == d0.ext * d1.ext &&
                                  p.d[2].ext <= 65535);
    if (t01 && batch_ok) {
      int64_t A = d1.ext, B = d0.ext, C = p.nd == 3 ? p.d[2].ext : 1;  // in [A,B,C] -> out [B,A,C]
      const int64_t V = 16 / elem_size;
      if (elem_size <= 8 && A % (16 * V) == 0 && B % 64 == 0 && B / 64 <= 65535 && C <= 65535 &&
          (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
        dim3 vgrid((unsigned)(A / (16 * V)), (unsigned)(B / 64), (unsigned)C);
        TCR_DISPATCH_ELEM(elem_size, E, TCR_LAUNCH((transpose_vec_kernel<E>), vgrid, 256, 0, (const E*)in, (E*)out, A, B));
        TCR_CHECK_LAUNCH();
        return TCR_OK;
      }
      dim3 grid((unsigned)ceil_div(A, 32), (unsigned)ceil_div(B, 32), (unsigned)C);
      if (grid.y <= 65535) {
        TCR_DISPATCH_ELEM(elem_size, E, TCR_LAUNCH((transpose_kernel<E>), grid, dim3(32, 8), 0, (const E*)in, (E*)out, A, B));
        TCR_CHECK_LAUNCH();
        return TCR_OK;
      }
    }
  }
  // 16-byte path: when the fastest effective rank is contiguous on both sides and everything is
  // 16-byte aligned, move uint4 "elements" — the coordinate arithmetic is amortised over 16 bytes
  if (elem_size < 16) {
    const int64_t f = 16 / elem_size;
    const MapDim& d0 = p.d[0];
    bool ok = d0.mul == 1 && d0.div == 1 && d0.in_stride == 1 && d0.ext % f == 0 && d0.add % f == 0 && d0.in_dim % f == 0 &&
              p.base % f == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    for (int k = 1; ok && k < p.nd; ++k) ok = p.d[k].in_stride % f == 0;
    if (ok) {
      MapPlan q = p;
      q.d[0].ext /= f; q.d[0].add /= f; q.d[0].in_dim /= f;
      for (int k = 1; k < q.nd; ++k) q.d[k].in_stride /= f;
      q.base /= f;
      q.n_out /= f;
      return launch_map<uint4>(in, out, q, n_in / f);
    }
  }
  TCR_DISPATCH_ELEM(elem_size, E, return launch_map<E>(in, out, p, n_in));
  return TCR_OK;
}

int tcr_extend(const void* in, void* out, const int64_t in_shape[8], const int64_t bcast[8], int elem_size) {
  tcr_map_desc d;
  int64_t out_shape[8];
  for (int k = 0; k < 8; ++k) {
    TCR_ARG(bcast[k] >= 1, "tcr_extend: cannot extend using zero dimensions");
    TCR_ARG(!(bcast[k] > 1 && in_shape[k] > 1), "tcr_extend: cannot extend non-singular dimension %d", k);
    out_shape[k] = in_shape[k] * bcast[k];
  }
  identity_desc(&d, in_shape, out_shape);
  for (int k = 0; k < 8; ++k)
    if (bcast[k] > 1) d.mul[k] = 0;
  return tcr_map_copy(in, out, &d, elem_size);
}

int tcr_permute(const void* in, void* out, const int64_t in_shape[8], const int32_t order[8], int elem_size) {
  tcr_map_desc d;
  int64_t out_shape[8];
  for (int k = 0; k < 8; ++k) {
    TCR_ARG(order[k] >= 0 && order[k] < 8, "tcr_permute: order[%d] out of range", k);
    out_shape[k] = in_shape[order[k]];
  }
  identity_desc(&d, in_shape, out_shape);
  for (int k = 0; k < 8; ++k) d.perm[k] = order[k];
  return tcr_map_copy(in, out, &d, elem_size);
}

int tcr_slice(const void* in, void* out, const int64_t in_shape[8], const int64_t offsets[8], const int64_t extents[8], int elem_size) {
  tcr_map_desc d;
  for (int k = 0; k < 8; ++k)
    TCR_ARG(offsets[k] >= 0 && extents[k] >= 1 && offsets[k] + extents[k] <= in_shape[k], "tcr_slice: box out of range at rank %d", k);
  {
    int rank = -1;
    TCR_REQUIRE_DEVICE();
    if (single_rank(in_shape, extents, &rank) && launch_row_window(in, out, in_shape, rank, offsets[rank], extents[rank], 0, 0, elem_size)) {
      TCR_CHECK_LAUNCH();
      return TCR_OK;
    }
  }
  identity_desc(&d, in_shape, extents);
  for (int k = 0; k < 8; ++k) d.add[k] = offsets[k];
  return tcr_map_copy(in, out, &d, elem_size);
}

int tcr_pad(const void* in, void* out, const int64_t in_shape[8], const int64_t pad_lo[8], const int64_t pad_hi[8], int elem_size) {
  tcr_map_desc d;
  int64_t out_shape[8];
  for (int k = 0; k < 8; ++k) {
    TCR_ARG(pad_lo[k] >= 0 && pad_hi[k] >= 0, "tcr_pad: negative padding");
    out_shape[k] = in_shape[k] + pad_lo[k] + pad_hi[k];
  }
  {
    int rank = -1;
    TCR_REQUIRE_DEVICE();
    if (single_rank(in_shape, out_shape, &rank) && launch_row_window(in, out, in_shape, rank, 0, in_shape[rank], pad_lo[rank], pad_hi[rank], elem_size)) {
      TCR_CHECK_LAUNCH();
      return TCR_OK;
    }
  }
  identity_desc(&d, in_shape, out_shape);
  for (int k = 0; k < 8; ++k) d.add[k] = -pad_lo[k];
  return tcr_map_copy(in, out, &d, elem_size);
}

int tcr_stride(const void* in, void* out, const int64_t in_shape[8], const int64_t incrs[8], int elem_size) {
  tcr_map_desc d;
  int64_t out_shape[8];
  for (int k = 0; k < 8; ++k) {
    TCR_ARG(incrs[k] >= 1, "tcr_stride: increments must be >= 1");
    out_shape[k] = (in_shape[k] + incrs[k] - 1) / incrs[k];  // Eigen TensorStridingOp: ceil(dim / stride)
  }
  identity_desc(&d, in_shape, out_shape);
  for (int k = 0; k < 8; ++k) d.mul[k] = incrs[k];
  return tcr_map_copy(in, out, &d, elem_size);
}

int tcr_scatter(const void* in, void* out, const int64_t in_shape[8], const int64_t out_shape[8], const int64_t incrs[8], int elem_size) {
  tcr_map_desc d;
  identity_desc(&d, in_shape, out_shape);
  for (int k = 0; k < 8; ++k) {
    TCR_ARG(incrs[k] >= 1, "tcr_scatter: increments must be >= 1");
    d.div[k] = incrs[k];
  }
  return tcr_map_copy(in, out, &d, elem_size);
}

int tcr_reverse(const void* in, void* out, const int64_t shape[8], uint32_t reverse_mask, int elem_size) {
  tcr_map_desc d;
  identity_desc(&d, shape, shape);
  for (int k = 0; k < 8; ++k)
    if ((reverse_mask >> k) & 1u) { d.mul[k] = -1; d.add[k] = shape[k] - 1; }
  return tcr_map_copy(in, out, &d, elem_size);
}

int tcr_copy2d_batched(const tcr_copy2d_item* items, int count) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(items != nullptr && count >= 0, "tcr_copy2d_batched: bad arguments");
  for (int k0 = 0; k0 < count; k0 += COPY2D_MAX) {
    static Copy2dBatch b;  // 16 KB: not on the stack
    b.count = count - k0 < COPY2D_MAX ? count - k0 : COPY2D_MAX;
    bool vec = true;
    int64_t most = 0;
    for (int k = 0; k < b.count; ++k) {
      const tcr_copy2d_item& it = items[k0 + k];
      TCR_ARG(it.dst != nullptr && it.src != nullptr && it.row_bytes >= 0 && it.rows >= 0 && it.row_bytes < (1ll << 31) && it.rows < (1ll << 31),
              "tcr_copy2d_batched: item %d is malformed", k0 + k);
      b.it[k].dst = (char*)it.dst; b.it[k].src = (const char*)it.src;
      b.it[k].row_bytes = (uint32_t)it.row_bytes; b.it[k].rows = (uint32_t)it.rows;
      b.it[k].dst_pitch = it.dst_pitch; b.it[k].src_pitch = it.src_pitch;
      vec = vec && ((((uintptr_t)it.dst | (uintptr_t)it.src | (uintptr_t)it.dst_pitch | (uintptr_t)it.src_pitch | (uintptr_t)it.row_bytes) & 15) == 0);
      most = std::max<int64_t>(most, it.row_bytes * it.rows);
    }
    if (most == 0) continue;
    const int unit = vec ? 16 : 1;
    int per_item = (int)ceil_div(most / unit, (int64_t)(256 * 8));
    if (per_item < 1) per_item = 1;
    if (per_item > 64) per_item = 64;
    dim3 grid((unsigned)per_item, (unsigned)b.count);
    if (vec) TCR_LAUNCH((copy2d_batched_kernel<uint4>), grid, 256, 0, b);
    else TCR_LAUNCH((copy2d_batched_kernel<uint8_t>), grid, 256, 0, b);
  }
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_concat(const void* const* args, const int64_t* shapes, int nargs, void* out, int axis, int elem_size) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(args && shapes && out, "tcr_concat: null argument");
  TCR_ARG(nargs >= 2, "tcr_concat: needs at least two arguments");
  TCR_ARG(axis >= 0 && axis < 8, "tcr_concat: axis %d out of range", axis);
  const int64_t* s0 = shapes;
  int64_t inner = 1, outer = 1, out_ext = 0;
  for (int k = 0; k < axis; ++k) inner *= s0[k];
  for (int k = axis + 1; k < 8; ++k) outer *= s0[k];
  for (int a = 0; a < nargs; ++a) {
    const int64_t* s = shapes + 8 * a;
    for (int k = 0; k < 8; ++k)
      TCR_ARG(k == axis || s[k] == s0[k], "tcr_concat: argument %d shape mismatch at rank %d", a, k);
    if (nargs > 2) TCR_ARG(s[axis] == 1, "tcr_concat: cannot group concat shapes with dimension that is not one");
    out_ext += s[axis];
  }
  if (inner * outer * out_ext == 0) return TCR_OK;
  if (nargs > 2) {
    for (int k0 = 0; k0 < nargs; k0 += 32) {
      ConcatArgs ca;
      ca.n = nargs - k0 < 32 ? nargs - k0 : 32;
      for (int k = 0; k < ca.n; ++k) ca.ptr[k] = args[k0 + k];
      int grid = wave_grid(inner * outer * ca.n, 256, 8);
      TCR_DISPATCH_ELEM(elem_size, E, TCR_LAUNCH((concat_many_kernel<E>), grid, 256, 0, ca, (E*)out, inner, outer, k0, nargs));
    }
    TCR_CHECK_LAUNCH();
    return TCR_OK;
  }
  if (nargs == 2 && outer > 1) {  // the recurrent cells' [x_t | h] input: one launch instead of one per argument
    const int64_t row0 = inner * shapes[axis], row1 = inner * shapes[8 + axis];
    const int64_t f = 16 / elem_size;
    if (elem_size < 16 && row0 % f == 0 && row1 % f == 0 && (((uintptr_t)args[0] | (uintptr_t)args[1] | (uintptr_t)out) & 15) == 0) {
      int grid = wave_grid((row0 + row1) / f * outer, 256, 8);
      TCR_LAUNCH((concat_pair_kernel<uint4>), grid, 256, 0, (const uint4*)args[0], (const uint4*)args[1], (uint4*)out, row0 / f, row1 / f, outer);
    } else {
      int grid = wave_grid((row0 + row1) * outer, 256, 8);
      TCR_DISPATCH_ELEM(elem_size, E, TCR_LAUNCH((concat_pair_kernel<E>), grid, 256, 0, (const E*)args[0], (const E*)args[1], (E*)out, row0, row1, outer));
    }
    TCR_CHECK_LAUNCH();
    return TCR_OK;
  }
  int64_t off = 0;
  for (int a = 0; a < nargs; ++a) {
    int64_t ext = shapes[8 * a + axis];
    if (outer == 1) {  // contiguous block
      int rc = tcr_d2d((char*)out + inner * off * elem_size, args[a], (size_t)(inner * ext) * elem_size);
      if (rc) return rc;
    } else {
      const int64_t f = 16 / elem_size;
      if (elem_size < 16 && inner % f == 0 && (((uintptr_t)args[a] | (uintptr_t)out) & 15) == 0) {
        int grid = wave_grid(inner / f * ext * outer, 256, 8);
        TCR_LAUNCH((concat_one_kernel<uint4>), grid, 256, 0, (const uint4*)args[a], (uint4*)out, inner / f, ext, outer, off, out_ext);
      } else {
        int grid = wave_grid(inner * ext * outer, 256, 8);
        TCR_DISPATCH_ELEM(elem_size, E, TCR_LAUNCH((concat_one_kernel<E>), grid, 256, 0, (const E*)args[a], (E*)out, inner, ext, outer, off, out_ext));
      }
    }
    off += ext;
  }
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

}  // extern "C"
