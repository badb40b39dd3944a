// elementwise.cu — fused, vectorised, coalesced elementwise kernels (HBM-bound).
//
// Replaces the cwise factories of the reference (internal/eigen/operator.hpp:377-987,
// select :1050-1067, cast :1239-1260, assign* :1190-1237, rand_uniform :993-1044).
// Two kernel families:
//   * ew_direct_kernel<T, F>: one op, 16-byte vector loads/stores, 2 vectors in flight
//     per thread, grid sized to whole waves of the SM count;
//   * ew_vm_kernel<T>: a per-thread register machine that executes a fused chain
//     (tcr_ew_program) over 4 elements at a time; operands may be broadcasts
//     (scalar / leading block / trailing block) so EXTEND never materialises.
#include <math.h>

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace tcr {

// ------------------------------------------------------------------ scalar math
template <typename T> struct Compute { using type = T; };            // float, double
template <> struct Compute<int32_t> { using type = double; };         // transcendental on ints go through double
template <> struct Compute<int64_t> { using type = double; };

template <typename T> __device__ __forceinline__ T m_exp(T x);
template <> __device__ __forceinline__ float m_exp(float x) { return expf(x); }
template <> __device__ __forceinline__ double m_exp(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T m_log(T x);
template <> __device__ __forceinline__ float m_log(float x) { return logf(x); }
template <> __device__ __forceinline__ double m_log(double x) { return log(x); }
template <typename T> __device__ __forceinline__ T m_sin(T x);
template <> __device__ __forceinline__ float m_sin(float x) { return sinf(x); }
template <> __device__ __forceinline__ double m_sin(double x) { return sin(x); }
template <typename T> __device__ __forceinline__ T m_cos(T x);
template <> __device__ __forceinline__ float m_cos(float x) { return cosf(x); }
template <> __device__ __forceinline__ double m_cos(double x) { return cos(x); }
template <typename T> __device__ __forceinline__ T m_tan(T x);
template <> __device__ __forceinline__ float m_tan(float x) { return tanf(x); }
template <> __device__ __forceinline__ double m_tan(double x) { return tan(x); }
template <typename T> __device__ __forceinline__ T m_sqrt(T x);
template <> __device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T m_round(T x);
template <> __device__ __forceinline__ float m_round(float x) { return roundf(x); }
template <> __device__ __forceinline__ double m_round(double x) { return round(x); }
template <typename T> __device__ __forceinline__ T m_tanh(T x);
template <> __device__ __forceinline__ float m_tanh(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double m_tanh(double x) { return tanh(x); }
template <typename T> __device__ __forceinline__ T m_pow(T x, T y);
template <> __device__ __forceinline__ float m_pow(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double m_pow(double x, double y) { return pow(x, y); }
template <typename T> __device__ __forceinline__ T m_abs(T x) { return x < T(0) ? -x : x; }
template <> __device__ __forceinline__ float m_abs(float x) { return fabsf(x); }
template <> __device__ __forceinline__ double m_abs(double x) { return fabs(x); }
// Eigen scalar_sigmoid_op: 1 / (1 + exp(-x))
template <typename T> __device__ __forceinline__ T m_sigmoid(T x) { return T(1) / (T(1) + m_exp<T>(-x)); }
// fp32: the denominator is >= 1, so the approximate reciprocal (<= 2 ulp, no slow path) is safe;
// 1 / inf = 0 is the correct limit for x -> -inf
template <> __device__ __forceinline__ float m_sigmoid(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }

// Heavy libm bodies (Payne-Hanek range reduction, pow) are kept out of line in the multi-opcode
// kernels: inlined once per element per opcode case they made those kernels > 150 KB of SASS and
// the instruction cache, not HBM, set their speed. The single-op direct kernels inline them.
template <typename T> __device__ __noinline__ T m_sin_ool(T x) { return m_sin<T>(x); }
template <typename T> __device__ __noinline__ T m_cos_ool(T x) { return m_cos<T>(x); }
template <typename T> __device__ __noinline__ T m_tan_ool(T x) { return m_tan<T>(x); }
template <typename T> __device__ __noinline__ T m_pow_ool(T x, T y) { return m_pow<T>(x, y); }

template <typename T, bool IsFloat = (sizeof(typename Compute<T>::type) == sizeof(T) && !std::is_integral<T>::value)>
struct Ops;

template <typename T>
struct Ops<T, true> {  // float / double
  static __device__ __forceinline__ T un(int op, T a) {
    switch (op) {
      case TCR_EW_ABS: return m_abs<T>(a);
      case TCR_EW_NEG: return -a;
      case TCR_EW_SIN: return m_sin<T>(a);
      case TCR_EW_COS: return m_cos<T>(a);
      case TCR_EW_TAN: return m_tan<T>(a);
      case TCR_EW_EXP: return m_exp<T>(a);
      case TCR_EW_LOG: return m_log<T>(a);
      case TCR_EW_SQRT: return m_sqrt<T>(a);
      case TCR_EW_ROUND: return m_round<T>(a);
      case TCR_EW_SIGMOID: return m_sigmoid<T>(a);
      case TCR_EW_TANH: return m_tanh<T>(a);
      case TCR_EW_SQUARE: return a * a;
      case TCR_EW_CUBE: return a * a * a;
      default: return a;
    }
  }
  static __device__ __forceinline__ T bin(int op, T a, T b) {
    switch (op) {
      case TCR_EW_POW: return m_pow<T>(a, b);
      case TCR_EW_ADD: return a + b;
      case TCR_EW_SUB: return a - b;
      case TCR_EW_MUL: return a * b;
      case TCR_EW_DIV: return a / b;
      case TCR_EW_MIN: return b < a ? b : a;  // std::min / Eigen cwiseMin
      case TCR_EW_MAX: return a < b ? b : a;
      case TCR_EW_EQ: return T(a == b);
      case TCR_EW_NEQ: return T(a != b);
      case TCR_EW_LT: return T(a < b);
      case TCR_EW_GT: return T(a > b);
      default: return a;
    }
  }
};

template <typename T>
struct Ops<T, false> {  // integers: transcendental ops evaluate in double and truncate
  using C = double;
  static __device__ __forceinline__ T un(int op, T a) {
    switch (op) {
      case TCR_EW_ABS: return a < T(0) ? T(-a) : a;
      case TCR_EW_NEG: return T(-a);
      case TCR_EW_SIN: return T(sin((C)a));
      case TCR_EW_COS: return T(cos((C)a));
      case TCR_EW_TAN: return T(tan((C)a));
      case TCR_EW_EXP: return T(exp((C)a));
      case TCR_EW_LOG: return T(log((C)a));
      case TCR_EW_SQRT: return T(sqrt((C)a));
      case TCR_EW_ROUND: return a;
      case TCR_EW_SIGMOID: return T(1.0 / (1.0 + exp(-(C)a)));
      case TCR_EW_TANH: return T(tanh((C)a));
      case TCR_EW_SQUARE: return T(a * a);
      case TCR_EW_CUBE: return T(a * a * a);
      default: return a;
    }
  }
  static __device__ __forceinline__ T bin(int op, T a, T b) {
    switch (op) {
      case TCR_EW_POW: return T(pow((C)a, (C)b));
      case TCR_EW_ADD: return T(a + b);
      case TCR_EW_SUB: return T(a - b);
      case TCR_EW_MUL: return T(a * b);
      case TCR_EW_DIV: return b == T(0) ? T(0) : T(a / b);
      case TCR_EW_MIN: return b < a ? b : a;
      case TCR_EW_MAX: return a < b ? b : a;
      case TCR_EW_EQ: return T(a == b);
      case TCR_EW_NEQ: return T(a != b);
      case TCR_EW_LT: return T(a < b);
      case TCR_EW_GT: return T(a > b);
      default: return a;
    }
  }
};

// ------------------------------------------------------------------ vector access
template <typename T> struct alignas(16) Vec {  // 16-byte vector of T
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

template <typename T>
__device__ __forceinline__ Vec<T> ld16(const T* p) {
  Vec<T> r;
  *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);  // plain load: operands may alias the output (ASSIGN_*)
  return r;
}
template <typename T>
__device__ __forceinline__ void st16(T* p, const Vec<T>& r) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(r.v);
}

// ------------------------------------------------------------------ direct kernels
// NIN inputs of the same size n, one output; OP is a compile-time opcode.
// mode: 0 = unary/binary cwise, 1 = in-place assign flavour handled by caller through pointers.
template <typename T, int OP, int NIN>
__global__ void __launch_bounds__(256) ew_direct_kernel(const T* a, const T* b, const T* c, T* out, int64_t n) {
  TCR_PDL_ENTER();
  constexpr int N = Vec<T>::N;
  constexpr int UNROLL = 2;
  const int64_t nvec = n / N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride * (UNROLL - 1) < nvec; i += stride * UNROLL) {
    Vec<T> va[UNROLL], vb[UNROLL], vc[UNROLL], vo[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      va[u] = ld16(a + (i + u * stride) * N);
      if (NIN > 1) vb[u] = ld16(b + (i + u * stride) * N);
      if (NIN > 2) vc[u] = ld16(c + (i + u * stride) * N);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (NIN == 1) vo[u].v[k] = Ops<T>::un(OP, va[u].v[k]);
        else if (NIN == 2) vo[u].v[k] = Ops<T>::bin(OP, va[u].v[k], vb[u].v[k]);
        else vo[u].v[k] = (va[u].v[k] != T(0)) ? vb[u].v[k] : vc[u].v[k];
      }
      st16(out + (i + u * stride) * N, vo[u]);
    }
  }
  for (; i < nvec; i += stride) {
    Vec<T> va = ld16(a + i * N), vb, vc, vo;
    if (NIN > 1) vb = ld16(b + i * N);
    if (NIN > 2) vc = ld16(c + i * N);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (NIN == 1) vo.v[k] = Ops<T>::un(OP, va.v[k]);
      else if (NIN == 2) vo.v[k] = Ops<T>::bin(OP, va.v[k], vb.v[k]);
      else vo.v[k] = (va.v[k] != T(0)) ? vb.v[k] : vc.v[k];
    }
    st16(out + i * N, vo);
  }
  // scalar tail
  int64_t t = nvec * N + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    if (NIN == 1) out[t] = Ops<T>::un(OP, a[t]);
    else if (NIN == 2) out[t] = Ops<T>::bin(OP, a[t], b[t]);
    else out[t] = (a[t] != T(0)) ? b[t] : c[t];
  }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <typename T, int OP, int NIN>
static int launch_direct(const void* a, const void* b, const void* c, void* out, int64_t n) {
  int grid = wave_grid(ceil_div(n, Vec<T>::N * 2), 256, 8);
  TCR_LAUNCH((ew_direct_kernel<T, OP, NIN>), grid, 256, 0, (const T*)a, (const T*)b, (const T*)c, (T*)out, n);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

#define UN_CASE(OP) case OP: return launch_direct<T, OP, 1>(a, nullptr, nullptr, out, n);
#define BIN_CASE(OP) case OP: return launch_direct<T, OP, 2>(a, b, nullptr, out, n);

template <typename T>
static int direct_unary(int op, const void* a, void* out, int64_t n) {
  switch (op) {
    UN_CASE(TCR_EW_ABS) UN_CASE(TCR_EW_NEG) UN_CASE(TCR_EW_SIN) UN_CASE(TCR_EW_COS) UN_CASE(TCR_EW_TAN)
    UN_CASE(TCR_EW_EXP) UN_CASE(TCR_EW_LOG) UN_CASE(TCR_EW_SQRT) UN_CASE(TCR_EW_ROUND) UN_CASE(TCR_EW_SIGMOID)
    UN_CASE(TCR_EW_TANH) UN_CASE(TCR_EW_SQUARE) UN_CASE(TCR_EW_CUBE)
    default: set_error("tcr_unary: opcode %d is not unary", op); return TCR_ERR_ARG;
  }
}
template <typename T>
static int direct_binary(int op, const void* a, const void* b, void* out, int64_t n) {
  switch (op) {
    BIN_CASE(TCR_EW_POW) BIN_CASE(TCR_EW_ADD) BIN_CASE(TCR_EW_SUB) BIN_CASE(TCR_EW_MUL) BIN_CASE(TCR_EW_DIV)
    BIN_CASE(TCR_EW_MIN) BIN_CASE(TCR_EW_MAX) BIN_CASE(TCR_EW_EQ) BIN_CASE(TCR_EW_NEQ) BIN_CASE(TCR_EW_LT)
    BIN_CASE(TCR_EW_GT)
    default: set_error("tcr_binary: opcode %d is not binary", op); return TCR_ERR_ARG;
  }
}

// opcode known at compile time inside a `case`: heavy bodies go out of line (see m_*_ool)
template <typename T, int OP> __device__ __forceinline__ T vm_un(T a) {
  using C = typename Compute<T>::type;
  if (OP == TCR_EW_SIN) return T(m_sin_ool<C>((C)a));
  if (OP == TCR_EW_COS) return T(m_cos_ool<C>((C)a));
  if (OP == TCR_EW_TAN) return T(m_tan_ool<C>((C)a));
  return Ops<T>::un(OP, a);
}
template <typename T, int OP> __device__ __forceinline__ T vm_bin(T a, T b) {
  using C = typename Compute<T>::type;
  if (OP == TCR_EW_POW) return T(m_pow_ool<C>((C)a, (C)b));
  return Ops<T>::bin(OP, a, b);
}

// ------------------------------------------------------------------ register machine
struct VmInput {
  const void* ptr;
  int32_t dtype;
  int32_t mode;  // 0 full, 1 scalar, 2 general 3-segment broadcast
  uint8_t bcast[3];
};
struct VmParams {
  int32_t n_inputs, n_outputs, n_instrs;
  int64_t n;
  int64_t d0, d1;  // segment extents (d2 implied)
  VmInput in[TCR_EW_MAX_INPUTS];
  tcr_ew_output out[TCR_EW_MAX_OUTPUTS];
  tcr_ew_instr ins[TCR_EW_MAX_INSTRS];
  // map-reduce form (tcr_elementwise_reduce): output 0 is not stored but summed over all elements
  void* red_out;       // one element of the compute type
  void* red_partial;   // gridDim.x elements
  int* red_counter;    // ticket, zero between launches (self-resetting)
  int32_t red_post;    // 0 none, 1 divide by red_imm, 2 multiply by red_imm
  double red_imm;
};

template <typename T>
__device__ __forceinline__ T load_any(const void* p, int dtype, int64_t j) {
  switch (dtype) {
    case TCR_FLOAT: return (T)((const float*)p)[j];
    case TCR_DOUBLE: return (T)((const double*)p)[j];
    case TCR_INT8: return (T)((const int8_t*)p)[j];
    case TCR_UINT8: return (T)((const uint8_t*)p)[j];
    case TCR_INT16: return (T)((const int16_t*)p)[j];
    case TCR_UINT16: return (T)((const uint16_t*)p)[j];
    case TCR_INT32: return (T)((const int32_t*)p)[j];
    case TCR_UINT32: return (T)((const uint32_t*)p)[j];
    case TCR_INT64: return (T)((const int64_t*)p)[j];
    case TCR_UINT64: return (T)((const uint64_t*)p)[j];
    default: return T(0);
  }
}
template <typename T>
__device__ __forceinline__ void store_any(void* p, int dtype, int64_t j, T v) {
  switch (dtype) {
    case TCR_FLOAT: ((float*)p)[j] = (float)v; break;
    case TCR_DOUBLE: ((double*)p)[j] = (double)v; break;
    case TCR_INT8: ((int8_t*)p)[j] = (int8_t)v; break;
    case TCR_UINT8: ((uint8_t*)p)[j] = (uint8_t)v; break;
    case TCR_INT16: ((int16_t*)p)[j] = (int16_t)v; break;
    case TCR_UINT16: ((uint16_t*)p)[j] = (uint16_t)v; break;
    case TCR_INT32: ((int32_t*)p)[j] = (int32_t)v; break;
    case TCR_UINT32: ((uint32_t*)p)[j] = (uint32_t)v; break;
    case TCR_INT64: ((int64_t*)p)[j] = (int64_t)v; break;
    case TCR_UINT64: ((uint64_t*)p)[j] = (uint64_t)v; break;
    default: break;
  }
}

constexpr int VM_V = 4;  // elements per thread per iteration

template <typename T> struct alignas(16) V4 { T v[VM_V]; };
template <typename T> struct VmCfg { static constexpr int THREADS = sizeof(T) == 4 ? 256 : 128; };

// The virtual registers live in shared memory, one 4-element slot per (register, thread):
// operand fetch is an LDS.128 with a computed address instead of a branch tree over
// hardware registers, which keeps the kernel at ~40 registers (full occupancy) and makes
// the cost of a VM instruction two shared loads, one uniform opcode branch and one store.
template <typename T> __device__ __forceinline__ T shfl_down_any(T v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
template <> __device__ __forceinline__ int64_t shfl_down_any(int64_t v, int o) { return (int64_t)__shfl_down_sync(0xffffffffu, (long long)v, o); }

// The register machine in three pieces: one input of one chunk (VM_V consecutive elements) from memory, the instruction list over
// the shared-memory register file (register r of this thread at regs[r * TH + threadIdx.x]), the outputs of one chunk to memory.
template <typename T, bool ALIGNED, typename I>
__device__ __forceinline__ V4<T> vm_load(const VmParams& p, const VmInput& in, const I ch) {
  const I n = (I)p.n;
  const I d0 = (I)p.d0, d1 = (I)p.d1;
  const I base = ch * VM_V;
  const bool full = base + VM_V <= n;
  V4<T> x;
  if (in.mode == 1) {
    T s = load_any<T>(in.ptr, in.dtype, 0);
#pragma unroll
    for (int v = 0; v < VM_V; ++v) x.v[v] = s;
  } else if (in.mode == 0) {
    if (ALIGNED && full && in.dtype == DTypeOf<T>::value) {
      const T* src = (const T*)in.ptr + base;
      if (sizeof(T) == 4) {
        *reinterpret_cast<uint4*>(x.v) = *reinterpret_cast<const uint4*>(src);
      } else {
        *reinterpret_cast<uint4*>(x.v) = *reinterpret_cast<const uint4*>(src);
        *reinterpret_cast<uint4*>(x.v + 2) = *reinterpret_cast<const uint4*>(src + 2);
      }
    } else if (ALIGNED && full && in.dtype == TCR_UINT8 && (reinterpret_cast<uintptr_t>(in.ptr) & 3) == 0) {
      // byte inputs (image pixels cast on the device): one 32-bit load per chunk
      const uint32_t w = *reinterpret_cast<const uint32_t*>((const uint8_t*)in.ptr + base);
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = (T)((w >> (8 * v)) & 0xffu);
    } else {
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = (base + v < n) ? load_any<T>(in.ptr, in.dtype, base + v) : T(0);
    }
  } else if (in.mode == 3) {
    // 3-segment broadcast with D0 % 4 == 0: the chunk stays inside one run of segment 0
    const I e0 = in.bcast[0] ? 1 : d0, e1 = in.bcast[1] ? 1 : d1;
    I i0 = base % d0, t = base / d0, i1 = t % d1, i2 = t / d1;
    I j = (in.bcast[0] ? 0 : i0) + e0 * ((in.bcast[1] ? 0 : i1) + e1 * (in.bcast[2] ? 0 : i2));
    if (in.bcast[0]) {
      T s = load_any<T>(in.ptr, in.dtype, j);
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = s;
    } else if (ALIGNED && in.dtype == DTypeOf<T>::value) {
      const T* src = (const T*)in.ptr + j;
      *reinterpret_cast<uint4*>(x.v) = *reinterpret_cast<const uint4*>(src);
      if (sizeof(T) == 8) *reinterpret_cast<uint4*>(x.v + 2) = *reinterpret_cast<const uint4*>(src + 2);
    } else {
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = load_any<T>(in.ptr, in.dtype, j + v);
    }
  } else {
    const I e0 = in.bcast[0] ? 1 : d0, e1 = in.bcast[1] ? 1 : d1;
#pragma unroll
    for (int v = 0; v < VM_V; ++v) {
      I i = base + v;
      if (i >= n) { x.v[v] = T(0); continue; }
      I i0 = i % d0, t = i / d0, i1 = t % d1, i2 = t / d1;
      I j = (in.bcast[0] ? 0 : i0) + e0 * ((in.bcast[1] ? 0 : i1) + e1 * (in.bcast[2] ? 0 : i2));
      x.v[v] = load_any<T>(in.ptr, in.dtype, j);
    }
  }
  return x;
}

template <typename T, int TH>
__device__ __forceinline__ void vm_exec(const VmParams& p, V4<T>* regs) {
  for (int pc = 0; pc < p.n_instrs; ++pc) {
    const tcr_ew_instr& ins = p.ins[pc];
    const int op = ins.op;
    V4<T> a = regs[ins.a * TH + threadIdx.x], d;
    if (op >= TCR_EW_POW && op <= TCR_EW_GT) {
      V4<T> b = regs[ins.b * TH + threadIdx.x];
      switch (op) {
#define VMB(OP) case OP: _Pragma("unroll") for (int v = 0; v < VM_V; ++v) d.v[v] = vm_bin<T, OP>(a.v[v], b.v[v]); break;
        VMB(TCR_EW_ADD) VMB(TCR_EW_SUB) VMB(TCR_EW_MUL) VMB(TCR_EW_DIV) VMB(TCR_EW_POW) VMB(TCR_EW_MIN)
        VMB(TCR_EW_MAX) VMB(TCR_EW_EQ) VMB(TCR_EW_NEQ) VMB(TCR_EW_LT) VMB(TCR_EW_GT)
#undef VMB
        default: d = a;
      }
    } else if (op == TCR_EW_CONST) {
#pragma unroll
      for (int v = 0; v < VM_V; ++v) d.v[v] = (T)ins.imm;
    } else if (op == TCR_EW_SELECT) {
      V4<T> b = regs[ins.b * TH + threadIdx.x], c = regs[ins.c * TH + threadIdx.x];
#pragma unroll
      for (int v = 0; v < VM_V; ++v) d.v[v] = (a.v[v] != T(0)) ? b.v[v] : c.v[v];
    } else {
      switch (op) {
#define VMU(OP) case OP: _Pragma("unroll") for (int v = 0; v < VM_V; ++v) d.v[v] = vm_un<T, OP>(a.v[v]); break;
        VMU(TCR_EW_SIGMOID) VMU(TCR_EW_TANH) VMU(TCR_EW_EXP) VMU(TCR_EW_NEG) VMU(TCR_EW_SQUARE) VMU(TCR_EW_LOG)
        VMU(TCR_EW_SQRT) VMU(TCR_EW_ABS) VMU(TCR_EW_SIN) VMU(TCR_EW_COS) VMU(TCR_EW_TAN) VMU(TCR_EW_ROUND)
        VMU(TCR_EW_CUBE)
#undef VMU
        default: d = a;  // MOV
      }
    }
    regs[ins.dst * TH + threadIdx.x] = d;
  }
}

template <typename T, bool ALIGNED, typename I, bool RED, int TH>
__device__ __forceinline__ void vm_store(const VmParams& p, V4<T>* regs, const I ch, T& red_acc) {
  const I n = (I)p.n;
  const I base = ch * VM_V;
  const bool full = base + VM_V <= n;
  if (RED) {
    const V4<T> y = regs[p.out[0].reg * TH + threadIdx.x];
#pragma unroll
    for (int v = 0; v < VM_V; ++v)
      if (base + v < n) red_acc += y.v[v];
  }
  for (int k = RED ? 1 : 0; k < p.n_outputs; ++k) {
    const tcr_ew_output& o = p.out[k];
    V4<T> y = regs[o.reg * TH + threadIdx.x];
    if (ALIGNED && full && o.dtype == DTypeOf<T>::value) {
      T* dst = (T*)o.ptr + base;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(y.v);
      if (sizeof(T) == 8) *reinterpret_cast<uint4*>(dst + 2) = *reinterpret_cast<const uint4*>(y.v + 2);
    } else {
#pragma unroll
      for (int v = 0; v < VM_V; ++v)
        if (base + v < n) store_any<T>(o.ptr, o.dtype, base + v, y.v[v]);
    }
  }
}

// one chunk of one program: load, execute, store
template <typename T, bool ALIGNED, typename I, bool RED>
__device__ __forceinline__ void vm_chunk(const VmParams& p, V4<T>* regs, const I ch, T& red_acc) {
  constexpr int TH = VmCfg<T>::THREADS;
  for (int k = 0; k < p.n_inputs; ++k) regs[k * TH + threadIdx.x] = vm_load<T, ALIGNED, I>(p, p.in[k], ch);
  vm_exec<T, TH>(p, regs);
  vm_store<T, ALIGNED, I, RED, TH>(p, regs, ch, red_acc);
}

template <typename T, bool ALIGNED, typename I, bool RED = false>
__global__ void __launch_bounds__(VmCfg<T>::THREADS) ew_vm_kernel(const __grid_constant__ VmParams p) {
  TCR_PDL_ENTER();
  constexpr int THREADS = VmCfg<T>::THREADS;
  __shared__ V4<T> regs[TCR_EW_NREGS * THREADS];
  T red_acc = T(0);  // RED: this thread's share of the sum of output 0
  const I nchunks = ((I)p.n + VM_V - 1) / VM_V;
  const I stride = (I)gridDim.x * THREADS;
  for (I ch = (I)blockIdx.x * THREADS + threadIdx.x; ch < nchunks; ch += stride) vm_chunk<T, ALIGNED, I, RED>(p, regs, ch, red_acc);
  if (RED) {
    // block sum in a fixed order (shuffle tree, then warp order), one partial per block; the last block to arrive adds the
    // partials in block order and applies the post-op: deterministic for a given grid, one launch
    __shared__ T warp_sum[THREADS / 32];
    __shared__ int last_block;
    T v = red_acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_down_any<T>(v, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      T b = T(0);
      for (int w = 0; w < THREADS / 32; ++w) b += warp_sum[w];
      ((T*)p.red_partial)[blockIdx.x] = b;
      __threadfence();
      const int ticket = atomicAdd(p.red_counter, 1);
      last_block = ticket == (int)gridDim.x - 1;
      if (last_block) *p.red_counter = 0;
    }
    __syncthreads();
    if (last_block && threadIdx.x < 32) {
      __threadfence();
      T t = T(0);
      for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += __ldcg((const T*)p.red_partial + b);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += shfl_down_any<T>(t, o);
      if (threadIdx.x == 0) {
        if (p.red_post == 1) t = t / (T)p.red_imm;
        else if (p.red_post == 2) t = t * (T)p.red_imm;
        *(T*)p.red_out = t;
      }
    }
  }
}

// Several programs over the SAME iteration space in one launch: a thread runs program 0, 1, ... on its chunk before moving to
// the next chunk. A later program may read what an earlier one stored at the same index (the thread's own stores are visible
// to it), which is how a chain of elementwise results that each have several readers — the LSTM cell's backward: dh, dc and
// four gate pre-activation gradients (tenncor/eteq/backprop.hpp:136-142 over cfg/tenncor/layer.yml:716-768) — becomes one launch
// without a register machine wide enough to hold all of them at once.
constexpr int VM_MULTI_MAX = 8;
constexpr int VM_MULTI_THREADS = 128;
struct VmMulti {
  int32_t count;
  int32_t n_ext;                                          // inputs read from memory, over all programs
  // where input j of program k comes from: >= 0 slot of the prefetched external inputs; < 0: output 0 of program -(src + 1)
  int16_t src[VM_MULTI_MAX][TCR_EW_MAX_INPUTS];
  uint8_t store[VM_MULTI_MAX];                            // 0: nobody outside the launch reads this program's result
  VmParams p[VM_MULTI_MAX];
};

// Dynamic shared memory: [n_ext + count][THREADS] chunks. Every input that comes from memory is fetched first, for all programs
// at once (one memory round trip for the launch instead of one per program); a program's result stays in its forward slot for
// the programs after it and goes to memory only if somebody outside the launch reads it.
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(VM_MULTI_THREADS) ew_vm_multi_kernel(const __grid_constant__ VmMulti pm) {
  TCR_PDL_ENTER();
  constexpr int TH = VM_MULTI_THREADS;
  __shared__ V4<T> regs[TCR_EW_NREGS * TH];
  extern __shared__ __align__(16) uint8_t vm_multi_dyn[];
  V4<T>* ext = reinterpret_cast<V4<T>*>(vm_multi_dyn);
  V4<T>* fwd = ext + pm.n_ext * TH;
  T unused = T(0);
  const uint32_t nchunks = ((uint32_t)pm.p[0].n + VM_V - 1) / VM_V;
  const uint32_t stride = gridDim.x * TH;
  for (uint32_t ch = blockIdx.x * TH + threadIdx.x; ch < nchunks; ch += stride) {
    const bool full = (ch + 1) * VM_V <= (uint32_t)pm.p[0].n;
    for (int k = 0; k < pm.count; ++k)
      for (int j = 0; j < pm.p[k].n_inputs; ++j) {
        const int sl = pm.src[k][j];
        if (sl < 0) continue;
        const VmInput& in = pm.p[k].in[j];
        if (ALIGNED && full && in.mode == 0 && in.dtype == DTypeOf<T>::value) {
          // whole, aligned, already of the compute type: global -> shared without a register in between, so all of these are in
          // flight together (a load followed by its own shared store would serialise the launch on one L2 latency per input)
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ext[sl * TH + threadIdx.x]);
          const T* src = (const T*)in.ptr + (size_t)ch * VM_V;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          if (sizeof(T) == 8) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 2) : "memory");
        } else {
          ext[sl * TH + threadIdx.x] = vm_load<T, ALIGNED, uint32_t>(pm.p[k], in, ch);
        }
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    for (int k = 0; k < pm.count; ++k) {
      const VmParams& p = pm.p[k];
      for (int j = 0; j < p.n_inputs; ++j) {
        const int sl = pm.src[k][j];
        regs[j * TH + threadIdx.x] = sl >= 0 ? ext[sl * TH + threadIdx.x] : fwd[(-sl - 1) * TH + threadIdx.x];
      }
      vm_exec<T, TH>(p, regs);
      fwd[k * TH + threadIdx.x] = regs[p.out[0].reg * TH + threadIdx.x];
      if (pm.store[k]) vm_store<T, ALIGNED, uint32_t, false, TH>(p, regs, ch, unused);
    }
  }
}

// ------------------------------------------------------------------ byte pixels -> float (CAST [* constant])
// The input of an image model arrives as UINT8 and is cast (and scaled by 1/255) on the device (tenncor/eteq/caster.hpp:10-44 +
// a MUL by a constant): 16 pixels per thread, one 16-byte load and four 16-byte stores, instead of four byte loads per thread.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n, float scale, int scaled) {
  TCR_PDL_ENTER();
  const int64_t nvec = n / 16, stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t word[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 f;
      f.x = (float)(word[q] & 0xffu); f.y = (float)((word[q] >> 8) & 0xffu); f.z = (float)((word[q] >> 16) & 0xffu); f.w = (float)(word[q] >> 24);
      if (scaled) { f.x = __fmul_rn(f.x, scale); f.y = __fmul_rn(f.y, scale); f.z = __fmul_rn(f.z, scale); f.w = __fmul_rn(f.w, scale); }
      reinterpret_cast<float4*>(out)[i * 4 + q] = f;
    }
  }
  for (int64_t i = nvec * 16 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    const float f = (float)in[i];
    out[i] = scaled ? __fmul_rn(f, scale) : f;
  }
}

// program == (float)u8 [* c] ?  (registers traced symbolically: 0 unknown, 1 the input, 2 a constant, 3 input * constant)
static bool match_pixel_cast(const tcr_ew_program* prog, float* scale, int* scaled) {
  if (prog->dtype != TCR_FLOAT || prog->n_inputs != 1 || prog->n_outputs != 1 || prog->n_instrs > 4) return false;
  const tcr_ew_input& in = prog->inputs[0];
  if (in.dtype != TCR_UINT8 || prog->outputs[0].dtype != TCR_FLOAT) return false;
  if ((in.bcast[0] && prog->dims[0] > 1) || (in.bcast[1] && prog->dims[1] > 1) || (in.bcast[2] && prog->dims[2] > 1)) return false;
  int kind[TCR_EW_NREGS] = {0};
  float cst[TCR_EW_NREGS] = {0};
  kind[0] = 1;
  for (int i = 0; i < prog->n_instrs; ++i) {
    const tcr_ew_instr& ins = prog->instrs[i];
    if (ins.dst >= TCR_EW_NREGS || ins.a >= TCR_EW_NREGS || ins.b >= TCR_EW_NREGS) return false;
    if (ins.op == TCR_EW_CONST) { kind[ins.dst] = 2; cst[ins.dst] = (float)ins.imm; }
    else if (ins.op == TCR_EW_MOV) { const int k = kind[ins.a]; const float c = cst[ins.a]; kind[ins.dst] = k; cst[ins.dst] = c; }
    else if (ins.op == TCR_EW_MUL && ((kind[ins.a] == 1 && kind[ins.b] == 2) || (kind[ins.a] == 2 && kind[ins.b] == 1))) {
      const float c = kind[ins.a] == 2 ? cst[ins.a] : cst[ins.b];
      kind[ins.dst] = 3; cst[ins.dst] = c;
    } else return false;
  }
  const int r = prog->outputs[0].reg;
  if (r >= TCR_EW_NREGS || (kind[r] != 1 && kind[r] != 3)) return false;
  *scaled = kind[r] == 3;
  *scale = cst[r];
  return aligned16(in.ptr) && aligned16(prog->outputs[0].ptr);
}

// ------------------------------------------------------------------ gated-cell backward (tcr_cell_backward)
struct CellBwdParams {
  uint32_t n;
  int32_t n_gates;
  const float *s_a, *s_b, *c_x, *c_y, *c_z;
  float *s_out, *c_out;
  int32_t kind[6], sel[6];
  const float *x[6], *y[6];
  float* out[6];
};

template <int V>
struct CellVec { float v[V]; };
template <int V> __device__ __forceinline__ CellVec<V> cell_ld(const float* p, uint32_t i) {
  CellVec<V> r;
  if (V == 4) *reinterpret_cast<float4*>(r.v) = *reinterpret_cast<const float4*>(p + i);
  else r.v[0] = p[i];
  return r;
}
template <int V> __device__ __forceinline__ void cell_st(float* p, uint32_t i, const CellVec<V>& r) {
  if (V == 4) *reinterpret_cast<float4*>(p + i) = *reinterpret_cast<const float4*>(r.v);
  else p[i] = r.v[0];
}

// every product / sum is a separately rounded operation (__fmul_rn / __fadd_rn keep ptxas from contracting them into FMAs):
// the results are bit-identical to the chain of single-functor launches
template <int V>
__global__ void __launch_bounds__(128) cell_backward_kernel(const __grid_constant__ CellBwdParams p) {
  TCR_PDL_ENTER();
  const uint32_t i = (blockIdx.x * 128u + threadIdx.x) * V;
  if (i >= p.n) return;
  // all loads first: one memory round trip for the launch
  const CellVec<V> sa = cell_ld<V>(p.s_a, i), sb = cell_ld<V>(p.s_b, i), cx = cell_ld<V>(p.c_x, i), cy = cell_ld<V>(p.c_y, i), cz = cell_ld<V>(p.c_z, i);
  CellVec<V> gx[6], gy[6];
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (k < p.n_gates) { gx[k] = cell_ld<V>(p.x[k], i); gy[k] = cell_ld<V>(p.y[k], i); }
  CellVec<V> s, c;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    s.v[v] = __fadd_rn(sa.v[v], sb.v[v]);
    c.v[v] = __fadd_rn(__fmul_rn(cx.v[v], cy.v[v]), __fmul_rn(cz.v[v], s.v[v]));
  }
  if (p.s_out != nullptr) cell_st<V>(p.s_out, i, s);
  if (p.c_out != nullptr) cell_st<V>(p.c_out, i, c);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    if (k >= p.n_gates) break;
    CellVec<V> o;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float x = gx[k].v[v];
      const float local = p.kind[k] == 1 ? __fmul_rn(x, __fsub_rn(1.0f, x)) : __fsub_rn(1.0f, __fmul_rn(x, x));
      o.v[v] = __fmul_rn(local, __fmul_rn(gy[k].v[v], p.sel[k] ? c.v[v] : s.v[v]));
    }
    cell_st<V>(p.out[k], i, o);
  }
}

static int op_arity(int op) {
  if (op >= TCR_EW_ABS && op <= TCR_EW_CUBE) return 1;
  if (op >= TCR_EW_POW && op <= TCR_EW_GT) return 2;
  if (op == TCR_EW_SELECT) return 3;
  if (op == TCR_EW_MOV) return 1;
  if (op == TCR_EW_CONST) return 0;
  return -1;
}

int* counter_ring_take(int n);  // runtime.cu

struct VmReduce {
  void* out = nullptr;
  int post = 0;
  double imm = 0;
};

// tcr_ew_program -> kernel parameters; `aligned` is cleared when a vector access would be misaligned
static int fill_vm(const tcr_ew_program* prog, VmParams& p, bool& aligned, const VmReduce* red) {
  memset(&p, 0, sizeof(p));
  p.n_inputs = prog->n_inputs;
  p.n_outputs = prog->n_outputs;
  p.n_instrs = prog->n_instrs;
  p.d0 = prog->dims[0];
  p.d1 = prog->dims[1];
  p.n = prog->dims[0] * prog->dims[1] * prog->dims[2];
  for (int k = 0; k < prog->n_inputs; ++k) {
    const tcr_ew_input& in = prog->inputs[k];
    TCR_ARG(in.ptr != nullptr, "tcr_elementwise: input %d is null", k);
    TCR_ARG(dtype_size(in.dtype) != 0, "tcr_elementwise: input %d has bad dtype %d", k, in.dtype);
    p.in[k].ptr = in.ptr;
    p.in[k].dtype = in.dtype;
    // a broadcast flag on an extent-1 segment is a no-op
    bool b0 = in.bcast[0] && prog->dims[0] > 1, b1 = in.bcast[1] && prog->dims[1] > 1,
         b2 = in.bcast[2] && prog->dims[2] > 1;
    bool all = (b0 || prog->dims[0] == 1) && (b1 || prog->dims[1] == 1) && (b2 || prog->dims[2] == 1);
    p.in[k].bcast[0] = b0; p.in[k].bcast[1] = b1; p.in[k].bcast[2] = b2;
    p.in[k].mode = (!b0 && !b1 && !b2) ? 0 : (all ? 1 : 2);
    if (p.in[k].mode == 2 && prog->dims[0] % VM_V == 0) p.in[k].mode = 3;
    if ((p.in[k].mode == 0 || (p.in[k].mode == 3 && !b0)) && !aligned16(in.ptr)) aligned = false;
  }
  for (int k = 0; k < prog->n_outputs; ++k) {
    if (red != nullptr && k == 0) {  // summed, not stored
      TCR_ARG(prog->outputs[0].reg < TCR_EW_NREGS, "tcr_elementwise_reduce: output register out of range");
      p.out[0] = prog->outputs[0];
      continue;
    }
    TCR_ARG(prog->outputs[k].ptr != nullptr, "tcr_elementwise: output %d is null", k);
    TCR_ARG(prog->outputs[k].reg < TCR_EW_NREGS, "tcr_elementwise: output %d register out of range", k);
    TCR_ARG(dtype_size(prog->outputs[k].dtype) != 0, "tcr_elementwise: output %d has bad dtype", k);
    p.out[k] = prog->outputs[k];
    if (!aligned16(p.out[k].ptr)) aligned = false;
  }
  for (int k = 0; k < prog->n_instrs; ++k) {
    const tcr_ew_instr& ins = prog->instrs[k];
    TCR_ARG(op_arity(ins.op) >= 0, "tcr_elementwise: instr %d has bad opcode %d", k, (int)ins.op);
    TCR_ARG(ins.dst < TCR_EW_NREGS && ins.a < TCR_EW_NREGS && ins.b < TCR_EW_NREGS && ins.c < TCR_EW_NREGS,
            "tcr_elementwise: instr %d register out of range", k);
    p.ins[k] = ins;
  }
  return TCR_OK;
}

template <typename T>
static int run_vm(const tcr_ew_program* prog, const VmReduce* red = nullptr) {
  VmParams p;
  bool aligned = true;
  int frc = fill_vm(prog, p, aligned, red);
  if (frc) return frc;
  if (p.n == 0 && red == nullptr) return TCR_OK;
  constexpr int THREADS = VmCfg<T>::THREADS;
  int grid = wave_grid(ceil_div(p.n, VM_V), THREADS, sizeof(T) == 4 ? 6 : 6);
  const bool small = p.n < (1ll << 31);
  if (red != nullptr) {
    void* partial = nullptr;
    int rc = tcr_alloc(&partial, sizeof(T) * (size_t)grid);
    if (rc) return rc;
    p.red_out = red->out;
    p.red_partial = partial;
    p.red_counter = counter_ring_take(1);
    TCR_ARG(p.red_counter != nullptr, "tcr_elementwise_reduce: no ticket counters (tcr_init not called?)");
    p.red_post = red->post;
    p.red_imm = red->imm;
    if (small) {
      if (aligned) TCR_LAUNCH((ew_vm_kernel<T, true, uint32_t, true>), grid, THREADS, 0, p);
      else TCR_LAUNCH((ew_vm_kernel<T, false, uint32_t, true>), grid, THREADS, 0, p);
    } else {
      if (aligned) TCR_LAUNCH((ew_vm_kernel<T, true, int64_t, true>), grid, THREADS, 0, p);
      else TCR_LAUNCH((ew_vm_kernel<T, false, int64_t, true>), grid, THREADS, 0, p);
    }
    tcr_free(partial);
    TCR_CHECK_LAUNCH();
    return TCR_OK;
  }
  if (small) {
    if (aligned) TCR_LAUNCH((ew_vm_kernel<T, true, uint32_t>), grid, THREADS, 0, p);
    else TCR_LAUNCH((ew_vm_kernel<T, false, uint32_t>), grid, THREADS, 0, p);
  } else {
    if (aligned) TCR_LAUNCH((ew_vm_kernel<T, true, int64_t>), grid, THREADS, 0, p);
    else TCR_LAUNCH((ew_vm_kernel<T, false, int64_t>), grid, THREADS, 0, p);
  }
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

// ------------------------------------------------------------------ chain kernel
// ncu showed the register machine above bound by instruction issue (72 % of issue slots, ~410
// instructions per 4 elements, 85 % of them decode / address arithmetic / branches, see
// profiles/r1_ncu_ew_vm.md). Most fused regions the planner emits are expression trees that a
// two-deep stack evaluates (Strahler number <= 2: activations, their gradients, optimiser updates,
// LSTM cell updates). Those run here: the program is re-coded on the host as <= 8 straight-line
// steps over an accumulator `acc` and one temporary `tmp`; a step may consume one leaf.
//   * intermediate values never leave hardware registers (no shared-memory register file);
//   * leaves are staged global -> shared with cp.async into thread-private 16-byte slots, double
//     buffered: the loads of iteration i+1 are in flight while iteration i computes, so bytes in
//     flight per SM (~100 KB at 1024 threads x 96 B) are not capped by the register file — a first
//     version that staged leaves in registers ran at 50 % of HBM with 16 warps per SM;
//   * one non-unrolled step loop = one copy of the opcode switch (an unrolled loop was 360 KB of SASS).
// step forms: every binary step is acc = op(acc, leaf); a swapped operand order is folded into the
// opcode (TCR_CH_REV bit); PUSH parks acc in a thread-private shared slot that a later POP reads back
// as its leaf, so there is no second live register array and no operand select
enum { CH_NONE = 0, CH_INIT, CH_PUSH, CH_UN, CH_BIN, CH_POP };
constexpr int TCR_CH_REV = 0x80;
enum { SL_NONE = 0, SL_FULL, SL_CONSTANT, SL_CHUNK, SL_GENERAL };                  // leaf kinds
constexpr int CHAIN_MAXN = 8;

struct ChainParams {
  int32_t n_steps, n_staged, n_buf;  // n_buf staging buffers: loads run n_buf - 1 iterations ahead
  int32_t has_chunk;                 // some leaf is a partial broadcast: track the split index
  int64_t n, d0, d1;
  uint8_t form[CHAIN_MAXN], op[CHAIN_MAXN], kind[CHAIN_MAXN];
  uint8_t stage[CHAIN_MAXN];  // staging slot of the leaf of step i (kinds FULL / CHUNK)
  uint8_t bcast[CHAIN_MAXN][3];
  int32_t dtype[CHAIN_MAXN];
  const void* ptr[CHAIN_MAXN];
  double imm[CHAIN_MAXN];
  void* out;
  int32_t out_dtype;
};

template <typename T, int CH>
__device__ __forceinline__ void chain_apply(int op, bool unary, V4<T> (&acc)[CH], const V4<T> (&r)[CH]) {
  if (unary) {
    switch (op) {
#define CHU(OP) case OP: _Pragma("unroll") for (int c = 0; c < CH; ++c) _Pragma("unroll") for (int v = 0; v < VM_V; ++v) acc[c].v[v] = vm_un<T, OP>(acc[c].v[v]); break;
      CHU(TCR_EW_SIGMOID) CHU(TCR_EW_TANH) CHU(TCR_EW_EXP) CHU(TCR_EW_NEG) CHU(TCR_EW_SQUARE) CHU(TCR_EW_LOG)
      CHU(TCR_EW_SQRT) CHU(TCR_EW_ABS) CHU(TCR_EW_SIN) CHU(TCR_EW_COS) CHU(TCR_EW_TAN) CHU(TCR_EW_ROUND) CHU(TCR_EW_CUBE)
#undef CHU
      default: break;
    }
  } else {
    switch (op) {
#define CHB(OP) case OP: _Pragma("unroll") for (int c = 0; c < CH; ++c) _Pragma("unroll") for (int v = 0; v < VM_V; ++v) acc[c].v[v] = vm_bin<T, OP>(acc[c].v[v], r[c].v[v]); break;
#define CHR(OP) case (OP | TCR_CH_REV): _Pragma("unroll") for (int c = 0; c < CH; ++c) _Pragma("unroll") for (int v = 0; v < VM_V; ++v) acc[c].v[v] = vm_bin<T, OP>(r[c].v[v], acc[c].v[v]); break;
      CHB(TCR_EW_ADD) CHB(TCR_EW_SUB) CHB(TCR_EW_MUL) CHB(TCR_EW_DIV) CHB(TCR_EW_POW) CHB(TCR_EW_MIN)
      CHB(TCR_EW_MAX) CHB(TCR_EW_EQ) CHB(TCR_EW_NEQ) CHB(TCR_EW_LT) CHB(TCR_EW_GT)
      CHR(TCR_EW_SUB) CHR(TCR_EW_DIV) CHR(TCR_EW_POW) CHR(TCR_EW_LT) CHR(TCR_EW_GT)
#undef CHB
#undef CHR
      default: break;
    }
  }
}

// rare paths (mixed-type or unaligned leaves, ragged tail), out of line to keep the hot loop small
template <typename T>
__device__ __noinline__ V4<T> chain_load_general(const void* ptr, int dtype, bool b0, bool b1, bool b2, uint32_t base, uint32_t n,
                                                 uint32_t d0, uint32_t d1) {
  V4<T> x;
  if (!b0 && !b1 && !b2 && base + VM_V <= n) {
    // un-broadcast leaf of another element type (a CAST folded into its consumer): no index arithmetic; byte pixels in one load
    if (dtype == TCR_UINT8 && (reinterpret_cast<uintptr_t>(ptr) & 3) == 0) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>((const uint8_t*)ptr + base);
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = (T)((w >> (8 * v)) & 0xffu);
      return x;
    }
#pragma unroll
    for (int v = 0; v < VM_V; ++v) x.v[v] = load_any<T>(ptr, dtype, base + v);
    return x;
  }
  const uint32_t e0 = b0 ? 1 : d0, e1 = b1 ? 1 : d1;
#pragma unroll 1
  for (int v = 0; v < VM_V; ++v) {
    const uint32_t e = base + v;
    T val = T(0);
    if (e < n) {
      const uint32_t i0 = e % d0, t = e / d0, i1 = t % d1, i2 = t / d1;
      val = load_any<T>(ptr, dtype, (b0 ? 0 : i0) + e0 * ((b1 ? 0 : i1) + e1 * (b2 ? 0 : i2)));
    }
    x.v[v] = val;
  }
  return x;
}
template <typename T>
__device__ __noinline__ void chain_store_general(void* out, int dtype, uint32_t base, uint32_t n, V4<T> y) {
#pragma unroll 1
  for (int v = 0; v < VM_V; ++v)
    if (base + v < n) store_any<T>(out, dtype, base + v, y.v[v]);
}

template <typename T, int CH>
__global__ void __launch_bounds__(256, 4) ew_chain_kernel(const __grid_constant__ ChainParams p) {
  TCR_PDL_ENTER();
  constexpr int THREADS = 256;
  using I = uint32_t;
  extern __shared__ __align__(16) unsigned char chain_smem[];
  // stage[buf][slot][chunk][thread]
  V4<T>* const mine = reinterpret_cast<V4<T>*>(chain_smem) + threadIdx.x;
  const int n_staged = p.n_staged;
  const int buf_stride = n_staged * CH * THREADS;
  const I n = (I)p.n, nchunks = (n + VM_V - 1) / VM_V;
  const I stride = (I)gridDim.x * (THREADS * CH);
  const I d0 = (I)p.d0, d1 = (I)p.d1;

  // (i0, i1, i2) of each chunk at the prefetch position, advanced by mixed-radix addition every
  // iteration: broadcast leaves need the split index, and three integer divisions per chunk per
  // iteration were 40 % of the instructions of a bias + sigmoid kernel
  I pi0[CH], pi1[CH], pi2[CH];
  I step0 = 0, step1 = 0, step2 = 0;  // the iteration stride, split the same way
#pragma unroll
  for (int c = 0; c < CH; ++c) pi0[c] = pi1[c] = pi2[c] = 0;
  const bool has_chunk = p.has_chunk != 0;
  if (has_chunk) {
    const I se = stride * VM_V;
    step0 = se % d0;
    const I t = se / d0;
    step1 = t % d1;
    step2 = t / d1;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const I base = ((I)blockIdx.x * (THREADS * CH) + threadIdx.x + (I)c * THREADS) * VM_V;
      pi0[c] = base % d0;
      const I u = base / d0;
      pi1[c] = u % d1;
      pi2[c] = u / d1;
    }
  }
  // issue the asynchronous copies of one iteration into buffer `buf` (called for consecutive iterations)
  auto prefetch = [&](I ch0, int buf) {
    if (ch0 < nchunks) {
#pragma unroll 1
      for (int i = 0; i < p.n_steps; ++i) {
        const int kind = p.kind[i];
        if (kind != SL_FULL && kind != SL_CHUNK) continue;
        V4<T>* slot = mine + buf * buf_stride + p.stage[i] * (CH * THREADS);
        const bool b0 = p.bcast[i][0], b1 = p.bcast[i][1], b2 = p.bcast[i][2];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const I base = (ch0 + (I)c * THREADS) * VM_V;
          if (base + VM_V > n) continue;  // ragged tail: loaded at use
          I j = base;
          if (kind == SL_CHUNK) {
            const I e0 = b0 ? 1 : d0, e1 = b1 ? 1 : d1;
            j = (b0 ? 0 : pi0[c]) + e0 * ((b1 ? 0 : pi1[c]) + e1 * (b2 ? 0 : pi2[c]));
          }
          const T* src = (const T*)p.ptr[i] + j;
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot + c * THREADS);
          if (kind == SL_CHUNK && b0) {
            if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
          } else if (kind == SL_CHUNK) {  // re-read by many threads: keep it in L1
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 2) : "memory");
          } else {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            if (sizeof(T) == 8) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 2) : "memory");
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (has_chunk)
#pragma unroll
    for (int c = 0; c < CH; ++c) {  // advance to the next iteration's chunk
      pi0[c] += step0;
      I carry = 0;
      if (pi0[c] >= d0) { pi0[c] -= d0; carry = 1; }
      pi1[c] += step1 + carry;
      carry = 0;
      if (pi1[c] >= d1) { pi1[c] -= d1; carry = 1; }
      pi2[c] += step2 + carry;
    }
  };

  I ch0 = (I)blockIdx.x * (THREADS * CH) + threadIdx.x;
  const int n_buf = p.n_buf;
  for (int k = 0; k < n_buf - 1; ++k) prefetch(ch0 + (I)k * stride, k);
  int buf = 0, ahead = n_buf - 1;  // buffer of the current iteration / buffer the next prefetch fills
  for (; ch0 < nchunks; ch0 += stride) {
    prefetch(ch0 + (I)(n_buf - 1) * stride, ahead);
    // everything but the newest n_buf - 1 groups has landed
    if (n_buf == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else if (n_buf == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else asm volatile("cp.async.wait_group 3;" ::: "memory");
    V4<T> acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int v = 0; v < VM_V; ++v) acc[c].v[v] = T(0);
    V4<T>* const park = mine + n_buf * buf_stride;  // PUSH / POP slot (one pending value: depth-2 trees)
#pragma unroll 1
    for (int i = 0; i < p.n_steps; ++i) {
      const int form = p.form[i], kind = p.kind[i];
      V4<T> leaf[CH];
      if (form == CH_POP) {
#pragma unroll
        for (int c = 0; c < CH; ++c) leaf[c] = park[c * THREADS];
      } else if (kind == SL_FULL || kind == SL_CHUNK) {
        const V4<T>* slot = mine + buf * buf_stride + p.stage[i] * (CH * THREADS);
        const bool splat = kind == SL_CHUNK && p.bcast[i][0];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const I base = (ch0 + (I)c * THREADS) * VM_V;
          if (base + VM_V <= n) {
            leaf[c] = slot[c * THREADS];
            if (splat) {
#pragma unroll
              for (int v = 1; v < VM_V; ++v) leaf[c].v[v] = leaf[c].v[0];
            }
          } else {
            leaf[c] = chain_load_general<T>(p.ptr[i], p.dtype[i], p.bcast[i][0], p.bcast[i][1], p.bcast[i][2], base, n, d0, d1);
          }
        }
      } else if (kind == SL_CONSTANT) {
        const T s = p.ptr[i] ? load_any<T>(p.ptr[i], p.dtype[i], 0) : (T)p.imm[i];
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int v = 0; v < VM_V; ++v) leaf[c].v[v] = s;
      } else if (kind == SL_GENERAL) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
          leaf[c] = chain_load_general<T>(p.ptr[i], p.dtype[i], p.bcast[i][0], p.bcast[i][1], p.bcast[i][2],
                                          (ch0 + (I)c * THREADS) * VM_V, n, d0, d1);
      }
      if (form == CH_INIT) {
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = leaf[c];
      } else if (form == CH_PUSH) {
#pragma unroll
        for (int c = 0; c < CH; ++c) { park[c * THREADS] = acc[c]; acc[c] = leaf[c]; }
      } else {
        chain_apply<T, CH>(p.op[i], form == CH_UN, acc, leaf);
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const I base = (ch0 + (I)c * THREADS) * VM_V;
      if (base >= n) continue;
      if (base + VM_V <= n && p.out_dtype == DTypeOf<T>::value) {
        T* dst = (T*)p.out + base;
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(acc[c].v);
        if (sizeof(T) == 8) *reinterpret_cast<uint4*>(dst + 2) = *reinterpret_cast<const uint4*>(acc[c].v + 2);
      } else {
        chain_store_general<T>(p.out, p.out_dtype, base, n, acc[c]);
      }
    }
    buf = buf + 1 == n_buf ? 0 : buf + 1;
    ahead = ahead + 1 == n_buf ? 0 : ahead + 1;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// Host: re-code a tcr_ew_program as a chain. Returns false when the program is not a depth-2 tree.
namespace {
struct ChainExpr {
  int kind;  // 0 input, 1 imm, 2 op
  int op = 0, a = -1, b = -1, input = -1, uses = 0;
  double imm = 0;
};
struct ChainCode {
  int n = 0;
  uint8_t form[8], op[8];
  int leaf[8];  // expr id of the leaf consumed by step i, or -1
};
struct ChainGen {
  std::vector<ChainExpr> ex;
  ChainCode code;
  bool emit(int form, int op, int leaf) {
    if (code.n >= 8) return false;
    code.form[code.n] = (uint8_t)form;
    code.op[code.n] = (uint8_t)op;
    code.leaf[code.n] = leaf;
    ++code.n;
    return true;
  }
  bool is_leaf(int e) const { return ex[e].kind != 2; }
  static bool commutes(int op) {
    return op == TCR_EW_ADD || op == TCR_EW_MUL || op == TCR_EW_MIN || op == TCR_EW_MAX || op == TCR_EW_EQ || op == TCR_EW_NEQ;
  }
  // depth 0: acc free; depth 1: acc holds a live value that PUSH parks until the matching POP
  bool gen(int e, int depth) {
    const ChainExpr& x = ex[e];
    if (x.kind != 2) return emit(depth == 0 ? CH_INIT : CH_PUSH, 0, e);
    if (x.uses > 1) return false;  // shared interior value: needs a named register
    if (x.b < 0) return gen(x.a, depth) && emit(CH_UN, x.op, -1);
    if (is_leaf(x.b)) return gen(x.a, depth) && emit(CH_BIN, x.op, x.b);                                  // acc op leaf
    if (is_leaf(x.a)) return gen(x.b, depth) && emit(CH_BIN, commutes(x.op) ? x.op : (x.op | TCR_CH_REV), x.a);  // leaf op acc
    if (depth != 0) return false;
    // acc = a; PUSH: park a, acc = b; POP: acc = parked(a) op acc(b) -> reversed
    return gen(x.a, 0) && gen(x.b, 1) && emit(CH_POP, commutes(x.op) ? x.op : (x.op | TCR_CH_REV), -1);
  }
};
}  // namespace

template <typename T, int CH>
static int launch_chain(const ChainParams& p) {
  bool parks = false;
  for (int i = 0; i < p.n_steps; ++i) parks |= p.form[i] == CH_PUSH;
  size_t slots = (size_t)p.n_buf * p.n_staged + (parks ? 1 : 0);  // staging ring (+ the PUSH/POP slot)
  if (slots == 0) slots = 1;
  const size_t smem = slots * CH * 256 * sizeof(V4<T>);
  static size_t configured = 0;
  if (smem > configured) {
    TCR_CUDA(cudaFuncSetAttribute(ew_chain_kernel<T, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  int per_sm = (int)(220 * 1024 / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int grid = wave_grid(ceil_div(p.n, VM_V), 256 * CH, per_sm);
  TCR_LAUNCH((ew_chain_kernel<T, CH>), grid, 256, smem, p);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

template <typename T>
static bool try_chain(const tcr_ew_program* prog, int* rc) {
  static const int enabled = std::getenv("TCR_EW_CHAIN") ? std::atoi(std::getenv("TCR_EW_CHAIN")) : 1;
  if (!enabled || prog->n_outputs != 1 || prog->n_instrs < 1) return false;
  const int64_t n = prog->dims[0] * prog->dims[1] * prog->dims[2];
  if (n < (1 << 15) || n >= (1ll << 31) - (1 << 22)) return false;  // small tensors: the register machine has the shorter prologue
  if (!aligned16(prog->outputs[0].ptr)) return false;
  // registers -> expressions
  ChainGen g;
  int reg[TCR_EW_NREGS];
  for (int r = 0; r < TCR_EW_NREGS; ++r) reg[r] = -1;
  for (int k = 0; k < prog->n_inputs; ++k) {
    ChainExpr e;
    e.kind = 0;
    e.input = k;
    g.ex.push_back(e);
    reg[k] = (int)g.ex.size() - 1;
  }
  for (int i = 0; i < prog->n_instrs; ++i) {
    const tcr_ew_instr& ins = prog->instrs[i];
    const int ar = op_arity(ins.op);
    ChainExpr e;
    if (ins.op == TCR_EW_CONST) {
      e.kind = 1;
      e.imm = ins.imm;
    } else if (ins.op == TCR_EW_MOV) {
      if (reg[ins.a] < 0) return false;
      reg[ins.dst] = reg[ins.a];
      continue;
    } else if (ar == 1 || ar == 2) {
      e.kind = 2;
      e.op = ins.op;
      e.a = reg[ins.a];
      e.b = ar == 2 ? reg[ins.b] : -1;
      if (e.a < 0 || (ar == 2 && e.b < 0)) return false;
    } else {
      return false;  // SELECT
    }
    g.ex.push_back(e);
    reg[ins.dst] = (int)g.ex.size() - 1;
  }
  const int root = reg[prog->outputs[0].reg];
  if (root < 0 || g.ex[root].kind != 2) return false;
  // use counts over the tree reachable from the root
  std::vector<int> stack{root};
  g.ex[root].uses = 1;
  while (!stack.empty()) {
    const int e = stack.back();
    stack.pop_back();
    for (int child : {g.ex[e].a, g.ex[e].b}) {
      if (child < 0) continue;
      if (++g.ex[child].uses == 1 && g.ex[child].kind == 2) stack.push_back(child);
    }
  }
  if (!g.gen(root, 0)) return false;
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.n_steps = g.code.n;
  p.n = n;
  p.d0 = prog->dims[0];
  p.d1 = prog->dims[1];
  p.out = prog->outputs[0].ptr;
  p.out_dtype = prog->outputs[0].dtype;
  int staged_of_input[TCR_EW_MAX_INPUTS];
  for (int k = 0; k < TCR_EW_MAX_INPUTS; ++k) staged_of_input[k] = -1;
  for (int i = 0; i < g.code.n; ++i) {
    p.form[i] = g.code.form[i];
    p.op[i] = g.code.op[i];
    p.kind[i] = SL_NONE;
    const int leaf = g.code.leaf[i];
    if (leaf < 0) continue;
    const ChainExpr& e = g.ex[leaf];
    if (e.kind == 1) {
      p.kind[i] = SL_CONSTANT;
      p.imm[i] = e.imm;
      continue;
    }
    const tcr_ew_input& in = prog->inputs[e.input];
    if (in.ptr == nullptr || dtype_size(in.dtype) == 0) return false;
    const bool b0 = in.bcast[0] && prog->dims[0] > 1, b1 = in.bcast[1] && prog->dims[1] > 1, b2 = in.bcast[2] && prog->dims[2] > 1;
    const bool all = (b0 || prog->dims[0] == 1) && (b1 || prog->dims[1] == 1) && (b2 || prog->dims[2] == 1);
    p.ptr[i] = in.ptr;
    p.dtype[i] = in.dtype;
    p.bcast[i][0] = b0; p.bcast[i][1] = b1; p.bcast[i][2] = b2;
    const bool same = in.dtype == DTypeOf<T>::value;
    if (all) p.kind[i] = SL_CONSTANT;
    else if (!b0 && !b1 && !b2) p.kind[i] = (same && aligned16(in.ptr)) ? SL_FULL : SL_GENERAL;
    else if (same && prog->dims[0] % VM_V == 0 && (b0 || aligned16(in.ptr))) p.kind[i] = SL_CHUNK;
    else p.kind[i] = SL_GENERAL;
    if (p.kind[i] == SL_FULL || p.kind[i] == SL_CHUNK) {
      // an input read by several steps is staged once
      if (staged_of_input[e.input] < 0) staged_of_input[e.input] = p.n_staged++;
      p.stage[i] = (uint8_t)staged_of_input[e.input];
    }
  }
  // enough 16-byte loads in flight per thread (~6) whatever the number of streamed inputs
  const int ch = (sizeof(T) == 4 && p.n_staged <= 4) ? 2 : 1;
  int n_full = 0;  // streamed from HBM (broadcast leaves hit in cache)
  {
    bool seen[CHAIN_MAXN] = {false};
    for (int i = 0; i < p.n_steps; ++i) {
      if (p.kind[i] == SL_CHUNK) p.has_chunk = 1;
      if (p.kind[i] == SL_FULL && !seen[p.stage[i]]) { seen[p.stage[i]] = true; ++n_full; }
    }
  }
  const int per_iter = ch * (n_full > 0 ? n_full : 1) * (sizeof(T) == 8 ? 2 : 1);
  p.n_buf = 1 + (6 + per_iter - 1) / per_iter;
  if (p.n_buf < 2) p.n_buf = 2;
  if (p.n_buf > 4) p.n_buf = 4;
  if (ch == 2) *rc = launch_chain<T, 2>(p);
  else *rc = launch_chain<T, 1>(p);
  return true;
}

// ------------------------------------------------------------------ rand (Philox4x32-10)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

// Device-resident generator state for tcr_rand_unif_stream: {seed, offset}. Keeping the offset on the
// device lets a captured CUDA graph draw fresh numbers at every replay (the kernel reads the offset,
// a one-thread kernel behind it advances it), in the same sequence an eager run produces.
__device__ uint64_t g_rand_state[2];
__global__ void rand_seed_kernel(uint64_t seed, uint64_t offset) {
  TCR_PDL_ENTER(); g_rand_state[0] = seed; g_rand_state[1] = offset; }
__global__ void rand_advance_kernel(uint64_t n) {
  TCR_PDL_ENTER(); g_rand_state[1] += n; }

template <typename T, bool STREAM = false>
__global__ void __launch_bounds__(256) rand_unif_kernel(const T* __restrict__ lo, const T* __restrict__ hi,
                                                        T* __restrict__ out, int64_t n, uint64_t seed,
                                                        uint64_t offset) {
  TCR_PDL_ENTER();
  if (STREAM) {
    seed = g_rand_state[0];
    offset = g_rand_state[1];
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t ctr = offset + (uint64_t)i;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    T a = lo[i], b = hi[i];
    if (std::is_integral<T>::value) {
      // closed range [a, b] like std::uniform_int_distribution (global/random.hpp:87-98)
      uint64_t span = (uint64_t)((int64_t)b - (int64_t)a) + 1ull;
      uint64_t r64 = ((uint64_t)c[0] << 32) | c[1];
      out[i] = span == 0 ? T(r64) : T((int64_t)a + (int64_t)(r64 % span));
    } else {
      // 53-bit uniform in [0,1) then affine map like std::uniform_real_distribution<double>
      uint64_t r64 = ((uint64_t)c[0] << 32) | c[1];
      double u = (double)(r64 >> 11) * (1.0 / 9007199254740992.0);
      double v = (double)a + u * ((double)b - (double)a);
      T t = (T)v;
      if (t >= b && b > a) t = a;  // keep the half-open range after rounding to T
      out[i] = t;
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
  TCR_PDL_ENTER();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (TO)in[i];
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* __restrict__ buf, int64_t n, double s) {
  TCR_PDL_ENTER();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = (T)((double)buf[i] * s);
}
template <>
__global__ void __launch_bounds__(256) scale_kernel<float>(float* __restrict__ buf, int64_t n, double s) {
  TCR_PDL_ENTER();
  const float fs = (float)s;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] *= fs;
}

}  // namespace tcr

using namespace tcr;

template <typename T>
static int run_vm_multi(const tcr_ew_program* progs, int count, const uint8_t* keep) {
  static VmMulti pm;  // 7 KB: not on the stack; launches are serialised by the caller (one library stream)
  pm.count = count;
  pm.n_ext = 0;
  bool aligned = true;
  for (int k = 0; k < count; ++k) {
    int rc = fill_vm(&progs[k], pm.p[k], aligned, nullptr);
    if (rc) return rc;
    pm.store[k] = keep == nullptr || keep[k] != 0;
    for (int j = 0; j < progs[k].n_inputs; ++j) {
      const VmInput& in = pm.p[k].in[j];
      int from = -1;
      // the value an earlier program of this launch computes: same address, whole, same element type
      for (int q = k - 1; q >= 0 && from < 0; --q)
        if (in.mode == 0 && in.ptr == progs[q].outputs[0].ptr && in.dtype == progs[q].outputs[0].dtype && in.dtype == progs[k].dtype) from = q;
      pm.src[k][j] = from >= 0 ? (int16_t)(-(from + 1)) : (int16_t)pm.n_ext++;
    }
  }
  for (int k = 0; k < count; ++k)
    for (int q = 0; q < k; ++q)
      for (int j = 0; j < progs[k].n_inputs; ++j) {
        // a BROADCAST or re-typed read of an earlier result cannot be served by this thread's own stores
        const VmInput& in = pm.p[k].in[j];
        TCR_ARG(!(in.ptr == progs[q].outputs[0].ptr && pm.src[k][j] >= 0), "tcr_elementwise_multi: program %d reads program %d's result broadcast or re-typed", k, q);
      }
  if (pm.p[0].n == 0) return TCR_OK;
  const size_t dyn = sizeof(V4<T>) * (size_t)VM_MULTI_THREADS * (size_t)(pm.n_ext + count);
  static size_t configured[2] = {0, 0};
  if (dyn > 48 * 1024 - sizeof(V4<T>) * TCR_EW_NREGS * VM_MULTI_THREADS && dyn > configured[aligned]) {
    if (aligned) TCR_CUDA(cudaFuncSetAttribute(ew_vm_multi_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    else TCR_CUDA(cudaFuncSetAttribute(ew_vm_multi_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    configured[aligned] = dyn;
  }
  const int grid = wave_grid(ceil_div(pm.p[0].n, VM_V), VM_MULTI_THREADS, 3);
  if (aligned) TCR_LAUNCH((ew_vm_multi_kernel<T, true>), grid, VM_MULTI_THREADS, dyn, pm);
  else TCR_LAUNCH((ew_vm_multi_kernel<T, false>), grid, VM_MULTI_THREADS, dyn, pm);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

// ------------------------------------------------------------------ vector-along-segment-0 broadcast: out = un(bin(x, v))
// `x OP EXTEND(v)` (+ one unary) where v spans segment 0 and repeats over every row — a dense layer's bias + activation when it is
// not a GEMM epilogue (cfg/tenncor/layer.yml dense: CONTRACT, EXTEND, ADD). A thread keeps its 16 bytes of v in registers and
// streams rows: no index arithmetic and no per-element interpreter step, four 16-byte loads in flight per thread. The chain
// interpreter ran this shape at 0.45 of the copy bandwidth (issue-bound). Values are those of the separate functors (same
// Ops<T> bodies, each result rounded).
__device__ __forceinline__ float rowvec_unary(int un, float a) {
  switch (un) {
#define RVU(OP) case OP: return vm_un<float, OP>(a);
    RVU(TCR_EW_SIGMOID) RVU(TCR_EW_TANH) RVU(TCR_EW_EXP) RVU(TCR_EW_NEG) RVU(TCR_EW_SQUARE) RVU(TCR_EW_LOG)
    RVU(TCR_EW_SQRT) RVU(TCR_EW_ABS) RVU(TCR_EW_ROUND) RVU(TCR_EW_CUBE)
#undef RVU
    default: return a;
  }
}
template <int BIN, bool REV>
__global__ void __launch_bounds__(256) ew_rowvec_kernel(const float* x, const float* __restrict__ vec, float* out, uint32_t d0v,
                                                        uint32_t colblocks, uint32_t rows_per_block, int64_t rows, int un) {
  TCR_PDL_ENTER();
  constexpr int U = 4;
  const uint32_t cb = blockIdx.x % colblocks, rb = blockIdx.x / colblocks, nrb = gridDim.x / colblocks;
  uint32_t col, lane;
  if (rows_per_block == 1) { col = cb * 256u + threadIdx.x; lane = 0; }
  else { col = threadIdx.x % d0v; lane = threadIdx.x / d0v; }
  if (col >= d0v) return;
  const Vec<float> b = ld16(vec + 4 * (size_t)col);
  const int64_t rstride = (int64_t)nrb * rows_per_block;
  int64_t r = (int64_t)rb * rows_per_block + lane;
  auto apply = [&](Vec<float>& v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v.v[k] = REV ? vm_bin<float, BIN>(b.v[k], v.v[k]) : vm_bin<float, BIN>(v.v[k], b.v[k]);
    if (un != TCR_EW_NOP) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v.v[k] = rowvec_unary(un, v.v[k]);
    }
  };
  for (; r + (U - 1) * rstride < rows; r += U * rstride) {
    Vec<float> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld16(x + 4 * ((r + u * rstride) * d0v + col));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      apply(v[u]);
      st16(out + 4 * ((r + u * rstride) * d0v + col), v[u]);
    }
  }
  for (; r < rows; r += rstride) {
    Vec<float> v = ld16(x + 4 * (r * d0v + col));
    apply(v);
    st16(out + 4 * (r * d0v + col), v);
  }
}

// true when the program is `un(bin(full, vector))` over floats and was launched (or failed to launch: *rc)
static bool try_rowvec(const tcr_ew_program* prog, int* rc) {
  static const int enabled = std::getenv("TCR_EW_ROWVEC") ? std::atoi(std::getenv("TCR_EW_ROWVEC")) : 1;
  if (!enabled || prog->dtype != TCR_FLOAT || prog->n_inputs != 2 || prog->n_outputs != 1 || prog->n_instrs < 1 || prog->n_instrs > 2) return false;
  if (prog->outputs[0].dtype != TCR_FLOAT || prog->inputs[0].dtype != TCR_FLOAT || prog->inputs[1].dtype != TCR_FLOAT) return false;
  const int64_t d0 = prog->dims[0], rows = prog->dims[1] * prog->dims[2];
  if (d0 < 4 || (d0 & 3) || rows < 2 || d0 * rows < (1 << 15) || d0 / 4 >= (1ll << 31)) return false;
  const int64_t d0v = d0 / 4;
  if (d0v < 256 && 256 % d0v != 0) return false;
  int full = -1, vec = -1;
  for (int k = 0; k < 2; ++k) {
    const tcr_ew_input& in = prog->inputs[k];
    const bool b0 = in.bcast[0] != 0, b1 = in.bcast[1] && prog->dims[1] > 1, b2 = in.bcast[2] && prog->dims[2] > 1;
    const bool rep1 = b1 || prog->dims[1] == 1, rep2 = b2 || prog->dims[2] == 1;
    if (!b0 && !b1 && !b2) full = k;
    else if (!b0 && rep1 && rep2) vec = k;
  }
  if (full < 0 || vec < 0) return false;
  const tcr_ew_instr& i0 = prog->instrs[0];
  if (!((i0.a == full && i0.b == vec) || (i0.a == vec && i0.b == full))) return false;
  const bool rev = i0.a == vec;
  int un = TCR_EW_NOP;
  uint8_t result = i0.dst;
  if (prog->n_instrs == 2) {
    const tcr_ew_instr& i1 = prog->instrs[1];
    if (i1.a != result) return false;
    switch (i1.op) {
      case TCR_EW_SIGMOID: case TCR_EW_TANH: case TCR_EW_EXP: case TCR_EW_NEG: case TCR_EW_SQUARE: case TCR_EW_LOG:
      case TCR_EW_SQRT: case TCR_EW_ABS: case TCR_EW_ROUND: case TCR_EW_CUBE: break;
      default: return false;
    }
    un = i1.op;
    result = i1.dst;
  }
  if (prog->outputs[0].reg != result) return false;
  const float* x = (const float*)prog->inputs[full].ptr;
  const float* v = (const float*)prog->inputs[vec].ptr;
  float* out = (float*)prog->outputs[0].ptr;
  if (!aligned16(x) || !aligned16(v) || !aligned16(out)) return false;
  if ((const char*)v < (const char*)out + 4 * d0 * rows && (const char*)out < (const char*)v + 4 * d0) return false;  // vector inside the output
  const uint32_t colblocks = d0v >= 256 ? (uint32_t)ceil_div(d0v, 256) : 1u, rpb = d0v >= 256 ? 1u : (uint32_t)(256 / d0v);
  int64_t nrb = ceil_div(rows, (int64_t)rpb * 4);
  const int64_t cap = std::max<int64_t>(1, (int64_t)state().sm_count * 8 / colblocks);
  if (nrb > cap) nrb = cap;
  const int grid = (int)(nrb * colblocks);
#define RV_LAUNCH(OP)                                                                                                                \
  case OP:                                                                                                                           \
    if (rev) TCR_LAUNCH((ew_rowvec_kernel<OP, true>), grid, 256, 0, x, v, out, (uint32_t)d0v, colblocks, rpb, rows, un);              \
    else TCR_LAUNCH((ew_rowvec_kernel<OP, false>), grid, 256, 0, x, v, out, (uint32_t)d0v, colblocks, rpb, rows, un);                 \
    break;
  switch (i0.op) {
    RV_LAUNCH(TCR_EW_ADD) RV_LAUNCH(TCR_EW_SUB) RV_LAUNCH(TCR_EW_MUL) RV_LAUNCH(TCR_EW_DIV) RV_LAUNCH(TCR_EW_MIN) RV_LAUNCH(TCR_EW_MAX)
    default: return false;
  }
#undef RV_LAUNCH
  const cudaError_t e = cudaPeekAtLastError();
  *rc = e == cudaSuccess ? TCR_OK : fail_cuda(e, "kernel launch", __FILE__, __LINE__);
  return true;
}

extern "C" {

int tcr_elementwise(const tcr_ew_program* prog) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(prog != nullptr, "tcr_elementwise: null program");
  TCR_ARG(prog->n_inputs >= 0 && prog->n_inputs <= TCR_EW_MAX_INPUTS, "tcr_elementwise: bad n_inputs %d", prog->n_inputs);
  TCR_ARG(prog->n_outputs >= 1 && prog->n_outputs <= TCR_EW_MAX_OUTPUTS, "tcr_elementwise: bad n_outputs %d", prog->n_outputs);
  TCR_ARG(prog->n_instrs >= 0 && prog->n_instrs <= TCR_EW_MAX_INSTRS, "tcr_elementwise: bad n_instrs %d", prog->n_instrs);
  TCR_ARG(prog->dims[0] >= 0 && prog->dims[1] >= 0 && prog->dims[2] >= 0, "tcr_elementwise: negative dims");
  // single-op programs over same-type full operands take the direct vector kernels
  if (prog->n_instrs == 1 && prog->n_outputs == 1 && prog->outputs[0].dtype == prog->dtype &&
      prog->outputs[0].reg == prog->instrs[0].dst) {
    const tcr_ew_instr& ins = prog->instrs[0];
    int ar = op_arity(ins.op);
    bool ok = ar >= 1 && ar <= 3 && ins.op != TCR_EW_MOV && prog->n_inputs == ar && aligned16(prog->outputs[0].ptr);
    for (int k = 0; ok && k < prog->n_inputs; ++k) {
      const tcr_ew_input& in = prog->inputs[k];
      bool bc = (in.bcast[0] && prog->dims[0] > 1) || (in.bcast[1] && prog->dims[1] > 1) || (in.bcast[2] && prog->dims[2] > 1);
      ok = !bc && in.dtype == prog->dtype && aligned16(in.ptr);
    }
    ok = ok && ins.a == 0 && (ar < 2 || ins.b == 1) && (ar < 3 || ins.c == 2);
    if (ok) {
      int64_t n = prog->dims[0] * prog->dims[1] * prog->dims[2];
      if (n == 0) return TCR_OK;
      void* out = prog->outputs[0].ptr;
      const void* a = prog->inputs[0].ptr;
      const void* b = ar > 1 ? prog->inputs[1].ptr : nullptr;
      const void* c = ar > 2 ? prog->inputs[2].ptr : nullptr;
      TCR_DISPATCH_COMPUTE(prog->dtype, T, {
        if (ar == 1) return direct_unary<T>(ins.op, a, out, n);
        if (ar == 2) return direct_binary<T>(ins.op, a, b, out, n);
        return launch_direct<T, TCR_EW_SELECT, 3>(a, b, c, out, n);
      });
    }
  }
  {
    float scale = 1.f;
    int scaled = 0;
    if (match_pixel_cast(prog, &scale, &scaled)) {
      const int64_t n = prog->dims[0] * prog->dims[1] * prog->dims[2];
      if (n == 0) return TCR_OK;
      const int grid = wave_grid(ceil_div(n, (int64_t)16), 256, 8);
      TCR_LAUNCH(u8_to_f32_kernel, grid, 256, 0, (const uint8_t*)prog->inputs[0].ptr, (float*)prog->outputs[0].ptr, n, scale, scaled);
      TCR_CHECK_LAUNCH();
      return TCR_OK;
    }
  }
  {
    int rc = TCR_OK;
    if (try_rowvec(prog, &rc)) return rc;
  }
  TCR_DISPATCH_COMPUTE(prog->dtype, T, {
    int rc = TCR_OK;
    if (try_chain<T>(prog, &rc)) return rc;
    return run_vm<T>(prog);
  });
  return TCR_OK;
}

int tcr_elementwise_multi(const tcr_ew_program* progs, int count, const uint8_t* keep) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(progs != nullptr && count >= 1 && count <= VM_MULTI_MAX, "tcr_elementwise_multi: 1..%d programs (got %d)", VM_MULTI_MAX, count);
  if (count == 1) return tcr_elementwise(progs);
  const int64_t n = progs[0].dims[0] * progs[0].dims[1] * progs[0].dims[2];
  TCR_ARG(n < (1ll << 31), "tcr_elementwise_multi: iteration space too large");
  for (int k = 0; k < count; ++k) {
    const tcr_ew_program& q = progs[k];
    TCR_ARG(q.n_inputs >= 0 && q.n_inputs <= TCR_EW_MAX_INPUTS && q.n_outputs >= 1 && q.n_outputs <= TCR_EW_MAX_OUTPUTS && q.n_instrs >= 0 &&
                q.n_instrs <= TCR_EW_MAX_INSTRS, "tcr_elementwise_multi: program %d has bad counts", k);
    TCR_ARG(q.dtype == progs[0].dtype, "tcr_elementwise_multi: program %d computes in another type", k);
    TCR_ARG(q.dims[0] * q.dims[1] * q.dims[2] == n, "tcr_elementwise_multi: program %d has another iteration space", k);
  }
  switch (progs[0].dtype) {
    case TCR_FLOAT: return run_vm_multi<float>(progs, count, keep);
    case TCR_DOUBLE: return run_vm_multi<double>(progs, count, keep);
    case TCR_INT32: return run_vm_multi<int32_t>(progs, count, keep);
    case TCR_INT64: return run_vm_multi<int64_t>(progs, count, keep);
    default: set_error("tcr_elementwise_multi: no compute kernels for dtype %d", progs[0].dtype); return TCR_ERR_ARG;
  }
}

int tcr_cell_backward(const tcr_cell_backward_desc* d) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(d != nullptr && d->n >= 0 && d->n < (1ll << 31), "tcr_cell_backward: bad element count");
  TCR_ARG(d->n_gates >= 0 && d->n_gates <= 6, "tcr_cell_backward: 0..6 gates (got %d)", d->n_gates);
  TCR_ARG(d->s_a && d->s_b && d->c_x && d->c_y && d->c_z, "tcr_cell_backward: null operand");
  CellBwdParams p;
  memset(&p, 0, sizeof(p));
  p.n = (uint32_t)d->n;
  p.n_gates = d->n_gates;
  p.s_a = (const float*)d->s_a; p.s_b = (const float*)d->s_b;
  p.c_x = (const float*)d->c_x; p.c_y = (const float*)d->c_y; p.c_z = (const float*)d->c_z;
  p.s_out = (float*)d->s_out; p.c_out = (float*)d->c_out;
  bool vec = (d->n % 4) == 0 && aligned16(d->s_a) && aligned16(d->s_b) && aligned16(d->c_x) && aligned16(d->c_y) && aligned16(d->c_z) &&
             aligned16(d->s_out) && aligned16(d->c_out);
  for (int k = 0; k < d->n_gates; ++k) {
    TCR_ARG(d->kind[k] == 1 || d->kind[k] == 2, "tcr_cell_backward: gate %d has kind %d (1 SIGMOID, 2 TANH)", k, d->kind[k]);
    TCR_ARG(d->x[k] && d->y[k] && d->out[k], "tcr_cell_backward: gate %d has a null pointer", k);
    p.kind[k] = d->kind[k]; p.sel[k] = d->sel[k] ? 1 : 0;
    p.x[k] = (const float*)d->x[k]; p.y[k] = (const float*)d->y[k]; p.out[k] = (float*)d->out[k];
    vec = vec && aligned16(d->x[k]) && aligned16(d->y[k]) && aligned16(d->out[k]);
  }
  if (d->n == 0) return TCR_OK;
  if (vec) TCR_LAUNCH(cell_backward_kernel<4>, (int)ceil_div(d->n / 4, (int64_t)128), 128, 0, p);
  else TCR_LAUNCH(cell_backward_kernel<1>, (int)ceil_div(d->n, (int64_t)128), 128, 0, p);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_elementwise_reduce(const tcr_ew_program* prog, void* out, int post_op, double post_imm) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(prog != nullptr && out != nullptr, "tcr_elementwise_reduce: null argument");
  TCR_ARG(prog->n_inputs >= 0 && prog->n_inputs <= TCR_EW_MAX_INPUTS, "tcr_elementwise_reduce: bad n_inputs %d", prog->n_inputs);
  TCR_ARG(prog->n_outputs >= 1 && prog->n_outputs <= TCR_EW_MAX_OUTPUTS, "tcr_elementwise_reduce: bad n_outputs %d", prog->n_outputs);
  TCR_ARG(prog->n_instrs >= 0 && prog->n_instrs <= TCR_EW_MAX_INSTRS, "tcr_elementwise_reduce: bad n_instrs %d", prog->n_instrs);
  TCR_ARG(post_op >= 0 && post_op <= 2, "tcr_elementwise_reduce: bad post_op %d", post_op);
  TCR_ARG(prog->dtype == TCR_FLOAT || prog->dtype == TCR_DOUBLE, "tcr_elementwise_reduce: floating-point programs only");
  TCR_ARG(prog->outputs[0].dtype == prog->dtype, "tcr_elementwise_reduce: the summed output has the compute type");
  VmReduce red;
  red.out = out;
  red.post = post_op;
  red.imm = post_imm;
  if (prog->dtype == TCR_FLOAT) return run_vm<float>(prog, &red);
  return run_vm<double>(prog, &red);
}

static void prog1(tcr_ew_program* p, int op, int nin, const void* a, const void* b, const void* c, void* out,
                  int64_t n, int dtype) {
  memset(p, 0, sizeof(*p));
  p->dtype = dtype;
  p->n_inputs = nin;
  p->n_outputs = 1;
  p->n_instrs = 1;
  p->dims[0] = n; p->dims[1] = 1; p->dims[2] = 1;
  const void* ptrs[3] = {a, b, c};
  for (int k = 0; k < nin; ++k) { p->inputs[k].ptr = ptrs[k]; p->inputs[k].dtype = dtype; }
  p->outputs[0].ptr = out; p->outputs[0].dtype = dtype; p->outputs[0].reg = 7;
  p->instrs[0].op = (uint8_t)op; p->instrs[0].dst = 7; p->instrs[0].a = 0; p->instrs[0].b = 1; p->instrs[0].c = 2;
}

int tcr_unary(int opcode, const void* in, void* out, int64_t n, int dtype) {
  TCR_ARG(opcode >= TCR_EW_ABS && opcode <= TCR_EW_CUBE, "tcr_unary: opcode %d is not a unary op", opcode);
  tcr_ew_program p;
  prog1(&p, opcode, 1, in, nullptr, nullptr, out, n, dtype);
  return tcr_elementwise(&p);
}

int tcr_binary(int opcode, const void* a, const void* b, void* out, int64_t n, int dtype) {
  TCR_ARG(opcode >= TCR_EW_POW && opcode <= TCR_EW_GT, "tcr_binary: opcode %d is not a binary op", opcode);
  tcr_ew_program p;
  prog1(&p, opcode, 2, a, b, nullptr, out, n, dtype);
  return tcr_elementwise(&p);
}

int tcr_select(const void* cond, const void* then_, const void* else_, void* out, int64_t n, int dtype) {
  tcr_ew_program p;
  prog1(&p, TCR_EW_SELECT, 3, cond, then_, else_, out, n, dtype);
  return tcr_elementwise(&p);
}

int tcr_nnary(int opcode, const void* const* args, int nargs, void* out, int64_t n, int dtype) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(opcode == TCR_OP_ADD || opcode == TCR_OP_MUL, "tcr_nnary: opcode %d is not ADD/MUL", opcode);
  TCR_ARG(nargs >= 1, "tcr_nnary: no arguments");
  if (n == 0) return TCR_OK;
  if (nargs == 1) return tcr_d2d(out, args[0], (size_t)n * dtype_size(dtype));
  if (nargs == 2) return tcr_binary(opcode, args[0], args[1], out, n, dtype);
  if (nargs <= TCR_EW_MAX_INPUTS) {  // fused register-machine chain: one pass over memory
    tcr_ew_program p;
    memset(&p, 0, sizeof(p));
    p.dtype = dtype; p.n_inputs = nargs; p.n_outputs = 1; p.n_instrs = nargs - 1;
    p.dims[0] = n; p.dims[1] = 1; p.dims[2] = 1;
    for (int k = 0; k < nargs; ++k) { p.inputs[k].ptr = args[k]; p.inputs[k].dtype = dtype; }
    for (int k = 1; k < nargs; ++k) { p.instrs[k - 1].op = (uint8_t)opcode; p.instrs[k - 1].dst = 0; p.instrs[k - 1].a = 0; p.instrs[k - 1].b = (uint8_t)k; }
    p.outputs[0].ptr = out; p.outputs[0].dtype = dtype; p.outputs[0].reg = 0;
    return tcr_elementwise(&p);
  }
  // many operands (e.g. the 128 per-step weight gradients of an unrolled LSTM summed by
  // one n-ary ADD, internal/teq/src/derive.cpp:49-51): chunks of 8 registers, the
  // running result re-enters as operand 0 of the next chunk
  const void* chunk[TCR_EW_MAX_INPUTS];
  int done = 0;
  while (done < nargs) {
    int take = 0;
    if (done > 0) chunk[take++] = out;
    while (take < TCR_EW_MAX_INPUTS && done < nargs) chunk[take++] = args[done++];
    int rc = tcr_nnary(opcode, chunk, take, out, n, dtype);
    if (rc) return rc;
  }
  return TCR_OK;
}

int tcr_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n) {
  TCR_REQUIRE_DEVICE();
  if (n == 0) return TCR_OK;
  if (in_dtype == out_dtype) return tcr_d2d(out, in, (size_t)n * dtype_size(in_dtype));
  int grid = wave_grid(n, 256, 8);
  TCR_DISPATCH_ALL(in_dtype, TI, {
    TCR_DISPATCH_ALL(out_dtype, TO, TCR_LAUNCH((cast_kernel<TI, TO>), grid, 256, 0, (const TI*)in, (TO*)out, n));
  });
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_assign(int opcode, void* dst, const void* src, int64_t n, int dtype) {
  TCR_REQUIRE_DEVICE();
  switch (opcode) {
    case TCR_OP_ASSIGN: return tcr_d2d(dst, src, (size_t)n * dtype_size(dtype));
    case TCR_OP_ASSIGN_ADD: return tcr_binary(TCR_EW_ADD, dst, src, dst, n, dtype);
    case TCR_OP_ASSIGN_SUB: return tcr_binary(TCR_EW_SUB, dst, src, dst, n, dtype);
    case TCR_OP_ASSIGN_MUL: return tcr_binary(TCR_EW_MUL, dst, src, dst, n, dtype);
    case TCR_OP_ASSIGN_DIV: return tcr_binary(TCR_EW_DIV, dst, src, dst, n, dtype);
    default: set_error("tcr_assign: opcode %d is not an ASSIGN op", opcode); return TCR_ERR_ARG;
  }
}

int tcr_rand_unif(const void* lo, const void* hi, void* out, int64_t n, int dtype, uint64_t seed, uint64_t offset) {
  TCR_REQUIRE_DEVICE();
  if (n == 0) return TCR_OK;
  int grid = wave_grid(n, 256, 8);
  TCR_DISPATCH_COMPUTE(dtype, T, TCR_LAUNCH((rand_unif_kernel<T>), grid, 256, 0, (const T*)lo, (const T*)hi, (T*)out, n, seed, offset));
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_rand_seed(uint64_t seed, uint64_t offset) {
  TCR_REQUIRE_DEVICE();
  TCR_LAUNCH(rand_seed_kernel, 1, 1, 0, seed, offset);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_rand_unif_stream(const void* lo, const void* hi, void* out, int64_t n, int dtype) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(lo && hi && out, "tcr_rand_unif_stream: null argument");
  TCR_ARG(n >= 0, "tcr_rand_unif_stream: negative size");
  if (n == 0) return TCR_OK;
  int grid = wave_grid(n, 256, 8);
  TCR_DISPATCH_COMPUTE(dtype, T, TCR_LAUNCH((rand_unif_kernel<T, true>), grid, 256, 0, (const T*)lo, (const T*)hi, (T*)out, n, 0ull, 0ull));
  TCR_LAUNCH(rand_advance_kernel, 1, 1, 0, (uint64_t)n);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

int tcr_scale_inplace(void* buf, int64_t n, int dtype, double scale) {
  TCR_REQUIRE_DEVICE();
  if (n == 0) return TCR_OK;
  int grid = wave_grid(n, 256, 8);
  TCR_DISPATCH_COMPUTE(dtype, T, TCR_LAUNCH((scale_kernel<T>), grid, 256, 0, (T*)buf, n, scale));
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

}  // extern "C"
