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
  // packed instruction: op [0,8) | VM_* flags [8,16) | dst [16,20) | a [20,24) | b [24,28) | c [28,32)
  uint32_t insw[TCR_EW_MAX_INSTRS];
  double imm[TCR_EW_MAX_INSTRS];
  uint8_t out_fwd[TCR_EW_MAX_OUTPUTS];  // output k is the value of the last instruction
};
// result forwarding: the value an instruction produces stays in hardware registers for the next
// instruction; it is written to its shared-memory slot only when something later still reads it
enum { VM_FWD_A = 1, VM_FWD_B = 2, VM_FWD_C = 4, VM_STORE = 8 };

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

// The virtual registers live in shared memory, one 4-element slot per (register, chunk, thread):
// operand fetch is an LDS.128 with a computed address instead of a branch tree over hardware
// registers. ncu (profiles/r1_ncu_ew_vm.md) showed the first version issue-bound (72 % issue
// slots, ~410 instructions per 4 elements, 85 % of them decode / address arithmetic / branches),
// so: (1) an instruction is one packed 32-bit word (op, flags, register numbers) fetched with a
// single constant load; (2) each thread runs the program over VM_CH chunks per decode, halving
// the per-element cost of decode, dispatch and loop control; (3) the value an instruction
// produces is forwarded in hardware registers to the next instruction and only written to its
// shared-memory slot when something later reads it.
template <typename T> struct VmChunks { static constexpr int N = sizeof(T) == 4 ? 2 : 1; };

template <typename T, typename I>
__device__ __forceinline__ V4<T> vm_load_input(const VmInput& in, I base, I n, I d0, I d1, bool full, bool aligned) {
  V4<T> x;
  if (in.mode == 1) {
    const T s = load_any<T>(in.ptr, in.dtype, 0);
#pragma unroll
    for (int v = 0; v < VM_V; ++v) x.v[v] = s;
  } else if (in.mode == 0) {
    if (aligned && full && in.dtype == DTypeOf<T>::value) {
      const T* src = (const T*)in.ptr + base;
      *reinterpret_cast<uint4*>(x.v) = *reinterpret_cast<const uint4*>(src);
      if (sizeof(T) == 8) *reinterpret_cast<uint4*>(x.v + 2) = *reinterpret_cast<const uint4*>(src + 2);
    } else {
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = (base + v < n) ? load_any<T>(in.ptr, in.dtype, base + v) : T(0);
    }
  } else if (in.mode == 3) {
    // 3-segment broadcast with D0 % 4 == 0: the chunk stays inside one run of segment 0
    const I e0 = in.bcast[0] ? 1 : d0, e1 = in.bcast[1] ? 1 : d1;
    const I i0 = base % d0, t = base / d0, i1 = t % d1, i2 = t / d1;
    const I j = (in.bcast[0] ? 0 : i0) + e0 * ((in.bcast[1] ? 0 : i1) + e1 * (in.bcast[2] ? 0 : i2));
    if (base >= n) {
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = T(0);
    } else if (in.bcast[0]) {
      const T s = load_any<T>(in.ptr, in.dtype, j);
#pragma unroll
      for (int v = 0; v < VM_V; ++v) x.v[v] = s;
    } else if (aligned && in.dtype == DTypeOf<T>::value) {
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
      const I i = base + v;
      if (i >= n) { x.v[v] = T(0); continue; }
      const I i0 = i % d0, t = i / d0, i1 = t % d1, i2 = t / d1;
      const I j = (in.bcast[0] ? 0 : i0) + e0 * ((in.bcast[1] ? 0 : i1) + e1 * (in.bcast[2] ? 0 : i2));
      x.v[v] = load_any<T>(in.ptr, in.dtype, j);
    }
  }
  return x;
}

template <typename T, bool ALIGNED, typename I>
__global__ void __launch_bounds__(VmCfg<T>::THREADS) ew_vm_kernel(const __grid_constant__ VmParams p) {
  constexpr int THREADS = VmCfg<T>::THREADS, CH = VmChunks<T>::N;
  extern __shared__ __align__(16) unsigned char vm_smem[];
  // slot of (register r, chunk c, thread t) = ((r * CH + c) * THREADS + t)
  V4<T>* const mine = reinterpret_cast<V4<T>*>(vm_smem) + threadIdx.x;
  constexpr int RSTRIDE = CH * THREADS;  // slots between consecutive registers
  const I n = (I)p.n;
  const I nchunks = (n + VM_V - 1) / VM_V;
  const I stride = (I)gridDim.x * (THREADS * CH);
  const I d0 = (I)p.d0, d1 = (I)p.d1;
  const int n_inputs = p.n_inputs, n_instrs = p.n_instrs, n_outputs = p.n_outputs;
  for (I ch0 = (I)blockIdx.x * (THREADS * CH) + threadIdx.x; ch0 < nchunks; ch0 += stride) {
    // ---- load inputs into registers 0..n_inputs-1 (all global loads are issued before the first use)
    for (int k = 0; k < n_inputs; ++k) {
      const VmInput& in = p.in[k];
      V4<T> x[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const I base = (ch0 + (I)c * THREADS) * VM_V;
        x[c] = vm_load_input<T, I>(in, base, n, d0, d1, base + VM_V <= n, ALIGNED);
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) mine[(k * CH + c) * THREADS] = x[c];
    }
    // ---- execute
    V4<T> last[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int v = 0; v < VM_V; ++v) last[c].v[v] = T(0);
    for (int pc = 0; pc < n_instrs; ++pc) {
      const uint32_t w = p.insw[pc];
      const int op = w & 0xff, fl = (w >> 8) & 0xff;
      const V4<T>* ra = mine + ((w >> 20) & 0xf) * RSTRIDE;
      const V4<T>* rb = mine + ((w >> 24) & 0xf) * RSTRIDE;
      V4<T> a[CH], d[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (fl & VM_FWD_A) a[c] = last[c];
        else a[c] = ra[c * THREADS];  // (CONST reads a slot it ignores: cheaper than another branch)
      }
      if (op >= TCR_EW_POW && op <= TCR_EW_GT) {
        V4<T> b[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          if (fl & VM_FWD_B) b[c] = last[c];
          else b[c] = rb[c * THREADS];
        }
        switch (op) {
#define VMB(OP) case OP: _Pragma("unroll") for (int c = 0; c < CH; ++c) _Pragma("unroll") for (int v = 0; v < VM_V; ++v) d[c].v[v] = Ops<T>::bin(OP, a[c].v[v], b[c].v[v]); break;
          VMB(TCR_EW_ADD) VMB(TCR_EW_SUB) VMB(TCR_EW_MUL) VMB(TCR_EW_DIV) VMB(TCR_EW_POW) VMB(TCR_EW_MIN)
          VMB(TCR_EW_MAX) VMB(TCR_EW_EQ) VMB(TCR_EW_NEQ) VMB(TCR_EW_LT) VMB(TCR_EW_GT)
#undef VMB
          default:
#pragma unroll
            for (int c = 0; c < CH; ++c) d[c] = a[c];
        }
      } else if (op == TCR_EW_CONST) {
        const T imm = (T)p.imm[pc];
#pragma unroll
        for (int c = 0; c < CH; ++c)
#pragma unroll
          for (int v = 0; v < VM_V; ++v) d[c].v[v] = imm;
      } else if (op == TCR_EW_SELECT) {
        const V4<T>* rc = mine + ((w >> 28) & 0xf) * RSTRIDE;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          V4<T> b, cc;
          if (fl & VM_FWD_B) b = last[c];
          else b = rb[c * THREADS];
          if (fl & VM_FWD_C) cc = last[c];
          else cc = rc[c * THREADS];
#pragma unroll
          for (int v = 0; v < VM_V; ++v) d[c].v[v] = (a[c].v[v] != T(0)) ? b.v[v] : cc.v[v];
        }
      } else {
        switch (op) {
#define VMU(OP) case OP: _Pragma("unroll") for (int c = 0; c < CH; ++c) _Pragma("unroll") for (int v = 0; v < VM_V; ++v) d[c].v[v] = Ops<T>::un(OP, a[c].v[v]); break;
          VMU(TCR_EW_SIGMOID) VMU(TCR_EW_TANH) VMU(TCR_EW_EXP) VMU(TCR_EW_NEG) VMU(TCR_EW_SQUARE) VMU(TCR_EW_LOG)
          VMU(TCR_EW_SQRT) VMU(TCR_EW_ABS) VMU(TCR_EW_SIN) VMU(TCR_EW_COS) VMU(TCR_EW_TAN) VMU(TCR_EW_ROUND)
          VMU(TCR_EW_CUBE)
#undef VMU
          default:  // MOV
#pragma unroll
            for (int c = 0; c < CH; ++c) d[c] = a[c];
        }
      }
      if (fl & VM_STORE) {
        V4<T>* rd = mine + ((w >> 16) & 0xf) * RSTRIDE;
#pragma unroll
        for (int c = 0; c < CH; ++c) rd[c * THREADS] = d[c];
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) last[c] = d[c];
    }
    // ---- store outputs
    for (int k = 0; k < n_outputs; ++k) {
      const tcr_ew_output& o = p.out[k];
      const bool fwd = p.out_fwd[k];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const I base = (ch0 + (I)c * THREADS) * VM_V;
        if (base >= n) continue;
        V4<T> y;
        if (fwd) y = last[c];
        else y = mine[(o.reg * CH + c) * THREADS];
        if (ALIGNED && base + VM_V <= n && o.dtype == DTypeOf<T>::value) {
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

template <typename T>
static int run_vm(const tcr_ew_program* prog) {
  VmParams p;
  memset(&p, 0, sizeof(p));
  p.n_inputs = prog->n_inputs;
  p.n_outputs = prog->n_outputs;
  p.n_instrs = prog->n_instrs;
  p.d0 = prog->dims[0];
  p.d1 = prog->dims[1];
  p.n = prog->dims[0] * prog->dims[1] * prog->dims[2];
  bool aligned = true;
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
  }
  // pack the instructions and compute the forwarding flags
  const tcr_ew_instr* ins = prog->instrs;
  for (int i = 0; i < prog->n_instrs; ++i) {
    const int ar = op_arity(ins[i].op);
    uint32_t fl = 0;
    if (i > 0) {
      const uint8_t prev = ins[i - 1].dst;
      if (ar >= 1 && ins[i].a == prev) fl |= VM_FWD_A;
      if (ar >= 2 && ins[i].b == prev) fl |= VM_FWD_B;
      if (ar >= 3 && ins[i].c == prev) fl |= VM_FWD_C;
    }
    bool need = false, overwritten = false;
    for (int j = i + 1; j < prog->n_instrs && !overwritten; ++j) {
      const tcr_ew_instr& u = ins[j];
      const int aj = op_arity(u.op);
      const bool reads = (aj >= 1 && u.a == ins[i].dst) || (aj >= 2 && u.b == ins[i].dst) || (aj >= 3 && u.c == ins[i].dst);
      if (reads && j != i + 1) need = true;  // j == i + 1 takes the forwarded copy
      if (u.dst == ins[i].dst) overwritten = true;
    }
    if (!overwritten)
      for (int k = 0; k < prog->n_outputs; ++k)
        if (prog->outputs[k].reg == ins[i].dst && i != prog->n_instrs - 1) need = true;
    if (need) fl |= VM_STORE;
    p.insw[i] = (uint32_t)ins[i].op | (fl << 8) | ((uint32_t)ins[i].dst << 16) | ((uint32_t)ins[i].a << 20) |
                ((uint32_t)ins[i].b << 24) | ((uint32_t)ins[i].c << 28);
    p.imm[i] = ins[i].imm;
  }
  for (int k = 0; k < prog->n_outputs; ++k)
    p.out_fwd[k] = prog->n_instrs > 0 && prog->outputs[k].reg == ins[prog->n_instrs - 1].dst;
  if (p.n == 0) return TCR_OK;
  constexpr int THREADS = VmCfg<T>::THREADS, CH = VmChunks<T>::N;
  constexpr size_t SMEM = (size_t)TCR_EW_NREGS * CH * THREADS * sizeof(V4<T>);
  static bool configured = false;
  if (!configured) {
    TCR_CUDA(cudaFuncSetAttribute(ew_vm_kernel<T, true, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    TCR_CUDA(cudaFuncSetAttribute(ew_vm_kernel<T, false, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    TCR_CUDA(cudaFuncSetAttribute(ew_vm_kernel<T, true, int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    TCR_CUDA(cudaFuncSetAttribute(ew_vm_kernel<T, false, int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  const int per_sm = (int)(220 * 1024 / SMEM) < 6 ? (int)(220 * 1024 / SMEM) : 6;
  int grid = wave_grid(ceil_div(p.n, VM_V), THREADS * CH, per_sm);
  const bool small = p.n < (1ll << 31) - (int64_t)4 * THREADS * CH * 148 * 8;
  if (small) {
    if (aligned) TCR_LAUNCH((ew_vm_kernel<T, true, uint32_t>), grid, THREADS, SMEM, p);
    else TCR_LAUNCH((ew_vm_kernel<T, false, uint32_t>), grid, THREADS, SMEM, p);
  } else {
    if (aligned) TCR_LAUNCH((ew_vm_kernel<T, true, int64_t>), grid, THREADS, SMEM, p);
    else TCR_LAUNCH((ew_vm_kernel<T, false, int64_t>), grid, THREADS, SMEM, p);
  }
  TCR_CHECK_LAUNCH();
  return TCR_OK;
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

template <typename T>
__global__ void __launch_bounds__(256) rand_unif_kernel(const T* __restrict__ lo, const T* __restrict__ hi,
                                                        T* __restrict__ out, int64_t n, uint64_t seed,
                                                        uint64_t offset) {
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
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (TO)in[i];
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* __restrict__ buf, int64_t n, double s) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = (T)((double)buf[i] * s);
}
template <>
__global__ void __launch_bounds__(256) scale_kernel<float>(float* __restrict__ buf, int64_t n, double s) {
  const float fs = (float)s;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] *= fs;
}

}  // namespace tcr

using namespace tcr;

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
  TCR_DISPATCH_COMPUTE(prog->dtype, T, return run_vm<T>(prog));
  return TCR_OK;
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

int tcr_scale_inplace(void* buf, int64_t n, int dtype, double scale) {
  TCR_REQUIRE_DEVICE();
  if (n == 0) return TCR_OK;
  int grid = wave_grid(n, 256, 8);
  TCR_DISPATCH_COMPUTE(dtype, T, TCR_LAUNCH((scale_kernel<T>), grid, 256, 0, (T*)buf, n, scale));
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

}  // extern "C"
