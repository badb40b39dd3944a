// gemm_rnn.cu — grouped / segmented fp32 product for SMALL-BATCH recurrent steps on the tcgen05 tensor cores.
//
// What it replaces. The reference unrolls layer.lstm / layer.gru (cfg/tenncor/layer.yml:716-813) into, per time step,
//   4 x SIGMOID|TANH(ADD(CONTRACT(CONCAT(x_t, h_{t-1}), W_g), EXTEND(b_g)))          forward: one product per gate
//   ADD(CONTRACT(dpre_g, W_g) for g in gates)                                         backward: sum of products
// (internal/eigen/operator.hpp:1069-1139 for CONTRACT, :336-368 CONCAT, :716-731 n-ary ADD). With the batch as the only
// "large" row extent (64 rows at BASELINE's C4), each of those is a weight-streaming product: 1152 x 1024 weights per gate
// against 64 rows. The general GEMM kernels pad the batch to a 128-row tile and need a split-K workspace + second launch.
//
// This kernel computes, for up to 4 output groups g and up to 4 K-segments s,
//     out_g[m, n] = act_g( sum_s A_s[m, :] . B_{g,s}[:, n] + bias_g[n] )         m = batch rows (any), n = units
// in ONE launch:
//   * roles are swapped on the tensor core: the weights are the MMA "A" operand (128 rows per CTA = `groups` row groups
//     of 128/groups units each, so one CTA holds ALL gates of its units), the batch is the MMA "N" extent (64 per CTA);
//   * K-segments have their own tensor maps: CONCAT(x_t, h_{t-1}) is never materialised (forward), and the sum over gates
//     of the backward products is one accumulation in TMEM (backward);
//   * split-K runs inside a thread-block CLUSTER (1,1,C): CTA c streams its share of the k-blocks, then the partial
//     accumulators are exchanged through distributed shared memory (reduce-scatter over the batch columns, fixed order =>
//     deterministic) — no workspace, no second kernel;
//   * optional LSTM cell epilogue: c_t = cand * in + c_{t-1} * forget, h_t = c_t * out (layer.yml:758-760) computed by the
//     CTA that owns all four gates of its units.
// Pipeline per CTA: warp 0 = TMA producer, warp 1 = MMA issuer (tcgen05.mma kind::tf32, 128 x 64 x 8, fp32 accumulator in
// TMEM), warps 2-5 = 3xTF32 lo-part converters, then epilogue.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace tcr {

int make_tf32_map(CUtensorMap* map, const float* base, int64_t dim0, int64_t dim1, int64_t pitch, uint32_t box0, uint32_t box1, bool mn_major);  // gemm_tc.cu

namespace {

constexpr int RM = 128, RN = 64, RK = 32;
constexpr int W_TILE = RM * RK * 4;  // 16 KiB
constexpr int X_TILE = RN * RK * 4;  // 8 KiB
constexpr int THREADS = 256;
constexpr int PRODUCERS = 3;  // warps 0, 6, 7: a warp issues the five boxes of a k-block in ~700 cycles (ELECT + R2UR round trip per lane), so k-blocks are dealt to three of them
constexpr int NACC = 1;  // TMEM accumulators used round-robin by the k-steps of a k-block and summed in the epilogue. Measured: 4 independent
                         // accumulators do not speed the MMAs up (the ~83 cycles per 128 x 64 x 8 MMA are not an accumulator dependency) and
                         // cost ~300 cycles of extra TMEM loads, so one is used.
constexpr int RED_BYTES = RM * RN * 4;        // 32 KiB: [src CTA][column][row]
constexpr int CELL_BYTES = 4 * RN * 32 * 4;   // 32 KiB: [gate][column][unit <= 32]
constexpr uint32_t SPIN_LIMIT = 1u << 28;

struct alignas(64) RnnMaps {
  CUtensorMap w[8];  // [group * nseg + segment]
  CUtensorMap x[4];  // [segment]
};

struct RnnParams {
  int64_t m, n;
  int groups, rows_per_group, nseg;
  int seg_kb_end[4];  // running count of k-blocks after segment s
  int w_k0[4];        // k coordinate at which segment s starts inside its W map
  int w_mn_major;     // 1: W element (k, n) has n contiguous (forward); 0: k contiguous (backward)
  int kb_total, kb_per_cta;
  float* out[4];
  const float* bias[4];
  int act[4];
  int64_t out_pitch;
  int accumulate;
  int cell, role_cand, role_in, role_forget, role_out;
  const float* c_prev;
  float* c_out;
  float* h_out;
  int64_t state_pitch;
  int dbg_mode;    // TCR_RNN_EXPERIMENT: 1 = no MMAs (TMA only), 2 = no TMA loads (MMAs on whatever shared memory holds), 4 = X tile not loaded, 8 = no lo-part conversion, 16 = lo parts of the activation tile only
  long long* dbg;  // TCR_RNN_DEBUG: SM-clock stamps of CTA (0,0,0), see tcr_rnn_debug_read
};

// One launch over T consecutive time steps (tcr_gemm_grouped_seq_*): what differs from step to step
struct alignas(64) RnnStep {
  float* out[4];
  const float* c_prev;
  float* c_out;
  float* h_out;
};
struct RnnSeq {
  const CUtensorMap* xmaps;  // [T][nseg] activation maps in device memory
  const RnnStep* steps;      // [T]
  int T;
  uint32_t dep_segs;         // bit s: segment s of step t reads what step t - 1 wrote (h_{t-1}): its loads wait for the grid barrier
  unsigned* bar;             // [0] arrivals (monotonic over the launch), [1] CTAs that finished; both zero between launches
  unsigned num_ctas;
};

#define RNN_STAMP(slot) do { if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[slot] = clock64(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int a_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                           // accumulator F32
  d |= 2u << 7;                           // A = TF32
  d |= 2u << 10;                          // B = TF32
  d |= (uint32_t)(a_mn_major & 1) << 15;  // A major
  d |= (uint32_t)(RN >> 3) << 17;         // N
  d |= (uint32_t)(RM >> 4) << 24;         // M
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float act_f(int act, float x) {
  if (act == TCR_EW_SIGMOID) return __fdividef(1.0f, 1.0f + expf(-x));
  if (act == TCR_EW_TANH) return tanhf(x);
  return x;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// Epilogue of one CTA of a cluster of CS = 64 / NC CTAs (warps 2..5, 128 threads; thread t owns accumulator row t):
//   1. its 64 accumulator columns go to their owners: column c belongs to CTA c / NC and lands in that CTA's
//      red[(sender * NC + c % NC) * 128 + t]  (st.shared::cluster, 128-byte coalesced per warp);
//   2. cluster barrier (release / acquire);
//   3. the NC columns this CTA owns are summed over the senders in rank order (deterministic), bias + activation are
//      applied and the rows are stored; with the LSTM cell the four gate values of a (unit, batch row) meet through shared
//      memory and c_t, h_t are written as well.
// Everything that does not depend on the accumulator (bias, c_{t-1}) is fetched BEFORE waiting for it.
template <int NC>
__device__ __forceinline__ void rnn_epilogue(const RnnParams& p, const RnnStep& sp, uint64_t* tmem_full, uint32_t full_parity, uint64_t* tmem_empty,
                                             uint32_t tmem_base, uint32_t red_s, uint32_t cell_s, int num_kb, int u0, int m0, uint32_t crank) {
  constexpr int CS = RN / NC;
  constexpr int CELLS = NC >= 4 ? NC / 4 : 1;  // (unit, column) pairs per thread in the cell phase
  const int t = threadIdx.x - 64, lane = threadIdx.x & 31, q = (threadIdx.x >> 5) & 3;
  const int row = 32 * q + lane;  // TMEM lane this thread reads (warp w may only touch lanes 32 (w % 4) ..)
  // ---- operands of the final phase that are already in memory
  const int g = t / p.rows_per_group, ul = t % p.rows_per_group;
  const int u = u0 + ul;
  const bool u_ok = u < p.n;
  const float bias = (u_ok && p.bias[g] != nullptr) ? __ldg(p.bias[g] + u) : 0.f;
  const int act = p.act[g];
  float* const out = sp.out[g];
  const int b0 = m0 + (int)crank * NC;  // first batch row this CTA finishes
  const int ul2 = t & 31, u2 = u0 + ul2;
  float cprev[CELLS];
#pragma unroll
  for (int i = 0; i < CELLS; ++i) {
    const int b = b0 + (t >> 5) + 4 * i;
    // plain load (not the read-only path): in a multi-step launch this is what the same thread stored one step earlier
    cprev[i] = (p.cell && sp.c_prev != nullptr && u2 < p.n && b < p.m) ? sp.c_prev[(int64_t)b * p.state_pitch + u2] : 0.f;
  }
  float old[NC];
  if (p.accumulate) {
#pragma unroll
    for (int j = 0; j < NC; ++j) old[j] = (u_ok && b0 + j < p.m && out != nullptr) ? out[(int64_t)(b0 + j) * p.out_pitch + u] : 0.f;
  }
  // ---- 1. scatter the partial sums
  {
    float v[RN];
    if (num_kb > 0) {
      mbar_wait(tmem_full, full_parity);
      if (threadIdx.x == 64) RNN_STAMP(6);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int j = 0; j < RN; ++j) v[j] = 0.f;
#pragma unroll
      for (int part = 0; part < 2 * NACC; ++part) {  // NACC accumulators x two 32-column halves
        uint32_t r[32];
        const int half = part & 1;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((part >> 1) * RN + half * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) v[half * 32 + j] += __uint_as_float(r[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < RN; ++j) v[j] = 0.f;
    }
    if (tmem_empty != nullptr) {  // multi-step launch: the accumulator may be overwritten by the next step's first MMA
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
    const uint32_t mine = red_s + (uint32_t)(((int)crank * NC) * RM + row) * 4u;
    if (CS == 1) {
#pragma unroll
      for (int j = 0; j < NC; ++j) sts_f32(mine + (uint32_t)(j * RM) * 4u, v[j]);
    } else {
#pragma unroll
      for (int dst = 0; dst < CS; ++dst) {
        const uint32_t base = map_to_cta(mine, (uint32_t)dst);
#pragma unroll
        for (int j = 0; j < NC; ++j) st_cluster_f32(base + (uint32_t)(j * RM) * 4u, v[dst * NC + j]);
      }
    }
    if (threadIdx.x == 64) RNN_STAMP(7);
  }
  // ---- 2.
  cluster_arrive();  // #2 (release): my partial sums are in their owners' shared memory
  cluster_wait();    //    (acquire): everybody's are in mine
  if (threadIdx.x == 64) RNN_STAMP(8);
  // ---- 3. rows of the owned columns
  float x[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) x[j] = 0.f;
#pragma unroll
  for (int src = 0; src < CS; ++src)
#pragma unroll
    for (int j = 0; j < NC; ++j) x[j] += lds_f32(red_s + (uint32_t)((src * NC + j) * RM + t) * 4u);
  if (act == TCR_EW_SIGMOID) {
#pragma unroll
    for (int j = 0; j < NC; ++j) x[j] = __fdividef(1.0f, 1.0f + expf(-(x[j] + bias)));
  } else if (act == TCR_EW_TANH) {
#pragma unroll
    for (int j = 0; j < NC; ++j) x[j] = tanhf(x[j] + bias);
  } else {
#pragma unroll
    for (int j = 0; j < NC; ++j) x[j] += bias;
  }
  if (u_ok && out != nullptr) {
#pragma unroll
    for (int j = 0; j < NC; ++j)
      if (b0 + j < p.m) out[(int64_t)(b0 + j) * p.out_pitch + u] = p.accumulate ? old[j] + x[j] : x[j];
  }
  if (p.cell) {  // groups == 4, rows_per_group == 32: gate g of unit ul, column j at cell[(g * NC + j) * 32 + ul]
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the cell buffer IS the reduce buffer: everybody has read its sums out of it
#pragma unroll
    for (int j = 0; j < NC; ++j) sts_f32(cell_s + (uint32_t)((g * NC + j) * 32 + ul) * 4u, x[j]);
    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
    for (int i = 0; i < CELLS; ++i) {
      const int j = (t >> 5) + 4 * i;
      const int b = b0 + j;
      if (j < NC && u2 < p.n && b < p.m) {
        const float cand = lds_f32(cell_s + (uint32_t)((p.role_cand * NC + j) * 32 + ul2) * 4u);
        const float in = lds_f32(cell_s + (uint32_t)((p.role_in * NC + j) * 32 + ul2) * 4u);
        const float forget = lds_f32(cell_s + (uint32_t)((p.role_forget * NC + j) * 32 + ul2) * 4u);
        const float og = lds_f32(cell_s + (uint32_t)((p.role_out * NC + j) * 32 + ul2) * 4u);
        const int64_t at = (int64_t)b * p.state_pitch + u2;
        const float c = __fadd_rn(__fmul_rn(cand, in), __fmul_rn(cprev[i], forget));  // ADD(MUL(gate, in), MUL(state, forget)): no fma contraction
        sp.c_out[at] = c;
        sp.h_out[at] = c * og;
      }
    }
  }
}

template <int MODE>
struct RnnCfg {
  // The main loop is bound by the turn-around of a ring stage (TMA land ~1160 cycles + lo-part conversion + 12 MMAs + commit,
  // ~2750 cycles), not by the MMAs (384 cycles per k-block): a k-block costs turn-around / STAGES. A fourth 48 KB stage fits
  // in 3xTF32 once the cell-exchange buffer shares the (already consumed) reduce buffer: 3 stages ~1000 cycles per k-block.
  static constexpr int STAGES = MODE == 2 ? 4 : 6;
  static constexpr int STAGE_BYTES = (MODE == 2 ? 2 : 1) * (W_TILE + X_TILE);
  static constexpr int SMEM = STAGES * STAGE_BYTES + RED_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(CELL_BYTES <= RED_BYTES, "the cell buffer lives inside the reduce buffer");
  static_assert(SMEM <= 232448, "dynamic shared memory of one CTA");
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// SEQ = false: one product (tcr_gemm_grouped). SEQ = true: T dependent time steps in ONE launch (tcr_gemm_grouped_seq_*): the
// CTAs stay resident, barriers / TMEM / tensor maps are set up once, every role keeps its ring position across steps, and
// between steps the grid meets at a counter in global memory — the epilogue threads of a CTA publish their h_t rows and
// arrive, the one lane that fetches activation tiles waits before its first load of a segment that reads h_{t-1} (weight
// tiles of the next step are already on their way by then). All CTAs must be co-resident (checked by the host).
template <int MODE, bool SEQ>
__global__ void __launch_bounds__(THREADS, 1)
gemm_rnn_kernel(const __grid_constant__ RnnMaps maps, const __grid_constant__ RnnParams p, const __grid_constant__ RnnSeq sq) {
  constexpr int STAGES = RnnCfg<MODE>::STAGES, STAGE_BYTES = RnnCfg<MODE>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* red = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  float* cellbuf = red;  // reused once every thread has summed its columns out of `red` (bar.sync in rnn_epilogue)
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES + RED_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* ready = bars + 2 * STAGES;
  uint64_t* tmem_full = bars + 3 * STAGES;
  uint64_t* tmem_empty = bars + 3 * STAGES + 1;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) RNN_STAMP(0);
  // the cluster is (1, 1, C) and the grid is C deep: rank and size are blockIdx.z / gridDim.z, which — unlike the %cluster_*
  // registers read through inline assembly — the compiler knows to be warp-uniform
  const uint32_t crank = blockIdx.z, csize = gridDim.z;
  const int u0 = (int)blockIdx.x * p.rows_per_group;  // first unit (column of every out_g) of this CTA
  const int m0 = (int)blockIdx.y * RN;                // first batch row
  const int kb_begin = (int)crank * p.kb_per_cta;
  const int num_kb = max(0, min(p.kb_per_cta, p.kb_total - kb_begin));
  const int T = SEQ ? sq.T : 1;

  auto tile_w = [&](int s) { return smem + s * STAGE_BYTES; };

  if (threadIdx.x == 32) {  // tensor maps live in kernel parameter space: fetch them now, not inside the first TMA
    for (int i = 0; i < p.groups * p.nseg; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.w[i]) : "memory");
    if (!SEQ)
      for (int i = 0; i < p.nseg; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.x[i]) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&ready[s], 4);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(RN * NACC));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  TCR_PDL_ENTER();  // everything above touched no global memory: the previous kernel may still be running
  if (threadIdx.x == 0) RNN_STAMP(1);
  cluster_arrive();  // #1: "this CTA is running" — awaited before anybody writes into a peer's shared memory

  if (warp == 0 || warp >= 6) {
    // ================= TMA producers =================
    const uint32_t pw = warp == 0 ? 0u : (uint32_t)(warp - 5);  // k-block `it` (counted over all steps) belongs to producer it % PRODUCERS
    // One lane per box: lanes 0..3 fetch the weight boxes, lane 4 the activation tile; lane 0 also arms the barrier. (One thread
    // issuing five TMAs with their address arithmetic cost ~950 cycles per k-block: the loop was bound by that thread, not by L2.)
    {
      // per-lane constants of the weight box this lane fetches
      const int chunks_per_group = 4 / p.groups;
      int my_map = 0, c0 = 0, c1 = 0;         // tensor-map index (without the segment), fixed coordinates
      uint32_t my_off = 0;                    // byte offset of the box inside the weight tile
      bool w_lane = false;
      if (p.w_mn_major) {                      // four [32 k][32 units] boxes
        w_lane = lane < 4;
        const int g = lane / chunks_per_group, sub = lane % chunks_per_group;
        my_map = g * p.nseg;
        c0 = u0 + 32 * sub;
        my_off = (uint32_t)lane * 4096u;
      } else {                                 // one [rows_per_group][32 k] box per group
        w_lane = lane < p.groups;
        my_map = lane * p.nseg;
        c1 = u0;
        my_off = (uint32_t)(lane * p.rows_per_group) * 128u;
      }
      const bool x_lane = lane == 4 && !(p.dbg_mode & 4);
      const uint32_t tx_bytes = (p.dbg_mode & 4) ? W_TILE : W_TILE + X_TILE;
      uint32_t it = 0;  // k-blocks issued by this CTA so far, over all steps: ring position
      for (int t = 0; t < T; ++t) {
        int seg = 0, kl = kb_begin;  // segment of the current k-block and its index inside the segment
        while (seg + 1 < p.nseg && kb_begin >= p.seg_kb_end[seg]) ++seg;
        if (seg > 0) kl = kb_begin - p.seg_kb_end[seg - 1];
        bool met = !SEQ || t == 0;  // step 0 reads what was there before the launch
        for (int i = 0; i < num_kb; ++i, ++it) {
          const bool mine = it % PRODUCERS == pw;
          const int st = (int)(it % STAGES);
          const uint32_t phase = (it / STAGES) & 1u;
          if (mine && it >= (uint32_t)STAGES && lane < 5) mbar_wait(&empty[st], phase ^ 1);
          if (!mine) {
          } else if (p.dbg_mode & 2) {
            if (lane == 0) mbar_arrive(&full[st]);
          } else {
            if (lane == 0) mbar_expect_tx(&full[st], tx_bytes);
            const int32_t wk = kl * RK;
            uint8_t* const wt = tile_w(st);
            if (w_lane) {
              if (p.w_mn_major) tma_load_2d(&maps.w[my_map + seg], &full[st], wt + my_off, c0, wk);
              else tma_load_2d(&maps.w[my_map + seg], &full[st], wt + my_off, wk, c1);
            }
            if (x_lane) {
              if (SEQ && !met && ((sq.dep_segs >> seg) & 1u)) {
                // every CTA has published its rows of h_{t-1}: arrivals >= t * CTAs (the weight boxes of this stage, and of the
                // stages behind it, are already in flight)
                const unsigned target = (unsigned)t * sq.num_ctas;
                uint32_t spins = 0;
                while (ld_acquire_gpu(sq.bar) < target)
                  if (++spins > SPIN_LIMIT) __trap();
                asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores of other SMs -> this TMA read
                met = true;
              }
              tma_load_2d(SEQ ? &sq.xmaps[t * p.nseg + seg] : &maps.x[seg], &full[st], wt + W_TILE, wk, m0);
            }
          }
          if (t == 0 && i == 0 && lane == 0 && pw == 0) RNN_STAMP(2);
          ++kl;
          if (seg + 1 < p.nseg && kb_begin + i + 1 >= p.seg_kb_end[seg]) { ++seg; kl = 0; }
        }
        if (t == 0 && lane == 0 && pw == (uint32_t)((it - 1) % PRODUCERS) && num_kb > 0) RNN_STAMP(3);
        if (SEQ) {  // the cluster barrier of this step's exchange counts every thread of the cluster
          __syncwarp();
          if (t == 0) cluster_wait();  // #1
          cluster_arrive();
          cluster_wait();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the loop (uniform control flow keeps the descriptor arithmetic in uniform registers, which is where
    // UTCHMMA reads its operands from); one elected lane issues. Descriptors are (constant fields | address >> 4), so the loop
    // adds precomputed 16-byte-unit offsets instead of re-encoding them. Measured per MMA issued by a lone divergent thread:
    // ~83 cycles of issue against a 32-cycle execution floor (128 x 64 x 8).
    {
      const uint32_t idesc = make_idesc(p.w_mn_major);
      const uint32_t a_lbo = p.w_mn_major ? 4096 : 16, a_sbo = p.w_mn_major ? 512 : 1024, a_kstep = p.w_mn_major ? 1024 : 32;
      const uint32_t a_lt = p.w_mn_major ? 1 : 2;
      const uint64_t a_base = make_desc(smem_u32(smem), a_lbo, a_sbo, a_lt);            // weight tile of stage 0
      const uint64_t b_base = make_desc(smem_u32(smem) + W_TILE, 16, 1024, 2);           // activation tile of stage 0
      const uint64_t a_step = a_kstep >> 4, b_step = 32 >> 4, stage_step = STAGE_BYTES >> 4, lo_step = (W_TILE + X_TILE) >> 4;
      const bool issuer = lane == 0;
      int st = 0;
      uint32_t phase = 0;
      uint64_t da0 = a_base, db0 = b_base;
      for (int t = 0; t < T; ++t) {
        if (SEQ && t > 0) {  // the epilogue has read the previous step's accumulator out of TMEM
          mbar_wait(tmem_empty, (uint32_t)(t - 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int i = 0; i < num_kb; ++i) {
          mbar_wait(MODE == 2 ? &ready[st] : &full[st], phase);
          if (t == 0 && i == 0 && issuer) RNN_STAMP(4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (p.dbg_mode & 1) {
            if (issuer) mbar_arrive(&empty[st]);
          } else {
            const uint32_t acc = i != 0;
            if (issuer) {
#pragma unroll
              for (int k8 = 0; k8 < RK / 8; ++k8)  // k-step k8 accumulates into accumulator k8: consecutive MMAs are independent
                umma_tf32(tmem_base + (k8 % NACC) * RN, (MODE == 2 ? lo_step : 0) + da0 + k8 * a_step, db0 + k8 * b_step, idesc, acc | (uint32_t)(k8 >= NACC));  // 3xTF32: small terms first
              if (MODE == 2) {
#pragma unroll
                for (int k8 = 0; k8 < RK / 8; ++k8) umma_tf32(tmem_base + (k8 % NACC) * RN, da0 + k8 * a_step, db0 + k8 * b_step + lo_step, idesc, 1);
#pragma unroll
                for (int k8 = 0; k8 < RK / 8; ++k8) umma_tf32(tmem_base + (k8 % NACC) * RN, da0 + k8 * a_step, db0 + k8 * b_step, idesc, 1);
              }
              umma_commit(&empty[st]);
            }
            __syncwarp();
          }
          da0 += stage_step;
          db0 += stage_step;
          if (++st == STAGES) { st = 0; phase ^= 1; da0 = a_base; db0 = b_base; }
        }
        if (issuer) {
          if (p.dbg_mode & 1) mbar_arrive(tmem_full); else umma_commit(tmem_full);
          if (t == 0) RNN_STAMP(5);
        }
        if (SEQ) {
          __syncwarp();
          if (t == 0) cluster_wait();  // #1
          cluster_arrive();
          cluster_wait();
        }
      }
    }
    __syncwarp();
  } else {
    // ================= 3xTF32 converters (warps 2..5), then the epilogue of the step =================
    const int ct = threadIdx.x - 64;  // 0..127
    const uint32_t red_s = smem_u32(red), cell_s = smem_u32(cellbuf);
    uint32_t it = 0;
    for (int t = 0; t < T; ++t) {
      if (MODE == 2) {
        for (int i = 0; i < num_kb; ++i, ++it) {
          const int st = (int)(it % STAGES);
          const uint32_t round = it / STAGES;
          mbar_wait(&full[st], round & 1);
          // the landed fp32 words are the hi operand as they are (the tensor core reads their top 19 bits); lo = x - hi
          const uint4* src = reinterpret_cast<const uint4*>(tile_w(st));
          uint4* dlo = reinterpret_cast<uint4*>(tile_w(st) + W_TILE + X_TILE);
          // timing experiments (results are wrong): 8 = no conversion at all, 16 = only the activation tile is converted
          const int e_begin = (p.dbg_mode & 8) ? (W_TILE + X_TILE) / 16 : (p.dbg_mode & 16) ? W_TILE / 16 : 0;
#pragma unroll 4
          for (int e = e_begin + ct; e < (W_TILE + X_TILE) / 16; e += 128) {
            const uint4 v = src[e];
            uint4 lo;
            lo.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
            lo.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
            lo.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
            lo.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
            dlo[e] = lo;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[st]);
        }
      }
      // ---- exchange of the partial accumulators + final sum (see rnn_epilogue)
      if (t == 0) cluster_wait();  // #1: every CTA of the cluster has started
      RnnStep sp;
      if (SEQ) {
        sp = sq.steps[t];
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) sp.out[g] = p.out[g];
        sp.c_prev = p.c_prev; sp.c_out = p.c_out; sp.h_out = p.h_out;
      }
      uint64_t* const te = SEQ ? tmem_empty : nullptr;
      const uint32_t par = (uint32_t)t & 1u;
      switch (csize) {
        case 1: rnn_epilogue<64>(p, sp, tmem_full, par, te, tmem_base, red_s, cell_s, num_kb, u0, m0, crank); break;
        case 2: rnn_epilogue<32>(p, sp, tmem_full, par, te, tmem_base, red_s, cell_s, num_kb, u0, m0, crank); break;
        case 4: rnn_epilogue<16>(p, sp, tmem_full, par, te, tmem_base, red_s, cell_s, num_kb, u0, m0, crank); break;
        case 8: rnn_epilogue<8>(p, sp, tmem_full, par, te, tmem_base, red_s, cell_s, num_kb, u0, m0, crank); break;
        default: rnn_epilogue<4>(p, sp, tmem_full, par, te, tmem_base, red_s, cell_s, num_kb, u0, m0, crank); break;
      }
      if (SEQ && t + 1 < T) {
        // this CTA's rows of h_t (and c_t, the gate activations) are stored: publish them and arrive at the grid barrier
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          __threadfence();
          atomicAdd(sq.bar, 1u);
        }
      }
    }
  }
  if (!SEQ && (warp < 2 || warp >= 6)) {
    cluster_wait();    // #1
    cluster_arrive();  // #2
    cluster_wait();
  }
  if (threadIdx.x == 64) RNN_STAMP(9);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(RN * NACC));
  if (SEQ && threadIdx.x == 0) {
    // the last CTA to finish leaves both counters at zero for the next launch (everybody has passed every barrier by then)
    __threadfence();
    if (atomicAdd(sq.bar + 1, 1u) == sq.num_ctas - 1) {
      sq.bar[0] = 0;
      sq.bar[1] = 0;
      __threadfence();
    }
  }
  if (threadIdx.x == 0) RNN_STAMP(10);
}

long long* g_rnn_dbg = nullptr;

int check_desc(const tcr_gemm_group_desc* d, bool need_device_ptrs) {
  TCR_ARG(d != nullptr, "tcr_gemm_grouped: null descriptor");
  TCR_ARG(d->m >= 1 && d->n >= 1, "tcr_gemm_grouped: empty output %lld x %lld", (long long)d->m, (long long)d->n);
  TCR_ARG(d->groups == 1 || d->groups == 2 || d->groups == 4, "tcr_gemm_grouped: groups must be 1, 2 or 4 (got %d)", d->groups);
  TCR_ARG(d->segments >= 1 && d->segments <= 4, "tcr_gemm_grouped: 1..4 K-segments (got %d)", d->segments);
  TCR_ARG(d->groups * d->segments <= 8, "tcr_gemm_grouped: groups x segments <= 8");
  TCR_ARG(d->precision == TCR_GEMM_TF32 || d->precision == TCR_GEMM_3XTF32, "tcr_gemm_grouped: tensor-core precisions only");
  TCR_ARG(d->b_pitch > 0 && d->b_pitch % 4 == 0, "tcr_gemm_grouped: B row pitch must be a multiple of 4 elements (TMA)");
  TCR_ARG(d->out_pitch >= d->n, "tcr_gemm_grouped: output pitch smaller than n");
  for (int s = 0; s < d->segments; ++s) {
    TCR_ARG(d->seg_k[s] >= 1, "tcr_gemm_grouped: empty K-segment %d", s);
    TCR_ARG(d->a_pitch[s] >= d->seg_k[s] && d->a_pitch[s] % 4 == 0, "tcr_gemm_grouped: A pitch of segment %d must be >= K and a multiple of 4", s);
    if (need_device_ptrs) {
      TCR_ARG(d->a[s] != nullptr && (((uintptr_t)d->a[s]) & 15) == 0, "tcr_gemm_grouped: A segment %d must be 16-byte aligned", s);
      for (int g = 0; g < d->groups; ++g)
        TCR_ARG(d->b[g][s] != nullptr && (((uintptr_t)d->b[g][s]) & 15) == 0, "tcr_gemm_grouped: B[%d][%d] must be 16-byte aligned", g, s);
    }
  }
  if (d->cell) {
    TCR_ARG(d->groups == 4, "tcr_gemm_grouped: the LSTM cell epilogue needs the four gates in one launch");
    const int roles = (1 << d->role_cand) | (1 << d->role_in) | (1 << d->role_forget) | (1 << d->role_out);
    TCR_ARG(roles == 15, "tcr_gemm_grouped: gate roles must be a permutation of 0..3");
    TCR_ARG(d->state_pitch >= d->n, "tcr_gemm_grouped: state pitch smaller than n");
    if (need_device_ptrs) TCR_ARG(d->c_out != nullptr && d->h_out != nullptr, "tcr_gemm_grouped: cell outputs missing");
  } else if (need_device_ptrs) {
    for (int g = 0; g < d->groups; ++g) TCR_ARG(d->out[g] != nullptr, "tcr_gemm_grouped: output %d missing", g);
  }
  return TCR_OK;
}

template <int MODE, bool SEQ>
int launch_rnn(const RnnMaps& maps, const RnnParams& p, const RnnSeq& sq, dim3 grid, int cluster, bool check_residency = false) {
  static bool configured = false;
  if (!configured) {
    TCR_CUDA(cudaFuncSetAttribute(gemm_rnn_kernel<MODE, SEQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, RnnCfg<MODE>::SMEM));
    TCR_CUDA(cudaFuncSetAttribute(gemm_rnn_kernel<MODE, SEQ>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = RnnCfg<MODE>::SMEM;
  cfg.stream = state().stream;
  cudaLaunchAttribute attr[2];
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)cluster;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (check_residency) {
    // the grid barrier of the multi-step kernel needs every cluster on the machine at once
    int max_clusters = 0;
    cfg.numAttrs = 1;
    TCR_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, gemm_rnn_kernel<MODE, SEQ>, &cfg));
    const int64_t need = (int64_t)grid.x * grid.y;
    TCR_ARG(max_clusters >= need, "tcr_gemm_grouped_seq: %lld clusters of %d CTAs do not fit the device at once (%d)", (long long)need, cluster, max_clusters);
    return TCR_OK;
  }
  TCR_CUDA(cudaLaunchKernelEx(&cfg, gemm_rnn_kernel<MODE, SEQ>, maps, p, sq));
  state().launches.fetch_add(1, std::memory_order_relaxed);
  return TCR_OK;
}

// everything of a launch that does not change from time step to time step
int build_common(const tcr_gemm_group_desc* d, RnnMaps& maps, RnnParams& p, dim3& grid, int& cluster, bool with_x_maps) {
  std::memset(&p, 0, sizeof(p));
  p.m = d->m;
  p.n = d->n;
  p.groups = d->groups;
  p.rows_per_group = RM / d->groups;
  p.nseg = d->segments;
  p.w_mn_major = d->b_trans ? 0 : 1;
  int kb = 0, rc = TCR_OK;
  for (int s = 0; s < d->segments; ++s) {
    kb += (int)ceil_div(d->seg_k[s], RK);
    p.seg_kb_end[s] = kb;
    p.w_k0[s] = 0;
    if (with_x_maps) {
      // activations: dims {K_s, m}, box {32 k, 64 rows}; rows beyond m are zero-filled by the TMA unit
      rc = make_tf32_map(&maps.x[s], (const float*)d->a[s], d->seg_k[s], d->m, d->a_pitch[s], RK, RN, false);
      if (rc) return rc;
    }
    for (int g = 0; g < d->groups; ++g) {
      CUtensorMap* wm = &maps.w[g * d->segments + s];
      if (p.w_mn_major) rc = make_tf32_map(wm, (const float*)d->b[g][s], d->n, d->seg_k[s], d->b_pitch, 32, RK, true);  // [K_s][n], n contiguous
      else rc = make_tf32_map(wm, (const float*)d->b[g][s], d->seg_k[s], d->n, d->b_pitch, RK, (uint32_t)p.rows_per_group, false);  // [n][K_s], k contiguous
      if (rc) return rc;
    }
  }
  p.kb_total = kb;
  for (int g = 0; g < 4; ++g) {
    p.out[g] = g < d->groups ? (float*)d->out[g] : nullptr;
    p.bias[g] = g < d->groups ? (const float*)d->bias[g] : nullptr;
    p.act[g] = g < d->groups ? d->act[g] : 0;
  }
  p.out_pitch = d->out_pitch;
  p.accumulate = d->accumulate;
  p.cell = d->cell;
  p.role_cand = d->role_cand; p.role_in = d->role_in; p.role_forget = d->role_forget; p.role_out = d->role_out;
  p.c_prev = (const float*)d->c_prev;
  p.c_out = (float*)d->c_out;
  p.h_out = (float*)d->h_out;
  p.state_pitch = d->state_pitch;
  const int64_t tiles = ceil_div(d->n, p.rows_per_group) * ceil_div(d->m, RN);
  TCR_ARG(tiles <= 65535, "tcr_gemm_grouped: output too large for this kernel (%lld tiles)", (long long)tiles);
  // cluster size = split-K factor: as many CTAs as fit one wave, at least two k-blocks each
  static const int forced = std::getenv("TCR_RNN_CLUSTER") ? std::atoi(std::getenv("TCR_RNN_CLUSTER")) : 0;
  const int sms = state().sm_count;
  cluster = 1;
  for (int c = 2; c <= 8; c *= 2)  // 16 (non-portable) measured slower than 8 on the 64 x 1024 x 4096 sum (20.5 vs 15.6 us)
    if (tiles * c <= sms && kb / c >= 2) cluster = c;
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) cluster = forced;
  while (cluster > 1 && kb < cluster) cluster /= 2;
  p.kb_per_cta = (int)ceil_div(kb, cluster);
  grid = dim3((unsigned)ceil_div(d->n, p.rows_per_group), (unsigned)ceil_div(d->m, RN), (unsigned)cluster);
  return TCR_OK;
}

struct SeqHandle {
  RnnMaps maps;
  RnnParams p;
  RnnSeq sq;
  dim3 grid;
  int cluster = 1, precision = 0;
  void* dev = nullptr;  // one allocation: [x maps][steps][2 counters]
};

}  // namespace

}  // namespace tcr

using namespace tcr;

extern "C" {

int tcr_gemm_grouped_check(const tcr_gemm_group_desc* d) { return check_desc(d, false); }

/* TCR_RNN_DEBUG=1: SM-clock stamps of CTA (0,0,0) of the last tcr_gemm_grouped launch (profiling aid, tools/rnn_gemm_bench.py):
 * 0 entry, 1 set-up done, 2 first TMA issued, 3 last TMA issued, 4 first stage landed, 5 last MMA committed, 6 accumulator complete,
 * 7 partial sums sent, 8 cluster exchange complete, 9 outputs stored, 10 exit */
int tcr_rnn_debug_read(long long out[16]) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(g_rnn_dbg != nullptr, "tcr_rnn_debug_read: run with TCR_RNN_DEBUG=1");
  TCR_CUDA(cudaStreamSynchronize(state().stream));
  TCR_CUDA(cudaMemcpy(out, g_rnn_dbg, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return TCR_OK;
}

int tcr_gemm_grouped(const tcr_gemm_group_desc* d) {
  TCR_REQUIRE_DEVICE();
  int rc = check_desc(d, true);
  if (rc) return rc;
  RnnMaps maps;
  RnnParams p;
  dim3 grid;
  int cluster = 1;
  rc = build_common(d, maps, p, grid, cluster, true);
  if (rc) return rc;
  static const bool debug = std::getenv("TCR_RNN_DEBUG") != nullptr;
  if (debug && g_rnn_dbg == nullptr) {
    TCR_CUDA(cudaMalloc(&g_rnn_dbg, 16 * sizeof(long long)));
    TCR_CUDA(cudaMemset(g_rnn_dbg, 0, 16 * sizeof(long long)));
  }
  p.dbg = debug ? g_rnn_dbg : nullptr;
  static const int experiment = std::getenv("TCR_RNN_EXPERIMENT") ? std::atoi(std::getenv("TCR_RNN_EXPERIMENT")) : 0;
  p.dbg_mode = experiment;
  RnnSeq none;
  std::memset(&none, 0, sizeof(none));
  return d->precision == TCR_GEMM_TF32 ? launch_rnn<1, false>(maps, p, none, grid, cluster) : launch_rnn<2, false>(maps, p, none, grid, cluster);
}

int tcr_gemm_grouped_seq_prepare(const tcr_gemm_group_desc* descs, int count, void** handle) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(descs != nullptr && handle != nullptr && count >= 2 && count <= 4096, "tcr_gemm_grouped_seq_prepare: 2..4096 steps");
  *handle = nullptr;
  const tcr_gemm_group_desc& d0 = descs[0];
  uint32_t dep = 0;
  for (int t = 0; t < count; ++t) {
    const tcr_gemm_group_desc& d = descs[t];
    int rc = check_desc(&d, true);
    if (rc) return rc;
    TCR_ARG(d.cell == 1 && d.accumulate == 0, "tcr_gemm_grouped_seq_prepare: steps carry the cell epilogue and do not accumulate");
    bool same = d.m == d0.m && d.n == d0.n && d.groups == d0.groups && d.segments == d0.segments && d.b_pitch == d0.b_pitch && d.b_trans == d0.b_trans &&
                d.precision == d0.precision && d.out_pitch == d0.out_pitch && d.state_pitch == d0.state_pitch && d.role_cand == d0.role_cand &&
                d.role_in == d0.role_in && d.role_forget == d0.role_forget && d.role_out == d0.role_out;
    for (int s = 0; s < d0.segments && same; ++s) {
      same = d.seg_k[s] == d0.seg_k[s] && d.a_pitch[s] == d0.a_pitch[s];
      for (int g = 0; g < d0.groups && same; ++g) same = d.b[g][s] == d0.b[g][s];
    }
    for (int g = 0; g < d0.groups && same; ++g) same = d.bias[g] == d0.bias[g] && d.act[g] == d0.act[g];
    TCR_ARG(same, "tcr_gemm_grouped_seq_prepare: step %d differs from step 0 in more than its activations and outputs", t);
    if (t > 0) {
      // which segments read the previous step's h; the state must be the previous step's c (the same thread re-reads what it stored)
      uint32_t mine = 0;
      for (int s = 0; s < d.segments; ++s)
        if (d.a[s] == descs[t - 1].h_out) mine |= 1u << s;
      TCR_ARG(t == 1 || mine == dep, "tcr_gemm_grouped_seq_prepare: step %d chains through other segments than step 1", t);
      dep = mine;
      TCR_ARG(d.c_prev == descs[t - 1].c_out, "tcr_gemm_grouped_seq_prepare: step %d does not continue the state of step %d", t, t - 1);
      for (int s = 0; s < d.segments; ++s)
        for (int q = 0; q < t - 1; ++q)
          TCR_ARG(d.a[s] != descs[q].h_out && d.a[s] != descs[q].c_out, "tcr_gemm_grouped_seq_prepare: step %d reads the output of step %d", t, q);
    }
  }
  TCR_ARG(dep != 0, "tcr_gemm_grouped_seq_prepare: the steps are not chained");
  SeqHandle* h = new SeqHandle();
  int rc = build_common(&d0, h->maps, h->p, h->grid, h->cluster, false);
  if (rc) { delete h; return rc; }
  h->precision = d0.precision;
  const size_t maps_bytes = sizeof(CUtensorMap) * (size_t)count * d0.segments, steps_bytes = sizeof(RnnStep) * (size_t)count;
  std::vector<CUtensorMap> xm((size_t)count * d0.segments);
  std::vector<RnnStep> st((size_t)count);
  for (int t = 0; t < count; ++t) {
    const tcr_gemm_group_desc& d = descs[t];
    for (int s = 0; s < d.segments; ++s) {
      rc = make_tf32_map(&xm[(size_t)t * d.segments + s], (const float*)d.a[s], d.seg_k[s], d.m, d.a_pitch[s], RK, RN, false);
      if (rc) { delete h; return rc; }
    }
    for (int g = 0; g < 4; ++g) st[t].out[g] = g < d.groups ? (float*)d.out[g] : nullptr;
    st[t].c_prev = (const float*)d.c_prev;
    st[t].c_out = (float*)d.c_out;
    st[t].h_out = (float*)d.h_out;
  }
  char* dev = nullptr;
  if (cudaMalloc(&dev, maps_bytes + steps_bytes + 256) != cudaSuccess) { delete h; set_error("tcr_gemm_grouped_seq_prepare: out of device memory"); return TCR_ERR_CUDA; }
  h->dev = dev;
  cudaMemcpy(dev, xm.data(), maps_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dev + maps_bytes, st.data(), steps_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dev + maps_bytes + steps_bytes, 0, 256);
  h->sq.xmaps = reinterpret_cast<const CUtensorMap*>(dev);
  h->sq.steps = reinterpret_cast<const RnnStep*>(dev + maps_bytes);
  h->sq.T = count;
  h->sq.dep_segs = dep;
  h->sq.bar = reinterpret_cast<unsigned*>(dev + maps_bytes + steps_bytes);
  h->sq.num_ctas = h->grid.x * h->grid.y * h->grid.z;
  rc = h->precision == TCR_GEMM_TF32 ? launch_rnn<1, true>(h->maps, h->p, h->sq, h->grid, h->cluster, true)
                                     : launch_rnn<2, true>(h->maps, h->p, h->sq, h->grid, h->cluster, true);
  if (rc) { cudaFree(dev); delete h; return rc; }
  *handle = h;
  return TCR_OK;
}

int tcr_gemm_grouped_seq_launch(void* handle) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(handle != nullptr, "tcr_gemm_grouped_seq_launch: null handle");
  SeqHandle* h = (SeqHandle*)handle;
  return h->precision == TCR_GEMM_TF32 ? launch_rnn<1, true>(h->maps, h->p, h->sq, h->grid, h->cluster) : launch_rnn<2, true>(h->maps, h->p, h->sq, h->grid, h->cluster);
}

int tcr_gemm_grouped_seq_destroy(void* handle) {
  if (handle == nullptr) return TCR_OK;
  SeqHandle* h = (SeqHandle*)handle;
  if (h->dev) cudaFree(h->dev);
  delete h;
  return TCR_OK;
}

}  // extern "C"
