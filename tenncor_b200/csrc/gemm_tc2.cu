// gemm_tc2.cu — persistent CTA-pair fp32 GEMM on tcgen05 (cta_group::2), TF32 and 3xTF32.
//
// Why a second kernel: ncu on the single-CTA 128x128 kernel (profiles/r1_ncu_gemm_tc_tf32_4096.md)
// shows tensor pipe 27 %, DRAM 14 %, L2 27 % — it is bound by shared-memory bandwidth (fp32
// operands: per k-block the MMA reads 32 KiB while TMA writes 32 KiB in the 256 cycles the MMAs
// take) and pays a fixed prologue/epilogue bubble per tile. This kernel
//   * pairs two SMs on one 256x256 output tile (tcgen05.mma.cta_group::2, M = 256, N = 256): each
//     CTA stages only its 128 rows of A and its 128 columns of B, halving smem traffic per flop;
//   * is persistent: one CTA pair per SM pair walks a static tile schedule, the accumulator is
//     double-buffered in TMEM (2 x 256 columns = all 512) so the epilogue of tile i overlaps the
//     main loop of tile i+1, and barrier setup / TMEM allocation happen once per launch;
//   * keeps the 3xTF32 hi/lo split in shared memory (converter warps write only `lo`; the tensor core
//     truncates the raw tile to its `hi` part), 3 MMAs per k-step;
//   * supports split-K work items (weight gradients: K = batch) with a deterministic second pass;
//   * stream-K when the tile count does not fill whole waves (8192 x 1024: 128 tiles on 74 pairs = 1.73 waves): the tiles x
//     k-blocks are dealt to the pairs as contiguous ranges of equal length, so a pair computes the last k-blocks of one tile, whole
//     tiles, and the first k-blocks of another. The pair holding a tile's FIRST k-blocks owns it; the next pair computes the rest
//     first thing in its walk, leaves the raw accumulator in a workspace slot and raises a flag per epilogue warp; the owner
//     reaches that tile last, adds the partial (own + partial, a fixed order) and runs the epilogue. Nobody waits on a chain.
//     (Cutting N instead was measured and rejected: with the lo-part conversion and both operand tiles re-read by three MMAs per
//     k-step the kernel sits at ~1500 of 2196 cycles per k-block of shared-memory traffic, so a 192-column piece costs 0.85 and
//     a 128-column piece 0.72 of a 256-column one, profiles/r2_tc2_nslice.md.)
//
// Warp roles (per CTA):  w0 TMA producer | w1 MMA issuer (leader CTA only) | w2 TMEM alloc + TF32
// forwarder | w3 idle | 3xTF32: w4-7 converters, w8-11 epilogue | TF32: w4-7 epilogue.
// Barrier protocol (s = smem stage, b = TMEM buffer):
//   full[s]   local, 1 arrival + 32 KiB tx : this CTA's A/B tiles landed
//   ready[s]  leader, 2 (TF32) or 8 (3x) arrivals from both CTAs : stage may be consumed by the MMA
//   empty[s]  local, tcgen05.commit multicast to both CTAs : stage may be refilled
//   tfull[b]  local, commit multicast : accumulator b complete
//   tempty[b] leader, 8 arrivals (4 epilogue warps x 2 CTAs) : accumulator b drained
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace tcr {

namespace {

constexpr int BM_CTA = 128, BN_CTA = 128, BN = 256, BK = 32;
constexpr int TILE_BYTES = 128 * BK * 4;  // 16 KiB per operand tile per CTA
constexpr uint32_t SPIN_LIMIT = 1u << 28;
constexpr int EPI_PITCH = 36;  // floats per staged row: 16-byte aligned, conflict-free for 128-bit accesses

struct Tc2Params {
  int64_t m, n, k;
  int64_t c_sm, c_sn;
  float* c;
  const float* bias;
  int epilogue, activation, accumulate;
  int a_mn_major, b_mn_major;
  int tiles_m, tiles_n, splits, kb_per_split;  // work item = (tile_m, tile_n, split); split-K partials go to c + split*m*n
  int streamk;                                 // 1: pairs walk equal ranges of tiles x k-blocks (splits == 1)
  float* sk_ws;                                // [pair][256][256] raw partial of the pair's head fragment
  int* sk_flags;                               // [pair][8]: one per (CTA, epilogue warp), zero between launches
  long long* dbg;                              // TCR_TC2_TRACE: [pair][8] globaltimer stamps (see gemm_tc2_dispatch)
};

__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC2_STAMP(slot) do { if (p.dbg != nullptr && leader && lane == 0) p.dbg[cluster_id * 8 + (slot)] = globaltimer_ns(); } while (0)

// first (tile, k-block) unit of pair c. No snapping to tile edges: moving a boundary by 2-3 k-blocks to avoid a short fragment
// costs more balance (3 us on the busiest pair at 8192 x 1024 x 784) than the fragment does
__host__ __device__ inline int64_t tc2_sk_boundary(int64_t T, int c, int C, int total_kb) {
  (void)total_kb;
  return T * c / C;
}

struct Tc2Piece {
  int64_t m0, n0;
  int kb0, nkb, split;
  int kind;  // 0 whole k range of its work item; 1 head fragment (contributor: partial to the workspace); 2 tail fragment (owner: adds the next pair's partial)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ uint32_t make_idesc(int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;   // c_format F32
  d |= 2u << 7;   // a_format TF32
  d |= 2u << 10;  // b_format TF32
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(BN >> 3) << 17;   // N = 256
  d |= (uint32_t)(256 >> 4) << 24;  // M = 256 across the CTA pair
  return d;
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit to the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ float act_f(int act, float x) {
  if (act == TCR_EW_SIGMOID) return 1.0f / (1.0f + expf(-x));
  if (act == TCR_EW_TANH) return tanhf(x);
  return x;
}

template <int MODE> struct Cfg {
  static constexpr int STAGES = MODE == 2 ? 3 : 6;
  static constexpr int STAGE_BYTES = (MODE == 2 ? 4 : 2) * TILE_BYTES;
  static constexpr int CONV_WARP0 = 4;                       // 3xTF32 converters: warps 4..7
  static constexpr int EPI_WARP0 = MODE == 2 ? 8 : 4;        // epilogue: 4 warps
  static constexpr int THREADS = (EPI_WARP0 + 4) * 32;
  static constexpr int READY_COUNT = MODE == 2 ? 8 : 2;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 512 + 4 * 32 * EPI_PITCH * 4;
};

template <int MODE>
__global__ void __launch_bounds__(Cfg<MODE>::THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ Tc2Params p) {
  using C = Cfg<MODE>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* ready = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* tfull = bars + 3 * STAGES;
  uint64_t* tempty = bars + 3 * STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 4);
  float* epi_stage = (float*)(smem + STAGES * C::STAGE_BYTES + 512);  // 4 warps x 32 rows x EPI_PITCH

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int total_kb = (int)((p.k + BK - 1) / BK);
  const int num_items = p.tiles_m * p.tiles_n * p.splits;

  auto tile_a = [&](int s) { return smem + s * C::STAGE_BYTES; };
  auto tile_b = [&](int s) { return smem + s * C::STAGE_BYTES + TILE_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], C::READY_COUNT);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();  // the peer's barriers exist before anyone arrives on them remotely
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  TCR_PDL_ENTER();  // everything above touched no global memory: the previous kernel may still be running

  // the pair's walk over its pieces, identical in every warp role
  const int64_t sk_total = (int64_t)p.tiles_m * p.tiles_n * total_kb;
  // stream-K: the range [sk_begin, sk_end) of (tile, k-block) units is walked as  head fragment, tail fragment, whole tiles.  The
  // tail (owned) comes second so that its epilogue — which waits for the next pair's head — hides under the whole tiles, and
  // the un-overlapped last epilogue is an ordinary one.
  const int64_t sk_begin = p.streamk ? tc2_sk_boundary(sk_total, cluster_id, num_clusters, total_kb) : 0;
  const int64_t sk_end = p.streamk ? tc2_sk_boundary(sk_total, cluster_id + 1, num_clusters, total_kb) : 0;
  const int64_t sk_first_full = (sk_begin + total_kb - 1) / total_kb, sk_full_end = sk_end / total_kb;  // tile indices
  const int sk_head = sk_begin < sk_first_full * total_kb ? 1 : 0, sk_tail = sk_full_end * total_kb < sk_end ? 1 : 0;
  const int64_t walk_begin = p.streamk ? 0 : cluster_id;
  const int64_t walk_end = p.streamk ? sk_head + sk_tail + (sk_full_end - sk_first_full) : num_items;
  auto next_piece = [&](int64_t& cur, Tc2Piece& pc) -> bool {
    if (cur >= walk_end) return false;
    int t;
    if (!p.streamk) {
      const int item = (int)cur;
      cur += num_clusters;
      pc.split = item % p.splits;
      t = item / p.splits;
      pc.kb0 = pc.split * p.kb_per_split;
      pc.nkb = min(p.kb_per_split, total_kb - pc.kb0);
      pc.kind = 0;
    } else {
      const int j = (int)cur++;
      pc.split = 0;
      if (j < sk_head) {
        t = (int)(sk_begin / total_kb);
        pc.kb0 = (int)(sk_begin - (int64_t)t * total_kb);
        pc.nkb = total_kb - pc.kb0;
        pc.kind = 1;
      } else if (j < sk_head + sk_tail) {
        t = (int)sk_full_end;
        pc.kb0 = 0;
        pc.nkb = (int)(sk_end - sk_full_end * total_kb);
        pc.kind = 2;
      } else {
        t = (int)(sk_first_full + (j - sk_head - sk_tail));
        pc.kb0 = 0;
        pc.nkb = total_kb;
        pc.kind = 0;
      }
    }
    const int tn = t % p.tiles_n, tm = t / p.tiles_n;
    pc.m0 = (int64_t)tm * 256;
    pc.n0 = (int64_t)tn * 256;
    return true;
  };
  Tc2Piece pc;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    TC2_STAMP(0);
    if (lane == 0) {
      uint32_t it = 0;  // global k-block counter -> stage / phase
      for (int64_t cur = walk_begin; next_piece(cur, pc);) {
        const int kb0 = pc.kb0, nkb = pc.nkb;
        const int32_t am = (int32_t)(pc.m0 + 128 * rank), bn = (int32_t)(pc.n0 + 128 * rank);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t use = it / STAGES;
          mbar_wait(&empty[s], (use & 1) ^ 1);  // passes immediately on the first use of a stage
          mbar_expect_tx(&full[s], 2 * TILE_BYTES);
          const int32_t k0 = (kb0 + kb) * BK;
          if (!p.a_mn_major) tma_load_2d(&map_a, &full[s], tile_a(s), k0, am);
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(&map_a, &full[s], tile_a(s) + j * 4096, am + 32 * j, k0);
          }
          if (!p.b_mn_major) tma_load_2d(&map_b, &full[s], tile_b(s), k0, bn);
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(&map_b, &full[s], tile_b(s) + j * 4096, bn + 32 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA, one thread) =================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(p.a_mn_major, p.b_mn_major);
      const uint32_t a_lbo = p.a_mn_major ? 4096 : 16, a_sbo = p.a_mn_major ? 512 : 1024, a_kstep = p.a_mn_major ? 1024 : 32;
      const uint32_t b_lbo = p.b_mn_major ? 4096 : 16, b_sbo = p.b_mn_major ? 512 : 1024, b_kstep = p.b_mn_major ? 1024 : 32;
      const uint32_t a_lt = p.a_mn_major ? 1 : 2, b_lt = p.b_mn_major ? 1 : 2;
      uint32_t it = 0, tile_it = 0;
      for (int64_t cur = walk_begin; next_piece(cur, pc); ++tile_it) {
        const int nkb = pc.nkb;
        const uint32_t b = tile_it & 1, buse = tile_it >> 1;
        mbar_wait(&tempty[b], (buse & 1) ^ 1);  // epilogues of both CTAs drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + b * BN;
        uint32_t accumulate = 0;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t use = it / STAGES;
          mbar_wait(&ready[s], use & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(tile_a(s)), b_addr = smem_u32(tile_b(s));
#pragma unroll
          for (int k8 = 0; k8 < BK / 8; ++k8) {
            const uint64_t da = make_desc(a_addr + k8 * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t db = make_desc(b_addr + k8 * b_kstep, b_lbo, b_sbo, b_lt);
            if (MODE == 2) {
              const uint64_t da_lo = make_desc(a_addr + 2 * TILE_BYTES + k8 * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t db_lo = make_desc(b_addr + 2 * TILE_BYTES + k8 * b_kstep, b_lbo, b_sbo, b_lt);
              umma2_tf32(tmem_d, da_lo, db, idesc, accumulate);
              umma2_tf32(tmem_d, da, db_lo, idesc, 1);
              umma2_tf32(tmem_d, da, db, idesc, 1);
            } else {
              umma2_tf32(tmem_d, da, db, idesc, accumulate);
            }
            accumulate = 1;
          }
          umma2_commit(&empty[s]);
        }
        umma2_commit(&tfull[b]);
      }
      TC2_STAMP(1);  // last MMA issued
    }
  } else if (warp == 2 && MODE == 1) {
    // ================= TF32 forwarder: local "tiles landed" -> leader's ready barrier =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t cur = walk_begin; next_piece(cur, pc);) {
        const int nkb = pc.nkb;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          mbar_arrive_cluster(&ready[s], 0);
        }
      }
    }
  } else if (MODE == 2 && warp >= C::CONV_WARP0 && warp < C::CONV_WARP0 + 4) {
    // ================= 3xTF32 converters: lo = x - trunc_tf32(x) =================
    const int ct = threadIdx.x - C::CONV_WARP0 * 32;  // 0..127
    uint32_t it = 0;
    for (int64_t cur = walk_begin; next_piece(cur, pc);) {
      const int nkb = pc.nkb;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        const uint4* src = reinterpret_cast<const uint4*>(tile_a(s));
        uint4* dlo = reinterpret_cast<uint4*>(tile_a(s) + 2 * TILE_BYTES);
#pragma unroll 4
        for (int i = ct; i < 2 * TILE_BYTES / 16; i += 128) {
          const uint4 v = src[i];
          uint4 lo;
          lo.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
          lo.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
          lo.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
          lo.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
          dlo[i] = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&ready[s], 0);
      }
    }
  } else if (warp >= C::EPI_WARP0) {
    // ================= epilogue: TMEM -> registers -> global (both CTAs, own 128 rows) =================
    const int q = warp & 3;
    uint32_t tile_it = 0;
    float* const my_stage = epi_stage + (warp - C::EPI_WARP0) * (32 * EPI_PITCH);
    for (int64_t cur = walk_begin; next_piece(cur, pc); ++tile_it) {
      const int64_t m0 = pc.m0, n0 = pc.n0;
      const int split = pc.split;
      const uint32_t b = tile_it & 1, buse = tile_it >> 1;
      mbar_wait(&tfull[b], buse & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (pc.kind == 1) {
        // ---- stream-K contributor: this warp's 32 rows of the raw accumulator go to the pair's workspace slot, then its flag rises
        float* slot = p.sk_ws + (int64_t)cluster_id * (256 * 256) + (int64_t)(128 * rank + 32 * q) * 256;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          __syncwarp();
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * BN + c * 32);
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
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(my_stage + lane * EPI_PITCH + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
          __syncwarp();
          const int sub = lane >> 3, col = (lane & 7) * 4;
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) {
            const int row = rr + sub;
            __stcg(reinterpret_cast<float4*>(slot + row * 256 + c * 32 + col), *reinterpret_cast<const float4*>(my_stage + row * EPI_PITCH + col));
          }
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + cluster_id * 8 + 4 * (int)rank + q), "r"(1) : "memory");
        }
        if (q == 0) TC2_STAMP(2);  // head partial published
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tempty[b], 0);
        continue;
      }
      const float* partial = nullptr;
      if (pc.kind == 2) {
        // ---- stream-K owner: the next pair computed the tile's remaining k-blocks at the start of its walk
        int* flag = p.sk_flags + (cluster_id + 1) * 8 + 4 * (int)rank + q;
        if (q == 0) TC2_STAMP(3);  // owner: accumulator complete, waiting for the partial
        if (lane == 0) {
          int seen = 0;
          uint32_t spins = 0;
          while (true) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
            if (seen) break;
            if (++spins > SPIN_LIMIT) __trap();
          }
          *flag = 0;  // nobody else reads it: ready for the next launch
        }
        __syncwarp();
        if (q == 0) TC2_STAMP(4);
        partial = p.sk_ws + (int64_t)(cluster_id + 1) * (256 * 256) + (int64_t)(128 * rank + 32 * q) * 256;  // this warp's 32 rows
      }
      float* c_out = p.c + (p.splits > 1 ? (int64_t)split * p.m * p.n : 0);
      const int64_t m = m0 + 128 * rank + 32 * q + lane;
      const bool m_ok = m < p.m;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        __syncwarp();
        if (partial != nullptr) {
          // the other pair's partial of this chunk: four full 128-byte row segments per load instruction, through the staging tile
          const int sub = lane >> 3, col = (lane & 7) * 4;
          float4 pv[8];
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) pv[rr / 4] = __ldcg(reinterpret_cast<const float4*>(partial + (rr + sub) * 256 + c * 32 + col));
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) *reinterpret_cast<float4*>(my_stage + (rr + sub) * EPI_PITCH + col) = pv[rr / 4];
          __syncwarp();
        }
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * BN + c * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int64_t nb = n0 + c * 32;
        if (nb >= p.n) continue;  // warp-uniform
        const int64_t mw = m0 + 128 * rank + 32 * q;  // first row of this warp
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (partial != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t = *reinterpret_cast<const float4*>(my_stage + lane * EPI_PITCH + j);
            v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
          }
          __syncwarp();  // the staging tile is rewritten below
        }
        // fast path (warp-uniform): unit-stride, 16-byte aligned rows, whole 32-column chunk inside C.
        // Rows are exchanged through shared memory so every store instruction writes four full
        // 128-byte row segments instead of 32 scattered 16-byte pieces.
        const bool fast = p.c_sn == 1 && nb + 32 <= p.n && (p.c_sm & 3) == 0 && ((((uintptr_t)c_out) & 15) == 0) && ((nb & 3) == 0) &&
                          !(p.accumulate && (p.epilogue != TCR_EPI_NONE || p.activation));  // that mix keeps the scalar order: (x + C) + bias, then activation
        if (fast) {
          if (m_ok && (p.epilogue == TCR_EPI_BIAS_M || p.activation || p.epilogue == TCR_EPI_BIAS_N)) {
            const float bm = p.epilogue == TCR_EPI_BIAS_M ? p.bias[m] : 0.f;
            // the opcode tests sit outside the element loops: 32 independent chains per lane
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bm;
            if (p.epilogue == TCR_EPI_BIAS_N) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + nb);
              if ((((uintptr_t)b4) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 t = __ldg(b4 + j);
                  v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += p.bias[nb + j];
              }
            }
            if (p.activation == TCR_EW_SIGMOID) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.0f, 1.0f + expf(-v[j]));
            } else if (p.activation == TCR_EW_TANH) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
            }
          }
          float* stage = my_stage;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stage + lane * EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const int sub = lane >> 3, col = (lane & 7) * 4;
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) {
            const int row = rr + sub;
            const int64_t gm = mw + row;
            if (gm < p.m) {
              float4 x = *reinterpret_cast<const float4*>(stage + row * EPI_PITCH + col);
              float* dst = c_out + gm * p.c_sm + nb + col;
              if (p.accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w;
              }
              *reinterpret_cast<float4*>(dst) = x;
            }
          }
        } else if (m_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int64_t n = nb + j;
            if (n < p.n) {
              float x = v[j];
              float* dst = c_out + m * p.c_sm + n * p.c_sn;
              if (p.accumulate) x += *dst;
              if (p.epilogue == TCR_EPI_BIAS_N) x += p.bias[n];
              else if (p.epilogue == TCR_EPI_BIAS_M) x += p.bias[m];
              if (p.activation) x = act_f(p.activation, x);
              *dst = x;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[b], 0);  // this warp's quarter of accumulator b is free
    }
  }
  if (warp == C::EPI_WARP0) TC2_STAMP(5);  // epilogue done
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

__global__ void __launch_bounds__(256) splitk2_reduce_kernel(const float* __restrict__ ws, int splits, Tc2Params p) {
  TCR_PDL_ENTER();
  const int64_t total = p.m * p.n, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / p.n, n = i % p.n;
    float x = 0.f;
    for (int z = 0; z < splits; ++z) x += ws[(int64_t)z * total + i];
    float* dst = p.c + m * p.c_sm + n * p.c_sn;
    if (p.accumulate) x += *dst;
    if (p.epilogue == TCR_EPI_BIAS_N) x += p.bias[n];
    else if (p.epilogue == TCR_EPI_BIAS_M) x += p.bias[m];
    if (p.activation) x = act_f(p.activation, x);
    *dst = x;
  }
}

template <int MODE>
int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb, const Tc2Params& p, int num_clusters) {
  using C = Cfg<MODE>;
  static bool configured = false;
  if (!configured) {
    TCR_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * num_clusters);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = state().stream;
  cudaLaunchAttribute attr[2];
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  TCR_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<MODE>, ma, mb, p));
  state().launches.fetch_add(1, std::memory_order_relaxed);
  return TCR_OK;
}

}  // namespace

int make_tf32_map(CUtensorMap* map, const float* base, int64_t dim0, int64_t dim1, int64_t pitch, uint32_t box0, uint32_t box1, bool mn_major);  // gemm_tc.cu

// 2-CTA path: returns handled = false when the problem should go to the single-CTA kernel
int* counter_ring_take(int n);  // runtime.cu

int gemm_tc2_dispatch(const void* a, const void* b, void* c, const tcr_gemm_desc* d, int a_mn, int b_mn, int64_t a_pitch, int64_t b_pitch, bool* handled) {
  *handled = false;
  static const int enabled = std::getenv("TCR_GEMM_2CTA") ? std::atoi(std::getenv("TCR_GEMM_2CTA")) : 1;
  if (!enabled || d->batch != 1) return TCR_OK;
  if (d->m <= 128 || d->n <= 128) return TCR_OK;  // a 256x256 pair tile would be mostly padding
  if (d->k < 8 * BK) return TCR_OK;                // too few k-blocks to amortise the persistent prologue (measured: 1152x1024x64)
  CUtensorMap ma, mb;
  int rc = a_mn ? make_tf32_map(&ma, (const float*)a, d->m, d->k, a_pitch, 32, 32, true) : make_tf32_map(&ma, (const float*)a, d->k, d->m, a_pitch, 32, 128, false);
  if (rc) return rc;
  rc = b_mn ? make_tf32_map(&mb, (const float*)b, d->n, d->k, b_pitch, 32, 32, true) : make_tf32_map(&mb, (const float*)b, d->k, d->n, b_pitch, 32, 128, false);
  if (rc) return rc;
  Tc2Params p;
  p.m = d->m; p.n = d->n; p.k = d->k;
  p.c_sm = d->c_sm; p.c_sn = d->c_sn;
  p.c = (float*)c;
  p.bias = (const float*)d->bias;
  p.epilogue = d->epilogue; p.activation = d->activation; p.accumulate = d->accumulate;
  p.a_mn_major = a_mn; p.b_mn_major = b_mn;
  p.tiles_m = (int)ceil_div(d->m, 256);
  p.tiles_n = (int)ceil_div(d->n, 256);
  const int total_kb = (int)ceil_div(d->k, BK);
  const int pairs = state().sm_count / 2;
  static const int cap = std::getenv("TCR_TC2_CLUSTERS") ? std::atoi(std::getenv("TCR_TC2_CLUSTERS")) : 0;  // experiment knob
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  int splits = 1;
  if (tiles * 2 <= pairs && total_kb >= 16) {
    double best = (double)tiles / (double)(ceil_div(tiles, pairs) * pairs);
    for (int sp = 2; sp <= 64 && total_kb / sp >= 8; ++sp) {
      double eff = (double)(tiles * sp) / (double)(ceil_div(tiles * sp, pairs) * pairs);
      if (eff > best + 0.05) { best = eff; splits = sp; }
    }
  }
  p.kb_per_split = (int)ceil_div(total_kb, splits);
  splits = (int)ceil_div(total_kb, p.kb_per_split);
  p.splits = splits;
  p.streamk = 0;
  p.sk_ws = nullptr;
  p.sk_flags = nullptr;
  p.dbg = nullptr;
  static const bool trace = std::getenv("TCR_TC2_TRACE") != nullptr;  // eager launches only: the dispatch synchronises and prints
  static long long* trace_buf = nullptr;
  if (trace) {
    if (!trace_buf) TCR_CUDA(cudaMalloc(&trace_buf, sizeof(long long) * 8 * 80));
    TCR_CUDA(cudaMemsetAsync(trace_buf, 0, sizeof(long long) * 8 * 80, state().stream));
    p.dbg = trace_buf;
  }
  void* sk_ws = nullptr;
  static const int sk_enabled = std::getenv("TCR_TC2_STREAMK") ? std::atoi(std::getenv("TCR_TC2_STREAMK")) : 1;
  if (sk_enabled && splits == 1 && tiles > pairs && total_kb >= 8) {
    // whole waves of tiles cost ceil(tiles / pairs) tile-lengths; equal ranges cost tiles / pairs (+ a fragment's overheads).
    // Every range must be at least one tile long so that a tile is shared by two pairs at most. Measured (profiles/r2_tc2_streamk.md):
    // the gain is far below the MMA count saved on the busiest pair — 4096^3: -13.5 % k-blocks, -4 % time; 8192 x 1024 x 784:
    // -13.5 % k-blocks, +1 % time — because with every pair busy to the end the per-k-block time rises from ~0.95-1.0 to 1.08 us
    // (the trace shows all pairs resident from t = 0 and nobody waiting: the 3-MMA main loop is at the chip's sustained
    // tensor rate, not at a per-pair schedule limit). Hence the 12 % threshold: C3's first layer keeps the plain two-wave walk.
    const double share = (double)tiles * total_kb / pairs;
    const double waves = (double)ceil_div(tiles, (int64_t)pairs) * total_kb;
    if (share >= total_kb + 1 && waves > 1.12 * (share + 3) && (cap <= 0 || cap >= pairs)) {
      int* flags = counter_ring_take(8 * (pairs + 1));
      if (flags != nullptr) {
        rc = tcr_alloc(&sk_ws, sizeof(float) * 256 * 256 * (size_t)pairs);
        if (rc) return rc;
        p.streamk = 1;
        p.sk_ws = (float*)sk_ws;
        p.sk_flags = flags;
      }
    }
  }
  void* ws = nullptr;
  Tc2Params pk = p;
  if (splits > 1) {
    rc = tcr_alloc(&ws, sizeof(float) * (size_t)splits * d->m * d->n);
    if (rc) return rc;
    pk.c = (float*)ws; pk.c_sm = d->n; pk.c_sn = 1;
    pk.epilogue = TCR_EPI_NONE; pk.activation = 0; pk.accumulate = 0; pk.bias = nullptr;
  }
  const int64_t items = tiles * splits;
  int num_clusters = (int)(items < pairs ? items : pairs);
  if (cap > 0 && num_clusters > cap) num_clusters = cap;
  static const bool dbg = std::getenv("TCR_TC2_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr, "tc2: %lld x %lld x %lld %s tiles=%lld splits=%d kb/split=%d pairs=%d streamk=%d\n", (long long)d->m, (long long)d->n, (long long)d->k,
            d->precision == TCR_GEMM_TF32 ? "tf32" : "3xtf32", (long long)tiles, splits, p.kb_per_split, num_clusters, p.streamk);
  rc = d->precision == TCR_GEMM_TF32 ? launch_tc2<1>(ma, mb, pk, num_clusters) : launch_tc2<2>(ma, mb, pk, num_clusters);
  if (sk_ws) tcr_free(sk_ws);
  if (rc) return rc;
  if (trace) {
    static long long host[8 * 80];
    TCR_CUDA(cudaStreamSynchronize(state().stream));
    TCR_CUDA(cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost));
    long long t0 = 0;
    for (int c = 0; c < num_clusters; ++c)
      if (host[c * 8] && (!t0 || host[c * 8] < t0)) t0 = host[c * 8];
    fprintf(stderr, "tc2 trace (us since the first pair started): pair start lastMMA headPublished ownerWait ownerGo epilogueDone\n");
    for (int c = 0; c < num_clusters; ++c) {
      fprintf(stderr, "  %2d", c);
      for (int k = 0; k < 6; ++k) fprintf(stderr, " %8.2f", host[c * 8 + k] ? (host[c * 8 + k] - t0) * 1e-3 : -1.0);
      fprintf(stderr, "\n");
    }
  }
  if (splits > 1) {
    int grid = wave_grid(d->m * d->n, 256, 8);
    TCR_LAUNCH(splitk2_reduce_kernel, grid, 256, 0, (const float*)ws, splits, p);
    TCR_CHECK_LAUNCH();
    tcr_free(ws);
  }
  *handled = true;
  return TCR_OK;
}

}  // namespace tcr
