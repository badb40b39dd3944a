// gemm_tc.cu — fp32 GEMM on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// MATMUL / CONTRACT of the reference (internal/eigen/operator.hpp:1069-1139) for FLOAT
// operands, in two precisions (BASELINE.json north_star):
//   TCR_GEMM_TF32    one tcgen05.mma.kind::tf32 pass (operands truncated to 10-bit mantissa)
//   TCR_GEMM_3XTF32  a = hi + lo split, D += Ahi*Bhi + Ahi*Blo + Alo*Bhi: fp32-grade accuracy
//                    (relative error ~2^-21 per product) at 3 MMAs per k-step
//
// Structure (one 128x128 output tile per CTA, BLOCK_K = 32 fp32 = one 128-byte swizzle row):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d into a ring of smem stages (SWIZZLE_128B),
//               arming full[s] with the expected byte count
//   warp 1      allocates TMEM (128 fp32 columns); one elected lane issues tcgen05.mma from
//               shared-memory descriptors into the TMEM accumulator; tcgen05.commit frees the
//               stage (empty[s]) and finally signals the epilogue (tmem_full)
//   warps 2-5   3xTF32 only: split each landed tile into hi (in place) and lo (second buffer),
//               fence.proxy.async, arrive on ready[s]; then all four run the epilogue:
//               tcgen05.ld 32 lanes x 32 columns -> bias / activation -> global stores
// Both operand majors are supported through the UMMA descriptors (K-major and MN-major), so
// the NN / TN / NT / TT variants produced by the backward contractions
// (tenncor/eteq/backprop.hpp:269-359) need no transposes.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace tcr {

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int TILE_BYTES = BM * BK * 4;  // 16 KiB (A tile == B tile)
constexpr int NUM_THREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 28;  // bounded waits: a protocol bug traps instead of hanging the GPU
constexpr int EPI_PITCH = 36;  // floats per staged epilogue row: 16-byte aligned, conflict-free for 128-bit accesses

struct TcParams {
  int64_t m, n, k;
  int64_t c_sm, c_sn;
  float* c;
  const float* bias;
  int epilogue, activation, accumulate;
  int a_mn_major, b_mn_major;  // 0: K-major (K contiguous), 1: MN-major (M / N contiguous)
  int kb_per_split;            // k-blocks handled by one blockIdx.z (split-K); partial tiles go to `c` + z*m*n
  int raw_hi;                  // 3xTF32: leave the landed tile untouched (the tensor core truncates it to tf32 = hi) and only write lo
  // GATHER: A is the patch matrix of a conv2d that is never written to memory (tcr_gemm_patches). Row `pos` = valid window position
  // (x, y, image) of an image [C, W, H, images] (C fastest); column k = c + C * (kx + kw * ky): for a fixed ky the kw * C values
  // are contiguous in the image, so with (kw * C) % 32 == 0 every 32-k block of a row is 128 contiguous bytes
  const float* g_img;
  int g_c, g_w, g_h;           // image extents (elements)
  int g_pw, g_ph;              // valid positions along x / y
  int g_run;                   // kw * C: contiguous run of one window row (elements)
  // fused split-K reduction: the last CTA of a tile to finish (ticket counter) sums the partial tiles
  // in split order and applies the epilogue — deterministic, and no second launch
  int* counters;               // one per output tile, zero between launches (self-resetting); null = separate reduce kernel
  float* fin_c;
  int64_t fin_sm, fin_sn;
  const float* fin_bias;
  int fin_epilogue, fin_activation, fin_accumulate;
};

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

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address [0,14), leading
// byte offset [16,30), stride byte offset [32,46) — all >> 4 —, version = 1 at [46,48),
// layout type at [61,64): SWIZZLE_128B = 2 (K-major tiles), SWIZZLE_128B_BASE32B = 1 (the only
// swizzled layout the hardware accepts for MN-major tf32 operands; TMA writes it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc(int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                          // c_format = F32
  d |= 2u << 7;                          // a_format = TF32
  d |= 2u << 10;                         // b_format = TF32
  d |= (uint32_t)(a_mn_major & 1) << 15; // a_major
  d |= (uint32_t)(b_mn_major & 1) << 16; // b_major
  d |= (uint32_t)(BN >> 3) << 17;        // n_dim
  d |= (uint32_t)(BM >> 4) << 24;        // m_dim
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

__device__ __forceinline__ float act_f(int act, float x) {
  if (act == TCR_EW_SIGMOID) return 1.0f / (1.0f + expf(-x));
  if (act == TCR_EW_TANH) return tanhf(x);
  return x;
}

// MODE 1: TF32 (STAGES x 32 KiB), MODE 2: 3xTF32 (STAGES x 64 KiB)
template <int MODE, int STAGES, bool GATHER>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms need 1024-byte alignment
  constexpr int STAGE_BYTES = (MODE == 2 ? 4 : 2) * TILE_BYTES;
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* ready = bars + 2 * STAGES;
  uint64_t* tmem_full = bars + 3 * STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int total_kb = (int)((p.k + BK - 1) / BK);
  const int kb_begin = (int)blockIdx.z * p.kb_per_split;
  const int num_kb = min(p.kb_per_split, total_kb - kb_begin);  // >= 1 by construction of the grid
  float* const c_out = p.c + (gridDim.z > 1 ? (int64_t)blockIdx.z * p.m * p.n : 0);

  auto tile_a = [&](int s) { return smem + s * STAGE_BYTES; };
  auto tile_b = [&](int s) { return smem + s * STAGE_BYTES + TILE_BYTES; };
  auto tile_a_lo = [&](int s) { return smem + s * STAGE_BYTES + 2 * TILE_BYTES; };
  auto tile_b_lo = [&](int s) { return smem + s * STAGE_BYTES + 3 * TILE_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], GATHER ? 33 : 1);  // GATHER: + one arrival per lane of the producer warp when its cp.async copies have landed
      mbar_init(&empty[s], 1);
      mbar_init(&ready[s], 4);  // one arrival per converter warp
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  TCR_PDL_ENTER();  // everything above touched no global memory: the previous kernel may still be running

  if (GATHER && warp == 0) {
    // ================= producer: B tile by TMA, A tile gathered from the image =================
    // lane l owns rows l, l + 32, l + 64, l + 96 of the tile: their 128-byte k-block is copied with 8 x cp.async (16 bytes) into the
    // SWIZZLE_128B arrangement the TMA would have produced (16-byte chunk j of row r sits at chunk j ^ (r & 7)); rows beyond the
    // last position are zero-filled (src-size 0)
    const float* rowp[4];
    uint32_t rowok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t pos = m0 + lane + 32 * i;
      rowok[i] = pos < p.m ? 16u : 0u;
      const int64_t q = pos < p.m ? pos : 0;
      const int64_t x = q % p.g_pw, t = q / p.g_pw, y = t % p.g_ph, img = t / p.g_ph;
      rowp[i] = p.g_img + (int64_t)p.g_c * (x + (int64_t)p.g_w * (y + (int64_t)p.g_h * img));
    }
    const int64_t row_stride = (int64_t)p.g_w * p.g_c;  // one image row down
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t round = kb / STAGES;
      if (kb >= STAGES) mbar_wait(&empty[s], (round - 1) & 1);
      const int32_t k0 = (kb_begin + kb) * BK;
      if (lane == 0) {
        mbar_expect_tx(&full[s], TILE_BYTES);
        if (!p.b_mn_major) {
          tma_load_2d(&map_b, &full[s], tile_b(s), k0, (int32_t)n0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(&map_b, &full[s], tile_b(s) + j * 4096, (int32_t)n0 + 32 * j, k0);
        }
      }
      const int ky = k0 / p.g_run;
      const int64_t off = (int64_t)ky * row_stride + (k0 - ky * p.g_run);
      const uint32_t tile = smem_u32(tile_a(s));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t r = (uint32_t)lane + 32u * i;
        const float* src = rowp[i] + off;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
          const uint32_t dst = tile + r * 128u + ((j ^ (r & 7u)) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + 4 * j), "r"(rowok[i]) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t round = kb / STAGES;
        if (kb >= STAGES) mbar_wait(&empty[s], (round - 1) & 1);
        mbar_expect_tx(&full[s], 2 * TILE_BYTES);
        const int32_t k0 = (kb_begin + kb) * BK;
        if (!p.a_mn_major) {
          tma_load_2d(&map_a, &full[s], tile_a(s), k0, (int32_t)m0);  // [128 rows][32 k]
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(&map_a, &full[s], tile_a(s) + j * 4096, (int32_t)m0 + 32 * j, k0);  // 4 x [32 k][32 m]
        }
        if (!p.b_mn_major) {
          tma_load_2d(&map_b, &full[s], tile_b(s), k0, (int32_t)n0);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(&map_b, &full[s], tile_b(s) + j * 4096, (int32_t)n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.a_mn_major, p.b_mn_major);
      // K-major (SWIZZLE_128B): rows of 128 B, 8-row groups 1024 B apart (SBO); one MMA consumes
      //   8 k = 32 B of each row.
      // MN-major (SWIZZLE_128B_BASE32B): k-rows of 128 B holding 32 consecutive m/n; the swizzle atom is
      //   4 k-rows (512 B, SBO); 32-element chunks along M/N are 4096 B apart (LBO); one MMA consumes
      //   8 k-rows = 1024 B.
      const uint32_t a_lbo = p.a_mn_major ? 4096 : 16, a_sbo = p.a_mn_major ? 512 : 1024, a_kstep = p.a_mn_major ? 1024 : 32;
      const uint32_t b_lbo = p.b_mn_major ? 4096 : 16, b_sbo = p.b_mn_major ? 512 : 1024, b_kstep = p.b_mn_major ? 1024 : 32;
      const uint32_t a_lt = p.a_mn_major ? 1 : 2, b_lt = p.b_mn_major ? 1 : 2;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t round = kb / STAGES;
        mbar_wait(MODE == 2 ? &ready[s] : &full[s], round & 1);
        if (GATHER && MODE != 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async writes (generic proxy) -> MMA operand reads
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(tile_a(s)), b_addr = smem_u32(tile_b(s));
#pragma unroll
        for (int k8 = 0; k8 < BK / 8; ++k8) {
          const uint64_t da = make_desc(a_addr + k8 * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t db = make_desc(b_addr + k8 * b_kstep, b_lbo, b_sbo, b_lt);
          if (MODE == 2) {
            const uint64_t da_lo = make_desc(a_addr + 2 * TILE_BYTES + k8 * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t db_lo = make_desc(b_addr + 2 * TILE_BYTES + k8 * b_kstep, b_lbo, b_sbo, b_lt);
            umma_tf32(tmem_base, da_lo, db, idesc, accumulate);  // small terms first
            umma_tf32(tmem_base, da, db_lo, idesc, 1);
            umma_tf32(tmem_base, da, db, idesc, 1);
          } else {
            umma_tf32(tmem_base, da, db, idesc, accumulate);
          }
          accumulate = 1;
        }
        umma_commit(&empty[s]);  // stage is reusable once these MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    // ================= converter (3xTF32) + epilogue: warps 2..5 =================
    const int ct = threadIdx.x - 64;  // 0..127
    if (MODE == 2) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t round = kb / STAGES;
        mbar_wait(&full[s], round & 1);
        // hi/lo split is elementwise, so the swizzle is irrelevant: A and B tiles are adjacent (32 KiB)
        uint4* src = reinterpret_cast<uint4*>(tile_a(s));
        uint4* dlo = reinterpret_cast<uint4*>(tile_a_lo(s));
#pragma unroll 4
        for (int i = ct; i < 2 * TILE_BYTES / 16; i += 128) {
          uint4 v = src[i], hi, lo;
          const uint32_t* x = reinterpret_cast<const uint32_t*>(&v);
          uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
          uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t hb;
            if (p.raw_hi) hb = x[e] & 0xFFFFE000u;  // what the tensor core keeps of the raw fp32 word
            else asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(__uint_as_float(x[e])));
            h[e] = hb;
            l[e] = __float_as_uint(__uint_as_float(x[e]) - __uint_as_float(hb));
          }
          if (!p.raw_hi) src[i] = hi;
          dlo[i] = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
      }
    }
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int64_t m = m0 + 32 * q + lane;
    const bool m_ok = m < p.m;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the masked stores of the previous chunk
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(c * 32);
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
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      // fast path (warp-uniform): rows go through shared memory (the pipeline stages are idle once
      // tmem_full fired) so each store instruction writes four full 128-byte row segments
      const bool fast = p.c_sn == 1 && nb + 32 <= p.n && (p.c_sm & 3) == 0 && ((((uintptr_t)c_out) & 15) == 0) && ((nb & 3) == 0) &&
                        !(p.accumulate && (p.epilogue != TCR_EPI_NONE || p.activation));
      if (fast) {
        if (m_ok && (p.epilogue != TCR_EPI_NONE || p.activation)) {
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
        float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * EPI_PITCH);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * EPI_PITCH + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        const int sub = lane >> 3, col = (lane & 7) * 4;
        const int64_t mw = m0 + 32 * q;
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
  }
  if (gridDim.z > 1 && p.counters != nullptr) __threadfence();  // partial tile visible device-wide before the ticket
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  if (gridDim.z > 1 && p.counters != nullptr) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
      int* ctr = p.counters + blockIdx.y * gridDim.x + blockIdx.x;
      const int ticket = atomicAdd(ctr, 1);
      s_last = ticket == (int)gridDim.z - 1;
      if (s_last) *ctr = 0;  // every split has arrived: ready for the next launch that is handed this counter
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      const int64_t plane = p.m * p.n;
      const int splits = (int)gridDim.z;
      const bool vec = (p.n & 3) == 0;
      for (int idx = threadIdx.x; idx < BM * (BN / 4); idx += NUM_THREADS) {
        const int64_t m = m0 + idx / (BN / 4), n = n0 + (idx % (BN / 4)) * 4;
        if (m >= p.m || n >= p.n) continue;
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        const float* src = p.c + m * p.n + n;  // p.c = workspace, partial z at + z * plane
        if (vec) {  // n + 4 <= p.n follows from n % 4 == 0 and p.n % 4 == 0
          for (int z = 0; z < splits; ++z) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)z * plane));
            x[0] += t.x; x[1] += t.y; x[2] += t.z; x[3] += t.w;
          }
        } else {
          for (int z = 0; z < splits; ++z)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < p.n) x[j] += __ldcg(src + (int64_t)z * plane + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j >= p.n) continue;
          float v = x[j];
          float* dst = p.fin_c + m * p.fin_sm + (n + j) * p.fin_sn;
          if (p.fin_accumulate) v += *dst;
          if (p.fin_epilogue == TCR_EPI_BIAS_N) v += p.fin_bias[n + j];
          else if (p.fin_epilogue == TCR_EPI_BIAS_M) v += p.fin_bias[m];
          if (p.fin_activation) v = act_f(p.fin_activation, v);
          *dst = v;
        }
      }
    }
  }
}

// deterministic split-K: out(m,n) = epilogue(sum_z ws[z][m][n])
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, TcParams p) {
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so that the library loads on machines without libcuda.so
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

// 2-D fp32 tensor map: dim0 = contiguous extent, dim1 = strided extent (pitch in elements)
int make_map(CUtensorMap* map, const float* base, int64_t dim0, int64_t dim1, int64_t pitch, uint32_t box0, uint32_t box1, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return TCR_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for dims %lld x %lld pitch %lld", (int)r, (long long)dim0, (long long)dim1, (long long)pitch);
    return TCR_ERR_CUDA;
  }
  return TCR_OK;
}

template <int MODE, int STAGES, bool GATHER = false>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int splits) {
  constexpr int STAGE_BYTES = (MODE == 2 ? 4 : 2) * TILE_BYTES;
  constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static bool configured = false;
  if (!configured) {
    TCR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE, STAGES, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(p.n, BN), (unsigned)ceil_div(p.m, BM), (unsigned)splits);
  TCR_LAUNCH((gemm_tc_kernel<MODE, STAGES, GATHER>), grid, NUM_THREADS, SMEM, ma, mb, p);
  TCR_CHECK_LAUNCH();
  return TCR_OK;
}

}  // namespace

int make_tf32_map(CUtensorMap* map, const float* base, int64_t dim0, int64_t dim1, int64_t pitch, uint32_t box0, uint32_t box1, bool mn_major) {
  return make_map(map, base, dim0, dim1, pitch, box0, box1, mn_major);
}

int gemm_tc2_dispatch(const void* a, const void* b, void* c, const tcr_gemm_desc* d, int a_mn, int b_mn, int64_t a_pitch, int64_t b_pitch, bool* handled);  // gemm_tc2.cu
int* counter_ring_take(int n);  // runtime.cu

int gemm_tc_dispatch(const void* a, const void* b, void* c, const tcr_gemm_desc* d, bool* handled) {
  *handled = false;
  if (d->dtype != TCR_FLOAT || d->k <= 0) return TCR_OK;
  // tiny problems do not fill a 128x128 tile: the SIMT kernel is faster and exact
  // ... unless the reduction is long (conv kernel gradients: K = positions x batch): split-K over many CTAs then
  // gives the parallelism one SIMT tile cannot (measured: 27 x 64 x 262144 took 21 ms on the SIMT kernel)
  const bool deep = d->k >= 4096 && d->m * d->n >= 256;
  if ((d->m * d->n < 64 * 64 && !deep) || d->k < 16 || d->m * d->n * d->k < (1ll << 20)) return TCR_OK;
  if (d->batch > 16) return TCR_OK;
  const bool a_k = d->a_sk == 1 || d->k == 1, a_m = d->a_sm == 1 || d->m == 1;
  const bool b_k = d->b_sk == 1 || d->k == 1, b_n = d->b_sn == 1 || d->n == 1;
  if (!(a_k || a_m) || !(b_k || b_n)) return TCR_OK;
  const int a_mn = !(d->a_sk == 1) && a_m ? 1 : (a_k ? 0 : 1);
  const int b_mn = !(d->b_sk == 1) && b_n ? 1 : (b_k ? 0 : 1);
  const int64_t a_pitch = a_mn ? d->a_sk : d->a_sm, b_pitch = b_mn ? d->b_sk : d->b_sn;
  // TMA: 16-byte aligned base and pitch
  auto ok16 = [](const void* ptr, int64_t pitch) { return (((uintptr_t)ptr) & 15) == 0 && pitch > 0 && (pitch % 4) == 0; };
  if (!ok16(a, a_pitch) || !ok16(b, b_pitch)) return TCR_OK;
  if ((d->a_sb % 4) != 0 || (d->b_sb % 4) != 0) return TCR_OK;
  if (a_pitch < (a_mn ? d->m : d->k) || b_pitch < (b_mn ? d->n : d->k)) return TCR_OK;

  {
    // large outputs: persistent CTA-pair kernel (256x256 tiles, cta_group::2)
    int rc2 = gemm_tc2_dispatch(a, b, c, d, a_mn, b_mn, a_pitch, b_pitch, handled);
    if (rc2 || *handled) return rc2;
  }
  for (int64_t bi = 0; bi < d->batch; ++bi) {
    const float* ap = (const float*)a + bi * d->a_sb;
    const float* bp = (const float*)b + bi * d->b_sb;
    CUtensorMap ma, mb;
    int rc;
    // K-major: dims {K, rows}, box {32 k, 128 rows};  MN-major: dims {cols, K}, box {32 cols, 32 k}
    rc = a_mn ? make_map(&ma, ap, d->m, d->k, a_pitch, 32, 32, true) : make_map(&ma, ap, d->k, d->m, a_pitch, 32, 128, false);
    if (rc) return rc;
    rc = b_mn ? make_map(&mb, bp, d->n, d->k, b_pitch, 32, 32, true) : make_map(&mb, bp, d->k, d->n, b_pitch, 32, 128, false);
    if (rc) return rc;
    TcParams p = TcParams();
    p.m = d->m; p.n = d->n; p.k = d->k;
    p.c_sm = d->c_sm; p.c_sn = d->c_sn;
    p.c = (float*)c + bi * d->c_sb;
    p.bias = (const float*)d->bias;
    p.epilogue = d->epilogue; p.activation = d->activation; p.accumulate = d->accumulate;
    p.a_mn_major = a_mn; p.b_mn_major = b_mn;
    {
      // measured on B200: the tensor core truncates fp32 words to tf32, so the raw tile IS the hi part
      // (3xTF32 error unchanged: 5.7e-5 vs 4.9e-5 abs at K = 256); TCR_3X_RAWHI=0 restores the explicit rna split
      static const int raw_hi = std::getenv("TCR_3X_RAWHI") ? std::atoi(std::getenv("TCR_3X_RAWHI")) : 1;
      p.raw_hi = raw_hi;
    }
    // split-K when the output has fewer tiles than SMs (weight gradients: K = batch): pick the
    // split count whose CTA count best fills whole waves of the machine
    const int total_kb = (int)ceil_div(d->k, BK);
    const int64_t tiles = ceil_div(d->m, BM) * ceil_div(d->n, BN);
    const int sms = state().sm_count;
    int splits = 1;
    static const int min_kb = std::getenv("TCR_GEMM_SPLIT_MIN_KB") ? std::atoi(std::getenv("TCR_GEMM_SPLIT_MIN_KB")) : 8;  // experiment knob; 0 = never split
    if (min_kb > 0 && tiles * 2 <= sms && total_kb >= 16) {
      double best = (double)tiles / (double)(ceil_div(tiles, sms) * sms);
      const int sp_cap = total_kb >= sms * 16 ? sms : 32;  // very long reductions may use every SM for one output tile
      // The floor of 8 k-blocks per split stays even for small latency-bound outputs. Alone in a graph the LSTM's 64 x 1024 x 1152
      // gate product gets faster with more splits (13.1 us at >= 8 k-blocks, 9.7 us at >= 2, profiles/r1_sweep_c4_gemm.txt), but the
      // step runs four of them on concurrent lanes: 4 x 144 CTAs no longer fit one wave and the whole C4 step went from 15.7 ms to
      // 20.2 ms (>= 4 k-blocks: 16.4 ms). Measured with TCR_GEMM_SPLIT_MIN_KB, profiles/r1_bench_c4_split_floor.md.
      for (int sp = 2; sp <= sp_cap && total_kb / sp >= min_kb; ++sp) {
        double eff = (double)(tiles * sp) / (double)(ceil_div(tiles * sp, sms) * sms);
        if (eff > best + 0.05) { best = eff; splits = sp; }
      }
    }
    p.kb_per_split = (int)ceil_div(total_kb, splits);
    splits = (int)ceil_div(total_kb, p.kb_per_split);
    void* ws = nullptr;
    TcParams pk = p;
    pk.counters = nullptr;
    // measured in a CUDA graph on the LSTM gate product (64 x 1024 x 1152, 4 splits): fused 22.1 us vs separate reduce 13.0 us —
    // one CTA per tile summing the partials has too little parallelism, so the fused form stays opt-in
    static const int fused_reduce = std::getenv("TCR_GEMM_FUSED_REDUCE") ? std::atoi(std::getenv("TCR_GEMM_FUSED_REDUCE")) : 0;
    if (splits > 1) {
      rc = tcr_alloc(&ws, sizeof(float) * (size_t)splits * d->m * d->n);
      if (rc) return rc;
      pk.c = (float*)ws; pk.c_sm = d->n; pk.c_sn = 1;
      pk.epilogue = TCR_EPI_NONE; pk.activation = 0; pk.accumulate = 0; pk.bias = nullptr;
      if (fused_reduce) {
        pk.counters = counter_ring_take((int)tiles);
        pk.fin_c = p.c; pk.fin_sm = p.c_sm; pk.fin_sn = p.c_sn; pk.fin_bias = p.bias;
        pk.fin_epilogue = p.epilogue; pk.fin_activation = p.activation; pk.fin_accumulate = p.accumulate;
      }
    }
    // Few k-blocks per CTA and many CTAs (conv2d patch products, K = 9 x channels): the 6-stage ring never fills, and with 198 KB
    // of stages only one CTA fits an SM, so nothing hides its prologue and epilogue (profiles/r1_ncu_conv.md). In TF32 a 3-stage ring
    // (97 KB, two CTAs per SM) is 1.37x faster there (65536 x 64 x 288: 25.3 -> 18.5 us). The 3xTF32 ring is already 3 stages; a
    // single-stage variant that would fit two CTAs was slower everywhere (41.7 -> 46.8 us; LSTM gate product 13.9 -> 19.5 us;
    // profiles/r1_sweep_shortk.txt) and is not built. TCR_GEMM_SHORTK=0 disables the shallow ring.
    static const int shortk = std::getenv("TCR_GEMM_SHORTK") ? std::atoi(std::getenv("TCR_GEMM_SHORTK")) : 1;
    if (shortk && d->precision == TCR_GEMM_TF32 && p.kb_per_split <= 12 && tiles * splits >= 2 * sms)
      rc = launch_tc<1, 3>(ma, mb, pk, splits);
    else
    rc = d->precision == TCR_GEMM_TF32 ? launch_tc<1, 6>(ma, mb, pk, splits) : launch_tc<2, 3>(ma, mb, pk, splits);
    if (rc) return rc;
    if (splits > 1) {
      if (pk.counters == nullptr) {
        int grid = wave_grid(d->m * d->n, 256, 8);
        TCR_LAUNCH(splitk_reduce_kernel, grid, 256, 0, (const float*)ws, splits, p);
        TCR_CHECK_LAUNCH();
      }
      tcr_free(ws);
    }
  }
  *handled = true;
  return TCR_OK;
}

// The conv2d composite's forward product with the patch matrix gathered inside the kernel (see TcParams::g_*). Returns
// TCR_ERR_UNSUPPORTED (nothing launched) when the view or the operands do not fit: the caller keeps tcr_im2col + tcr_gemm.
int gemm_tc_patches(const void* image, const void* b, void* c, const tcr_gemm_desc* d, const int64_t img[8], const int64_t win[8]) {
  auto unsupported = [](const char* why) { set_error("tcr_gemm_patches: %s", why); return TCR_ERR_UNSUPPORTED; };
  if (d->dtype != TCR_FLOAT || d->batch != 1 || d->accumulate || d->post_op) return unsupported("fp32, one problem, no accumulation / post-op");
  if (d->precision != TCR_GEMM_TF32 && d->precision != TCR_GEMM_3XTF32) return unsupported("tensor-core precisions only");
  for (int r = 3; r < 8; ++r)
    if (win[r] != 1) return unsupported("the window covers channels, x and y only");
  const int64_t C = img[0], W = img[1], H = img[2], kw = win[1], kh = win[2];
  int64_t images = 1;
  for (int r = 3; r < 8; ++r) images *= img[r];
  if (win[0] != C || kw > W || kh > H) return unsupported("the window spans every channel");
  const int64_t run = kw * C, K = run * kh, pw = W - kw + 1, ph = H - kh + 1, M = pw * ph * images;
  if (run % BK != 0 || C % 4 != 0) return unsupported("kw * channels must be a multiple of 32");
  if (d->k != K || d->m != M) return unsupported("product extents do not match the patch view");
  if (W * C >= (1ll << 31) || M >= (1ll << 31) || (((uintptr_t)image) & 15) != 0) return unsupported("image too large or misaligned");
  const bool b_k = d->b_sk == 1 || d->k == 1, b_n = d->b_sn == 1 || d->n == 1;
  if (!(b_k || b_n)) return unsupported("kernel operand is strided in both ranks");
  const int b_mn = !(d->b_sk == 1) && b_n ? 1 : (b_k ? 0 : 1);
  const int64_t b_pitch = b_mn ? d->b_sk : d->b_sn;
  if ((((uintptr_t)b) & 15) != 0 || b_pitch <= 0 || (b_pitch % 4) != 0 || b_pitch < (b_mn ? d->n : d->k)) return unsupported("kernel operand not addressable by TMA");
  CUtensorMap ma, mb;
  std::memset(&ma, 0, sizeof(ma));
  int rc = b_mn ? make_map(&mb, (const float*)b, d->n, d->k, b_pitch, 32, 32, true) : make_map(&mb, (const float*)b, d->k, d->n, b_pitch, 32, 128, false);
  if (rc) return rc;
  TcParams p = TcParams();
  p.m = M; p.n = d->n; p.k = K;
  p.c_sm = d->c_sm; p.c_sn = d->c_sn;
  p.c = (float*)c;
  p.bias = (const float*)d->bias;
  p.epilogue = d->epilogue; p.activation = d->activation; p.accumulate = 0;
  p.a_mn_major = 0; p.b_mn_major = b_mn;
  p.raw_hi = 1;
  p.kb_per_split = (int)(K / BK);
  p.g_img = (const float*)image;
  p.g_c = (int)C; p.g_w = (int)W; p.g_h = (int)H; p.g_pw = (int)pw; p.g_ph = (int)ph; p.g_run = (int)run;
  return d->precision == TCR_GEMM_TF32 ? launch_tc<1, 3, true>(ma, mb, p, 1) : launch_tc<2, 3, true>(ma, mb, p, 1);
}

}  // namespace tcr

extern "C" int tcr_gemm_patches(const void* image, const void* b, void* c, const tcr_gemm_desc* desc, const int64_t img_shape[8],
                                const int64_t win_shape[8]) {
  TCR_REQUIRE_DEVICE();
  TCR_ARG(image && b && c && desc && img_shape && win_shape, "tcr_gemm_patches: null argument");
  return tcr::gemm_tc_patches(image, b, c, desc, img_shape, win_shape);
}
