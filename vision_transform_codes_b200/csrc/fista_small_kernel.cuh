// Small-batch ISTA/FISTA in the Gram form with EVERYTHING resident on chip for all iterations (sm_100a).
//
// For problems like BASELINE configs[0] (batch 250, 256 atoms: ista_fista.py:100-133 run 300 times on a 9.8 GFLOP
// problem) an iteration is microseconds of work, and a launch (or a pass of the state through L2) per iteration costs
// more than the arithmetic: 334 launches and 3.6 ms per call with the tiled kernels. Here one CTA pair owns a slice of
// 32 patches for the whole run and nothing leaves the SMs between iterations:
//
//   acc[atom, patch] = sum_k G[atom, k] * y[patch, k]          (the Gram form  y G - b  transposed: atoms on the M axis)
//
//   A = G = Phi Phi^T (S <= 256 atoms x 256, bf16 parts): each CTA of the pair keeps its 128 rows in shared memory
//       for the whole launch (K-major, SWIZZLE_64B, loaded once by TMA);
//   B = y_k (32 patches x 256 atoms, bf16 parts), MN-major (patches contiguous): a thread of the epilogue owns ONE atom
//       (one TMEM lane) and all 32 patches, so the 16 values it produces for either CTA's half of the patches are 32
//       contiguous bytes of that CTA's B tile -- written straight from registers, the peer's half through
//       distributed shared memory; no transpose, no global memory;
//   D = 256 x 32 fp32 in TMEM (32 columns); a_{k-1}, a_{k-2} and the drive b stay in REGISTERS (96 per thread).
//
// Per iteration: the leader's MMA thread waits for the y barrier (one release.cluster arrival per epilogue warp of the
// pair, after every lane has written and fenced its values), issues K / 16 x products tcgen05.mma.cta_group::2 and commits to the accumulator
// barrier of both CTAs; the epilogue threads read their lane, apply the same fused update as every other schedule
// (fista_update16, bit-identical arithmetic), write y_{k+1} and arrive. Rows (patches) are independent, so pairs never
// talk to each other: the grid is ceil(B / 32) pairs.
//
// Limits (the host falls back to the tiled schedule otherwise): Gram form, S <= 256, scalar prox (no group
// shrinkage), no early stopping (its statistic is global), at most two parts (bf16, bf16x3).
#pragma once
#include "fista_iter_kernel.cuh"

namespace vtc {

constexpr int SM_ROWS = 32;        // patches per CTA pair (UMMA N)
constexpr int SM_HALF = 16;        // ... of which each CTA holds 16 in its B tile
constexpr int SM_K = 256;          // padded atom count (K of the contraction, 2 x 128 = M)
constexpr int SM_BK = 32;          // K extent of one TMA box of G (64-byte rows, SWIZZLE_64B)

template <int P>
struct SmallCfg {
  static_assert(P == 1 || P == 2, "parts");
  static constexpr int THREADS = 192;                       // warps 0-3 epilogue (one per TMEM lane quarter), 4 MMA, 5 TMA
  static constexpr int A_BOX = BLOCK_M * SM_BK * 2;         // 8 KB: 128 rows x 32 columns of one part of G
  static constexpr int A_BYTES = P * (SM_K / SM_BK) * A_BOX;   // 64 / 128 KB
  static constexpr int B_TILE = SM_K * SM_HALF * 2;         // 8 KB: one part of this CTA's 16 patches, [k][16] bf16
  static constexpr int B_BYTES = P * B_TILE;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + A_BYTES;
  static constexpr int OFF_BAR = OFF_B + B_BYTES;
  static constexpr int B_A_FULL = 0;      // G has landed (this CTA's own barrier)
  static constexpr int B_Y_READY = 1;     // leader's: 8 warp arrivals, y_k is complete in both CTAs' B tiles
  static constexpr int B_ACC_FULL = 2;    // both CTAs': multicast commit, the accumulator of iteration k is complete
  static constexpr int NUM_BARRIERS = 3;
  static constexpr int SMEM_TOTAL = OFF_BAR + NUM_BARRIERS * 8 + 16;
  static constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;
  static constexpr int NPAIRS = (P == 1) ? 1 : 3;
  static_assert(SMEM_ALLOC <= 232448, "over the 227 KB shared memory limit");
};

struct SmallParams {
  CUtensorMap tmG;        // G_op (S x parts*Sp bf16, row-major): box 32 x 128, SWIZZLE_64B
  int g_part_stride;      // Sp (a multiple of 64, <= 256): columns of one part
  int k_blocks;           // Sp / 32: K blocks of G that exist (atoms at or beyond S are zero in G and in y)
  const float* b;         // drive x Phi^T, tile-contiguous fp32 [ceil(S/16)][B][16]
  const float* init;      // starting point a_0, row-major (B x S) with pitch ld_init, or nullptr for zero
  long long ld_init;
  float* out;             // codes, row-major (B x S), pitch ld_out
  long long ld_out;
  int B, S;
  int num_iters;
  const float* betas;     // betas[k] = momentum coefficient of iteration k, betas[0] = 0
  int prox, use_momentum;
  const float* scalars;   // [0] = eta, [1] = theta
};

// Instruction descriptor for kind::f16 with B MN-major (bit 16): bf16 x bf16 -> fp32, A K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16_b_mn(uint32_t M, uint32_t N) {
  return make_idesc_bf16(M, N) | (1u << 16);
}
// MN-major operand tile, SWIZZLE_32B: 16 elements (32 bytes) contiguous along MN per k, eight k of a group 32 bytes
// apart, groups SBO = 256 bytes apart; a single 16-element span along MN, so LBO is not used. version 1 (sm_100).
__device__ __forceinline__ uint64_t make_mnmajor_sw32_desc(uint32_t smem_addr) {
  const uint32_t hi = ((256u >> 4) & 0x3FFF) | (1u << 14) | (6u << 29);
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFF) | (((32u >> 4) & 0x3FFF) << 16);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_cluster() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait_cluster_acquire(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
  return ok != 0;
}

template <int P>
__global__ void __launch_bounds__(SmallCfg<P>::THREADS, 1) vtc_fista_small_kernel(const __grid_constant__ SmallParams p) {
  using C = SmallCfg<P>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = sbase + C::OFF_A, sB = sbase + C::OFF_B;
  const uint32_t bar0 = sbase + C::OFF_BAR;
  auto bar = [&](int idx) { return bar0 + 8 * idx; };
  const uint32_t tmem_slot = bar0 + C::NUM_BARRIERS * 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1;
  const int row0 = pair * SM_ROWS;          // first patch of this pair's slice

  if (warp == 5 && lane == 0) tma_prefetch_desc(&p.tmG);
  if (warp == 4 && lane == 0) {
    mbar_init(bar(C::B_A_FULL), 1);
    mbar_init(bar(C::B_Y_READY), 2 * 4);         // every epilogue warp of both CTAs
    mbar_init(bar(C::B_ACC_FULL), 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 5) {
    tmem_alloc_pair(tmem_slot, 32);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 5) {
    // ================================ G: this CTA's 128 rows, once ================================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(bar(C::B_A_FULL), P * p.k_blocks * C::A_BOX);
#pragma unroll
      for (int q = 0; q < P; ++q)
        for (int kb = 0; kb < p.k_blocks; ++kb)
          tma_load_2d(sA + (q * (SM_K / SM_BK) + kb) * C::A_BOX, &p.tmG, bar(C::B_A_FULL),
                      q * p.g_part_stride + kb * SM_BK, cta_rank * BLOCK_M, kEvictLast);
    }
    __syncwarp();
  } else if (warp == 4) {
    // ================================ MMA issuer (leader CTA) ================================
    // (the peer's G tile is known to have landed: its epilogue threads wait for it before their first arrival on
    // the y barrier)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16_b_mn(PAIR_M, SM_ROWS);
      mbar_wait(bar(C::B_A_FULL), 0);
      for (int k = 1; k <= p.num_iters; ++k) {
        // y_{k-1} is complete in both B tiles (release.cluster arrivals, acquired here)
        uint32_t spins = 0;
        while (!mbar_try_wait_cluster_acquire(bar(C::B_Y_READY), (k - 1) & 1)) {
          if (++spins > (1u << 22)) {
            printf("vtc_b200: small kernel, y barrier timeout (block %d, iteration %d)\n", (int)blockIdx.x, k);
            __trap();
          }
        }
        fence_proxy_async_cluster();   // generic-proxy writes of y (local and remote) -> the tensor core's async reads
        tc_fence_after();
        if (elect_one_sync()) {
          uint32_t accumulate = 0;
          // K block by K block, every product of a block before the next block: the accumulation order of the tiled
          // Gram-form kernel (gemm_kernel.cuh), so the two schedules give bit-identical iterates
          for (int kb = 0; kb < p.k_blocks; ++kb) {
#pragma unroll
            for (int pr = 0; pr < C::NPAIRS; ++pr) {
              // (operands are swapped against the tiled kernel, where y is A and G is B: the part of y is pair_a, the
              // part of G is pair_b, so that the three products are accumulated in the same order)
              const uint32_t a_box = sA + (pair_b(P, pr) * (SM_K / SM_BK) + kb) * C::A_BOX;
              const uint32_t b_part = sB + pair_a(P, pr) * C::B_TILE;
#pragma unroll
              for (int k2 = 0; k2 < SM_BK / UMMA_K; ++k2) {
                const int ks = kb * (SM_BK / UMMA_K) + k2;
                // A: 32 bytes further inside the 64-byte rows of the box per step; B: 16 k-rows of 32 bytes per step
                const uint64_t adesc = make_kmajor_desc(a_box, SM_BK * 2) + 2 * k2;
                const uint64_t bdesc = make_mnmajor_sw32_desc(b_part + ks * (UMMA_K * SM_HALF * 2));
                umma_bf16_pair(tmem_base, adesc, bdesc, idesc, accumulate);
                accumulate = 1;
              }
            }
          }
          umma_commit_pair(bar(C::B_ACC_FULL), 3);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue: one atom per thread, 32 patches ================================
    const int atom = cta_rank * BLOCK_M + warp * 32 + lane;     // TMEM lane = warp * 32 + lane
    const bool atom_ok = atom < p.S;
    UpdateArgs ua;
    ua.prox = p.prox, ua.group = 1, ua.use_momentum = p.use_momentum;
    ua.eta = __ldg(p.scalars + 0), ua.theta = __ldg(p.scalars + 1);
    ua.want_stat = false;
    float stat_unused = 0.f;
    // state in registers: a1 = a_{k-1}, a2 = a_{k-2}, bd = drive. Rows beyond the batch and atoms beyond S are zero.
    float a1[SM_ROWS], a2[SM_ROWS], bd[SM_ROWS];
#pragma unroll
    for (int n = 0; n < SM_ROWS; ++n) {
      const long long row = static_cast<long long>(row0) + n;
      const bool ok = atom_ok && row < p.B;
      bd[n] = ok ? __ldg(p.b + (static_cast<long long>(atom / EPI_COLS) * p.B + row) * EPI_COLS + atom % EPI_COLS) : 0.f;
      a1[n] = (ok && p.init != nullptr) ? __ldg(p.init + row * p.ld_init + atom) : 0.f;
      a2[n] = a1[n];
    }
    // B tiles: [part][k = atom][16 patches] bf16, SWIZZLE_32B (16-byte chunk c of k-row r sits at c ^ ((r >> 2) & 1))
    const uint32_t sw = (atom >> 2) & 1;
    const uint32_t my_row = sB + atom * (SM_HALF * 2);
    const uint32_t peer_row = mapa_cluster(my_row, cta_rank ^ 1);
    const uint32_t y_ready = mapa_cluster(bar(C::B_Y_READY), 0);
    // the 16 values of this atom for the patches of CTA h: one 32-byte k-row of that CTA's tile, per part
    auto write_half = [&](int h, const float (&y16)[16]) {
      const bool local = (h == cta_rank);
      const uint32_t base = local ? my_row : peer_row;
      split_parts16(y16, P, [&](int part, const uint32_t (&w32)[8]) {
        const uint32_t dst = base + part * C::B_TILE;
        if (local) {
          sts128u(dst + ((0 ^ sw) << 4), w32[0], w32[1], w32[2], w32[3]);
          sts128u(dst + ((1 ^ sw) << 4), w32[4], w32[5], w32[6], w32[7]);
        } else {
          st_cluster_v4(dst + ((0 ^ sw) << 4), w32[0], w32[1], w32[2], w32[3]);
          st_cluster_v4(dst + ((1 ^ sw) << 4), w32[4], w32[5], w32[6], w32[7]);
        }
      });
    };
    // y_k complete for this warp: every lane's writes (its own and the peer's tile) ordered before ONE release at
    // cluster scope (the peer's tensor core reads what this CTA wrote through distributed shared memory)
    auto publish_y = [&]() {
      fence_proxy_async_cluster();
      __syncwarp();
      if (lane == 0) mbar_arrive_release_cluster(y_ready);
    };
    mbar_wait(bar(C::B_A_FULL), 0);   // this CTA's G is in place before the first arrival that lets the MMAs start
    {
      float y16[16];
#pragma unroll
      for (int h = 0; h < 2; ++h) {   // y_0 = a_0
#pragma unroll
        for (int x = 0; x < 16; ++x) y16[x] = a1[h * 16 + x];
        write_half(h, y16);
      }
      publish_y();
    }
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int k = 1; k <= p.num_iters; ++k) {
      ua.beta_prev = __ldg(p.betas + k - 1);
      ua.beta_next = __ldg(p.betas + k);
      const bool has_prev2 = p.use_momentum != 0 && ua.beta_prev != 0.f;
      ua.in_mask = has_prev2 ? 7 : 3;
      mbar_wait(bar(C::B_ACC_FULL), (k - 1) & 1);
      tc_fence_after();
      uint32_t v0[16], v1[16];
      tmem_ld16(t_lane, v0);
      tmem_ld16(t_lane + 16, v1);
      tmem_ld_wait();
      tc_fence_before();   // the accumulator has been read: the MMAs of the next iteration may overwrite it
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float in[3][16], outv[16], partv[16];
#pragma unroll
        for (int x = 0; x < 16; ++x) in[0][x] = a1[h * 16 + x], in[1][x] = bd[h * 16 + x], in[2][x] = a2[h * 16 + x];
        fista_update16(ua, h == 0 ? v0 : v1, in, outv, partv, stat_unused);
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          a2[h * 16 + x] = a1[h * 16 + x];
          a1[h * 16 + x] = outv[x];
        }
        if (k < p.num_iters) write_half(h, partv);
      }
      if (k < p.num_iters) publish_y();
    }
    // the last iterate, straight to the caller's row-major codes
    if (atom_ok) {
#pragma unroll
      for (int n = 0; n < SM_ROWS; ++n) {
        const long long row = static_cast<long long>(row0) + n;
        if (row < p.B) p.out[row * p.ld_out + atom] = a1[n];
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 5) tmem_dealloc_pair(tmem_base, 32);
}

}  // namespace vtc
