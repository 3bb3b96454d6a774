// Panel-resident ISTA/FISTA iterations, second generation (sm_100a): the default for S > 2 D, D <= 256, bf16 / bf16x3.
// Same algorithm, job structure, persistent schedule and bit-exact arithmetic as fista_iter_kernel.cuh (read that
// header first); what changed is how the epilogue is fed, after the round-2 ablations (profiles/README.md) showed
// that the first kernel was bound by its own handshakes and by the latency its shallow rings exposed (two G stages:
// an L2 round trip per K block; two input stages per math group, each held from the load through the math to the end
// of the store), not by HBM, the tensor pipe or shared-memory bandwidth:
//
//   * a_{k-2} does not go through shared memory: the math warps read it with coalesced ld.global.cg (prefetched one
//     sub-tile ahead into registers) from a state layout made for it -- every sub-tile is one contiguous 8 KB block
//     [4 column quads][128 rows][4 floats], so a warp's 128-bit load covers 512 contiguous bytes. Only a_{k-1} is
//     staged (1-D bulk copy, no tensor map) and a_k is written over it in place (each thread touches its own row
//     only): in/out stages are 8 KB instead of 16;
//   * the shared memory that frees goes where the latency was exposed: nine input stages (three per math group
//     instead of two) and THREE G stages instead of two (bf16x3; plain bf16: twelve input stages);
//   * setmaxnreg moves registers from the eight non-math warps (40 each) to the math warps (128 each with three
//     groups): room for the prefetch registers without spills;
//   * G and R MMAs are issued by two warps that block on their own barriers; the operand "full" barriers take one
//     arrival (the leader's expect_tx for both CTAs' bytes) instead of a cross-CTA arrive per stage;
//   * starting from zero nothing is initialised: iteration 1 loads no state, every block is written before it is read.
//
// (The first version of this kernel used 32-atom chunks as the epilogue's work unit -- one y-ring handshake per chunk --
// and read x with plain loads at the panel end: 9 % SLOWER than the first generation, because four chunks on three
// groups made one group the critical path of every tile and the panel end took twice as long. Work units are 16-column
// sub-tiles again and the panel end is staged by TMA.)
//
// State buffers (IterParams2::state) are "quad-blocked": [col block cb = atom / 16][row block rb = row / 128][quad][row % 128][4].
// The caller converts a warm start into that layout and the result back to row-major (block_quad / unblock_quad
// kernels in aux_kernels.cuh); rows beyond the batch and atoms beyond S are zero and stay zero.
//
// Warp roles (4 NG + 8 warps, NG = 3 or 4 math groups):
//   0          TMA producer of the G operand ring (r_op K blocks + Phi tile halves), both CTAs
//   1          tcgen05.mma issuer of G (leader CTA)
//   2          TMEM allocator, then bulk / TMA stores (a_k sub-tiles, r_op parts)
//   3          bulk loader of a_{k-1} (x at the panel end)
//   4 ..       epilogue math, NG groups of four warps (one warp per TMEM lane quarter) on sub-tiles round-robin
//   4 NG + 4   TMA producer of the Phi^T chunk ring (B operand of R), both CTAs
//   4 NG + 5   tcgen05.mma issuer of R (leader CTA)
//   last two   none (they complete the last warpgroup: setmaxnreg is warpgroup-wide)
#pragma once
#include "fista_iter_kernel.cuh"

namespace vtc {

template <int P, int NG = 3>
struct Iter2Cfg {
  static_assert(P == 1 || P == 2, "parts");
  static_assert(NG == 3 || NG == 4, "math groups");
  static constexpr int GROUPS = NG;
  static constexpr int MATH_WARPS = 4 * GROUPS;
  static constexpr int PT_WARP = 4 + MATH_WARPS;
  static constexpr int R_WARP = PT_WARP + 1;
  // whole warpgroups (setmaxnreg is a warpgroup-wide instruction): 20 / 24 warps, the last two have no role
  static constexpr int THREADS = 32 * (MATH_WARPS + 8);
  static constexpr int LAUNCH_REGS = (65536 / THREADS) / 8 * 8;   // what the launch gives every thread (96 / 80)
  // The launch gives every warp the compiled 96 registers (5 warps per scheduler): the CTA's pool is 20 x 32 x 96 =
  // 61440 registers. The eight non-math warps hand most of theirs back and the twelve math warps take 128:
  // (8 x 40 + 12 x 128) x 32 = 59392 <= 61440 (setmaxnreg.inc blocks for ever when the pool cannot cover it).
  static constexpr int REGS_OTHER = 40;
  static constexpr int REGS_MATH = ((THREADS * LAUNCH_REGS - 8 * 32 * REGS_OTHER) / (MATH_WARPS * 32)) / 8 * 8;  // 128 / 96
  static_assert((8 * REGS_OTHER + MATH_WARPS * REGS_MATH) * 32 <= THREADS * LAUNCH_REGS, "register pool of the CTA");
  static constexpr int BK = (P == 1) ? 64 : 32;          // K extent of a G stage
  static constexpr int SPAN = BK * 2;
  static constexpr int A_TILE = BLOCK_M * SPAN;           // one part of this CTA's 128 rows of r_op
  static constexpr int B_TILE = (IT_BN / 2) * SPAN;       // one part of this CTA's 64 atoms of the Phi tile
  static constexpr int G_STAGE = P * (A_TILE + B_TILE);
  static constexpr int CHUNK = 32;                        // atoms per epilogue unit = K extent of one group of R MMAs
  static constexpr int Y_TILE = BLOCK_M * CHUNK * 2;      // one part of y: 128 rows x 32 atoms, SWIZZLE_64B
  static constexpr int Y_STAGE = P * Y_TILE;
  static constexpr int PT_TILE = (IT_RN / 2) * CHUNK * 2; // one part of this CTA's 128 pixel rows of Phi^T
  static constexpr int PT_STAGE = P * PT_TILE;
  static constexpr int IN_STAGE = EPI_ARRAY_BYTES;        // a_{k-1} in, a_k (or the r parts of a panel-end sub-tile) out
  // three groups: 3 input stages each, 3 G stages, 3 y stages; four groups: 2 input stages each, 4 G stages, 2 y
  // stages (a y stage then always has the same two writer groups). Plain bf16 (half the operand bytes) spends what is
  // left on a deeper input ring: it is HBM-bound
#ifndef VTC_IT2_IN3
#define VTC_IT2_IN3 9   // (tuning builds: tools/ab_build.sh with NVCC_EXTRA=-DVTC_IT2_IN3=.. -DVTC_IT2_G3=..)
#endif
#ifndef VTC_IT2_G3
#define VTC_IT2_G3 3
#endif
  static constexpr int IN_STAGES = (NG == 4) ? (P == 1 ? 12 : 8) : (P == 1 ? 12 : VTC_IT2_IN3);
  static constexpr int G_STAGES = (NG == 4 && P == 2) ? 4 : (P == 2 ? VTC_IT2_G3 : 3);
#ifndef VTC_IT2_Y
#define VTC_IT2_Y 3
#endif
  static constexpr int Y_STAGES = (NG == 4) ? 2 : (P == 2 ? VTC_IT2_Y : 3);
#ifndef VTC_IT2_PT
#define VTC_IT2_PT 2
#endif
  static constexpr int PT_STAGES = (P == 2) ? (NG == 3 ? VTC_IT2_PT : 2) : 4;
  // panel-end sub-tiles are padded to a multiple of this, so that the running sub-tile index (math group, in/out stage)
  // and the running y chunk index stay congruent from job to job: every y stage always has the same writer groups
  static constexpr int PANEL_END_PAD = (NG == 4) ? 4 : 6;
  static constexpr int OFF_G = 0;
  static constexpr int OFF_Y = OFF_G + G_STAGES * G_STAGE;
  static constexpr int OFF_PT = OFF_Y + Y_STAGES * Y_STAGE;
  static constexpr int OFF_IN = OFF_PT + PT_STAGES * PT_STAGE;
  static constexpr int OFF_BAR = OFF_IN + IN_STAGES * IN_STAGE;
  static constexpr int B_G_FULL = 0;
  static constexpr int B_G_EMPTY = B_G_FULL + G_STAGES;
  static constexpr int B_PT_FULL = B_G_EMPTY + G_STAGES;
  static constexpr int B_PT_EMPTY = B_PT_FULL + PT_STAGES;
  static constexpr int B_Y_FULL = B_PT_EMPTY + PT_STAGES;
  static constexpr int B_Y_EMPTY = B_Y_FULL + Y_STAGES;
  static constexpr int B_ACCG_FULL = B_Y_EMPTY + Y_STAGES;
  static constexpr int B_ACCG_EMPTY = B_ACCG_FULL + 2;
  static constexpr int B_ACCR_FULL = B_ACCG_EMPTY + 2;
  static constexpr int B_ACCR_EMPTY = B_ACCR_FULL + 1;
  static constexpr int B_IN_FULL = B_ACCR_EMPTY + 1;
  static constexpr int B_IN_FREE = B_IN_FULL + IN_STAGES;
  static constexpr int B_OUT_FULL = B_IN_FREE + IN_STAGES;
  static constexpr int NUM_BARRIERS = B_OUT_FULL + IN_STAGES;
  static constexpr int SMEM_TOTAL = OFF_BAR + NUM_BARRIERS * 8 + 16;
  static constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;
  static constexpr int NPAIRS = (P == 1) ? 1 : 3;
  static constexpr int STORES_IN_FLIGHT = 1;
  // a stage is always consumed by the same group (stage e belongs to group (e % 6) / 2): a group sees the phases of
  // its stages' barriers strictly in order
  static_assert(IN_STAGES % GROUPS == 0, "input stages must have a fixed owner group");
  static_assert(P * EPI_PART_BYTES <= IN_STAGE, "r parts must fit a stage");
  static_assert(SMEM_ALLOC <= 232448, "over the 227 KB shared memory limit");
};

struct IterParams2 {
  CUtensorMap tmR;      // r_op (bf16 parts, tile-contiguous [part][Dp/BK][rows][BK]): box BK x 128, A operand of G
  CUtensorMap tmPhi;    // phi_op (S x parts*Dp, row-major): box BK x 64, B operand of G
  CUtensorMap tmPhiT;   // phiT_op (D x parts*Sp, row-major): box 32 x 128, SWIZZLE_64B, B operand of R
  CUtensorMap tmROut;   // r_op as a store target: box 16 x 128, SWIZZLE_32B
  CUtensorMap tmX;      // images (B x D fp32, row-major), box 16 x 128, SWIZZLE_64B
  float* state[3];      // quad-blocked fp32 code arrays: [0] the starting point a_0 (unused when init_zero), [1] a_k
                        // for odd k, [2] a_k for even k (a_k overwrites a_{k-2} in place; the final iterate too)
  const float* x;       // images, row-major (B x D), pitch ld_x floats (16-byte aligned rows)
  long long ld_x;
  int B, D;
  int init_zero;        // a_0 = 0: iteration 1 loads no state at all
  int num_panels;       // ceil(B / 256)
  int S;
  int num_n_tiles;      // ceil(S / 128)
  int kb_g;             // Dp / BK: K blocks of G (= column blocks per part of r_op)
  int phi_part_stride;  // Dp
  int phiT_part_stride; // Sp
  int nsub_r;           // Dp / 16: sub-tiles of r written at the panel end
  int r_block_w;        // BK of r_op's layout
  int k_first, k_count; // this launch runs iterations k_first .. k_first + k_count - 1 of every panel
  int k_final;          // the iteration whose output is the result: no r_k produced (INT_MAX: the caller decides)
  const float* betas;   // device: betas[k] = FISTA momentum coefficient of iteration k, betas[0] = 0
  int* done;            // device, per panel: CTAs that have completed a job of this launch on it (nullptr: k_count == 1)
  int prox, group, use_momentum;
  const float* scalars;  // device: [0] = eta, [1] = theta
  double* stat;          // optional: += sum |a_k - a_{k-1}|
  unsigned long long* trace;
  int ablate;            // timing experiments only (AblateBits; results are wrong when non-zero)
  int l2_prefetch;       // sub-tiles of a_{k-2} prefetched into L2 by the loader alongside the a_{k-1} loads
};

// ---- 1-D bulk copies (no tensor map): contiguous global <-> shared, sizes and addresses multiples of 16 bytes
// VTC_IT2_STATE_HINT (tuning builds): 1 = the state streams carry an L2 evict-first policy (they are read once per
// iteration and never hit: 0.8 GB per iteration through a 126 MB L2 that also has to keep r_op and the dictionary)
#ifndef VTC_IT2_STATE_HINT
#define VTC_IT2_STATE_HINT 0
#endif
// VTC_IT2_R_HINT (tuning builds): 1 = r_op (67 MB at configs[1], rewritten in place every iteration and read eight
// times per job) is stored and loaded with an L2 evict-last policy, so that it can stay in the 126 MB L2 instead of
// making the round trip through HBM
#ifndef VTC_IT2_R_HINT
#define VTC_IT2_R_HINT 0
#endif
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
#if VTC_IT2_STATE_HINT
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar), "l"(kEvictFirst) : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, uint32_t smem_src, uint32_t bytes) {
#if VTC_IT2_STATE_HINT
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(gdst), "r"(smem_src), "r"(bytes), "l"(kEvictFirst) : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
               : "memory");
#endif
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// state written by another SM earlier in this launch: L2 is the point of coherence, never a (possibly stale) L1 line
__device__ __forceinline__ float4 ldg_cg_v4(const float* p) {
  float4 v;
#if VTC_IT2_STATE_HINT
  asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(kEvictFirst));
#else
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
#endif
  return v;
}
__device__ __forceinline__ float4 ldg_nc_v4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <int P, int NG>
__global__ void __launch_bounds__((Iter2Cfg<P, NG>::THREADS), 1) vtc_fista_iter2_kernel(const __grid_constant__ IterParams2 p) {
  using C = Iter2Cfg<P, NG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sG = sbase + C::OFF_G, sY = sbase + C::OFF_Y, sPT = sbase + C::OFF_PT;
  const uint32_t sIn = sbase + C::OFF_IN;
  const uint32_t bar0 = sbase + C::OFF_BAR;
  auto bar = [&](int idx) { return bar0 + 8 * idx; };
  const uint32_t tmem_slot = bar0 + C::NUM_BARRIERS * 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int NT = p.num_n_tiles;
  const long long total_jobs = static_cast<long long>(p.k_count) * p.num_panels;
  const int my_jobs = total_jobs > cluster_id
                          ? static_cast<int>((total_jobs - cluster_id + num_clusters - 1) / num_clusters) : 0;
  const int my_tiles = my_jobs * NT;
  struct Job {
    int k, m0, panel, wait_target;
    bool do_r, has_prev, has_prev2;
    int prev, prev2, out;   // indices into state
    float beta_prev, beta_next;
  };
  auto job_at = [&](int ji) {
    Job j;
    const long long ticket = cluster_id + static_cast<long long>(ji) * num_clusters;
    const int ko = static_cast<int>(ticket / p.num_panels);
    j.k = p.k_first + ko;
    j.panel = static_cast<int>(ticket - static_cast<long long>(ko) * p.num_panels);
    j.m0 = j.panel * PAIR_M + cta_rank * BLOCK_M;
    j.wait_target = 2 * ko;   // both CTAs of every earlier job of this launch on this panel
    j.do_r = j.k < p.k_final;
    j.beta_prev = __ldg(p.betas + j.k - 1);
    j.beta_next = __ldg(p.betas + j.k);
    j.has_prev = !(j.k == 1 && p.init_zero);
    j.has_prev2 = p.use_momentum != 0 && j.beta_prev != 0.f;
    j.prev = (j.k == 1) ? 0 : (((j.k - 1) & 1) ? 1 : 2);
    j.prev2 = (j.k & 1) ? 1 : 2;     // only read when has_prev2 (k >= 3)
    j.out = (j.k & 1) ? 1 : 2;
    return j;
  };
  // the data of job (k - 1, panel) must be complete before anything of job (k, panel) is read
  auto wait_for_previous = [&](const Job& j, bool for_tma) {
    if (p.done != nullptr && j.wait_target > 0) {
      if (lane == 0) {
        // The predecessor runs on another CTA pair of this launch (the whole grid is resident: launch_iter_p). Almost
        // never taken; when it is, back off instead of hammering L2, and give up only after about a minute -- other
        // work sharing the GPU (another stream, NCCL, MPS) may legitimately delay a pair for a long time, and a trap
        // takes the whole CUDA context down.
        uint32_t spins = 0;
        while (ld_acquire_gpu(p.done + j.panel) < j.wait_target) {
          if (++spins > 1024u) asm volatile("nanosleep.u32 256;" ::: "memory");
          if (spins > (1u << 28)) {
            printf("vtc_b200: job (k %d, panel %d) never saw its predecessor (block %d)\n", j.k, j.panel, (int)blockIdx.x);
            __trap();
          }
        }
      }
      __syncwarp();
      if (for_tma) fence_proxy_async_all();   // the acquired data is read through TMA / bulk copies (async proxy)
    }
  };
  // chunks of tile nt (whole chunks; atoms at or beyond S are zero everywhere)
  auto tile_chunks = [&](int nt) { return (min(IT_BN, p.S - nt * IT_BN) + C::CHUNK - 1) / C::CHUNK; };
  int state_subs = 0;    // 16-atom sub-tiles of one job (whole 32-atom chunks)
  for (int nt = 0; nt < NT; ++nt) state_subs += 2 * tile_chunks(nt);
  // panel-end sub-tiles padded to a multiple of 6, so that the running sub-tile index (math group, in/out stage) and
  // the running y chunk index stay congruent from job to job: every y stage always has the same two writer groups
  const int nsub_r_pad = (p.nsub_r + C::PANEL_END_PAD - 1) / C::PANEL_END_PAD * C::PANEL_END_PAD;
  const long long row_blocks = static_cast<long long>(p.num_panels) * 2;
  // float offset of sub-tile (col block cb, this CTA's row block of panel) in a quad-blocked state array
  auto state_offset = [&](int panel, int cb) {
    return (static_cast<long long>(cb) * row_blocks + panel * 2 + cta_rank) * (EPI_ARRAY_BYTES / 4);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmR);
    tma_prefetch_desc(&p.tmPhi);
    tma_prefetch_desc(&p.tmPhiT);
    tma_prefetch_desc(&p.tmROut);
    tma_prefetch_desc(&p.tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::G_STAGES; ++s) {
      // one arrival: the leader's expect_tx for BOTH CTAs' bytes (the peer's loads complete their bytes on the same
      // barrier; bytes that land before the expect_tx leave the transaction count negative until it is posted)
      mbar_init(bar(C::B_G_FULL + s), 1);
      mbar_init(bar(C::B_G_EMPTY + s), 1);   // multicast tcgen05.commit
    }
    for (int s = 0; s < C::PT_STAGES; ++s) {
      mbar_init(bar(C::B_PT_FULL + s), 1);
      mbar_init(bar(C::B_PT_EMPTY + s), 1);
    }
    for (int s = 0; s < C::Y_STAGES; ++s) {
      mbar_init(bar(C::B_Y_FULL + s), 2 * 2 * 4);  // 2 CTAs x the two sub-tiles of a chunk x 4 warps (leader's barrier)
      mbar_init(bar(C::B_Y_EMPTY + s), 1);        // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(C::B_ACCG_FULL + a), 1);
      mbar_init(bar(C::B_ACCG_EMPTY + a), 2 * C::MATH_WARPS);
    }
    mbar_init(bar(C::B_ACCR_FULL), 1);
    mbar_init(bar(C::B_ACCR_EMPTY), 2 * C::MATH_WARPS);
    for (int e = 0; e < C::IN_STAGES; ++e) {
      mbar_init(bar(C::B_IN_FULL + e), 1);
      mbar_init(bar(C::B_IN_FREE + e), 1);    // the storer, once the store of the stage's result has read it
      mbar_init(bar(C::B_OUT_FULL + e), 4);   // the four math warps that wrote the result
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_launch_dependents();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const bool is_math = warp >= 4 && warp < 4 + C::MATH_WARPS;

  if (!is_math) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REGS_OTHER));
  if (warp == 0) {
    // ================================ G operand producer (both CTAs) ================================
    Tracer trace(p, 0, lane == 0);
    uint32_t it = 0;
    for (int ji = 0; ji < my_jobs; ++ji) {
      const Job job = job_at(ji);
      const int m0 = job.m0;
      wait_for_previous(job, true);   // r_op[panel] is the previous iteration's output
      for (int nt = 0; nt < NT; ++nt) {
        const int n0 = nt * IT_BN;
        for (int kb = 0; kb < p.kb_g; ++kb, ++it) {
          const int s = it % C::G_STAGES;
          const uint32_t ph = (it / C::G_STAGES) & 1;
          mbar_wait(bar(C::B_G_EMPTY + s), ph ^ 1);
          if (elect_one_sync()) {
            const uint32_t full = bar(C::B_G_FULL + s);
            const bool load_b = !(ablate_of(p) & ABL_G_LOAD);
            const bool load_a = load_b && !((ablate_of(p) & ABL_R_STREAM) && nt > 0);
            if (leader) {
              if (load_b) mbar_arrive_expect_tx(full, load_a ? 2 * C::G_STAGE : 2 * P * C::B_TILE);
              else mbar_arrive(full);
            }
            const uint32_t dst = sG + s * C::G_STAGE;
            trace(TR_G_LOAD, nt * p.kb_g + kb);
#pragma unroll
            for (int q = 0; q < P; ++q)
              if (load_a) tma_load_3d_pair(dst + q * C::A_TILE, &p.tmR, full, 0, m0, q * p.kb_g + kb,
                                           VTC_IT2_R_HINT ? kEvictLast : kEvictNormal);
#pragma unroll
            for (int q = 0; q < P; ++q)
              if (load_b) tma_load_2d_pair(dst + P * C::A_TILE + q * C::B_TILE, &p.tmPhi, full,
                                           q * p.phi_part_stride + kb * C::BK, n0 + cta_rank * (IT_BN / 2), kEvictLast);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == C::PT_WARP) {
    // ================================ Phi^T chunk producer (both CTAs) ================================
    uint32_t it = 0;
    for (int ji = 0; ji < my_jobs; ++ji) {
      if (!job_at(ji).do_r) continue;
      for (int nt = 0; nt < NT; ++nt) {
        const int nch = tile_chunks(nt);
        for (int c = 0; c < nch; ++c, ++it) {
          const int s = it % C::PT_STAGES;
          const uint32_t ph = (it / C::PT_STAGES) & 1;
          mbar_wait(bar(C::B_PT_EMPTY + s), ph ^ 1);
          if (elect_one_sync()) {
            const uint32_t full = bar(C::B_PT_FULL + s);
            const bool load_pt = !(ablate_of(p) & ABL_PT_LOAD);
            if (leader) {
              if (load_pt) mbar_arrive_expect_tx(full, 2 * C::PT_STAGE);
              else mbar_arrive(full);
            }
#pragma unroll
            for (int q = 0; q < P; ++q)
              if (load_pt) tma_load_2d_pair(sPT + s * C::PT_STAGE + q * C::PT_TILE, &p.tmPhiT, full,
                                            q * p.phiT_part_stride + nt * IT_BN + c * C::CHUNK, cta_rank * (IT_RN / 2),
                                            kEvictLast);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ G MMA issuer (leader CTA) ================================
    if (leader) {
      constexpr uint32_t idesc_g = make_idesc_bf16(PAIR_M, IT_BN);
      Tracer trace(p, 1, lane == 0);
      uint32_t g_it = 0;
      for (int g_tile = 0; g_tile < my_tiles; ++g_tile) {
        const int acc = g_tile & 1;
        mbar_wait(bar(C::B_ACCG_EMPTY + acc), ((g_tile >> 1) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + acc * IT_BN;
        uint32_t accumulate = 0;
        for (int g_kb = 0; g_kb < p.kb_g; ++g_kb, ++g_it) {
          const int s = g_it % C::G_STAGES;
          mbar_wait(bar(C::B_G_FULL + s), (g_it / C::G_STAGES) & 1);
          tc_fence_after();
          const uint32_t stage = sG + s * C::G_STAGE;
          const bool last = (g_kb == p.kb_g - 1);
          if (elect_one_sync()) {
            if (g_kb == 0) trace(TR_G_BEGIN, g_tile);
            if (last) trace(TR_G_END, g_tile);
#pragma unroll
            for (int pr = 0; pr < C::NPAIRS; ++pr) {
              const uint64_t adesc = make_kmajor_desc(stage + pair_a(P, pr) * C::A_TILE, C::SPAN);
              const uint64_t bdesc = make_kmajor_desc(stage + P * C::A_TILE + pair_b(P, pr) * C::B_TILE, C::SPAN);
#pragma unroll
              for (int k = 0; k < C::BK / UMMA_K; ++k) {
                if (!(ablate_of(p) & ABL_G_MMA)) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_g, accumulate);
                accumulate = 1;
              }
            }
            umma_commit_pair(bar(C::B_G_EMPTY + s), 3);
            if (last) umma_commit_pair(bar(C::B_ACCG_FULL + acc), 3);
          }
          accumulate = 1;
          __syncwarp();
        }
      }
    }
  } else if (warp == C::R_WARP) {
    // ================================ R MMA issuer (leader CTA) ================================
    if (leader) {
      constexpr uint32_t idesc_r = make_idesc_bf16(PAIR_M, IT_RN);
      const uint32_t acc_r = tmem_base + 2 * IT_BN;
      uint32_t r_it = 0;  // running chunk index: y stage r_it % Y_STAGES, Phi^T stage r_it % PT_STAGES
      int r_jobs = 0;     // jobs with an R so far (parity of the acc_r barriers)
      for (int pi = 0; pi < my_jobs; ++pi) {
        if (!job_at(pi).do_r) continue;   // the final iteration has no R
        for (int nt = 0; nt < NT; ++nt) {
          const int nch = tile_chunks(nt);
          for (int c = 0; c < nch; ++c, ++r_it) {
            const bool first = (nt == 0 && c == 0);
            const int ys = r_it % C::Y_STAGES, ps = r_it % C::PT_STAGES;
            mbar_wait(bar(C::B_Y_FULL + ys), (r_it / C::Y_STAGES) & 1);
            mbar_wait(bar(C::B_PT_FULL + ps), (r_it / C::PT_STAGES) & 1);
            // the panel-end epilogue of the previous job with an R must have drained acc_r before it is overwritten
            if (first) mbar_wait(bar(C::B_ACCR_EMPTY), (r_jobs & 1) ^ 1);
            tc_fence_after();
            const uint32_t ystage = sY + ys * C::Y_STAGE, pstage = sPT + ps * C::PT_STAGE;
            const bool last = (nt == NT - 1) && (c == nch - 1);
            if (elect_one_sync()) {
              uint32_t accumulate = first ? 0u : 1u;
#pragma unroll
              for (int pr = 0; pr < C::NPAIRS; ++pr) {
                const uint64_t adesc = make_kmajor_desc(ystage + pair_a(P, pr) * C::Y_TILE, C::CHUNK * 2);
                const uint64_t bdesc = make_kmajor_desc(pstage + pair_b(P, pr) * C::PT_TILE, C::CHUNK * 2);
#pragma unroll
                for (int k = 0; k < C::CHUNK / UMMA_K; ++k) {
                  if (!(ablate_of(p) & ABL_R_MMA)) umma_bf16_pair(acc_r, adesc + 2 * k, bdesc + 2 * k, idesc_r, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit_pair(bar(C::B_Y_EMPTY + ys), 3);
              umma_commit_pair(bar(C::B_PT_EMPTY + ps), 3);
              if (last) umma_commit_pair(bar(C::B_ACCR_FULL), 3);
            }
            __syncwarp();
          }
        }
        ++r_jobs;
      }
    }
  } else if (warp == 3) {
    // ================================ loader of a_{k-1} (and x at the panel end) ================================
    uint32_t q = 0;   // running sub-tile index: stage q % IN_STAGES, math group q % GROUPS
    for (int ji = 0; ji < my_jobs; ++ji) {
      const Job job = job_at(ji);
      wait_for_previous(job, true);   // a_{k-1} (and the block a_k overwrites) belong to the previous iterations
      const float* src = p.state[job.prev];
      const float* src2 = p.state[job.prev2];
      const int nsub = state_subs + (job.do_r ? nsub_r_pad : 0);
      for (int j = 0; j < nsub; ++j, ++q) {
        const int e = q % C::IN_STAGES;
        mbar_wait(bar(C::B_IN_FREE + e), ((q / C::IN_STAGES) & 1) ^ 1);
        if (elect_one_sync()) {
          const uint32_t full = bar(C::B_IN_FULL + e);
          if (j < state_subs) {
            if (job.has_prev && !(ablate_of(p) & ABL_STATE_LOAD)) {
              const long long off = state_offset(job.panel, j);
              mbar_arrive_expect_tx(full, EPI_ARRAY_BYTES);
              bulk_load_1d(sIn + e * C::IN_STAGE, src + off, EPI_ARRAY_BYTES, full);
              if (p.l2_prefetch && job.has_prev2) bulk_prefetch_l2(src2 + off, EPI_ARRAY_BYTES);
            } else {
              mbar_arrive(full);   // iteration 1 from zero: nothing staged
            }
          } else if (j - state_subs < p.nsub_r) {
            mbar_arrive_expect_tx(full, EPI_ARRAY_BYTES);
            tma_load_2d(sIn + e * C::IN_STAGE, &p.tmX, full, (j - state_subs) * EPI_COLS, job.m0, kEvictNormal);
          } else {
            mbar_arrive(full);     // padding sub-tile of the panel end
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ================================ storer ================================
    uint32_t q = 0;
    for (int ji = 0; ji < my_jobs; ++ji) {
      const Job job = job_at(ji);
      float* dst = p.state[job.out];
      const int nsub = state_subs + (job.do_r ? nsub_r_pad : 0);
      for (int j = 0; j < nsub; ++j, ++q) {
        const int e = q % C::IN_STAGES;
        mbar_wait(bar(C::B_OUT_FULL + e), (q / C::IN_STAGES) & 1);
        const uint32_t src = sIn + e * C::IN_STAGE;
        if (elect_one_sync()) {
          if (j < state_subs) {
            if (!(ablate_of(p) & ABL_STATE_STORE)) bulk_store_1d(dst + state_offset(job.panel, j), src, EPI_ARRAY_BYTES);
          } else if (j - state_subs < p.nsub_r) {
            const int col = (j - state_subs) * EPI_COLS;
#pragma unroll
            for (int part = 0; part < P; ++part)
#if VTC_IT2_R_HINT
              tma_store_3d_hint(&p.tmROut, src + part * EPI_PART_BYTES, col % p.r_block_w, job.m0,
                                part * p.kb_g + col / p.r_block_w, kEvictLast);
#else
              tma_store_3d(&p.tmROut, src + part * EPI_PART_BYTES, col % p.r_block_w, job.m0,
                           part * p.kb_g + col / p.r_block_w);
#endif
          }
          bulk_commit();
          if (q >= C::STORES_IN_FLIGHT) {
            bulk_wait_read<C::STORES_IN_FLIGHT>();
            mbar_arrive(bar(C::B_IN_FREE + (q - C::STORES_IN_FLIGHT) % C::IN_STAGES));
          }
        }
        __syncwarp();
      }
      if (p.done != nullptr) {
        // every store of this job has been performed: publish it to the pair that runs the panel's next iteration
        if (elect_one_sync()) {
          bulk_wait<0>();
          fence_proxy_async_all();
          red_release_gpu_add(p.done + job.panel, 1);
        }
        __syncwarp();
      }
    }
    if (elect_one_sync()) bulk_wait<0>();
    __syncwarp();
  }
  } else {
    // ================================ epilogue math ================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REGS_MATH));
    // this role's own copies of the thread index and the shared-memory base, pinned in registers (pin_u32)
    const uint32_t tid = pin_u32(threadIdx.x);
    const int warp = static_cast<int>(tid >> 5), lane = static_cast<int>(tid & 31);
    const uint32_t sbase_m = pin_u32(sbase);
    const uint32_t sY = sbase_m + C::OFF_Y, sIn = sbase_m + C::OFF_IN, bar0 = sbase_m + C::OFF_BAR;
    auto bar = [&](int idx) { return bar0 + 8 * idx; };
    const int group = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t sw64 = (row >> 1) & 3;
    const uint32_t sw32 = (row >> 2) & 1;
    UpdateArgs ua;
    ua.prox = p.prox, ua.group = p.group, ua.use_momentum = p.use_momentum;
    ua.eta = __ldg(p.scalars + 0), ua.theta = __ldg(p.scalars + 1);
    ua.want_stat = p.stat != nullptr;
    float stat_local = 0.f;
    Tracer trace(p, 2, warp == 4 && lane == 0);
    uint32_t q = 0;     // running sub-tile index (all sub-tiles of all jobs): group q % 3, stage q % IN_STAGES
    uint32_t yc = 0;    // running chunk index of the jobs with an R: y stage yc % 3
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t row_in = row * 16;            // this thread's 16 bytes inside a quad plane of a stage
    const long long cb_stride = row_blocks * (EPI_ARRAY_BYTES / 4);   // floats between column blocks of the state
    const bool fast_soft = ua.prox == 0 && ua.group <= 1;
    int t = 0;          // running atom tile: accumulator t & 1
    int r_jobs = 0;
    for (int pi = 0; pi < my_jobs; ++pi) {
      const Job job = job_at(pi);
      const bool has_prev = job.has_prev, has_prev2 = job.has_prev2;
      ua.in_mask = has_prev2 ? 5 : 1;
      ua.beta_prev = job.beta_prev, ua.beta_next = job.beta_next;
      if (has_prev2) wait_for_previous(job, false);   // a_{k-2} is read with plain loads by this warp
      const float bp_fast = has_prev2 ? ua.beta_prev : 0.f, bn_fast = ua.use_momentum ? ua.beta_next : 0.f;
      const bool do_r = job.do_r;
      // a_{k-2} of this group's NEXT sub-tile, loaded one sub-tile ahead (its sub-tiles inside a job are s_first,
      // s_first + 3, ...; sub-tile s is column block s of the panel's state: one pointer per job, one multiply-add per
      // sub-tile). Without an a_{k-2} (first iterations, ISTA) the registers stay zero and beta_prev is 0.
      const float* prev2 = pin_ptr(p.state[job.prev2] + row * 4 + state_offset(job.panel, 0));
      float4 pf[4];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) pf[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
      auto prefetch_prev2 = [&](int cb) {
        const float* src = prev2 + cb * cb_stride;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) pf[ch] = ldg_cg_v4(src + ch * (BLOCK_M * 4));
      };
      {
        const int s_first = (group + C::GROUPS - static_cast<int>(q % C::GROUPS)) % C::GROUPS;
        if (has_prev2 && s_first < state_subs) prefetch_prev2(s_first);
      }
      int s0 = 0;   // first sub-tile of the current tile inside the job
      for (int nt = 0; nt < NT; ++nt, ++t) {
        const int nsub = 2 * tile_chunks(nt);
        const int acc = t & 1;
        mbar_wait(bar(C::B_ACCG_FULL + acc), (t >> 1) & 1);
        tc_fence_after();
        trace(TR_E_BEGIN, pi * (NT + 1) + nt);
        const uint32_t t_row = lane_base + acc * IT_BN;
        const uint32_t drained_bar = bar(C::B_ACCG_EMPTY + acc);
        const uint32_t q0 = q;
        const int j_first = static_cast<int>((group + C::GROUPS - q0 % C::GROUPS) % C::GROUPS);
        const int j_last = (j_first < nsub) ? j_first + ((nsub - 1 - j_first) / C::GROUPS) * C::GROUPS : -1;
        if (j_last < 0) {
          __syncwarp();
          if (elect_one_sync()) {
            tc_fence_before();
            mbar_arrive_remote(drained_bar, 0);
          }
        }
        q = q0 + nsub;
        for (int j = j_first; j < nsub; j += C::GROUPS) {
          const uint32_t qq = q0 + j;
          const int e = qq % C::IN_STAGES;
          const int s = s0 + j;                    // sub-tile inside the job = column block
          const uint32_t chunk = yc + (s >> 1);
          const int ys = chunk % C::Y_STAGES;
          uint32_t v[16];
          trace(TR_E_SUB, j);
          if (!(ablate_of(p) & ABL_TMEM_LD)) tmem_ld16(t_row + j * EPI_COLS, v);
          else {
#pragma unroll
            for (int x = 0; x < 16; ++x) v[x] = 0u;
          }
          float in[3][16];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const float4 b = pf[ch];
            in[2][4 * ch + 0] = b.x, in[2][4 * ch + 1] = b.y, in[2][4 * ch + 2] = b.z, in[2][4 * ch + 3] = b.w;
          }
          if (has_prev2 && s + C::GROUPS < state_subs) prefetch_prev2(s + C::GROUPS);
          mbar_wait(bar(C::B_IN_FULL + e), (qq / C::IN_STAGES) & 1);
          trace(TR_E_IN, j);
          const uint32_t in_stage = sIn + e * C::IN_STAGE + row_in;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_prev && !(ablate_of(p) & ABL_STATE_LSU)) a = lds128(in_stage + ch * (BLOCK_M * 16));
            in[0][4 * ch + 0] = a.x, in[0][4 * ch + 1] = a.y, in[0][4 * ch + 2] = a.z, in[0][4 * ch + 3] = a.w;
            in[1][4 * ch + 0] = 0.f, in[1][4 * ch + 1] = 0.f, in[1][4 * ch + 2] = 0.f, in[1][4 * ch + 3] = 0.f;
          }
          tmem_ld_wait();
          trace(TR_E_LD, j);
          if (j == j_last) {   // this warp has drained its share of the accumulator
            __syncwarp();
            if (elect_one_sync()) {
              tc_fence_before();
              mbar_arrive_remote(drained_bar, 0);
            }
          }
          float outv[16], partv[16];
          if (fast_soft && !ua.want_stat)
            soft_update16<false, true, true, 0>(v, in, ua.eta, ua.theta, bp_fast, bn_fast, outv, partv, stat_local);
          else if (fast_soft)
            soft_update16<false, true, true, 1>(v, in, ua.eta, ua.theta, bp_fast, bn_fast, outv, partv, stat_local);
          else
            fista_update16(ua, v, in, outv, partv, stat_local);
          trace(TR_E_CMP, j);
          // a_k over a_{k-1}, in place: every thread reads and writes its own 4 x 16 bytes only
          if (!(ablate_of(p) & ABL_STATE_LSU)) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              sts128(in_stage + ch * (BLOCK_M * 16), outv[4 * ch], outv[4 * ch + 1], outv[4 * ch + 2], outv[4 * ch + 3]);
          }
          uint32_t yfull = 0;
          if (do_r) {
            // y_k parts straight into the A operand of R: K-major 64-byte rows (SWIZZLE_64B), this sub-tile is the
            // 32-byte half s & 1 of the row
            mbar_wait(bar(C::B_Y_EMPTY + ys), ((chunk / C::Y_STAGES) & 1) ^ 1);
            trace(TR_E_YW, j);
            const uint32_t ystage = sY + ys * C::Y_STAGE + row * (C::CHUNK * 2);
            const uint32_t c2 = 2 * (s & 1);
            if (!(ablate_of(p) & ABL_Y_STS)) split_parts16(partv, P, [&](int part, const uint32_t (&w32)[8]) {
              const uint32_t prow = ystage + part * C::Y_TILE;
              sts128u(prow + (((c2 + 0) ^ sw64) << 4), w32[0], w32[1], w32[2], w32[3]);
              sts128u(prow + (((c2 + 1) ^ sw64) << 4), w32[4], w32[5], w32[6], w32[7]);
            });
            yfull = bar(C::B_Y_FULL + ys);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one_sync()) {
            mbar_arrive(bar(C::B_OUT_FULL + e));
            if (do_r) mbar_arrive_remote(yfull, 0);
          }
          trace(TR_E_ARR, j);
        }
        trace(TR_E_END, pi * (NT + 1) + nt);
        s0 += nsub;
      }
      if (do_r) {
        yc += state_subs / 2;
        // ---- panel end: r_k = acc_r - x -> bf16 parts -> r_op[panel]; x staged by TMA like a state sub-tile
        mbar_wait(bar(C::B_ACCR_FULL), r_jobs & 1);
        tc_fence_after();
        trace(TR_E_BEGIN, pi * (NT + 1) + NT);
        ++r_jobs;
        const uint32_t t_row = lane_base + 2 * IT_BN;
        const uint32_t drained_bar = bar(C::B_ACCR_EMPTY);
        const uint32_t q0 = q;
        const int j_first = static_cast<int>((group + C::GROUPS - q0 % C::GROUPS) % C::GROUPS);
        const int j_last = (j_first < p.nsub_r) ? j_first + ((p.nsub_r - 1 - j_first) / C::GROUPS) * C::GROUPS : -1;
        if (j_last < 0) {
          __syncwarp();
          if (elect_one_sync()) {
            tc_fence_before();
            mbar_arrive_remote(drained_bar, 0);
          }
        }
        q = q0 + nsub_r_pad;
        for (int j = j_first; j < nsub_r_pad; j += C::GROUPS) {
          const uint32_t qq = q0 + j;
          const int e = qq % C::IN_STAGES;
          if (j >= p.nsub_r) {   // padding sub-tile: pass the stage on
            mbar_wait(bar(C::B_IN_FULL + e), (qq / C::IN_STAGES) & 1);
            __syncwarp();
            if (elect_one_sync()) mbar_arrive(bar(C::B_OUT_FULL + e));
            continue;
          }
          uint32_t v[16];
          tmem_ld16(t_row + j * EPI_COLS, v);
          mbar_wait(bar(C::B_IN_FULL + e), (qq / C::IN_STAGES) & 1);
          const uint32_t stage = sIn + e * C::IN_STAGE;
          float xin[16];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const float4 a = lds128(stage + row * 64 + ((ch ^ sw64) << 4));   // TMA box 16 x 128 fp32, SWIZZLE_64B
            xin[4 * ch + 0] = a.x, xin[4 * ch + 1] = a.y, xin[4 * ch + 2] = a.z, xin[4 * ch + 3] = a.w;
          }
          // the r parts go over the same stage, where a part row of this thread lies on other threads' x rows: all
          // four warps of the group must have read their x first
          named_bar_sync(1 + group, 128);
          tmem_ld_wait();
          if (j == j_last) {
            __syncwarp();
            if (elect_one_sync()) {
              tc_fence_before();
              mbar_arrive_remote(drained_bar, 0);
            }
          }
          float partv[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) partv[x] = __uint_as_float(v[x]) - xin[x];  // r_k = y_k Phi - x
          split_parts16(partv, P, [&](int part, const uint32_t (&w32)[8]) {
            const uint32_t prow = stage + part * EPI_PART_BYTES + row * 32;
            sts128u(prow + ((0 ^ sw32) << 4), w32[0], w32[1], w32[2], w32[3]);
            sts128u(prow + ((1 ^ sw32) << 4), w32[4], w32[5], w32[6], w32[7]);
          });
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one_sync()) mbar_arrive(bar(C::B_OUT_FULL + e));
        }
        trace(TR_E_END, pi * (NT + 1) + NT);
      }
    }
    if (p.stat) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) stat_local += __shfl_xor_sync(0xffffffffu, stat_local, o);
      if (lane == 0) atomicAdd(p.stat, static_cast<double>(stat_local));
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace vtc
