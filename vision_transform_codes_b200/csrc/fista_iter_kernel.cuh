// Panel-resident ISTA/FISTA iterations (sm_100a): the synthesis form with the operand y_k kept on chip; one launch runs
// one iteration or (persistent schedule, the default) all of them
//
//   grad = r_{k-1} Phi^T                       (analysis contraction,  K = D, N = S)        ista_fista.py:105-106
//   a_k  = prox(y_{k-1} - eta * grad) ,  y_k = a_k + beta_k (a_k - a_{k-1})                 ista_fista.py:107-133
//   r_k  = y_k Phi - x                         (synthesis contraction, K = S, N = D)
//
// with the next operand y_k never leaving the chip. A job = one iteration of one 256-row panel of patches, run by a CTA
// pair (cluster of 2, cta_group::2) that walks the panel's atoms in tiles of 128; the jobs (iteration, panel) of a launch
// are dealt round-robin to the pairs, and a panel's iteration k waits for the completion counter of its iteration k - 1
// (IterParams::done), which may have run on another pair:
//
//   tensor pipe   G(t): acc_g[t&1] (256 x 128, TMEM) = r_op[panel] * Phi[tile]^T           operands by TMA
//                 R(t): acc_r (256 x 256, TMEM)     += y_k[panel, tile] * Phi[tile]         A operand written to shared
//                                                                                            memory by the epilogue
//   epilogue      per 16-atom sub-tile: tcgen05.ld acc_g, fused update (shared with gemm_kernel.cuh), a_k -> TMA store,
//                 bf16 parts of y_k -> the y ring in the UMMA K-major SWIZZLE_64B layout -> R(t)
//   panel end     acc_r - x -> bf16 parts -> r_op[panel] (in place: every G of the panel has completed), read by the
//                 next launch
//
// Against the two-launch schedule (gemm_kernel.cuh: EPI_STORE then EPI_FISTA) this removes the HBM round trip of y_op
// (2 P bytes written + 2 P bytes read per code element and iteration) and hides the synthesis MMAs under the state
// stream. The MMA thread issues whichever of G(t+1) (K block by K block, as operands land) and R(t) (chunk by chunk, as
// the epilogue of tile t writes y_k) has its inputs ready, so both run while the epilogue of tile t streams its state;
// acc_g is double buffered (2 x 128 TMEM columns), acc_r takes the other 256.
//
// The fp32 inputs of a sub-tile (a_{k-1}, a_{k-2}: 2 x 8 KB) and its result share one ring: the math warps write a_k
// over the a_{k-2} slot of the stage they just read and the TMA store leaves from there, so there is no separate
// output ring to wait for and the whole budget goes into bytes in flight from HBM.
//
// Requirements (the host falls back to the two-launch schedule otherwise): D <= 256, P <= 2.
//
// Warp roles (4 + 4 * GROUPS + 2 warps):
//   0        TMA producer of the G operand ring (r_op K blocks + Phi tile halves), both CTAs
//   1        tcgen05.mma issuer of G (leader CTA)
//   2        TMEM allocator, then TMA stores (a_k sub-tiles, r_op parts)
//   3        TMA loader of the epilogue inputs (a_{k-1}, a_{k-2}; x at the panel end)
//   4 ..     epilogue math, GROUPS groups of four warps on sub-tiles round-robin
//   last - 1 TMA producer of the Phi^T chunk ring (B operand of R), both CTAs
//   last     tcgen05.mma issuer of R (leader CTA)
#pragma once
#include "gemm_kernel.cuh"

namespace vtc {

constexpr int IT_BN = 128;     // atoms per gradient tile (UMMA N of G)
constexpr int IT_RN = 256;     // UMMA N of R = padded pixel count
constexpr int IT_VARIANTS = 4;  // tuning variants (stage counts / math warps), VTC_B200_ITER_VARIANT

template <int P, int V>
struct IterCfg {
  static_assert(P == 1 || P == 2, "parts");
  static_assert(V >= 0 && V < IT_VARIANTS, "variant");
  // tuning variants (VTC_B200_ITER_VARIANT; default 0 for bf16x3, 1 for bf16); see profiles/README.md for what each measured
  //                     math groups   chunk   G (P=2 / P=1)   y   Phi^T   in/out stages
  //   0                     3          32       2 / 3          3   2       6
  //   1                     2          32       4 / 2          2   2       4 / 8
  //   2                     3          16       3 / 4          3   3       6
  //   3                     3          16       2 / 2          -   2 (x32) 9     y parts in the in/out stage (YIN)
  // epilogue math: GROUPS groups of four warps (one warp per TMEM lane quarter) on sub-tiles round-robin
  static constexpr int GROUPS = (V == 1) ? 2 : 3;
  static constexpr int MATH_WARPS = 4 * GROUPS;
  static constexpr int PT_WARP = 4 + MATH_WARPS;           // Phi^T chunk producer
  // bf16x3: G and R MMAs are issued by two warps (this one more); plain bf16 keeps the single polling issuer (same-box
  // A/B, profiles/README.md: two issuers -4 % for bf16x3, +17 % for plain bf16)
  static constexpr bool SPLIT_ISSUE = (P == 2);
  static constexpr int R_WARP = PT_WARP + 1;               // issuer of the R MMAs: the last warp (SPLIT_ISSUE)
  static constexpr int THREADS = 32 * (PT_WARP + 1 + (SPLIT_ISSUE ? 1 : 0));
  static constexpr int BK = (P == 1) ? 64 : 32;          // K extent of a G stage
  static constexpr int SPAN = BK * 2;
  static constexpr int A_TILE = BLOCK_M * SPAN;           // one part of this CTA's 128 rows of r_op
  static constexpr int B_TILE = (IT_BN / 2) * SPAN;       // one part of this CTA's 64 atoms of the Phi tile
  static constexpr int G_STAGE = P * (A_TILE + B_TILE);
  // atoms per y / Phi^T chunk (K extent of one group of R MMAs): 32 = two sub-tiles (64-byte rows, SWIZZLE_64B) or
  // 16 = one sub-tile (32-byte rows, SWIZZLE_32B: half the ring bytes, twice the handshakes)
  static constexpr int CHUNK = (V == 2 || V == 3) ? 16 : 32;
  static constexpr int SUBS = CHUNK / EPI_COLS;           // sub-tiles per chunk
  static constexpr int Y_TILE = BLOCK_M * CHUNK * 2;      // one part of y: 128 rows x CHUNK atoms
  static constexpr int Y_STAGE = P * Y_TILE;
  // atoms per Phi^T chunk: CHUNK, except YIN, which keeps 32-atom chunks (two 16-atom R chunks read one each: half the
  // TMA loads and ring handshakes) -- the A and B operands of an MMA may use different swizzle spans
  static constexpr int PT_CHUNK = (V == 3) ? 32 : CHUNK;
  static constexpr int PT_TILE = (IT_RN / 2) * PT_CHUNK * 2;  // one part of this CTA's 128 pixel rows of Phi^T
  static constexpr int PT_STAGE = P * PT_TILE;
  static constexpr int IN_STAGE = 2 * EPI_ARRAY_BYTES;
  static constexpr int OUT_SLOT = EPI_ARRAY_BYTES;        // the result (a_k fp32, or P bf16 part sub-tiles of r) is written
                                                          // over the second 8 KB of its input stage and stored from there
  static constexpr int STORES_IN_FLIGHT = 1;              // TMA stores whose shared-memory reads may still be pending
  // shared memory split between the rings (P = 2: G 24 KB, y 16 KB, Phi^T 16 KB, in/out 16 KB per stage)
  static constexpr int G_STAGES = (V == 0) ? (P == 2 ? 2 : 3) : (V == 1) ? (P == 2 ? 4 : 2) : (V == 3) ? 2 : (V == 2 && P == 2) ? 3 : 4;
  // YIN: no y ring. The y parts of a sub-tile (P x 128 rows x 32 B, SWIZZLE_32B) are written over the a_{k-1} slot of
  // the sub-tile's own in/out stage once the group has read it, and R takes its A operand from there; the stage goes
  // back to the loader when both the TMA store of a_k and the R MMAs have read it.
  static constexpr bool YIN = (V == 3);
  // every y stage must always be written by the same math groups (a group then sees the phases of the stage's
  // barrier strictly in order, like the in/out stages): Y_STAGES * SUBS is a multiple of GROUPS
  static constexpr int Y_STAGES = YIN ? 0 : (V == 1) ? 2 : 3;
  static constexpr int PT_STAGES = (V == 2) ? 3 : 2;
  static constexpr int IN_STAGES = (V == 1) ? (P == 2 ? 4 : 8) : (V == 3) ? 9 : 6;      // a multiple of GROUPS: fixed owner group per stage
  static constexpr int OFF_G = 0;
  static constexpr int OFF_Y = OFF_G + G_STAGES * G_STAGE;
  static constexpr int OFF_PT = OFF_Y + Y_STAGES * Y_STAGE;
  static constexpr int OFF_IN = OFF_PT + PT_STAGES * PT_STAGE;
  static constexpr int OFF_BAR = OFF_IN + IN_STAGES * IN_STAGE;
  // barrier indices
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
  static constexpr int B_OUT_FULL = B_IN_FREE + IN_STAGES;   // one per in/out stage
  static constexpr int B_YS_FULL = B_OUT_FULL + IN_STAGES;   // YIN: the y parts of the stage are written (leader's barrier)
  static constexpr int NUM_BARRIERS = B_YS_FULL + IN_STAGES;
  static constexpr int SMEM_TOTAL = OFF_BAR + NUM_BARRIERS * 8 + 16;
  static constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;
  static constexpr int NPAIRS = (P == 1) ? 1 : 3;
  static_assert(IN_STAGES % GROUPS == 0, "input stages must have a fixed owner group");
  static_assert((Y_STAGES * SUBS) % GROUPS == 0, "y stages must have fixed writer groups");
  // the panel-end sub-tiles are padded to a multiple of this with empty ones, so that the running sub-tile index (which
  // selects the math group and the in/out stage) and the running y chunk index stay congruent from panel to panel
  static constexpr int PANEL_END_PAD = 6;
  static_assert(PANEL_END_PAD % GROUPS == 0, "padding must keep the group assignment aligned");
  static_assert(P * EPI_PART_BYTES <= OUT_SLOT, "r parts must fit the output slot");
  static_assert(!YIN || (SUBS == 1 && P * BLOCK_M * 32 <= EPI_ARRAY_BYTES), "y parts must fit the a_{k-1} slot");
  static_assert(SMEM_ALLOC <= 232448, "over the 227 KB shared memory limit");
};

struct IterParams {
  CUtensorMap tmR;      // r_op (bf16 parts, tile-contiguous [part][Dp/BK][rows][BK]): box BK x 128, A operand of G
  CUtensorMap tmPhi;    // phi_op (S x parts*Dp, row-major): box BK x 64, B operand of G
  CUtensorMap tmPhiT;   // phiT_op (D x parts*Sp, row-major): box 32 x 128, SWIZZLE_64B, B operand of R
  CUtensorMap tmState[4];  // fp32 code arrays, box 16 x 128: [0] the starting point a_0, [1] a_k for odd k, [2] a_k for
                           // even k (a_k overwrites a_{k-2} in place), [3] where the final iterate goes
  CUtensorMap tmX;      // images (B x D fp32, row-major), box 16 x 128
  CUtensorMap tmROut;   // r_op as a store target: box 16 x 128, SWIZZLE_32B
  int num_panels;       // ceil(B / 256)
  int S;
  int num_n_tiles;      // ceil(S / 128)
  int kb_g;             // Dp / BK: K blocks of G (= column blocks per part of r_op)
  int phi_part_stride;  // Dp
  int phiT_part_stride; // Sp
  int nsub_r;           // Dp / 16: sub-tiles of r written at the panel end
  int r_block_w;        // BK of r_op's layout
  int state_blocked[4]; // tmState[i] is tile-contiguous (3-D map) instead of row-major
  // One launch runs iterations k_first .. k_first + k_count - 1 of every panel. The jobs (iteration, panel) are dealt
  // round-robin to the CTA pairs in iteration-major order, so every pair gets the same number of jobs whatever the panel
  // count (no quantisation of panels onto pairs), and a panel's state simply moves from pair to pair through HBM / L2;
  // job (k, panel) waits until both CTAs of job (k - 1, panel) have published `done[panel]`.
  int k_first, k_count;
  int k_final;          // the iteration whose output is the result: written to tmState[3], no r_k produced (nothing
                        // consumes it); INT_MAX when the caller stops on its own criterion
  const float* betas;   // device: betas[k] = FISTA momentum coefficient of iteration k, betas[0] = 0
  int* done;            // device, per panel: CTAs that have completed a job of this launch on it (nullptr: k_count == 1)
  int prox, group, use_momentum;
  const float* scalars;  // device: [0] = eta, [1] = theta
  double* stat;          // optional: += sum |a_k - a_{k-1}|
  unsigned long long* trace;  // optional (tools/iter_trace.py): 4 regions of 2048 words, [0] = event count, then
                              // (event id << 48 | SM clock) words, written by four threads of CTA 0
  int ablate;            // timing experiments only (VTC_B200_ABLATE, results are WRONG when non-zero): see AblateBits
};
// What a timing experiment leaves out of the kernel (tools/ablate.sh): the time that disappears with a piece is what
// that piece costs in situ.
// The switches exist only in a library built with -DVTC_ABLATE (tools/ab_build.sh): the shipped kernels carry none of
// the tests (they cost the plain-bf16 kernel 7-17 % when they were run-time flags).
#ifdef VTC_ABLATE
template <typename Params> __device__ __forceinline__ int ablate_of(const Params& p) { return p.ablate; }
#else
template <typename Params> __device__ __forceinline__ constexpr int ablate_of(const Params&) { return 0; }
#endif
enum AblateBits {
  ABL_STATE_LOAD = 1,    // no TMA loads of a_{k-1}, a_{k-2}
  ABL_STATE_STORE = 2,   // no TMA stores of a_k
  ABL_R_STREAM = 4,      // r_op K blocks loaded for the first atom tile of a job only (as if r_k were resident)
  ABL_STATE_LSU = 8,     // math warps neither read the state from shared memory nor write a_k to it
  ABL_R_MMA = 16,        // no R MMAs (commits only)
  ABL_G_MMA = 32,        // no G MMAs (commits only)
  ABL_Y_STS = 64,        // math warps do not write the y parts
  ABL_TMEM_LD = 128,     // math warps do not read the accumulator
  ABL_G_LOAD = 256,      // no TMA loads into the G operand ring at all (producer arrives only)
  ABL_PT_LOAD = 512,     // no TMA loads of the Phi^T chunks
};

// timeline events of CTA 0 for tools/iter_trace.py: id = kind << 8 | index
enum TraceKind { TR_G_BEGIN = 1, TR_G_END = 2, TR_R_ISSUE = 3, TR_E_BEGIN = 4, TR_E_SUB = 5, TR_E_END = 6, TR_G_LOAD = 7,
                 TR_START = 8, TR_STOP = 9, TR_E_IN = 10, TR_E_LD = 11, TR_E_CMP = 12, TR_E_YW = 13, TR_E_ARR = 14 };
// Each tracing thread owns a region of 2048 words (role = 0 G producer, 1 issuer, 2 math warp 4, 3 thread 0) and a private
// counter: plain stores, no atomics, so that the trace perturbs the timeline as little as possible.
// Compiled in only with -DVTC_TRACE (tools/ab_build.sh): a run-time "is tracing on" test at every trace point cost the
// math warps ~45 of their ~450 instructions per sub-tile and a tenth of their time (ncu stall samples on the
// BSSY / BRA / BSYNC of the tests), so the shipped kernels carry none.
#ifdef VTC_TRACE
struct Tracer {
  unsigned long long* base;
  uint32_t n;
  template <typename Params>
  __device__ __forceinline__ Tracer(const Params& p, int role, bool mine)
      : base((p.trace != nullptr && blockIdx.x == 0 && mine) ? p.trace + role * 2048 : nullptr), n(0) {}
  __device__ __forceinline__ void operator()(int kind, int index) {
    if (base != nullptr && n < 2047) {
      base[++n] = (static_cast<unsigned long long>((kind << 8) | (index & 255)) << 48) |
                  (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFFFFFull);
      base[0] = n;
    }
  }
};
#else
struct Tracer {
  template <typename Params>
  __device__ __forceinline__ Tracer(const Params&, int, bool) {}
  __device__ __forceinline__ void operator()(int, int) const {}
};
#endif

// completion flags between jobs of one launch (a panel's iteration k may run on another CTA pair than k - 1)
__device__ __forceinline__ int ld_acquire_gpu(const int* ptr) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* ptr, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// non-blocking tests (the MMA issuer polls several barriers)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
template <int P, int V>
__global__ void __launch_bounds__(IterCfg<P, V>::THREADS, 1) vtc_fista_iter_kernel(const __grid_constant__ IterParams p) {
  using C = IterCfg<P, V>;
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
  // jobs of this pair: tickets cluster_id, cluster_id + num_clusters, ... of the k_count x num_panels grid, k slowest
  const long long total_jobs = static_cast<long long>(p.k_count) * p.num_panels;
  const int my_jobs = total_jobs > cluster_id
                          ? static_cast<int>((total_jobs - cluster_id + num_clusters - 1) / num_clusters) : 0;
  const int my_tiles = my_jobs * NT;
  struct Job {
    int k, m0, panel, wait_target;
    bool do_r, has_prev2;
    int prev, prev2, out;   // indices into tmState
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
    j.has_prev2 = p.use_momentum != 0 && j.beta_prev != 0.f;
    j.prev = (j.k == 1) ? 0 : (((j.k - 1) & 1) ? 1 : 2);
    j.prev2 = (j.k <= 2) ? 0 : ((j.k & 1) ? 1 : 2);
    j.out = (j.k == p.k_final) ? 3 : ((j.k & 1) ? 1 : 2);
    return j;
  };
  // the data of job (k - 1, panel) must be complete before anything of job (k, panel) is read
  auto wait_for_previous = [&](const Job& j) {
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
      fence_proxy_async_all();   // the acquired data is read through TMA (async proxy)
    }
  };
  const int nsub_r_pad = (p.nsub_r + C::PANEL_END_PAD - 1) / C::PANEL_END_PAD * C::PANEL_END_PAD;
  // sub-tiles of tile nt: an even number (whole chunks); columns at or beyond S are zero everywhere (TMA fill)
  auto tile_chunks = [&](int nt) { return (min(IT_BN, p.S - nt * IT_BN) + C::CHUNK - 1) / C::CHUNK; };
  int job_tile_subs = 0;   // sub-tiles of the atom tiles of one job (without its panel end)
  for (int nt = 0; nt < NT; ++nt) job_tile_subs += C::SUBS * tile_chunks(nt);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmR);
    tma_prefetch_desc(&p.tmPhi);
    tma_prefetch_desc(&p.tmPhiT);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmState[i]);
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmROut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::G_STAGES; ++s) {
      // the leader's producer arrives (with the expected bytes of BOTH CTAs' loads); the peer's loads complete their
      // bytes on the same barrier and need no arrival of their own: a peer load of phase n + 1 can only be issued after
      // the commit that ended phase n, and bytes that land before the leader's expect_tx just leave the transaction
      // count negative until it is posted. (A remote arrive per K block put a cross-CTA round trip into every stage cycle.)
      mbar_init(bar(C::B_G_FULL + s), 1);
      mbar_init(bar(C::B_G_EMPTY + s), 1);   // multicast tcgen05.commit
    }
    for (int s = 0; s < C::PT_STAGES; ++s) {
      mbar_init(bar(C::B_PT_FULL + s), 1);
      mbar_init(bar(C::B_PT_EMPTY + s), 1);
    }
    for (int s = 0; s < C::Y_STAGES; ++s) {
      mbar_init(bar(C::B_Y_FULL + s), 2 * C::SUBS * 4);  // 2 CTAs x sub-tiles of a chunk x 4 warps (leader's barrier)
      mbar_init(bar(C::B_Y_EMPTY + s), 1);         // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar(C::B_ACCG_FULL + a), 1);
      mbar_init(bar(C::B_ACCG_EMPTY + a), 2 * C::MATH_WARPS);
    }
    mbar_init(bar(C::B_ACCR_FULL), 1);
    mbar_init(bar(C::B_ACCR_EMPTY), 2 * C::MATH_WARPS);
    for (int e = 0; e < C::IN_STAGES; ++e) {
      mbar_init(bar(C::B_IN_FULL + e), 1);
      // the storer, once the TMA store of the stage's result has read it (+ YIN: the commit of the R MMAs that read y)
      mbar_init(bar(C::B_IN_FREE + e), C::YIN ? 2 : 1);
      mbar_init(bar(C::B_YS_FULL + e), 2 * 4);   // YIN: 2 CTAs x the 4 warps of the stage's group, every use of the stage
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
  Tracer trace0(p, 3, threadIdx.x == 0);
  trace0(TR_START, 0);
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================================ G operand producer (both CTAs) ================================
    Tracer trace(p, 0, lane == 0);
    uint32_t it = 0;
    for (int ji = 0; ji < my_jobs; ++ji) {
      const Job job = job_at(ji);
      const int m0 = job.m0;
      wait_for_previous(job);   // r_op[panel] is the previous iteration's output
      for (int nt = 0; nt < NT; ++nt) {
        const int n0 = nt * IT_BN;
        for (int kb = 0; kb < p.kb_g; ++kb, ++it) {
          const int s = it % C::G_STAGES;
          const uint32_t ph = (it / C::G_STAGES) & 1;
          mbar_wait(bar(C::B_G_EMPTY + s), ph ^ 1);
          if (elect_one_sync()) {
            const uint32_t full = bar(C::B_G_FULL + s);
            const bool load_a = !((ablate_of(p) & ABL_R_STREAM) && nt > 0) && !(ablate_of(p) & ABL_G_LOAD);
            const bool load_b = !(ablate_of(p) & ABL_G_LOAD);
            if (leader) {
              if (load_b) mbar_arrive_expect_tx(full, load_a ? 2 * C::G_STAGE : 2 * P * C::B_TILE);
              else mbar_arrive(full);
            }
            const uint32_t dst = sG + s * C::G_STAGE;
            trace(TR_G_LOAD, nt * p.kb_g + kb);
#pragma unroll
            for (int q = 0; q < P; ++q)
              if (load_a) tma_load_3d_pair(dst + q * C::A_TILE, &p.tmR, full, 0, m0, q * p.kb_g + kb, kEvictNormal);
#pragma unroll
            for (int q = 0; q < P; ++q)
              if (load_b) tma_load_2d_pair(dst + P * C::A_TILE + q * C::B_TILE, &p.tmPhi, full, q * p.phi_part_stride + kb * C::BK,
                               n0 + cta_rank * (IT_BN / 2), kEvictLast);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == C::PT_WARP) {
    // ================================ Phi^T chunk producer (both CTAs) ================================
    {
      uint32_t it = 0;
      for (int ji = 0; ji < my_jobs; ++ji) {
        if (!job_at(ji).do_r) continue;
        for (int nt = 0; nt < NT; ++nt) {
          const int nch = (tile_chunks(nt) * C::CHUNK + C::PT_CHUNK - 1) / C::PT_CHUNK;
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
                                 q * p.phiT_part_stride + nt * IT_BN + c * C::PT_CHUNK, cta_rank * (IT_RN / 2), kEvictLast);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 && !C::SPLIT_ISSUE) {
    // ================================ MMA issuer of G and R (leader CTA), plain bf16 ================================
    if (leader) {
      constexpr uint32_t idesc_g = make_idesc_bf16(PAIR_M, IT_BN);
      constexpr uint32_t idesc_r = make_idesc_bf16(PAIR_M, IT_RN);
      const uint32_t acc_r = tmem_base + 2 * IT_BN;
      Tracer trace(p, 1, lane == 0);
      // Two instruction streams share the (in-order) tensor pipe: G(t), K block by K block as its operands land, and
      // R(t), chunk by chunk as the epilogue of tile t writes y_k. Whichever has its inputs ready is issued next
      // (non-blocking barrier tests), so a G waiting for an L2 round trip never holds up the R chunks that free the
      // y ring, and vice versa. G runs at most two tiles ahead of the epilogue (acc_g is double buffered).
      int g_tile = 0, g_kb = 0;
      uint32_t g_it = 0, g_accumulate = 0;
      int r_tile = 0, r_chunk = 0;
      uint32_t r_it = 0;
      uint32_t r_pt = 0;  // running Phi^T chunk index
      uint32_t r_q = 0;   // running sub-tile index of the next R chunk (YIN: its in/out stage is r_q % IN_STAGES)
      uint32_t idle = 0;
      int r_job = -1;
      bool r_job_do_r = false;
      while (g_tile < my_tiles || r_tile < my_tiles) {
        bool progressed = false;
        if (r_tile < my_tiles && r_tile / NT != r_job) {   // entering a job: does it produce r_k at all?
          r_job = r_tile / NT;
          r_job_do_r = job_at(r_job).do_r;
        }
        if (r_tile < my_tiles && !r_job_do_r) {
          r_tile += NT;   // the final iteration has no R
          r_q += job_tile_subs;
          continue;
        }
        if (r_tile < my_tiles) {
          const int pi = r_tile / NT, nt = r_tile % NT;
          const bool first = (nt == 0 && r_chunk == 0);
          constexpr int R_PER_PT = C::PT_CHUNK / C::CHUNK;   // R chunks served by one Phi^T chunk
          const int ps = r_pt % C::PT_STAGES;
          // A operand of this chunk: a stage of the y ring, or (YIN) the a_{k-1} slot of the sub-tile's in/out stage
          uint32_t ystage, a_part_bytes, y_release;
          bool ready;
          if constexpr (C::YIN) {
            const int e = r_q % C::IN_STAGES;
            ready = mbar_test_wait(bar(C::B_YS_FULL + e), (r_q / C::IN_STAGES) & 1);
            ystage = sIn + e * C::IN_STAGE, a_part_bytes = BLOCK_M * 32, y_release = bar(C::B_IN_FREE + e);
          } else {
            const int ys = r_it % C::Y_STAGES;
            ready = mbar_test_wait(bar(C::B_Y_FULL + ys), (r_it / C::Y_STAGES) & 1);
            ystage = sY + ys * C::Y_STAGE, a_part_bytes = C::Y_TILE, y_release = bar(C::B_Y_EMPTY + ys);
          }
          ready = ready && mbar_test_wait(bar(C::B_PT_FULL + ps), (r_pt / C::PT_STAGES) & 1);
          // the panel-end epilogue of the previous panel must have drained acc_r before it is overwritten
          if (ready && first) ready = mbar_test_wait(bar(C::B_ACCR_EMPTY), (pi & 1) ^ 1);
          if (ready) {
            tc_fence_after();
            const uint32_t pstage = sPT + ps * C::PT_STAGE;
            const bool last = (nt == NT - 1) && (r_chunk == tile_chunks(nt) - 1);
            // the Phi^T chunk is done with after its last R chunk (the last chunk of a tile may use only its first half)
            const bool pt_done = (r_chunk % R_PER_PT == R_PER_PT - 1) || (r_chunk == tile_chunks(nt) - 1);
            const uint32_t pt_koff = 2 * (C::CHUNK / UMMA_K) * (r_chunk % R_PER_PT);   // in 16-byte units
            if (elect_one_sync()) {
              trace(TR_R_ISSUE, r_tile * 4 + r_chunk);
              uint32_t accumulate = first ? 0u : 1u;
#pragma unroll
              for (int pr = 0; pr < C::NPAIRS; ++pr) {
                const uint64_t adesc = make_kmajor_desc(ystage + pair_a(P, pr) * a_part_bytes, C::CHUNK * 2);
                const uint64_t bdesc = make_kmajor_desc(pstage + pair_b(P, pr) * C::PT_TILE, C::PT_CHUNK * 2) + pt_koff;
#pragma unroll
                for (int k = 0; k < C::CHUNK / UMMA_K; ++k) {
                  umma_bf16_pair(acc_r, adesc + 2 * k, bdesc + 2 * k, idesc_r, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit_pair(y_release, 3);
              if (pt_done) umma_commit_pair(bar(C::B_PT_EMPTY + ps), 3);
              if (last) umma_commit_pair(bar(C::B_ACCR_FULL), 3);
            }
            __syncwarp();
            ++r_it;
            ++r_q;
            if (pt_done) ++r_pt;
            if (++r_chunk == tile_chunks(nt)) {
              r_chunk = 0, ++r_tile;
              if (r_tile % NT == 0) r_q += nsub_r_pad;   // the sub-tiles of this job's panel end carry no y
            }
            progressed = true;
          }
        }
        if (g_tile < my_tiles) {
          const int acc = g_tile & 1;
          const int s = g_it % C::G_STAGES;
          bool ready = mbar_test_wait(bar(C::B_G_FULL + s), (g_it / C::G_STAGES) & 1);
          if (ready && g_kb == 0) ready = mbar_test_wait(bar(C::B_ACCG_EMPTY + acc), ((g_tile >> 1) & 1) ^ 1);
          if (ready) {
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * IT_BN;
            const uint32_t stage = sG + s * C::G_STAGE;
            if (g_kb == 0) g_accumulate = 0;
            const bool last = (g_kb == p.kb_g - 1);
            if (elect_one_sync()) {
              if (g_kb == 0) trace(TR_G_BEGIN, g_tile);
              if (last) trace(TR_G_END, g_tile);
              uint32_t accumulate = g_accumulate;
#pragma unroll
              for (int pr = 0; pr < C::NPAIRS; ++pr) {
                const uint64_t adesc = make_kmajor_desc(stage + pair_a(P, pr) * C::A_TILE, C::SPAN);
                const uint64_t bdesc = make_kmajor_desc(stage + P * C::A_TILE + pair_b(P, pr) * C::B_TILE, C::SPAN);
#pragma unroll
                for (int k = 0; k < C::BK / UMMA_K; ++k) {
                  umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_g, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit_pair(bar(C::B_G_EMPTY + s), 3);
              if (last) umma_commit_pair(bar(C::B_ACCG_FULL + acc), 3);
            }
            __syncwarp();
            g_accumulate = 1;
            ++g_it;
            if (++g_kb == p.kb_g) g_kb = 0, ++g_tile;
            progressed = true;
          }
        }
        if (progressed) {
          idle = 0;
        } else {
          // nothing ready: back off for a few tens of ns so the polling does not take issue slots from the math warps
          asm volatile("nanosleep.u32 32;" ::: "memory");
        }
        if (!progressed && ++idle > (1u << 24)) {
          if (lane == 0)
            printf("vtc_b200: MMA issuer stalled (block %d, G tile %d kb %d, R tile %d chunk %d of %d tiles)\n",
                   (int)blockIdx.x, g_tile, g_kb, r_tile, r_chunk, my_tiles);
          __trap();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ G MMA issuer (leader CTA) ================================
    // G(t) = r_{k-1}[panel] * Phi[tile]^T into acc_g[t & 1], K block by K block as the operands land; at most two tiles
    // ahead of the epilogue (acc_g is double buffered). G and R are issued by two different warps (this one and
    // R_WARP): each blocks on its own barriers (a hardware-suspended try_wait wakes in ~60 cycles, a test_wait poll
    // costs ~150), where a single thread polling both streams spent ~1400 cycles per (R chunk, G K block) pair on
    // barrier tests and commits -- 5.6 us per atom tile with nothing else to do (profiles/README.md, round 2
    // ablations). The two accumulate into different TMEM columns, so no order between the two streams is needed;
    // tcgen05.commit tracks the MMAs of the issuing thread only.
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
  } else if (C::SPLIT_ISSUE && warp == C::R_WARP) {
    // ================================ R MMA issuer (leader CTA) ================================
    // R(t): acc_r += y_k[panel, tile] * Phi[tile], chunk by chunk as the epilogue of tile t writes y_k
    if (leader) {
      constexpr uint32_t idesc_r = make_idesc_bf16(PAIR_M, IT_RN);
      constexpr int R_PER_PT = C::PT_CHUNK / C::CHUNK;   // R chunks served by one Phi^T chunk
      const uint32_t acc_r = tmem_base + 2 * IT_BN;
      Tracer trace(p, 1, false);
      uint32_t r_it = 0;  // running R chunk index (y ring)
      uint32_t r_pt = 0;  // running Phi^T chunk index
      uint32_t r_q = 0;   // running sub-tile index of the next R chunk (YIN: its in/out stage is r_q % IN_STAGES)
      int r_jobs = 0;     // jobs with an R so far (parity of the acc_r barriers)
      for (int pi = 0; pi < my_jobs; ++pi) {
        if (!job_at(pi).do_r) {   // the final iteration has no R (and no panel end)
          r_q += job_tile_subs;
          continue;
        }
        for (int nt = 0; nt < NT; ++nt) {
          const int nch = tile_chunks(nt);
          for (int r_chunk = 0; r_chunk < nch; ++r_chunk, ++r_it, ++r_q) {
            const bool first = (nt == 0 && r_chunk == 0);
            const int ps = r_pt % C::PT_STAGES;
            // A operand of this chunk: a stage of the y ring, or (YIN) the a_{k-1} slot of the sub-tile's in/out stage
            uint32_t ystage, a_part_bytes, y_release;
            if constexpr (C::YIN) {
              const int e = r_q % C::IN_STAGES;
              mbar_wait(bar(C::B_YS_FULL + e), (r_q / C::IN_STAGES) & 1);
              ystage = sIn + e * C::IN_STAGE, a_part_bytes = BLOCK_M * 32, y_release = bar(C::B_IN_FREE + e);
            } else {
              const int ys = r_it % C::Y_STAGES;
              mbar_wait(bar(C::B_Y_FULL + ys), (r_it / C::Y_STAGES) & 1);
              ystage = sY + ys * C::Y_STAGE, a_part_bytes = C::Y_TILE, y_release = bar(C::B_Y_EMPTY + ys);
            }
            if (r_chunk % R_PER_PT == 0) mbar_wait(bar(C::B_PT_FULL + ps), (r_pt / C::PT_STAGES) & 1);
            // the panel-end epilogue of the previous job with an R must have drained acc_r before it is overwritten
            if (first) mbar_wait(bar(C::B_ACCR_EMPTY), (r_jobs & 1) ^ 1);
            tc_fence_after();
            const uint32_t pstage = sPT + ps * C::PT_STAGE;
            const bool last = (nt == NT - 1) && (r_chunk == nch - 1);
            // the Phi^T chunk is done with after its last R chunk (the last chunk of a tile may use only its first half)
            const bool pt_done = (r_chunk % R_PER_PT == R_PER_PT - 1) || (r_chunk == nch - 1);
            const uint32_t pt_koff = 2 * (C::CHUNK / UMMA_K) * (r_chunk % R_PER_PT);   // in 16-byte units
            if (elect_one_sync()) {
              uint32_t accumulate = first ? 0u : 1u;
#pragma unroll
              for (int pr = 0; pr < C::NPAIRS; ++pr) {
                const uint64_t adesc = make_kmajor_desc(ystage + pair_a(P, pr) * a_part_bytes, C::CHUNK * 2);
                const uint64_t bdesc = make_kmajor_desc(pstage + pair_b(P, pr) * C::PT_TILE, C::PT_CHUNK * 2) + pt_koff;
#pragma unroll
                for (int k = 0; k < C::CHUNK / UMMA_K; ++k) {
                  if (!(ablate_of(p) & ABL_R_MMA)) umma_bf16_pair(acc_r, adesc + 2 * k, bdesc + 2 * k, idesc_r, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit_pair(y_release, 3);
              if (pt_done) umma_commit_pair(bar(C::B_PT_EMPTY + ps), 3);
              if (last) umma_commit_pair(bar(C::B_ACCR_FULL), 3);
            }
            __syncwarp();
            if (pt_done) ++r_pt;
          }
        }
        r_q += nsub_r_pad;   // the sub-tiles of this job's panel end carry no y
        ++r_jobs;
      }
    }
  } else if (warp == 3) {
    // ================================ epilogue input loader ================================
    uint32_t q = 0;
    for (int ji = 0; ji < my_jobs; ++ji) {
      const Job job = job_at(ji);
      const int m0 = job.m0;
      const uint32_t state_bytes = (job.has_prev2 ? 2 : 1) * EPI_ARRAY_BYTES;
      wait_for_previous(job);   // a_{k-1} (and the slot a_k overwrites) belong to the previous iterations
      for (int nt = 0; nt <= NT; ++nt) {
        const bool panel_end = (nt == NT);
        if (panel_end && !job.do_r) break;
        const int nsub = panel_end ? nsub_r_pad : C::SUBS * tile_chunks(nt);
        for (int j = 0; j < nsub; ++j, ++q) {
          const int e = q % C::IN_STAGES;
          mbar_wait(bar(C::B_IN_FREE + e), ((q / C::IN_STAGES) & 1) ^ 1);
          if (elect_one_sync()) {
            const uint32_t full = bar(C::B_IN_FULL + e);
            const uint32_t dst = sIn + e * C::IN_STAGE;
            if (panel_end && j >= p.nsub_r) {
              mbar_arrive(full);  // padding sub-tile: nothing to load
            } else if (panel_end) {
              mbar_arrive_expect_tx(full, EPI_ARRAY_BYTES);
              tma_load_2d(dst, &p.tmX, full, j * EPI_COLS, m0, kEvictNormal);
            } else if (ablate_of(p) & ABL_STATE_LOAD) {
              mbar_arrive(full);
            } else {
              mbar_arrive_expect_tx(full, state_bytes);
              const int col = nt * IT_BN + j * EPI_COLS;
              for (int i = 0; i < (job.has_prev2 ? 2 : 1); ++i) {
                const int src = i == 0 ? job.prev : job.prev2;
                const uint32_t d = dst + i * EPI_ARRAY_BYTES;
                if (p.state_blocked[src]) tma_load_3d(d, &p.tmState[src], full, 0, m0, col / EPI_COLS, kEvictNormal);
                else tma_load_2d(d, &p.tmState[src], full, col, m0, kEvictNormal);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    // ================================ epilogue storer ================================
    uint32_t q = 0;
    uint32_t prev_has_y = 0;   // bit i: did the sub-tile i + 1 stores ago put y parts for R into its stage (YIN)
    for (int ji = 0; ji < my_jobs; ++ji) {
      const Job job = job_at(ji);
      const int m0 = job.m0;
      for (int nt = 0; nt <= NT; ++nt) {
        const bool panel_end = (nt == NT);
        if (panel_end && !job.do_r) break;
        const int nsub = panel_end ? nsub_r_pad : C::SUBS * tile_chunks(nt);
        for (int j = 0; j < nsub; ++j, ++q) {
          const int e = q % C::IN_STAGES;
          mbar_wait(bar(C::B_OUT_FULL + e), (q / C::IN_STAGES) & 1);
          const uint32_t src = sIn + e * C::IN_STAGE + EPI_ARRAY_BYTES;
          if (elect_one_sync()) {
            if (panel_end && j >= p.nsub_r) {
              // padding sub-tile: nothing to store (the empty bulk group below keeps the recycling uniform)
            } else if (panel_end) {
              const int col = j * EPI_COLS;
#pragma unroll
              for (int part = 0; part < P; ++part)
                tma_store_3d(&p.tmROut, src + part * EPI_PART_BYTES, col % p.r_block_w, m0,
                             part * p.kb_g + col / p.r_block_w);
            } else if (!(ablate_of(p) & ABL_STATE_STORE)) {
              const int col = nt * IT_BN + j * EPI_COLS;
              if (p.state_blocked[job.out]) tma_store_3d(&p.tmState[job.out], src, 0, m0, col / EPI_COLS);
              else tma_store_2d(&p.tmState[job.out], src, col, m0);
            }
            bulk_commit();
            if (q >= C::STORES_IN_FLIGHT) {
              // all but the STORES_IN_FLIGHT most recent stores have read their stage: hand the oldest back to the loader
              bulk_wait_read<C::STORES_IN_FLIGHT>();
              const uint32_t free_bar = bar(C::B_IN_FREE + (q - C::STORES_IN_FLIGHT) % C::IN_STAGES);
              mbar_arrive(free_bar);
              // YIN: a stage whose a_{k-1} slot holds no y for R (panel end, final iteration) gets the MMA's arrival here
              if (C::YIN && !((prev_has_y >> (C::STORES_IN_FLIGHT - 1)) & 1u)) mbar_arrive(free_bar);
            }
          }
          __syncwarp();
          prev_has_y = (prev_has_y << 1) | ((!panel_end && job.do_r) ? 1u : 0u);
        }
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
  } else if (warp >= 4 && warp < 4 + C::MATH_WARPS) {
    // ================================ epilogue math ================================
    const uint32_t group = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t sw64 = (row >> 1) & 3;
    const uint32_t sw32 = (row >> 2) & 1;
    UpdateArgs ua;
    ua.prox = p.prox, ua.group = p.group, ua.use_momentum = p.use_momentum;
    ua.eta = __ldg(p.scalars + 0), ua.theta = __ldg(p.scalars + 1);
    ua.want_stat = p.stat != nullptr;
    Tracer trace(p, 2, warp == 4 && lane == 0);
    float stat_local = 0.f;
    uint32_t q = 0, yc = 0;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int t = 0;
    for (int pi = 0; pi < my_jobs; ++pi) {
      const Job job = job_at(pi);
      const bool has_prev = job.has_prev2;
      ua.in_mask = has_prev ? 5 : 1;
      ua.beta_prev = job.beta_prev, ua.beta_next = job.beta_next;
      for (int nt = 0; nt <= NT; ++nt) {
        const bool panel_end = (nt == NT);
        if (panel_end && !job.do_r) break;
        const int nsub = panel_end ? p.nsub_r : C::SUBS * tile_chunks(nt);   // real sub-tiles
        const int nsub_all = panel_end ? nsub_r_pad : nsub;                  // + padding at the panel end
        uint32_t t_row, drained_bar;
        if (panel_end) {
          mbar_wait(bar(C::B_ACCR_FULL), pi & 1);
          t_row = lane_base + 2 * IT_BN;
          drained_bar = bar(C::B_ACCR_EMPTY);
        } else {
          const int acc = t & 1;
          mbar_wait(bar(C::B_ACCG_FULL + acc), (t >> 1) & 1);
          t_row = lane_base + acc * IT_BN;
          drained_bar = bar(C::B_ACCG_EMPTY + acc);
          ++t;
        }
        tc_fence_after();
        trace(TR_E_BEGIN, pi * (NT + 1) + nt);
        // this group's sub-tiles of the tile: j_first, j_first + GROUPS, ... (running index q0 + j congruent to the
        // group), walked directly -- testing every j cost 8 % of the math warps' time in round 1's profile
        const uint32_t q0 = q;
        const int j_first = static_cast<int>((group + C::GROUPS - q0 % C::GROUPS) % C::GROUPS);
        // last REAL sub-tile of this tile that belongs to this group (-1: none)
        const int j_last = (j_first < nsub) ? j_first + ((nsub - 1 - j_first) / C::GROUPS) * C::GROUPS : -1;
        if (j_last < 0) {
          __syncwarp();
          if (lane == 0) {
            tc_fence_before();
            mbar_arrive_remote(drained_bar, 0);
          }
        }
        q = q0 + nsub_all;   // (the loop below uses its own running index qq)
        for (int j = j_first; j < nsub_all; j += C::GROUPS) {
          const uint32_t qq = q0 + j;
          const uint32_t chunk = yc + j / C::SUBS;  // y chunk of this sub-tile (state tiles only)
          const int e = qq % C::IN_STAGES;
          if (j >= nsub) {  // padding sub-tile of the panel end: pass the stage on
            mbar_wait(bar(C::B_IN_FULL + e), (qq / C::IN_STAGES) & 1);
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(bar(C::B_OUT_FULL + e));
              if (C::YIN) mbar_arrive_remote(bar(C::B_YS_FULL + e), 0);   // every use of a stage completes a phase
            }
            continue;
          }
          uint32_t v[16];
          trace(TR_E_SUB, j);
          if (!(ablate_of(p) & ABL_TMEM_LD)) tmem_ld16(t_row + j * EPI_COLS, v);
          else {
#pragma unroll
            for (int x = 0; x < 16; ++x) v[x] = 0u;
          }
          mbar_wait(bar(C::B_IN_FULL + e), (qq / C::IN_STAGES) & 1);
          trace(TR_E_IN, j);
          tmem_ld_wait();
          trace(TR_E_LD, j);
          if (j == j_last) {
            __syncwarp();
            if (lane == 0) {
              tc_fence_before();
              mbar_arrive_remote(drained_bar, 0);
            }
          }
          const uint32_t in_stage = sIn + e * C::IN_STAGE;
          float in[3][16];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!(ablate_of(p) & ABL_STATE_LSU) || panel_end) a = lds128(in_stage + row * 64 + ((ch ^ sw64) << 4));
            in[0][4 * ch + 0] = a.x, in[0][4 * ch + 1] = a.y, in[0][4 * ch + 2] = a.z, in[0][4 * ch + 3] = a.w;
            in[1][4 * ch + 0] = 0.f, in[1][4 * ch + 1] = 0.f, in[1][4 * ch + 2] = 0.f, in[1][4 * ch + 3] = 0.f;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_prev && !panel_end && !(ablate_of(p) & ABL_STATE_LSU)) b = lds128(in_stage + EPI_ARRAY_BYTES + row * 64 + ((ch ^ sw64) << 4));
            in[2][4 * ch + 0] = b.x, in[2][4 * ch + 1] = b.y, in[2][4 * ch + 2] = b.z, in[2][4 * ch + 3] = b.w;
          }
          float outv[16], partv[16];
          if (panel_end) {
#pragma unroll
            for (int x = 0; x < 16; ++x) partv[x] = __uint_as_float(v[x]) - in[0][x];  // r_k = y_k Phi - x
          } else {
            // the common case (scalar soft threshold, synthesis form: the accumulator is the whole gradient) without
            // the run-time dispatch of fista_update16: one branch-free path, where "no a_{k-2}" / "no momentum" are
            // beta = 0 (a + 0 * (a - 0) is a, bit for bit), which in[2] = 0 and the job's betas already encode
            if (ua.prox == 0 && ua.group <= 1)
              soft_update16<false, true, true>(v, in, ua.eta, ua.theta, has_prev ? ua.beta_prev : 0.f,
                                               ua.use_momentum ? ua.beta_next : 0.f, outv, partv, stat_local,
                                               ua.want_stat);
            else
              fista_update16(ua, v, in, outv, partv, stat_local);
          }
          trace(TR_E_CMP, j);
          // the result goes over the second input slot of the same stage (every thread has read its own row of it):
          // no output ring to wait for; the storer hands the stage back to the loader once the TMA store has read it
          const uint32_t out_stage = in_stage + EPI_ARRAY_BYTES;
          if (panel_end) {
            split_parts16(partv, P, [&](int part, const uint32_t (&w32)[8]) {
              const uint32_t prow = out_stage + part * EPI_PART_BYTES + row * 32;
              sts128u(prow + ((0 ^ sw32) << 4), w32[0], w32[1], w32[2], w32[3]);
              sts128u(prow + ((1 ^ sw32) << 4), w32[4], w32[5], w32[6], w32[7]);
            });
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(bar(C::B_OUT_FULL + e));
              if (C::YIN) mbar_arrive_remote(bar(C::B_YS_FULL + e), 0);
            }
          } else {
            if (!(ablate_of(p) & ABL_STATE_LSU)) {
#pragma unroll
              for (int ch = 0; ch < 4; ++ch)
                sts128(out_stage + row * 64 + ((ch ^ sw64) << 4), outv[4 * ch], outv[4 * ch + 1], outv[4 * ch + 2],
                       outv[4 * ch + 3]);
            }
            if constexpr (C::YIN) {
              if (job.do_r) {
                // y_k parts over the a_{k-1} slot of this stage, as the A operand of R (K-major, 32-byte rows,
                // SWIZZLE_32B). A y row lands on another thread's a_{k-1} row: all four warps of the group must have
                // read their inputs first.
                named_bar_sync(1 + group, 128);
                trace(TR_E_YW, j);
                const uint32_t yrow = in_stage + row * 32;
                split_parts16(partv, P, [&](int part, const uint32_t (&w32)[8]) {
                  const uint32_t prow = yrow + part * (BLOCK_M * 32);
                  sts128u(prow + ((0 ^ sw32) << 4), w32[0], w32[1], w32[2], w32[3]);
                  sts128u(prow + ((1 ^ sw32) << 4), w32[4], w32[5], w32[6], w32[7]);
                });
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(bar(C::B_OUT_FULL + e));
                mbar_arrive_remote(bar(C::B_YS_FULL + e), 0);
              }
            } else {
            uint32_t yfull = 0;
            if (job.do_r) {
              // y_k parts straight into the A operand of R, K-major with rows of CHUNK * 2 bytes in the matching TMA /
              // UMMA swizzle: 64-byte rows (this sub-tile is the 32-byte half j % 2 of the row) or 32-byte rows
              const int ys = chunk % C::Y_STAGES;
              mbar_wait(bar(C::B_Y_EMPTY + ys), ((chunk / C::Y_STAGES) & 1) ^ 1);
              trace(TR_E_YW, j);
              const uint32_t ystage = sY + ys * C::Y_STAGE + row * (C::CHUNK * 2);
              const uint32_t c0 = 2 * (j % C::SUBS);
              const uint32_t sw = (C::CHUNK == 32) ? sw64 : sw32;
              if (!(ablate_of(p) & ABL_Y_STS)) split_parts16(partv, P, [&](int part, const uint32_t (&w32)[8]) {
                const uint32_t prow = ystage + part * C::Y_TILE;
                sts128u(prow + (((c0 + 0) ^ sw) << 4), w32[0], w32[1], w32[2], w32[3]);
                sts128u(prow + (((c0 + 1) ^ sw) << 4), w32[4], w32[5], w32[6], w32[7]);
              });
              yfull = bar(C::B_Y_FULL + ys);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(bar(C::B_OUT_FULL + e));
              if (job.do_r) mbar_arrive_remote(yfull, 0);
            }
            }
            trace(TR_E_ARR, j);
          }
        }
        trace(TR_E_END, pi * (NT + 1) + nt);
        if (!panel_end) yc += nsub / C::SUBS;
      }
    }
    if (p.stat) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) stat_local += __shfl_xor_sync(0xffffffffu, stat_local, o);
      if (lane == 0) atomicAdd(p.stat, static_cast<double>(stat_local));
    }
  }

  __syncwarp();
  trace0(TR_STOP, 0);
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace vtc
