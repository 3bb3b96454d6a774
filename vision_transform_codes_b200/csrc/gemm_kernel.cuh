// Persistent, warp-specialised tcgen05 GEMM on CTA pairs (cta_group::2) with a fused, TMA-staged epilogue (sm_100a).
//
//   D[M,N] = sum over (pa, pb) in PAIRS  A_pa[M,K] * B_pb[N,K]^T            bf16 operands, fp32 accumulate in TMEM
//
// A and B are K-major bf16 "parts" matrices: each row holds P parts of an fp32 matrix side by side (hi | mid | lo of a
// bf16 split, each padded to a multiple of 64 columns). P = 1: plain bf16 (1 product), P = 2: bf16x3 (hi*lo + lo*hi +
// hi*hi, error ~2^-17), P = 3: bf16x6 (~fp32). For every K block ALL parts of A and B are staged once and every
// product of the list is issued from that stage, so the x3 / x6 modes move 2x / 3x the operand bytes of plain bf16,
// not 3x / 6x.
//
// One CTA pair (cluster of 2, one CTA per SM) owns a 256 x 256 output tile: each CTA holds its own 128 rows of A and
// half (128 rows) of the B tile, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) and both tensor cores read
// both halves of B. Accumulators are double-buffered in TMEM (2 x 256 columns).
//
// Warp roles (384 threads):
//   warp 0 lane 0 : TMA producer of the operand ring (full / empty mbarriers; full lives in the leader CTA)
//   warp 1 lane 0 : tcgen05.mma issuer (leader CTA only); commits are multicast to both CTAs
//   warp 2        : TMEM allocator; lane 0 then drives the TMA stores of the epilogue (out ring)
//   warp 3 lane 0 : TMA loader of the epilogue's fp32 state tiles (in ring), running ahead of the math warps
//   warps 4..11   : epilogue math, two groups of four warps taking 16-column sub-tiles round-robin:
//                   tcgen05.ld -> fused update -> swizzled st.shared -> TMA store
//
// Epilogues:
//   EPI_STORE : out = acc [- in0]; optionally also emitted as bf16 parts (split of the fp32 value)
//   EPI_FISTA : one ISTA/FISTA iteration of vision_transform_codes/analysis_transforms/fully_connected/
//               ista_fista.py:105-133 (and subspace_ista_fista.py:144-169 for group shrinkage):
//                 y   = a_k + beta_prev * (a_k - a_km1)            (rebuilt in fp32, never stored)
//                 u   = y - eta * (acc - b)                        acc = y_op * G^T from the tensor cores
//                 a   = prox(u, theta)
//                 y'  = a + beta_next * (a - a_k)  -> bf16 parts   (next iteration's A operand)
#pragma once
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace vtc {

constexpr int BLOCK_M = 128;        // rows per CTA
constexpr int PAIR_M = 256;         // rows per CTA pair (UMMA M)
constexpr int BLOCK_N = 256;        // default UMMA N (a kernel template parameter BN: 256 or 128); each CTA stages half
constexpr int UMMA_K = 16;
constexpr int EPI_COLS = 16;        // epilogue sub-tile width (fp32 columns)
constexpr int MAX_PARTS = 3;
constexpr int MAX_SEGMENTS = 16;     // kernel taps per image in units of the stride (e.g. 16x16 kernel, stride 8: 4)
constexpr int NUM_MATH_GROUPS = 2;   // groups of four warps (one per TMEM lane quarter) taking sub-tiles round-robin
                                    // (three groups measured no faster: the fused launch is not issue-bound)
constexpr int NUM_MATH_WARPS = 4 * NUM_MATH_GROUPS;
constexpr int GEMM_THREADS = 128 + 32 * NUM_MATH_WARPS;
constexpr int EPI_ARRAY_BYTES = BLOCK_M * EPI_COLS * 4;  // 8 KB: one fp32 [128 x 16] sub-tile
constexpr int EPI_PART_BYTES = BLOCK_M * EPI_COLS * 2;   // 4 KB: one bf16 [128 x 16] sub-tile

// Compile-time configuration for P staged parts per operand and NIN fp32 epilogue inputs per sub-tile. Shared memory
// (227 KB) is split between the operand ring and the epilogue's input ring according to what bounds the launch:
// NIN = 1 (plain GEMMs, K large): deep operand ring; NIN = 2 (fused update after the short K = D contraction of the
// synthesis form, HBM-bound): shallow operand ring, 5-8 input stages in flight; NIN = 3 (Gram-form fused update).
// RES > 0: "resident B" -- the launch has a single column of tiles (N <= BN) and at most RES K blocks, so this CTA's half
// of B is the same for every tile: it is loaded once into its own region and the ring carries A only (the
// convolutional launches: B is the 16-64 KB dictionary operand, re-fetching it per tile was a third of the L2->SM bytes).
// NBANDS > 0 (with RES > 0): "halo" staging of a segmented A. The taps of one kernel row (qy fixed, qx = 0..tx-1) read
// the same operand rows shifted by qx: instead of one 128-row tile per tap, each stage holds NBANDS bands of HALO_ROWS =
// 128 + halo rows (one per kernel row qy) and every tap's UMMA descriptor starts a few rows into its band -- the
// swizzle is a function of the shared-memory address, so a row offset is just a byte offset. 16x16 kernels at stride 8:
// 2 bands of 144 rows instead of 4 tiles of 128 rows per K column block (0.56x the A bytes through L2).
constexpr int HALO_ROWS = 144;
constexpr int HALO_LEAD = 8;        // rows in front of the tile in a band that serves negative tap shifts
template <int P, int NIN, int BN, int RES = 0, int NBANDS = 0>
struct Cfg {
  // epilogue math groups (four warps each). (A third group for the 64-wide halo tiles, which have only four sub-tiles
  // each, measured neutral: 0.238 against 0.232 ms per convolutional iteration, it costs an operand stage.)
  static constexpr int GROUPS = NUM_MATH_GROUPS;
  // RES < 0 ("deep operand ring", NIN = 2 only): the fused update behind a LONG K = D contraction (D = 1024, configs[3]).
  // The default NIN = 2 split (2 operand stages, 6 input stages) is made for K = 256, where the launch is HBM-bound; with
  // 32 K blocks per tile two operand stages make every K block an L2 round trip -- ncu: tensor pipe 55 % active against
  // 96 % for the r = y Phi - x launch of the same flops (profiles/r02_config3_ncu_details.txt). Deep: 4 operand stages,
  // 4 input stages, 2 output stages. (A third math group measured neutral there: not epilogue-bound.)
  static constexpr bool DEEP = (RES < 0);
  static_assert(!DEEP || (NIN == 2 && NBANDS == 0), "the deep operand ring is a variant of the NIN = 2 split");
  static constexpr int MATH_WARPS = 4 * GROUPS;
  static constexpr int THREADS = 128 + 32 * MATH_WARPS;
  static_assert(NBANDS == 0 || RES > 0, "halo staging needs the resident B");
  static_assert(BN == 256 || BN == 128 || BN == 64, "tile width");
  static constexpr int HALF_N = BN / 2;                         // B rows staged by each CTA of the pair
  static constexpr int TMEM_COLS = 2 * BN;                      // two accumulators
  static constexpr int BK = (P == 1) ? 64 : 32;                 // K extent of a stage
  static constexpr int SPAN = BK * 2;                           // bytes per operand row = swizzle span (128 / 64)
  static constexpr int TILE_BYTES = BLOCK_M * SPAN;             // one part tile of A
  static constexpr int B_TILE_BYTES = HALF_N * SPAN;            // one part tile of this CTA's half of B
  static constexpr int BAND_BYTES = HALO_ROWS * SPAN;           // one part of one band
  static constexpr int STAGE_BYTES = NBANDS > 0 ? NBANDS * P * BAND_BYTES
                                   : RES > 0 ? P * TILE_BYTES : P * (TILE_BYTES + B_TILE_BYTES);
  static constexpr int BRES_KB_BYTES = P * B_TILE_BYTES;        // one K block of the resident B
  static constexpr int BRES_BYTES = (RES > 0 ? RES : 0) * BRES_KB_BYTES;
  static constexpr int IN_STAGE_BYTES = NIN * EPI_ARRAY_BYTES;
  static constexpr int OP_STAGES = NBANDS > 0 ? (NIN == 1 ? 3 : 2)
                                 : RES > 0 ? (NIN == 1 ? 6 : (P == 3 ? 2 : 3))
                                 : NIN == 3 ? ((P == 3) ? 2 : 3)
                                 : NIN == 2 ? (DEEP ? (P == 3 ? 2 : 4) : 2)
                                            : ((P == 3) ? 3 : 4);
  // IN_STAGES is a multiple of NUM_MATH_GROUPS: every input stage is always consumed by the same math group, so a
  // group sees the phases of "its" stages strictly in order (parity waits must never run a whole phase ahead).
  static constexpr int IN_STAGES = NBANDS > 0 ? (NIN == 1 ? 6 : 4)
                                 : NIN == 3 ? ((P == 3) ? 2 : 4)
                                 : NIN == 2 ? (DEEP ? 4 : (P == 3) ? 4 : 6)
                                            : ((P == 3) ? 2 : 6);
  static constexpr int OUT_STAGES = (NBANDS > 0 && NIN == 1) ? 2 : (NIN == 3 && P != 3) ? 2 : DEEP ? 2 : 3;
  static_assert(IN_STAGES % GROUPS == 0, "input stages must have a fixed owner group");
  // output stages: with two groups the in-order storer keeps a stage's barrier at most one phase behind any waiter
  // (three stages are fine); with three groups a slow group can leave it two phases behind, so every stage then needs a
  // fixed writer group as well
  static_assert(GROUPS == 2 || OUT_STAGES % GROUPS == 0, "output stages must have a fixed writer group");
  static constexpr int OUT_STAGE_BYTES = EPI_ARRAY_BYTES + P * EPI_PART_BYTES;
  static constexpr int NPAIRS = (P == 1) ? 1 : (P == 2) ? 3 : 6;
  static constexpr int OFF_OP = 0;
  static constexpr int OFF_BRES = OFF_OP + OP_STAGES * STAGE_BYTES;
  static constexpr int OFF_IN = OFF_BRES + BRES_BYTES;
  static constexpr int OFF_OUT = OFF_IN + IN_STAGES * IN_STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_OUT + OUT_STAGES * OUT_STAGE_BYTES;
  static constexpr int NUM_BARRIERS = 2 * OP_STAGES + 4 + 2 * IN_STAGES + 2 * OUT_STAGES + 1;  // + resident B full
  static constexpr int SMEM_TOTAL = OFF_BAR + NUM_BARRIERS * 8 + 16;
  static constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;          // slack to align the dynamic base to 1024 B
  static_assert(SMEM_ALLOC <= 232448, "over the 227 KB shared memory limit");
};
// (A part, B part) products of each mode, smallest magnitude first so fp32 accumulation loses the least:
//   P = 2: (hi,lo) (lo,hi) (hi,hi)        P = 3: (0,2) (2,0) (1,1) (0,1) (1,0) (0,0)
__host__ __device__ constexpr int pair_a(int P, int i) {
  return P == 1 ? 0 : P == 2 ? (i == 1 ? 1 : 0) : (i == 1 ? 2 : (i == 2 || i == 4) ? 1 : 0);
}
__host__ __device__ constexpr int pair_b(int P, int i) {
  return P == 1 ? 0 : P == 2 ? (i == 0 ? 1 : 0) : (i == 0 ? 2 : (i == 2 || i == 3) ? 1 : 0);
}

enum EpiKind { EPI_STORE = 0, EPI_FISTA = 1 };
enum ProxFlags { PROX_HARD = 1, PROX_NONNEG = 2 };
// Tuning switches (VTC_B200_FLAGS). All default to off; the measured effect of each on configs[1] is in
// profiles/README.md. (Bits 2 and 4 were TMA L2 prefetch cursors for operands / state tiles: measured 8-37 % slower,
// removed.)
enum TuneFlags {
  TUNE_CONTIGUOUS = 1,          // contiguous tile range per cluster instead of round-robin        (measured: slower)
  TUNE_STATE_EVICT_FIRST = 8,   // evict-first hint on state loads: neighbours fetched in the same 128-byte line are
                                // dropped before their own sub-tile asks for them (+45 % HBM reads) (measured: slower)
  TUNE_PROMO_256 = 16,          // 256-byte L2 promotion on the fp32 tensor maps                     (measured: slower)
  TUNE_NO_PDL = 32,             // launch without programmatic stream serialization
};

struct GemmParams {
  CUtensorMap tmA, tmB;    // bf16 operands, 2-D (cols, rows), box BK x 128, swizzle = row span
  CUtensorMap tmIn[3];     // fp32 epilogue inputs, 2-D, box 16 x 128, SWIZZLE_64B
  CUtensorMap tmOut;       // fp32 output, 2-D, box 16 x 128, SWIZZLE_64B
  CUtensorMap tmParts;     // bf16 parts output, 2-D over (n_parts * Kp) columns, box 16 x 128, SWIZZLE_32B
  int M, N;
  int num_m_blocks, num_n_blocks;   // in units of 256 x 256 pair tiles
  int k_blocks;            // ceil(K / BK)
  int a_part_stride, b_part_stride;   // column offset between parts of A / B (their padded K)
  int out_part_stride;     // column offset between parts of the bf16 output
  int ksplits, kb_per_split;
  int out_rows_per_split;  // fp32 partial outputs are stacked along rows
  int in_mask;             // bit i set: epilogue input slot i is loaded (FISTA: 0 = a_k, 1 = b, 2 = a_{k-1})
  int n_parts, store_out;
  int prox, group;         // ProxFlags; group size for subspace shrinkage (1 = scalar prox)
  int use_momentum;        // FISTA (1) or ISTA (0)
  float beta_prev, beta_next;
  const float* scalars;    // device: [0]=eta, [1]=theta
  double* stat;            // optional: += sum |a_new - a_k| (early stopping statistic)
  int flags;               // tuning switches, see TuneFlags
  // Blocked ("tile-contiguous") global layouts: instead of row-major, a matrix is stored as [column block][row][W] so
  // that a (128 rows x W columns) TMA box is ONE contiguous span of HBM. bit i (0..2): tmIn[i]; bit 3: tmOut;
  // bit 4: tmParts; bit 5: tmA. Such maps are 3-D (W, rows, blocks); fp32 state uses W = 16, operands W = BK.
  int blocked_mask;
  int a_blocks_per_part;      // blocked A: column blocks per part
  int parts_block_w;          // blocked parts output: block width (the consumer's BK) ...
  int parts_blocks_per_part;  // ... and column blocks per part
  // Segmented K (strided convolutions as GEMMs over image blocks, see vtc_fista_conv): K is nseg segments of seg_kb
  // K blocks; segment q reads the SAME A matrix (columns (kb % seg_kb) * BK) with its rows shifted by seg_shift[q],
  // against B columns kb * BK as usual. Rows shifted outside the matrix read as zero (TMA fill). seg_kb = 0: off.
  int seg_kb;
  int seg_shift[MAX_SEGMENTS];
  // halo staging (Cfg NBANDS > 0): band b of a stage holds rows [m0 + halo_band_row[b], + HALO_ROWS) of A (tmAh, box
  // HALO_ROWS rows); tap q reads band halo_tap_band[q] from row halo_tap_row[q] on
  CUtensorMap tmAh;
  int halo_band_row[4];
  int halo_tap_band[MAX_SEGMENTS], halo_tap_row[MAX_SEGMENTS];
  // Rows are (image, i, j) on a grid_h x grid_w grid of stride-sized image blocks (grid_w = 0: no grid).
  //   EPI_FISTA: code positions exist for i < code_h, j < code_w; the other rows are padding and stay exactly zero.
  //   EPI_STORE: columns are (channel, dy, dx) of a blk_sy x blk_sx block; outputs whose pixel (i*sy+dy, j*sx+dx) lies
  //              outside [pix_y0, pix_y1) x [pix_x0, pix_x1) are forced to zero (utils/convolutions.py:17-24 create_mask).
  int grid_h, grid_w, code_h, code_w;
  int blk_sy, blk_sx, pix_y0, pix_y1, pix_x0, pix_x1;
};
enum BlockedBits { BLK_IN0 = 1, BLK_OUT = 8, BLK_PARTS = 16, BLK_A = 32 };

struct TileCoord {
  int m0, n0, kb0, kb1, out_row0, nsub;
};
template <int BN>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int w, int cta_rank) {
  const int n_blk = w % p.num_n_blocks;
  const int t = w / p.num_n_blocks;
  const int m_blk = t % p.num_m_blocks;
  const int z = t / p.num_m_blocks;
  TileCoord c;
  c.m0 = m_blk * PAIR_M + cta_rank * BLOCK_M;
  c.n0 = n_blk * BN;
  c.kb0 = z * p.kb_per_split;
  c.kb1 = min(p.k_blocks, c.kb0 + p.kb_per_split);
  c.out_row0 = z * p.out_rows_per_split + c.m0;
  const int ncols = min(BN, p.N - c.n0);
  c.nsub = (ncols + EPI_COLS - 1) / EPI_COLS;
  return c;
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(lo)) | (static_cast<uint32_t>(__bfloat16_as_ushort(hi)) << 16);
}

// Subspace shrinkage of subspace_ista_fista.py:149-156 over G adjacent columns: a_g = u_g * max(1 - theta/||u_g||, 0),
// with ||u_g|| = 0 replaced by 1. Fully unrolled so the 16-wide register arrays never become local memory.
template <int G>
__device__ __forceinline__ void group_shrink(const float (&u)[16], float (&o)[16], float theta) {
#pragma unroll
  for (int g0 = 0; g0 < 16; g0 += G) {
    float ss = 0.f;
#pragma unroll
    for (int x = 0; x < G; ++x) ss = __fadd_rn(ss, __fmul_rn(u[g0 + x], u[g0 + x]));
    float nrm = sqrtf(ss);
    if (nrm == 0.f) nrm = 1.f;
    const float scale = fmaxf(__fsub_rn(1.f, __fdiv_rn(theta, nrm)), 0.f);
#pragma unroll
    for (int x = 0; x < G; ++x) o[g0 + x] = __fmul_rn(u[g0 + x], scale);
  }
}

// Scalar soft-threshold update of 16 columns with the iteration's structure fixed at compile time (the common case):
// no per-element selects on run-time flags. HASB: the drive b is an input (Gram form); PREV: a_{k-1} is an input
// (momentum term non-zero); MOM: FISTA extrapolation of the new iterate.
// The sums, differences and products run two columns at a time on the packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2 of
// sm_100): every packed operation rounds each half exactly like the scalar one it replaces (a - b is issued as
// fma(b, -1, a): the product is exact, one rounding), so the result is bit-identical to the unpacked sequence and to
// the reference's separate multiply and add; only the issue slots are halved.
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
// STAT: 1 / 0 = the convergence statistic is (not) accumulated, decided at compile time; -1 = want_stat decides.
template <bool HASB, bool PREV, bool MOM, int STAT = -1>
__device__ __forceinline__ void soft_update16(const uint32_t (&v)[16], const float (&in)[3][16], float eta, float theta,
                                              float beta_prev, float beta_next, float (&outv)[16],
                                              float (&partv)[16], float& stat_local, bool want_stat = false) {
  const float2 eta2 = make_float2(eta, eta), bp2 = make_float2(beta_prev, beta_prev);
  const float2 bn2 = make_float2(beta_next, beta_next);
#pragma unroll
  for (int x = 0; x < 16; x += 2) {
    const float2 ak = make_float2(in[0][x], in[0][x + 1]);
    float2 y = ak;
    if (PREV) y = __fadd2_rn(ak, __fmul2_rn(bp2, sub2(ak, make_float2(in[2][x], in[2][x + 1]))));
    float2 g = make_float2(__uint_as_float(v[x]), __uint_as_float(v[x + 1]));
    if (HASB) g = sub2(g, make_float2(in[1][x], in[1][x + 1]));
    const float2 u = sub2(y, __fmul2_rn(eta2, g));
    float2 a;
    a.x = copysignf(fmaxf(__fsub_rn(fabsf(u.x), theta), 0.f), u.x);
    a.y = copysignf(fmaxf(__fsub_rn(fabsf(u.y), theta), 0.f), u.y);
    outv[x] = a.x;
    outv[x + 1] = a.y;
    const float2 d = sub2(a, ak);
    const float2 yn = MOM ? __fadd2_rn(a, __fmul2_rn(bn2, d)) : a;
    partv[x] = yn.x;
    partv[x + 1] = yn.y;
    if (STAT > 0 || (STAT < 0 && want_stat)) stat_local += fabsf(d.x) + fabsf(d.y);
  }
}

// One ISTA/FISTA update of a 16-column sub-tile (shared by the GEMM epilogue below and the panel-resident iteration
// kernel in fista_iter_kernel.cuh). v = accumulator (gradient, or y G in the Gram form), in[0] = a_{k-1}, in[1] = b (Gram
// form only), in[2] = a_{k-2} (only when the momentum term is non-zero). outv = a_k, partv = y_k (next operand).
struct UpdateArgs {
  int in_mask, prox, group, use_momentum;
  float beta_prev, beta_next, eta, theta;
  bool want_stat;
};
__device__ __forceinline__ void fista_update16(const UpdateArgs& p, const uint32_t (&v)[16], const float (&in)[3][16],
                                               float (&outv)[16], float (&partv)[16], float& stat_local) {
  if (p.prox == 0 && p.group <= 1) {
    // fast paths: scalar soft threshold (ista_fista.py:117-120) with the iteration structure known at compile time
    const bool prev = (p.in_mask & 4) != 0;
#define VTC_SOFT(HASB, PREV, MOM) \
  soft_update16<HASB, PREV, MOM>(v, in, p.eta, p.theta, p.beta_prev, p.beta_next, outv, partv, stat_local, p.want_stat)
    if (p.in_mask & 2) {  // Gram form: b is an input
      if (!p.use_momentum) VTC_SOFT(true, false, false);
      else if (!prev) VTC_SOFT(true, false, true);
      else VTC_SOFT(true, true, true);
    } else {              // synthesis form: the accumulator already is the whole gradient
      if (!p.use_momentum) VTC_SOFT(false, false, false);
      else if (!prev) VTC_SOFT(false, false, true);
      else VTC_SOFT(false, true, true);
    }
#undef VTC_SOFT
    return;
  }
  // general path. in[1] absent -> 0: the accumulator already is the full gradient
  float u[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) {
    const float ak = in[0][x];
    float y = ak;
    if (p.in_mask & 4) y = __fadd_rn(ak, __fmul_rn(p.beta_prev, __fsub_rn(ak, in[2][x])));
    const float g = __fsub_rn(__uint_as_float(v[x]), in[1][x]);
    u[x] = __fsub_rn(y, __fmul_rn(p.eta, g));
  }
  if (p.group <= 1) {
#pragma unroll
    for (int x = 0; x < 16; ++x) {
      const float ux = u[x];
      float a;
      if (p.prox & PROX_HARD) {
        const float mag = (p.prox & PROX_NONNEG) ? ux : fabsf(ux);
        a = (mag < p.theta) ? 0.f : ux;
      } else {
        a = fmaxf(__fsub_rn(ux, p.theta), 0.f);
      }
      outv[x] = a;
    }
  } else {
    switch (p.group) {
      case 2: group_shrink<2>(u, outv, p.theta); break;
      case 4: group_shrink<4>(u, outv, p.theta); break;
      case 8: group_shrink<8>(u, outv, p.theta); break;
      default: group_shrink<16>(u, outv, p.theta); break;
    }
  }
#pragma unroll
  for (int x = 0; x < 16; ++x) {
    const float a = outv[x];
    const float d = __fsub_rn(a, in[0][x]);
    partv[x] = p.use_momentum ? __fadd_rn(a, __fmul_rn(p.beta_next, d)) : a;
    if (p.want_stat) stat_local += fabsf(d);
  }
}

// bf16 split of 16 fp32 values into n_parts parts (hi, then the residual's hi, ...): emit(part, w32) receives the 16
// bf16 of one part packed as 8 words (column 2x in the low half of word x).
template <typename Emit>
__device__ __forceinline__ void split_parts16(const float (&partv)[16], int n_parts, Emit emit) {
  float r[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) r[x] = partv[x];
#pragma unroll
  for (int part = 0; part < MAX_PARTS; ++part) {
    if (part >= n_parts) break;
    uint32_t w32[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      // one cvt.rn.bf16x2.f32 packs two columns; the bf16 -> fp32 widening is a shift / mask of that word
      const __nv_bfloat162 h = __floats2bfloat162_rn(r[2 * x], r[2 * x + 1]);
      w32[x] = *reinterpret_cast<const uint32_t*>(&h);
      if (part + 1 < n_parts) {
        r[2 * x] = __fsub_rn(r[2 * x], __uint_as_float(w32[x] << 16));
        r[2 * x + 1] = __fsub_rn(r[2 * x + 1], __uint_as_float(w32[x] & 0xffff0000u));
      }
    }
    emit(part, w32);
  }
}

template <int EPI, int P, int NIN, int BN, int RES = 0, int NBANDS = 0>
__global__ void __launch_bounds__((Cfg<P, NIN, BN, RES, NBANDS>::THREADS), 1)
vtc_gemm_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<P, NIN, BN, RES, NBANDS>;
  constexpr int IN_STAGE_BYTES = C::IN_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // swizzled TMA / UMMA tiles need a 1024-byte aligned base; the offset is identical in both CTAs of the pair
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sOp = sbase + C::OFF_OP, sIn = sbase + C::OFF_IN, sOut = sbase + C::OFF_OUT;
  const uint32_t bar0 = sbase + C::OFF_BAR;
  auto full_bar = [&](int s) { return bar0 + 8 * s; };
  auto empty_bar = [&](int s) { return bar0 + 8 * (C::OP_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar0 + 8 * (2 * C::OP_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar0 + 8 * (2 * C::OP_STAGES + 2 + a); };
  auto in_full_bar = [&](int e) { return bar0 + 8 * (2 * C::OP_STAGES + 4 + e); };
  auto in_free_bar = [&](int e) { return bar0 + 8 * (2 * C::OP_STAGES + 4 + C::IN_STAGES + e); };
  auto out_full_bar = [&](int o) { return bar0 + 8 * (2 * C::OP_STAGES + 4 + 2 * C::IN_STAGES + o); };
  auto out_free_bar = [&](int o) { return bar0 + 8 * (2 * C::OP_STAGES + 4 + 2 * C::IN_STAGES + C::OUT_STAGES + o); };
  const uint32_t bres_full_bar = bar0 + 8 * (C::NUM_BARRIERS - 1);
  const uint32_t sBres = sbase + C::OFF_BRES;
  const uint32_t tmem_slot = bar0 + C::NUM_BARRIERS * 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int total_tiles = p.num_m_blocks * p.num_n_blocks * p.ksplits;
  // Each cluster walks a contiguous range of tiles (n fastest): the 256-row A panel of an m block is fetched from
  // HBM once, for the first of its n tiles, and is L2-resident for the others.
  const bool contiguous = (p.flags & TUNE_CONTIGUOUS) != 0;
  const int tiles_per_cluster = (total_tiles + num_clusters - 1) / num_clusters;
  const int w_begin = contiguous ? min(total_tiles, cluster_id * tiles_per_cluster) : cluster_id;
  const int w_end = contiguous ? min(total_tiles, w_begin + tiles_per_cluster) : total_tiles;
  const int w_step = contiguous ? 1 : num_clusters;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < 3; ++i)
      if (p.in_mask & (1 << i)) tma_prefetch_desc(&p.tmIn[i]);
    if (p.store_out) tma_prefetch_desc(&p.tmOut);
    if (p.n_parts) tma_prefetch_desc(&p.tmParts);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::OP_STAGES; ++s) {
      mbar_init(full_bar(s), 2);   // one arrive per CTA's producer (used in the leader only)
      mbar_init(empty_bar(s), 1);  // multicast tcgen05.commit
    }
    mbar_init(bres_full_bar, 2);   // resident B: one arrive per CTA's producer (used in the leader only)
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);                     // multicast tcgen05.commit
      mbar_init(tmem_empty_bar(a), 2 * C::MATH_WARPS);   // every math warp of both CTAs (used in the leader only)
    }
    for (int e = 0; e < C::IN_STAGES; ++e) {
      mbar_init(in_full_bar(e), 1);
      mbar_init(in_free_bar(e), 4);   // the four warps of the math group that consumed the stage
    }
    for (int o = 0; o < C::OUT_STAGES; ++o) {
      mbar_init(out_full_bar(o), 4);
      mbar_init(out_free_bar(o), 1);
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the tail of the previous launch
  // in the stream; global memory written by it is only touched after this point.
  pdl_launch_dependents();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================================ operand producer (both CTAs) ================================
    const bool a_blocked = (p.blocked_mask & BLK_A) != 0;
    if (RES > 0 && w_begin < w_end) {
      // resident B: every K block of this CTA's half of the (single) B tile, once per launch
      if (elect_one_sync()) {
        if (leader) mbar_arrive_expect_tx(bres_full_bar, 2 * p.k_blocks * C::BRES_KB_BYTES);
        else mbar_arrive_remote(bres_full_bar, 0);
        for (int kb = 0; kb < p.k_blocks; ++kb)
#pragma unroll
          for (int q = 0; q < P; ++q)
            tma_load_2d_pair(sBres + kb * C::BRES_KB_BYTES + q * C::B_TILE_BYTES, &p.tmB, bres_full_bar,
                             q * p.b_part_stride + kb * C::BK, cta_rank * C::HALF_N, kEvictLast);
      }
      __syncwarp();
    }
    uint32_t it = 0;
    for (int w = w_begin; w < w_end; w += w_step) {
      const TileCoord c = decode_tile<BN>(p, w, cta_rank);
      if (NBANDS > 0) {
        // halo staging: one stage per K column block, NBANDS bands of HALO_ROWS rows, all parts
        for (int kba = 0; kba < p.seg_kb; ++kba, ++it) {
          const int s = it % C::OP_STAGES;
          const uint32_t ph = (it / C::OP_STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * C::STAGE_BYTES);
            else mbar_arrive_remote(full_bar(s), 0);
            const uint32_t dst = sOp + s * C::STAGE_BYTES;
#pragma unroll
            for (int b = 0; b < NBANDS; ++b)
#pragma unroll
              for (int q = 0; q < P; ++q)
                tma_load_3d_pair(dst + (b * P + q) * C::BAND_BYTES, &p.tmAh, full_bar(s), 0, c.m0 + p.halo_band_row[b],
                                 q * p.a_blocks_per_part + kba, kEvictNormal);
          }
          __syncwarp();
        }
      }
      for (int kb = c.kb0; NBANDS == 0 && kb < c.kb1; ++kb, ++it) {
        const int s = it % C::OP_STAGES;
        const uint32_t ph = (it / C::OP_STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        if (elect_one_sync()) {
          if (leader) mbar_arrive_expect_tx(full_bar(s), 2 * C::STAGE_BYTES);
          else mbar_arrive_remote(full_bar(s), 0);
          const uint32_t dst = sOp + s * C::STAGE_BYTES;
          // segmented K: this K block of A = column block (kb % seg_kb) of the rows shifted by the segment's offset
          const int kba = p.seg_kb ? kb % p.seg_kb : kb;
          const int arow = p.seg_kb ? c.m0 + p.seg_shift[kb / p.seg_kb] : c.m0;
#pragma unroll
          for (int q = 0; q < P; ++q) {
            if (a_blocked)
              tma_load_3d_pair(dst + q * C::TILE_BYTES, &p.tmA, full_bar(s), 0, arow, q * p.a_blocks_per_part + kba,
                               kEvictNormal);
            else
              tma_load_2d_pair(dst + q * C::TILE_BYTES, &p.tmA, full_bar(s), q * p.a_part_stride + kba * C::BK, arow,
                               kEvictNormal);
          }
          if (RES <= 0) {
#pragma unroll
            for (int q = 0; q < P; ++q)
              tma_load_2d_pair(dst + P * C::TILE_BYTES + q * C::B_TILE_BYTES, &p.tmB, full_bar(s),
                               q * p.b_part_stride + kb * C::BK, c.n0 + cta_rank * C::HALF_N, kEvictLast);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA) ================================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR_M, BN);
      uint32_t it = 0, tile_iter = 0;
      if (RES > 0 && w_begin < w_end) {
        mbar_wait(bres_full_bar, 0);
        tc_fence_after();
      }
      for (int w = w_begin; w < w_end; w += w_step, ++tile_iter) {
        const TileCoord c = decode_tile<BN>(p, w, cta_rank);
        const int acc = tile_iter & 1;
        const uint32_t acc_ph = (tile_iter >> 1) & 1;
        mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        if (NBANDS > 0) {
          const int nseg = p.k_blocks / p.seg_kb;
          for (int kba = 0; kba < p.seg_kb; ++kba, ++it) {
            const int s = it % C::OP_STAGES;
            mbar_wait(full_bar(s), (it / C::OP_STAGES) & 1);
            tc_fence_after();
            const uint32_t stage = sOp + s * C::STAGE_BYTES;
            if (elect_one_sync()) {
              for (int q = 0; q < nseg; ++q) {
                // tap q: rows halo_tap_row[q].. of its band; B K block (q, kba) from the resident operand
                const uint32_t abase = stage + p.halo_tap_band[q] * (P * C::BAND_BYTES) + p.halo_tap_row[q] * C::SPAN;
                const uint32_t bbase = sBres + (q * p.seg_kb + kba) * C::BRES_KB_BYTES;
#pragma unroll
                for (int pr = 0; pr < C::NPAIRS; ++pr) {
                  const uint64_t adesc = make_kmajor_desc(abase + pair_a(P, pr) * C::BAND_BYTES, C::SPAN);
                  const uint64_t bdesc = make_kmajor_desc(bbase + pair_b(P, pr) * C::B_TILE_BYTES, C::SPAN);
#pragma unroll
                  for (int k = 0; k < C::BK / UMMA_K; ++k) {
                    umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                    accumulate = 1;
                  }
                }
              }
              umma_commit_pair(empty_bar(s), 3);
            }
            accumulate = 1;
            __syncwarp();
          }
        }
        for (int kb = c.kb0; NBANDS == 0 && kb < c.kb1; ++kb, ++it) {
          const int s = it % C::OP_STAGES;
          const uint32_t ph = (it / C::OP_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t stage = sOp + s * C::STAGE_BYTES;
          if (elect_one_sync()) {
#pragma unroll
            for (int pr = 0; pr < C::NPAIRS; ++pr) {
              const uint64_t adesc = make_kmajor_desc(stage + pair_a(P, pr) * C::TILE_BYTES, C::SPAN);
              const uint32_t bbase = RES > 0 ? sBres + kb * C::BRES_KB_BYTES : stage + P * C::TILE_BYTES;
              const uint64_t bdesc = make_kmajor_desc(bbase + pair_b(P, pr) * C::B_TILE_BYTES, C::SPAN);
#pragma unroll
              for (int k = 0; k < C::BK / UMMA_K; ++k) {
                // +32 bytes (16 bf16) per K step inside the swizzle span -> +2 in the (address >> 4) field
                umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
                accumulate = 1;
              }
            }
            umma_commit_pair(empty_bar(s), 3);  // both CTAs' slots are free once these MMAs have drained
          }
          accumulate = 1;
          __syncwarp();
        }
        if (elect_one_sync()) umma_commit_pair(tmem_full_bar(acc), 3);  // accumulator ready for both epilogues
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ================================ epilogue loader ================================
    const uint32_t in_bytes = __popc(p.in_mask) * EPI_ARRAY_BYTES;
    uint32_t q = 0;
    for (int w = w_begin; w < w_end; w += w_step) {
      const TileCoord c = decode_tile<BN>(p, w, cta_rank);
      for (int j = 0; j < c.nsub; ++j, ++q) {
        const int e = q % C::IN_STAGES;
        const uint32_t ph = (q / C::IN_STAGES) & 1;
        mbar_wait(in_free_bar(e), ph ^ 1);
        if (elect_one_sync()) {
          if (p.in_mask != 0) {
            mbar_arrive_expect_tx(in_full_bar(e), in_bytes);
            const uint64_t hint = (p.flags & TUNE_STATE_EVICT_FIRST) ? kEvictFirst : kEvictNormal;
            for (int i = 0; i < 3; ++i) {
              if (!(p.in_mask & (1 << i))) continue;  // inputs are packed into the stage in slot order
              const uint32_t dst = sIn + e * IN_STAGE_BYTES + __popc(p.in_mask & ((1 << i) - 1)) * EPI_ARRAY_BYTES;
              const int col = c.n0 + j * EPI_COLS;
              if (p.blocked_mask & (BLK_IN0 << i)) tma_load_3d(dst, &p.tmIn[i], in_full_bar(e), 0, c.m0, col / EPI_COLS, hint);
              else tma_load_2d(dst, &p.tmIn[i], in_full_bar(e), col, c.m0, hint);
            }
          } else {
            mbar_arrive(in_full_bar(e));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ================================ epilogue storer ================================
    uint32_t q = 0;
    for (int w = w_begin; w < w_end; w += w_step) {
      const TileCoord c = decode_tile<BN>(p, w, cta_rank);
      for (int j = 0; j < c.nsub; ++j, ++q) {
        const int o = q % C::OUT_STAGES;
        const uint32_t ph = (q / C::OUT_STAGES) & 1;
        mbar_wait(out_full_bar(o), ph);  // a math group has written {out | parts} of sub-tile q
        const uint32_t src = sOut + o * C::OUT_STAGE_BYTES;
        if (elect_one_sync()) {  // bulk groups are per thread: the same elected lane issues, commits and waits
          const int col = c.n0 + j * EPI_COLS;
          if (p.store_out) {
            if (p.blocked_mask & BLK_OUT) tma_store_3d(&p.tmOut, src, 0, c.out_row0, col / EPI_COLS);
            else tma_store_2d(&p.tmOut, src, col, c.out_row0);
          }
          for (int part = 0; part < p.n_parts; ++part) {
            const uint32_t psrc = src + EPI_ARRAY_BYTES + part * EPI_PART_BYTES;
            if (p.blocked_mask & BLK_PARTS)
              tma_store_3d(&p.tmParts, psrc, col % p.parts_block_w, c.m0,
                           part * p.parts_blocks_per_part + col / p.parts_block_w);
            else
              tma_store_2d(&p.tmParts, psrc, part * p.out_part_stride + col, c.m0);
          }
          bulk_commit();
          if (q >= C::OUT_STAGES - 1) {
            // all but the OUT_STAGES-1 most recent store groups have left shared memory: recycle the oldest stage
            bulk_wait_read<C::OUT_STAGES - 1>();
            mbar_arrive(out_free_bar((q - (C::OUT_STAGES - 1)) % C::OUT_STAGES));
          }
        }
        __syncwarp();
      }
    }
    if (elect_one_sync()) bulk_wait<0>();
    __syncwarp();
  } else if (warp >= 4) {
    // ================================ epilogue math ================================
    const uint32_t group = (warp - 4) >> 2;  // which sub-tiles (q mod NUM_MATH_GROUPS) this warp works on
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;     // row inside this CTA's 128-row tile
    const uint32_t sw64 = (row >> 1) & 3;    // SWIZZLE_64B: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)
    const uint32_t sw32 = (row >> 2) & 1;    // SWIZZLE_32B: chunk c of row r lives at chunk c ^ ((r >> 2) & 1)
    float eta = 0.f, theta = 0.f;
    if (EPI == EPI_FISTA) {
      eta = __ldg(p.scalars + 0);
      theta = __ldg(p.scalars + 1);
    }
    float stat_local = 0.f;
    uint32_t q = 0, tile_iter = 0;
    for (int w = w_begin; w < w_end; w += w_step, ++tile_iter) {
      const TileCoord c = decode_tile<BN>(p, w, cta_rank);
      const int acc = tile_iter & 1;
      const uint32_t acc_ph = (tile_iter >> 1) & 1;
      // grid position of this thread's row (convolutional launches), once per tile: the integer divisions are not cheap.
      // mask_mode: 0 = the row's whole pixel block is inside the un-masked region (or padding row handling is off),
      // 1 = wholly outside (every output zero), 2 = straddles the border (per-pixel test)
      int gi = 0, gj = 0, mask_mode = 0;
      bool padding_row = false;
      if (p.grid_w > 0) {
        const int cell = (c.m0 + row) % (p.grid_h * p.grid_w);
        gi = cell / p.grid_w;
        gj = cell - gi * p.grid_w;
        if (EPI == EPI_FISTA) {
          padding_row = gi >= p.code_h || gj >= p.code_w;
        } else {
          const int py0 = gi * p.blk_sy, px0 = gj * p.blk_sx;
          const bool inside = py0 >= p.pix_y0 && py0 + p.blk_sy <= p.pix_y1 && px0 >= p.pix_x0 && px0 + p.blk_sx <= p.pix_x1;
          const bool outside = py0 + p.blk_sy <= p.pix_y0 || py0 >= p.pix_y1 || px0 + p.blk_sx <= p.pix_x0 || px0 >= p.pix_x1;
          mask_mode = inside ? 0 : outside ? 1 : 2;
        }
      }
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      // last sub-tile of this tile that belongs to this warp's group (-1: none)
      int j_last = -1;
      for (int j = c.nsub - 1; j >= 0 && j >= c.nsub - C::GROUPS; --j)
        if ((q + j) % C::GROUPS == group) {
          j_last = j;
          break;
        }
      if (j_last < 0) {
        __syncwarp();
        if (lane == 0) {
          tc_fence_before();
          mbar_arrive_remote(tmem_empty_bar(acc), 0);
        }
      }
      for (int j = 0; j < c.nsub; ++j, ++q) {
        if (q % C::GROUPS != group) continue;
        const int e = q % C::IN_STAGES;
        const uint32_t in_ph = (q / C::IN_STAGES) & 1;
        const int o = q % C::OUT_STAGES;
        const uint32_t out_ph = (q / C::OUT_STAGES) & 1;
        uint32_t v[16];
        tmem_ld16(t_row + j * EPI_COLS, v);
        mbar_wait(in_full_bar(e), in_ph);
        tmem_ld_wait();
        if (j == j_last) {  // this warp has drained its share of the accumulator
          __syncwarp();
          if (lane == 0) {
            tc_fence_before();
            mbar_arrive_remote(tmem_empty_bar(acc), 0);
          }
        }
        const uint32_t in_stage = sIn + e * IN_STAGE_BYTES;
        float in[3][16];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (p.in_mask & (1 << i)) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float4 t = lds128(in_stage + __popc(p.in_mask & ((1 << i) - 1)) * EPI_ARRAY_BYTES + row * 64 +
                                      ((ch ^ sw64) << 4));
              in[i][4 * ch + 0] = t.x; in[i][4 * ch + 1] = t.y; in[i][4 * ch + 2] = t.z; in[i][4 * ch + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x) in[i][x] = 0.f;
          }
        }
        float outv[16];   // fp32 result
        float partv[16];  // value whose bf16 split is emitted as parts
        if (EPI == EPI_FISTA && padding_row) {
          // padding row of the code grid: its inputs are zero and stay zero (prox(0) = 0 for every variant)
#pragma unroll
          for (int x = 0; x < 16; ++x) v[x] = 0u;
        }
        if (EPI == EPI_STORE) {
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            outv[x] = __uint_as_float(v[x]) - in[0][x];
            partv[x] = outv[x];
          }
          if (mask_mode == 1) {
#pragma unroll
            for (int x = 0; x < 16; ++x) outv[x] = 0.f, partv[x] = 0.f;
          } else if (mask_mode == 2) {
            // reconstruction mask: zero outside the un-padded image; column -> (dy, dx) walks incrementally
            const int col0 = c.n0 + j * EPI_COLS;
            int dx = col0 % p.blk_sx;
            int dy = (col0 / p.blk_sx) % p.blk_sy;
            const int py0 = gi * p.blk_sy, px0 = gj * p.blk_sx;
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const int py = py0 + dy, px = px0 + dx;
              if (py < p.pix_y0 || py >= p.pix_y1 || px < p.pix_x0 || px >= p.pix_x1) outv[x] = 0.f, partv[x] = 0.f;
              if (++dx == p.blk_sx) {
                dx = 0;
                if (++dy == p.blk_sy) dy = 0;
              }
            }
          }
        } else {
          UpdateArgs ua;
          ua.in_mask = p.in_mask, ua.prox = p.prox, ua.group = p.group, ua.use_momentum = p.use_momentum;
          ua.beta_prev = p.beta_prev, ua.beta_next = p.beta_next, ua.eta = eta, ua.theta = theta;
          ua.want_stat = p.stat != nullptr;
          fista_update16(ua, v, in, outv, partv, stat_local);
        }
        // every lane has consumed its inputs: hand the in stage back to the loader
        __syncwarp();
        if (lane == 0) mbar_arrive(in_free_bar(e));
        mbar_wait(out_free_bar(o), out_ph ^ 1);
        const uint32_t out_stage = sOut + o * C::OUT_STAGE_BYTES;
        if (p.store_out) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            sts128(out_stage + row * 64 + ((ch ^ sw64) << 4), outv[4 * ch], outv[4 * ch + 1], outv[4 * ch + 2],
                   outv[4 * ch + 3]);
        }
        if (p.n_parts) {
          split_parts16(partv, p.n_parts, [&](int part, const uint32_t (&w32)[8]) {
            const uint32_t prow = out_stage + EPI_ARRAY_BYTES + part * EPI_PART_BYTES + row * 32;
            sts128u(prow + ((0 ^ sw32) << 4), w32[0], w32[1], w32[2], w32[3]);
            sts128u(prow + ((1 ^ sw32) << 4), w32[4], w32[5], w32[6], w32[7]);
          });
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_full_bar(o));
      }
    }
    if (EPI == EPI_FISTA && p.stat) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) stat_local += __shfl_xor_sync(0xffffffffu, stat_local, o);
      if (lane == 0) atomicAdd(p.stat, static_cast<double>(stat_local));
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
}

}  // namespace vtc
