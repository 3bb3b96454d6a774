// Persistent, warp-specialised tcgen05 GEMM with a fused, TMA-staged epilogue (sm_100a).
//
//   D[M,N] = sum_seg  A[:, a_koff[seg] : +K] * B[:, b_koff[seg] : +K]^T        (bf16 operands, fp32 accumulate in TMEM)
//
// Both operands are K-major bf16 matrices whose rows hold several "parts" of an fp32 matrix side by side
// (hi | mid | lo of a bf16 split, each padded to a multiple of 64 columns). A list of (A part, B part) segment
// pairs selects the precision: 1 segment = plain bf16, 3 = bf16x3 (~2^-17), 6 = bf16x6 (~fp32).
//
// Warp roles (256 threads, one CTA per SM, static round-robin tile schedule):
//   warp 0 lane 0 : TMA producer for the A/B operand ring            (full/empty mbarriers)
//   warp 1 lane 0 : tcgen05.mma issuer, accumulators double-buffered in TMEM (tmem_full/tmem_empty)
//   warp 2        : TMEM allocator / deallocator
//   warp 3 lane 0 : epilogue DMA: TMA-loads the fp32 state tiles a sub-tile ahead, TMA-stores the results
//   warps 4..7    : epilogue math: tcgen05.ld 16 columns -> fused update -> swizzled st.shared
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

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;     // 64 bf16 = one 128-byte swizzle span
constexpr int UMMA_K = 16;
constexpr int NUM_STAGES = 3;   // operand ring
constexpr int EPI_COLS = 16;    // epilogue sub-tile width (fp32 columns)
constexpr int EPI_STAGES = 3;   // epilogue state ring
constexpr int MAX_SEG = 6;
constexpr int MAX_PARTS = 3;

constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;                 // 16 KB
constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;                 // 32 KB
constexpr int EPI_ARRAY_BYTES = BLOCK_M * EPI_COLS * 4;              // 8 KB : one fp32 [128 x 16] sub-tile
constexpr int EPI_STAGE_BYTES = 3 * EPI_ARRAY_BYTES;                 // 24 KB: {in0,in1,in2} then {out | parts}
constexpr int SMEM_A_OFF = 0;
constexpr int SMEM_B_OFF = SMEM_A_OFF + NUM_STAGES * A_STAGE_BYTES;
constexpr int SMEM_EPI_OFF = SMEM_B_OFF + NUM_STAGES * B_STAGE_BYTES;
constexpr int SMEM_BAR_OFF = SMEM_EPI_OFF + EPI_STAGES * EPI_STAGE_BYTES;
constexpr int NUM_BARRIERS = 2 * NUM_STAGES + 4 + 2 * EPI_STAGES;
constexpr int SMEM_TOTAL = SMEM_BAR_OFF + NUM_BARRIERS * 8 + 16;
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;  // slack to align the dynamic base to 1024 B
constexpr int GEMM_THREADS = 256;
constexpr int TMEM_COLS = 2 * BLOCK_N;         // two accumulators

enum EpiKind { EPI_STORE = 0, EPI_FISTA = 1 };
enum ProxFlags { PROX_HARD = 1, PROX_NONNEG = 2 };

struct GemmParams {
  CUtensorMap tmA, tmB;    // bf16 operands, 2-D (cols, rows), box 64 x 128 / 64 x 256, SWIZZLE_128B
  CUtensorMap tmIn[3];     // fp32 epilogue inputs, 2-D, box 16 x 128, SWIZZLE_64B
  CUtensorMap tmOut;       // fp32 output, 2-D, box 16 x 128, SWIZZLE_64B
  CUtensorMap tmParts;     // bf16 parts output, 3-D (cols, parts, rows), box 16 x n_parts x 128, no swizzle
  int M, N;
  int num_m_blocks, num_n_blocks;
  int k_blocks;            // ceil(K / 64) per segment
  int nseg;
  int a_koff[MAX_SEG], b_koff[MAX_SEG];
  int ksplits, kb_per_split;
  int out_rows_per_split;  // fp32 partial outputs are stacked along rows
  int n_in, n_parts, store_out;
  int prox, group;         // ProxFlags; group size for subspace shrinkage (1 = scalar prox)
  int use_momentum;        // FISTA (1) or ISTA (0)
  float beta_prev, beta_next;
  const float* scalars;    // device: [0]=eta, [1]=theta
  double* stat;            // optional: += sum |a_new - a_k| (early stopping statistic)
};

struct TileCoord {
  int m0, n0, kb0, kb1, out_row0, nsub;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int w) {
  const int n_blk = w % p.num_n_blocks;
  const int t = w / p.num_n_blocks;
  const int m_blk = t % p.num_m_blocks;
  const int z = t / p.num_m_blocks;
  TileCoord c;
  c.m0 = m_blk * BLOCK_M;
  c.n0 = n_blk * BLOCK_N;
  c.kb0 = z * p.kb_per_split;
  c.kb1 = min(p.k_blocks, c.kb0 + p.kb_per_split);
  c.out_row0 = z * p.out_rows_per_split + c.m0;
  const int ncols = min(BLOCK_N, p.N - c.n0);
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

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) vtc_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need a 1024-byte aligned base.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sA = sbase + SMEM_A_OFF, sB = sbase + SMEM_B_OFF, sE = sbase + SMEM_EPI_OFF;
  const uint32_t bar0 = sbase + SMEM_BAR_OFF;
  auto full_bar = [&](int s) { return bar0 + 8 * s; };
  auto empty_bar = [&](int s) { return bar0 + 8 * (NUM_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar0 + 8 * (2 * NUM_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar0 + 8 * (2 * NUM_STAGES + 2 + a); };
  auto epi_full_bar = [&](int e) { return bar0 + 8 * (2 * NUM_STAGES + 4 + e); };
  auto epi_done_bar = [&](int e) { return bar0 + 8 * (2 * NUM_STAGES + 4 + EPI_STAGES + e); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR_OFF + NUM_BARRIERS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.num_m_blocks * p.num_n_blocks * p.ksplits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < p.n_in; ++i) tma_prefetch_desc(&p.tmIn[i]);
    if (p.store_out) tma_prefetch_desc(&p.tmOut);
    if (p.n_parts) tma_prefetch_desc(&p.tmParts);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), 128);
    }
    for (int e = 0; e < EPI_STAGES; ++e) {
      mbar_init(epi_full_bar(e), 1);
      mbar_init(epi_done_bar(e), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0 && lane == 0) {
    // ================================ operand producer ================================
    uint32_t it = 0;
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x) {
      const TileCoord c = decode_tile(p, w);
      for (int seg = 0; seg < p.nseg; ++seg) {
        for (int kb = c.kb0; kb < c.kb1; ++kb, ++it) {
          const int s = it % NUM_STAGES;
          const uint32_t ph = (it / NUM_STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), A_STAGE_BYTES + B_STAGE_BYTES);
          tma_load_2d(sA + s * A_STAGE_BYTES, &p.tmA, full_bar(s), p.a_koff[seg] + kb * BLOCK_K, c.m0, kEvictNormal);
          tma_load_2d(sB + s * B_STAGE_BYTES, &p.tmB, full_bar(s), p.b_koff[seg] + kb * BLOCK_K, c.n0, kEvictLast);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
    uint32_t it = 0, tile_iter = 0;
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x, ++tile_iter) {
      const TileCoord c = decode_tile(p, w);
      const int acc = tile_iter & 1;
      const uint32_t acc_ph = (tile_iter >> 1) & 1;
      mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      uint32_t accumulate = 0;
      for (int seg = 0; seg < p.nseg; ++seg) {
        for (int kb = c.kb0; kb < c.kb1; ++kb, ++it) {
          const int s = it % NUM_STAGES;
          const uint32_t ph = (it / NUM_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t adesc = make_kmajor_sw128_desc(sA + s * A_STAGE_BYTES);
          const uint64_t bdesc = make_kmajor_sw128_desc(sB + s * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // +32 bytes (UMMA_K bf16) per step inside the 128-byte swizzle span -> +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(empty_bar(s));  // smem slot is free once these MMAs have drained
        }
      }
      umma_commit(tmem_full_bar(acc));  // accumulator ready for the epilogue
    }
  } else if (warp == 3 && lane == 0) {
    // ================================ epilogue DMA ================================
    const uint32_t in_bytes = p.n_in * EPI_ARRAY_BYTES;
    // load cursor
    int wl = blockIdx.x, jl = 0;
    TileCoord cl = decode_tile(p, wl < total_tiles ? wl : 0);
    uint32_t ql = 0;
    auto issue_load = [&]() {  // arm stage (ql % EPI_STAGES) with sub-tile (wl, jl); returns false past the end
      if (wl >= total_tiles) return false;
      const int e = ql % EPI_STAGES;
      const uint32_t dst = sE + e * EPI_STAGE_BYTES;
      if (p.n_in > 0) {
        mbar_arrive_expect_tx(epi_full_bar(e), in_bytes);
        for (int i = 0; i < p.n_in; ++i)
          tma_load_2d(dst + i * EPI_ARRAY_BYTES, &p.tmIn[i], epi_full_bar(e), cl.n0 + jl * EPI_COLS, cl.m0,
                      kEvictFirst);
      } else {
        mbar_arrive(epi_full_bar(e));  // nothing to load: just hand the stage to the math warps
      }
      ++ql;
      if (++jl == cl.nsub) {
        jl = 0;
        wl += gridDim.x;
        if (wl < total_tiles) cl = decode_tile(p, wl);
      }
      return true;
    };
    for (int i = 0; i < EPI_STAGES; ++i) issue_load();
    uint32_t q = 0;
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x) {
      const TileCoord c = decode_tile(p, w);
      for (int j = 0; j < c.nsub; ++j, ++q) {
        const int e = q % EPI_STAGES;
        const uint32_t ph = (q / EPI_STAGES) & 1;
        mbar_wait(epi_done_bar(e), ph);  // math warps have written {out | parts} of sub-tile q
        const uint32_t src = sE + e * EPI_STAGE_BYTES;
        if (p.store_out) tma_store_2d(&p.tmOut, src, c.n0 + j * EPI_COLS, c.out_row0);
        if (p.n_parts) tma_store_3d(&p.tmParts, src + EPI_ARRAY_BYTES, c.n0 + j * EPI_COLS, 0, c.m0);
        bulk_commit();
        bulk_wait_read<0>();  // stage memory may be overwritten again
        issue_load();         // refill this stage with sub-tile q + EPI_STAGES
      }
    }
    bulk_wait<0>();
  } else if (warp >= 4) {
    // ================================ epilogue math ================================
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;     // row inside the 128-row tile
    const uint32_t sw = (row >> 1) & 3;      // SWIZZLE_64B: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)
    float eta = 0.f, theta = 0.f;
    if (EPI == EPI_FISTA) {
      eta = __ldg(p.scalars + 0);
      theta = __ldg(p.scalars + 1);
    }
    float stat_local = 0.f;
    uint32_t q = 0, tile_iter = 0;
    for (int w = blockIdx.x; w < total_tiles; w += gridDim.x, ++tile_iter) {
      const TileCoord c = decode_tile(p, w);
      const int acc = tile_iter & 1;
      const uint32_t acc_ph = (tile_iter >> 1) & 1;
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N;
      for (int j = 0; j < c.nsub; ++j, ++q) {
        const int e = q % EPI_STAGES;
        const uint32_t ph = (q / EPI_STAGES) & 1;
        uint32_t v[16];
        tmem_ld16(t_row + j * EPI_COLS, v);
        mbar_wait(epi_full_bar(e), ph);
        tmem_ld_wait();
        if (j == c.nsub - 1) {  // accumulator fully drained into registers: hand TMEM back to the MMA warp
          tc_fence_before();
          mbar_arrive(tmem_empty_bar(acc));
        }
        uint8_t* stage = smem + SMEM_EPI_OFF + e * EPI_STAGE_BYTES;
        float in[3][16];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          if (i < p.n_in) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const float4 t = *reinterpret_cast<const float4*>(stage + i * EPI_ARRAY_BYTES + row * 64 + ((ch ^ sw) << 4));
              in[i][4 * ch + 0] = t.x; in[i][4 * ch + 1] = t.y; in[i][4 * ch + 2] = t.z; in[i][4 * ch + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int x = 0; x < 16; ++x) in[i][x] = 0.f;
          }
        }
        float outv[16];   // fp32 result
        float partv[16];  // value whose bf16 split is emitted as parts
        if (EPI == EPI_STORE) {
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            outv[x] = __uint_as_float(v[x]) - in[0][x];
            partv[x] = outv[x];
          }
        } else {
          // in[0] = a_k, in[1] = b, in[2] = a_{k-1} (only loaded when the momentum term is non-zero)
          float u[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float ak = in[0][x];
            float y = ak;
            if (p.n_in == 3) y = __fadd_rn(ak, __fmul_rn(p.beta_prev, __fsub_rn(ak, in[2][x])));
            const float g = __fsub_rn(__uint_as_float(v[x]), in[1][x]);
            u[x] = __fsub_rn(y, __fmul_rn(eta, g));
          }
          if (p.group <= 1) {
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const float ux = u[x];
              float a;
              if (p.prox & PROX_HARD) {
                const float mag = (p.prox & PROX_NONNEG) ? ux : fabsf(ux);
                a = (mag < theta) ? 0.f : ux;
              } else if (p.prox & PROX_NONNEG) {
                a = fmaxf(__fsub_rn(ux, theta), 0.f);
              } else {
                a = copysignf(fmaxf(__fsub_rn(fabsf(ux), theta), 0.f), ux);
              }
              outv[x] = a;
            }
          } else {
            // subspace shrinkage over `group` adjacent columns (group divides 16)
            switch (p.group) {
              case 2: group_shrink<2>(u, outv, theta); break;
              case 4: group_shrink<4>(u, outv, theta); break;
              case 8: group_shrink<8>(u, outv, theta); break;
              default: group_shrink<16>(u, outv, theta); break;
            }
          }
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const float a = outv[x];
            const float d = __fsub_rn(a, in[0][x]);
            partv[x] = p.use_momentum ? __fadd_rn(a, __fmul_rn(p.beta_next, d)) : a;
            if (p.stat) stat_local += fabsf(d);
          }
        }
        named_bar_sync(1, 128);  // every math thread has finished reading this stage's inputs
        if (p.store_out) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            *reinterpret_cast<float4*>(stage + row * 64 + ((ch ^ sw) << 4)) =
                make_float4(outv[4 * ch], outv[4 * ch + 1], outv[4 * ch + 2], outv[4 * ch + 3]);
        }
        if (p.n_parts) {
          // smem box layout [row][part][16 bf16]
          uint8_t* prow = stage + EPI_ARRAY_BYTES + row * (p.n_parts * 32);
          float r[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) r[x] = partv[x];
#pragma unroll
          for (int part = 0; part < MAX_PARTS; ++part) {
            if (part >= p.n_parts) break;
            uint32_t w32[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const __nv_bfloat16 h0 = __float2bfloat16_rn(r[2 * x]);
              const __nv_bfloat16 h1 = __float2bfloat16_rn(r[2 * x + 1]);
              r[2 * x] = __fsub_rn(r[2 * x], __bfloat162float(h0));
              r[2 * x + 1] = __fsub_rn(r[2 * x + 1], __bfloat162float(h1));
              w32[x] = pack_bf16x2(h0, h1);
            }
            *reinterpret_cast<uint4*>(prow + part * 32) = make_uint4(w32[0], w32[1], w32[2], w32[3]);
            *reinterpret_cast<uint4*>(prow + part * 32 + 16) = make_uint4(w32[4], w32[5], w32[6], w32[7]);
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(epi_done_bar(e));
      }
    }
    if (EPI == EPI_FISTA && p.stat) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) stat_local += __shfl_xor_sync(0xffffffffu, stat_local, o);
      if (lane == 0) atomicAdd(p.stat, static_cast<double>(stat_local));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace vtc
