// Small HBM-bound helper kernels around the tcgen05 GEMM: bf16 operand splitting, transposes, the Lipschitz constant
// (largest eigenvalue of dictionary^T dictionary) in fp64, the dictionary apply step, gathers/scatters for subspaces.
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vtc {

// ------------------------------------------------------------------------------------------------------------
// fp32 (R x C, pitch ld) -> bf16 parts (R x nparts*Cp): part p holds bf16(residual after removing parts < p);
// columns [C, Cp) of every part are written as zero so that K-tail tiles contribute nothing.
// block == 0: row-major, part p in columns [p*Cp, (p+1)*Cp). block > 0 (even, divides Cp): tile-contiguous layout
// [part][Cp/block][R][block], so that a (128 rows x block columns) operand tile is one contiguous span.
__global__ void split_rows_kernel(const float* __restrict__ in, int64_t ld, int64_t R, int64_t C, int64_t Cp,
                                  int nparts, int block, __nv_bfloat16* __restrict__ out, float scale = 1.f) {
  const int64_t half = Cp / 2;
  const int64_t total = R * half;
  const int64_t pitch = static_cast<int64_t>(nparts) * Cp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / half;
    const int64_t c = (i - r * half) * 2;
    float v0 = (c < C) ? scale * in[r * ld + c] : 0.f;       // scale = +-1: exact
    float v1 = (c + 1 < C) ? scale * in[r * ld + c + 1] : 0.f;
    for (int p = 0; p < nparts; ++p) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      v0 = __fsub_rn(v0, __bfloat162float(h0));
      v1 = __fsub_rn(v1, __bfloat162float(h1));
      __nv_bfloat162 pk;
      pk.x = h0;
      pk.y = h1;
      const int64_t idx = block ? ((p * (Cp / block) + c / block) * R + r) * block + c % block : r * pitch + p * Cp + c;
      *reinterpret_cast<__nv_bfloat162*>(out + idx) = pk;
    }
  }
}

// tile-contiguous fp32 [ceil(C/16)][R][16]  ->  row-major (R x C, pitch ld)
__global__ void unblock_f32_kernel(const float* __restrict__ in, int64_t R, int64_t C, float* __restrict__ out,
                                   int64_t ld) {
  const int64_t total = R * C;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / C, c = i - r * C;
    out[r * ld + c] = in[((c / 16) * R + r) * 16 + (c % 16)];
  }
}

// "Quad-blocked" fp32 state of the second-generation iteration kernel (fista_iter2_kernel.cuh): sub-tile (col block cb
// of 16 atoms, row block rb of 128 rows) is one contiguous 8 KB block [4 quads][128 rows][4 floats] at
// ((cb * RB + rb) * 2048) floats. Rows at or beyond R and columns at or beyond C are zero.
__global__ void block_quad_kernel(const float* __restrict__ in, int64_t ld, int64_t R, int64_t C, int64_t RB, int64_t CB,
                                  float* __restrict__ out) {
  const int64_t total = CB * RB * 512;   // float4 slots
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t blk = i / 512;
    const int w = static_cast<int>(i - blk * 512);
    const int quad = w / 128, row = w % 128;
    const int64_t cb = blk / RB, rb = blk - cb * RB;
    const int64_t r = rb * 128 + row, c = cb * 16 + quad * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R) {
      if (c + 0 < C) v.x = in[r * ld + c + 0];
      if (c + 1 < C) v.y = in[r * ld + c + 1];
      if (c + 2 < C) v.z = in[r * ld + c + 2];
      if (c + 3 < C) v.w = in[r * ld + c + 3];
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}
// The way back: one float4 slot per thread. Inside a warp four lanes take the four quads of one row and eight rows
// follow each other, so the reads are four 128-byte runs of the block and the writes eight 64-byte runs of the rows
// (whole sectors on both sides; the slot-per-float version this replaces moved 16 bytes per 32-byte sector and took
// 0.42 ms for the 268 MB of configs[1]); blocks are walked column block fastest, so that the two halves of a row's
// 128-byte line are written back to back.
__global__ void unblock_quad_kernel(const float* __restrict__ in, int64_t R, int64_t C, int64_t RB,
                                    float* __restrict__ out, int64_t ld) {
  const int64_t CB = (C + 15) / 16;
  const int64_t total = CB * RB * 512;   // float4 slots
  const bool vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ld & 3) == 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t blk = i / 512;
    const int w = static_cast<int>(i - blk * 512);
    const int quad = w & 3, row = (w >> 5) * 8 + ((w & 31) >> 2);
    const int64_t rb = blk / CB, cb = blk - rb * CB;
    const int64_t r = rb * 128 + row, c = cb * 16 + quad * 4;
    if (r >= R || c >= C) continue;
    const float4 v = reinterpret_cast<const float4*>(in)[(cb * RB + rb) * 512 + quad * 128 + row];
    float* dst = out + r * ld + c;
    if (vec && c + 3 < C) {
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      dst[0] = v.x;
      if (c + 1 < C) dst[1] = v.y;
      if (c + 2 < C) dst[2] = v.z;
      if (c + 3 < C) dst[3] = v.w;
    }
  }
}

// fp32 (R x C, pitch ld) -> bf16 parts of the TRANSPOSE: out is (C x nparts*Rp), zero in columns [R, Rp).
// 32x32 tiles through shared memory so that both the read and the write are coalesced.
__global__ void transpose_split_kernel(const float* __restrict__ in, int64_t ld, int64_t R, int64_t C, int64_t Rp,
                                       int nparts, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;  // over Rp
  const int64_t c0 = static_cast<int64_t>(blockIdx.y) * 32;  // over C
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[r * ld + c] : 0.f;
  }
  __syncthreads();
  const int64_t pitch = static_cast<int64_t>(nparts) * Rp;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < Rp) {
      float v = tile[threadIdx.x][i];
      for (int p = 0; p < nparts; ++p) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        v = __fsub_rn(v, __bfloat162float(h));
        out[c * pitch + p * Rp + r] = h;
      }
    }
  }
}

// fp32 (R x C, pitch ld) -> fp32 transpose (C x R, pitch ldo)
__global__ void transpose_f32_kernel(const float* __restrict__ in, int64_t ld, int64_t R, int64_t C,
                                     float* __restrict__ out, int64_t ldo) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int64_t c0 = static_cast<int64_t>(blockIdx.y) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) out[c * ldo + r] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Lipschitz constant L = lambda_max(Phi^T Phi) (ista_fista.py:72-74), in fp64, without a host round trip:
//   M = Phi^T Phi (n x n, n = D padded to 32);  A_0 = M;  A_{j+1} = (A_j / tr A_j)^2
// After p squarings A_p / tr A_p is the projector onto the top eigenspace up to (lambda_2/lambda_1)^(2^p), and
//   L = <A_p, M> / tr A_p   is the eigenvalue average under that weight (a Rayleigh quotient).
__global__ void gram_fp64_kernel(const float* __restrict__ phi, int64_t S, int64_t D, int n, double* __restrict__ M,
                                 double* __restrict__ trace_out) {
  __shared__ double As[32][33], Bs[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  double acc[4] = {0, 0, 0, 0};
  for (int64_t s0 = 0; s0 < S; s0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const int64_t s = s0 + r;
      As[r][tx] = (s < S && i0 + tx < D) ? static_cast<double>(phi[s * D + i0 + tx]) : 0.0;
      Bs[r][tx] = (s < S && j0 + tx < D) ? static_cast<double>(phi[s * D + j0 + tx]) : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const double b = Bs[k][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += As[k][ty + 8 * q] * b;
    }
    __syncthreads();
  }
  double tr = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + ty + 8 * q, j = j0 + tx;
    M[static_cast<int64_t>(i) * n + j] = acc[q];
    if (i == j) tr += acc[q];
  }
  if (i0 == j0 && tr != 0.0) atomicAdd(trace_out, tr);
}

// C = (A / *trace_in)^2 for symmetric A (n x n, n % 32 == 0); accumulates tr C into *trace_out.
// Squaring number `step` >= 1 reads *trace_in = tr (A_{step-1} / tr A_{step-1})^2 = sum_i w_i^2 of the normalised
// eigenvalue weights w_i: once that is within kLipschitzConverged of 1 a single eigen-direction carries all the weight
// (the Rayleigh quotient is then exact to ~(1 - sum w_i^2) / 2 relative), the step index is recorded in *stop and this
// and every later squaring returns at once. A (nearly) degenerate top eigenvalue never triggers it and runs them all.
constexpr double kLipschitzConverged = 1e-11;
__global__ void square_fp64_kernel(const double* __restrict__ A, int n, const double* __restrict__ trace_in,
                                   double* __restrict__ C, double* __restrict__ trace_out, int step,
                                   double* __restrict__ stop) {
  __shared__ double As[32][33], Bs[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (*stop != 0.0) return;
  if (step >= 1 && 1.0 - *trace_in < kLipschitzConverged) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && tx == 0 && ty == 0) *stop = static_cast<double>(step);
    return;
  }
  const double scale = 1.0 / *trace_in;
  double acc[4] = {0, 0, 0, 0};
  for (int k0 = 0; k0 < n; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      // A symmetric: A[i][k] read as row k0+r of column block i0 (coalesced)
      As[r][tx] = A[static_cast<int64_t>(k0 + r) * n + i0 + tx] * scale;
      Bs[r][tx] = A[static_cast<int64_t>(k0 + r) * n + j0 + tx] * scale;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const double b = Bs[k][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += As[k][ty + 8 * q] * b;
    }
    __syncthreads();
  }
  double tr = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + ty + 8 * q, j = j0 + tx;
    C[static_cast<int64_t>(i) * n + j] = acc[q];
    if (i == j) tr += acc[q];
  }
  if (i0 == j0 && tr != 0.0) atomicAdd(trace_out, tr);
}

// The Gram matrix and ALL squarings in one cooperative launch (grid-wide barriers between the steps) instead of one
// launch per step: the thirty dependent launches were a third of a small call (BASELINE configs[0]). Same tiles, same
// arithmetic and the same stopping rule as gram_fp64_kernel / square_fp64_kernel above; the Rayleigh quotient stays in
// lipschitz_finalize_kernel (one block, fixed summation order). Launched with cudaLaunchCooperativeKernel: the whole
// grid is resident or the launch fails (the caller then takes the per-step launches).
__global__ void lipschitz_squarings_kernel(const float* __restrict__ phi, int64_t S, int64_t D, int n,
                                           double* __restrict__ M, double* __restrict__ A0, double* __restrict__ A1,
                                           double* __restrict__ traces, int squarings) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double As[32][33], Bs[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  {
    double acc[4] = {0, 0, 0, 0};
    for (int64_t s0 = 0; s0 < S; s0 += 32) {
      for (int r = ty; r < 32; r += 8) {
        const int64_t s = s0 + r;
        As[r][tx] = (s < S && i0 + tx < D) ? static_cast<double>(phi[s * D + i0 + tx]) : 0.0;
        Bs[r][tx] = (s < S && j0 + tx < D) ? static_cast<double>(phi[s * D + j0 + tx]) : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const double b = Bs[k][tx];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += As[k][ty + 8 * q] * b;
      }
      __syncthreads();
    }
    double tr = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + ty + 8 * q, j = j0 + tx;
      M[static_cast<int64_t>(i) * n + j] = acc[q];
      if (i == j) tr += acc[q];
    }
    if (i0 == j0 && tr != 0.0) atomicAdd(traces + 0, tr);
  }
  grid.sync();
  const double* src = M;
  double* dst = A0;
  for (int step = 0; step < squarings; ++step) {
    const double trace_in = traces[step];
    if (step >= 1 && 1.0 - trace_in < kLipschitzConverged) {   // the same value in every block: a uniform exit
      if (blockIdx.x == 0 && blockIdx.y == 0 && tx == 0 && ty == 0) traces[squarings + 1] = static_cast<double>(step);
      break;
    }
    const double scale = 1.0 / trace_in;
    double acc[4] = {0, 0, 0, 0};
    for (int k0 = 0; k0 < n; k0 += 32) {
      for (int r = ty; r < 32; r += 8) {
        As[r][tx] = src[static_cast<int64_t>(k0 + r) * n + i0 + tx] * scale;
        Bs[r][tx] = src[static_cast<int64_t>(k0 + r) * n + j0 + tx] * scale;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const double b = Bs[k][tx];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += As[k][ty + 8 * q] * b;
      }
      __syncthreads();
    }
    double tr = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + ty + 8 * q, j = j0 + tx;
      dst[static_cast<int64_t>(i) * n + j] = acc[q];
      if (i == j) tr += acc[q];
    }
    if (i0 == j0 && tr != 0.0) atomicAdd(traces + step + 1, tr);
    grid.sync();
    src = dst;
    dst = (dst == A0) ? A1 : A0;
  }
}

// scalars[0] = eta = 1/L, [1] = theta = lambda * eta, [2] = L, [3] = status bits (1 = non-finite / non-positive L)
// traces[j] = trace of iterate j (A_0 = M; iterate j >= 1 lives in A0 for odd j, A1 for even j); traces[squarings + 1] =
// the iterate the squarings stopped at (0: all of them ran).
__global__ void lipschitz_finalize_kernel(const double* __restrict__ A0, const double* __restrict__ A1,
                                          const double* __restrict__ M, int n, const double* __restrict__ traces,
                                          int squarings, float sparsity_weight, float* __restrict__ scalars,
                                          float* __restrict__ lipschitz_out) {
  __shared__ double red[1024];
  const double stop = traces[squarings + 1];
  const int last = (stop != 0.0) ? static_cast<int>(stop) : squarings;
  const double* __restrict__ Ap = (last & 1) ? A0 : A1;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int64_t total = static_cast<int64_t>(n) * n;
  const int64_t bd = blockDim.x;
  for (int64_t i = threadIdx.x; i < total; i += 4 * bd) {   // four independent chains of loads per thread
    s0 += Ap[i] * M[i];
    if (i + bd < total) s1 += Ap[i + bd] * M[i + bd];
    if (i + 2 * bd < total) s2 += Ap[i + 2 * bd] * M[i + 2 * bd];
    if (i + 3 * bd < total) s3 += Ap[i + 3 * bd] * M[i + 3 * bd];
  }
  red[threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float L = static_cast<float>(red[0] / traces[last]);
    if (lipschitz_out) *lipschitz_out = L;
    if (scalars) {
      const float eta = 1.f / L;  // stepsize = 1. / lipschitz_constant  (ista_fista.py:80), float32 arithmetic
      scalars[0] = eta;
      scalars[1] = sparsity_weight * eta;  // sparsity_weight * stepsize  (ista_fista.py:119)
      scalars[2] = L;
      scalars[3] = (isfinite(L) && L > 0.f) ? 0.f : 1.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// grad[s][d] = sum_z partial[z][s][d]   (fixed order -> deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nsplit, int64_t rows_per_split,
                                       int64_t ldp, int64_t S, int64_t D, float* __restrict__ grad) {
  const int64_t total = S * D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t s = i / D, d = i - s * D;
    float acc = 0.f;
    for (int z = 0; z < nsplit; ++z) acc += partial[(z * rows_per_split + s) * ldp + d];
    grad[i] = acc;
  }
}

// Gradient of the within-group alignment penalty sum_{i != j in group} |cos(phi_i, phi_j)|
// (subspace_sc_cheap_quadratic_descent.py:91-127). One block per group; slots[g*W + i] is the atom of member i or -1.
//   normalised dictionary:  grad_i = sum_j sign(c_ij) (phi_j - c_ij phi_i),            c = phi phi^T
//   general:                grad_i = sum_j sign(c_ij) (phi_j / (n_i n_j) - c_ij / n_i^2 phi_i),  c_ij = phi_i.phi_j / (n_i n_j)
// One block per ATOM: it walks the groups in order and, for every group the atom belongs to, adds that group's
// contribution to the atom's row -- the order of the reference's loop (:66-70: accum[group] = accum[group] + ...), no
// atomics, so the result is bit-reproducible whatever the group structure (an atom in three or more groups made the
// earlier atomicAdd version order-dependent, which broke the bit-identity of data-parallel replicas).
__global__ void alignment_grad_kernel(const float* __restrict__ dict, int64_t D, const int32_t* __restrict__ slots,
                                      int num_groups, int W, int normalized, float* __restrict__ accum) {
  extern __shared__ float sm[];
  float* rows = sm;                                // [W][D]: the rows of the current group
  float* dots = sm + static_cast<size_t>(W) * D;   // [W]: <phi_m, phi_j> for this atom m
  float* norms = dots + W;                         // [W]
  __shared__ int member_pos;
  const int32_t atom = static_cast<int32_t>(blockIdx.x);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int g = 0; g < num_groups; ++g) {
    const int32_t* members = slots + static_cast<int64_t>(g) * W;
    if (threadIdx.x == 0) {
      int pos = -1;
      for (int j = 0; j < W; ++j)
        if (members[j] == atom) {
          pos = j;
          break;
        }
      member_pos = pos;
    }
    __syncthreads();
    const int m = member_pos;
    if (m < 0) {
      __syncthreads();   // (member_pos is rewritten by the next group)
      continue;
    }
    for (int64_t i = threadIdx.x; i < static_cast<int64_t>(W) * D; i += blockDim.x) {
      const int j = static_cast<int>(i / D);
      const int32_t a = members[j];
      rows[i] = (a >= 0) ? dict[static_cast<int64_t>(a) * D + (i - static_cast<int64_t>(j) * D)] : 0.f;
    }
    __syncthreads();
    for (int j = warp; j < W; j += nwarps) {
      float acc = 0.f, nn = 0.f;
      for (int64_t d = lane; d < D; d += 32) {
        acc += rows[m * D + d] * rows[j * D + d];
        nn += rows[j * D + d] * rows[j * D + d];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
      }
      if (lane == 0) dots[j] = acc, norms[j] = sqrtf(nn);
    }
    __syncthreads();
    for (int64_t d = threadIdx.x; d < D; d += blockDim.x) {
      float grad = 0.f;
      for (int j = 0; j < W; ++j) {
        if (members[j] < 0) continue;
        float c, t;
        if (normalized) {
          c = dots[j];
          t = rows[j * D + d] - c * rows[m * D + d];
        } else {
          const float nn = norms[m] * norms[j];
          c = dots[j] / nn;
          t = rows[j * D + d] / nn - (c / (norms[m] * norms[m])) * rows[m * D + d];
        }
        const float sgn = (c > 0.f) ? 1.f : (c < 0.f) ? -1.f : 0.f;
        grad += sgn * t;
      }
      accum[static_cast<int64_t>(atom) * D + d] += grad;   // this block owns the row
    }
    __syncthreads();
  }
}

// One block per atom (row): sc_cheap_quadratic_descent.py:43-48 / sc_steepest_descent.py:37-41 after the contraction.
__global__ void dict_apply_kernel(float* __restrict__ dict, const float* __restrict__ grad, const float* __restrict__ h,
                                  const float* __restrict__ reg, float penalty, int64_t D, float batch, float stepsize,
                                  float lowest, int normalize) {
  __shared__ float red[256];
  const int64_t s = blockIdx.x;
  float* row = dict + s * D;
  const float* g = grad + s * D;
  const float denom = h ? (h[s] + lowest) : 1.f;
  float ss = 0.f;
  for (int64_t d = threadIdx.x; d < D; d += blockDim.x) {
    float data_term = __fdiv_rn(g[d], batch);               // mm / codes.size(0)
    if (reg) data_term = __fadd_rn(data_term, __fmul_rn(penalty, reg[s * D + d]));  // + alignment_penalty * reg. grads
    float u = __fmul_rn(stepsize, data_term);
    if (h) u = __fdiv_rn(u, denom);                                    // dict_update.div_(h[:, None] + lowest)
    const float v = __fsub_rn(row[d], u);                              // dictionary.sub_(dict_update)
    row[d] = v;
    ss += v * v;
  }
  if (!normalize) return;
  red[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float nrm = sqrtf(red[0]);
  for (int64_t d = threadIdx.x; d < D; d += blockDim.x) row[d] = __fdiv_rn(row[d], nrm);
}

// code_sq_sum[s] += sum over this block's rows of codes[r][s]^2
__global__ void col_sq_sum_kernel(const float* __restrict__ codes, int64_t ld, int64_t B, int64_t S,
                                  int64_t rows_per_block, float* __restrict__ out) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(B, r0 + rows_per_block);
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float v = codes[r * ld + s];
    acc += v * v;
  }
  atomicAdd(out + s, acc);
}
// h <- 0.99 h + (sum / batch) / 100      (training/sparse_coding.py:154)
__global__ void hessian_ema_kernel(float* __restrict__ h, const float* __restrict__ sq_sum, int64_t S, float batch) {
  const int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (s < S) h[s] = __fadd_rn(__fmul_rn(h[s], 0.99f), __fdiv_rn(__fdiv_rn(sq_sum[s], batch), 100.f));
}

// ------------------------------------------------------------------------------------------------------------
// Subspace regrouping (subspace_ista_fista.py:94-111, :184-190); index[j] = atom of slot j, or -1 for padding.
__global__ void gather_rows_kernel(const float* __restrict__ src, int64_t ld, const int32_t* __restrict__ index,
                                   int64_t n_slots, int64_t D, float* __restrict__ dst) {
  const int64_t total = n_slots * D;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t j = i / D, d = i - j * D;
    const int32_t a = index[j];
    dst[i] = (a >= 0) ? src[static_cast<int64_t>(a) * ld + d] : 0.f;
  }
}
__global__ void gather_cols_kernel(const float* __restrict__ src, int64_t ld, const int32_t* __restrict__ index,
                                   int64_t B, int64_t n_slots, float* __restrict__ dst, int64_t ldd) {
  const int64_t total = B * n_slots;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / n_slots, j = i - r * n_slots;
    const int32_t a = index[j];
    dst[r * ldd + j] = (a >= 0) ? src[r * ld + a] : 0.f;
  }
}
// dst (B x S) must be zeroed by the caller; duplicates of an atom in several groups are summed.
__global__ void scatter_add_cols_kernel(const float* __restrict__ src, int64_t ld, const int32_t* __restrict__ index,
                                        int64_t B, int64_t n_slots, float* __restrict__ dst, int64_t ldd) {
  const int64_t total = B * n_slots;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / n_slots, j = i - r * n_slots;
    const int32_t a = index[j];
    if (a >= 0) atomicAdd(dst + r * ldd + a, src[r * ld + j]);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Convolutional path (vtc_fista_conv): a strided convolution whose kernel is (ty x tx) strides large is a GEMM over
// "image blocks". Rows of every internal matrix are (image, i, j) on a grid of gh x gw stride-sized blocks; columns are
// either the S code channels or the db = c * sy * sx pixels (channel, dy, dx) of one block.
struct ConvGeom {
  int b, c, h, w;        // padded images
  int s, kh, kw;         // kernels
  int sy, sx, ty, tx;    // stride; kernel extent in strides (kh / sy, kw / sx)
  int gh, gw;            // grid of blocks: h / sy, w / sx
  int ch, cw;            // code grid: gh - ty + 1, gw - tx + 1
  int db;                // pixels per block: c * sy * sx
};

// images (b, c, h, w) -> blocks (b*gh*gw x db, pitch ld)
__global__ void conv_image_to_blocks_kernel(const float* __restrict__ img, ConvGeom g, float* __restrict__ out,
                                            int64_t ld) {
  const int64_t rows = static_cast<int64_t>(g.b) * g.gh * g.gw;
  const int64_t total = rows * g.db;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / g.db;
    const int pix = static_cast<int>(i - row * g.db);
    const int dx = pix % g.sx, dy = (pix / g.sx) % g.sy, ch = pix / (g.sx * g.sy);
    const int gj = static_cast<int>(row % g.gw), gi = static_cast<int>((row / g.gw) % g.gh);
    const int64_t bi = row / (static_cast<int64_t>(g.gw) * g.gh);
    out[row * ld + pix] = img[((bi * g.c + ch) * g.h + gi * g.sy + dy) * g.w + gj * g.sx + dx];
  }
}

// codes (b, s, ch, cw) -> grid layout (b*gh*gw x s, pitch ld); rows outside the code grid are zero
__global__ void conv_codes_to_grid_kernel(const float* __restrict__ codes, ConvGeom g, float* __restrict__ out,
                                          int64_t ld) {
  const int64_t rows = static_cast<int64_t>(g.b) * g.gh * g.gw;
  const int64_t total = rows * g.s;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // consecutive threads walk the source's fastest axis (j) so that the read is coalesced
    const int gj = static_cast<int>(i % g.gw);
    const int gi = static_cast<int>((i / g.gw) % g.gh);
    const int sc = static_cast<int>((i / (static_cast<int64_t>(g.gw) * g.gh)) % g.s);
    const int64_t bi = i / (static_cast<int64_t>(g.gw) * g.gh * g.s);
    const int64_t row = (bi * g.gh + gi) * g.gw + gj;
    float v = 0.f;
    if (gi < g.ch && gj < g.cw) v = codes[((bi * g.s + sc) * g.ch + gi) * g.cw + gj];
    out[row * ld + sc] = v;
  }
}

// grid layout (row-major, pitch ld) -> codes (b, s, ch, cw)
__global__ void conv_grid_to_codes_kernel(const float* __restrict__ grid, int64_t ld, ConvGeom g,
                                          float* __restrict__ codes) {
  const int64_t total = static_cast<int64_t>(g.b) * g.s * g.ch * g.cw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cj = static_cast<int>(i % g.cw);
    const int ci = static_cast<int>((i / g.cw) % g.ch);
    const int sc = static_cast<int>((i / (static_cast<int64_t>(g.cw) * g.ch)) % g.s);
    const int64_t bi = i / (static_cast<int64_t>(g.cw) * g.ch * g.s);
    codes[i] = grid[((bi * g.gh + ci) * g.gw + cj) * ld + sc];
  }
}

// dictionary (s, c, kh, kw) -> the two B operands, each as bf16 parts with the taps q = (qy, qx) side by side along K:
//   analysis  (s  x nparts * nq*dbp):  [sc][part][q * dbp + pix] = Phi[sc, ch, qy*sy+dy, qx*sx+dx]
//   synthesis (db x nparts * nq*sp ):  [pix][part][q * sp + sc]  = the same value
// (pix = (ch*sy + dy)*sx + dx). Both buffers are zeroed by the caller (padding columns stay zero).
__global__ void conv_dict_operands_kernel(const float* __restrict__ dict, ConvGeom g, int nparts, int64_t dbp, int64_t sp,
                                          __nv_bfloat16* __restrict__ analysis, __nv_bfloat16* __restrict__ synthesis) {
  const int64_t per_kernel = static_cast<int64_t>(g.c) * g.kh * g.kw;
  const int64_t total = g.s * per_kernel;
  const int nq = g.ty * g.tx;
  const int64_t ka = nq * dbp, ks = nq * sp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t sc = i / per_kernel;
    const int e = static_cast<int>(i - sc * per_kernel);
    const int v_ = e % g.kw, u_ = (e / g.kw) % g.kh, ch = e / (g.kw * g.kh);
    const int q = (u_ / g.sy) * g.tx + v_ / g.sx;
    const int pix = (ch * g.sy + u_ % g.sy) * g.sx + v_ % g.sx;
    float v = dict[i];
    for (int p = 0; p < nparts; ++p) {
      const __nv_bfloat16 hpart = __float2bfloat16_rn(v);
      v = __fsub_rn(v, __bfloat162float(hpart));
      analysis[sc * (nparts * ka) + p * ka + q * dbp + pix] = hpart;
      synthesis[static_cast<int64_t>(pix) * (nparts * ks) + p * ks + q * sp + sc] = hpart;
    }
  }
}

// per-tap gradient blocks (nq matrices of s x ld, tap q = rows of codes^T against blocks shifted by tap q)
//   -> gradient in dictionary layout (s, c, kh, kw)
__global__ void conv_grad_to_dict_layout_kernel(const float* __restrict__ taps, int64_t tap_stride, int64_t ld,
                                                ConvGeom g, float* __restrict__ grad) {
  const int64_t per_kernel = static_cast<int64_t>(g.c) * g.kh * g.kw;
  const int64_t total = g.s * per_kernel;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t sc = i / per_kernel;
    const int e = static_cast<int>(i - sc * per_kernel);
    const int v_ = e % g.kw, u_ = (e / g.kw) % g.kh, ch = e / (g.kw * g.kh);
    const int q = (u_ / g.sy) * g.tx + v_ / g.sx;
    const int pix = (ch * g.sy + u_ % g.sy) * g.sx + v_ % g.sx;
    grad[i] = taps[q * tap_stride + sc * ld + pix];
  }
}

// sum over images and positions of codes^2 per channel: codes (b, s, n) -> out (s,)   (training/sparse_coding.py:158-161)
__global__ void conv_channel_sq_sum_kernel(const float* __restrict__ codes, int64_t b, int64_t s, int64_t n,
                                           float* __restrict__ out) {
  __shared__ double red[256];
  const int64_t sc = blockIdx.x;
  double acc = 0.0;
  for (int64_t bi = 0; bi < b; ++bi) {
    const float* src = codes + (bi * s + sc) * n;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += static_cast<double>(src[i]) * src[i];
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[sc] = static_cast<float>(red[0]);
}

// Convolutional dictionary apply step, in place (dict_update_rules/convolutional/sc_cheap_quadratic_descent.py:72-79,
// sc_steepest_descent.py:66-72):  U = grad_sum / batch;  U /= (h + lowest) if h;  U *= ||Phi|| / ||U||  (Frobenius norms
// over the whole dictionary);  Phi -= stepsize * U;  every kernel divided by its own norm if normalize.
// One block (the dictionary is a few thousand to a few hundred thousand floats); reductions in double, fixed order.
__global__ void conv_dict_apply_kernel(float* __restrict__ dict, const float* __restrict__ grad_sum,
                                       const float* __restrict__ hessian, int64_t s, int64_t per_kernel, float batch,
                                       float stepsize, float lowest, int normalize) {
  __shared__ double red[1024];
  __shared__ double total_phi, total_u;
  const int64_t total = s * per_kernel;
  auto block_sum = [&](double v) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  auto update_of = [&](int64_t i) {
    float u = grad_sum[i] / batch;
    if (hessian != nullptr) u = u / (hessian[i / per_kernel] + lowest);
    return u;
  };
  double a = 0.0, b = 0.0;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const float u = update_of(i);
    a += static_cast<double>(dict[i]) * dict[i];
    b += static_cast<double>(u) * u;
  }
  const double sp = block_sum(a), su = block_sum(b);
  if (threadIdx.x == 0) total_phi = sp, total_u = su;
  __syncthreads();
  const float scale = static_cast<float>(sqrt(total_phi)) / static_cast<float>(sqrt(total_u));
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x)
    dict[i] = __fsub_rn(dict[i], __fmul_rn(stepsize, __fmul_rn(update_of(i), scale)));
  __syncthreads();
  if (normalize) {
    // one warp per kernel at a time
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int64_t sc = warp; sc < s; sc += nwarps) {
      float* row = dict + sc * per_kernel;
      double acc = 0.0;
      for (int64_t i = lane; i < per_kernel; i += 32) acc += static_cast<double>(row[i]) * row[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const float nrm = static_cast<float>(sqrt(acc));
      for (int64_t i = lane; i < per_kernel; i += 32) row[i] = row[i] / nrm;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Data feed: crops of (ph x pw) pixels out of device-resident images (n, h, w, c), flattened in (y, x, c)
// order like utils/dataset_generation.py:212-218 (all_patches[p] = img[v:v+ph, u:u+pw]) followed by its reshape(N, -1).
// corners (B, 3) int32: image index, top row, left column. One warp-sized group of threads per patch row segment;
// consecutive threads read consecutive (x, c) elements of an image row: coalesced on both sides.
__global__ void extract_patches_kernel(const float* __restrict__ images, int64_t h, int64_t w, int64_t c,
                                       const int32_t* __restrict__ corners, int64_t B, int64_t ph, int64_t pw,
                                       float* __restrict__ patches, int64_t ld) {
  const int64_t row_elems = pw * c;
  const int64_t per_patch = ph * row_elems;
  const int64_t total = B * per_patch;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / per_patch;
    const int64_t e = i - b * per_patch;
    const int64_t y = e / row_elems, xc = e - y * row_elems;
    const int64_t img = corners[3 * b], top = corners[3 * b + 1], left = corners[3 * b + 2];
    patches[b * ld + e] = images[((img * h + top + y) * w + left) * c + xc];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Validation metrics (training/sparse_coding.py:177-229) from device-resident residuals and codes:
// one record of METRIC_FIELDS floats per item (patch or image):
//   [0] sum of squared residuals   [1] l1 norm of the codes, or the sum of the group l2 norms (subspace)
//   [2] number of non-zero codes   [3] min and [4] max of the item's (cropped) pixels
// Every sum has a fixed order (lane-strided partials, then a shuffle/shared-memory tree), so results are
// reproducible run to run.
constexpr int METRIC_FIELDS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// fully connected: one warp per patch. resid (B x D, pitch ld_r) = codes * dictionary - images.
// slots != nullptr: (num_groups x group_width) atom indices, -1 = padding; field [1] becomes sum_g ||a_g||_2.
__global__ void fc_metrics_rows_kernel(const float* __restrict__ resid, int64_t ld_r, const float* __restrict__ images,
                                       int64_t ld_x, const float* __restrict__ codes, int64_t ld_a, int64_t B,
                                       int64_t S, int64_t D, const int32_t* __restrict__ slots, int64_t num_groups,
                                       int group_width, float* __restrict__ items) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t b = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    float r2 = 0.f, lo = INFINITY, hi = -INFINITY;
    for (int64_t d = lane; d < D; d += 32) {
      const float r = resid[b * ld_r + d], x = images[b * ld_x + d];
      r2 = fmaf(r, r, r2);
      lo = fminf(lo, x), hi = fmaxf(hi, x);
    }
    float l1 = 0.f, l0 = 0.f;
    for (int64_t s = lane; s < S; s += 32) {
      const float a = codes[b * ld_a + s];
      l1 += fabsf(a);
      l0 += (a != 0.f) ? 1.f : 0.f;
    }
    if (slots) {
      l1 = 0.f;
      for (int64_t g = lane; g < num_groups; g += 32) {
        float n2 = 0.f;
        for (int j = 0; j < group_width; ++j) {
          const int32_t s = slots[g * group_width + j];
          if (s >= 0) {
            const float a = codes[b * ld_a + s];
            n2 = fmaf(a, a, n2);
          }
        }
        l1 += sqrtf(n2);
      }
    }
    r2 = warp_sum(r2), l1 = warp_sum(l1), l0 = warp_sum(l0), lo = warp_min(lo), hi = warp_max(hi);
    if (lane == 0) {
      float* o = items + b * METRIC_FIELDS;
      o[0] = r2, o[1] = l1, o[2] = l0, o[3] = lo, o[4] = hi;
    }
  }
}

// convolutional, stage 1: grid (chunks, images); block j of image b takes the elements i = j*threads + t (+ k*chunks*
// threads) of the image's masked residual blocks (rows x db, pitch ld_r), of its codes (contiguous n_codes) and of the
// un-padded crop of its pixels, and writes one partial record.
__global__ void conv_metrics_partial_kernel(const float* __restrict__ resid, int64_t ld_r, const float* __restrict__ img,
                                            const float* __restrict__ codes, ConvGeom g, int pad_t, int pad_b,
                                            int pad_l, int pad_r, float* __restrict__ partials) {
  __shared__ float red[5][32];
  const int64_t b = blockIdx.y;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t rows = static_cast<int64_t>(g.gh) * g.gw;
  float r2 = 0.f, l1 = 0.f, l0 = 0.f, lo = INFINITY, hi = -INFINITY;
  for (int64_t i = first; i < rows * g.db; i += stride) {
    const int64_t r = i / g.db, c = i - r * g.db;
    const float v = resid[(b * rows + r) * ld_r + c];
    r2 = fmaf(v, v, r2);
  }
  const int64_t n_codes = static_cast<int64_t>(g.s) * g.ch * g.cw;
  for (int64_t i = first; i < n_codes; i += stride) {
    const float a = codes[b * n_codes + i];
    l1 += fabsf(a);
    l0 += (a != 0.f) ? 1.f : 0.f;
  }
  const int64_t hc = g.h - pad_t - pad_b, wc = g.w - pad_l - pad_r;
  for (int64_t i = first; i < static_cast<int64_t>(g.c) * hc * wc; i += stride) {
    const int64_t ch = i / (hc * wc), e = i - ch * hc * wc;
    const int64_t y = e / wc, x = e - y * wc;
    const float v = img[((b * g.c + ch) * g.h + pad_t + y) * g.w + pad_l + x];
    lo = fminf(lo, v), hi = fmaxf(hi, v);
  }
  r2 = warp_sum(r2), l1 = warp_sum(l1), l0 = warp_sum(l0), lo = warp_min(lo), hi = warp_max(hi);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[0][warp] = r2, red[1][warp] = l1, red[2][warp] = l0, red[3][warp] = lo, red[4][warp] = hi;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    r2 = warp_sum(lane < nw ? red[0][lane] : 0.f);
    l1 = warp_sum(lane < nw ? red[1][lane] : 0.f);
    l0 = warp_sum(lane < nw ? red[2][lane] : 0.f);
    lo = warp_min(lane < nw ? red[3][lane] : INFINITY);
    hi = warp_max(lane < nw ? red[4][lane] : -INFINITY);
    if (lane == 0) {
      float* o = partials + (b * gridDim.x + blockIdx.x) * METRIC_FIELDS;
      o[0] = r2, o[1] = l1, o[2] = l0, o[3] = lo, o[4] = hi;
    }
  }
}
// stage 2: one warp per image folds its chunk partials into the image's record
__global__ void conv_metrics_fold_kernel(const float* __restrict__ partials, int chunks, int64_t B,
                                         float* __restrict__ items) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float r2 = 0.f, l1 = 0.f, l0 = 0.f, lo = INFINITY, hi = -INFINITY;
  for (int j = lane; j < chunks; j += 32) {
    const float* p = partials + (b * chunks + j) * METRIC_FIELDS;
    r2 += p[0], l1 += p[1], l0 += p[2], lo = fminf(lo, p[3]), hi = fmaxf(hi, p[4]);
  }
  r2 = warp_sum(r2), l1 = warp_sum(l1), l0 = warp_sum(l0), lo = warp_min(lo), hi = warp_max(hi);
  if (lane == 0) {
    float* o = items + b * METRIC_FIELDS;
    o[0] = r2, o[1] = l1, o[2] = l0, o[3] = lo, o[4] = hi;
  }
}

// batch totals over the item records, one block, fp64:
//   out[0] = sum_b 0.5 * r2_b            (LASSO l2 component, summed)
//   out[1] = sum_b l1_b                  (lagrange component / sparsity_weight, summed)
//   out[2] = sum_b l0_b / codes_per_item (normalised l0, summed)
//   out[3] = sum over items with mse != 0 of log10(mse_b), mse_b = r2_b / pixels_per_item;  out[4] = their number
//   out[5], out[6] = min, max pixel of the batch;  out[7] = number of items
__global__ void metrics_totals_kernel(const float* __restrict__ items, int64_t B, double pixels_per_item,
                                      double codes_per_item, double* __restrict__ out) {
  __shared__ double red[7][32];
  double l2 = 0.0, l1 = 0.0, l0 = 0.0, lg = 0.0, nf = 0.0, lo = INFINITY, hi = -INFINITY;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const float* p = items + b * METRIC_FIELDS;
    l2 += 0.5 * static_cast<double>(p[0]);
    l1 += static_cast<double>(p[1]);
    l0 += static_cast<double>(p[2]) / codes_per_item;
    // the reference forms the mean squared error in fp32 (numpy mean of an fp32 array)
    const float mse = static_cast<float>(static_cast<double>(p[0]) / pixels_per_item);
    if (mse != 0.f) lg += log10(static_cast<double>(mse)), nf += 1.0;
    lo = fmin(lo, static_cast<double>(p[3])), hi = fmax(hi, static_cast<double>(p[4]));
  }
  double v[7] = {l2, l1, l0, lg, nf, lo, hi};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int f = 0; f < 7; ++f) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = __shfl_xor_sync(0xffffffffu, v[f], o);
      v[f] = (f == 5) ? fmin(v[f], other) : (f == 6) ? fmax(v[f], other) : v[f] + other;
    }
    if (lane == 0) red[f][warp] = v[f];
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int f = 0; f < 7; ++f) {
      double x = (lane < nw) ? red[f][lane] : (f == 5 ? INFINITY : f == 6 ? -INFINITY : 0.0);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, x, o);
        x = (f == 5) ? fmin(x, other) : (f == 6) ? fmax(x, other) : x + other;
      }
      if (lane == 0) out[f] = x;
    }
    if (lane == 0) out[7] = static_cast<double>(B);
  }
}

// mean over the pixels of |dictionary - previous| per dictionary element (training/sparse_coding.py:226-228)
__global__ void dict_change_kernel(const float* __restrict__ dict, const float* __restrict__ prev, int64_t S,
                                   int64_t per_kernel, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (s >= S) return;
  float acc = 0.f;
  for (int64_t i = lane; i < per_kernel; i += 32) acc += fabsf(dict[s * per_kernel + i] - prev[s * per_kernel + i]);
  acc = warp_sum(acc);
  if (lane == 0) out[s] = acc / static_cast<float>(per_kernel);
}

// ------------------------------------------------------------------------------------------------------------
// Data feed: centre-surround whitening of whole images in the DFT domain
// (utils/image_processing.py:267-308). The transfer function is real and depends on the DFT size only:
//   raw(i, j) = max(|f|, low) * exp(-(|f| / (0.5 * high))^order),  |f| = hypot(fftfreq(h)[i], fftfreq(w)[j])
// normalised to a maximum of 1 and floored at 1e-3 when norm_and_threshold (:300-302). fp64 like numpy; the FFTs
// themselves are cuFFT (torch.fft), only the filter and its application are kernels here.
__device__ __forceinline__ double whitening_raw(int64_t i, int64_t j, int64_t h, int64_t w, double low, double high,
                                                double order) {
  // np.fft.fftfreq(n)[i] = i / n for i < (n - 1) / 2 + 1, (i - n) / n after that
  const double fv = static_cast<double>(i < (h - 1) / 2 + 1 ? i : i - h) / static_cast<double>(h);
  const double fh = static_cast<double>(j < (w - 1) / 2 + 1 ? j : j - w) / static_cast<double>(w);
  const double mag = sqrt(fv * fv + fh * fh);
  return fmax(mag, low) * exp(-pow(mag / (0.5 * high), order));
}
// *max_bits = max over the grid of raw (a non-negative double compares like its bit pattern)
__global__ void whitening_filter_max_kernel(int64_t h, int64_t w, double low, double high, double order,
                                            unsigned long long* __restrict__ max_bits) {
  double m = 0.0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < h * w;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    m = fmax(m, whitening_raw(i / w, i % w, h, w, low, high, order));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits, static_cast<unsigned long long>(__double_as_longlong(m)));
}
__global__ void whitening_filter_kernel(int64_t h, int64_t w, double low, double high, double order,
                                        const unsigned long long* __restrict__ max_bits, int norm_and_threshold,
                                        float* __restrict__ out) {
  const double mx = norm_and_threshold ? __longlong_as_double(static_cast<long long>(*max_bits)) : 1.0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < h * w;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    double v = whitening_raw(i / w, i % w, h, w, low, high, order);
    if (norm_and_threshold) {
      v /= mx;
      if (v < 1e-3) v = 1e-3;
    }
    out[i] = static_cast<float>(v);
  }
}
// spectrum (n, h*w, c) complex64, in place: every channel of every image times the real transfer function (h*w)
__global__ void spectrum_filter_kernel(float2* __restrict__ spectrum, int64_t n, int64_t hw, int64_t c,
                                       const float* __restrict__ filter) {
  const int64_t total = n * hw * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float f = filter[(i / c) % hw];
    float2 v = spectrum[i];
    v.x *= f, v.y *= f;
    spectrum[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Subspace (group) shrinkage for groups WIDER than one epilogue sub-tile (more than 16 atoms; subspace_ista_fista.py:
// 94-96 takes any width): the update of one iteration as a pass of its own over row-major arrays, one warp per
// (patch, group):   y = a1 + beta_prev (a1 - a2);  u = y - eta * grad;  a = u * max(1 - theta / ||u_g||, 0)  with
// ||u_g|| = 0 -> 1 (subspace_ista_fista.py:144-156);  y' = a + beta_next (a - a1)  -> bf16 parts (the next operand).
// a is written over a2 (every element is read before it is written, by the same thread). W = group width (a multiple
// of 32), ld = pitch of the fp32 arrays, y parts in the operand layout of split_rows_kernel.
__global__ void wide_group_prox_kernel(const float* __restrict__ grad, const float* __restrict__ a1, float* a2_out,
                                       int64_t ld, int64_t B, int64_t S, int W, const float* __restrict__ scalars,
                                       float beta_prev, float beta_next, int use_momentum, int has_prev2,
                                       __nv_bfloat16* __restrict__ yparts, int64_t Kp, int nparts, int block,
                                       double* __restrict__ stat) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t groups_per_row = S / W;
  const float eta = scalars[0], theta = scalars[1];
  float stat_local = 0.f;
  for (int64_t item = warp; item < B * groups_per_row; item += nwarps) {
    const int64_t r = item / groups_per_row;
    const int64_t c0 = (item - r * groups_per_row) * W;
    float ss = 0.f;
    for (int c = lane; c < W; c += 32) {
      const int64_t i = r * ld + c0 + c;
      const float ak = a1[i];
      float y = ak;
      if (has_prev2) y = __fadd_rn(ak, __fmul_rn(beta_prev, __fsub_rn(ak, a2_out[i])));
      const float u = __fsub_rn(y, __fmul_rn(eta, grad[i]));
      ss = __fadd_rn(ss, __fmul_rn(u, u));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    float nrm = sqrtf(ss);
    if (nrm == 0.f) nrm = 1.f;
    const float scale = fmaxf(__fsub_rn(1.f, __fdiv_rn(theta, nrm)), 0.f);
    for (int c = lane; c < W; c += 32) {
      const int64_t col = c0 + c;
      const int64_t i = r * ld + col;
      const float ak = a1[i];
      float y = ak;
      if (has_prev2) y = __fadd_rn(ak, __fmul_rn(beta_prev, __fsub_rn(ak, a2_out[i])));
      const float u = __fsub_rn(y, __fmul_rn(eta, grad[i]));
      const float a = __fmul_rn(u, scale);
      const float d = __fsub_rn(a, ak);
      float v = use_momentum ? __fadd_rn(a, __fmul_rn(beta_next, d)) : a;
      a2_out[i] = a;
      stat_local += fabsf(d);
      if (yparts != nullptr) {
        for (int p = 0; p < nparts; ++p) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          v = __fsub_rn(v, __bfloat162float(h));
          const int64_t idx = block ? ((p * (Kp / block) + col / block) * B + r) * block + col % block
                                    : r * (static_cast<int64_t>(nparts) * Kp) + p * Kp + col;
          yparts[idx] = h;
        }
      }
    }
  }
  if (stat != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) stat_local += __shfl_xor_sync(0xffffffffu, stat_local, o);
    if (lane == 0) atomicAdd(stat, static_cast<double>(stat_local));
  }
}

// ------------------------------------------------------------------------------------------------------------
// FISTA momentum table (ista_fista.py:123-125): t_1 = 1, t_{k+1} = (1 + sqrt(1 + 4 t_k^2)) / 2,
// betas[k] = float((t_k - 1) / t_{k+1}), betas[0] = 0; all zero for ISTA. The recurrence is sequential: one thread, IEEE
// double operations without contraction, so the table is bit-identical to the host-side double arithmetic of the
// reference. (A device table instead of a host-to-device copy keeps a call capturable into a CUDA graph.)
// Window form: entries k_lo .. num_iters are written to betas[k - k_lo] (runs longer than the table take it in windows).
__global__ void fista_betas_kernel(float* __restrict__ betas, int num_iters, int fista, int k_lo = 0) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (k_lo == 0) betas[0] = 0.f;
  double t = 1.0;
  for (int k = 1; k <= num_iters; ++k) {
    const double t_next = __ddiv_rn(__dadd_rn(1.0, __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(__dmul_rn(4.0, t), t)))), 2.0);
    if (k >= k_lo) betas[k - k_lo] = fista ? static_cast<float>(__ddiv_rn(__dadd_rn(t, -1.0), t_next)) : 0.f;
    t = t_next;
  }
}

}  // namespace vtc
