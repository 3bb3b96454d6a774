// C ABI (include/vtc_b200.h) and host-side orchestration of the sm_100a kernels.
#include "../../include/vtc_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "aux_kernels.cuh"
#include "gemm_kernel.cuh"
#include "fista_iter_kernel.cuh"
#include "fista_iter2_kernel.cuh"
#include "fista_small_kernel.cuh"

namespace {

using namespace vtc;

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (expr);                                                                            \
    if (e__ != cudaSuccess) return fail(VTC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define TRY(expr)               \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != VTC_OK) return rc__; \
  } while (0)

// every kernel launch of this library goes through COUNT_LAUNCH so that callers can report how many ran
long long g_launches = 0;
#define COUNT_LAUNCH() (++g_launches)

constexpr int kProfSamples = 16;
constexpr int kProfHistory = 64;   // profiled vtc_fista_fc calls kept (ring): begin / iter_begin / iter_end / end events
struct Profile {
  bool on = false, valid = false;
  cudaEvent_t begin = nullptr, iter_begin = nullptr, iter_end = nullptr;   // of the last profiled call (ring slots)
  cudaEvent_t h_begin[kProfHistory] = {}, h_iter_begin[kProfHistory] = {}, h_iter_end[kProfHistory] = {},
              h_end[kProfHistory] = {};
  long long calls = 0;     // profiled calls since profiling was switched on
  bool events_made = false;
  cudaEvent_t k1_begin[kProfSamples], k1_end[kProfSamples], k2_end[kProfSamples];  // sampled iterations
  int iter_launches = 0, iters = 0, samples = 0;
  int launch_iters = 1;   // iterations per sampled launch (the persistent schedule runs all of them in one)
  bool two_launches = false;
};
Profile g_prof;

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---------------------------------------------------------------------------------------------- device properties
struct DeviceInfo {
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  bool ok = false;
};
int device_info(DeviceInfo* out) {
  static thread_local DeviceInfo cache[64];
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(VTC_ERR_CUDA, "device ordinal %d out of range", dev);
  if (!cache[dev].ok) {
    CUDA_TRY(cudaDeviceGetAttribute(&cache[dev].sm_count, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&cache[dev].cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&cache[dev].cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    cache[dev].ok = true;
  }
  *out = cache[dev];
  return VTC_OK;
}
int require_sm100(DeviceInfo* info) {
  TRY(device_info(info));
  if (info->cc_major != 10)
    return fail(VTC_ERR_CUDA, "vtc_b200 needs an sm_100 (B200) device, found sm_%d%d; there is no fallback path",
                info->cc_major, info->cc_minor);
  return VTC_OK;
}

int tune_flags();

// ---------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int get_encode(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) return fail(VTC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    cached = reinterpret_cast<EncodeTiledFn>(p);
  }
  *fn = cached;
  return VTC_OK;
}
int encode(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw, const char* what) {
  EncodeTiledFn fn;
  TRY(get_encode(&fn));
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(VTC_ERR_ARG, "%s: base pointer not 16-byte aligned", what);
  for (int i = 0; i < rank - 1; ++i)
    if (gstr[i] % 16 != 0) return fail(VTC_ERR_ARG, "%s: stride %llu not a multiple of 16 bytes", what, (unsigned long long)gstr[i]);
  // operands (re-read from L2 by several tiles): 256-byte promotion; fp32 state tiles (64-byte rows, read once): none
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 && !(tune_flags() & TUNE_PROMO_256)) promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  CUresult r = fn(m, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VTC_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return VTC_OK;
}

// A bf16 "parts" matrix: rows x (parts * Kp), part p = columns [p*Kp, p*Kp + K), zero padded to Kp (multiple of 64).
struct PartsMat {
  const void* ptr = nullptr;
  int64_t rows = 0, K = 0, Kp = 0;
  int parts = 0;
  int block = 0;  // 0: row-major; else tile-contiguous [part][Kp/block][rows][block] (same size in bytes)
  int64_t pitch_elems() const { return static_cast<int64_t>(parts) * Kp; }
  size_t bytes() const { return static_cast<size_t>(rows) * pitch_elems() * 2; }
};
struct F32Mat {
  const void* ptr = nullptr;
  int64_t rows = 0, cols = 0, ld = 0;
  bool blocked = false;  // tile-contiguous [ceil(cols/16)][rows][16] instead of row-major with pitch ld
};

int map_operand(CUtensorMap* m, const PartsMat& a, int block_k, const char* what, int box_rows = BLOCK_M,
                int blocked_box_rows = BLOCK_M) {
  if (a.block) {
    if (a.block != block_k) return fail(VTC_ERR_ARG, "%s: blocked operand of width %d read with K block %d", what, a.block, block_k);
    const uint64_t dims[3] = {static_cast<uint64_t>(a.block), static_cast<uint64_t>(a.rows),
                              static_cast<uint64_t>(a.parts * (a.Kp / a.block))};
    const uint64_t str[2] = {static_cast<uint64_t>(a.block) * 2, static_cast<uint64_t>(a.rows) * a.block * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(block_k), static_cast<uint32_t>(blocked_box_rows), 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a.ptr, dims, str, box,
                  block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, what);
  }
  const uint64_t dims[2] = {static_cast<uint64_t>(a.pitch_elems()), static_cast<uint64_t>(a.rows)};
  const uint64_t str[1] = {static_cast<uint64_t>(a.pitch_elems()) * 2};
  const uint32_t box[2] = {static_cast<uint32_t>(block_k), static_cast<uint32_t>(box_rows)};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.ptr, dims, str, box,
                block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : block_k == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, what);
}
int map_f32(CUtensorMap* m, const F32Mat& a, const char* what) {
  if (a.blocked) {
    const uint64_t dims[3] = {EPI_COLS, static_cast<uint64_t>(a.rows), static_cast<uint64_t>(ceil_div(a.cols, EPI_COLS))};
    const uint64_t str[2] = {EPI_COLS * 4, static_cast<uint64_t>(a.rows) * EPI_COLS * 4};
    const uint32_t box[3] = {EPI_COLS, BLOCK_M, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a.ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, what);
  }
  const uint64_t dims[2] = {static_cast<uint64_t>(a.cols), static_cast<uint64_t>(a.rows)};
  const uint64_t str[1] = {static_cast<uint64_t>(a.ld) * 4};
  const uint32_t box[2] = {EPI_COLS, BLOCK_M};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B, what);
}
// The epilogue's bf16 split store: part p of sub-tile (row0, col0) goes to columns p * Kp + col0. Columns in a part's
// zero padding [K, Kp) may be written, always with zeros (the accumulator and every input are zero there).
int map_parts_out(CUtensorMap* m, const PartsMat& a, const char* what) {
  if (a.block) {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.block), static_cast<uint64_t>(a.rows),
                              static_cast<uint64_t>(a.parts * (a.Kp / a.block))};
    const uint64_t str[2] = {static_cast<uint64_t>(a.block) * 2, static_cast<uint64_t>(a.rows) * a.block * 2};
    const uint32_t box[3] = {EPI_COLS, BLOCK_M, 1};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a.ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B, what);
  }
  const uint64_t dims[2] = {static_cast<uint64_t>(a.pitch_elems()), static_cast<uint64_t>(a.rows)};
  const uint64_t str[1] = {static_cast<uint64_t>(a.pitch_elems()) * 2};
  const uint32_t box[2] = {EPI_COLS, BLOCK_M};
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.ptr, dims, str, box, CU_TENSOR_MAP_SWIZZLE_32B, what);
}

// ---------------------------------------------------------------------------------------------- precision
int parts_for(int precision) { return precision == VTC_PRECISION_BF16 ? 1 : precision == VTC_PRECISION_BF16X3 ? 2 : 3; }
bool valid_precision(int p) { return p == VTC_PRECISION_BF16 || p == VTC_PRECISION_BF16X3 || p == VTC_PRECISION_BF16X6; }
// Tuning switches of the GEMM kernel (TuneFlags); VTC_B200_FLAGS overrides the default for experiments.
int tune_flags() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("VTC_B200_FLAGS");
    cached = e ? atoi(e) : 0;
  }
  return cached;
}

// How one ISTA/FISTA iteration is contracted:
//   gram      : acc = y G,  G = Phi Phi^T, b = x Phi^T precomputed        2*S*S flops per patch, 1 launch
//   synthesis : r = y Phi - x, then acc = r Phi^T (the reference's own form, ista_fista.py:105-106)
//                                                                          4*S*D flops per patch, 2 launches
// auto picks the cheaper one: synthesis when S > 2 D.
enum Formulation { FORM_AUTO = 0, FORM_GRAM = 1, FORM_SYNTHESIS = 2 };
int g_formulation = -1;
int formulation_for(int64_t S, int64_t D) {
  if (g_formulation < 0) {
    const char* e = getenv("VTC_B200_FORMULATION");
    g_formulation = e ? atoi(e) : FORM_AUTO;
  }
  if (g_formulation == FORM_GRAM || g_formulation == FORM_SYNTHESIS) return g_formulation;
  return S > 2 * D ? FORM_SYNTHESIS : FORM_GRAM;
}

// K blocks the kernel walks for a padded K extent: 64 columns per stage for plain bf16, 32 for the split modes.
int64_t k_blocks_for(int64_t Kp, int precision) { return Kp / (parts_for(precision) == 1 ? 64 : 32); }

// ---------------------------------------------------------------------------------------------- GEMM launch
struct GemmCall {
  PartsMat A, B;         // D[M,N] = A[M,K] * B[N,K]^T
  int precision = VTC_PRECISION_BF16X3;
  int64_t M = 0, N = 0, K = 0;
  F32Mat in[3];
  int in_mask = 0;       // bit i: in[i] is loaded by the epilogue
  F32Mat out;            // fp32 output (optional)
  bool store_out = false;
  PartsMat parts_out;    // bf16 split output (optional)
  int n_parts = 0;
  int ksplits = 1;       // >1: fp32 partials stacked along rows of `out`, out_rows_per_split apart
  int64_t out_rows_per_split = 0;
  // FISTA epilogue
  int prox = 0, group = 1, use_momentum = 0;
  float beta_prev = 0.f, beta_next = 0.f;
  const float* scalars = nullptr;
  double* stat = nullptr;
  int max_pairs = 0;     // cap on the SM pairs used (0 = all): concurrent chains share the chip
  // segmented K and grid masks of the convolutional path (GemmParams has the semantics)
  int seg_kb = 0, nseg = 0;
  int seg_shift[MAX_SEGMENTS] = {};
  // halo staging of the taps (GemmParams): bands per stage (0 = off), row of each band, band and row of each tap
  int halo_bands = 0;
  int halo_band_row[4] = {};
  int halo_tap_band[MAX_SEGMENTS] = {}, halo_tap_row[MAX_SEGMENTS] = {};
  int grid_h = 0, grid_w = 0, code_h = 0, code_w = 0;
  int blk_sy = 1, blk_sx = 1, pix_y0 = 0, pix_y1 = 0, pix_x0 = 0, pix_x1 = 0;
};

template <int EPI, int P, int NIN, int BN, int RES = 0, int NBANDS = 0>
int launch_gemm_p(const GemmCall& c, const DeviceInfo& info, cudaStream_t stream) {
  using Cf = Cfg<P, NIN, BN, RES, NBANDS>;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  TRY(map_operand(&p.tmA, c.A, Cf::BK, "A operand"));
  if (NBANDS > 0) {
    if (c.halo_bands != NBANDS || !c.A.block) return fail(VTC_ERR_ARG, "halo tile variant used outside its range");
    TRY(map_operand(&p.tmAh, c.A, Cf::BK, "A operand (halo bands)", BLOCK_M, HALO_ROWS));
    for (int b = 0; b < NBANDS; ++b) p.halo_band_row[b] = c.halo_band_row[b];
    for (int q = 0; q < c.nseg; ++q) p.halo_tap_band[q] = c.halo_tap_band[q], p.halo_tap_row[q] = c.halo_tap_row[q];
  }
  TRY(map_operand(&p.tmB, c.B, Cf::BK, "B operand", Cf::HALF_N));
  for (int i = 0; i < 3; ++i)
    if (c.in_mask & (1 << i)) TRY(map_f32(&p.tmIn[i], c.in[i], "epilogue input"));
  if (c.store_out) TRY(map_f32(&p.tmOut, c.out, "fp32 output"));
  if (c.n_parts) TRY(map_parts_out(&p.tmParts, c.parts_out, "bf16 parts output"));
  p.M = static_cast<int>(c.M);
  p.N = static_cast<int>(c.N);
  p.num_m_blocks = static_cast<int>(ceil_div(c.M, PAIR_M));
  p.num_n_blocks = static_cast<int>(ceil_div(c.N, BN));
  p.k_blocks = static_cast<int>(ceil_div(c.B.Kp, Cf::BK));  // Kp is a multiple of 64; the padding is zero
  if (c.nseg > 0) {
    // A holds one segment's columns; B holds all nseg segments side by side
    p.seg_kb = static_cast<int>(c.A.Kp / Cf::BK);
    if (c.nseg > MAX_SEGMENTS || c.B.Kp != c.A.Kp * c.nseg) return fail(VTC_ERR_ARG, "bad K segmentation");
    for (int i = 0; i < c.nseg; ++i) p.seg_shift[i] = c.seg_shift[i];
  }
  p.grid_h = c.grid_h, p.grid_w = c.grid_w, p.code_h = c.code_h, p.code_w = c.code_w;
  p.blk_sy = c.blk_sy, p.blk_sx = c.blk_sx;
  p.pix_y0 = c.pix_y0, p.pix_y1 = c.pix_y1, p.pix_x0 = c.pix_x0, p.pix_x1 = c.pix_x1;
  p.a_part_stride = static_cast<int>(c.A.Kp);
  p.b_part_stride = static_cast<int>(c.B.Kp);
  p.out_part_stride = static_cast<int>(c.parts_out.Kp);
  p.kb_per_split = static_cast<int>(ceil_div(p.k_blocks, c.ksplits));
  p.ksplits = static_cast<int>(ceil_div(p.k_blocks, p.kb_per_split));
  p.out_rows_per_split = static_cast<int>(c.out_rows_per_split);
  p.in_mask = c.in_mask;
  p.n_parts = c.n_parts;
  p.store_out = c.store_out ? 1 : 0;
  p.prox = c.prox;
  p.group = c.group;
  p.use_momentum = c.use_momentum;
  p.beta_prev = c.beta_prev;
  p.beta_next = c.beta_next;
  p.scalars = c.scalars;
  p.stat = c.stat;
  p.flags = tune_flags();
  for (int i = 0; i < 3; ++i)
    if ((c.in_mask & (1 << i)) && c.in[i].blocked) p.blocked_mask |= (BLK_IN0 << i);
  if (c.store_out && c.out.blocked) p.blocked_mask |= BLK_OUT;
  if (c.n_parts && c.parts_out.block) {
    p.blocked_mask |= BLK_PARTS;
    p.parts_block_w = c.parts_out.block;
    p.parts_blocks_per_part = static_cast<int>(c.parts_out.Kp / c.parts_out.block);
  }
  if (c.A.block) {
    p.blocked_mask |= BLK_A;
    p.a_blocks_per_part = static_cast<int>(c.A.Kp / c.A.block);
  }
  if (RES > 0 && (p.num_n_blocks != 1 || p.k_blocks > RES || p.ksplits != 1))
    return fail(VTC_ERR_ARG, "resident-B tile variant used outside its range");
  const long long tiles = 1ll * p.num_m_blocks * p.num_n_blocks * p.ksplits;
  if (tiles > 0x7fffffffll) return fail(VTC_ERR_ARG, "too many tiles");
  static bool attr_set_dev[64] = {};  // the attribute is per device
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  bool& attr_set = attr_set_dev[dev & 63];
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(vtc_gemm_kernel<EPI, P, NIN, BN, RES, NBANDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM_ALLOC));
    attr_set = true;
  }
  long long max_pairs = info.sm_count / 2;
  if (c.max_pairs > 0 && c.max_pairs < max_pairs) max_pairs = c.max_pairs;
  const int pairs = static_cast<int>(tiles < max_pairs ? tiles : max_pairs);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(Cf::THREADS);
  cfg.dynamicSmemBytes = Cf::SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // prologue overlaps the previous launch's tail
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (tune_flags() & TUNE_NO_PDL) ? 1 : 2;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, vtc_gemm_kernel<EPI, P, NIN, BN, RES, NBANDS>, p));
  COUNT_LAUNCH();
  return VTC_OK;
}

// Convolutional launches whose whole B operand (N <= 64 outputs, <= 8 K blocks) stays resident in shared memory:
// 64-wide tiles, the ring carries A only. VTC_B200_CONV_RESIDENT=0 keeps the 128-wide streaming tiles.
bool conv_resident_b(const GemmCall& c, int P) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("VTC_B200_CONV_RESIDENT");
    enabled = e ? (atoi(e) != 0) : 1;
  }
  const int64_t bk = (P == 1) ? 64 : 32;
  return enabled && c.nseg > 0 && P <= 2 && c.N <= 64 && c.ksplits == 1 && c.B.Kp / bk <= 8;
}

template <int EPI>
int launch_gemm(const GemmCall& c, cudaStream_t stream) {
  DeviceInfo info;
  TRY(require_sm100(&info));
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) return fail(VTC_ERR_ARG, "empty GEMM %lld x %lld x %lld", (long long)c.M, (long long)c.N, (long long)c.K);
  if (c.nseg == 0 && (c.A.K != c.K || c.B.K != c.K || c.A.Kp != c.B.Kp)) return fail(VTC_ERR_ARG, "operand K mismatch");
  const int P = parts_for(c.precision);
  if (c.A.parts < P || c.B.parts < P) return fail(VTC_ERR_ARG, "operand has too few bf16 parts for precision %d", c.precision);
  if (c.n_parts > MAX_PARTS || (c.n_parts && c.parts_out.parts < c.n_parts)) return fail(VTC_ERR_ARG, "bad parts output");
  if (c.n_parts > P) return fail(VTC_ERR_ARG, "at most %d output parts at precision %d", P, c.precision);
  // epilogue input slots in use: plain GEMMs at most 1, fused update 2 (synthesis form) or 3 (Gram form, reads b)
  const int nin = __builtin_popcount(c.in_mask);
  if (EPI == EPI_STORE) {
    if (nin > 1) return fail(VTC_ERR_ARG, "EPI_STORE takes at most one epilogue input");
    // (a 128-wide tile variant, Cfg<.., 128>, fills the 74 SM pairs better for N = 256 -- 512 instead of 256 tiles --
    //  but measured 10-15 % slower: the A panel is staged twice and the N = 128 MMA is shared-memory bound)
    // narrow outputs (the convolutional path: N = code channels or pixels per block) take the 128-wide tile so that
    // fewer MMA columns are spent on zero padding
    if (conv_resident_b(c, P)) {
      if (c.halo_bands == 2)
        return P == 1 ? launch_gemm_p<EPI_STORE, 1, 1, 64, 8, 2>(c, info, stream)
                      : launch_gemm_p<EPI_STORE, 2, 1, 64, 8, 2>(c, info, stream);
      return P == 1 ? launch_gemm_p<EPI_STORE, 1, 1, 64, 8>(c, info, stream)
                    : launch_gemm_p<EPI_STORE, 2, 1, 64, 8>(c, info, stream);
    }
    if (c.N <= 128 && c.nseg > 0) {
      switch (P) {
        case 1: return launch_gemm_p<EPI_STORE, 1, 1, 128>(c, info, stream);
        case 2: return launch_gemm_p<EPI_STORE, 2, 1, 128>(c, info, stream);
        default: return launch_gemm_p<EPI_STORE, 3, 1, 128>(c, info, stream);
      }
    }
    switch (P) {
      case 1: return launch_gemm_p<EPI_STORE, 1, 1, 256>(c, info, stream);
      case 2: return launch_gemm_p<EPI_STORE, 2, 1, 256>(c, info, stream);
      default: return launch_gemm_p<EPI_STORE, 3, 1, 256>(c, info, stream);
    }
  }
  if (nin <= 2) {
    if (conv_resident_b(c, P)) {
      if (c.halo_bands == 2)
        return P == 1 ? launch_gemm_p<EPI_FISTA, 1, 2, 64, 8, 2>(c, info, stream)
                      : launch_gemm_p<EPI_FISTA, 2, 2, 64, 8, 2>(c, info, stream);
      return P == 1 ? launch_gemm_p<EPI_FISTA, 1, 2, 64, 8>(c, info, stream)
                    : launch_gemm_p<EPI_FISTA, 2, 2, 64, 8>(c, info, stream);
    }
    if (c.N <= 128 && c.nseg > 0) {
      switch (P) {
        case 1: return launch_gemm_p<EPI_FISTA, 1, 2, 128>(c, info, stream);
        case 2: return launch_gemm_p<EPI_FISTA, 2, 2, 128>(c, info, stream);
        default: return launch_gemm_p<EPI_FISTA, 3, 2, 128>(c, info, stream);
      }
    }
    // a long contraction (D >= 512, configs[3]): the deep operand ring (Cfg RES = -1); K = 256 keeps the split made
    // for the HBM-bound case
    if (c.K >= 512 && c.nseg == 0 && P <= 2)
      return P == 1 ? launch_gemm_p<EPI_FISTA, 1, 2, 256, -1>(c, info, stream)
                    : launch_gemm_p<EPI_FISTA, 2, 2, 256, -1>(c, info, stream);
    switch (P) {
      case 1: return launch_gemm_p<EPI_FISTA, 1, 2, 256>(c, info, stream);
      case 2: return launch_gemm_p<EPI_FISTA, 2, 2, 256>(c, info, stream);
      default: return launch_gemm_p<EPI_FISTA, 3, 2, 256>(c, info, stream);
    }
  }
  // Gram-form iteration. Small batches (BASELINE configs[0]: 250 patches = one row of tiles) are latency-bound: 64-wide
  // tiles put four times as many SM pairs to work on the same launch and quarter every pair's epilogue
  if (P <= 2 && ceil_div(c.M, PAIR_M) * ceil_div(c.N, 256) * 4 <= info.sm_count / 2) {
    return P == 1 ? launch_gemm_p<EPI_FISTA, 1, 3, 64>(c, info, stream) : launch_gemm_p<EPI_FISTA, 2, 3, 64>(c, info, stream);
  }
  switch (P) {
    case 1: return launch_gemm_p<EPI_FISTA, 1, 3, 256>(c, info, stream);
    case 2: return launch_gemm_p<EPI_FISTA, 2, 3, 256>(c, info, stream);
    default: return launch_gemm_p<EPI_FISTA, 3, 3, 256>(c, info, stream);
  }
}

// ---------------------------------------------------------------------------------------------- fused iteration
// One launch = one ISTA/FISTA iteration of the synthesis form with the operand y_k kept on chip (fista_iter_kernel.cuh).
struct IterCall {
  PartsMat r_op, phi_op, phiT_op;
  int precision = VTC_PRECISION_BF16X3;
  int64_t B = 0, S = 0, D = 0;
  F32Mat state[4];       // a_0 (starting point), a_k for odd k, a_k for even k, where the final iterate goes
  F32Mat x;
  int k_first = 1, k_count = 1;   // iterations run by this launch
  int k_final = 0x7fffffff;       // iteration whose output is the result (none: the caller stops on its own criterion)
  const float* betas = nullptr;   // device table, betas[k] = momentum coefficient of iteration k, betas[0] = 0
  int* done = nullptr;            // device per-panel completion counters, zeroed (required when k_count > 1)
  int prox = 0, group = 1, use_momentum = 0;
  const float* scalars = nullptr;
  double* stat = nullptr;
  int max_pairs = 0;
  // second-generation kernel (fista_iter2_kernel.cuh): quad-blocked state arrays and the images read directly
  float* qstate[3] = {nullptr, nullptr, nullptr};
  bool init_zero = false;
};

unsigned long long* g_iter_trace = nullptr;  // vtc_debug_iter_trace

// VTC_B200_FUSED_ITER=0 keeps the two-launch schedule (for comparison and as the reference of the parity tests)
int g_fused_iter = -1;
bool fused_iter_enabled() {
  if (g_fused_iter < 0) {
    const char* e = getenv("VTC_B200_FUSED_ITER");
    g_fused_iter = e ? (atoi(e) != 0) : 1;
  }
  return g_fused_iter != 0;
}
// VTC_B200_PERSISTENT=0: one launch per iteration instead of one launch for all of them
bool persistent_iterations_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("VTC_B200_PERSISTENT");
    cached = e ? (atoi(e) != 0) : 1;
  }
  return cached != 0;
}
bool fused_iter_ok(int64_t S, int64_t D, int precision) {
  return fused_iter_enabled() && formulation_for(S, D) == FORM_SYNTHESIS && D <= IT_RN && parts_for(precision) <= 2;
}

// The panel-resident kernel: default 2 = fista_iter2_kernel.cuh (a_{k-2} read directly from a coalescing-friendly state
// layout, 8 KB in/out stages, three G stages: -9.5 % against the first generation on configs[1] bf16x3, same-box,
// profiles/README.md); VTC_B200_ITER_GEN=1 = fista_iter_kernel.cuh
int iter_generation() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("VTC_B200_ITER_GEN");
    cached = e ? atoi(e) : 2;
    if (cached != 1) cached = 2;
  }
  return cached;
}
int ablate_flags() {
  static int ablate = -1;   // timing experiments (tools/ablate.sh): pieces of the kernel left out, results wrong
  if (ablate < 0) {
    const char* e = getenv("VTC_B200_ABLATE");
    ablate = e ? atoi(e) : 0;
  }
  return ablate;
}

// tuning variant of the fused iteration kernel (stage counts / math warps); VTC_B200_ITER_VARIANT overrides the
// default: bf16x3 is bound by shared-memory bandwidth and does best with three math groups (variant 0), plain bf16 is
// bound by the state bytes in flight and does best with the deepest input ring (variant 1: 8 stages, two math groups)
int iter_variant(int parts) {
  static int cached = -2;
  if (cached == -2) {
    const char* e = getenv("VTC_B200_ITER_VARIANT");
    cached = e ? atoi(e) : -1;
    if (cached >= IT_VARIANTS) cached = -1;
  }
  return cached >= 0 ? cached : (parts == 1 ? 1 : 0);
}

template <int P, int V>
int launch_iter_p(const IterCall& c, const DeviceInfo& info, cudaStream_t stream) {
  using Cf = IterCfg<P, V>;
  IterParams p;
  memset(&p, 0, sizeof(p));
  if (c.r_op.block != Cf::BK) return fail(VTC_ERR_ARG, "fused iteration: r_op must be tile-contiguous with block %d", Cf::BK);
  TRY(map_operand(&p.tmR, c.r_op, Cf::BK, "r operand"));
  TRY(map_operand(&p.tmPhi, c.phi_op, Cf::BK, "dictionary operand", IT_BN / 2));
  TRY(map_operand(&p.tmPhiT, c.phiT_op, Cf::PT_CHUNK, "transposed dictionary operand", IT_RN / 2));
  for (int i = 0; i < 4; ++i) {
    TRY(map_f32(&p.tmState[i], c.state[i], "code array"));
    p.state_blocked[i] = c.state[i].blocked ? 1 : 0;
  }
  if (c.k_count > 1 && c.done == nullptr) return fail(VTC_ERR_ARG, "fused iteration: several iterations per launch need the completion counters");
  p.k_first = c.k_first, p.k_count = c.k_count, p.k_final = c.k_final;
  p.betas = c.betas;
  p.done = c.k_count > 1 ? c.done : nullptr;
  TRY(map_f32(&p.tmX, c.x, "images"));
  TRY(map_parts_out(&p.tmROut, c.r_op, "r parts output"));
  p.num_panels = static_cast<int>(ceil_div(c.B, PAIR_M));
  p.S = static_cast<int>(c.S);
  p.num_n_tiles = static_cast<int>(ceil_div(c.S, IT_BN));
  p.kb_g = static_cast<int>(c.r_op.Kp / Cf::BK);
  p.phi_part_stride = static_cast<int>(c.phi_op.Kp);
  p.phiT_part_stride = static_cast<int>(c.phiT_op.Kp);
  p.nsub_r = static_cast<int>(c.r_op.Kp / EPI_COLS);
  p.r_block_w = Cf::BK;
  p.prox = c.prox, p.group = c.group, p.use_momentum = c.use_momentum;
  p.scalars = c.scalars;
  p.stat = c.stat;
  p.trace = g_iter_trace;
  g_iter_trace = nullptr;  // one shot: only the next launch is traced
  p.ablate = ablate_flags();
  static bool attr_set_dev[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  bool& attr_set = attr_set_dev[dev & 63];
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(vtc_fista_iter_kernel<P, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM_ALLOC));
    attr_set = true;
  }
  long long max_pairs = info.sm_count / 2;
  if (c.max_pairs > 0 && c.max_pairs < max_pairs) max_pairs = c.max_pairs;
  const int pairs = static_cast<int>(p.num_panels < max_pairs ? p.num_panels : max_pairs);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(Cf::THREADS);
  cfg.dynamicSmemBytes = Cf::SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (tune_flags() & TUNE_NO_PDL) ? 1 : 2;
  if (c.k_count > 1) {
    // jobs of a persistent launch wait on jobs of other pairs: every pair of the grid must be resident at once
    static int max_clusters_dev[64] = {};
    int& max_clusters = max_clusters_dev[dev & 63];
    if (max_clusters == 0) {
      cudaLaunchConfig_t query = cfg;
      query.numAttrs = 1;  // cluster dimension only
      CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, vtc_fista_iter_kernel<P, V>, &query));
      if (max_clusters < 1) return fail(VTC_ERR_CUDA, "the iteration kernel does not fit on this device");
    }
    if (pairs > max_clusters) cfg.gridDim = dim3(2 * max_clusters);
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(stream, &capturing));
    // (a stream being captured into a graph cannot wait on an event recorded outside the capture; graph launches are
    // ordered by the graph's own stream)
    if (c.max_pairs == 0 && capturing == cudaStreamCaptureStatusNone) {
      // A launch that may take every SM pair: two of them on different streams could each become half resident and spin
      // on their missing halves for ever. They are ordered through an event instead (each would use the whole device
      // anyway); launches restricted to a share of the pairs (the two-chain mode) are sized to fit side by side.
      static std::mutex mu;
      static cudaEvent_t last_dev[64] = {};
      std::lock_guard<std::mutex> lock(mu);
      cudaEvent_t& last = last_dev[dev & 63];
      if (!last) CUDA_TRY(cudaEventCreateWithFlags(&last, cudaEventDisableTiming));
      CUDA_TRY(cudaStreamWaitEvent(stream, last, 0));   // no-op until the event has been recorded once
      CUDA_TRY(cudaLaunchKernelEx(&cfg, vtc_fista_iter_kernel<P, V>, p));
      COUNT_LAUNCH();
      CUDA_TRY(cudaEventRecord(last, stream));
      return VTC_OK;
    }
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, vtc_fista_iter_kernel<P, V>, p));
  COUNT_LAUNCH();
  return VTC_OK;
}
// Persistent launches that fill the device and wait on one another must not share it with a second such launch on
// another stream: they are ordered through an event (see launch_iter_p).
template <typename Kernel, typename Params>
int launch_persistent(cudaLaunchConfig_t& cfg, Kernel kernel, const Params& p, int pairs, bool whole_device, int dev,
                      int& max_clusters, cudaStream_t stream) {
  if (max_clusters == 0) {
    cudaLaunchConfig_t query = cfg;
    query.numAttrs = 1;  // cluster dimension only
    CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &query));
    if (max_clusters < 1) return fail(VTC_ERR_CUDA, "the iteration kernel does not fit on this device");
  }
  if (pairs > max_clusters) cfg.gridDim = dim3(2 * max_clusters);
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  CUDA_TRY(cudaStreamIsCapturing(stream, &capturing));
  if (whole_device && capturing == cudaStreamCaptureStatusNone) {
    static std::mutex mu;
    static cudaEvent_t last_dev[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    cudaEvent_t& last = last_dev[dev & 63];
    if (!last) CUDA_TRY(cudaEventCreateWithFlags(&last, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamWaitEvent(stream, last, 0));   // no-op until the event has been recorded once
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, p));
    COUNT_LAUNCH();
    CUDA_TRY(cudaEventRecord(last, stream));
    return VTC_OK;
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, p));
  COUNT_LAUNCH();
  return VTC_OK;
}

template <int P, int NG>
int launch_iter2_p(const IterCall& c, const DeviceInfo& info, cudaStream_t stream) {
  using Cf = Iter2Cfg<P, NG>;
  IterParams2 p;
  memset(&p, 0, sizeof(p));
  if (c.r_op.block != Cf::BK) return fail(VTC_ERR_ARG, "fused iteration: r_op must be tile-contiguous with block %d", Cf::BK);
  TRY(map_operand(&p.tmR, c.r_op, Cf::BK, "r operand"));
  TRY(map_operand(&p.tmPhi, c.phi_op, Cf::BK, "dictionary operand", IT_BN / 2));
  TRY(map_operand(&p.tmPhiT, c.phiT_op, Cf::CHUNK, "transposed dictionary operand", IT_RN / 2));
  TRY(map_parts_out(&p.tmROut, c.r_op, "r parts output"));
  TRY(map_f32(&p.tmX, c.x, "images"));
  for (int i = 0; i < 3; ++i) p.state[i] = c.qstate[i];
  p.init_zero = c.init_zero ? 1 : 0;
  if (!p.state[1] || !p.state[2] || (!p.init_zero && !p.state[0])) return fail(VTC_ERR_ARG, "fused iteration: state arrays missing");
  p.x = static_cast<const float*>(c.x.ptr);
  p.ld_x = c.x.ld;
  if ((reinterpret_cast<uintptr_t>(p.x) & 15) != 0 || (p.ld_x % 4) != 0)
    return fail(VTC_ERR_ARG, "fused iteration: images must be 16-byte aligned with a pitch that is a multiple of 4");
  p.B = static_cast<int>(c.B), p.D = static_cast<int>(c.D);
  if (c.k_count > 1 && c.done == nullptr) return fail(VTC_ERR_ARG, "fused iteration: several iterations per launch need the completion counters");
  p.k_first = c.k_first, p.k_count = c.k_count, p.k_final = c.k_final;
  p.betas = c.betas;
  p.done = c.k_count > 1 ? c.done : nullptr;
  p.num_panels = static_cast<int>(ceil_div(c.B, PAIR_M));
  p.S = static_cast<int>(c.S);
  p.num_n_tiles = static_cast<int>(ceil_div(c.S, IT_BN));
  p.kb_g = static_cast<int>(c.r_op.Kp / Cf::BK);
  p.phi_part_stride = static_cast<int>(c.phi_op.Kp);
  p.phiT_part_stride = static_cast<int>(c.phiT_op.Kp);
  p.nsub_r = static_cast<int>(c.r_op.Kp / EPI_COLS);
  p.r_block_w = Cf::BK;
  p.prox = c.prox, p.group = c.group, p.use_momentum = c.use_momentum;
  p.scalars = c.scalars;
  p.stat = c.stat;
  p.ablate = ablate_flags();
  p.trace = g_iter_trace;
  g_iter_trace = nullptr;  // one shot: only the next launch is traced
  {
    static int l2pf = -1;
    if (l2pf < 0) {
      const char* e = getenv("VTC_B200_ITER_L2PF");
      l2pf = e ? atoi(e) : 0;
    }
    p.l2_prefetch = l2pf;
  }
  static bool attr_set_dev[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  bool& attr_set = attr_set_dev[dev & 63];
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(vtc_fista_iter2_kernel<P, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM_ALLOC));
    attr_set = true;
  }
  long long max_pairs = info.sm_count / 2;
  if (c.max_pairs > 0 && c.max_pairs < max_pairs) max_pairs = c.max_pairs;
  {
    // measurement aid: fewer SM pairs (how the iteration time scales with the SMs used tells a per-SM bound from a
    // chip-wide one, profiles/README.md)
    static int limit = -1;
    if (limit < 0) {
      const char* e = getenv("VTC_B200_ITER_PAIRS");
      limit = e ? atoi(e) : 0;
    }
    if (limit > 0 && limit < max_pairs) max_pairs = limit;
  }
  const int pairs = static_cast<int>(p.num_panels < max_pairs ? p.num_panels : max_pairs);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(Cf::THREADS);
  cfg.dynamicSmemBytes = Cf::SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (tune_flags() & TUNE_NO_PDL) ? 1 : 2;
  if (c.k_count > 1) {
    static int max_clusters_dev[64] = {};   // per instantiation and device
    return launch_persistent(cfg, vtc_fista_iter2_kernel<P, NG>, p, pairs, c.max_pairs == 0, dev, max_clusters_dev[dev & 63],
                             stream);
  }
  CUDA_TRY(cudaLaunchKernelEx(&cfg, vtc_fista_iter2_kernel<P, NG>, p));
  COUNT_LAUNCH();
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- small-batch kernel
// VTC_B200_SMALL=0 keeps the tiled Gram-form schedule for small problems (one launch per iteration)
constexpr int64_t kSmallMaxRows = 74 * SM_ROWS;   // one pair per 32 patches, at most one wave of pairs
int g_small_kernel = -1;   // vtc_set_small_batch_kernel / VTC_B200_SMALL
bool small_kernel_ok(int64_t B, int64_t S, int64_t D, int precision, int group_size, bool early) {
  if (g_small_kernel < 0) {
    const char* e = getenv("VTC_B200_SMALL");
    g_small_kernel = e ? (atoi(e) != 0) : 1;
  }
  return g_small_kernel && formulation_for(S, D) == FORM_GRAM && S <= SM_K && B <= kSmallMaxRows && group_size == 1 &&
         !early && parts_for(precision) <= 2;
}
struct SmallCall {
  PartsMat G_op;
  const float* b = nullptr;
  const float* init = nullptr;
  int64_t ld_init = 0;
  float* out = nullptr;
  int64_t ld_out = 0;
  int64_t B = 0, S = 0;
  int num_iters = 0, prox = 0, use_momentum = 0, precision = VTC_PRECISION_BF16X3;
  const float* betas = nullptr;
  const float* scalars = nullptr;
};
template <int P>
int launch_small_p(const SmallCall& c, cudaStream_t stream) {
  using Cf = SmallCfg<P>;
  SmallParams p;
  memset(&p, 0, sizeof(p));
  if (c.G_op.block != 0 || c.G_op.Kp > SM_K || c.G_op.Kp % 64 != 0)
    return fail(VTC_ERR_ARG, "small kernel: G operand must be row-major with at most %d padded columns", SM_K);
  TRY(map_operand(&p.tmG, c.G_op, SM_BK, "Gram operand"));
  p.g_part_stride = static_cast<int>(c.G_op.Kp);
  p.k_blocks = static_cast<int>(c.G_op.Kp / SM_BK);
  p.b = c.b, p.init = c.init, p.ld_init = c.ld_init, p.out = c.out, p.ld_out = c.ld_out;
  p.B = static_cast<int>(c.B), p.S = static_cast<int>(c.S);
  p.num_iters = c.num_iters, p.betas = c.betas, p.prox = c.prox, p.use_momentum = c.use_momentum;
  p.scalars = c.scalars;
  static bool attr_set_dev[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!attr_set_dev[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(vtc_fista_small_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM_ALLOC));
    attr_set_dev[dev & 63] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * static_cast<unsigned>(ceil_div(c.B, SM_ROWS)));
  cfg.blockDim = dim3(Cf::THREADS);
  cfg.dynamicSmemBytes = Cf::SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (tune_flags() & TUNE_NO_PDL) ? 1 : 2;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, vtc_fista_small_kernel<P>, p));
  COUNT_LAUNCH();
  return VTC_OK;
}

int launch_iter(const IterCall& c, cudaStream_t stream) {
  DeviceInfo info;
  TRY(require_sm100(&info));
  if (c.D > IT_RN) return fail(VTC_ERR_ARG, "fused iteration needs D <= %d", IT_RN);
  if (c.qstate[1] != nullptr) {
    static int groups = -1;   // VTC_B200_ITER_GROUPS: math groups of the second-generation kernel (3 or 4)
    if (groups < 0) {
      const char* e = getenv("VTC_B200_ITER_GROUPS");
      groups = (e && atoi(e) == 4) ? 4 : 3;
    }
    if (groups == 4)
      return parts_for(c.precision) == 1 ? launch_iter2_p<1, 4>(c, info, stream) : launch_iter2_p<2, 4>(c, info, stream);
    return parts_for(c.precision) == 1 ? launch_iter2_p<1, 3>(c, info, stream) : launch_iter2_p<2, 3>(c, info, stream);
  }
  const bool one = parts_for(c.precision) == 1;
  switch (iter_variant(parts_for(c.precision))) {
    case 1: return one ? launch_iter_p<1, 1>(c, info, stream) : launch_iter_p<2, 1>(c, info, stream);
    case 2: return one ? launch_iter_p<1, 2>(c, info, stream) : launch_iter_p<2, 2>(c, info, stream);
    case 3: return one ? launch_iter_p<1, 3>(c, info, stream) : launch_iter_p<2, 3>(c, info, stream);
    default: return one ? launch_iter_p<1, 0>(c, info, stream) : launch_iter_p<2, 0>(c, info, stream);
  }
}

// ---------------------------------------------------------------------------------------------- small launches
int grid_for(int64_t work, int threads, int sm_count) {
  int64_t g = ceil_div(work, threads);
  const int64_t cap = static_cast<int64_t>(sm_count) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}
int split_rows(const float* in, int64_t ld, int64_t R, int64_t C, const PartsMat& out, cudaStream_t st,
               float scale = 1.f) {
  DeviceInfo info;
  TRY(device_info(&info));
  split_rows_kernel<<<grid_for(R * out.Kp / 2, 256, info.sm_count), 256, 0, st>>>(
      in, ld, R, C, out.Kp, out.parts, out.block, reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(out.ptr)), scale);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}
// out = parts of in^T : out.rows == C, out.K == R
int transpose_split(const float* in, int64_t ld, int64_t R, int64_t C, const PartsMat& out, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>(out.Kp / 32), static_cast<unsigned>(ceil_div(C, 32)));
  if (grid.y > 65535) return fail(VTC_ERR_ARG, "transpose_split: too many columns (%lld)", (long long)C);
  transpose_split_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ld, R, C, out.Kp, out.parts,
                                                       reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(out.ptr)));
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}
int transpose_f32(const float* in, int64_t ld, int64_t R, int64_t C, float* out, int64_t ldo, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>(ceil_div(R, 32)), static_cast<unsigned>(ceil_div(C, 32)));
  if (grid.y > 65535) return fail(VTC_ERR_ARG, "transpose: too many columns (%lld)", (long long)C);
  transpose_f32_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ld, R, C, out, ldo);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- workspace carving
struct Carver {
  uint8_t* base;
  size_t off = 0, cap;
  bool dry;
  Carver(void* b, size_t c) : base(static_cast<uint8_t*>(b)), cap(c), dry(b == nullptr) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~size_t(1023);
    void* p = dry ? nullptr : base + off;
    off += bytes;
    return p;
  }
  bool fits() const { return dry || off <= cap; }
};
PartsMat carve_parts(Carver& cv, int64_t rows, int64_t K, int parts, int block = 0) {
  PartsMat m;
  m.rows = rows;
  m.K = K;
  m.Kp = round_up(K, 64);
  m.parts = parts;
  m.block = block;
  m.ptr = cv.take(m.bytes());
  return m;
}

constexpr int kSquarings = 28;
struct LipschitzWs {
  double *M, *A0, *A1, *traces;
  int n;
};
LipschitzWs carve_lipschitz(Carver& cv, int64_t D) {
  LipschitzWs w;
  w.n = static_cast<int>(round_up(D, 32));
  const size_t mat = static_cast<size_t>(w.n) * w.n * sizeof(double);
  w.M = static_cast<double*>(cv.take(mat));
  w.A0 = static_cast<double*>(cv.take(mat));
  w.A1 = static_cast<double*>(cv.take(mat));
  w.traces = static_cast<double*>(cv.take((kSquarings + 2) * sizeof(double)));
  return w;
}
// scalars (device float[4]) <- eta, theta, L, status ; lipschitz_dev (device float, optional) <- L
int run_lipschitz(const float* dict, int64_t S, int64_t D, const LipschitzWs& w, float sparsity_weight, float* scalars,
                  float* lipschitz_dev, cudaStream_t st) {
  CUDA_TRY(cudaMemsetAsync(w.traces, 0, (kSquarings + 2) * sizeof(double), st));
  const dim3 grid(w.n / 32, w.n / 32), block(32, 8);
  // One cooperative launch for the Gram matrix and every squaring when the whole grid can be resident (and the stream is
  // not being captured into a graph: a refused launch would invalidate the capture); else one launch per step.
  bool fused = false;
  {
    static int coop_dev[64] = {};   // 0 unknown, 1 usable, -1 not
    static int blocks_per_sm_dev[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    int& coop = coop_dev[dev & 63];
    if (coop == 0) {
      int supported = 0;
      CUDA_TRY(cudaDeviceGetAttribute(&supported, cudaDevAttrCooperativeLaunch, dev));
      const char* e = getenv("VTC_B200_LIPSCHITZ_FUSED");
      coop = (supported && !(e && atoi(e) == 0)) ? 1 : -1;
      if (coop == 1)
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm_dev[dev & 63], lipschitz_squarings_kernel,
                                                               256, 0));
    }
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(st, &capturing));
    DeviceInfo info;
    TRY(device_info(&info));
    if (coop == 1 && capturing == cudaStreamCaptureStatusNone &&
        static_cast<long long>(grid.x) * grid.y <= static_cast<long long>(blocks_per_sm_dev[dev & 63]) * info.sm_count) {
      int n = w.n, squarings = kSquarings;
      int64_t S_ = S, D_ = D;
      double *M = w.M, *A0 = w.A0, *A1 = w.A1, *traces = w.traces;
      void* args[] = {&dict, &S_, &D_, &n, &M, &A0, &A1, &traces, &squarings};
      if (cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lipschitz_squarings_kernel), grid, block, args, 0, st) ==
          cudaSuccess) {
        COUNT_LAUNCH();
        fused = true;
      } else {
        (void)cudaGetLastError();   // refused (e.g. the device is shared): the per-step launches below
        CUDA_TRY(cudaMemsetAsync(w.traces, 0, (kSquarings + 2) * sizeof(double), st));
      }
    }
  }
  if (!fused) {
    gram_fp64_kernel<<<grid, block, 0, st>>>(dict, S, D, w.n, w.M, w.traces + 0);
    COUNT_LAUNCH();
    const double* src = w.M;
    double* dst = w.A0;
    for (int j = 0; j < kSquarings; ++j) {
      square_fp64_kernel<<<grid, block, 0, st>>>(src, w.n, w.traces + j, dst, w.traces + j + 1, j,
                                                 w.traces + kSquarings + 1);
      COUNT_LAUNCH();
      src = dst;
      dst = (dst == w.A0) ? w.A1 : w.A0;
    }
  }
  lipschitz_finalize_kernel<<<1, 1024, 0, st>>>(w.A0, w.A1, w.M, w.n, w.traces, kSquarings, sparsity_weight,
                                                       scalars, lipschitz_dev);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- FISTA
constexpr int kMaxFusedIters = 16384;   // iterations one persistent launch can run (size of the momentum table)
constexpr int kStatSlots = 4096;        // ring of early-stopping statistics (one double per iteration)
struct FistaWs {
  float* scalars;
  double* stats;
  float* betas;   // device: momentum coefficient of every iteration (fused schedule)
  int* done;      // device: per-panel completion counters of the persistent launch
  LipschitzWs lip;
  PartsMat phi_op, x_op, G_op, yop[2], phiT_op, r_op;
  // bvec, X1, X2: tile-contiguous fp32 state [ceil(S/16)][B][16]; init_pad / out_pad / x_pad: row-major staging for
  // caller buffers that TMA cannot address directly (unaligned base or pitch)
  float *bvec, *X1, *X2, *init_pad, *out_pad, *x_pad;
  float *Q0, *Q1, *Q2;   // quad-blocked state of the second-generation iteration kernel (Q0: converted warm start)
  int64_t q_row_blocks, q_col_blocks;
  int64_t ldS, ldD;
};
bool iter2_ok(int64_t S, int64_t D, int precision);
FistaWs carve_fista(Carver& cv, int64_t B, int64_t S, int64_t D, int precision) {
  FistaWs w;
  const int P = parts_for(precision);
  const int bk = (P == 1) ? 64 : 32;  // K block of the iteration GEMMs = block width of the streamed A operands
  w.scalars = static_cast<float*>(cv.take(64));
  w.stats = static_cast<double*>(cv.take(8 * kStatSlots));
  w.betas = static_cast<float*>(cv.take(sizeof(float) * (kMaxFusedIters + 1)));
  w.done = static_cast<int*>(cv.take(sizeof(int) * ceil_div(B, PAIR_M)));
  w.lip = carve_lipschitz(cv, D);
  w.phi_op = carve_parts(cv, S, D, 3);
  const bool gram = formulation_for(S, D) == FORM_GRAM;
  w.ldS = round_up(S, 4);
  w.ldD = round_up(D, 4);
  const size_t state_rm = static_cast<size_t>(B) * w.ldS * 4;                 // row-major, padded pitch
  const size_t state_blk = static_cast<size_t>(B) * round_up(S, EPI_COLS) * 4;  // tile-contiguous
  if (gram) {
    w.x_op = carve_parts(cv, B, D, 3);
    w.G_op = carve_parts(cv, S, S, P);
    w.bvec = static_cast<float*>(cv.take(state_blk));
    w.x_pad = nullptr;
  } else {
    w.phiT_op = carve_parts(cv, D, S, 3);
    w.r_op = carve_parts(cv, B, D, P, bk);
    w.x_pad = static_cast<float*>(cv.take(static_cast<size_t>(B) * w.ldD * 4));
    w.bvec = nullptr;
  }
  w.Q0 = w.Q1 = w.Q2 = nullptr;
  w.q_row_blocks = 2 * ceil_div(B, PAIR_M);
  w.q_col_blocks = round_up(S, 32) / EPI_COLS;
  if (iter2_ok(S, D, precision)) {
    // second-generation kernel: three quad-blocked state arrays (whole 128-row x 32-atom blocks) and, for a warm start
    // only, the operand parts of the starting point (r_0 = y_0 Phi - x by one GEMM)
    const size_t qbytes = static_cast<size_t>(w.q_row_blocks) * w.q_col_blocks * EPI_ARRAY_BYTES;
    w.yop[0] = carve_parts(cv, B, S, P, bk);
    w.yop[1] = PartsMat();
    w.Q0 = static_cast<float*>(cv.take(qbytes));
    w.Q1 = static_cast<float*>(cv.take(qbytes));
    w.Q2 = static_cast<float*>(cv.take(qbytes));
    w.X1 = w.X2 = w.init_pad = w.out_pad = nullptr;
    return w;
  }
  w.yop[0] = carve_parts(cv, B, S, P, bk);
  w.yop[1] = carve_parts(cv, B, S, P, bk);
  w.X1 = static_cast<float*>(cv.take(state_blk));
  w.X2 = static_cast<float*>(cv.take(state_blk));
  w.init_pad = static_cast<float*>(cv.take(state_rm));
  w.out_pad = static_cast<float*>(cv.take(state_rm));
  return w;
}

bool tma_ok(const void* p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 4) == 0; }
bool iter2_ok(int64_t S, int64_t D, int precision) { return fused_iter_ok(S, D, precision) && iter_generation() == 2; }

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int vtc_version(void) { return 100; }
long long vtc_launch_count(void) { return g_launches; }
int vtc_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.valid = false;
  if (on) g_prof.calls = 0;
  return VTC_OK;
}
int vtc_set_formulation(int formulation) {
  if (formulation < 0 || formulation > 2) return fail(VTC_ERR_ARG, "formulation must be 0 (auto), 1 (gram) or 2 (synthesis)");
  g_formulation = formulation;
  return VTC_OK;
}
int vtc_get_formulation(int64_t S, int64_t D) { return formulation_for(S, D); }
int vtc_set_fused_iteration(int on) {
  g_fused_iter = on != 0;
  return VTC_OK;
}
int vtc_set_small_batch_kernel(int on) {
  g_small_kernel = on != 0;
  return VTC_OK;
}
int vtc_debug_iter_trace(void* device_buffer) {
#ifdef VTC_TRACE
  g_iter_trace = static_cast<unsigned long long*>(device_buffer);
  return VTC_OK;
#else
  (void)device_buffer;
  return fail(VTC_ERR_ARG, "vtc_debug_iter_trace: this library was built without -DVTC_TRACE (tools/ab_build.sh)");
#endif
}
int vtc_get_fused_iteration(int64_t S, int64_t D, int precision) {
  return valid_precision(precision) && fused_iter_ok(S, D, precision) ? 1 : 0;
}
int vtc_profile_last(float* setup_ms, float* iter_ms, int* iter_launches, int* iters, float* fused_launch_ms,
                     float* first_launch_ms) {
  if (!g_prof.valid) return fail(VTC_ERR_ARG, "vtc_profile_last: no profiled vtc_fista_fc call");
  CUDA_TRY(cudaEventSynchronize(g_prof.iter_end));
  float a = 0.f, b = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&a, g_prof.begin, g_prof.iter_begin));
  CUDA_TRY(cudaEventElapsedTime(&b, g_prof.iter_begin, g_prof.iter_end));
  if (setup_ms) *setup_ms = a;
  if (iter_ms) *iter_ms = b;
  if (iter_launches) *iter_launches = g_prof.iter_launches;
  if (iters) *iters = g_prof.iters;
  // mean duration of the sampled launches: the fused ISTA/FISTA launch, and (synthesis form) the launch before it
  float fused = 0.f, first = 0.f;
  for (int i = 0; i < g_prof.samples; ++i) {
    float t1 = 0.f, t2 = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t1, g_prof.k1_begin[i], g_prof.k1_end[i]));
    CUDA_TRY(cudaEventElapsedTime(&t2, g_prof.k1_end[i], g_prof.k2_end[i]));
    if (g_prof.two_launches) first += t1, fused += t2;
    else fused += t1;
  }
  if (g_prof.samples > 0) fused /= g_prof.samples, first /= g_prof.samples;
  fused /= g_prof.launch_iters;   // per iteration
  if (fused_launch_ms) *fused_launch_ms = fused;
  if (first_launch_ms) *first_launch_ms = first;
  return VTC_OK;
}
// Every profiled vtc_fista_fc call since vtc_profile_enable(1), oldest first (a ring of kProfHistory calls): device
// time of the setup, of the iteration launches and of the finishing copy, and the idle time of the stream between the
// end of this call and the beginning of the next one (0 for the last).
int vtc_profile_history_count(void) {
  return static_cast<int>(g_prof.calls < kProfHistory ? g_prof.calls : kProfHistory);
}
int vtc_profile_history(int index, float* setup_ms, float* iter_ms, float* finish_ms, float* gap_to_next_ms) {
  const int n = vtc_profile_history_count();
  if (index < 0 || index >= n) return fail(VTC_ERR_ARG, "vtc_profile_history: index %d outside [0, %d)", index, n);
  const long long first = g_prof.calls - n;
  const int slot = static_cast<int>((first + index) % kProfHistory);
  CUDA_TRY(cudaEventSynchronize(g_prof.h_end[slot]));
  float a = 0.f, b = 0.f, c = 0.f, g = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&a, g_prof.h_begin[slot], g_prof.h_iter_begin[slot]));
  CUDA_TRY(cudaEventElapsedTime(&b, g_prof.h_iter_begin[slot], g_prof.h_iter_end[slot]));
  CUDA_TRY(cudaEventElapsedTime(&c, g_prof.h_iter_end[slot], g_prof.h_end[slot]));
  if (index + 1 < n) {
    const int next = static_cast<int>((first + index + 1) % kProfHistory);
    CUDA_TRY(cudaEventSynchronize(g_prof.h_begin[next]));
    CUDA_TRY(cudaEventElapsedTime(&g, g_prof.h_end[slot], g_prof.h_begin[next]));
  }
  if (setup_ms) *setup_ms = a;
  if (iter_ms) *iter_ms = b;
  if (finish_ms) *finish_ms = c;
  if (gap_to_next_ms) *gap_to_next_ms = g;
  return VTC_OK;
}
const char* vtc_last_error(void) { return g_err; }

int vtc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  DeviceInfo info;
  TRY(device_info(&info));
  if (sm_count) *sm_count = info.sm_count;
  if (cc_major) *cc_major = info.cc_major;
  if (cc_minor) *cc_minor = info.cc_minor;
  return VTC_OK;
}

size_t vtc_lipschitz_workspace_bytes(int64_t S, int64_t D) {
  (void)S;
  Carver cv(nullptr, 0);
  carve_lipschitz(cv, D);
  return cv.off + 1024;
}
int vtc_lipschitz(const float* dictionary, int64_t S, int64_t D, float* lipschitz_dev, void* workspace,
                  size_t workspace_bytes, vtc_stream_t stream) {
  if (!dictionary || !lipschitz_dev || S <= 0 || D <= 0) return fail(VTC_ERR_ARG, "vtc_lipschitz: bad argument");
  Carver cv(workspace, workspace_bytes);
  LipschitzWs w = carve_lipschitz(cv, D);
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_lipschitz: workspace too small");
  return run_lipschitz(dictionary, S, D, w, 0.f, nullptr, lipschitz_dev, static_cast<cudaStream_t>(stream));
}

// ---- chains -----------------------------------------------------------------------------------------------------
// In the synthesis form an iteration is a tensor-bound launch (r = y Phi - x) followed by an HBM-bound one
// (r Phi^T + fused update). Rows are independent, so a large batch is cut into two "chains" (contiguous halves) that
// run on two streams with half of the SM pairs each: while one chain is in its tensor-bound launch the other streams
// its state through HBM. Launches are issued alternately so that neither chain starts late.
}  // extern "C"  (helpers below have C++ linkage)

namespace {

struct FistaCommon {
  const float* dictionary;
  int64_t S, D;
  float sparsity_weight;
  int num_iters, variant, prox, group_size, precision, P;
  bool gram, early;
  bool fused_iter;  // one launch per iteration (fista_iter_kernel.cuh) instead of two
  bool iter2;       // ... by the second-generation kernel (fista_iter2_kernel.cuh, quad-blocked state)
  bool small;       // the whole run in one launch with everything on chip (fista_small_kernel.cuh: S <= 256, small batch)
};

struct FistaChain {
  // caller's buffers for this chain's rows
  const float* images;
  const float* initial_codes;
  float* codes_out;
  int64_t ld_images, ld_codes, B;
  cudaStream_t st;
  int max_pairs;       // SM pairs this chain's GEMM launches may occupy (0 = all)
  FistaWs w;
  // state: a_k lives in the tile-contiguous buffer X1 for odd k and X2 for even k; a_0 is `init` (row-major caller
  // data, or X2 zeroed); the last iterate is written row-major to `final_out`
  F32Mat X1, X2, init, final_out;
  const float* x_in;
  int64_t ld_x;
};

int chains_for(int64_t B, int64_t S, int64_t D) {
  static int requested = -1;
  if (requested < 0) {
    const char* e = getenv("VTC_B200_CHAINS");
    requested = e ? atoi(e) : 1;  // measured neutral on configs[1] (the fused launch is bound per SM): off by default
    if (requested < 1) requested = 1;
    if (requested > 2) requested = 2;
  }
  // only where the two launches of an iteration stress different resources, and each chain still fills its SMs
  if (requested == 2 && formulation_for(S, D) == FORM_SYNTHESIS && B >= 2 * 37 * PAIR_M) return 2;
  return 1;
}
int64_t chain_rows(int64_t B, int chains, int c) {
  if (chains == 1) return B;
  const int64_t first = round_up((B + 1) / 2, PAIR_M);
  return c == 0 ? first : B - first;
}

int chain_setup(const FistaCommon& cm, FistaChain& ch) {
  FistaWs& w = ch.w;
  cudaStream_t st = ch.st;
  const int64_t B = ch.B, S = cm.S, D = cm.D;
  // step size and operand splits; Gram form additionally G = Phi Phi^T (as bf16 parts) and b = x Phi^T
  TRY(run_lipschitz(cm.dictionary, S, D, w.lip, cm.sparsity_weight, w.scalars, nullptr, st));
  TRY(split_rows(cm.dictionary, D, S, D, w.phi_op, st));
  ch.x_in = ch.images;  // synthesis form: the fp32 images are an epilogue input of the first contraction
  ch.ld_x = ch.ld_images;
  if (cm.gram) {
    TRY(split_rows(ch.images, ch.ld_images, B, D, w.x_op, st));
    {
      GemmCall g;
      g.A = w.phi_op, g.B = w.phi_op;
      g.precision = VTC_PRECISION_BF16X6;
      g.M = S, g.N = S, g.K = D;
      g.parts_out = w.G_op, g.n_parts = cm.P;
      g.max_pairs = ch.max_pairs;
      CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.G_op.ptr), 0, w.G_op.bytes(), st));
      TRY(launch_gemm<EPI_STORE>(g, st));
    }
    {
      GemmCall g;
      g.A = w.x_op, g.B = w.phi_op;
      g.precision = VTC_PRECISION_BF16X6;
      g.M = B, g.N = S, g.K = D;
      g.out = F32Mat{w.bvec, B, S, 0, true}, g.store_out = true;
      g.max_pairs = ch.max_pairs;
      TRY(launch_gemm<EPI_STORE>(g, st));
    }
    if (cm.small) return VTC_OK;   // the resident kernel needs G, b and the step size only: no state buffers
  } else {
    TRY(transpose_split(cm.dictionary, D, S, D, w.phiT_op, st));
    if (!(cm.fused_iter && !ch.initial_codes))   // (from zero the fused schedule writes every element of r_0 itself)
      CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.r_op.ptr), 0, w.r_op.bytes(), st));
    if (!tma_ok(ch.images, ch.ld_images)) {
      CUDA_TRY(cudaMemcpy2DAsync(w.x_pad, w.ldD * 4, ch.images, ch.ld_images * 4, D * 4, B, cudaMemcpyDeviceToDevice, st));
      ch.x_in = w.x_pad, ch.ld_x = w.ldD;
    }
  }
  if (cm.iter2) {
    // second-generation panel-resident kernel: quad-blocked state. From zero nothing is initialised at all (iteration
    // 1 reads no state, every block is written before it is read) and r_0 = -x is a split of the images; a warm
    // start is converted once and r_0 = y_0 Phi - x comes from one GEMM.
    if (ch.initial_codes) {
      DeviceInfo info;
      TRY(device_info(&info));
      const int64_t slots = w.q_row_blocks * w.q_col_blocks * 512;
      block_quad_kernel<<<grid_for(slots, 256, info.sm_count), 256, 0, st>>>(ch.initial_codes, ch.ld_codes, B, S,
                                                                             w.q_row_blocks, w.q_col_blocks, w.Q0);
      COUNT_LAUNCH();
      CUDA_TRY(cudaGetLastError());
      TRY(split_rows(ch.initial_codes, ch.ld_codes, B, S, w.yop[0], st));
      GemmCall r;
      r.precision = cm.precision;
      r.max_pairs = ch.max_pairs;
      r.A = w.yop[0], r.B = w.phiT_op;
      r.M = B, r.N = D, r.K = S;
      r.in[0] = F32Mat{ch.x_in, B, D, ch.ld_x}, r.in_mask = 1;
      r.parts_out = w.r_op, r.n_parts = cm.P;
      TRY(launch_gemm<EPI_STORE>(r, st));
    } else {
      TRY(split_rows(ch.x_in, ch.ld_x, B, D, w.r_op, st, -1.f));
    }
    if (cm.early) CUDA_TRY(cudaMemsetAsync(w.stats, 0, sizeof(double) * kStatSlots, st));
    return VTC_OK;
  }
  // state buffers
  ch.X1 = F32Mat{w.X1, B, S, 0, true};
  ch.X2 = F32Mat{w.X2, B, S, 0, true};
  ch.final_out = tma_ok(ch.codes_out, ch.ld_codes) ? F32Mat{ch.codes_out, B, S, ch.ld_codes, false}
                                                    : F32Mat{w.out_pad, B, S, w.ldS, false};
  if (!cm.fused_iter) CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.yop[1].ptr), 0, w.yop[1].bytes(), st));
  if (ch.initial_codes) {
    if (tma_ok(ch.initial_codes, ch.ld_codes)) {
      ch.init = F32Mat{ch.initial_codes, B, S, ch.ld_codes, false};
    } else {
      CUDA_TRY(cudaMemcpy2DAsync(w.init_pad, w.ldS * 4, ch.initial_codes, ch.ld_codes * 4, S * 4, B, cudaMemcpyDeviceToDevice, st));
      ch.init = F32Mat{w.init_pad, B, S, w.ldS, false};
    }
    TRY(split_rows(ch.initial_codes, ch.ld_codes, B, S, w.yop[0], st));
  } else {
    // a_0 = 0: X2 doubles as a_0 (it is only overwritten, in place, when a_2 is produced)
    CUDA_TRY(cudaMemsetAsync(w.X2, 0, static_cast<size_t>(B) * round_up(S, EPI_COLS) * 4, st));
    // (the fused schedule starts from r_0 = -x and never reads the operand parts of y_0 = 0)
    if (!cm.fused_iter) CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.yop[0].ptr), 0, w.yop[0].bytes(), st));
    ch.init = ch.X2;
  }
  if (cm.early) CUDA_TRY(cudaMemsetAsync(w.stats, 0, sizeof(double) * kStatSlots, st));
  if (cm.fused_iter && !ch.initial_codes) {
    // the fused schedule carries r_{k-1} = y_{k-1} Phi - x between launches; from zero r_0 = -x is a split of the
    // images (no K = S contraction of an all-zero operand)
    TRY(split_rows(ch.x_in, ch.ld_x, B, D, w.r_op, st, -1.f));
  } else if (cm.fused_iter) {
    // r_0 from the warm start, once
    GemmCall r;
    r.precision = cm.precision;
    r.max_pairs = ch.max_pairs;
    r.A = w.yop[0], r.B = w.phiT_op;
    r.M = B, r.N = D, r.K = S;
    r.in[0] = F32Mat{ch.x_in, B, D, ch.ld_x}, r.in_mask = 1;
    r.parts_out = w.r_op, r.n_parts = cm.P;
    TRY(launch_gemm<EPI_STORE>(r, st));
  }
  return VTC_OK;
}

// Iterations k_first .. k_first + k_count - 1 of one chain in ONE launch of the panel-resident kernel. k_count > 1: the
// persistent schedule (jobs = (iteration, panel) dealt round-robin to the SM pairs, no launch boundary and no panel
// quantisation); the momentum table w.betas must have been uploaded, w.done is zeroed here.
int chain_iterate_fused(const FistaCommon& cm, FistaChain& ch, int k_first, int k_count, int betas_k_lo = 0) {
  FistaWs& w = ch.w;
  IterCall c;
  c.r_op = w.r_op, c.phi_op = w.phi_op, c.phiT_op = w.phiT_op;
  c.precision = cm.precision;
  c.B = ch.B, c.S = cm.S, c.D = cm.D;
  c.state[0] = ch.init, c.state[1] = ch.X1, c.state[2] = ch.X2, c.state[3] = ch.final_out;
  c.x = F32Mat{ch.x_in, ch.B, cm.D, ch.ld_x};
  c.k_first = k_first, c.k_count = k_count;
  c.k_final = cm.early ? 0x7fffffff : cm.num_iters;
  c.betas = w.betas - betas_k_lo;   // the table holds entries betas_k_lo ..: indexed by absolute iteration
  if (k_count > 1) {
    CUDA_TRY(cudaMemsetAsync(w.done, 0, sizeof(int) * ceil_div(ch.B, PAIR_M), ch.st));
    c.done = w.done;
  }
  c.prox = cm.prox, c.group = cm.group_size, c.use_momentum = (cm.variant == VTC_VARIANT_FISTA);
  c.scalars = w.scalars;
  c.stat = cm.early ? w.stats + (k_first - 1) % kStatSlots : nullptr;
  c.max_pairs = ch.max_pairs;
  if (cm.iter2) {
    c.qstate[0] = w.Q0, c.qstate[1] = w.Q1, c.qstate[2] = w.Q2;
    c.init_zero = ch.initial_codes == nullptr;
  }
  return launch_iter(c, ch.st);
}

// iteration k (1-based) of one chain; `sample` >= 0 records profile events around its launches
int chain_iterate(const FistaCommon& cm, FistaChain& ch, int k, float beta_prev, float beta_k, int sample,
                  int betas_k_lo = 0) {
  FistaWs& w = ch.w;
  cudaStream_t st = ch.st;
  const int64_t B = ch.B, S = cm.S, D = cm.D;
  const F32Mat& a_prev = (k == 1) ? ch.init : ((k - 1) & 1) ? ch.X1 : ch.X2;  // a_{k-1}
  const F32Mat& a_prev2 = (k <= 2) ? ch.init : (k & 1) ? ch.X1 : ch.X2;       // a_{k-2}
  // a_k overwrites a_{k-2} in place (same tile reads before it writes); the last iterate of a run without early
  // stopping goes straight to the caller's row-major buffer
  const F32Mat& a_out = (k == cm.num_iters && !cm.early) ? ch.final_out : (k & 1) ? ch.X1 : ch.X2;
  if (sample >= 0) CUDA_TRY(cudaEventRecord(g_prof.k1_begin[sample], st));
  if (cm.fused_iter) {
    (void)a_prev, (void)a_prev2, (void)a_out, (void)beta_prev, (void)beta_k;
    TRY(chain_iterate_fused(cm, ch, k, 1, betas_k_lo));
    if (sample >= 0) {
      CUDA_TRY(cudaEventRecord(g_prof.k1_end[sample], st));
      CUDA_TRY(cudaEventRecord(g_prof.k2_end[sample], st));
    }
    return VTC_OK;
  }
  GemmCall g;
  g.precision = cm.precision;
  g.max_pairs = ch.max_pairs;
  g.in[0] = a_prev;
  g.in_mask = 1;
  if (cm.gram) {
    g.A = w.yop[(k - 1) & 1], g.B = w.G_op;
    g.M = B, g.N = S, g.K = S;
    g.in[1] = F32Mat{w.bvec, B, S, 0, true};
    g.in_mask |= 2;
  } else {
    // r = y Phi - x, emitted as bf16 parts; then the fused contraction acc = r Phi^T is the whole gradient
    GemmCall r;
    r.precision = cm.precision;
    r.max_pairs = ch.max_pairs;
    r.A = w.yop[(k - 1) & 1], r.B = w.phiT_op;
    r.M = B, r.N = D, r.K = S;
    r.in[0] = F32Mat{ch.x_in, B, D, ch.ld_x}, r.in_mask = 1;
    r.parts_out = w.r_op, r.n_parts = cm.P;
    TRY(launch_gemm<EPI_STORE>(r, st));
    if (sample >= 0) CUDA_TRY(cudaEventRecord(g_prof.k1_end[sample], st));
    g.A = w.r_op, g.B = w.phi_op;
    g.M = B, g.N = S, g.K = D;
  }
  if (cm.variant == VTC_VARIANT_FISTA && beta_prev != 0.f) {
    g.in[2] = a_prev2;
    g.in_mask |= 4;
  }
  g.out = a_out, g.store_out = true;
  if (k < cm.num_iters) g.parts_out = w.yop[k & 1], g.n_parts = cm.P;
  g.prox = cm.prox;
  g.group = cm.group_size;
  g.use_momentum = (cm.variant == VTC_VARIANT_FISTA);
  g.beta_prev = beta_prev, g.beta_next = beta_k;
  g.scalars = w.scalars;
  g.stat = cm.early ? w.stats + (k - 1) % kStatSlots : nullptr;
  TRY(launch_gemm<EPI_FISTA>(g, st));
  if (sample >= 0) {
    if (cm.gram) CUDA_TRY(cudaEventRecord(g_prof.k1_end[sample], st));
    CUDA_TRY(cudaEventRecord(g_prof.k2_end[sample], st));
  }
  return VTC_OK;
}

int chain_finish(const FistaCommon& cm, FistaChain& ch, int k_done) {
  const int64_t B = ch.B, S = cm.S;
  if (cm.iter2) {
    // the last iterate sits quad-blocked in the buffer of its parity: one pass back to the caller's row-major codes
    DeviceInfo info;
    TRY(device_info(&info));
    unblock_quad_kernel<<<grid_for(ch.w.q_row_blocks * ((S + 15) / 16) * 512, 256, info.sm_count), 256, 0, ch.st>>>(
        (k_done & 1) ? ch.w.Q1 : ch.w.Q2, B, S, ch.w.q_row_blocks, ch.codes_out, ch.ld_codes);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    return VTC_OK;
  }
  if (cm.early) {
    // the stopping iteration was not known in advance: the result sits in a tile-contiguous state buffer
    DeviceInfo info;
    TRY(device_info(&info));
    const F32Mat& res = (k_done & 1) ? ch.X1 : ch.X2;
    unblock_f32_kernel<<<grid_for(B * S, 256, info.sm_count), 256, 0, ch.st>>>(
        static_cast<const float*>(res.ptr), B, S, ch.codes_out, ch.ld_codes);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
  } else if (ch.final_out.ptr != ch.codes_out) {
    CUDA_TRY(cudaMemcpy2DAsync(ch.codes_out, ch.ld_codes * 4, ch.final_out.ptr, ch.final_out.ld * 4, S * 4, B,
                               cudaMemcpyDeviceToDevice, ch.st));
  }
  return VTC_OK;
}

// side stream + fork/join events of the second chain, one set per device
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
int side_stream(SideStream** out) {
  static SideStream per_device[64];
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  SideStream& s = per_device[dev & 63];
  if (!s.stream) {
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  *out = &s;
  return VTC_OK;
}

}  // namespace

namespace {
// ---- subspace groups wider than one epilogue sub-tile (more than 16 atoms; subspace_ista_fista.py:94-96 takes any
// width): the group norm spans several sub-tiles, so the update is a pass of its own (wide_group_prox_kernel) after the
// two contractions of the synthesis form -- three launches per iteration, any S and D. Rare in practice (the
// reference's examples use groups of 2 to 8), kept simple.
struct WideWs {
  float* scalars;
  double* stats;
  LipschitzWs lip;
  PartsMat phi_op, phiT_op, r_op, yop[2];
  float *x_pad, *acc, *A0, *A1;
  int64_t ldS, ldD;
};
WideWs carve_wide(Carver& cv, int64_t B, int64_t S, int64_t D, int precision) {
  WideWs w;
  const int P = parts_for(precision);
  const int bk = (P == 1) ? 64 : 32;
  w.scalars = static_cast<float*>(cv.take(64));
  w.stats = static_cast<double*>(cv.take(8 * kStatSlots));
  w.lip = carve_lipschitz(cv, D);
  w.phi_op = carve_parts(cv, S, D, 3);
  w.phiT_op = carve_parts(cv, D, S, 3);
  w.r_op = carve_parts(cv, B, D, P, bk);
  w.yop[0] = carve_parts(cv, B, S, P, bk);
  w.yop[1] = carve_parts(cv, B, S, P, bk);
  w.ldS = round_up(S, 4), w.ldD = round_up(D, 4);
  w.x_pad = static_cast<float*>(cv.take(static_cast<size_t>(B) * w.ldD * 4));
  const size_t state = static_cast<size_t>(B) * w.ldS * 4;
  w.acc = static_cast<float*>(cv.take(state));
  w.A0 = static_cast<float*>(cv.take(state));
  w.A1 = static_cast<float*>(cv.take(state));
  return w;
}
bool wide_groups(int group_size) { return group_size > EPI_COLS; }

int run_wide_groups(const float* images, int64_t ld_images, const float* dictionary, const float* initial_codes,
                    float* codes_out, int64_t ld_codes, int64_t B, int64_t S, int64_t D, float sparsity_weight,
                    int num_iters, int variant, int W, float early_stopping_epsilon, int precision, void* workspace,
                    size_t workspace_bytes, int* iters_run, float* lipschitz_out, cudaStream_t st) {
  if (W % 32 != 0 || S % W != 0)
    return fail(VTC_ERR_UNSUPPORTED, "wide groups: the (padded) group width must be a multiple of 32 dividing the grouped code size; got %d", W);
  Carver cv(workspace, workspace_bytes);
  WideWs w = carve_wide(cv, B, S, D, precision);
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_fista_fc: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
  DeviceInfo info;
  TRY(require_sm100(&info));
  const int P = parts_for(precision);
  const bool early = early_stopping_epsilon >= 0.f;
  const bool fista = variant == VTC_VARIANT_FISTA;
  TRY(run_lipschitz(dictionary, S, D, w.lip, sparsity_weight, w.scalars, nullptr, st));
  TRY(split_rows(dictionary, D, S, D, w.phi_op, st));
  TRY(transpose_split(dictionary, D, S, D, w.phiT_op, st));
  const float* x_in = images;
  int64_t ld_x = ld_images;
  if (!tma_ok(images, ld_images)) {
    CUDA_TRY(cudaMemcpy2DAsync(w.x_pad, w.ldD * 4, images, ld_images * 4, D * 4, B, cudaMemcpyDeviceToDevice, st));
    x_in = w.x_pad, ld_x = w.ldD;
  }
  const size_t state = static_cast<size_t>(B) * w.ldS * 4;
  CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.r_op.ptr), 0, w.r_op.bytes(), st));
  CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.yop[1].ptr), 0, w.yop[1].bytes(), st));
  CUDA_TRY(cudaMemsetAsync(w.A1, 0, state, st));
  if (initial_codes) {
    CUDA_TRY(cudaMemsetAsync(w.A0, 0, state, st));
    CUDA_TRY(cudaMemcpy2DAsync(w.A0, w.ldS * 4, initial_codes, ld_codes * 4, S * 4, B, cudaMemcpyDeviceToDevice, st));
    TRY(split_rows(initial_codes, ld_codes, B, S, w.yop[0], st));
  } else {
    CUDA_TRY(cudaMemsetAsync(w.A0, 0, state, st));
    CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.yop[0].ptr), 0, w.yop[0].bytes(), st));
  }
  if (early) CUDA_TRY(cudaMemsetAsync(w.stats, 0, sizeof(double) * kStatSlots, st));
  float eta_host = 0.f;
  if (early || lipschitz_out) {
    float sc_host[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(sc_host, w.scalars, sizeof(sc_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    eta_host = sc_host[0];
    if (lipschitz_out) *lipschitz_out = sc_host[2];
    if (sc_host[3] != 0.f || !isfinite(sc_host[2]))
      return fail(VTC_ERR_NONFINITE, "largest eigenvalue of dictionary^T dictionary is %g: a dictionary element overflowed", sc_host[2]);
  }
  double t_k = 1.0;
  float beta_prev = 0.f;
  int k_done = 0;
  for (int k = 1; k <= num_iters; ++k) {
    const double t_next = (1.0 + sqrt(1.0 + 4.0 * t_k * t_k)) / 2.0;
    const float beta_k = fista ? static_cast<float>((t_k - 1.0) / t_next) : 0.f;
    t_k = t_next;
    float* prev = (k & 1) ? w.A0 : w.A1;    // a_{k-1}
    float* other = (k & 1) ? w.A1 : w.A0;   // a_{k-2}, overwritten by a_k
    GemmCall r;   // r = y_{k-1} Phi - x, as bf16 parts
    r.precision = precision;
    r.A = w.yop[(k - 1) & 1], r.B = w.phiT_op;
    r.M = B, r.N = D, r.K = S;
    r.in[0] = F32Mat{x_in, B, D, ld_x}, r.in_mask = 1;
    r.parts_out = w.r_op, r.n_parts = P;
    TRY(launch_gemm<EPI_STORE>(r, st));
    GemmCall g;   // gradient = r Phi^T, fp32
    g.precision = precision;
    g.A = w.r_op, g.B = w.phi_op;
    g.M = B, g.N = S, g.K = D;
    g.out = F32Mat{w.acc, B, S, w.ldS, false}, g.store_out = true;
    TRY(launch_gemm<EPI_STORE>(g, st));
    if (early && k > 1 && (k - 1) % kStatSlots == 0) CUDA_TRY(cudaMemsetAsync(w.stats, 0, sizeof(double) * kStatSlots, st));
    const PartsMat& yo = w.yop[k & 1];
    wide_group_prox_kernel<<<grid_for(B * (S / W) * 32, 256, info.sm_count), 256, 0, st>>>(
        w.acc, prev, other, w.ldS, B, S, W, w.scalars, beta_prev, beta_k, fista ? 1 : 0,
        (fista && beta_prev != 0.f) ? 1 : 0,
        k < num_iters ? reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(yo.ptr)) : nullptr, yo.Kp, yo.parts, yo.block,
        early ? w.stats + (k - 1) % kStatSlots : nullptr);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    beta_prev = beta_k;
    k_done = k;
    if (early) {
      double sum_abs = 0.0;
      CUDA_TRY(cudaMemcpyAsync(&sum_abs, w.stats + (k - 1) % kStatSlots, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      const double avg = sum_abs / (static_cast<double>(B) * static_cast<double>(S)) / static_cast<double>(eta_host);
      if (avg < static_cast<double>(early_stopping_epsilon) && k > 1) break;
    }
  }
  const float* result = (k_done & 1) ? w.A1 : w.A0;
  CUDA_TRY(cudaMemcpy2DAsync(codes_out, ld_codes * 4, result, w.ldS * 4, S * 4, B, cudaMemcpyDeviceToDevice, st));
  if (iters_run) *iters_run = k_done;
  return VTC_OK;
}
}  // namespace

extern "C" {

size_t vtc_fista_workspace_bytes(int64_t B, int64_t S, int64_t D, int precision) {
  if (!valid_precision(precision) || B <= 0 || S <= 0 || D <= 0) return 0;
  // enough for either schedule: one chain over all rows (always used with early stopping) or two half-batch chains
  Carver one(nullptr, 0);
  carve_fista(one, B, S, D, precision);
  Carver two(nullptr, 0);
  const int chains = chains_for(B, S, D);
  for (int c = 0; c < chains; ++c) carve_fista(two, chain_rows(B, chains, c), S, D, precision);
  Carver wide(nullptr, 0);   // (the group width is not an argument here: room for the wide-group schedule as well)
  carve_wide(wide, B, S, D, precision);
  size_t need = one.off > two.off ? one.off : two.off;
  if (wide.off > need) need = wide.off;
  return need + 2048;
}

int vtc_get_chains(int64_t B, int64_t S, int64_t D) { return chains_for(B, S, D); }

int vtc_fista_fc(const float* images, int64_t ld_images, const float* dictionary, const float* initial_codes,
                 float* codes_out, int64_t ld_codes, int64_t B, int64_t S, int64_t D, float sparsity_weight,
                 int num_iters, int variant, int nonnegative_only, int hard_threshold, int group_size,
                 float early_stopping_epsilon, int precision, void* workspace, size_t workspace_bytes,
                 int* iters_run, float* lipschitz_out, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!images || !dictionary || !codes_out) return fail(VTC_ERR_ARG, "vtc_fista_fc: null pointer");
  if (B <= 0 || S <= 0 || D <= 0 || ld_images < D || ld_codes < S) return fail(VTC_ERR_ARG, "vtc_fista_fc: bad shape");
  if (variant != VTC_VARIANT_ISTA && variant != VTC_VARIANT_FISTA) return fail(VTC_ERR_ARG, "variant must be ista or fista");
  if (!valid_precision(precision)) return fail(VTC_ERR_ARG, "precision must be 1, 3 or 6");
  if (num_iters < 1) return fail(VTC_ERR_ARG, "num_iters must be >= 1");
  if (group_size < 1) return fail(VTC_ERR_ARG, "group_size must be >= 1");
  if (group_size > 1) {
    if (hard_threshold) return fail(VTC_ERR_UNSUPPORTED, "hard threshold is not implemented for the subspace variant");
    if (wide_groups(group_size))   // the norm spans several epilogue sub-tiles: the three-launch schedule
      return run_wide_groups(images, ld_images, dictionary, initial_codes, codes_out, ld_codes, B, S, D, sparsity_weight,
                             num_iters, variant, group_size, early_stopping_epsilon, precision, workspace,
                             workspace_bytes, iters_run, lipschitz_out, st);
    if (EPI_COLS % group_size != 0 || S % group_size != 0)
      return fail(VTC_ERR_UNSUPPORTED, "group_size must divide 16 (or be a multiple of 32) and divide the (grouped) code size; got %d", group_size);
  }
  if (S > (1 << 20) || B > (1ll << 31) - 256) return fail(VTC_ERR_ARG, "problem too large for 32-bit tile coordinates");
  DeviceInfo info;
  TRY(require_sm100(&info));
  FistaCommon cm;
  cm.dictionary = dictionary;
  cm.S = S, cm.D = D;
  cm.sparsity_weight = sparsity_weight;
  cm.num_iters = num_iters, cm.variant = variant;
  cm.prox = (hard_threshold ? PROX_HARD : 0) | (nonnegative_only ? PROX_NONNEG : 0);
  cm.group_size = group_size;
  cm.precision = precision, cm.P = parts_for(precision);
  cm.gram = formulation_for(S, D) == FORM_GRAM;
  cm.small = small_kernel_ok(B, S, D, precision, group_size, early_stopping_epsilon >= 0.f) && num_iters <= kMaxFusedIters;
  cm.iter2 = iter2_ok(S, D, precision);   // (takes runs longer than the momentum table in windows)
  cm.fused_iter = cm.iter2 || (fused_iter_ok(S, D, precision) && num_iters <= kMaxFusedIters);  // else: two launches
  cm.early = early_stopping_epsilon >= 0.f;

  // the early-stopping statistic is global over the batch: one chain then
  const int chains = cm.early ? 1 : chains_for(B, S, D);
  FistaChain ch[2];
  Carver cv(workspace, workspace_bytes);
  SideStream* side = nullptr;
  if (chains == 2) TRY(side_stream(&side));
  int64_t row0 = 0;
  for (int c = 0; c < chains; ++c) {
    const int64_t rows = chain_rows(B, chains, c);
    ch[c].images = images + row0 * ld_images;
    ch[c].initial_codes = initial_codes ? initial_codes + row0 * ld_codes : nullptr;
    ch[c].codes_out = codes_out + row0 * ld_codes;
    ch[c].ld_images = ld_images, ch[c].ld_codes = ld_codes, ch[c].B = rows;
    ch[c].st = (c == 0) ? st : side->stream;
    ch[c].max_pairs = (chains == 2) ? info.sm_count / 4 : 0;
    ch[c].w = carve_fista(cv, rows, S, D, precision);
    row0 += rows;
  }
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_fista_fc: workspace too small (%zu < %zu)", workspace_bytes, cv.off);

  int prof_slot = 0;
  if (g_prof.on) {
    if (!g_prof.events_made) {
      for (int i = 0; i < kProfHistory; ++i) {
        CUDA_TRY(cudaEventCreate(&g_prof.h_begin[i]));
        CUDA_TRY(cudaEventCreate(&g_prof.h_iter_begin[i]));
        CUDA_TRY(cudaEventCreate(&g_prof.h_iter_end[i]));
        CUDA_TRY(cudaEventCreate(&g_prof.h_end[i]));
      }
      g_prof.events_made = true;
      for (int i = 0; i < kProfSamples; ++i) {
        CUDA_TRY(cudaEventCreate(&g_prof.k1_begin[i]));
        CUDA_TRY(cudaEventCreate(&g_prof.k1_end[i]));
        CUDA_TRY(cudaEventCreate(&g_prof.k2_end[i]));
      }
    }
    prof_slot = static_cast<int>(g_prof.calls % kProfHistory);
    g_prof.begin = g_prof.h_begin[prof_slot];
    g_prof.iter_begin = g_prof.h_iter_begin[prof_slot];
    g_prof.iter_end = g_prof.h_iter_end[prof_slot];
    g_prof.samples = 0;
    g_prof.valid = false;
    CUDA_TRY(cudaEventRecord(g_prof.begin, st));
  }
  if (chains == 2) {  // fork: the side stream starts after everything already queued on the caller's stream
    CUDA_TRY(cudaEventRecord(side->fork, st));
    CUDA_TRY(cudaStreamWaitEvent(side->stream, side->fork, 0));
  }
  for (int c = 0; c < chains; ++c) TRY(chain_setup(cm, ch[c]));

  float eta_host = 0.f;
  float sc_host[4] = {0, 0, 0, 0};
  if (cm.early || lipschitz_out) {
    CUDA_TRY(cudaMemcpyAsync(sc_host, ch[0].w.scalars, sizeof(sc_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    eta_host = sc_host[0];
    if (lipschitz_out) *lipschitz_out = sc_host[2];
    if (sc_host[3] != 0.f || !isfinite(sc_host[2])) {
      if (chains == 2) cudaStreamSynchronize(side->stream);
      return fail(VTC_ERR_NONFINITE, "largest eigenvalue of dictionary^T dictionary is %g: a dictionary element overflowed", sc_host[2]);
    }
  }

  // ---- iterations. beta_k = (t_k - 1) / t_{k+1} in double on the host (ista_fista.py:124-125), applied as float32.
  double t_k = 1.0;
  float beta_prev = 0.f;
  int k_done = 0;
  const int table_iters = num_iters < kMaxFusedIters ? num_iters : kMaxFusedIters;
  if (cm.small) {
    // everything on chip for the whole run: one launch, result straight into the caller's codes
    FistaWs& w = ch[0].w;
    fista_betas_kernel<<<1, 32, 0, st>>>(w.betas, num_iters, variant == VTC_VARIANT_FISTA ? 1 : 0, 0);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
    if (g_prof.on) CUDA_TRY(cudaEventRecord(g_prof.iter_begin, st));
    SmallCall sc;
    sc.G_op = w.G_op, sc.b = w.bvec, sc.init = initial_codes, sc.ld_init = ld_codes, sc.out = codes_out, sc.ld_out = ld_codes;
    sc.B = B, sc.S = S, sc.num_iters = num_iters, sc.prox = cm.prox, sc.use_momentum = (variant == VTC_VARIANT_FISTA);
    sc.precision = precision, sc.betas = w.betas, sc.scalars = w.scalars;
    if (g_prof.on) CUDA_TRY(cudaEventRecord(g_prof.k1_begin[0], st));
    TRY(cm.P == 1 ? launch_small_p<1>(sc, st) : launch_small_p<2>(sc, st));
    if (g_prof.on) {
      CUDA_TRY(cudaEventRecord(g_prof.k1_end[0], st));
      CUDA_TRY(cudaEventRecord(g_prof.k2_end[0], st));
      CUDA_TRY(cudaEventRecord(g_prof.iter_end, st));
      CUDA_TRY(cudaEventRecord(g_prof.h_end[prof_slot], st));
      ++g_prof.calls;
      g_prof.samples = 1, g_prof.two_launches = false;
      g_prof.iter_launches = 1, g_prof.launch_iters = num_iters, g_prof.iters = num_iters;
      g_prof.valid = true;
    }
    if (iters_run) *iters_run = num_iters;
    return VTC_OK;
  }
  if (cm.fused_iter) {
    // the panel-resident kernel takes the momentum coefficients from a device table: betas[k], betas[0] = 0
    for (int c = 0; c < chains; ++c) {
      fista_betas_kernel<<<1, 32, 0, ch[c].st>>>(ch[c].w.betas, table_iters, variant == VTC_VARIANT_FISTA ? 1 : 0, 0);
      COUNT_LAUNCH();
    }
    CUDA_TRY(cudaGetLastError());
  }
  int betas_k_lo = 0;   // first iteration index held by the device table (runs longer than the table: windows)
  auto betas_window = [&](int k) -> int {   // makes the table cover iterations k - 1 .. ; returns 0 or an error
    if (k <= betas_k_lo + table_iters && k - 1 >= betas_k_lo) return VTC_OK;
    betas_k_lo = k - 1;
    const int hi = (num_iters - betas_k_lo < kMaxFusedIters) ? num_iters : betas_k_lo + kMaxFusedIters;
    for (int c = 0; c < chains; ++c) {
      fista_betas_kernel<<<1, 32, 0, ch[c].st>>>(ch[c].w.betas, hi, variant == VTC_VARIANT_FISTA ? 1 : 0, betas_k_lo);
      COUNT_LAUNCH();
    }
    CUDA_TRY(cudaGetLastError());
    return VTC_OK;
  };
  if (g_prof.on) CUDA_TRY(cudaEventRecord(g_prof.iter_begin, st));
  const bool persistent = cm.fused_iter && !cm.early && chains == 1 && persistent_iterations_enabled();
  if (persistent) {
    // every iteration of every panel in ONE launch (runs longer than the momentum table: one launch per window)
    if (g_prof.on) CUDA_TRY(cudaEventRecord(g_prof.k1_begin[0], st));
    int launches = 0;
    for (int k0 = 1; k0 <= num_iters; k0 += kMaxFusedIters, ++launches) {
      TRY(betas_window(k0));
      const int count = (num_iters - k0 + 1 < kMaxFusedIters) ? num_iters - k0 + 1 : kMaxFusedIters;
      TRY(chain_iterate_fused(cm, ch[0], k0, count, betas_k_lo));
    }
    if (g_prof.on) {
      CUDA_TRY(cudaEventRecord(g_prof.k1_end[0], st));
      CUDA_TRY(cudaEventRecord(g_prof.k2_end[0], st));
      g_prof.samples = 1;
      g_prof.two_launches = false;
    }
    k_done = num_iters;
  }
  for (int k = 1; k <= num_iters && !persistent; ++k) {
    const double t_next = (1.0 + sqrt(1.0 + 4.0 * t_k * t_k)) / 2.0;
    const float beta_k = (variant == VTC_VARIANT_FISTA) ? static_cast<float>((t_k - 1.0) / t_next) : 0.f;
    t_k = t_next;
    const int sample = (g_prof.on && k > num_iters / 2 && g_prof.samples < kProfSamples) ? g_prof.samples : -1;
    if (cm.fused_iter) TRY(betas_window(k));
    if (cm.early && k > 1 && (k - 1) % kStatSlots == 0)   // the ring of statistics wraps: every slot has been read
      CUDA_TRY(cudaMemsetAsync(ch[0].w.stats, 0, sizeof(double) * kStatSlots, st));
    for (int c = 0; c < chains; ++c) TRY(chain_iterate(cm, ch[c], k, beta_prev, beta_k, c == 0 ? sample : -1, betas_k_lo));
    if (sample >= 0) {
      g_prof.samples = sample + 1;
      g_prof.two_launches = !cm.gram && !cm.fused_iter;
    }
    beta_prev = beta_k;
    k_done = k;
    if (cm.early) {
      double sum_abs = 0.0;
      CUDA_TRY(cudaMemcpyAsync(&sum_abs, ch[0].w.stats + (k - 1) % kStatSlots, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      const double avg = sum_abs / (static_cast<double>(B) * static_cast<double>(S)) / static_cast<double>(eta_host);
      if (avg < static_cast<double>(early_stopping_epsilon) && k > 1) break;
    }
  }
  if (g_prof.on && chains == 1) CUDA_TRY(cudaEventRecord(g_prof.iter_end, st));
  for (int c = 0; c < chains; ++c) TRY(chain_finish(cm, ch[c], k_done));
  if (chains == 2) {  // join: the caller's stream continues only after the side chain is done
    CUDA_TRY(cudaEventRecord(side->join, side->stream));
    CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));
  }
  if (g_prof.on) {
    if (chains == 2) CUDA_TRY(cudaEventRecord(g_prof.iter_end, st));
    CUDA_TRY(cudaEventRecord(g_prof.h_end[prof_slot], st));
    ++g_prof.calls;
    g_prof.iter_launches = persistent ? 1 : k_done * ((cm.gram || cm.fused_iter) ? 1 : 2) * chains;
    g_prof.launch_iters = persistent ? k_done : 1;
    g_prof.iters = k_done;
    g_prof.valid = true;
  }
  if (iters_run) *iters_run = k_done;
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- dictionary update
namespace {
struct GradWs {
  PartsMat a_op, aT_op, phiT_op, RT_op;
  float* xT;
  int64_t ldB;
  float* partial;
  int ksplits;
  int64_t rows_per_split, ldD;
};
GradWs carve_grad(Carver& cv, int64_t B, int64_t S, int64_t D, int precision, int sm_count) {
  GradWs w;
  const int P = parts_for(precision);
  w.a_op = carve_parts(cv, B, S, P);
  w.aT_op = carve_parts(cv, S, B, P);
  w.phiT_op = carve_parts(cv, D, S, P);
  w.RT_op = carve_parts(cv, D, B, P);
  w.ldB = round_up(B, 4);
  w.xT = static_cast<float*>(cv.take(static_cast<size_t>(D) * w.ldB * 4));
  const int64_t tiles_mn = ceil_div(S, PAIR_M) * ceil_div(D, BLOCK_N);
  const int64_t kb = k_blocks_for(round_up(B, 64), precision);
  int64_t ks = (sm_count / 2) / tiles_mn;
  if (ks < 1) ks = 1;
  if (ks > kb) ks = kb;
  w.ksplits = static_cast<int>(ks);
  w.rows_per_split = ceil_div(S, PAIR_M) * PAIR_M;
  w.ldD = round_up(D, 4);
  w.partial = static_cast<float*>(cv.take(static_cast<size_t>(w.ksplits) * w.rows_per_split * w.ldD * 4));
  return w;
}
}  // namespace

size_t vtc_dict_grad_workspace_bytes(int64_t B, int64_t S, int64_t D, int precision) {
  if (!valid_precision(precision) || B <= 0 || S <= 0 || D <= 0) return 0;
  Carver cv(nullptr, 0);
  carve_grad(cv, B, S, D, precision, 148);
  return cv.off + 1024;
}

int vtc_sc_dict_grad(const float* images, int64_t ld_images, const float* dictionary, const float* codes,
                     int64_t ld_codes, float* grad_sum, int64_t B, int64_t S, int64_t D, int precision,
                     void* workspace, size_t workspace_bytes, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!images || !dictionary || !codes || !grad_sum) return fail(VTC_ERR_ARG, "vtc_sc_dict_grad: null pointer");
  if (B <= 0 || S <= 0 || D <= 0 || ld_images < D || ld_codes < S) return fail(VTC_ERR_ARG, "vtc_sc_dict_grad: bad shape");
  if (!valid_precision(precision)) return fail(VTC_ERR_ARG, "precision must be 1, 3 or 6");
  DeviceInfo info;
  TRY(require_sm100(&info));
  Carver cv(workspace, workspace_bytes);
  // ksplits must not depend on the SM count used for the size query (148 is the B200 count the query assumes)
  GradWs w = carve_grad(cv, B, S, D, precision, info.sm_count < 148 ? info.sm_count : 148);
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_sc_dict_grad: workspace too small (%zu < %zu)", workspace_bytes, cv.off);

  TRY(split_rows(codes, ld_codes, B, S, w.a_op, st));
  TRY(transpose_split(codes, ld_codes, B, S, w.aT_op, st));
  TRY(transpose_split(dictionary, D, S, D, w.phiT_op, st));
  TRY(transpose_f32(images, ld_images, B, D, w.xT, w.ldB, st));
  // R^T (D x B) = Phi^T a^T - x^T, emitted directly as bf16 parts (the K-major operand of the next contraction)
  {
    GemmCall g;
    g.A = w.phiT_op, g.B = w.a_op;
    g.precision = precision;
    g.M = D, g.N = B, g.K = S;
    g.in[0] = F32Mat{w.xT, D, B, w.ldB}, g.in_mask = 1;
    g.parts_out = w.RT_op, g.n_parts = w.RT_op.parts;
    CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.RT_op.ptr), 0, w.RT_op.bytes(), st));
    TRY(launch_gemm<EPI_STORE>(g, st));
  }
  // grad (S x D) = a^T R, contraction over the batch, split-K across the SMs
  {
    GemmCall g;
    g.A = w.aT_op, g.B = w.RT_op;
    g.precision = precision;
    g.M = S, g.N = D, g.K = B;
    g.ksplits = w.ksplits;
    g.out_rows_per_split = w.rows_per_split;
    g.out = F32Mat{w.partial, static_cast<int64_t>(w.ksplits) * w.rows_per_split, D, w.ldD}, g.store_out = true;
    TRY(launch_gemm<EPI_STORE>(g, st));
    // launch_gemm may have lowered the split count; recompute exactly as it does
    const int64_t kb = k_blocks_for(round_up(B, 64), precision);
    const int64_t per = ceil_div(kb, w.ksplits);
    const int nsplit = static_cast<int>(ceil_div(kb, per));
    reduce_partials_kernel<<<grid_for(S * D, 256, info.sm_count), 256, 0, st>>>(w.partial, nsplit, w.rows_per_split,
                                                                                w.ldD, S, D, grad_sum);
    COUNT_LAUNCH();
    CUDA_TRY(cudaGetLastError());
  }
  return VTC_OK;
}

int vtc_sc_dict_apply(float* dictionary, const float* grad_sum, const float* hessian_diagonal,
                      const float* alignment_grad, float alignment_penalty, int64_t S, int64_t D,
                      int64_t batch_global, float stepsize, float lowest_code_val, int normalize,
                      vtc_stream_t stream) {
  if (!dictionary || !grad_sum || S <= 0 || D <= 0 || batch_global <= 0) return fail(VTC_ERR_ARG, "vtc_sc_dict_apply: bad argument");
  dict_apply_kernel<<<static_cast<unsigned>(S), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dictionary, grad_sum, hessian_diagonal, alignment_grad, alignment_penalty, D, static_cast<float>(batch_global),
      stepsize, lowest_code_val, normalize);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_subspace_alignment_grad(const float* dictionary, int64_t S, int64_t D, const int32_t* group_slots,
                                int64_t num_groups, int64_t group_width, int dictionary_is_normalized,
                                float* alignment_grad, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!dictionary || !group_slots || !alignment_grad || S <= 0 || D <= 0 || num_groups <= 0 || group_width <= 0)
    return fail(VTC_ERR_ARG, "vtc_subspace_alignment_grad: bad argument");
  const size_t smem = (static_cast<size_t>(group_width) * D + 2 * group_width) * sizeof(float);
  if (smem > 200 * 1024) return fail(VTC_ERR_UNSUPPORTED, "group of %lld atoms x %lld pixels does not fit in shared memory", (long long)group_width, (long long)D);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(alignment_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  CUDA_TRY(cudaMemsetAsync(alignment_grad, 0, static_cast<size_t>(S) * D * sizeof(float), st));
  alignment_grad_kernel<<<static_cast<unsigned>(S), 256, smem, st>>>(
      dictionary, D, group_slots, static_cast<int>(num_groups), static_cast<int>(group_width), dictionary_is_normalized,
      alignment_grad);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_hessian_diag_update(const float* codes, int64_t ld_codes, int64_t B, int64_t S, int64_t batch_global,
                            float* code_sq_sum, float* hessian_diagonal, int apply_ema, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!codes || !code_sq_sum || B <= 0 || S <= 0 || batch_global <= 0) return fail(VTC_ERR_ARG, "vtc_hessian_diag_update: bad argument");
  if (apply_ema && !hessian_diagonal) return fail(VTC_ERR_ARG, "vtc_hessian_diag_update: hessian_diagonal required");
  DeviceInfo info;
  TRY(device_info(&info));
  CUDA_TRY(cudaMemsetAsync(code_sq_sum, 0, S * sizeof(float), st));
  const int64_t col_blocks = ceil_div(S, 128);
  int64_t row_blocks = (static_cast<int64_t>(info.sm_count) * 8) / col_blocks;
  if (row_blocks < 1) row_blocks = 1;
  if (row_blocks > B) row_blocks = B;
  const int64_t rows_per_block = ceil_div(B, row_blocks);
  row_blocks = ceil_div(B, rows_per_block);
  col_sq_sum_kernel<<<dim3(static_cast<unsigned>(col_blocks), static_cast<unsigned>(row_blocks)), 128, 0, st>>>(
      codes, ld_codes, B, S, rows_per_block, code_sq_sum);
  COUNT_LAUNCH();
  if (apply_ema) {
    hessian_ema_kernel<<<static_cast<unsigned>(ceil_div(S, 256)), 256, 0, st>>>(hessian_diagonal, code_sq_sum, S,
                                                                              static_cast<float>(batch_global));
    COUNT_LAUNCH();
  }
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_hessian_ema(float* hessian_diagonal, const float* code_sq_sum, int64_t S, int64_t batch_global,
                    vtc_stream_t stream) {
  if (!hessian_diagonal || !code_sq_sum || S <= 0 || batch_global <= 0) return fail(VTC_ERR_ARG, "vtc_hessian_ema: bad argument");
  hessian_ema_kernel<<<static_cast<unsigned>(ceil_div(S, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      hessian_diagonal, code_sq_sum, S, static_cast<float>(batch_global));
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- generic matmul
size_t vtc_matmul_nt_workspace_bytes(int64_t M, int64_t N, int64_t K, int precision) {
  if (!valid_precision(precision) || M <= 0 || N <= 0 || K <= 0) return 0;
  Carver cv(nullptr, 0);
  carve_parts(cv, M, K, parts_for(precision));
  carve_parts(cv, N, K, parts_for(precision));
  cv.take(static_cast<size_t>(M) * round_up(N, 4) * 4);
  cv.take(static_cast<size_t>(M) * round_up(N, 4) * 4);
  return cv.off + 1024;
}
int vtc_matmul_nt(const float* A, const float* Bm, const float* sub, float* out, int64_t M, int64_t N, int64_t K,
                  int precision, void* workspace, size_t workspace_bytes, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!A || !Bm || !out || M <= 0 || N <= 0 || K <= 0) return fail(VTC_ERR_ARG, "vtc_matmul_nt: bad argument");
  if (!valid_precision(precision)) return fail(VTC_ERR_ARG, "precision must be 1, 3 or 6");
  Carver cv(workspace, workspace_bytes);
  PartsMat a = carve_parts(cv, M, K, parts_for(precision));
  PartsMat b = carve_parts(cv, N, K, parts_for(precision));
  const int64_t ldN = round_up(N, 4);
  float* out_pad = static_cast<float*>(cv.take(static_cast<size_t>(M) * ldN * 4));
  float* sub_pad = static_cast<float*>(cv.take(static_cast<size_t>(M) * ldN * 4));
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_matmul_nt: workspace too small");
  TRY(split_rows(A, K, M, K, a, st));
  TRY(split_rows(Bm, K, N, K, b, st));
  GemmCall g;
  g.A = a, g.B = b;
  g.precision = precision;
  g.M = M, g.N = N, g.K = K;
  const bool direct = tma_ok(out, N);
  if (sub) {
    const float* s = sub;
    int64_t lds = N;
    if (!tma_ok(sub, N)) {
      CUDA_TRY(cudaMemcpy2DAsync(sub_pad, ldN * 4, sub, N * 4, N * 4, M, cudaMemcpyDeviceToDevice, st));
      s = sub_pad, lds = ldN;
    }
    g.in[0] = F32Mat{s, M, N, lds}, g.in_mask = 1;
  }
  g.out = direct ? F32Mat{out, M, N, N} : F32Mat{out_pad, M, N, ldN};
  g.store_out = true;
  TRY(launch_gemm<EPI_STORE>(g, st));
  if (!direct) CUDA_TRY(cudaMemcpy2DAsync(out, N * 4, out_pad, ldN * 4, N * 4, M, cudaMemcpyDeviceToDevice, st));
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- subspace helpers
int vtc_gather_rows(const float* src, int64_t ld_src, const int32_t* index, int64_t n_slots, int64_t D, float* dst,
                    vtc_stream_t stream) {
  if (!src || !index || !dst || n_slots <= 0 || D <= 0) return fail(VTC_ERR_ARG, "vtc_gather_rows: bad argument");
  DeviceInfo info;
  TRY(device_info(&info));
  gather_rows_kernel<<<grid_for(n_slots * D, 256, info.sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, ld_src, index, n_slots, D, dst);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}
int vtc_gather_cols(const float* src, int64_t ld_src, const int32_t* index, int64_t B, int64_t n_slots, float* dst,
                    int64_t ld_dst, vtc_stream_t stream) {
  if (!src || !index || !dst || n_slots <= 0 || B <= 0) return fail(VTC_ERR_ARG, "vtc_gather_cols: bad argument");
  DeviceInfo info;
  TRY(device_info(&info));
  gather_cols_kernel<<<grid_for(B * n_slots, 256, info.sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, ld_src, index, B, n_slots, dst, ld_dst);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}
int vtc_scatter_add_cols(const float* src, int64_t ld_src, const int32_t* index, int64_t B, int64_t n_slots,
                         float* dst, int64_t ld_dst, int64_t S, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!src || !index || !dst || n_slots <= 0 || B <= 0 || S <= 0) return fail(VTC_ERR_ARG, "vtc_scatter_add_cols: bad argument");
  DeviceInfo info;
  TRY(device_info(&info));
  CUDA_TRY(cudaMemset2DAsync(dst, ld_dst * 4, 0, S * 4, B, st));
  scatter_add_cols_kernel<<<grid_for(B * n_slots, 256, info.sm_count), 256, 0, st>>>(src, ld_src, index, B, n_slots,
                                                                                      dst, ld_dst);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

}  // extern "C"

// ================================================================================================ convolutional path
namespace {

struct ConvShape {
  ConvGeom g;
  int64_t rows;          // b * gh * gw
  int nq;                // kernel taps in units of the stride: ty * tx
  int pad_t, pad_b, pad_l, pad_r;
  int64_t per_kernel;    // c * kh * kw
};

int conv_shape(int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW, int64_t SY, int64_t SX,
               int pad_t, int pad_b, int pad_l, int pad_r, ConvShape* out) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || S <= 0 || KH <= 0 || KW <= 0 || SY <= 0 || SX <= 0)
    return fail(VTC_ERR_ARG, "convolutional path: bad shape");
  if (KH % SY != 0 || KW % SX != 0)
    return fail(VTC_ERR_UNSUPPORTED, "convolutional path: the kernel size (%lld x %lld) must be a multiple of the stride (%lld x %lld)",
                (long long)KH, (long long)KW, (long long)SY, (long long)SX);
  if (H < KH || W < KW || (H - KH) % SY != 0 || (W - KW) % SX != 0)
    return fail(VTC_ERR_ARG, "convolutional path: padded image (%lld x %lld) is not kernel + a whole number of strides "
                "(pad it with utils.convolutions.get_padding_amt)", (long long)H, (long long)W);
  if (pad_t < 0 || pad_b < 0 || pad_l < 0 || pad_r < 0 || pad_t + pad_b > H || pad_l + pad_r > W)
    return fail(VTC_ERR_ARG, "convolutional path: bad padding");
  ConvShape& cs = *out;
  ConvGeom& g = cs.g;
  g.b = (int)B, g.c = (int)C, g.h = (int)H, g.w = (int)W, g.s = (int)S, g.kh = (int)KH, g.kw = (int)KW;
  g.sy = (int)SY, g.sx = (int)SX, g.ty = (int)(KH / SY), g.tx = (int)(KW / SX);
  g.gh = (int)(H / SY), g.gw = (int)(W / SX);
  g.ch = g.gh - g.ty + 1, g.cw = g.gw - g.tx + 1;
  g.db = (int)(C * SY * SX);
  cs.nq = g.ty * g.tx;
  if (cs.nq > MAX_SEGMENTS) return fail(VTC_ERR_UNSUPPORTED, "convolutional path: at most %d kernel taps per stride cell (kernel / stride), got %d", MAX_SEGMENTS, cs.nq);
  cs.rows = B * g.gh * g.gw;
  if (cs.rows > (1ll << 31) - 4096) return fail(VTC_ERR_ARG, "convolutional path: too many grid rows for 32-bit tile coordinates");
  cs.pad_t = pad_t, cs.pad_b = pad_b, cs.pad_l = pad_l, cs.pad_r = pad_r;
  cs.per_kernel = C * KH * KW;
  return VTC_OK;
}

struct ConvWs {
  float* scalars;
  double* stats;
  LipschitzWs lip;
  PartsMat phiA_op, phiS_op, yop[2], r_op;
  float *xblk, *X1, *X2, *init_pad, *out_pad;
  int64_t ldS, ldD;
};
ConvWs carve_conv(Carver& cv, const ConvShape& cs, int precision) {
  ConvWs w;
  const ConvGeom& g = cs.g;
  const int P = parts_for(precision);
  const int bk = (P == 1) ? 64 : 32;
  const int64_t Sp = round_up(g.s, 64), Dbp = round_up(g.db, 64);
  w.scalars = static_cast<float*>(cv.take(64));
  w.stats = static_cast<double*>(cv.take(8 * kStatSlots));
  w.lip = carve_lipschitz(cv, cs.per_kernel);
  w.phiA_op = carve_parts(cv, g.s, cs.nq * Dbp, 3);
  w.phiS_op = carve_parts(cv, g.db, cs.nq * Sp, 3);
  w.ldS = round_up(g.s, 4);
  w.ldD = round_up(g.db, 4);
  w.xblk = static_cast<float*>(cv.take(static_cast<size_t>(cs.rows) * w.ldD * 4));
  w.yop[0] = carve_parts(cv, cs.rows, g.s, P, bk);
  w.yop[1] = carve_parts(cv, cs.rows, g.s, P, bk);
  w.r_op = carve_parts(cv, cs.rows, g.db, P, bk);
  const size_t state_blk = static_cast<size_t>(cs.rows) * round_up(g.s, EPI_COLS) * 4;
  w.X1 = static_cast<float*>(cv.take(state_blk));
  w.X2 = static_cast<float*>(cv.take(state_blk));
  w.init_pad = static_cast<float*>(cv.take(static_cast<size_t>(cs.rows) * w.ldS * 4));
  w.out_pad = static_cast<float*>(cv.take(static_cast<size_t>(cs.rows) * w.ldS * 4));
  return w;
}

// the two dictionary operands (analysis: N = code channels, synthesis: N = pixels of a block; taps along K)
int conv_dictionary_operands(const float* dictionary, const ConvShape& cs, const ConvWs& w, cudaStream_t st) {
  DeviceInfo info;
  TRY(device_info(&info));
  CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.phiA_op.ptr), 0, w.phiA_op.bytes(), st));
  CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.phiS_op.ptr), 0, w.phiS_op.bytes(), st));
  conv_dict_operands_kernel<<<grid_for(cs.g.s * cs.per_kernel, 256, info.sm_count), 256, 0, st>>>(
      dictionary, cs.g, 3, round_up(cs.g.db, 64), round_up(cs.g.s, 64),
      reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(w.phiA_op.ptr)),
      reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(w.phiS_op.ptr)));
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

// Halo staging of the taps: one band of rows per kernel row qy, the taps qx of that row start qx rows into it (analysis,
// positive shifts) or HALO_LEAD - qx rows into a band loaded HALO_LEAD rows early (synthesis, negative shifts).
// VTC_B200_CONV_HALO=0 keeps one tile per tap.
void conv_halo(GemmCall& c, const ConvGeom& g, bool negative) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("VTC_B200_CONV_HALO");
    enabled = e ? (atoi(e) != 0) : 1;
  }
  if (!enabled || g.ty != 2 || g.tx - 1 > HALO_ROWS - BLOCK_M - HALO_LEAD) return;
  c.halo_bands = g.ty;
  for (int qy = 0; qy < g.ty; ++qy) {
    c.halo_band_row[qy] = negative ? -(qy * g.gw) - HALO_LEAD : qy * g.gw;
    for (int qx = 0; qx < g.tx; ++qx) {
      c.halo_tap_band[qy * g.tx + qx] = qy;
      c.halo_tap_row[qy * g.tx + qx] = negative ? HALO_LEAD - qx : qx;
    }
  }
}

// r = mask * (conv_transpose(y) - x) on the block grid: synthesis contraction over (tap, channel), taps = row shifts
void conv_synthesis_call(GemmCall& r, const ConvShape& cs, const ConvWs& w, const PartsMat& y, int precision) {
  const ConvGeom& g = cs.g;
  r.precision = precision;
  r.A = y, r.B = w.phiS_op;
  r.M = cs.rows, r.N = g.db, r.K = static_cast<int64_t>(cs.nq) * g.s;
  r.nseg = cs.nq;
  for (int qy = 0; qy < g.ty; ++qy)
    for (int qx = 0; qx < g.tx; ++qx) r.seg_shift[qy * g.tx + qx] = -(qy * g.gw + qx);
  conv_halo(r, g, true);
  r.in[0] = F32Mat{w.xblk, cs.rows, g.db, w.ldD}, r.in_mask = 1;
  r.grid_h = g.gh, r.grid_w = g.gw;
  r.blk_sy = g.sy, r.blk_sx = g.sx;
  r.pix_y0 = cs.pad_t, r.pix_y1 = g.h - cs.pad_b, r.pix_x0 = cs.pad_l, r.pix_x1 = g.w - cs.pad_r;
}

}  // namespace

extern "C" {

size_t vtc_fista_conv_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW,
                                      int64_t SY, int64_t SX, int precision) {
  ConvShape cs;
  if (!valid_precision(precision) || conv_shape(B, C, H, W, S, KH, KW, SY, SX, 0, 0, 0, 0, &cs) != VTC_OK) return 0;
  Carver cv(nullptr, 0);
  carve_conv(cv, cs, precision);
  return cv.off + 2048;
}

int vtc_fista_conv(const float* images_padded, const float* dictionary, const float* initial_codes, float* codes_out,
                   int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW, int64_t SY,
                   int64_t SX, int pad_top, int pad_bottom, int pad_left, int pad_right, float sparsity_weight,
                   int num_iters, int variant, int nonnegative_only, int hard_threshold,
                   float early_stopping_epsilon, int precision, void* workspace, size_t workspace_bytes,
                   int* iters_run, float* lipschitz_out, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!images_padded || !dictionary || !codes_out) return fail(VTC_ERR_ARG, "vtc_fista_conv: null pointer");
  if (variant != VTC_VARIANT_ISTA && variant != VTC_VARIANT_FISTA) return fail(VTC_ERR_ARG, "variant must be ista or fista");
  if (!valid_precision(precision)) return fail(VTC_ERR_ARG, "precision must be 1, 3 or 6");
  if (num_iters < 1) return fail(VTC_ERR_ARG, "num_iters must be >= 1");
  ConvShape cs;
  TRY(conv_shape(B, C, H, W, S, KH, KW, SY, SX, pad_top, pad_bottom, pad_left, pad_right, &cs));
  const ConvGeom& g = cs.g;
  DeviceInfo info;
  TRY(require_sm100(&info));
  const bool early = early_stopping_epsilon >= 0.f;
  if (early && num_iters > 4096) return fail(VTC_ERR_UNSUPPORTED, "early stopping supports at most 4096 iterations");
  Carver cv(workspace, workspace_bytes);
  ConvWs w = carve_conv(cv, cs, precision);
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_fista_conv: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
  const int P = parts_for(precision);
  const int64_t R = cs.rows;

  // ---- setup: step size from the Gram matrix of the flattened kernels (convolutional/ista_fista.py:104-113),
  //      dictionary operands, images as blocks, starting point
  TRY(run_lipschitz(dictionary, S, cs.per_kernel, w.lip, sparsity_weight, w.scalars, nullptr, st));
  TRY(conv_dictionary_operands(dictionary, cs, w, st));
  conv_image_to_blocks_kernel<<<grid_for(R * g.db, 256, info.sm_count), 256, 0, st>>>(images_padded, g, w.xblk, w.ldD);
  COUNT_LAUNCH();
  CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.r_op.ptr), 0, w.r_op.bytes(), st));
  CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.yop[1].ptr), 0, w.yop[1].bytes(), st));
  F32Mat X1{w.X1, R, S, 0, true}, X2{w.X2, R, S, 0, true}, init, final_out{w.out_pad, R, S, w.ldS, false};
  if (initial_codes) {
    conv_codes_to_grid_kernel<<<grid_for(R * S, 256, info.sm_count), 256, 0, st>>>(initial_codes, g, w.init_pad, w.ldS);
    COUNT_LAUNCH();
    init = F32Mat{w.init_pad, R, S, w.ldS, false};
    TRY(split_rows(w.init_pad, w.ldS, R, S, w.yop[0], st));
  } else {
    CUDA_TRY(cudaMemsetAsync(w.X2, 0, static_cast<size_t>(R) * round_up(S, EPI_COLS) * 4, st));
    CUDA_TRY(cudaMemsetAsync(const_cast<void*>(w.yop[0].ptr), 0, w.yop[0].bytes(), st));
    init = X2;
  }
  if (early) CUDA_TRY(cudaMemsetAsync(w.stats, 0, sizeof(double) * num_iters, st));
  CUDA_TRY(cudaGetLastError());

  float eta_host = 0.f;
  float sc_host[4] = {0, 0, 0, 0};
  if (early || lipschitz_out) {
    CUDA_TRY(cudaMemcpyAsync(sc_host, w.scalars, sizeof(sc_host), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    eta_host = sc_host[0];
    if (lipschitz_out) *lipschitz_out = sc_host[2];
    if (sc_host[3] != 0.f || !isfinite(sc_host[2]))
      return fail(VTC_ERR_NONFINITE, "largest eigenvalue of the kernel Gram matrix is %g: a dictionary element overflowed", sc_host[2]);
  }

  // ---- iterations (convolutional/ista_fista.py:141-190): two launches each
  double t_k = 1.0;
  float beta_prev = 0.f;
  int k_done = 0;
  for (int k = 1; k <= num_iters; ++k) {
    const double t_next = (1.0 + sqrt(1.0 + 4.0 * t_k * t_k)) / 2.0;
    const float beta_k = (variant == VTC_VARIANT_FISTA) ? static_cast<float>((t_k - 1.0) / t_next) : 0.f;
    t_k = t_next;
    const F32Mat& a_prev = (k == 1) ? init : ((k - 1) & 1) ? X1 : X2;
    const F32Mat& a_prev2 = (k <= 2) ? init : (k & 1) ? X1 : X2;
    const F32Mat& a_out = (k == num_iters && !early) ? final_out : (k & 1) ? X1 : X2;
    {
      GemmCall r;
      conv_synthesis_call(r, cs, w, w.yop[(k - 1) & 1], precision);
      r.parts_out = w.r_op, r.n_parts = P;
      TRY(launch_gemm<EPI_STORE>(r, st));
    }
    GemmCall c;
    c.precision = precision;
    c.A = w.r_op, c.B = w.phiA_op;
    c.M = R, c.N = S, c.K = static_cast<int64_t>(cs.nq) * g.db;
    c.nseg = cs.nq;
    for (int qy = 0; qy < g.ty; ++qy)
      for (int qx = 0; qx < g.tx; ++qx) c.seg_shift[qy * g.tx + qx] = qy * g.gw + qx;
    conv_halo(c, g, false);
    c.grid_h = g.gh, c.grid_w = g.gw, c.code_h = g.ch, c.code_w = g.cw;
    c.in[0] = a_prev, c.in_mask = 1;
    if (variant == VTC_VARIANT_FISTA && beta_prev != 0.f) c.in[2] = a_prev2, c.in_mask |= 4;
    c.out = a_out, c.store_out = true;
    if (k < num_iters) c.parts_out = w.yop[k & 1], c.n_parts = P;
    c.prox = (hard_threshold ? PROX_HARD : 0) | (nonnegative_only ? PROX_NONNEG : 0);
    c.group = 1;
    c.use_momentum = (variant == VTC_VARIANT_FISTA);
    c.beta_prev = beta_prev, c.beta_next = beta_k;
    c.scalars = w.scalars;
    c.stat = early ? w.stats + (k - 1) : nullptr;
    TRY(launch_gemm<EPI_FISTA>(c, st));
    beta_prev = beta_k;
    k_done = k;
    if (early) {
      double sum_abs = 0.0;
      CUDA_TRY(cudaMemcpyAsync(&sum_abs, w.stats + (k - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      const double count = static_cast<double>(B) * S * g.ch * g.cw;
      if (sum_abs / count / static_cast<double>(eta_host) < static_cast<double>(early_stopping_epsilon) && k > 1) break;
    }
  }
  if (early) {
    const F32Mat& res = (k_done & 1) ? X1 : X2;
    unblock_f32_kernel<<<grid_for(R * S, 256, info.sm_count), 256, 0, st>>>(static_cast<const float*>(res.ptr), R, S,
                                                                            w.out_pad, w.ldS);
    COUNT_LAUNCH();
  }
  conv_grid_to_codes_kernel<<<grid_for(B * S * g.ch * g.cw, 256, info.sm_count), 256, 0, st>>>(w.out_pad, w.ldS, g, codes_out);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  if (iters_run) *iters_run = k_done;
  return VTC_OK;
}

// ---- convolutional dictionary update ------------------------------------------------------------------------------
}  // extern "C"
namespace {
struct ConvGradWs {
  ConvWs fw;             // reuses the inference carving for the operands, the image blocks and one y operand
  float* rblk;           // masked residual on the block grid (rows x ldD)
  PartsMat aT_op, RT_op; // codes^T and (tap-shifted) residual^T with the grid rows along K
  float* partial;
  int ksplits;
  int64_t rows_per_split, ldTap;
  float* taps;           // nq gradient blocks (s x ldTap each)
};
ConvGradWs carve_conv_grad(Carver& cv, const ConvShape& cs, int precision, int sm_count) {
  ConvGradWs w;
  const ConvGeom& g = cs.g;
  const int P = parts_for(precision);
  w.fw = carve_conv(cv, cs, precision);
  w.rblk = static_cast<float*>(cv.take(static_cast<size_t>(cs.rows) * w.fw.ldD * 4));
  w.aT_op = carve_parts(cv, g.s, cs.rows, P);
  w.RT_op = carve_parts(cv, g.db, cs.rows, P);
  const int64_t tiles_mn = ceil_div(g.s, PAIR_M) * ceil_div(g.db, BLOCK_N);
  const int64_t kb = k_blocks_for(w.aT_op.Kp, precision);
  int64_t ks = (sm_count / 2) / tiles_mn;
  if (ks < 1) ks = 1;
  if (ks > kb) ks = kb;
  w.ksplits = static_cast<int>(ks);
  w.rows_per_split = ceil_div(g.s, PAIR_M) * PAIR_M;
  w.ldTap = round_up(g.db, 4);
  w.partial = static_cast<float*>(cv.take(static_cast<size_t>(w.ksplits) * w.rows_per_split * w.ldTap * 4));
  w.taps = static_cast<float*>(cv.take(static_cast<size_t>(cs.nq) * g.s * w.ldTap * 4));
  return w;
}
}  // namespace
extern "C" {

size_t vtc_conv_dict_grad_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH,
                                          int64_t KW, int64_t SY, int64_t SX, int precision) {
  ConvShape cs;
  if (!valid_precision(precision) || conv_shape(B, C, H, W, S, KH, KW, SY, SX, 0, 0, 0, 0, &cs) != VTC_OK) return 0;
  Carver cv(nullptr, 0);
  carve_conv_grad(cv, cs, precision, 148);
  return cv.off + 2048;
}

int vtc_sc_conv_dict_grad(const float* images_padded, const float* dictionary, const float* codes, float* grad_sum,
                          int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW, int64_t SY,
                          int64_t SX, int pad_top, int pad_bottom, int pad_left, int pad_right, int precision,
                          void* workspace, size_t workspace_bytes, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!images_padded || !dictionary || !codes || !grad_sum) return fail(VTC_ERR_ARG, "vtc_sc_conv_dict_grad: null pointer");
  if (!valid_precision(precision)) return fail(VTC_ERR_ARG, "precision must be 1, 3 or 6");
  ConvShape cs;
  TRY(conv_shape(B, C, H, W, S, KH, KW, SY, SX, pad_top, pad_bottom, pad_left, pad_right, &cs));
  const ConvGeom& g = cs.g;
  DeviceInfo info;
  TRY(require_sm100(&info));
  Carver cv(workspace, workspace_bytes);
  ConvGradWs w = carve_conv_grad(cv, cs, precision, info.sm_count < 148 ? info.sm_count : 148);
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_sc_conv_dict_grad: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
  const int64_t R = cs.rows;
  // codes on the block grid (a^T as the operand of the gradient contraction; a as the operand of the synthesis)
  conv_codes_to_grid_kernel<<<grid_for(R * S, 256, info.sm_count), 256, 0, st>>>(codes, g, w.fw.init_pad, w.fw.ldS);
  COUNT_LAUNCH();
  TRY(split_rows(w.fw.init_pad, w.fw.ldS, R, S, w.fw.yop[0], st));
  TRY(transpose_split(w.fw.init_pad, w.fw.ldS, R, S, w.aT_op, st));
  TRY(conv_dictionary_operands(dictionary, cs, w.fw, st));
  conv_image_to_blocks_kernel<<<grid_for(R * g.db, 256, info.sm_count), 256, 0, st>>>(images_padded, g, w.fw.xblk, w.fw.ldD);
  COUNT_LAUNCH();
  // masked residual (convolutional/sc_cheap_quadratic_descent.py:65-69), fp32, then transposed into operand parts
  {
    GemmCall r;
    conv_synthesis_call(r, cs, w.fw, w.fw.yop[0], precision);
    r.out = F32Mat{w.rblk, R, g.db, w.fw.ldD}, r.store_out = true;
    TRY(launch_gemm<EPI_STORE>(r, st));
  }
  // one contraction over the grid rows per kernel tap: grad[s, tap, pix] = sum_m a[m, s] * r[m + shift(tap), pix];
  // the shift is taken while transposing the residual into its operand parts (rows past the end read as zero)
  const int64_t kb = k_blocks_for(w.aT_op.Kp, precision);
  const int64_t per = ceil_div(kb, w.ksplits);
  const int nsplit = static_cast<int>(ceil_div(kb, per));
  for (int qy = 0; qy < g.ty; ++qy)
    for (int qx = 0; qx < g.tx; ++qx) {
      const int q = qy * g.tx + qx;
      GemmCall c;
      c.A = w.aT_op, c.B = w.RT_op;
      c.precision = precision;
      c.M = S, c.N = g.db, c.K = w.aT_op.K;
      const int64_t shift = static_cast<int64_t>(qy) * g.gw + qx;
      TRY(transpose_split(w.rblk + shift * w.fw.ldD, w.fw.ldD, R - shift, g.db, w.RT_op, st));
      c.ksplits = w.ksplits;
      c.out_rows_per_split = w.rows_per_split;
      c.out = F32Mat{w.partial, static_cast<int64_t>(w.ksplits) * w.rows_per_split, g.db, w.ldTap}, c.store_out = true;
      TRY(launch_gemm<EPI_STORE>(c, st));
      reduce_partials_kernel<<<grid_for(S * g.db, 256, info.sm_count), 256, 0, st>>>(
          w.partial, nsplit, w.rows_per_split, w.ldTap, S, g.db, w.taps + static_cast<size_t>(q) * S * w.ldTap);
      COUNT_LAUNCH();
    }
  conv_grad_to_dict_layout_kernel<<<grid_for(S * cs.per_kernel, 256, info.sm_count), 256, 0, st>>>(
      w.taps, S * w.ldTap, w.ldTap, g, grad_sum);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_sc_conv_dict_apply(float* dictionary, const float* grad_sum, const float* hessian_diagonal, int64_t S,
                           int64_t per_kernel, int64_t batch_global, float stepsize, float lowest_code_val,
                           int normalize, vtc_stream_t stream) {
  if (!dictionary || !grad_sum || S <= 0 || per_kernel <= 0 || batch_global <= 0) return fail(VTC_ERR_ARG, "vtc_sc_conv_dict_apply: bad argument");
  conv_dict_apply_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      dictionary, grad_sum, hessian_diagonal, S, per_kernel, static_cast<float>(batch_global), stepsize,
      lowest_code_val, normalize);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_conv_hessian_diag_update(const float* codes, int64_t B, int64_t S, int64_t positions, int64_t batch_global,
                                 float* code_sq_sum, float* hessian_diagonal, int apply_ema, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!codes || !code_sq_sum || B <= 0 || S <= 0 || positions <= 0 || batch_global <= 0) return fail(VTC_ERR_ARG, "vtc_conv_hessian_diag_update: bad argument");
  if (apply_ema && !hessian_diagonal) return fail(VTC_ERR_ARG, "vtc_conv_hessian_diag_update: hessian_diagonal required");
  conv_channel_sq_sum_kernel<<<static_cast<unsigned>(S), 256, 0, st>>>(codes, B, S, positions, code_sq_sum);
  COUNT_LAUNCH();
  if (apply_ema) {
    hessian_ema_kernel<<<static_cast<unsigned>(ceil_div(S, 256)), 256, 0, st>>>(hessian_diagonal, code_sq_sum, S,
                                                                              static_cast<float>(batch_global));
    COUNT_LAUNCH();
  }
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- validation metrics
}  // extern "C"
namespace {
struct MetricsWs {
  PartsMat a_op, phiT_op;
  float *resid, *items;
  int64_t ldD;
};
MetricsWs carve_metrics(Carver& cv, int64_t B, int64_t S, int64_t D, int precision) {
  MetricsWs w;
  const int P = parts_for(precision);
  w.a_op = carve_parts(cv, B, S, P);
  w.phiT_op = carve_parts(cv, D, S, P);
  w.ldD = round_up(D, 4);
  w.resid = static_cast<float*>(cv.take(static_cast<size_t>(B) * w.ldD * 4));
  w.items = static_cast<float*>(cv.take(static_cast<size_t>(B) * METRIC_FIELDS * 4));
  return w;
}
constexpr int CONV_METRIC_CHUNKS = 64;
}  // namespace
extern "C" {

size_t vtc_sc_metrics_workspace_bytes(int64_t B, int64_t S, int64_t D, int precision) {
  if (!valid_precision(precision) || B <= 0 || S <= 0 || D <= 0) return 0;
  Carver cv(nullptr, 0);
  carve_metrics(cv, B, S, D, precision);
  return cv.off + 1024;
}

int vtc_sc_metrics(const float* images, int64_t ld_images, const float* dictionary, const float* codes,
                   int64_t ld_codes, int64_t B, int64_t S, int64_t D, const int32_t* group_slots, int64_t num_groups,
                   int64_t group_width, int precision, double* totals, void* workspace, size_t workspace_bytes,
                   vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!images || !dictionary || !codes || !totals) return fail(VTC_ERR_ARG, "vtc_sc_metrics: null pointer");
  if (B <= 0 || S <= 0 || D <= 0 || ld_images < D || ld_codes < S) return fail(VTC_ERR_ARG, "vtc_sc_metrics: bad shape");
  if (group_slots && (num_groups <= 0 || group_width <= 0)) return fail(VTC_ERR_ARG, "vtc_sc_metrics: bad group table");
  if (!valid_precision(precision)) return fail(VTC_ERR_ARG, "precision must be 1, 3 or 6");
  DeviceInfo info;
  TRY(require_sm100(&info));
  Carver cv(workspace, workspace_bytes);
  MetricsWs w = carve_metrics(cv, B, S, D, precision);
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_sc_metrics: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
  TRY(split_rows(codes, ld_codes, B, S, w.a_op, st));
  TRY(transpose_split(dictionary, D, S, D, w.phiT_op, st));
  // resid (B x D) = codes * dictionary - images   (training/sparse_coding.py:182, :196-197)
  {
    GemmCall g;
    g.A = w.a_op, g.B = w.phiT_op;
    g.precision = precision;
    g.M = B, g.N = D, g.K = S;
    if (tma_ok(images, ld_images)) {
      g.in[0] = F32Mat{images, B, D, ld_images};
    } else {
      CUDA_TRY(cudaMemcpy2DAsync(w.resid, w.ldD * 4, images, ld_images * 4, D * 4, B, cudaMemcpyDeviceToDevice, st));
      g.in[0] = F32Mat{w.resid, B, D, w.ldD};
    }
    g.in_mask = 1;
    g.out = F32Mat{w.resid, B, D, w.ldD}, g.store_out = true;
    TRY(launch_gemm<EPI_STORE>(g, st));
  }
  fc_metrics_rows_kernel<<<grid_for(B * 32, 256, info.sm_count), 256, 0, st>>>(
      w.resid, w.ldD, images, ld_images, codes, ld_codes, B, S, D, group_slots, num_groups,
      static_cast<int>(group_width), w.items);
  COUNT_LAUNCH();
  metrics_totals_kernel<<<1, 1024, 0, st>>>(w.items, B, static_cast<double>(D), static_cast<double>(S), totals);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

size_t vtc_sc_conv_metrics_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W, int64_t S, int64_t KH,
                                           int64_t KW, int64_t SY, int64_t SX, int precision) {
  ConvShape cs;
  if (!valid_precision(precision) || conv_shape(B, C, H, W, S, KH, KW, SY, SX, 0, 0, 0, 0, &cs) != VTC_OK) return 0;
  Carver cv(nullptr, 0);
  carve_conv(cv, cs, precision);
  cv.take(static_cast<size_t>(cs.rows) * round_up(cs.g.db, 4) * 4);
  cv.take(static_cast<size_t>(B) * (CONV_METRIC_CHUNKS + 1) * METRIC_FIELDS * 4);
  return cv.off + 2048;
}

int vtc_sc_conv_metrics(const float* images_padded, const float* dictionary, const float* codes, int64_t B, int64_t C,
                        int64_t H, int64_t W, int64_t S, int64_t KH, int64_t KW, int64_t SY, int64_t SX, int pad_top,
                        int pad_bottom, int pad_left, int pad_right, int precision, double* totals, void* workspace,
                        size_t workspace_bytes, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!images_padded || !dictionary || !codes || !totals) return fail(VTC_ERR_ARG, "vtc_sc_conv_metrics: null pointer");
  if (!valid_precision(precision)) return fail(VTC_ERR_ARG, "precision must be 1, 3 or 6");
  ConvShape cs;
  TRY(conv_shape(B, C, H, W, S, KH, KW, SY, SX, pad_top, pad_bottom, pad_left, pad_right, &cs));
  const ConvGeom& g = cs.g;
  if (pad_top + pad_bottom >= H || pad_left + pad_right >= W) return fail(VTC_ERR_ARG, "vtc_sc_conv_metrics: the padding leaves no pixels");
  if (B > 65535) return fail(VTC_ERR_ARG, "vtc_sc_conv_metrics: at most 65535 images per call");
  DeviceInfo info;
  TRY(require_sm100(&info));
  Carver cv(workspace, workspace_bytes);
  ConvWs w = carve_conv(cv, cs, precision);
  float* rblk = static_cast<float*>(cv.take(static_cast<size_t>(cs.rows) * w.ldD * 4));
  float* partials = static_cast<float*>(cv.take(static_cast<size_t>(B) * CONV_METRIC_CHUNKS * METRIC_FIELDS * 4));
  float* items = static_cast<float*>(cv.take(static_cast<size_t>(B) * METRIC_FIELDS * 4));
  if (!workspace || !cv.fits()) return fail(VTC_ERR_WORKSPACE, "vtc_sc_conv_metrics: workspace too small (%zu < %zu)", workspace_bytes, cv.off);
  const int64_t R = cs.rows;
  conv_codes_to_grid_kernel<<<grid_for(R * S, 256, info.sm_count), 256, 0, st>>>(codes, g, w.init_pad, w.ldS);
  COUNT_LAUNCH();
  TRY(split_rows(w.init_pad, w.ldS, R, S, w.yop[0], st));
  TRY(conv_dictionary_operands(dictionary, cs, w, st));
  conv_image_to_blocks_kernel<<<grid_for(R * g.db, 256, info.sm_count), 256, 0, st>>>(images_padded, g, w.xblk, w.ldD);
  COUNT_LAUNCH();
  // reconstruction minus image inside the un-padded region, zero outside it (training/sparse_coding.py:185-197)
  {
    GemmCall r;
    conv_synthesis_call(r, cs, w, w.yop[0], precision);
    r.out = F32Mat{rblk, R, g.db, w.ldD}, r.store_out = true;
    TRY(launch_gemm<EPI_STORE>(r, st));
  }
  conv_metrics_partial_kernel<<<dim3(CONV_METRIC_CHUNKS, static_cast<unsigned>(B)), 256, 0, st>>>(
      rblk, w.ldD, images_padded, codes, g, pad_top, pad_bottom, pad_left, pad_right, partials);
  COUNT_LAUNCH();
  conv_metrics_fold_kernel<<<static_cast<unsigned>(ceil_div(B * 32, 256)), 256, 0, st>>>(partials, CONV_METRIC_CHUNKS, B, items);
  COUNT_LAUNCH();
  const double pixels = static_cast<double>(C) * (H - pad_top - pad_bottom) * (W - pad_left - pad_right);
  metrics_totals_kernel<<<1, 1024, 0, st>>>(items, B, pixels, static_cast<double>(S) * g.ch * g.cw, totals);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_dict_change(const float* dictionary, const float* previous_dictionary, int64_t S, int64_t per_kernel,
                    float* mean_abs_change, vtc_stream_t stream) {
  if (!dictionary || !previous_dictionary || !mean_abs_change || S <= 0 || per_kernel <= 0)
    return fail(VTC_ERR_ARG, "vtc_dict_change: bad argument");
  dict_change_kernel<<<static_cast<unsigned>(ceil_div(S * 32, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dictionary, previous_dictionary, S, per_kernel, mean_abs_change);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

// ---------------------------------------------------------------------------------------------- data feed
int vtc_whitening_filter(int64_t h, int64_t w, double cutoff_low, double cutoff_high, double order,
                         int norm_and_threshold, float* filter_out, void* scratch8, vtc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!filter_out || !scratch8 || h <= 0 || w <= 0 || !(cutoff_high > 0.0) || !(order >= 1.0) || cutoff_low < 0.0)
    return fail(VTC_ERR_ARG, "vtc_whitening_filter: bad argument");
  DeviceInfo info;
  TRY(device_info(&info));
  unsigned long long* max_bits = static_cast<unsigned long long*>(scratch8);
  CUDA_TRY(cudaMemsetAsync(max_bits, 0, 8, st));
  const unsigned grid = grid_for(h * w, 256, info.sm_count);
  if (norm_and_threshold) {
    whitening_filter_max_kernel<<<grid, 256, 0, st>>>(h, w, cutoff_low, cutoff_high, order, max_bits);
    COUNT_LAUNCH();
  }
  whitening_filter_kernel<<<grid, 256, 0, st>>>(h, w, cutoff_low, cutoff_high, order, max_bits, norm_and_threshold,
                                                filter_out);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_spectrum_filter(void* spectrum, int64_t n, int64_t hw, int64_t c, const float* filter, vtc_stream_t stream) {
  if (!spectrum || !filter || n <= 0 || hw <= 0 || c <= 0) return fail(VTC_ERR_ARG, "vtc_spectrum_filter: bad argument");
  DeviceInfo info;
  TRY(device_info(&info));
  spectrum_filter_kernel<<<grid_for(n * hw * c, 256, info.sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float2*>(spectrum), n, hw, c, filter);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

int vtc_extract_patches(const float* images, int64_t n, int64_t h, int64_t w, int64_t c, const int32_t* corners,
                        int64_t B, int64_t ph, int64_t pw, float* patches, int64_t ld_patches, vtc_stream_t stream) {
  if (!images || !corners || !patches || n <= 0 || h <= 0 || w <= 0 || c <= 0 || B <= 0 || ph <= 0 || pw <= 0 ||
      ph > h || pw > w || ld_patches < ph * pw * c)
    return fail(VTC_ERR_ARG, "vtc_extract_patches: bad argument");
  DeviceInfo info;
  TRY(device_info(&info));
  extract_patches_kernel<<<grid_for(B * ph * pw * c, 256, info.sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      images, h, w, c, corners, B, ph, pw, patches, ld_patches);
  COUNT_LAUNCH();
  CUDA_TRY(cudaGetLastError());
  return VTC_OK;
}

}  // extern "C"
