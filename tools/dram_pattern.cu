// Microbenchmark behind two design decisions (see profiles/README.md): (1) does HBM deliver less for 64-byte pieces at
// a 4 KB stride than for contiguous blocks? (2) what does the fused launch's exact traffic mix (2 fp32 tiles read,
// 1 fp32 + 2 bf16 tiles written, 128 rows x 16 columns each) reach with plain LSU accesses?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void stream(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o,
                       long rows, int cols4, int piece4, int blocked) {
  const long pieces_per_row = cols4 / piece4;
  const long items = (rows / 128) * pieces_per_row;
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const long mb = it / pieces_per_row, pc = it % pieces_per_row;
    for (int e = threadIdx.x; e < 128 * piece4; e += blockDim.x) {
      const int r = e / piece4, c = e % piece4;
      long idx;
      if (blocked) idx = (pc * rows + mb * 128 + r) * piece4 + c;
      else idx = (mb * 128 + r) * (long)cols4 + pc * piece4 + c;
      float4 x = a[idx], y = b[idx];
      o[idx] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
  }
}
// the fused launch's mix: per (128-row, 16-column) sub-tile read a, b (64 B per row each), write o (64 B) and two bf16
// part tiles (32 B per row each; row pitch 2 * cols * 2 bytes, parts side by side)
__global__ void fused_mix(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o,
                          uint4* __restrict__ parts, long rows, int cols, int blocked) {
  const int cols4 = cols / 4;
  const long sub_per_row = cols / 16;
  const long items = (rows / 128) * sub_per_row;
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const long mb = it / sub_per_row, sc = it % sub_per_row;
    for (int e = threadIdx.x; e < 128 * 4; e += blockDim.x) {
      const int r = e / 4, c = e % 4;
      const long row = mb * 128 + r;
      // blocked: [column block of 16][row][16] -> every sub-tile is one contiguous 8 KB (fp32) / 4 KB (bf16) block
      const long idx = blocked ? (sc * rows + row) * 4 + c : row * cols4 + sc * 4 + c;
      float4 x = a[idx], y = b[idx];
      float4 s = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
      o[idx] = s;
      if (c < 2) {  // 2 x 16 B per part per row
        if (blocked) {
          parts[(sc * rows + row) * 2 + c] = make_uint4(__float_as_uint(s.x), __float_as_uint(s.y), 0, 0);
          parts[((sub_per_row + sc) * rows + row) * 2 + c] = make_uint4(__float_as_uint(s.z), __float_as_uint(s.w), 0, 0);
        } else {
          const long prow = row * (2 * cols * 2 / 16);  // row pitch in uint4
          parts[prow + sc * 2 + c] = make_uint4(__float_as_uint(s.x), __float_as_uint(s.y), 0, 0);
          parts[prow + cols * 2 / 16 + sc * 2 + c] = make_uint4(__float_as_uint(s.z), __float_as_uint(s.w), 0, 0);
        }
      }
    }
  }
}
int main() {
  const long rows = 65536; const int cols = 1024, cols4 = cols / 4;
  size_t bytes = rows * cols * 4;
  float4 *a, *b, *o; uint4* parts;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, bytes); cudaMalloc(&parts, bytes);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int piece4 : {4, 8, 16, 64}) for (int blocked = 0; blocked < 2; ++blocked) {
    for (int rep = 0; rep < 2; ++rep) stream<<<148 * 4, 512>>>(a, b, o, rows, cols4, piece4, blocked);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) stream<<<148 * 4, 512>>>(a, b, o, rows, cols4, piece4, blocked);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("piece %4d B %s: %.1f GB/s\n", piece4 * 16, blocked ? "blocked   " : "row-major ", 3.0 * bytes * 5 / ms / 1e6);
  }
  for (int blocked = 0; blocked < 2; ++blocked) {
    for (int rep = 0; rep < 2; ++rep) fused_mix<<<148 * 4, 512>>>(a, b, o, parts, rows, cols, blocked);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 5; ++rep) fused_mix<<<148 * 4, 512>>>(a, b, o, parts, rows, cols, blocked);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("fused mix (2 fp32 reads, 1 fp32 + 2 bf16 writes) %s: %.1f GB/s\n", blocked ? "blocked" : "row-major",
           4.0 * bytes * 5 / ms / 1e6);
  }
  return 0;
}
