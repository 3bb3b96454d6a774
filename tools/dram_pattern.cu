// Microbenchmark: does HBM deliver less for 64-byte pieces at a 4 KB stride (what a 16-column fp32 sub-tile of a
// row-major (B x 1024) state array looks like) than for contiguous 8 KB blocks? Reads 2 arrays, writes 1, like K2.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void stream(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o,
                       long rows, int cols4, int piece4, int blocked) {
  // work item = (128-row block, piece of piece4 float4 per row); 512 threads: 4 lanes... generic mapping below
  const long pieces_per_row = cols4 / piece4;
  const long items = (rows / 128) * pieces_per_row;
  for (long it = blockIdx.x; it < items; it += gridDim.x) {
    const long mb = it / pieces_per_row, pc = it % pieces_per_row;
    for (int e = threadIdx.x; e < 128 * piece4; e += blockDim.x) {
      const int r = e / piece4, c = e % piece4;
      long idx;
      if (blocked) idx = (pc * rows + mb * 128 + r) * piece4 + c;            // [piece][row][piece4]
      else idx = (mb * 128 + r) * (long)cols4 + pc * piece4 + c;             // row-major
      float4 x = a[idx], y = b[idx];
      o[idx] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
  }
}
int main() {
  const long rows = 65536; const int cols = 1024, cols4 = cols / 4;
  size_t bytes = rows * cols * 4;
  float4 *a, *b, *o;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, bytes);
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
  return 0;
}
