// What can the L2 deliver to the SMs? Every CTA reads the same L2-resident buffer (32 MB, < 126 MB L2) over and over with
// cp.async.bulk (the TMA path the kernels use: 8 KB pieces into shared memory, 4 in flight per CTA) and with ld.global.v4.
// Prints TB/s delivered to the SMs. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_probe tools/l2_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int STAGES, int PIECE>
__global__ void __launch_bounds__(128, 1) bulk_read(const uint8_t* __restrict__ buf, size_t bytes, int rounds) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const size_t pieces = bytes / PIECE;
  size_t idx = (static_cast<size_t>(blockIdx.x) * (pieces / gridDim.x) * 37) % pieces;   // CTAs start far apart
  const long total = static_cast<long>(rounds) * static_cast<long>(pieces / gridDim.x);
  for (long it = 0; it < total + STAGES; ++it) {
    const int s = it % STAGES;
    const uint32_t bar = smem_u32(&bars[s]);
    if (it >= STAGES) {
      const uint32_t parity = ((it / STAGES) - 1) & 1;
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    if (it < total) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(PIECE) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(smem + s * PIECE)), "l"(buf + idx * PIECE), "r"(PIECE), "r"(bar) : "memory");
      idx += 1;   // every CTA walks the whole buffer on its own: no two CTAs ask for a line at the same time
      if (idx >= pieces) idx -= pieces;
    }
  }
}

__global__ void __launch_bounds__(1024, 1) ldg_read(const uint4* __restrict__ buf, size_t n16, int rounds, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t per_block = n16 / gridDim.x;
  const size_t start = (static_cast<size_t>(blockIdx.x) * 37 % gridDim.x) * per_block;
  (void)stride;
  for (int r = 0; r < rounds; ++r)
    for (size_t i = threadIdx.x; i < per_block; i += blockDim.x) {
      size_t j = start + static_cast<size_t>(r) * 4099 * 64 + i;   // a different slice of the buffer every round
      j %= n16;
      const uint4 v = __ldcg(buf + j);
      acc.x ^= v.x, acc.y ^= v.y, acc.z ^= v.z, acc.w ^= v.w;
    }
  if (acc.x == 0x12345678u) *sink = acc;
}

int main() {
  const size_t bytes = 32u << 20;
  uint8_t* buf;
  uint4* sink;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&sink, 16);
  cudaMemset(buf, 1, bytes);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto report = [&](const char* name, double total_bytes) {
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-46s %.2f TB/s delivered to %d SMs (%.2f ms) %s\n", name, total_bytes / (ms * 1e-3) / 1e12, sms, ms,
           cudaGetErrorString(cudaGetLastError()));
  };
  const int rounds = 1500;
  {
    constexpr int ST = 4, PC = 8192;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      bulk_read<ST, PC><<<sms, 128, ST * PC>>>(buf, bytes, rounds);
      cudaEventRecord(e1);
      if (rep) report("cp.async.bulk 8 KB pieces, 4 in flight / SM", double(rounds) * (bytes / PC / sms) * sms * PC);
      else cudaEventSynchronize(e1);
    }
  }
  {
    constexpr int ST = 8, PC = 16384;
    cudaFuncSetAttribute(bulk_read<ST, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * PC);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      bulk_read<ST, PC><<<sms, 128, ST * PC>>>(buf, bytes, rounds);
      cudaEventRecord(e1);
      if (rep) report("cp.async.bulk 16 KB pieces, 8 in flight / SM", double(rounds) * (bytes / PC / sms) * sms * PC);
      else cudaEventSynchronize(e1);
    }
  }
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    ldg_read<<<sms, 1024>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, rounds, sink);
    cudaEventRecord(e1);
    if (rep) report("ld.global.cg.v4, 1024 threads / SM", double(rounds) * (bytes / 16 / sms) * sms * 16);
    else cudaEventSynchronize(e1);
  }
  return 0;
}
