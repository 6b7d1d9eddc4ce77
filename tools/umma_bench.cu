// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16) as a function of the shared-memory
// operand layout (no-swizzle K-major with various SBO / start alignments, and 128B-swizzle for
// reference) and of N.  Data is garbage; only timing matters.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../deepwmh_b200/csrc/tc_primitives.cuh"

using namespace dwmh;

struct Cfg { uint32_t a_off, a_sbo, a_lbo, b_off, b_sbo, b_lbo, N, layout, niter; };

__global__ void __launch_bounds__(128, 1) bench(const Cfg* cfgs, int ncfg, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t barv;
  const uint32_t base = tc::smem_u32(smem);
  const uint32_t bar = tc::smem_u32(&barv);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(tc::smem_u32(&tmem_ptr), 512);
  tc::fence_proxy_async();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x < 32) {
    const bool leader = tc::elect_one();
    uint32_t phase = 0;
    for (int c = 0; c < ncfg; ++c) {
      const Cfg g = cfgs[c];
      const uint32_t hi_a = ((g.a_sbo >> 4) & 0x3FFF) | (1u << 14) | (g.layout << 29);
      const uint32_t hi_b = ((g.b_sbo >> 4) & 0x3FFF) | (1u << 14) | (g.layout << 29);
      const uint32_t lo_a = ((base + g.a_off) >> 4) | (((g.a_lbo >> 4) & 0x3FFF) << 16);
      const uint32_t lo_b = ((base + 100 * 1024 + g.b_off) >> 4) | (((g.b_lbo >> 4) & 0x3FFF) << 16);
      const uint32_t idesc = tc::instr_desc_f16(0, 128, g.N);
      const uint64_t ad = ((uint64_t)hi_a << 32) | lo_a, bd = ((uint64_t)hi_b << 32) | lo_b;
      __syncwarp();
      const unsigned long long t0 = clock64();
      for (uint32_t i = 0; i < g.niter; ++i)
        if (leader) tc::umma_f16(tmem, ad, bd, idesc, 1u);
      if (leader) tc::umma_commit(bar);
      tc::mbar_wait(bar, phase, 99);
      phase ^= 1;
      const unsigned long long t1 = clock64();
      if (leader) out[c] = t1 - t0;
      __syncwarp();
    }
  }
  tc::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 512);
}

int main() {
  Cfg h[64]; int n = 0;
  const uint32_t NIT = 512;
  const uint32_t Ns[] = {32, 96, 192, 256};
  for (uint32_t N : Ns) {
    // A: conv layout (SBO 160, LBO 2880) at several tap offsets; B: packed (SBO 128, LBO = N*16)
    h[n++] = {0, 160, 2880, 0, 128, N * 16, N, 0, NIT};
    h[n++] = {16, 160, 2880, 0, 128, N * 16, N, 0, NIT};
    h[n++] = {176, 160, 2880, 0, 128, N * 16, N, 0, NIT};
    // A aligned & dense (SBO 128, LBO 2048)
    h[n++] = {0, 128, 2048, 0, 128, N * 16, N, 0, NIT};
    h[n++] = {16, 128, 2048, 0, 128, N * 16, N, 0, NIT};
    // A with SBO 256 (aligned groups, gaps)
    h[n++] = {0, 256, 4096, 0, 128, N * 16, N, 0, NIT};
    // canonical 128B swizzle, K-major: SBO 1024 (8 rows x 128 B), LBO unused(1)
    h[n++] = {0, 1024, 16, 0, 1024, 16, N, 2, NIT};
    // 128B swizzle with start offsets (shifted rows: +128 B, +32 B k-advance)
    h[n++] = {128, 1024, 16, 0, 1024, 16, N, 2, NIT};
    // 32B swizzle (layout 6): rows of 32 B, SBO 256
    h[n++] = {0, 256, 16, 0, 256, 16, N, 6, NIT};
    // 64B swizzle (layout 4): rows of 64 B, SBO 512
    h[n++] = {0, 512, 16, 0, 512, 16, N, 4, NIT};
  }
  Cfg* d; unsigned long long* o;
  cudaMalloc(&d, sizeof h); cudaMalloc(&o, 64 * 8);
  cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    bench<<<1, 128, 200 * 1024>>>(d, n, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  unsigned long long r[64];
  cudaMemcpy(r, o, 64 * 8, cudaMemcpyDeviceToHost);
  printf("%-6s %-6s %-6s %-6s %-6s %-7s %10s %8s\n", "N", "layout", "a_off", "a_sbo", "a_lbo", "b_lbo", "cyc/mma", "ideal");
  for (int i = 0; i < n; ++i)
    printf("%-6u %-6u %-6u %-6u %-6u %-7u %10.1f %8.1f\n", h[i].N, h[i].layout, h[i].a_off, h[i].a_sbo, h[i].a_lbo, h[i].b_lbo,
           (double)r[i] / h[i].niter, h[i].N / 2.0);
  return 0;
}
