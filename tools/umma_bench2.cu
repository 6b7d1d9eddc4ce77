// Micro-benchmark 2: what limits tcgen05.mma throughput inside the conv kernel?
// One CTA per SM (grid = 1 or 148).  Up to two issuing warps ("groups") each run a dependent chain of
// M=128, K=16 fp16 MMAs with the conv kernel's operand layout (A: SBO 160 / LBO 2880, B: SBO 128 / LBO N*16) into
// their own TMEM columns.  Optional background load: warps that read TMEM (tcgen05.ld, like the epilogue), warps that
// stream shared memory (LDS.128 + STS.128, like the transform warps) and bulk copies into shared memory (like TMA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench2 tools/umma_bench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../deepwmh_b200/csrc/tc_primitives.cuh"

using namespace dwmh;

struct Cfg { uint32_t N, issuers, tmem_readers, smem_streamers, bulk, niter, commit_every, random_data; };   // commit_every: tcgen05.commit (to a barrier nobody waits on) after every burst of 18 MMAs

__global__ void __launch_bounds__(384, 1) bench(Cfg g, const uint8_t* gsrc, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t barv[4];
  __shared__ volatile int stop;
  const uint32_t base = tc::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (g.random_data) {        // two fp16 values in [-1, 1): exponent 0x3800..0x3BFF pattern with random mantissa / sign
      uint32_t x = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
      v = (x & 0x83FF83FFu) | 0x38003800u;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(tc::smem_u32(&barv[i]), 1); stop = 0; tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_ptr), 512);
  tc::fence_proxy_async();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  // warps 0,1: issuers; 2..5: TMEM readers; 6..9: smem streamers; 10: bulk-copy producer
  if (warp < 2) {
    if (warp < (int)g.issuers) {
      const bool leader = tc::elect_one();
      const uint32_t bar = tc::smem_u32(&barv[warp]);
      const uint32_t hi_a = (160u >> 4) | (1u << 14), hi_b = (128u >> 4) | (1u << 14);
      const uint32_t a0 = ((base + warp * 46080u) >> 4) | ((2880u >> 4) << 16);             // 4 stages of 11.5 KB per group
      const uint32_t b0 = ((base + 96 * 1024 + warp * 0u) >> 4) | (((g.N * 16u) >> 4) << 16);   // weights shared
      const uint32_t idesc = tc::instr_desc_f16(0, 128, g.N);
      const uint32_t ntile = 49152u / (64u * g.N) ? 49152u / (64u * g.N) : 1u;
      __syncwarp();
      const unsigned long long t0 = clock64();
      for (uint32_t i = 0; i < g.niter; i += 18) {
        if (leader) {
#pragma unroll
          for (int s = 0; s < 18; ++s) {        // 9 taps x 2 k-steps, like one plane of a 32-channel layer
            const uint32_t sh = (s >> 1) / 3 * 10 + (s >> 1) % 3 + (s & 1) * (5760 >> 4) + ((i / 18) & 3) * (11520 >> 4);
            tc::umma_f16(tmem + warp * 256, ((uint64_t)hi_a << 32) | (a0 + sh), ((uint64_t)hi_b << 32) | (b0 + ((s >> 1) % ntile) * ((g.N * 32u * 2u) >> 4) + (s & 1) * ((g.N * 16u * 2u) >> 4)), idesc, 1u);
          }
        }
        if (g.commit_every && leader) tc::umma_commit(tc::smem_u32(&barv[3]));
        if (g.commit_every >= 3) tc::mbar_test_wait(tc::smem_u32(&barv[2]), 1);      // a barrier probe per burst (phase parity 1 = "previous phase": completes at once)
        if (g.commit_every >= 2) tc::tc_fence_after();
        __syncwarp();
      }
      if (leader) tc::umma_commit(bar);
      tc::mbar_wait(bar, 0, 99);
      const unsigned long long t1 = clock64();
      if (leader && blockIdx.x == 0) out[warp] = t1 - t0;
      __syncwarp();
      if (warp == 0 && lane == 0) { if (g.issuers == 2) { tc::mbar_wait(tc::smem_u32(&barv[1]), 0, 98); } stop = 1; }
    }
  } else if (warp < 6) {
    if (warp - 2 < (int)g.tmem_readers) {
      const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      uint32_t acc = 0;
      while (!stop) {
        uint32_t r[16];
        tc::tmem_ld16(tm_lane + 448 + (acc & 16), r);       // columns the MMAs do not touch
        tc::tmem_ld_wait();
        acc += r[0] + 16;
      }
      if (acc == 0xdeadbeef) out[8] = acc;
    }
  } else if (warp < 10) {
    if (warp - 6 < (int)g.smem_streamers) {
      uint4* buf = reinterpret_cast<uint4*>(smem + 150 * 1024 + (warp - 6) * 8192);
      uint32_t acc = 0;
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { uint4 v = buf[lane + 32 * i]; v.x += acc; buf[lane + 32 * i] = v; acc += v.y; }
      }
      if (acc == 0xdeadbeef) out[9] = acc;
    }
  } else if (warp == 10) {
    if (g.bulk) {
      const uint32_t bar = tc::smem_u32(&barv[2]);
      uint32_t phase = 0;
      const bool leader = tc::elect_one();
      uint32_t it = 0;
      while (!stop) {
        if (leader) {
          tc::mbar_arrive_expect_tx(bar, 11520);
          tc::bulk_load(base + 184 * 1024, gsrc + (size_t)((blockIdx.x * 977u + it * 131u) & 0xFFFFu) * 11520u, 11520, bar);
        }
        tc::mbar_wait(bar, phase, 97);
        phase ^= 1; ++it;
      }
    }
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  unsigned long long* o; uint8_t* src;
  cudaMalloc(&o, 16 * 8);
  cudaMalloc(&src, (size_t)65536 * 11520 + 65536);
  cudaMemset(src, 0, (size_t)65536 * 11520);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const uint32_t NIT = 18 * 2048;
  const Cfg cfgs[] = {
      {96, 1, 0, 0, 0, NIT, 0, 0}, {96, 2, 0, 0, 0, NIT, 0, 0}, {96, 2, 4, 0, 0, NIT, 0, 0}, {96, 2, 0, 4, 0, NIT, 0, 0}, {96, 2, 0, 0, 1, NIT, 0, 0}, {96, 2, 4, 4, 1, NIT, 0, 0},
      {96, 1, 0, 0, 0, NIT, 1, 0}, {96, 2, 0, 0, 0, NIT, 1, 0}, {96, 2, 4, 4, 1, NIT, 1, 0},
      {192, 1, 0, 0, 0, NIT, 0, 0}, {192, 2, 0, 0, 0, NIT, 0, 0}, {192, 2, 4, 4, 1, NIT, 1, 0}, {96, 2, 0, 0, 0, NIT, 2, 0}, {96, 2, 0, 0, 0, NIT, 3, 0}, {96, 1, 0, 0, 0, NIT, 3, 0}, {96, 2, 4, 4, 1, NIT, 3, 1}, {192, 2, 0, 0, 0, NIT, 0, 1}, {64, 2, 0, 0, 0, NIT, 0, 0}, {128, 2, 0, 0, 0, NIT, 0, 0}, {256, 2, 0, 0, 0, NIT, 0, 0}};
  printf("grid %d\n%-5s %-8s %-8s %-8s %-5s %-7s %-5s %12s %12s %8s\n", grid, "N", "issuers", "tmem_rd", "smem_st", "bulk", "commit", "rand", "cyc/mma(w0)", "cyc/mma(SM)", "N/2");
  for (const Cfg& c : cfgs) {
    cudaMemset(o, 0, 16 * 8);
    for (int rep = 0; rep < 2; ++rep) {
      bench<<<grid, 384, 200 * 1024>>>(c, src, o);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    unsigned long long r[16];
    cudaMemcpy(r, o, 16 * 8, cudaMemcpyDeviceToHost);
    const double per0 = (double)r[0] / c.niter;
    const double tot = (double)(c.issuers == 2 ? (r[0] > r[1] ? r[0] : r[1]) : r[0]) / (c.niter * c.issuers);
    printf("%-5u %-8u %-8u %-8u %-5u %-7u %-5u %12.1f %12.1f %8.1f\n", c.N, c.issuers, c.tmem_readers, c.smem_streamers, c.bulk, c.commit_every, c.random_data, per0, tot, c.N / 2.0);
  }
  return 0;
}
