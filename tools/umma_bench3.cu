// Micro-benchmark 3: the conv kernel's MMA / epilogue pipeline without any loads.
// Two groups per CTA; per "plane" the issuer waits for a free accumulator slot, issues the conv kernel's steady-state
// burst (first MMA split: N=64 accumulate + N=32 overwrite, then nkc*18-1 MMAs N=96 into rotating TMEM columns),
// commits acc_full; four epilogue warps per group wait acc_full, tcgen05.ld the finished slot (32 columns) and
// release it.  Switches remove one ingredient at a time.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench3 tools/umma_bench3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../deepwmh_b200/csrc/tc_primitives.cuh"

using namespace dwmh;

struct Cfg { int groups, nkc, rotate, split_first, epilogue, planes, same_b; };

constexpr int R = 8, CB = 32;

__global__ void __launch_bounds__(448, 1) bench(Cfg g, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t barv[2][2 * R];
  const uint32_t base = tc::smem_u32(smem);
  const int warp_abs = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp_abs / 7, warp = warp_abs % 7;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 2; ++q) for (int i = 0; i < R; ++i) { tc::mbar_init(tc::smem_u32(&barv[q][i]), 1); tc::mbar_init(tc::smem_u32(&barv[q][R + i]), 4); }
    tc::fence_barrier_init();
  }
  if (warp_abs == 1) tc::tmem_alloc(tc::smem_u32(&tmem_ptr), 512);
  tc::fence_proxy_async();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = tmem_ptr + grp * 256;
  const uint32_t acc_full = tc::smem_u32(&barv[grp][0]), acc_empty = tc::smem_u32(&barv[grp][R]);
  if (grp < g.groups) {
    if (warp == 1) {
      const bool leader = tc::elect_one();
      const uint32_t a_hi = (160u >> 4) | (1u << 14), b_hi = (128u >> 4) | (1u << 14);
      const uint32_t idesc0 = tc::instr_desc_f16(0, 128, 0);
      const uint32_t id1 = idesc0 | ((CB >> 3) << 17), id2 = idesc0 | (((2 * CB) >> 3) << 17), id3 = idesc0 | (((3 * CB) >> 3) << 17);
      const uint32_t b_lbo16 = 3 * CB, kstep_b = 2 * b_lbo16, tile16 = (32 * 3 * CB * 2) >> 4;
      const uint32_t b_lo_res = ((base + 90 * 1024) >> 4) | (b_lbo16 << 16);
      const uint32_t a_lbo_field = (2880u >> 4) << 16;
      uint32_t lo_slot = 0, fresh = 2, fresh_phase = 0, done = 0;     // slots 0,1 count as touched
      __syncwarp();
      const unsigned long long t0 = clock64();
      for (int t = 0; t < g.planes; ++t) {
        if (g.epilogue) tc::mbar_wait(acc_empty + 8 * fresh, fresh_phase ^ 1, 3);
        if (++fresh == R) { fresh = 0; fresh_phase ^= 1; }
        const bool contiguous = lo_slot + 3 <= R;
        const uint32_t col = tmem + (g.rotate && contiguous ? lo_slot * CB : 0);
        for (int kc = 0; kc < g.nkc; ++kc) {
          tc::tc_fence_after();
          if (leader) {
            const uint32_t a_lo0 = ((base + grp * 46080u + ((t * g.nkc + kc) & 3) * 11520u) >> 4) | a_lbo_field;
            uint32_t bl = b_lo_res + (g.same_b ? 0u : (uint32_t)kc * 9u * tile16);
#pragma unroll
            for (int sft = 0; sft < 9; ++sft) {
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                const uint64_t adesc = ((uint64_t)a_hi << 32) | (a_lo0 + (sft / 3) * 10 + (sft % 3) + kk * (5760 >> 4));
                if (sft == 0 && kk == 0 && kc == 0 && g.split_first) {
                  tc::umma_f16(col, adesc, ((uint64_t)b_hi << 32) | bl, id2, 1u);
                  tc::umma_f16(col + 2 * CB, adesc, ((uint64_t)b_hi << 32) | (bl + 2 * CB), id1, 0u);
                } else tc::umma_f16(col, adesc, ((uint64_t)b_hi << 32) | (bl + kk * kstep_b), id3, 1u);
              }
              if (!g.same_b) bl += tile16;
            }
          }
          __syncwarp();
        }
        if (leader) tc::umma_commit(acc_full + 8 * done);
        if (++done == R) done = 0;
        if (g.rotate) lo_slot = lo_slot + 1 == R ? 0 : lo_slot + 1;
        __syncwarp();
      }
      // drain: wait for the last commit
      const uint32_t last = (done + R - 1) % R;
      if (!g.epilogue) { /* nobody consumes: the phase of `last` flips every R planes */ tc::mbar_wait(acc_full + 8 * last, ((g.planes - 1) / R) & 1, 99); }
      const unsigned long long t1 = clock64();
      if (leader && blockIdx.x == 0) out[grp] = t1 - t0;
    } else if (warp >= 2 && warp < 6 && g.epilogue) {
      const int q = warp_abs & 3;
      const uint32_t tm_lane = tmem + ((uint32_t)(q * 32) << 16);
      uint32_t slot = 0, phase = 0, acc = 0;
      for (int t = 0; t < g.planes; ++t) {
        tc::mbar_wait(acc_full + 8 * slot, phase, 7);
        tc::tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t r[16];
          tc::tmem_ld16(tm_lane + (g.rotate ? slot * CB : 0) + ch * 16, r);
          tc::tmem_ld_wait();
          acc += r[3];
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(acc_empty + 8 * slot);
        if (++slot == R) { slot = 0; phase ^= 1; }
      }
      if (acc == 0xdeadbeef) out[8] = acc;
    }
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp_abs == 1) tc::tmem_dealloc(tmem_ptr, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  unsigned long long* o;
  cudaMalloc(&o, 16 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int P = 1024;
  const Cfg cfgs[] = {
      // groups nkc rotate split epilogue planes same_b
      {2, 2, 1, 1, 1, P, 0},   // the kernel's dec4-a pipeline
      {2, 2, 1, 1, 0, P, 0},   // no epilogue (no slot waits)
      {2, 2, 0, 1, 1, P, 0},   // fixed TMEM columns
      {2, 2, 1, 0, 1, P, 0},   // no split first MMA
      {2, 2, 0, 0, 0, P, 0},   // bare bursts
      {2, 2, 0, 0, 0, P, 1},   // bare bursts, one weight tile
      {1, 2, 1, 1, 1, P, 0},   // one group
      {2, 1, 1, 1, 1, P, 0},   // 32-channel layer (18 MMAs per plane)
      {2, 1, 0, 0, 0, P, 0},
  };
  printf("grid %d\n%-7s %-4s %-7s %-6s %-9s %-7s %12s %12s\n", grid, "groups", "nkc", "rotate", "split", "epilogue", "same_b", "cyc/plane", "cyc/mma(SM)");
  for (const Cfg& c : cfgs) {
    cudaMemset(o, 0, 16 * 8);
    for (int rep = 0; rep < 2; ++rep) {
      bench<<<grid, 448, 200 * 1024>>>(c, o);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    unsigned long long r[16];
    cudaMemcpy(r, o, 16 * 8, cudaMemcpyDeviceToHost);
    const double cyc = (double)(c.groups == 2 ? (r[0] > r[1] ? r[0] : r[1]) : r[0]) / c.planes;
    const int mmas = c.nkc * 18 + (c.split_first ? 1 : 0);
    printf("%-7d %-4d %-7d %-6d %-9d %-7d %12.1f %12.1f\n", c.groups, c.nkc, c.rotate, c.split_first, c.epilogue, c.same_b, cyc, cyc / (mmas * c.groups));
  }
  return 0;
}
