// conv3d 3x3x3 / stride 1 / pad 1 as a tcgen05 implicit GEMM (sm_100a), hand-written.
//
// GEMM view: M = output voxels, N = output channels, K = 27 taps x Cin.  One CTA owns a 16(H) x 8(W)
// output tile (128 GEMM rows = one UMMA_M) of one sample and one block of CB output channels and
// MARCHES ALONG D (the slowest spatial axis):
//   * TMA loads one haloed input plane chunk [KC/8][18][10][8ch] per step (x/y halo by TMA OOB zero
//     fill, the D halo by skipping out-of-range planes).  The chunked HBM layout makes the box land in
//     shared memory directly as a K-major, no-swizzle UMMA operand: 16 B per (voxel, 8-channel chunk),
//     SBO = 160 B between the 8-voxel rows, LBO = 2880 B between channel chunks.  The 9 in-plane taps
//     are 9 start addresses into the same plane (dy*160 + dx*16): each input byte is read from L2 once
//     and reused 27 x CB times from shared memory.
//   * the three D-taps are STACKED ALONG N: input plane d feeds output planes d-1, d, d+1, whose fp32
//     accumulators sit in neighbouring TMEM column blocks, so one tcgen05.mma of N = 3*CB consumes an
//     A window once for three taps (narrow layers, Cout = 32, would otherwise be bound by the
//     128 B/clk shared-memory operand bandwidth at ~40 % of the tensor peak).
//   * accumulators form a ring of R = 512/CB TMEM slots; the epilogue warps drain a finished plane
//     (tcgen05.ld -> fp16/bf16 pack -> coalesced 16-B stores, InstanceNorm sum / sum-of-squares by a
//     transposing warp-shuffle reduction), zero the slot (tcgen05.st) and hand it back, so every MMA
//     accumulates and no per-slot first-touch flag is needed.
//   * weights are pre-packed on the host as ready-made K-major operand tiles [kc][tap_yx][k8][3*CB][8]
//     and either stay resident in shared memory for the whole CTA or stream through a ring of bulk
//     copies.
// Warp roles: 0 = TMA producer (activations), 1 = MMA issuer + TMEM owner, 2..5 = epilogue,
// 6 = bulk-copy producer (weights).  All hand-offs are mbarriers; nothing spins on __syncthreads.
#pragma once
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>
#include "common.cuh"
#include "tc_primitives.cuh"

namespace dwmh {

constexpr int TC_THREADS = 224;
constexpr int TC_TH = 16, TC_TW = 8;
constexpr int TC_PH = 18, TC_PW = 10;
constexpr int TC_PLANE_BYTES = TC_PH * TC_PW * 16;     // one 8-channel chunk of a haloed plane
constexpr int TC_MAX_SA = 4, TC_MAX_NB = 20, TC_MAX_R = 16;
constexpr int TC_SMEM_MAX = 232448;                    // 227 KB opt-in limit
constexpr int TC_SMEM_RESERVED = 2048;                 // barriers + TMEM pointer + statistics

struct TcKParams {
  const void* wpack; void* out; double* sums;
  int C0, C1, Cout, CB, KC, nkc, nkc0;
  int D, H, W, tilesH, tilesW, ZB, nzb, ncb;
  int SA, NB, resident, R, fmt;
  uint32_t a_stage_bytes, b_tile_bytes, off_b, off_bar;
};

template <typename T>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const TcKParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = tc::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int wi = blockIdx.x;
  const int tw = wi % p.tilesW; wi /= p.tilesW;
  const int th = wi % p.tilesH; wi /= p.tilesH;
  const int zb = wi % p.nzb; wi /= p.nzb;
  const int cb = wi % p.ncb; wi /= p.ncb;
  const int n = wi;
  const int h0 = th * TC_TH, w0 = tw * TC_TW;
  const int z_lo = zb * p.ZB, z_end = min(p.D, z_lo + p.ZB);
  const int SA = p.SA, NB = p.NB, R = p.R, CB = p.CB;

  const uint32_t bar = smem_base + p.off_bar;
  auto a_full = [&](int s) { return bar + 8u * s; };
  auto a_empty = [&](int s) { return bar + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar + 8u * (2 * SA + s); };
  auto b_empty = [&](int s) { return bar + 8u * (2 * SA + NB + s); };
  auto acc_full = [&](int s) { return bar + 8u * (2 * SA + 2 * NB + s); };
  auto acc_empty = [&](int s) { return bar + 8u * (2 * SA + 2 * NB + R + s); };
  const uint32_t nbar = 2 * SA + 2 * NB + 2 * R;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * nbar);
  float* s_stat = reinterpret_cast<float*>(smem + p.off_bar + 8 * nbar + 16);       // [2][CB]

  if (threadIdx.x == 0) {
    tc::prefetch_tensormap(&tmA0);
    tc::prefetch_tensormap(&tmA1);
    for (int s = 0; s < SA; ++s) { tc::mbar_init(a_full(s), 1); tc::mbar_init(a_empty(s), 1); }
    for (int s = 0; s < NB; ++s) { tc::mbar_init(b_full(s), 1); tc::mbar_init(b_empty(s), 1); }
    for (int s = 0; s < R; ++s) { tc::mbar_init(acc_full(s), 1); tc::mbar_init(acc_empty(s), 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_ptr_smem), 512);
  for (int i = threadIdx.x; i < 2 * CB; i += TC_THREADS) s_stat[i] = 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  if (warp == 0) {
    // ---------------- activation producer: one TMA box per (input plane, channel chunk) -----------
    if (lane == 0) {
      uint32_t it = 0;
      for (int zi = z_lo - 1; zi <= z_end; ++zi) {
        if (zi < 0 || zi >= p.D) continue;
        for (int kc = 0; kc < p.nkc; ++kc, ++it) {
          const int s = it % SA;
          tc::mbar_wait(a_empty(s), ((it / SA) & 1) ^ 1, 1);
          tc::mbar_arrive_expect_tx(a_full(s), p.a_stage_bytes);
          const bool first = kc < p.nkc0;
          const int c8 = first ? n * (p.C0 >> 3) + kc * (p.KC >> 3) : n * (p.C1 >> 3) + (kc - p.nkc0) * (p.KC >> 3);
          tc::tma_load_4d(smem_base + s * p.a_stage_bytes, first ? &tmA0 : &tmA1, a_full(s), (w0 - 1) * 8, h0 - 1, zi, c8);
        }
      }
    }
  } else if (warp == 6) {
    // ---------------- weight producer: ready-made operand tiles, bulk copies ----------------------
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)cb * (9 * p.nkc) * p.b_tile_bytes;
      if (p.resident) {
        for (int t = 0; t < 9 * p.nkc; ++t) {
          tc::mbar_arrive_expect_tx(b_full(t), p.b_tile_bytes);
          tc::bulk_load(smem_base + p.off_b + t * p.b_tile_bytes, wsrc + (size_t)t * p.b_tile_bytes, p.b_tile_bytes, b_full(t));
        }
      } else {
        uint32_t it = 0;
        for (int zi = z_lo - 1; zi <= z_end; ++zi) {
          if (zi < 0 || zi >= p.D) continue;
          for (int t = 0; t < 9 * p.nkc; ++t, ++it) {
            const int s = it % NB;
            tc::mbar_wait(b_empty(s), ((it / NB) & 1) ^ 1, 2);
            tc::mbar_arrive_expect_tx(b_full(s), p.b_tile_bytes);
            tc::bulk_load(smem_base + p.off_b + s * p.b_tile_bytes, wsrc + (size_t)t * p.b_tile_bytes, p.b_tile_bytes, b_full(s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc0 = tc::instr_desc_f16(p.fmt, 128, 0);
      const uint32_t b_lbo = 3u * CB * 16u;
      const int last_zi = min(z_end, p.D - 1);
      uint32_t a_it = 0, b_it = 0;
      int next_fresh = z_lo, next_done = z_lo;
      for (int zi = z_lo - 1; zi <= z_end; ++zi) {
        if (zi < 0 || zi >= p.D) continue;
        const int zo_lo = max(zi - 1, z_lo), zo_hi = min(zi + 1, z_end - 1);
        while (next_fresh <= zo_hi) {          // slot must have been zeroed by the epilogue
          const int u = next_fresh - z_lo;
          tc::mbar_wait(acc_empty(u % R), (u / R) & 1, 3);
          ++next_fresh;
        }
        tc::tc_fence_after();
        int nseg = 0, prev_slot = -2;
        uint32_t seg_col[3], seg_n[3], seg_j[3];
        for (int zo = zo_lo; zo <= zo_hi; ++zo) {
          const int slot = (zo - z_lo) % R;
          if (nseg > 0 && slot == prev_slot + 1 && seg_n[nseg - 1] + CB <= 256) seg_n[nseg - 1] += CB;
          else { seg_col[nseg] = slot * CB; seg_n[nseg] = CB; seg_j[nseg] = zo - zi + 1; ++nseg; }
          prev_slot = slot;
        }
        for (int kc = 0; kc < p.nkc; ++kc, ++a_it) {
          const int sa = a_it % SA;
          tc::mbar_wait(a_full(sa), (a_it / SA) & 1, 4);
          tc::tc_fence_after();
          const uint32_t a_base = smem_base + sa * p.a_stage_bytes;
          for (int sft = 0; sft < 9; ++sft) {
            int sb;
            if (p.resident) { sb = kc * 9 + sft; tc::mbar_wait(b_full(sb), 0, 5); }
            else { sb = b_it % NB; tc::mbar_wait(b_full(sb), (b_it / NB) & 1, 6); }
            tc::tc_fence_after();
            const uint32_t b_base = smem_base + p.off_b + sb * p.b_tile_bytes;
            const uint32_t a_tap = a_base + (sft / 3) * (TC_PW * 16) + (sft % 3) * 16;
            for (int kk = 0; kk < (p.KC >> 4); ++kk) {
              const uint64_t adesc = tc::smem_desc_kmajor_noswizzle(a_tap + kk * 2 * TC_PLANE_BYTES, TC_PLANE_BYTES, TC_PW * 16);
              for (int g = 0; g < nseg; ++g) {
                const uint64_t bdesc = tc::smem_desc_kmajor_noswizzle(b_base + kk * 2 * b_lbo + seg_j[g] * CB * 16, b_lbo, 128);
                tc::umma_f16(tmem + seg_col[g], adesc, bdesc, idesc0 | ((seg_n[g] >> 3) << 17), 1u);
              }
            }
            if (!p.resident) { tc::umma_commit(b_empty(sb)); ++b_it; }
          }
          tc::umma_commit(a_empty(sa));
        }
        const int done_upto = (zi == last_zi) ? z_end - 1 : zi - 1;
        while (next_done <= done_upto) { tc::umma_commit(acc_full((next_done - z_lo) % R)); ++next_done; }
      }
    }
  } else {
    // ---------------- epilogue: warps 2..5, TMEM lane quarter = warp % 4 ---------------------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int h = h0 + (row >> 3), w = w0 + (row & 7);
    const bool valid = h < p.H && w < p.W;
    const uint32_t tm_lane = tmem + ((uint32_t)(q * 32) << 16);
    const int nch = CB >> 4;
    for (int c = 0; c < R * CB; c += 16) tc::tmem_st16_zero(tm_lane + c);
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncwarp();
    if (lane == 0) for (int s = 0; s < R; ++s) tc::mbar_arrive(acc_empty(s));
    float rs[8], rq[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    const size_t V = (size_t)p.D * p.H * p.W;
    uint4* out_base = reinterpret_cast<uint4*>(p.out) + ((size_t)n * (p.Cout >> 3) + (size_t)cb * (CB >> 3)) * V;
    for (int zo = z_lo; zo < z_end; ++zo) {
      const int u = zo - z_lo, slot = u % R;
      tc::mbar_wait(acc_full(slot), (u / R) & 1, 7);
      tc::tc_fence_after();
      uint4* outp = out_base + ((size_t)zo * p.H + h) * p.W + w;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        if (ch < nch) {
          uint32_t r[16];
          tc::tmem_ld16(tm_lane + slot * CB + ch * 16, r);
          tc::tmem_ld_wait();
          float a[16], b[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { a[i] = valid ? __uint_as_float(r[i]) : 0.f; b[i] = a[i] * a[i]; }
          if (valid) {
            outp[(size_t)(2 * ch) * V] = pack8<T>(a);
            outp[(size_t)(2 * ch + 1) * V] = pack8<T>(a + 8);
          }
          // transposing butterfly: afterwards lane l holds the 32-row total of column (l & 15)
#pragma unroll
          for (int k = 8; k >= 1; k >>= 1) {
            const bool up = (lane & k) != 0;
#pragma unroll
            for (int i = 0; i < k; ++i) {
              const float sa_ = up ? a[i] : a[i + k], ka_ = up ? a[i + k] : a[i];
              const float sb_ = up ? b[i] : b[i + k], kb_ = up ? b[i + k] : b[i];
              a[i] = ka_ + __shfl_xor_sync(0xffffffffu, sa_, k);
              b[i] = kb_ + __shfl_xor_sync(0xffffffffu, sb_, k);
            }
          }
          rs[ch] += a[0] + __shfl_xor_sync(0xffffffffu, a[0], 16);
          rq[ch] += b[0] + __shfl_xor_sync(0xffffffffu, b[0], 16);
        }
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        if (ch < nch) tc::tmem_st16_zero(tm_lane + slot * CB + ch * 16);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(acc_empty(slot));
    }
    if (lane < 16) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        if (ch < nch) { atomicAdd(&s_stat[ch * 16 + lane], rs[ch]); atomicAdd(&s_stat[CB + ch * 16 + lane], rq[ch]); }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * CB; i += TC_THREADS) {
    const int c = i % CB, which = i / CB;
    atomicAdd(p.sums + ((size_t)n * p.Cout + (size_t)cb * CB + c) * 2 + which, (double)s_stat[i]);
  }
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// host side: configuration, weight packing, tensor maps, launch
// ------------------------------------------------------------------------------------------------
struct TcLayer {
  bool enabled = false;
  CUtensorMap tm0, tm1;
  void* wpack = nullptr;
  TcKParams kp{};
  size_t smem_bytes = 0;
};

inline void tc_free(TcLayer& t) {
  if (t.wpack) cudaFree(t.wpack);
  t.wpack = nullptr;
  t.enabled = false;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tc_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(f);
  }
  return fn;
}

// tensor map over one activation tensor [maxN][C/8][D][H][W][8]: dims (W*8, H, D, maxN*C/8), box (80, 18, 1, KC/8)
inline bool tc_make_map(CUtensorMap* m, const void* base, int maxN, int C, int D, int H, int W, int KC, bool bf16, std::string* why) {
  PFN_tmapEncodeTiled enc = tc_encode_fn();
  if (!enc) { *why = "cuTensorMapEncodeTiled unavailable"; return false; }
  cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)maxN * (C / 8)};
  cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
  cuuint32_t box[4] = {TC_PW * 8, TC_PH, 1, (cuuint32_t)(KC / 8)};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { *why = "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r); return false; }
  return true;
}

inline uint16_t tc_to_bits(float v, bool bf16) {
  if (bf16) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// Returns 0 always; t.enabled says whether the layer runs on the tensor cores.  *why is set only on
// a hard failure of a layer that should have been supported.
inline int tc_prepare(TcLayer& t, const std::vector<float>& w, int c0, int c1, int cout, const int k[3], const int s[3],
                      const int in_sp[3], const int out_sp[3], int maxN, bool bf16, const void* in0, const void* in1,
                      void* out, std::string* why) {
  t.enabled = false;
  why->clear();
  if (k[0] != 3 || k[1] != 3 || k[2] != 3 || s[0] != 1 || s[1] != 1 || s[2] != 1) return 0;
  if (c0 % 16 || c1 % 16 || cout % 16) return 0;
  (void)in_sp;
  const int cin = c0 + c1;
  int KC = 64;
  while (KC > 16 && (c0 % KC || c1 % KC)) KC >>= 1;
  const int budget = TC_SMEM_MAX - TC_SMEM_RESERVED;
  TcKParams& kp = t.kp;
  int CB = 0, SA = 0, NB = 0, resident = 0;
  for (;;) {
    const int a_stage = KC * 360;
    const int nkc = cin / KC;
    // resident weights: largest CB >= 32 (or the whole layer) that leaves room for >= 2 activation stages
    if (9 * nkc <= TC_MAX_NB) {
      for (int cbt = std::min(cout, 128); cbt >= 16; cbt -= 16) {
        if (cout % cbt) continue;
        if (cbt < 32 && cbt != cout) break;
        const long long btot = 27LL * cin * cbt * 2;
        if (btot + 2LL * a_stage <= budget) {
          CB = cbt; resident = 1; NB = 9 * nkc;
          SA = (int)std::min<long long>(TC_MAX_SA, (budget - btot) / a_stage);
          break;
        }
      }
    }
    if (!CB) {      // streaming weights
      for (int cbt = std::min(cout, 128); cbt >= 16; cbt -= 16) {
        if (cout % cbt) continue;
        const int b_tile = 3 * cbt * KC * 2;
        const int nb = (budget - 3 * a_stage) / b_tile;
        if (nb >= 3) { CB = cbt; resident = 0; SA = 3; NB = std::min(nb, TC_MAX_NB); break; }
      }
    }
    if (CB || KC == 16) break;
    KC >>= 1;
  }
  if (!CB) return 0;
  kp.C0 = c0; kp.C1 = c1; kp.Cout = cout; kp.CB = CB; kp.KC = KC; kp.nkc = cin / KC; kp.nkc0 = c0 / KC;
  kp.D = out_sp[0]; kp.H = out_sp[1]; kp.W = out_sp[2];
  kp.tilesH = (kp.H + TC_TH - 1) / TC_TH; kp.tilesW = (kp.W + TC_TW - 1) / TC_TW;
  kp.ncb = cout / CB; kp.SA = SA; kp.NB = NB; kp.resident = resident;
  kp.R = std::min(TC_MAX_R, 512 / CB);
  kp.fmt = bf16 ? 1 : 0;
  kp.a_stage_bytes = KC * 360; kp.b_tile_bytes = 3 * CB * KC * 2;
  kp.off_b = SA * kp.a_stage_bytes;
  kp.off_bar = kp.off_b + NB * kp.b_tile_bytes;
  kp.off_bar = (kp.off_bar + 127) & ~127u;
  t.smem_bytes = kp.off_bar + TC_SMEM_RESERVED;
  if (t.smem_bytes > TC_SMEM_MAX) { *why = "internal: shared memory plan exceeds 227 KB"; return 1; }
  if (t.smem_bytes < 120 * 1024) t.smem_bytes = 120 * 1024;          // one CTA per SM: the CTA owns all 512 TMEM columns
  kp.out = out;
  // operand tiles: [cb][kc][tap_yx][k8][row = j*CB + co][8], j = 2 - kd (output plane d-1, d, d+1)
  const size_t tile_elems = (size_t)(KC / 8) * 3 * CB * 8;
  std::vector<uint16_t> pk((size_t)kp.ncb * kp.nkc * 9 * tile_elems);
  for (int cb = 0; cb < kp.ncb; ++cb)
    for (int kc = 0; kc < kp.nkc; ++kc)
      for (int sft = 0; sft < 9; ++sft) {
        uint16_t* tile = pk.data() + (((size_t)cb * kp.nkc + kc) * 9 + sft) * tile_elems;
        for (int k8 = 0; k8 < KC / 8; ++k8)
          for (int row = 0; row < 3 * CB; ++row)
            for (int e = 0; e < 8; ++e) {
              const int co = cb * CB + row % CB, kd = 2 - row / CB, ci = kc * KC + k8 * 8 + e;
              const float v = w[((size_t)co * cin + ci) * 27 + kd * 9 + sft];
              tile[((size_t)k8 * 3 * CB + row) * 8 + e] = tc_to_bits(v, bf16);
            }
      }
  if (cudaMalloc(&t.wpack, pk.size() * 2) != cudaSuccess) { *why = "cudaMalloc(wpack) failed"; return 1; }
  if (cudaMemcpy(t.wpack, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) { *why = "cudaMemcpy(wpack) failed"; return 1; }
  kp.wpack = t.wpack;
  if (!tc_make_map(&t.tm0, in0, maxN, c0, kp.D, kp.H, kp.W, KC, bf16, why)) return 1;
  if (c1 > 0) { if (!tc_make_map(&t.tm1, in1, maxN, c1, kp.D, kp.H, kp.W, KC, bf16, why)) return 1; }
  else t.tm1 = t.tm0;
  t.enabled = true;
  return 0;
}

inline int tc_init_attributes(bool bf16) {
  cudaError_t e = bf16 ? cudaFuncSetAttribute(conv3_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX)
                       : cudaFuncSetAttribute(conv3_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX);
  return e == cudaSuccess ? 0 : 1;
}

template <typename T>
int tc_launch(TcLayer& t, int nb, double* sums, int num_sms, cudaStream_t st, std::string* err) {
  TcKParams kp = t.kp;
  kp.sums = sums;
  const int tiles = kp.tilesH * kp.tilesW;
  int ZB = kp.D;
  while ((long long)nb * kp.ncb * tiles * ((kp.D + ZB - 1) / ZB) < 2LL * num_sms && ZB > 8) ZB = (ZB + 1) / 2;
  kp.ZB = ZB; kp.nzb = (kp.D + ZB - 1) / ZB;
  const unsigned grid = (unsigned)((long long)nb * kp.ncb * kp.nzb * tiles);
  conv3_tc_kernel<T><<<grid, TC_THREADS, t.smem_bytes, st>>>(t.tm0, t.tm1, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { if (err) *err = std::string("conv3_tc_kernel launch failed: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

}  // namespace dwmh
