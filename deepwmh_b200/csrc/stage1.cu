// the stage-1 NLL anomaly map of deepwmh/analysis/lesion_analysis.py:84-176 on the device.
// Voxel-parallel, HBM-bound kernels; fp32 storage, fp64 per-voxel arithmetic (the reference works in float64 once a
// volume has been z-scored).  Context-free entry points (no network involved): `device` + caller-owned buffers.
//
//   dwmh_s1_zscore            z_score (image_ops.py:172-179) [+ tissue-min fill, lesion_analysis.py:150-151,160-161]
//   dwmh_s1_mean_std_grid     mean_std_grid, order 1 (image_ops.py:56-170)
//   dwmh_s1_align_local_mean  x_i - x_i_local_mu + x_prime_local_mu (lesion_analysis.py:166-169)
//   dwmh_s1_group_nll         group_mean / group_std / nll (image_ops.py:197-231, lesion_analysis.py:84-113)
//   dwmh_s1_median_filter     median_filter(mode='constant', cval=0) behind median_3mm (image_ops.py:181-183,378-421)
#include "../../include/deepwmh_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"

extern "C" void dwmh_internal_set_error(const char* msg);   // api.cu (thread-local message behind dwmh_last_error)

namespace {

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  dwmh_internal_set_error(buf);
  return 1;
}
#define S1_CU(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
  return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

int grid_for(int device) {
  static int sms[64] = {0};
  if (device < 0 || device >= 64) return 148 * 8;
  if (!sms[device]) cudaDeviceGetAttribute(&sms[device], cudaDevAttrMultiProcessorCount, device);
  return (sms[device] > 0 ? sms[device] : 148) * 8;
}

using dwmh::warp_sum_d;

// order-preserving float <-> signed int (atomicMin on floats)
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---------------------------------------------------------------------------------------------------------------
// z_score.  ws: double[4] = {sum, sumsq, count, -} + int min (ordered) at byte 32.  Pass 1 reads x (+mask), pass 2
// reads x (+mask) and writes x: 12 B/voxel (+8 with a mask).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) s1_stats_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                       int64_t n, double* __restrict__ acc, int* __restrict__ minord) {
  double s = 0.0, ss = 0.0, cnt = 0.0;
  int mn = 0x7fffffff;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    if (!mask || mask[i] > 0.5f) { s += v; ss += (double)v * v; cnt += 1.0; mn = min(mn, f2ord(v)); }
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss); cnt = warp_sum_d(cnt);
  mn = __reduce_min_sync(0xffffffffu, mn);
  __shared__ double sh[3][8];
  __shared__ int shm[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s; sh[1][w] = ss; sh[2][w] = cnt; shm[w] = mn; }
  __syncthreads();
  if (w == 0) {
    s = l < 8 ? sh[0][l] : 0.0; ss = l < 8 ? sh[1][l] : 0.0; cnt = l < 8 ? sh[2][l] : 0.0; mn = l < 8 ? shm[l] : 0x7fffffff;
    s = warp_sum_d(s); ss = warp_sum_d(ss); cnt = warp_sum_d(cnt); mn = __reduce_min_sync(0xffffffffu, mn);
    if (l == 0) { atomicAdd(acc + 0, s); atomicAdd(acc + 1, ss); atomicAdd(acc + 2, cnt); atomicMin(minord, mn); }
  }
}

__global__ void __launch_bounds__(256) s1_zscore_apply_kernel(float* __restrict__ x, const float* __restrict__ mask, int64_t n,
                                                              const double* __restrict__ acc, const int* __restrict__ minord,
                                                              int fill_outside) {
  const double cnt = acc[2] > 0.0 ? acc[2] : 1.0;
  const double mean = acc[0] / cnt;
  double var = acc[1] / cnt - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const double sd = fmax(sqrt(var), 0.00001);                       // np.max([std, 1e-5])
  const float fill = (float)(((double)ord2f(*minord) - mean) / sd);   // z-scoring is monotone: min(z) = z(min)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float z = (float)(((double)x[i] - mean) / sd);
    x[i] = (fill_outside && mask && !(mask[i] >= 0.5f)) ? fill : z;   // np.where(m < 0.5, tissue_min, x)
  }
}

// ---------------------------------------------------------------------------------------------------------------
// mean_std_grid.  The blocks of the reference are patch-sized and half-overlapping, i.e. every block is the union of
// (up to) 2x2x2 step-sized cells: one CTA reduces one cell to {sum, sumsq, count} (the volume is read exactly once),
// a tiny kernel combines the cells into the zero-bordered mean / std grids, and the zoom kernel evaluates scipy's
// order-1 `zoom` of those grids at the voxels that survive the reference's cropping.
// ---------------------------------------------------------------------------------------------------------------
struct GridGeom {
  int X, Y, Z;          // data shape
  int st[3];            // step = patch / 2 (patch rounded up to even)
  int pad[3];           // padded shape (multiple of the patch)
  int g[3];             // cells per axis = pad / st = grid shape
};

__global__ void __launch_bounds__(256) s1_cell_sums_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                           GridGeom q, double* __restrict__ cells) {
  const int cz = blockIdx.x % q.g[2], cy = (blockIdx.x / q.g[2]) % q.g[1], cx = blockIdx.x / (q.g[2] * q.g[1]);
  const int x0 = cx * q.st[0], y0 = cy * q.st[1], z0 = cz * q.st[2];
  const int nx = max(0, min(q.st[0], q.X - x0)), ny = max(0, min(q.st[1], q.Y - y0)), nz = max(0, min(q.st[2], q.Z - z0));
  double s = 0.0, ss = 0.0, cnt = 0.0;
  const int rows = nx * ny;
  // a warp walks one z-row at a time: consecutive lanes read consecutive floats
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int r = w; r < rows; r += 8) {
    const int64_t base = ((int64_t)(x0 + r / ny) * q.Y + (y0 + r % ny)) * q.Z + z0;
    for (int k = l; k < nz; k += 32) {
      const float v = x[base + k];
      if (!mask || mask[base + k] > 0.5f) { s += v; ss += (double)v * v; cnt += 1.0; }
    }
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss); cnt = warp_sum_d(cnt);
  __shared__ double sh[3][8];
  if (l == 0) { sh[0][w] = s; sh[1][w] = ss; sh[2][w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { s += sh[0][i]; ss += sh[1][i]; cnt += sh[2][i]; }
    double* o = cells + (size_t)blockIdx.x * 3;
    o[0] = s; o[1] = ss; o[2] = cnt;
  }
}

// grids: [g0+2][g1+2][g2+2] doubles, borders zero (memset by the host)
__global__ void s1_grid_stats_kernel(const double* __restrict__ cells, GridGeom q, int masked,
                                     double* __restrict__ mean_grid, double* __restrict__ std_grid) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= q.g[0] * q.g[1] * q.g[2]) return;
  const int k = t % q.g[2], j = (t / q.g[2]) % q.g[1], i = t / (q.g[2] * q.g[1]);
  double s = 0.0, ss = 0.0, cnt = 0.0;
  for (int a = i; a < min(i + 2, q.g[0]); ++a)                       // the last block is clipped by the padded shape
    for (int b = j; b < min(j + 2, q.g[1]); ++b)
      for (int c = k; c < min(k + 2, q.g[2]); ++c) {
        const double* p = cells + ((size_t)(a * q.g[1] + b) * q.g[2] + c) * 3;
        s += p[0]; ss += p[1]; cnt += p[2];
      }
  double mu, sd;
  if (masked) {
    if (cnt > 0.0) { mu = s / cnt; const double v = ss / cnt - mu * mu; sd = sqrt(v > 0.0 ? v : 0.0); }
    else { mu = 0.0; sd = 0.00001; }
  } else {                                                           // zero padding counts as data
    const double nb = (double)(min(i + 2, q.g[0]) - i) * q.st[0] * (double)(min(j + 2, q.g[1]) - j) * q.st[1] *
                      (double)(min(k + 2, q.g[2]) - k) * q.st[2];
    mu = s / nb; const double v = ss / nb - mu * mu; sd = fmax(sqrt(v > 0.0 ? v : 0.0), 0.00001);
  }
  const size_t o = ((size_t)(i + 1) * (q.g[1] + 2) + (j + 1)) * (q.g[2] + 2) + (k + 1);
  mean_grid[o] = mu; std_grid[o] = sd;
}

// scipy.ndimage.zoom(grid, step, order=1): output index o <-> input coordinate o * (in - 1) / (out - 1)
__device__ __forceinline__ void zoom_coord(int v, int st, int g, int& i0, double& f) {
  const int in = g + 2, out = in * st;
  const double c = (double)(v + st / 2) * ((double)(in - 1) / (double)(out - 1));
  i0 = (int)floor(c);
  f = c - (double)i0;
  if (i0 >= in - 1) { i0 = in - 2; f = 1.0; }
}

__global__ void __launch_bounds__(256) s1_grid_zoom_kernel(const double* __restrict__ mean_grid, const double* __restrict__ std_grid,
                                                           GridGeom q, float* __restrict__ mean_out, float* __restrict__ std_out) {
  const int64_t V = (int64_t)q.X * q.Y * q.Z;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int G1 = q.g[1] + 2, G2 = q.g[2] + 2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < V; t += stride) {
    const int z = (int)(t % q.Z), y = (int)((t / q.Z) % q.Y), x = (int)(t / ((int64_t)q.Z * q.Y));
    int i0, j0, k0; double fx, fy, fz;
    zoom_coord(x, q.st[0], q.g[0], i0, fx); zoom_coord(y, q.st[1], q.g[1], j0, fy); zoom_coord(z, q.st[2], q.g[2], k0, fz);
    const size_t o = ((size_t)i0 * G1 + j0) * G2 + k0;
    const double w[2][2][2] = {{{(1 - fx) * (1 - fy) * (1 - fz), (1 - fx) * (1 - fy) * fz}, {(1 - fx) * fy * (1 - fz), (1 - fx) * fy * fz}},
                               {{fx * (1 - fy) * (1 - fz), fx * (1 - fy) * fz}, {fx * fy * (1 - fz), fx * fy * fz}}};
    double m = 0.0, s = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const size_t p = o + ((size_t)a * G1 + b) * G2 + c;
          m += w[a][b][c] * mean_grid[p];
          if (std_out) s += w[a][b][c] * std_grid[p];
        }
    mean_out[t] = (float)m;
    if (std_out) std_out[t] = (float)s;
  }
}

__global__ void __launch_bounds__(256) s1_align_kernel(float* __restrict__ x, const float* __restrict__ mu_i,
                                                       const float* __restrict__ mu_p, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = (float)(((double)x[i] - (double)mu_i[i]) + (double)mu_p[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// group mean / population std over the K reference volumes + NLL of the target, one pass: (K + 1) x 4 B read,
// 4 .. 12 B written per voxel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int S1_MAX_REFS = 32;
struct RefPtrs { const float* p[S1_MAX_REFS]; };

__global__ void __launch_bounds__(256) s1_group_nll_kernel(const float* __restrict__ xp, RefPtrs refs, int K, double min_std, int side,
                                                           const float* __restrict__ mul_mask, float* __restrict__ anomaly,
                                                           float* __restrict__ mu_out, float* __restrict__ sigma_out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += (double)__ldg(refs.p[k] + i);
    const double mu = s / K;
    double v = 0.0;
    for (int k = 0; k < K; ++k) { const double d = (double)__ldg(refs.p[k] + i) - mu; v += d * d; }   // second read hits L1/L2
    double sg = sqrt(v / K);
    sg = min_std < 0.0 ? sg + 1e-6 : (sg < min_std ? min_std : sg);
    const double x = (double)xp[i];
    double a = (x - mu) * (x - mu) / (2.0 * sg * sg) + log(sg * 2.506);
    if (a != a) a = 0.0;                                              // np.nan_to_num(nan=0.0)
    if (side > 0) a = x > mu ? a : 0.0;
    else if (side < 0) a = x < mu ? a : 0.0;
    if (mul_mask) a *= (double)mul_mask[i];
    if (anomaly) anomaly[i] = (float)a;
    if (mu_out) mu_out[i] = (float)mu;
    if (sigma_out) sigma_out[i] = (float)sg;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// median filter, scipy.ndimage.median_filter(size=(kx,ky,kz), mode='constant', cval=0): window [i - k/2, i - k/2 + k)
// per axis, rank (kx ky kz) / 2 of the ascending window.  One thread per voxel, CTA tile 2 x 4 x 32 (z fastest) staged
// with its halo in shared memory as order-preserving integer keys; the rank is found by a 32-step bitwise search on
// the key (count of window keys below the probe), so any window size costs 32 N shared-memory reads and no sort.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MT_X = 2, MT_Y = 4, MT_Z = 32;

__device__ __forceinline__ uint32_t f2key(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(256) s1_median_kernel(const float* __restrict__ in, float* __restrict__ out, int X, int Y, int Z,
                                                        int kx, int ky, int kz) {
  extern __shared__ uint32_t tile[];
  const int tx = MT_X + kx - 1, ty = MT_Y + ky - 1, tz = MT_Z + kz - 1;
  const int bz = blockIdx.x * MT_Z, by = blockIdx.y * MT_Y, bx = blockIdx.z * MT_X;
  const int ox = bx - kx / 2, oy = by - ky / 2, oz = bz - kz / 2;
  for (int t = threadIdx.x; t < tx * ty * tz; t += blockDim.x) {
    const int c = t % tz, b = (t / tz) % ty, a = t / (tz * ty);
    const int gx = ox + a, gy = oy + b, gz = oz + c;
    float v = 0.f;                                                   // cval
    if (gx >= 0 && gx < X && gy >= 0 && gy < Y && gz >= 0 && gz < Z) v = in[((int64_t)gx * Y + gy) * Z + gz];
    tile[t] = f2key(v);
  }
  __syncthreads();
  const int lz = threadIdx.x % MT_Z, ly = (threadIdx.x / MT_Z) % MT_Y, lx = threadIdx.x / (MT_Z * MT_Y);
  const int gx = bx + lx, gy = by + ly, gz = bz + lz;
  if (gx >= X || gy >= Y || gz >= Z) return;
  const int rank = (kx * ky * kz) / 2;
  uint32_t key = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t probe = key | (1u << bit);
    int below = 0;
    for (int a = 0; a < kx; ++a)
      for (int b = 0; b < ky; ++b) {
        const uint32_t* row = tile + ((lx + a) * ty + (ly + b)) * tz + lz;
        for (int c = 0; c < kz; ++c) below += row[c] < probe ? 1 : 0;
      }
    if (below <= rank) key = probe;                                  // largest key with (#window keys below it) <= rank
  }
  out[((int64_t)gx * Y + gy) * Z + gz] = key2f(key);
}

int geom(int X, int Y, int Z, const int32_t patch[3], GridGeom* q) {
  if (X <= 0 || Y <= 0 || Z <= 0) return fail("mean_std_grid: empty volume");
  q->X = X; q->Y = Y; q->Z = Z;
  const int sh[3] = {X, Y, Z};
  for (int a = 0; a < 3; ++a) {
    if (patch[a] <= 0) return fail("mean_std_grid: patch_size[%d] = %d", a, patch[a]);
    const int ps = 2 * ((patch[a] + 1) / 2);                         // 2 * ceil(p / 2)
    q->st[a] = ps / 2;
    q->pad[a] = ps * ((sh[a] + ps - 1) / ps);
    q->g[a] = q->pad[a] / q->st[a];
  }
  return 0;
}
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" int dwmh_s1_zscore(int32_t device, float* x, const float* mask, int64_t n, int32_t fill_outside, void* workspace,
                              double* stats_out, void* stream_) {
  if (!x || !workspace) return fail("dwmh_s1_zscore: null argument");
  if (n <= 0) return fail("dwmh_s1_zscore: empty volume");
  if (fill_outside && !mask) return fail("dwmh_s1_zscore: fill_outside needs a mask");
  cudaStream_t st = (cudaStream_t)stream_;
  S1_CU(cudaSetDevice(device));
  double* acc = (double*)workspace;
  int* mn = (int*)((char*)workspace + 32);
  S1_CU(cudaMemsetAsync(workspace, 0, 32, st));
  S1_CU(cudaMemsetAsync(mn, 0x7f, 4, st));                           // 0x7f7f7f7f: above every finite float's key
  const int grid = grid_for(device);
  s1_stats_kernel<<<grid, 256, 0, st>>>(x, mask, n, acc, mn);
  s1_zscore_apply_kernel<<<grid, 256, 0, st>>>(x, mask, n, acc, mn, fill_outside);
  S1_CU(cudaGetLastError());
  if (stats_out) {
    double h[4];
    S1_CU(cudaMemcpyAsync(h, acc, sizeof h, cudaMemcpyDeviceToHost, st));
    S1_CU(cudaStreamSynchronize(st));
    const double cnt = h[2] > 0 ? h[2] : 1.0, m = h[0] / cnt;
    double var = h[1] / cnt - m * m; if (var < 0) var = 0;
    stats_out[0] = m; stats_out[1] = sqrt(var); stats_out[2] = h[2];
  }
  return 0;
}

extern "C" int dwmh_s1_mean_std_grid_workspace(int32_t X, int32_t Y, int32_t Z, const int32_t patch_size[3], int64_t* bytes) {
  GridGeom q;
  if (!patch_size || !bytes) return fail("dwmh_s1_mean_std_grid_workspace: null argument");
  if (geom(X, Y, Z, patch_size, &q)) return 1;
  const size_t cells = (size_t)q.g[0] * q.g[1] * q.g[2], grid = (size_t)(q.g[0] + 2) * (q.g[1] + 2) * (q.g[2] + 2);
  *bytes = (int64_t)(align256(cells * 3 * sizeof(double)) + 2 * align256(grid * sizeof(double)));
  return 0;
}

extern "C" int dwmh_s1_mean_std_grid(int32_t device, const float* x, const float* mask, int32_t X, int32_t Y, int32_t Z,
                                     const int32_t patch_size[3], float* mean_out, float* std_out, void* workspace, void* stream_) {
  if (!x || !mean_out || !workspace || !patch_size) return fail("dwmh_s1_mean_std_grid: null argument");
  GridGeom q;
  if (geom(X, Y, Z, patch_size, &q)) return 1;
  cudaStream_t st = (cudaStream_t)stream_;
  S1_CU(cudaSetDevice(device));
  const size_t ncell = (size_t)q.g[0] * q.g[1] * q.g[2], ngrid = (size_t)(q.g[0] + 2) * (q.g[1] + 2) * (q.g[2] + 2);
  double* cells = (double*)workspace;
  double* mg = (double*)((char*)workspace + align256(ncell * 3 * sizeof(double)));
  double* sg = (double*)((char*)mg + align256(ngrid * sizeof(double)));
  S1_CU(cudaMemsetAsync(mg, 0, 2 * align256(ngrid * sizeof(double)), st));
  s1_cell_sums_kernel<<<(unsigned)ncell, 256, 0, st>>>(x, mask, q, cells);
  s1_grid_stats_kernel<<<(unsigned)((ncell + 127) / 128), 128, 0, st>>>(cells, q, mask ? 1 : 0, mg, sg);
  s1_grid_zoom_kernel<<<grid_for(device), 256, 0, st>>>(mg, sg, q, mean_out, std_out);
  S1_CU(cudaGetLastError());
  return 0;
}

extern "C" int dwmh_s1_align_local_mean(int32_t device, float* x, const float* local_mu, const float* target_local_mu, int64_t n, void* stream_) {
  if (!x || !local_mu || !target_local_mu) return fail("dwmh_s1_align_local_mean: null argument");
  S1_CU(cudaSetDevice(device));
  s1_align_kernel<<<grid_for(device), 256, 0, (cudaStream_t)stream_>>>(x, local_mu, target_local_mu, n);
  S1_CU(cudaGetLastError());
  return 0;
}

extern "C" int dwmh_s1_group_nll(int32_t device, const float* x_prime, const float* const* refs, int32_t k, double min_std, int32_t side,
                                 const float* mul_mask, float* anomaly, float* mu_out, float* sigma_out, int64_t n, void* stream_) {
  if (!x_prime || !refs) return fail("dwmh_s1_group_nll: null argument");
  if (k <= 0 || k > S1_MAX_REFS) return fail("dwmh_s1_group_nll: k = %d reference images (1..%d supported)", k, S1_MAX_REFS);
  if (side < -1 || side > 1) return fail("dwmh_s1_group_nll: side must be -1, 0 or +1");
  RefPtrs rp{};
  for (int i = 0; i < k; ++i) { if (!refs[i]) return fail("dwmh_s1_group_nll: refs[%d] is null", i); rp.p[i] = refs[i]; }
  S1_CU(cudaSetDevice(device));
  s1_group_nll_kernel<<<grid_for(device), 256, 0, (cudaStream_t)stream_>>>(x_prime, rp, k, min_std, side, mul_mask, anomaly, mu_out, sigma_out, n);
  S1_CU(cudaGetLastError());
  return 0;
}

extern "C" int dwmh_s1_median_filter(int32_t device, const float* in, float* out, int32_t X, int32_t Y, int32_t Z,
                                     const int32_t kernel_size[3], void* stream_) {
  if (!in || !out || !kernel_size) return fail("dwmh_s1_median_filter: null argument");
  if (in == out) return fail("dwmh_s1_median_filter: in and out must not alias");
  if (X <= 0 || Y <= 0 || Z <= 0) return fail("dwmh_s1_median_filter: empty volume");
  const int kx = kernel_size[0], ky = kernel_size[1], kz = kernel_size[2];
  if (kx < 1 || ky < 1 || kz < 1 || kx > 9 || ky > 9 || kz > 9) return fail("dwmh_s1_median_filter: kernel %dx%dx%d (1..9 per axis supported)", kx, ky, kz);
  S1_CU(cudaSetDevice(device));
  const size_t smem = (size_t)(MT_X + kx - 1) * (MT_Y + ky - 1) * (MT_Z + kz - 1) * sizeof(uint32_t);
  dim3 grid((Z + MT_Z - 1) / MT_Z, (Y + MT_Y - 1) / MT_Y, (X + MT_X - 1) / MT_X);
  if (grid.y > 65535 || grid.z > 65535) return fail("dwmh_s1_median_filter: volume too large");
  s1_median_kernel<<<grid, 256, smem, (cudaStream_t)stream_>>>(in, out, X, Y, Z, kx, ky, kz);
  S1_CU(cudaGetLastError());
  return 0;
}
