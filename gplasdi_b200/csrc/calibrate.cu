// calibrate.cu -- SINDy calibration: SBP finite-difference dZ/dt and the per-series least squares.
//
// Replaces SINDy.compute_time_derivative / SINDy.calibrate (reference
// src/lasdi/latent_dynamics/sindy.py:41-102) and the sparse operators of src/lasdi/fd.py:14-217.
//
// The reference applies a (T x T) sparse COO matrix; here the operator is what it is -- a banded
// stencil with dense boundary closures -- applied as a coalesced streaming kernel (HBM-bound:
// 4 B read + 4 B written per element).  The fit accumulates the (n+1)x(n+1) Gram matrix and the
// (n+1)xn right-hand side in fp64 straight from the series held in shared memory (Z is read from
// HBM exactly once), solves by Cholesky in fp64 and rounds to fp32: for cond([1|Z]) <~ 1e3 this
// matches LAPACK gelsy on fp32 data to ~4e-7 (SURVEY.md §7 "lstsq semantics").
#include "common.cuh"
#include "sbp_tables.h"

namespace glasdi {

struct SbpF {
  int halfwidth, depth;
  int widths[SBP_MAXD];
  float interior[SBP_MAXHW];
  float rows[SBP_MAXD][SBP_MAXW];
};
__constant__ SbpF c_sbp[SBP_NSCHEMES];
static bool g_sbp_uploaded[64] = {false};

static int upload_sbp(glasdi_handle* h) {
  if (h->device < 64 && g_sbp_uploaded[h->device]) return GLASDI_OK;
  SbpF host[SBP_NSCHEMES];
  for (int s = 0; s < SBP_NSCHEMES; ++s) {
    host[s].halfwidth = kSbpSchemes[s].halfwidth;
    host[s].depth = kSbpSchemes[s].depth;
    for (int r = 0; r < SBP_MAXD; ++r) {
      host[s].widths[r] = kSbpSchemes[s].widths[r];
      for (int c = 0; c < SBP_MAXW; ++c) host[s].rows[r][c] = (float)kSbpSchemes[s].rows[r][c];   // fd.py:54 fp32 cast
    }
    for (int k = 0; k < SBP_MAXHW; ++k) host[s].interior[k] = (float)kSbpSchemes[s].interior[k];
  }
  GL_CUDA(h, cudaMemcpyToSymbol(c_sbp, host, sizeof(host)));
  if (h->device < 64) g_sbp_uploaded[h->device] = true;
  return GLASDI_OK;
}

// derivative of component i at time row t of one series; z points at element [0][i], row stride = n
// accumulation in ascending column order, fp32 FMA (the reference's sparse.mm is an fp32 row sum).
template <typename LoadF>
__device__ __forceinline__ float sbp_row(const SbpF& sc, int t, int T, LoadF ld) {
  float acc = 0.f;
  if (t < sc.depth) {
    const int w = sc.widths[t];
    for (int c = 0; c < w; ++c) acc = fmaf(sc.rows[t][c], ld(c), acc);
  } else if (t >= T - sc.depth) {
    const int r = T - 1 - t;
    const int w = sc.widths[r];
    for (int c = w - 1; c >= 0; --c) acc = fmaf(-sc.rows[r][c], ld(T - 1 - c), acc);
  } else {
    for (int k = sc.halfwidth; k >= 1; --k) acc = fmaf(-sc.interior[k - 1], ld(t - k), acc);
    for (int k = 1; k <= sc.halfwidth; ++k) acc = fmaf(sc.interior[k - 1], ld(t + k), acc);
  }
  return acc;
}

__global__ void __launch_bounds__(256) fd_dzdt_kernel(const float* __restrict__ Z, int T, int n, int fd_id,
                                                      float inv_dt, float* __restrict__ dZ) {
  const SbpF& sc = c_sbp[fd_id];
  const size_t base = (size_t)blockIdx.y * T * n;
  const float* __restrict__ z = Z + base;
  const int total = T * n;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
    const int t = f / n, i = f - t * n;
    const float v = sbp_row(sc, t, T, [&](int row) { return __ldg(z + row * n + i); });
    dZ[base + f] = inv_dt * v;
  }
}

constexpr int kFitThreads = 256;

template <int N>
__global__ void __launch_bounds__(kFitThreads) sindy_fit_kernel(const float* __restrict__ Z, int T, int fd_id,
                                                                float inv_dt, int norm_order,
                                                                float* __restrict__ coefs_out,
                                                                float* __restrict__ loss_sindy_out,
                                                                float* __restrict__ loss_coef_out,
                                                                int32_t* __restrict__ info_out, int use_smem) {
  constexpr int NA = N + 1;                    // library size [1 | z]
  constexpr int NG = NA * (NA + 1) / 2;        // upper triangle of the Gram matrix
  constexpr int NR = NA * N;                   // right-hand side
  constexpr int NS = NG + NR;
  extern __shared__ float zs[];                // series [T][N] when it fits
  __shared__ double red[kFitThreads / 32][NS];
  __shared__ double sol[NA][N];
  __shared__ float solf[NA][N];
  __shared__ double red2[kFitThreads / 32];
  __shared__ int s_info;

  const SbpF& sc = c_sbp[fd_id];
  const int b = blockIdx.x;
  const float* __restrict__ zg = Z + (size_t)b * T * N;
  const float* z = zg;
  if (use_smem) {
    for (int f = threadIdx.x; f < T * N; f += blockDim.x) zs[f] = zg[f];
    __syncthreads();
    z = zs;
  }

  double acc[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) acc[q] = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float x[NA], d[N];
    x[0] = 1.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      x[1 + i] = z[t * N + i];
      d[i] = inv_dt * sbp_row(sc, t, T, [&](int row) { return z[row * N + i]; });
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = a; c < NA; ++c) acc[q++] += (double)x[a] * (double)x[c];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int i = 0; i < N; ++i) acc[q++] += (double)x[a] * (double)d[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double G[NA][NA], R[NA][N];
    int q = 0;
    for (int a = 0; a < NA; ++a)
      for (int c = a; c < NA; ++c) {
        double v = 0.0;
        for (int w = 0; w < kFitThreads / 32; ++w) v += red[w][q];
        G[a][c] = v; G[c][a] = v;
        ++q;
      }
    for (int a = 0; a < NA; ++a)
      for (int i = 0; i < N; ++i) {
        double v = 0.0;
        for (int w = 0; w < kFitThreads / 32; ++w) v += red[w][q];
        R[a][i] = v;
        ++q;
      }
    // Cholesky G = L L^T (in place, lower), with a relative pivot test
    int info = 0;
    double dmax = 0.0;
    for (int a = 0; a < NA; ++a) dmax = fmax(dmax, G[a][a]);
    for (int j = 0; j < NA && info == 0; ++j) {
      double s = G[j][j];
      for (int k = 0; k < j; ++k) s -= G[j][k] * G[j][k];
      if (!(s > 1e-13 * dmax)) { info = j + 1; break; }
      const double ljj = sqrt(s);
      G[j][j] = ljj;
      for (int i = j + 1; i < NA; ++i) {
        double v = G[i][j];
        for (int k = 0; k < j; ++k) v -= G[i][k] * G[j][k];
        G[i][j] = v / ljj;
      }
    }
    if (info == 0) {
      for (int c = 0; c < N; ++c) {
        double y[NA];
        for (int i = 0; i < NA; ++i) {
          double v = R[i][c];
          for (int k = 0; k < i; ++k) v -= G[i][k] * y[k];
          y[i] = v / G[i][i];
        }
        for (int i = NA - 1; i >= 0; --i) {
          double v = y[i];
          for (int k = i + 1; k < NA; ++k) v -= G[k][i] * sol[k][c];
          sol[i][c] = v / G[i][i];
        }
      }
    } else {
      for (int a = 0; a < NA; ++a)
        for (int c = 0; c < N; ++c) sol[a][c] = nan("");
    }
    double nrm = 0.0;
    for (int a = 0; a < NA; ++a)
      for (int c = 0; c < N; ++c) {
        const float cf = (float)sol[a][c];
        solf[a][c] = cf;
        coefs_out[(size_t)b * NR + a * N + c] = cf;     // row-major flatten of (n+1, n), sindy.py:80
        nrm += (norm_order == 1) ? fabs((double)cf) : (double)cf * (double)cf;
      }
    if (loss_coef_out) loss_coef_out[b] = (float)((norm_order == 1) ? nrm : sqrt(nrm));
    s_info = info;
    if (info_out) info_out[b] = info;
  }
  __syncthreads();
  if (!loss_sindy_out) return;
  // second pass: MSE(dZdt, [1|Z] R) with fp32 coefficients, sindy.py:75
  double se = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float d = inv_dt * sbp_row(sc, t, T, [&](int row) { return z[row * N + i]; });
      float pred = solf[0][i];
#pragma unroll
      for (int j = 0; j < N; ++j) pred = fmaf(z[t * N + j], solf[1 + j][i], pred);
      const float e = d - pred;
      se += (double)e * (double)e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if (lane == 0) red2[warp] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < kFitThreads / 32; ++w) v += red2[w];
    loss_sindy_out[b] = (float)(v / ((double)T * N));
  }
}

int launch_fd_dzdt(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt, float* dZ,
                   cudaStream_t st) {
  if (fd_id < 0 || fd_id >= SBP_NSCHEMES) return set_error(h, GLASDI_ERR_INVALID, "unknown fd scheme id");
  if (B <= 0 || n <= 0) return set_error(h, GLASDI_ERR_INVALID, "fd_dzdt: empty batch");
  if (T < 2 * kSbpSchemes[fd_id].depth) return set_error(h, GLASDI_ERR_INVALID, "fd_dzdt: series shorter than 2*boundary depth");
  if (B > 65535) return set_error(h, GLASDI_ERR_INVALID, "fd_dzdt: B > 65535");
  int rc = upload_sbp(h);
  if (rc) return rc;
  const int total = T * n;
  dim3 grid((total + 255) / 256, B);
  fd_dzdt_kernel<<<grid, 256, 0, st>>>(Z, T, n, fd_id, (float)(1.0 / dt), dZ);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "fd_dzdt launch");
}

template <int N>
static int launch_fit_n(glasdi_handle* h, const float* Z, int B, int T, int fd_id, double dt, int norm_order,
                        float* coefs, float* ls, float* lc, int32_t* info, cudaStream_t st) {
  size_t smem = (size_t)T * N * sizeof(float);
  int use_smem = smem <= 160 * 1024;
  if (!use_smem) smem = 0;
  if (smem > 40 * 1024)
    GL_CUDA(h, cudaFuncSetAttribute(sindy_fit_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sindy_fit_kernel<N><<<B, kFitThreads, smem, st>>>(Z, T, fd_id, (float)(1.0 / dt), norm_order, coefs, ls, lc, info,
                                                   use_smem);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "sindy_fit launch");
}

int launch_sindy_fit(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt, int norm_order,
                     float* coefs, float* ls, float* lc, int32_t* info, cudaStream_t st) {
  if (fd_id < 0 || fd_id >= SBP_NSCHEMES) return set_error(h, GLASDI_ERR_INVALID, "unknown fd scheme id");
  if (B <= 0) return set_error(h, GLASDI_ERR_INVALID, "sindy_fit: empty batch");
  if (T < 2 * kSbpSchemes[fd_id].depth) return set_error(h, GLASDI_ERR_INVALID, "sindy_fit: series shorter than 2*boundary depth");
  if (norm_order != 1 && norm_order != 2) return set_error(h, GLASDI_ERR_INVALID, "sindy_fit: norm_order must be 1 or 2 ('fro')");
  int rc = upload_sbp(h);
  if (rc) return rc;
  switch (n) {
    case 1: return launch_fit_n<1>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    case 2: return launch_fit_n<2>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    case 3: return launch_fit_n<3>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    case 4: return launch_fit_n<4>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    case 5: return launch_fit_n<5>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    case 6: return launch_fit_n<6>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    case 7: return launch_fit_n<7>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    case 8: return launch_fit_n<8>(h, Z, B, T, fd_id, dt, norm_order, coefs, ls, lc, info, st);
    default: return set_error(h, GLASDI_ERR_INVALID, "sindy_fit: latent dimension must be 1..8");
  }
}

}  // namespace glasdi
