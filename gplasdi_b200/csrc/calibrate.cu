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

// ---- rank-revealing solve of the normal equations = LAPACK xGELSY on [1|Z] --------------------------------------
// The reference solves `torch.linalg.lstsq(Z_i, dZdt)` (sindy.py:72) = LAPACK sgelsy with rcond = eps_fp32 * max(m, n):
// QR with column pivoting, effective rank from an incremental condition estimate of the leading triangle (xLAIC1), and
// the MINIMUM-NORM solution of the rank-truncated problem.  An untrained encoder produces series with cond([1|Z]) of
// 1e4..1e6, beyond 1/rcond = 8.4e3 at T = 1001: those are truncated by the reference, not solved.
// Here the same algorithm runs on the fp64 Gram matrix G = X^T X: Cholesky with diagonal pivoting of G IS the R factor
// of the column-pivoted QR of X (same pivot rule: largest remaining column norm), Q^T Y[:r] = R11^-T (P^T X^T Y)[:r], and
// the minimum-norm solution of R1 y = c, R1 = [R11 R12], is y = R1^T (R1 R1^T)^-1 c.  fp64 on G resolves R's diagonal to
// 1e-16 cond(X)^2: exact enough for a decision at 1e-4.
struct GelsyWork {   // per series, in shared memory; NA <= 9
  double M[9][9];    // lower Cholesky factor of R1 R1^T (rank < NA only)
  int perm[9];
  int rank;
};

// xLAIC1: one step of incremental condition estimation.  x (length j) estimates the extreme singular vector of the
// leading j x j triangle with singular value `sest`; [w; gamma] is the next column.  Returns sestpr and the rotation (s, c).
__host__ __device__ inline void laic1(bool want_max, int j, const double* x, double sest, const double* w, double gamma, double& sestpr,
                      double& s, double& c) {
  const double eps = 2.220446049250313e-16;
  double alpha = 0.0;
  for (int i = 0; i < j; ++i) alpha += x[i] * w[i];
  const double absalp = fabs(alpha), absgam = fabs(gamma), absest = fabs(sest);
  if (want_max) {
    if (sest == 0.0) {
      const double s1 = fmax(absgam, absalp);
      if (s1 == 0.0) { s = 0.0; c = 1.0; sestpr = 0.0; }
      else { s = alpha / s1; c = gamma / s1; const double tmp = sqrt(s * s + c * c); s /= tmp; c /= tmp; sestpr = s1 * tmp; }
    } else if (absgam <= eps * absest) {
      s = 1.0; c = 0.0;
      const double tmp = fmax(absest, absalp), s1 = absest / tmp, s2 = absalp / tmp;
      sestpr = tmp * sqrt(s1 * s1 + s2 * s2);
    } else if (absalp <= eps * absest) {
      if (absgam <= absest) { s = 1.0; c = 0.0; sestpr = absest; } else { s = 0.0; c = 1.0; sestpr = absgam; }
    } else if (absest <= eps * absalp || absest <= eps * absgam) {
      if (absgam <= absalp) {
        const double tmp = absgam / absalp; s = sqrt(1.0 + tmp * tmp); sestpr = absalp * s;
        c = (gamma / absalp) / s; s = copysign(1.0, alpha) / s;
      } else {
        const double tmp = absalp / absgam; c = sqrt(1.0 + tmp * tmp); sestpr = absgam * c;
        s = (alpha / absgam) / c; c = copysign(1.0, gamma) / c;
      }
    } else {
      const double z1 = alpha / absest, z2 = gamma / absest;
      const double bb = (1.0 - z1 * z1 - z2 * z2) * 0.5, cc = z1 * z1;
      const double t = bb > 0.0 ? cc / (bb + sqrt(bb * bb + cc)) : sqrt(bb * bb + cc) - bb;
      const double sine = -z1 / t, cosine = -z2 / (1.0 + t), tmp = sqrt(sine * sine + cosine * cosine);
      s = sine / tmp; c = cosine / tmp; sestpr = sqrt(t + 1.0) * absest;
    }
  } else {
    if (sest == 0.0) {
      sestpr = 0.0;
      double sine, cosine;
      if (fmax(absgam, absalp) == 0.0) { sine = 1.0; cosine = 0.0; } else { sine = -gamma; cosine = alpha; }
      const double s1 = fmax(fabs(sine), fabs(cosine));
      s = sine / s1; c = cosine / s1;
      const double tmp = sqrt(s * s + c * c); s /= tmp; c /= tmp;
    } else if (absgam <= eps * absest) {
      s = 0.0; c = 1.0; sestpr = absgam;
    } else if (absalp <= eps * absest) {
      if (absgam <= absest) { s = 0.0; c = 1.0; sestpr = absgam; } else { s = 1.0; c = 0.0; sestpr = absest; }
    } else if (absest <= eps * absalp || absest <= eps * absgam) {
      if (absgam <= absalp) {
        const double tmp = absgam / absalp; c = sqrt(1.0 + tmp * tmp); sestpr = absest * (tmp / c);
        s = -(gamma / absalp) / c; c = copysign(1.0, alpha) / c;
      } else {
        const double tmp = absalp / absgam; s = sqrt(1.0 + tmp * tmp); sestpr = absest / s;
        c = (alpha / absgam) / s; s = -copysign(1.0, gamma) / s;
      }
    } else {
      const double z1 = alpha / absest, z2 = gamma / absest;
      const double norma = fmax(1.0 + z1 * z1 + fabs(z1 * z2), fabs(z1 * z2) + z2 * z2);
      const double test = 1.0 + 2.0 * (z1 - z2) * (z1 + z2);
      double sine, cosine;
      if (test >= 0.0) {
        const double bb = (z1 * z1 + z2 * z2 + 1.0) * 0.5, cc = z2 * z2;
        const double t = cc / (bb + sqrt(fabs(bb * bb - cc)));
        sine = z1 / (1.0 - t); cosine = -z2 / t;
        sestpr = sqrt(t + 4.0 * eps * eps * norma) * absest;
      } else {
        const double bb = (z2 * z2 + z1 * z1 - 1.0) * 0.5, cc = z1 * z1;
        const double t = bb >= 0.0 ? -cc / (bb + sqrt(bb * bb + cc)) : bb - sqrt(bb * bb + cc);
        sine = -z1 / t; cosine = -z2 / (1.0 + t);
        sestpr = sqrt(1.0 + t + 4.0 * eps * eps * norma) * absest;
      }
      const double tmp = sqrt(sine * sine + cosine * cosine);
      s = sine / tmp; c = cosine / tmp;
    }
  }
}

// In: G symmetric (full storage).  Out: the upper triangle of G holds R (P^T G P = R^T R), W.perm / W.rank / W.M filled.
// Single thread; rolled loops.  Returns the effective rank (0 if G is not finite).
template <int NA>
__host__ __device__ int gelsy_factor(double (*G)[NA], GelsyWork& W, double rcond) {
  for (int j = 0; j < NA; ++j) W.perm[j] = j;
  double chk = 0.0;                                // NaN if any entry of G is inf / NaN (comparisons would skip them)
  for (int i = 0; i < NA; ++i)
    for (int k = 0; k < NA; ++k) chk = fma(G[i][k], 0.0, chk);
  if (chk != 0.0) { W.rank = 0; return 0; }
  int nr = NA;                                     // rows of R that exist (exactly singular remainder stops earlier)
#pragma unroll 1
  for (int j = 0; j < NA; ++j) {
    int pv = j;
    double dm = G[j][j];
    for (int k = j + 1; k < NA; ++k) if (G[k][k] > dm) { dm = G[k][k]; pv = k; }
    if (!(dm > 0.0) || !(dm <= 1.79e308)) { nr = j; break; }
    if (pv != j) {
      for (int i = 0; i < NA; ++i) { const double t = G[i][j]; G[i][j] = G[i][pv]; G[i][pv] = t; }
      for (int k = 0; k < NA; ++k) { const double t = G[j][k]; G[j][k] = G[pv][k]; G[pv][k] = t; }
      const int t = W.perm[j]; W.perm[j] = W.perm[pv]; W.perm[pv] = t;
    }
    const double rjj = sqrt(G[j][j]), inv = 1.0 / rjj;
    G[j][j] = rjj;
    for (int k = j + 1; k < NA; ++k) G[j][k] *= inv;
    for (int i = j + 1; i < NA; ++i)
      for (int k = i; k < NA; ++k) { const double v = G[i][k] - G[j][i] * G[j][k]; G[i][k] = v; G[k][i] = v; }
  }
  // effective rank (xGELSY): grow the leading triangle while smax * rcond <= smin (incremental estimates)
  int rank = 0;
  // Shortcut for the well-conditioned series of a trained model: the estimates satisfy smin_est >= sigma_min and
  // smax_est <= sigma_max <= |R|_F, and sigma_min(R) >= |det R| ((n - 1) / |R|_F^2)^((n - 1) / 2); when that lower bound
  // already clears rcond |R|_F, xLAIC1 would accept every column (the leading triangles are better conditioned still).
  if (nr == NA) {
    double det = 1.0, fro2 = 0.0;
    for (int i = 0; i < NA; ++i) {
      det *= G[i][i];
      for (int k = i; k < NA; ++k) fro2 += G[i][k] * G[i][k];
    }
    double lb = det;
    const double f = (double)(NA - 1) / fro2;
    for (int i = 0; i < NA - 1; ++i) lb *= sqrt(f);
    if (lb >= rcond * sqrt(fro2)) { W.rank = NA; return NA; }
  }
  if (nr > 0) {
    double xmin[NA], xmax[NA], col[NA];
    xmin[0] = 1.0; xmax[0] = 1.0;
    double smax = fabs(G[0][0]), smin = smax;
    rank = 1;
#pragma unroll 1
    while (rank < nr) {
      const int i = rank;
      for (int k = 0; k < i; ++k) col[k] = G[k][i];
      double sminpr, s1, c1, smaxpr, s2, c2;
      laic1(false, i, xmin, smin, col, G[i][i], sminpr, s1, c1);
      laic1(true, i, xmax, smax, col, G[i][i], smaxpr, s2, c2);
      if (!(smaxpr * rcond <= sminpr)) break;
      for (int k = 0; k < i; ++k) { xmin[k] *= s1; xmax[k] *= s2; }
      xmin[i] = c1; xmax[i] = c2;
      smin = sminpr; smax = smaxpr;
      ++rank;
    }
  }
  W.rank = rank;
  if (rank > 0 && rank < NA) {                     // M = chol(R1 R1^T), R1 = R[:rank, :]
    for (int i = 0; i < rank; ++i)
      for (int k = 0; k <= i; ++k) {
        double v = 0.0;
        for (int q = i; q < NA; ++q) v += G[i][q] * G[k][q];     // rows i >= k: columns >= i
        W.M[i][k] = v;
      }
    for (int j = 0; j < rank; ++j) {
      double d = W.M[j][j];
      for (int k = 0; k < j; ++k) d -= W.M[j][k] * W.M[j][k];
      const double ljj = sqrt(d);
      W.M[j][j] = ljj;
      for (int i = j + 1; i < rank; ++i) {
        double v = W.M[i][j];
        for (int k = 0; k < j; ++k) v -= W.M[i][k] * W.M[j][k];
        W.M[i][j] = v / ljj;
      }
    }
  }
  return rank;
}

// w := (R1 R1^T)^-1 w  (in place, length rank)
__host__ __device__ __forceinline__ void gelsy_msolve(const GelsyWork& W, double* w) {
  const int r = W.rank;
  for (int i = 0; i < r; ++i) { double v = w[i]; for (int k = 0; k < i; ++k) v -= W.M[i][k] * w[k]; w[i] = v / W.M[i][i]; }
  for (int i = r - 1; i >= 0; --i) { double v = w[i]; for (int k = i + 1; k < r; ++k) v -= W.M[k][i] * w[k]; w[i] = v / W.M[i][i]; }
}

// x := xGELSY solution for the right-hand side v = (X^T y): x = P R1^T (R1 R1^T)^-1 R11^-T (P^T v)[:rank].
// `v` / `x` are strided columns (element a at [a * stride]); x may alias v.
template <int NA>
__host__ __device__ void gelsy_solve(const double (*R)[NA], const GelsyWork& W, const double* v, double* x, int stride) {
  const int r = W.rank;
  double c[NA], y[NA];
  for (int i = 0; i < NA; ++i) c[i] = v[W.perm[i] * stride];
  for (int i = 0; i < r; ++i) { double t = c[i]; for (int k = 0; k < i; ++k) t -= R[k][i] * c[k]; c[i] = t / R[i][i]; }
  if (r == NA) {
    for (int i = NA - 1; i >= 0; --i) { double t = c[i]; for (int k = i + 1; k < NA; ++k) t -= R[i][k] * y[k]; y[i] = t / R[i][i]; }
  } else {
    gelsy_msolve(W, c);
    for (int k = 0; k < NA; ++k) { double t = 0.0; for (int i = 0; i < r && i <= k; ++i) t += R[i][k] * c[i]; y[k] = t; }
  }
  for (int k = 0; k < NA; ++k) x[W.perm[k] * stride] = y[k];
}

// x := G~^+ v with G~ = P R1^T R1 P^T, the Gram matrix of the rank-truncated problem (= G^-1 v at full rank); pass
// `null_out` to also get N v = v - G~^+ G~ v, the component in the discarded directions.  Strided like gelsy_solve.
template <int NA>
__host__ __device__ void gelsy_pinv(const double (*R)[NA], const GelsyWork& W, const double* v, double* x, double* null_out, int stride) {
  const int r = W.rank;
  if (r == NA) {
    gelsy_solve<NA>(R, W, v, x, stride);
    if (null_out) for (int k = 0; k < NA; ++k) null_out[k * stride] = 0.0;
    return;
  }
  double u[NA], t[NA], w[NA], y[NA];
  for (int i = 0; i < NA; ++i) u[i] = v[W.perm[i] * stride];
  for (int i = 0; i < r; ++i) { double a = 0.0; for (int k = i; k < NA; ++k) a += R[i][k] * u[k]; t[i] = a; }    // R1 P^T v
  for (int i = 0; i < r; ++i) w[i] = t[i];
  gelsy_msolve(W, w);                                                                                      // (R1 R1^T)^-1 R1 u
  if (null_out) {
    for (int k = 0; k < NA; ++k) { double a = 0.0; for (int i = 0; i < r && i <= k; ++i) a += R[i][k] * w[i]; y[k] = a; }
    for (int k = 0; k < NA; ++k) null_out[W.perm[k] * stride] = u[k] - y[k];
  }
  gelsy_msolve(W, w);
  for (int k = 0; k < NA; ++k) { double a = 0.0; for (int i = 0; i < r && i <= k; ++i) a += R[i][k] * w[i]; y[k] = a; }
  for (int k = 0; k < NA; ++k) x[W.perm[k] * stride] = y[k];
}

// rcond of the reference call: torch.linalg.lstsq default = eps(fp32) * max(m, n), m = T rows, n = N + 1 columns
__host__ __device__ __forceinline__ double gelsy_rcond(int T, int NA) { return 1.1920928955078125e-07 * (double)max(T, NA); }

// derivative of component i at time row t of one series; `ld(row)` returns Z[row][i].
// Accumulation in ascending column order, fp32 FMA (the reference's sparse.mm is an fp32 row sum).
// HW = compile-time interior half-width (1..4): the interior path is fully unrolled with its
// coefficients read once from the constant bank; only the 2*depth boundary rows take the generic path.
template <int HW, typename LoadF>
__device__ __forceinline__ float sbp_row(const SbpF& sc, int t, int T, LoadF ld) {
  const int depth = sc.depth;
  if (t >= depth && t < T - depth) {
    float acc = 0.f;
#pragma unroll
    for (int k = HW; k >= 1; --k) acc = fmaf(-sc.interior[k - 1], ld(t - k), acc);
#pragma unroll
    for (int k = 1; k <= HW; ++k) acc = fmaf(sc.interior[k - 1], ld(t + k), acc);
    return acc;
  }
  float acc = 0.f;
  if (t < depth) {
    const int w = sc.widths[t];
    for (int c = 0; c < w; ++c) acc = fmaf(sc.rows[t][c], ld(c), acc);
  } else {
    const int r = T - 1 - t;
    const int w = sc.widths[r];
    for (int c = w - 1; c >= 0; --c) acc = fmaf(-sc.rows[r][c], ld(T - 1 - c), acc);
  }
  return acc;
}

// dZ = (1/dt) D Z.  Flat index f = t*N + i inside a series: every warp load/store is one contiguous
// 128-byte line and the stencil neighbours sit +-k*N floats away (L1 hits).  Each thread produces
// kFdPerThread outputs strided by the block size (independent loads in flight); interior rows need no
// index arithmetic at all (row test on f), only the 2*depth boundary rows recover (t, i).
constexpr int kFdThreads = 256;
constexpr int kFdPerThread = 4;
template <int HW, int N>
__global__ void __launch_bounds__(kFdThreads) fd_dzdt_kernel(const float* __restrict__ Z, int T, int n_rt, int fd_id,
                                                             float inv_dt, float* __restrict__ dZ) {
  const SbpF& sc = c_sbp[fd_id];
  const int n = N > 0 ? N : n_rt;
  const size_t base = (size_t)blockIdx.y * T * n;
  const float* __restrict__ z = Z + base;
  float* __restrict__ out = dZ + base;
  const int total = T * n;
  const int lo = sc.depth * n, hi = (T - sc.depth) * n;      // interior rows: lo <= f < hi
  float cf[HW];
#pragma unroll
  for (int k = 0; k < HW; ++k) cf[k] = sc.interior[k];
  const int f0 = blockIdx.x * (kFdThreads * kFdPerThread) + threadIdx.x;
  float v[kFdPerThread];
#pragma unroll
  for (int j = 0; j < kFdPerThread; ++j) {
    const int f = f0 + j * kFdThreads;
    v[j] = 0.f;
    if (f >= lo && f < hi) {
      float acc = 0.f;
#pragma unroll
      for (int k = HW; k >= 1; --k) acc = fmaf(-cf[k - 1], __ldg(z + f - k * n), acc);
#pragma unroll
      for (int k = 1; k <= HW; ++k) acc = fmaf(cf[k - 1], __ldg(z + f + k * n), acc);
      v[j] = acc;
    } else if (f < total) {
      const int t = f / n, i = f - t * n;
      v[j] = sbp_row<HW>(sc, t, T, [&](int row) { return __ldg(z + row * n + i); });
    }
  }
#pragma unroll
  for (int j = 0; j < kFdPerThread; ++j) {
    const int f = f0 + j * kFdThreads;
    if (f < total) out[f] = inv_dt * v[j];
  }
}

// Vectorised variant for a compile-time latent dimension: the batch is one flat array of B*T*N floats, a thread owns
// one aligned float4 of outputs and reads the ALIGNED float4 vectors that cover its stencil reach (HW*N floats to
// either side): every load and store is a full 16-byte access (four times the bytes in flight of the scalar kernel,
// which ncu showed latency-bound: 35 long-scoreboard stalls per issue at 3.2 TB/s), the neighbours are picked out of
// registers at compile-time offsets.  Same FMA order per element as the scalar kernel.  Quads that touch a boundary
// closure, straddle two series or the end of the array take the scalar path (~2*depth*N + 8*Q of T*N elements).
template <int HW, int N, int Q>
__global__ void __launch_bounds__(kFdThreads) fd_dzdt_vec_kernel(const float* __restrict__ Z, long long total, int T,
                                                                 int fd_id, float inv_dt, float* __restrict__ dZ) {
  const SbpF& sc = c_sbp[fd_id];
  constexpr int R = HW * N;                    // stencil reach in floats
  constexpr int VL = (R + 3) / 4;              // aligned vectors to the left of the thread's Q quads
  constexpr int VR = (R + 3) / 4;              // ... to the right (covers the last output + R)
  const long long v = ((long long)blockIdx.x * kFdThreads + threadIdx.x) * Q;     // first of Q consecutive quads
  const long long f0 = v * 4;
  if (f0 >= total) return;
  const int L = T * N;
  const long long b = f0 / L;
  const int loc = (int)(f0 - b * L);
  const int lo = sc.depth * N, hi = (T - sc.depth) * N;
  float cf[HW];
#pragma unroll
  for (int k = 0; k < HW; ++k) cf[k] = sc.interior[k];
  if (loc >= lo && loc + 4 * Q - 1 < hi && (v + Q - 1 + VR) * 4 + 3 < total) {
    float w[(VL + Q + VR) * 4];
    const float4* __restrict__ zv = reinterpret_cast<const float4*>(Z) + (v - VL);
#pragma unroll
    for (int u = 0; u < VL + Q + VR; ++u) {
      const float4 q = __ldg(zv + u);
      w[4 * u + 0] = q.x; w[4 * u + 1] = q.y; w[4 * u + 2] = q.z; w[4 * u + 3] = q.w;
    }
#pragma unroll
    for (int qd = 0; qd < Q; ++qd) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = HW; k >= 1; --k) acc = fmaf(-cf[k - 1], w[(VL + qd) * 4 + j - k * N], acc);
#pragma unroll
        for (int k = 1; k <= HW; ++k) acc = fmaf(cf[k - 1], w[(VL + qd) * 4 + j + k * N], acc);
        o[j] = inv_dt * acc;
      }
      reinterpret_cast<float4*>(dZ)[v + qd] = make_float4(o[0], o[1], o[2], o[3]);
    }
    return;
  }
#pragma unroll 1
  for (int j = 0; j < 4 * Q; ++j) {
    const long long f = f0 + j;
    if (f >= total) break;
    const long long bj = f / L;
    const int lj = (int)(f - bj * L);
    const float* __restrict__ z = Z + bj * L;
    float acc;
    if (lj >= lo && lj < hi) {
      acc = 0.f;
#pragma unroll
      for (int k = HW; k >= 1; --k) acc = fmaf(-cf[k - 1], __ldg(z + lj - k * N), acc);
#pragma unroll
      for (int k = 1; k <= HW; ++k) acc = fmaf(cf[k - 1], __ldg(z + lj + k * N), acc);
    } else {
      const int t = lj / N, i = lj - t * N;
      acc = sbp_row<HW>(sc, t, T, [&](int row) { return __ldg(z + row * N + i); });
    }
    dZ[f] = inv_dt * acc;
  }
}

constexpr int kFitThreads = 128;             // 4 warps = 4 training series per CTA

// One WARP per training series.  Lane l owns time rows l, l+32, ...: it forms dZ/dt for its rows with the
// stencil straight from global memory (the series is 20 KB: L1/L2 resident, HBM sees it once) and
// accumulates the upper triangle of the (n+1)x(n+1) Gram matrix and the (n+1)xn right-hand side in fp64
// registers.  A shuffle butterfly sums over lanes; the tiny Cholesky factorisation / substitutions run as
// rolled loops on a 528-byte per-warp shared-memory scratch (lane 0 factorises, lanes 0..n-1 solve one
// right-hand side each), which keeps the kernel's code small enough to live in the instruction cache.
template <int N, int HW, bool SMEM>
__global__ void __launch_bounds__(kFitThreads) sindy_fit_kernel(const float* __restrict__ Z, int B, int T, int fd_id,
                                                                float inv_dt, int norm_order,
                                                                float* __restrict__ coefs_out,
                                                                float* __restrict__ loss_sindy_out,
                                                                float* __restrict__ loss_coef_out,
                                                                int32_t* __restrict__ info_out) {
  constexpr int NA = N + 1;                    // library size [1 | z]
  constexpr int NG = NA * (NA + 1) / 2;        // upper triangle of the Gram matrix, row-major
  constexpr int NR = NA * N;                   // right-hand side
  constexpr int NS = NG + NR;
  __shared__ double sG[kFitThreads / 32][NA][NA];
  __shared__ double sR[kFitThreads / 32][NA][N];
  __shared__ int sInfo[kFitThreads / 32];
  __shared__ GelsyWork sW[kFitThreads / 32];
  extern __shared__ float s_series[];          // [warps per CTA][T*N] when the series fit (else global reads)
  const SbpF& sc = c_sbp[fd_id];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int b = blockIdx.x * (kFitThreads / 32) + wib;
  if (b >= B) return;                          // whole warp leaves together (no block barriers below)
  const float* __restrict__ z = Z + (size_t)b * T * N;
  uint32_t zs = 0;                             // SMEM: shared-space address of the warp's copy (explicit ld.shared below)
  if constexpr (SMEM) {
    // one coalesced sweep: T*N/32 independent loads per lane in flight, HBM sees the series exactly once
    // asynchronous 4-byte copies (a series is T N floats at a 4-byte aligned offset: no 16-byte bulk copies): all
    // T N / 32 copies of a lane are in flight at once instead of one load -> store round trip after the other
    float* mine = s_series + (size_t)wib * T * N;
    for (int f = lane; f < T * N; f += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(mine + f)), "l"(z + f) : "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    zs = smem_u32(mine);
  }
  // a generic pointer that may be global or shared compiles to generic loads (15 % long-scoreboard stalls in the first ncu)
  auto ldz = [&](int idx) -> float {
    if constexpr (SMEM) {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(zs + 4u * (uint32_t)idx));
      return v;
    } else {
      return __ldg(z + idx);
    }
  };

  double acc[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) acc[q] = 0.0;
  for (int t = lane; t < T; t += 32) {
    float x[NA], d[N];
    x[0] = 1.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      x[1 + i] = ldz(t * N + i);
      d[i] = inv_dt * sbp_row<HW>(sc, t, T, [&](int row) { return ldz(row * N + i); });
    }
    double xd[NA], dd[N];
#pragma unroll
    for (int a = 0; a < NA; ++a) xd[a] = (double)x[a];
#pragma unroll
    for (int i = 0; i < N; ++i) dd[i] = (double)d[i];
    int q = 0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = a; c < NA; ++c) { acc[q] = fma(xd[a], xd[c], acc[q]); ++q; }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int i = 0; i < N; ++i) { acc[q] = fma(xd[a], dd[i], acc[q]); ++q; }
  }
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[q] = v;
  }
  if (lane == 0) {
    int q = 0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = a; c < NA; ++c) { sG[wib][a][c] = acc[q]; sG[wib][c][a] = acc[q]; ++q; }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int i = 0; i < N; ++i) sR[wib][a][i] = acc[q++];
    sInfo[wib] = NA - gelsy_factor<NA>(sG[wib], sW[wib], gelsy_rcond(T, NA));
  }
  __syncwarp();
  const int info = sInfo[wib];                 // number of truncated directions (0 = full rank); NA = not finite
  if (lane < N && info < NA) gelsy_solve<NA>(sG[wib], sW[wib], &sR[wib][0][lane], &sR[wib][0][lane], N);
  __syncwarp();
  float Rf[NA][N];
  double nrm = 0.0;
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      const float cf = info < NA ? (float)sR[wib][a][c] : nanf("");
      Rf[a][c] = cf;
      nrm += (norm_order == 1) ? fabs((double)cf) : (double)cf * (double)cf;
    }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = 0; c < N; ++c) coefs_out[(size_t)b * NR + a * N + c] = Rf[a][c];   // row-major (n+1, n), sindy.py:80
    if (loss_coef_out) loss_coef_out[b] = (float)((norm_order == 1) ? nrm : sqrt(nrm));
    if (info_out) info_out[b] = info;
  }
  if (!loss_sindy_out) return;
  // second pass: MSE(dZdt, [1|Z] R) with the fp32 coefficients, sindy.py:75
  double se = 0.0;
  for (int t = lane; t < T; t += 32) {
    float zr[N];
#pragma unroll
    for (int j = 0; j < N; ++j) zr[j] = ldz(t * N + j);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float d = inv_dt * sbp_row<HW>(sc, t, T, [&](int row) { return ldz(row * N + i); });
      float pred = Rf[0][i];
#pragma unroll
      for (int j = 0; j < N; ++j) pred = fmaf(zr[j], Rf[1 + j][i], pred);
      const float e = d - pred;
      se += (double)e * (double)e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if (lane == 0) loss_sindy_out[b] = (float)(se / ((double)T * N));
}

// ---- backward of the fit -------------------------------------------------------------------------------------
// D[r][c] of the SBP operator (without 1/dt): interior rows are antisymmetric bands, the first/last `depth` rows
// dense closures (right boundary = negated mirror of the left, reference src/lasdi/fd.py:39).
template <int HW>
__device__ __forceinline__ float sbp_entry(const SbpF& sc, int r, int c, int T) {
  const int depth = sc.depth;
  if (r < depth) return c < sc.widths[r] ? sc.rows[r][c] : 0.f;
  if (r >= T - depth) {
    const int rr = T - 1 - r, cc = T - 1 - c;
    return cc < sc.widths[rr] ? -sc.rows[rr][cc] : 0.f;
  }
  const int k = c - r;
  if (k == 0 || k > HW || k < -HW) return 0.f;
  return k > 0 ? sc.interior[k - 1] : -sc.interior[-k - 1];
}

constexpr int kBwdThreads = 128;
constexpr int kFitCtaMaxBatch = 296;         // up to two series per SM go to the CTA-cooperative fit kernel

// Gradient of  L = sum_b gs[b] * loss_sindy[b] + gc[b] * loss_coef[b]  with respect to Z, where (reference
// sindy.py:66-77, differentiated by torch autograd through sparse.mm / lstsq / MSELoss / norm in gplasdi.py:186-197)
//   Y = (1/dt) D Z,  X = [1 | Z],  R = argmin |Y - X R|,  E = Y - X R,
//   loss_sindy = |E|^2 / (T n),  loss_coef = |R|_1 or |R|_F.
// With M = (X^T X)^-1:  dR = M (dX^T E - X^T dX R + X^T dY), so for Rbar = dL/dR and Bm = M Rbar
//   Ybar = X Bm + c E,   Xbar = E Bm^T - Ybar R^T,   c = 2 gs / (T n),
//   Rbar = gc * d|R|/dR - c X^T E      (X^T E = 0 for the exact minimiser; kept because R is rounded to fp32),
//   Zbar = Xbar[:, 1:] + (1/dt) D^T Ybar.
// One CTA per series: the series and Ybar live in shared memory, X^T X / X^T E are accumulated and solved in fp64,
// D^T is applied as a gather in ascending row order (deterministic, no atomics).
template <int N, int HW>
__global__ void __launch_bounds__(kBwdThreads) sindy_fit_bwd_kernel(const float* __restrict__ Z,
                                                                    const float* __restrict__ coefs,
                                                                    const float* __restrict__ g_sindy,
                                                                    const float* __restrict__ g_coef, int T, int fd_id,
                                                                    float inv_dt, int norm_order,
                                                                    float* __restrict__ Zbar) {
  constexpr int NA = N + 1;
  constexpr int NG = NA * (NA + 1) / 2;
  constexpr int NR = NA * N;
  constexpr int NS = NG + NR;
  constexpr int NW = kBwdThreads / 32;
  __shared__ double sRed[NW][NS];
  __shared__ double sG[NA][NA];
  __shared__ double sB[NA][N];
  __shared__ float sBm[NA][N];
  __shared__ float sRf[NA][N];
  __shared__ int sInfo;
  __shared__ GelsyWork sW;
  __shared__ double sNull[NA][N], sPR[NA][N];
  __shared__ float sW3[NA][NA];
  extern __shared__ float s_bwd[];
  float* sZ = s_bwd;                    // [T][N] series, later Xbar[:, 1:]
  float* sY = s_bwd + (size_t)T * N;    // [T][N] residual E, later Ybar
  const SbpF& sc = c_sbp[fd_id];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* __restrict__ z = Z + (size_t)b * T * N;
  for (int f = tid; f < T * N; f += kBwdThreads)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sZ + f)), "l"(z + f) : "memory");
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (tid < NR) sRf[tid / N][tid % N] = __ldg(coefs + (size_t)b * NR + tid);
  __syncthreads();
  const float gs = g_sindy ? __ldg(g_sindy + b) : 0.f;
  const float gc = g_coef ? __ldg(g_coef + b) : 0.f;
  const float c = 2.f * gs / ((float)T * (float)N);

  float Rf[NA][N];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int i = 0; i < N; ++i) Rf[a][i] = sRf[a][i];

  // pass A: residual rows, X^T X (upper triangle) and X^T E in fp64
  double acc[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) acc[q] = 0.0;
  for (int t = tid; t < T; t += kBwdThreads) {
    float x[NA], e[N];
    x[0] = 1.f;
#pragma unroll
    for (int j = 0; j < N; ++j) x[1 + j] = sZ[t * N + j];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float d = inv_dt * sbp_row<HW>(sc, t, T, [&](int row) { return sZ[row * N + i]; });
      float pred = Rf[0][i];
#pragma unroll
      for (int j = 0; j < N; ++j) pred = fmaf(x[1 + j], Rf[1 + j][i], pred);
      e[i] = d - pred;
      sY[t * N + i] = e[i];
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int cc = a; cc < NA; ++cc) { acc[q] = fma((double)x[a], (double)x[cc], acc[q]); ++q; }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int i = 0; i < N; ++i) { acc[q] = fma((double)x[a], (double)e[i], acc[q]); ++q; }
  }
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sRed[warp][q] = v;
  }
  __syncthreads();
  if (tid < NS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += sRed[w][tid];
    sRed[0][tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    int q = 0;
    for (int a = 0; a < NA; ++a)
      for (int cc = a; cc < NA; ++cc) { sG[a][cc] = sRed[0][q]; sG[cc][a] = sRed[0][q]; ++q; }
    // Rbar = gc * d|R|_p/dR - c X^T E
    double nrm = 0.0;
    for (int a = 0; a < NA; ++a)
      for (int i = 0; i < N; ++i) nrm += (double)sRf[a][i] * (double)sRf[a][i];
    nrm = sqrt(nrm);
    for (int a = 0; a < NA; ++a)
      for (int i = 0; i < N; ++i) {
        const double r = (double)sRf[a][i];
        double dn;
        if (norm_order == 1) dn = (r > 0.0) - (r < 0.0);
        else dn = nrm > 0.0 ? r / nrm : 0.0;
        sB[a][i] = (double)gc * dn - (double)c * sRed[0][NG + a * N + i];
      }
    sInfo = NA - gelsy_factor<NA>(sG, sW, gelsy_rcond(T, NA));
  }
  __syncthreads();
  const int info = sInfo;                      // truncated directions of the forward's rank decision (same G, same rule)
  if (tid < N && info < NA) {
    // Bm = G~^+ Rbar; rank-truncated series also need N Rbar (discarded directions) and G~^+ R for the third term of the
    // pseudo-inverse derivative  dX^+ = -X^+ dX X^+ + G~^+ dX^T (I - X X^+) + (I - X^+ X) dX^T X^+T X^+
    gelsy_pinv<NA>(sG, sW, &sB[0][tid], &sB[0][tid], info ? &sNull[0][tid] : nullptr, N);
    for (int i = 0; i < NA; ++i) sBm[i][tid] = (float)sB[i][tid];
    if (info) {
      for (int a = 0; a < NA; ++a) sPR[a][tid] = (double)sRf[a][tid];
      gelsy_pinv<NA>(sG, sW, &sPR[0][tid], &sPR[0][tid], nullptr, N);
    }
  }
  __syncthreads();
  if (info && info < NA && tid < NA * NA) {    // W3 = (G~^+ R) (N Rbar)^T : Xbar += X W3
    const int a = tid / NA, bb = tid - a * NA;
    double v = 0.0;
    for (int i = 0; i < N; ++i) v += sPR[a][i] * sNull[bb][i];
    sW3[a][bb] = (float)v;
  }
  __syncthreads();
  float* __restrict__ out = Zbar + (size_t)b * T * N;
  if (info >= NA) {                      // non-finite series: the forward returned NaN coefficients
    for (int f = tid; f < T * N; f += kBwdThreads) out[f] = nanf("");
    return;
  }
  float Bm[NA][N];
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int i = 0; i < N; ++i) Bm[a][i] = sBm[a][i];
  // pass B: Ybar = X Bm + c E (over E), Xbar[:, 1:] = E Bm^T - Ybar R^T (over Z); every thread touches its own rows
  for (int t = tid; t < T; t += kBwdThreads) {
    float zr[N], e[N], yb[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { zr[j] = sZ[t * N + j]; e[j] = sY[t * N + j]; }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float v = Bm[0][i];
#pragma unroll
      for (int j = 0; j < N; ++j) v = fmaf(zr[j], Bm[1 + j][i], v);
      yb[i] = fmaf(c, e[i], v);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < N; ++i) v = fmaf(e[i], Bm[1 + j][i], v);
#pragma unroll
      for (int i = 0; i < N; ++i) v = fmaf(-yb[i], Rf[1 + j][i], v);
      if (info) {                        // + (X W3)[t, 1 + j]
        float w = sW3[0][1 + j];
#pragma unroll
        for (int a = 0; a < N; ++a) w = fmaf(zr[a], sW3[1 + a][1 + j], w);
        v += w;
      }
      sZ[t * N + j] = v;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) sY[t * N + i] = yb[i];
  }
  __syncthreads();
  // Zbar = Xbar[:, 1:] + (1/dt) D^T Ybar: column t of D gathers rows [t-HW, t+HW] of the interior band and, within
  // SBP_MAXW of an end, the dense boundary rows
  const int depth = sc.depth;
  for (int f = tid; f < T * N; f += kBwdThreads) {
    const int t = f / N, i = f - t * N;
    float a = 0.f;
    if (t < SBP_MAXW)
      for (int r = 0; r < depth; ++r) a = fmaf(sbp_entry<HW>(sc, r, t, T), sY[r * N + i], a);
    const int r0 = max(t - HW, depth), r1 = min(t + HW, T - depth - 1);
    for (int r = r0; r <= r1; ++r) a = fmaf(sbp_entry<HW>(sc, r, t, T), sY[r * N + i], a);
    if (T - 1 - t < SBP_MAXW)
      for (int r = T - depth; r < T; ++r) a = fmaf(sbp_entry<HW>(sc, r, t, T), sY[r * N + i], a);
    out[f] = fmaf(inv_dt, a, sZ[f]);
  }
}

template <int N, int HW>
static cudaError_t launch_bwd_nh(const float* Z, const float* coefs, const float* gs, const float* gc, int B, int T,
                                 int fd_id, float inv_dt, int norm_order, float* Zbar, cudaStream_t st) {
  const size_t smem = (size_t)2 * T * N * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sindy_fit_bwd_kernel<N, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  sindy_fit_bwd_kernel<N, HW><<<B, kBwdThreads, smem, st>>>(Z, coefs, gs, gc, T, fd_id, inv_dt, norm_order, Zbar);
  return cudaGetLastError();
}

template <int N>
static cudaError_t launch_bwd_n(const float* Z, const float* coefs, const float* gs, const float* gc, int B, int T,
                                int fd_id, float inv_dt, int norm_order, float* Zbar, cudaStream_t st) {
  switch (kSbpSchemes[fd_id].halfwidth) {
    case 1: return launch_bwd_nh<N, 1>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st);
    case 2: return launch_bwd_nh<N, 2>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st);
    case 3: return launch_bwd_nh<N, 3>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st);
    default: return launch_bwd_nh<N, 4>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st);
  }
}

int launch_sindy_fit_backward(glasdi_handle* h, const float* Z, const float* coefs, const float* gs, const float* gc,
                              int B, int T, int n, int fd_id, double dt, int norm_order, float* Zbar, cudaStream_t st) {
  if (fd_id < 0 || fd_id >= SBP_NSCHEMES) return set_error(h, GLASDI_ERR_INVALID, "unknown fd scheme id");
  if (B <= 0) return set_error(h, GLASDI_ERR_INVALID, "sindy_fit_backward: empty batch");
  if (T < 2 * kSbpSchemes[fd_id].depth)
    return set_error(h, GLASDI_ERR_INVALID, "sindy_fit_backward: series shorter than 2*boundary depth");
  if (norm_order != 1 && norm_order != 2)
    return set_error(h, GLASDI_ERR_INVALID, "sindy_fit_backward: norm_order must be 1 or 2 ('fro')");
  if ((size_t)2 * T * n * sizeof(float) > 200 * 1024)
    return set_error(h, GLASDI_ERR_INVALID, "sindy_fit_backward: series longer than 25600 values (T * n)");
  int rc = upload_sbp(h);
  if (rc) return rc;
  const float inv_dt = (float)(1.0 / dt);
  cudaError_t e;
  switch (n) {
    case 1: e = launch_bwd_n<1>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    case 2: e = launch_bwd_n<2>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    case 3: e = launch_bwd_n<3>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    case 4: e = launch_bwd_n<4>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    case 5: e = launch_bwd_n<5>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    case 6: e = launch_bwd_n<6>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    case 7: e = launch_bwd_n<7>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    case 8: e = launch_bwd_n<8>(Z, coefs, gs, gc, B, T, fd_id, inv_dt, norm_order, Zbar, st); break;
    default: return set_error(h, GLASDI_ERR_INVALID, "sindy_fit_backward: latent dimension must be 1..8");
  }
  h->launches++;
  return check_cuda(h, e, "sindy_fit_backward launch");
}


// ---- CTA-cooperative fit for the training loop's batch sizes --------------------------------------------------------
// The warp-per-series kernel above is a throughput design (8192 series in 0.33 ms); at n_train <= ~100 series a call is
// the serial time of ONE warp walking its series (58 us at T = 1001).  Here 128 threads share a series: 8 rows per
// thread instead of 32, fp64 partial sums reduced by shuffles and a fixed-order sum over the warps, the same Cholesky
// solve, the same second pass.  Same arithmetic per row; the fp64 sums are taken in a different order, so coefficients
// may differ from the warp kernel's in the last fp32 bit (both are within 1e-6 of LAPACK on the fixtures).
template <int N, int HW>
__global__ void __launch_bounds__(kBwdThreads) sindy_fit_cta_kernel(const float* __restrict__ Z, int T, int fd_id,
                                                                    float inv_dt, int norm_order,
                                                                    float* __restrict__ coefs_out,
                                                                    float* __restrict__ loss_sindy_out,
                                                                    float* __restrict__ loss_coef_out,
                                                                    int32_t* __restrict__ info_out) {
  constexpr int NA = N + 1;
  constexpr int NG = NA * (NA + 1) / 2;
  constexpr int NR = NA * N;
  constexpr int NS = NG + NR;
  constexpr int NW = kBwdThreads / 32;
  __shared__ double sRed[NW][NS];
  __shared__ double sG[NA][NA];
  __shared__ double sR[NA][N];
  __shared__ double sSe[NW];
  __shared__ int sInfo;
  __shared__ GelsyWork sW;
  extern __shared__ float s_fit[];                    // [T][N] the series
  const SbpF& sc = c_sbp[fd_id];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const float* __restrict__ zg = Z + (size_t)b * T * N;
  for (int f = tid; f < T * N; f += kBwdThreads)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_fit + f)), "l"(zg + f) : "memory");
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const float* z = s_fit;

  double acc[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) acc[q] = 0.0;
  for (int t = tid; t < T; t += kBwdThreads) {
    double xd[NA], dd[N];
    xd[0] = 1.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      xd[1 + i] = (double)z[t * N + i];
      dd[i] = (double)(inv_dt * sbp_row<HW>(sc, t, T, [&](int row) { return z[row * N + i]; }));
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = a; c < NA; ++c) { acc[q] = fma(xd[a], xd[c], acc[q]); ++q; }
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int i = 0; i < N; ++i) { acc[q] = fma(xd[a], dd[i], acc[q]); ++q; }
  }
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sRed[warp][q] = v;
  }
  __syncthreads();
  if (tid < NS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) v += sRed[w][tid];
    sRed[0][tid] = v;
  }
  __syncthreads();
  if (tid == 0) {
    int q = 0;
    for (int a = 0; a < NA; ++a)
      for (int c = a; c < NA; ++c) { sG[a][c] = sRed[0][q]; sG[c][a] = sRed[0][q]; ++q; }
    for (int a = 0; a < NA; ++a)
      for (int i = 0; i < N; ++i) sR[a][i] = sRed[0][q++];
    sInfo = NA - gelsy_factor<NA>(sG, sW, gelsy_rcond(T, NA));
  }
  __syncthreads();
  const int info = sInfo;                      // number of truncated directions (0 = full rank); NA = not finite
  if (tid < N && info < NA) gelsy_solve<NA>(sG, sW, &sR[0][tid], &sR[0][tid], N);
  __syncthreads();
  float Rf[NA][N];
  double nrm = 0.0;
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      const float cf = info < NA ? (float)sR[a][c] : nanf("");
      Rf[a][c] = cf;
      nrm += (norm_order == 1) ? fabs((double)cf) : (double)cf * (double)cf;
    }
  if (tid == 0) {
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int c = 0; c < N; ++c) coefs_out[(size_t)b * NR + a * N + c] = Rf[a][c];
    if (loss_coef_out) loss_coef_out[b] = (float)((norm_order == 1) ? nrm : sqrt(nrm));
    if (info_out) info_out[b] = info;
  }
  if (!loss_sindy_out) return;
  double se = 0.0;
  for (int t = tid; t < T; t += kBwdThreads) {
    float zr[N];
#pragma unroll
    for (int j = 0; j < N; ++j) zr[j] = z[t * N + j];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float d = inv_dt * sbp_row<HW>(sc, t, T, [&](int row) { return z[row * N + i]; });
      float pred = Rf[0][i];
#pragma unroll
      for (int j = 0; j < N; ++j) pred = fmaf(zr[j], Rf[1 + j][i], pred);
      const float e = d - pred;
      se += (double)e * (double)e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if (lane == 0) sSe[warp] = se;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) tot += sSe[w];
    loss_sindy_out[b] = (float)(tot / ((double)T * N));
  }
}

template <int HW>
static void launch_fd_hw(const float* Z, int B, int T, int n, int fd_id, float inv_dt, float* dZ, cudaStream_t st) {
  const int per_block = kFdThreads * kFdPerThread;
  dim3 grid((T * n + per_block - 1) / per_block, B);
  const bool aligned = ((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(dZ)) & 15) == 0;
  if (n == 5 && aligned) {
    const long long total = (long long)B * T * n;
    // quads per thread (measured: two help the narrow stencil, 79 -> 86 % of the HBM rate; for the wider ones more quads
    // per thread -- fewer halo loads per output -- were SLOWER: sbp24 0.071 -> 0.078 ms at Q = 2, sbp36 / sbp48 0.090 / 0.100
    // -> 0.164 / 0.170 ms at Q = 4 with 56-float windows in registers, round 2)
    constexpr int Q = HW == 1 ? 2 : 1;
    const long long per_block = (long long)kFdThreads * Q * 4;
    const unsigned int blocks = (unsigned int)((total + per_block - 1) / per_block);
    fd_dzdt_vec_kernel<HW, 5, Q><<<blocks, kFdThreads, 0, st>>>(Z, total, T, fd_id, inv_dt, dZ);
  } else if (n == 5) {
    fd_dzdt_kernel<HW, 5><<<grid, kFdThreads, 0, st>>>(Z, T, n, fd_id, inv_dt, dZ);
  } else {
    fd_dzdt_kernel<HW, 0><<<grid, kFdThreads, 0, st>>>(Z, T, n, fd_id, inv_dt, dZ);
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
  const float inv_dt = (float)(1.0 / dt);
  switch (kSbpSchemes[fd_id].halfwidth) {
    case 1: launch_fd_hw<1>(Z, B, T, n, fd_id, inv_dt, dZ, st); break;
    case 2: launch_fd_hw<2>(Z, B, T, n, fd_id, inv_dt, dZ, st); break;
    case 3: launch_fd_hw<3>(Z, B, T, n, fd_id, inv_dt, dZ, st); break;
    default: launch_fd_hw<4>(Z, B, T, n, fd_id, inv_dt, dZ, st); break;
  }
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "fd_dzdt launch");
}

template <int N, int HW>
static void launch_fit_nh(const float* Z, int B, int T, int fd_id, float inv_dt, int norm_order, float* coefs,
                          float* ls, float* lc, int32_t* info, cudaStream_t st) {
  const size_t smem_cta = (size_t)T * N * sizeof(float);
  if (B <= kFitCtaMaxBatch && smem_cta <= 200 * 1024) {   // training-loop batch sizes: latency, not throughput
    if (smem_cta > 40 * 1024)
      cudaFuncSetAttribute(sindy_fit_cta_kernel<N, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta);
    sindy_fit_cta_kernel<N, HW><<<B, kBwdThreads, smem_cta, st>>>(Z, T, fd_id, inv_dt, norm_order, coefs, ls, lc, info);
    return;
  }
  const int per_cta = kFitThreads / 32;
  size_t smem = (size_t)per_cta * T * N * sizeof(float);
  const int use_smem = smem <= 100 * 1024;            // 2 CTAs per SM keep their series resident (reading through
                                                      // L1/L2 instead, for more resident warps, measured 10 % slower)
  if (!use_smem) smem = 0;
  if (use_smem) {
    if (smem > 40 * 1024)
      cudaFuncSetAttribute(sindy_fit_kernel<N, HW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sindy_fit_kernel<N, HW, true><<<(B + per_cta - 1) / per_cta, kFitThreads, smem, st>>>(Z, B, T, fd_id, inv_dt, norm_order,
                                                                                          coefs, ls, lc, info);
  } else {
    sindy_fit_kernel<N, HW, false><<<(B + per_cta - 1) / per_cta, kFitThreads, 0, st>>>(Z, B, T, fd_id, inv_dt, norm_order,
                                                                                        coefs, ls, lc, info);
  }
}

template <int N>
static int launch_fit_n(glasdi_handle* h, const float* Z, int B, int T, int fd_id, double dt, int norm_order,
                        float* coefs, float* ls, float* lc, int32_t* info, cudaStream_t st) {
  const float inv_dt = (float)(1.0 / dt);
  switch (kSbpSchemes[fd_id].halfwidth) {
    case 1: launch_fit_nh<N, 1>(Z, B, T, fd_id, inv_dt, norm_order, coefs, ls, lc, info, st); break;
    case 2: launch_fit_nh<N, 2>(Z, B, T, fd_id, inv_dt, norm_order, coefs, ls, lc, info, st); break;
    case 3: launch_fit_nh<N, 3>(Z, B, T, fd_id, inv_dt, norm_order, coefs, ls, lc, info, st); break;
    default: launch_fit_nh<N, 4>(Z, B, T, fd_id, inv_dt, norm_order, coefs, ls, lc, info, st); break;
  }
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
