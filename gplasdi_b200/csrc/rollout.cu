// rollout.cu -- latent ODE rollout of the SINDy system: affine  dz/dt = A z + b  (src/lasdi) and the polynomial / sine
// candidate libraries of the legacy drivers (rollout_lib_kernel below).
//
// Replaces the python double loop over SINDy.simulate (reference src/lasdi/gplasdi.py:262-266,
// src/lasdi/latent_dynamics/sindy.py:104-117: scipy odeint/LSODA, fp64, rtol=atol~1.5e-8).
//
// One thread per trajectory (p, s).  The (n+1)*n coefficients never leave registers after the
// first read and all T steps are integrated without touching HBM except to emit the outputs:
//   EXPM : per trajectory, Phi = expm(A dt), g = dt*phi1(A dt) b  (Taylor + scaling/squaring, fp64),
//          then z_{k+1} = Phi z_k + g  -- exact for the affine system for ANY dt (no stability or
//          truncation limit), 30 DFMA per step at n = 5.
//   RK4  : classical RK4 in fp64 with `substeps` sub-steps per output step.
// Outputs: Zs fp32 [P][T][n][Spad] (sample-contiguous => every store/load of the ensemble is a
// coalesced 128 B line; this is what the decoder kernel consumes) and/or Zref fp64 [P,S,T,n]
// (the reference's layout, for the simulate / sample_roms drop-ins).
#include "common.cuh"

#include <algorithm>

namespace glasdi {

template <int N>
struct Affine {
  double A[N][N];
  double b[N];
};

// coefficient vector = row-major flatten of R (n+1, n); b = R[0,:], A[i][j] = R[1+j][i]
template <int N>
__device__ __forceinline__ void load_coefs(const double* __restrict__ c, Affine<N>& m) {
#pragma unroll
  for (int i = 0; i < N; ++i) m.b[i] = c[i];
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) m.A[i][j] = c[(1 + j) * N + i];
}

// Phi = exp(A dt), g = (int_0^dt exp(A s) ds) b
template <int N>
__device__ void affine_propagator(const Affine<N>& m, double dt, double (&Phi)[N][N], double (&g)[N]) {
  // scale: ||A dt||_inf / 2^s <= 0.25
  double nrm = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) r += fabs(m.A[i][j]);
    nrm = fmax(nrm, r);
  }
  nrm *= fabs(dt);
  int s = 0;
  while (nrm > 0.25 && s < 60) { nrm *= 0.5; ++s; }
  const double h = ldexp(dt, -s);

  // X = A h.  Horner for E = sum_{k<=KT} X^k / k!   and  F = sum_{k<=KT} X^k / (k+1)!  (phi1)
  //   E_j = I + X E_{j+1} / j ,  F_j = I/1 ... computed jointly through  F = phi1, E = I + X F.
  constexpr int KT = 14;   // 0.25^15/15! ~ 7e-22
  double F[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) F[i][j] = (i == j) ? 1.0 : 0.0;
  // phi1(X) = I + X/2 (I + X/3 (I + ... X/(KT+1)))
  for (int k = KT + 1; k >= 2; --k) {
    double Tm[N][N];
    const double sc = h / (double)k;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < N; ++l) acc = fma(m.A[i][l], F[l][j], acc);
        Tm[i][j] = acc * sc + ((i == j) ? 1.0 : 0.0);
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) F[i][j] = Tm[i][j];
  }
  // gs = h * F b ; Phi = I + h A F
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) acc = fma(F[i][j], m.b[j], acc);
    g[i] = acc * h;
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < N; ++l) acc = fma(m.A[i][l], F[l][j], acc);
      Phi[i][j] = acc * h + ((i == j) ? 1.0 : 0.0);
    }
  // squaring: Phi_{2h} = Phi_h^2 , g_{2h} = Phi_h g_h + g_h
  for (int q = 0; q < s; ++q) {
    double g2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = g[i];
#pragma unroll
      for (int j = 0; j < N; ++j) acc = fma(Phi[i][j], g[j], acc);
      g2[i] = acc;
    }
    double P2[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < N; ++l) acc = fma(Phi[i][l], Phi[l][j], acc);
        P2[i][j] = acc;
      }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      g[i] = g2[i];
#pragma unroll
      for (int j = 0; j < N; ++j) Phi[i][j] = P2[i][j];
    }
  }
}

template <int N>
__device__ __forceinline__ void rhs(const Affine<N>& m, const double (&z)[N], double (&f)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double acc = m.b[i];
#pragma unroll
    for (int j = 0; j < N; ++j) acc = fma(m.A[i][j], z[j], acc);
    f[i] = acc;
  }
}

template <int N>
__global__ void __launch_bounds__(128)
rollout_kernel(const double* __restrict__ coefs, const float* __restrict__ z0, int z0_per_sample,
               int P, int S, int T, double dt, int integrator, int substeps,
               float* __restrict__ Zs, int Spad, double* __restrict__ Zref) {
  // blockIdx.y = p ; threads cover samples so that stores into Zs[p][t][i][s] coalesce over s
  const int p = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const size_t traj = (size_t)p * S + s;

  Affine<N> m;
  load_coefs<N>(coefs + traj * (size_t)(N * (N + 1)), m);
  double z[N];
  const float* zp = z0_per_sample ? (z0 + traj * N) : (z0 + (size_t)p * N);
#pragma unroll
  for (int i = 0; i < N; ++i) z[i] = (double)zp[i];

  float* zs_base = Zs ? (Zs + ((size_t)p * T * N) * Spad + s) : nullptr;
  double* zr_base = Zref ? (Zref + traj * (size_t)T * N) : nullptr;

  auto emit = [&](int k) {
    if (zs_base) {
#pragma unroll
      for (int i = 0; i < N; ++i) zs_base[((size_t)k * N + i) * Spad] = (float)z[i];
    }
    if (zr_base) {
#pragma unroll
      for (int i = 0; i < N; ++i) zr_base[(size_t)k * N + i] = z[i];
    }
  };

  if (integrator == GLASDI_INT_EXPM) {
    double Phi[N][N], g[N];
    affine_propagator<N>(m, dt, Phi, g);
    emit(0);
    for (int k = 1; k < T; ++k) {
      double zn[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double acc = g[i];
#pragma unroll
        for (int j = 0; j < N; ++j) acc = fma(Phi[i][j], z[j], acc);
        zn[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) z[i] = zn[i];
      emit(k);
    }
  } else {
    const double hstep = dt / (double)substeps;
    emit(0);
    for (int k = 1; k < T; ++k) {
      for (int q = 0; q < substeps; ++q) {
        double k1[N], k2[N], k3[N], k4[N], zt[N];
        rhs<N>(m, z, k1);
#pragma unroll
        for (int i = 0; i < N; ++i) zt[i] = fma(0.5 * hstep, k1[i], z[i]);
        rhs<N>(m, zt, k2);
#pragma unroll
        for (int i = 0; i < N; ++i) zt[i] = fma(0.5 * hstep, k2[i], z[i]);
        rhs<N>(m, zt, k3);
#pragma unroll
        for (int i = 0; i < N; ++i) zt[i] = fma(hstep, k3[i], z[i]);
        rhs<N>(m, zt, k4);
#pragma unroll
        for (int i = 0; i < N; ++i)
          z[i] = fma(hstep / 6.0, (k1[i] + 2.0 * k2[i]) + (2.0 * k3[i] + k4[i]), z[i]);
      }
      emit(k);
    }
  }
}

// ---- non-affine candidate libraries (reference legacy/Vlasov1D1V/utils/utils_sindy.py:91-170) -------------------
// dz/dt = Theta(z) R with Theta = [1, z, z_i z_j (i <= j) if QUAD, sin z if SINE], R = [|Theta|, N] row-major.  One thread
// per trajectory; the |Theta| N fp64 coefficients (up to 424 at N = 8) live in shared memory as [coefficient][thread]
// (conflict-free, read at compile-time offsets), the state and the RK4 stages in registers.  fp64 RK4; the number of
// sub-steps of every output step comes from a row-sum bound of the Jacobian at the current state:
//   L(z) = max_i ( sum_j |A_ij| + sum_j |c_sin,ij| + sum_{j<=k} |q_i,jk| (|z_j| + |z_k|) ),   substeps = ceil(16 dt L)
// so h L <= 1/16: the local truncation error (h L)^5 / 120 per step keeps the trajectory within ~1e-7 L t of LSODA's.
template <int N, bool QUAD, bool SINE>
struct Lib {
  static constexpr int kQ = QUAD ? N * (N + 1) / 2 : 0;
  static constexpr int kTerms = 1 + N + kQ + (SINE ? N : 0);
  static constexpr int kCoefs = kTerms * N;
};

template <int N, bool QUAD, bool SINE>
__device__ __forceinline__ void rhs_lib(const volatile double* c, int bd, const double (&z)[N], double (&f)[N]) {
  using L = Lib<N, QUAD, SINE>;
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = c[i * bd];
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = fma(c[((1 + j) * N + i) * bd], z[j], f[i]);
  if constexpr (QUAD) {
    int q = 0;
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int k = j; k < N; ++k, ++q) {
        const double zz = z[j] * z[k];
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = fma(c[((1 + N + q) * N + i) * bd], zz, f[i]);
      }
  }
  if constexpr (SINE) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double sj = sin(z[j]);
#pragma unroll
      for (int i = 0; i < N; ++i) f[i] = fma(c[((1 + N + L::kQ + j) * N + i) * bd], sj, f[i]);
    }
  }
}

template <int N, bool QUAD, bool SINE>
__global__ void rollout_lib_kernel(const double* __restrict__ coefs, const float* __restrict__ z0, int z0_per_sample,
                                   int P, int S, int T, double dt_uniform, const double* __restrict__ dts, int min_substeps,
                                   float* __restrict__ Zs, int Spad, double* __restrict__ Zref,
                                   unsigned int* __restrict__ flags) {
  using L = Lib<N, QUAD, SINE>;
  extern __shared__ double cs[];                         // [kCoefs][blockDim.x]
  const int bd = blockDim.x;
  const int p = blockIdx.y;
  const int s0 = blockIdx.x * bd;
  const int s = s0 + threadIdx.x;
  // the block's trajectories are contiguous in `coefs`: coalesced read, transposed store
  const int nloc = min(bd, S - s0);
  const double* cblk = coefs + ((size_t)p * S + s0) * L::kCoefs;
  for (int e = threadIdx.x; e < nloc * L::kCoefs; e += bd) cs[(e % L::kCoefs) * bd + e / L::kCoefs] = cblk[e];
  __syncthreads();
  if (s >= S) return;
  const size_t traj = (size_t)p * S + s;
  const volatile double* c = cs + threadIdx.x;   // volatile: re-read at every use instead of being hoisted into (spilled) registers

  double z[N];
  const float* zp = z0_per_sample ? (z0 + traj * N) : (z0 + (size_t)p * N);
#pragma unroll
  for (int i = 0; i < N; ++i) z[i] = (double)zp[i];
  // state-independent part of the Jacobian bound
  double lin[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) a += fabs(c[((1 + j) * N + i) * bd]);
    if constexpr (SINE) {
#pragma unroll
      for (int j = 0; j < N; ++j) a += fabs(c[((1 + N + L::kQ + j) * N + i) * bd]);
    }
    lin[i] = a;
  }

  float* zs_base = Zs ? (Zs + ((size_t)p * T * N) * Spad + s) : nullptr;
  double* zr_base = Zref ? (Zref + traj * (size_t)T * N) : nullptr;
  auto emit = [&](int k) {
    if (zs_base) {
#pragma unroll
      for (int i = 0; i < N; ++i) zs_base[((size_t)k * N + i) * Spad] = (float)z[i];
    }
    if (zr_base) {
#pragma unroll
      for (int i = 0; i < N; ++i) zr_base[(size_t)k * N + i] = z[i];
    }
  };

  bool bad = false;
  emit(0);
  for (int k = 1; k < T; ++k) {
    const double dt = dts ? dts[k - 1] : dt_uniform;           // non-uniform output grids: dts[k-1] = t_k - t_{k-1}
    double lmax = 0.0;
    {
      double li[N];
#pragma unroll
      for (int i = 0; i < N; ++i) li[i] = lin[i];
      if constexpr (QUAD) {
        int q = 0;
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
          for (int kk = j; kk < N; ++kk, ++q) {
            const double w = fabs(z[j]) + fabs(z[kk]);
#pragma unroll
            for (int i = 0; i < N; ++i) li[i] = fma(fabs(c[((1 + N + q) * N + i) * bd]), w, li[i]);
          }
      }
#pragma unroll
      for (int i = 0; i < N; ++i) lmax = fmax(lmax, li[i]);
    }
    double chk = 0.0;                                          // NaN as soon as a component is inf / NaN (fmax drops NaNs)
#pragma unroll
    for (int i = 0; i < N; ++i) chk = fma(z[i], 0.0, chk);
    const double want = ceil(16.0 * fabs(dt) * lmax);
    int sub = min_substeps;
    if (chk != 0.0) { bad = true; }
    else if (!(want <= 16384.0)) { bad = true; sub = 16384; }
    else sub = max(sub, (int)want);
    const double hstep = dt / (double)sub;
    for (int q = 0; q < sub; ++q) {
      double k1[N], k2[N], k3[N], k4[N], zt[N];
      rhs_lib<N, QUAD, SINE>(c, bd, z, k1);
#pragma unroll
      for (int i = 0; i < N; ++i) zt[i] = fma(0.5 * hstep, k1[i], z[i]);
      rhs_lib<N, QUAD, SINE>(c, bd, zt, k2);
#pragma unroll
      for (int i = 0; i < N; ++i) zt[i] = fma(0.5 * hstep, k2[i], z[i]);
      rhs_lib<N, QUAD, SINE>(c, bd, zt, k3);
#pragma unroll
      for (int i = 0; i < N; ++i) zt[i] = fma(hstep, k3[i], z[i]);
      rhs_lib<N, QUAD, SINE>(c, bd, zt, k4);
#pragma unroll
      for (int i = 0; i < N; ++i)
        z[i] = fma(hstep / 6.0, (k1[i] + 2.0 * k2[i]) + (2.0 * k3[i] + k4[i]), z[i]);
    }
    emit(k);
  }
  {
    double chk = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) chk = fma(z[i], 0.0, chk);
    bad |= chk != 0.0;
  }
  if (bad && flags) atomicOr(flags, 1u);
}

template <int N, bool QUAD, bool SINE>
static int launch_rollout_lib_inst(glasdi_handle* h, const double* coefs, const float* z0, int z0ps, int P, int S, int T,
                                   double dt, float* Zs, int Spad, double* Zref, cudaStream_t st, const double* dts = nullptr) {
  using L = Lib<N, QUAD, SINE>;
  int bd = 128;
  while (bd > 32 && (size_t)bd * L::kCoefs * sizeof(double) > 100 * 1024) bd >>= 1;
  const size_t smem = (size_t)bd * L::kCoefs * sizeof(double);
  auto kern = rollout_lib_kernel<N, QUAD, SINE>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(h, e, "rollout_lib_kernel shared memory");
  }
  dim3 block(bd), grid((S + bd - 1) / bd, P);
  kern<<<grid, block, smem, st>>>(coefs, z0, z0ps, P, S, T, dt, dts, h->opt_substeps, Zs, Spad, Zref,
                                         reinterpret_cast<unsigned int*>(h->d_flags + 3));
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "rollout_lib_kernel launch");
}

template <int N>
static int launch_rollout_lib_n(glasdi_handle* h, int lib, const double* coefs, const float* z0, int z0ps, int P, int S,
                                int T, double dt, float* Zs, int Spad, double* Zref, cudaStream_t st, const double* dts = nullptr) {
  switch (lib) {
    case GLASDI_LIB_POLY1: return launch_rollout_lib_inst<N, false, false>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
    case GLASDI_LIB_POLY1_SINE: return launch_rollout_lib_inst<N, false, true>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
    case GLASDI_LIB_POLY2: return launch_rollout_lib_inst<N, true, false>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
    default: return launch_rollout_lib_inst<N, true, true>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
  }
}

int library_terms(int lib, int n) {
  if (lib < GLASDI_LIB_POLY1 || lib > GLASDI_LIB_POLY2_SINE || n < 1) return -1;
  const bool quad = lib == GLASDI_LIB_POLY2 || lib == GLASDI_LIB_POLY2_SINE;
  const bool sine = lib == GLASDI_LIB_POLY1_SINE || lib == GLASDI_LIB_POLY2_SINE;
  return 1 + n + (quad ? n * (n + 1) / 2 : 0) + (sine ? n : 0);
}

// caller-supplied trajectories: Zis fp64 [P,S,T,n] -> Zs fp32 [P][T][n][Spad]
// (the fp64 -> fp32 cast is `torch.Tensor(Zi)` of reference gplasdi.py:65)
__global__ void zis_to_soa_kernel(const double* __restrict__ Zis, int P, int S, int T, int n,
                                  float* __restrict__ Zs, int Spad) {
  // tile transpose through shared memory: reads contiguous in (t,i), writes contiguous in s
  __shared__ float tile[32][33];
  const int p = blockIdx.z;
  const int TN = T * n;
  const int s0 = blockIdx.y * 32, q0 = blockIdx.x * 32;   // q = t*n + i
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int s = s0 + r, q = q0 + threadIdx.x;
    float v = 0.f;
    if (s < S && q < TN) v = (float)Zis[((size_t)p * S + s) * TN + q];
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int q = q0 + r, s = s0 + threadIdx.x;
    if (q < TN && s < S) Zs[((size_t)p * TN + q) * Spad + s] = tile[threadIdx.x][r];
  }
}

template <int N>
static int launch_rollout_n(glasdi_handle* h, const double* coefs, const float* z0, int z0ps, int P, int S, int T,
                            double dt, float* Zs, int Spad, double* Zref, cudaStream_t st) {
  dim3 block(128), grid((S + 127) / 128, P);
  rollout_kernel<N><<<grid, block, 0, st>>>(coefs, z0, z0ps, P, S, T, dt, h->opt_integrator, h->opt_substeps, Zs,
                                            Spad, Zref);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "rollout_kernel launch");
}

int launch_rollout(glasdi_handle* h, const double* coefs, const float* z0, int z0ps, int P, int S, int n, int T,
                   double dt, float* Zs, int Spad, double* Zref, cudaStream_t st, const double* dts) {
  if (P <= 0 || S <= 0 || T <= 0) return set_error(h, GLASDI_ERR_INVALID, "rollout: empty shape");
  if (P > 65535) return set_error(h, GLASDI_ERR_INVALID, "rollout: P > 65535 per call (shard the test grid)");
  // a non-uniform output grid (dts) takes the step-adaptive RK4 kernel for every library, the affine one included: the
  // exact propagator would have to be rebuilt for every distinct step
  if (h->opt_library != GLASDI_LIB_POLY1 || dts != nullptr) {
    const int lib = h->opt_library;
    switch (n) {
      case 1: return launch_rollout_lib_n<1>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      case 2: return launch_rollout_lib_n<2>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      case 3: return launch_rollout_lib_n<3>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      case 4: return launch_rollout_lib_n<4>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      case 5: return launch_rollout_lib_n<5>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      case 6: return launch_rollout_lib_n<6>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      case 7: return launch_rollout_lib_n<7>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      case 8: return launch_rollout_lib_n<8>(h, lib, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st, dts);
      default: return set_error(h, GLASDI_ERR_INVALID, "rollout: latent dimension must be 1..8");
    }
  }
  switch (n) {
    case 1: return launch_rollout_n<1>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    case 2: return launch_rollout_n<2>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    case 3: return launch_rollout_n<3>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    case 4: return launch_rollout_n<4>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    case 5: return launch_rollout_n<5>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    case 6: return launch_rollout_n<6>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    case 7: return launch_rollout_n<7>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    case 8: return launch_rollout_n<8>(h, coefs, z0, z0ps, P, S, T, dt, Zs, Spad, Zref, st);
    default: return set_error(h, GLASDI_ERR_INVALID, "rollout: latent dimension must be 1..8");
  }
}

// columns [S, Spad) of every (p, t, i) row := the row's sample 0 (the pilot of the decode kernels): padding then
// cancels exactly against the pilot and the covariance-form producers need no bounds predicates
__global__ void pad_columns_kernel(float* __restrict__ Zs, long long rows, int S, int Spad) {
  const int npad = Spad - S;
  const long long total = rows * npad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / npad;
    const int c = (int)(e - row * npad);
    Zs[row * Spad + S + c] = Zs[row * Spad];
  }
}

int launch_pad_columns(glasdi_handle* h, float* Zs, long long rows, int S, int Spad, cudaStream_t st) {
  if (Spad <= S || rows <= 0) return GLASDI_OK;
  const long long total = rows * (Spad - S);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 16);
  pad_columns_kernel<<<blocks, 256, 0, st>>>(Zs, rows, S, Spad);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "pad_columns launch");
}

int launch_zis_to_soa(glasdi_handle* h, const double* Zis, int P, int S, int T, int n, float* Zs, int Spad,
                      cudaStream_t st) {
  dim3 block(32, 8), grid((T * n + 31) / 32, (S + 31) / 32, P);
  if (P > 65535 || grid.y > 65535) return set_error(h, GLASDI_ERR_INVALID, "zis_to_soa: grid too large");
  zis_to_soa_kernel<<<grid, block, 0, st>>>(Zis, P, S, T, n, Zs, Spad);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "zis_to_soa launch");
}

}  // namespace glasdi
