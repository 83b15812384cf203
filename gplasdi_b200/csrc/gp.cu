// gp.cu -- Gaussian-process posterior and coefficient sampling on the device.
//
// Replaces: eval_gp / sample_coefs (reference src/lasdi/gp.py:36-80), i.e. scikit-learn's
// GaussianProcessRegressor.predict(return_std=True) for the kernel the reference fits (ConstantKernel * RBF,
// gp.py:28) evaluated at every test parameter, and the scalar double loop of np.random.normal draws (gp.py:76-78).
//
// * gp_predict_kernel: one thread per (test parameter, coefficient GP), fp64, the same operation sequence as
//   scikit-learn: k* = c exp(-|x/l - x'/l|^2 / 2), mean = k* . alpha, v = L^-1 k* by forward substitution,
//   var = c - |v|^2 clipped at 0.  The training set is tiny (n_train <= a few dozen): the kernel is latency-only.
// * normal draws with numpy's LEGACY generator, bit-compatible with the reference's seeded runs: MT19937 words ->
//   53-bit doubles -> polar Box-Muller with rejection and a cached second value (numpy/random/src/legacy/
//   legacy-distributions.c: legacy_gauss; mt19937.c).  The word stream is inherently sequential (lag-227 linear
//   recurrence): ONE CTA advances it through shared memory, 227 words per barrier, and writes the raw stream to HBM.
//   Everything after it is parallel: every 4-word attempt is tested independently, an exclusive scan over the accept
//   flags gives each accepted pair its position in the output, and the state the host generator must continue from
//   (key block, position, cached gaussian) is reconstructed from the number of words consumed.
//   Values agree with numpy to the last bit except for the rounding of log() (device libm vs glibc: <= 1 ulp).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace glasdi {
namespace gp {

constexpr int kPredThreads = 64;

__global__ void __launch_bounds__(kPredThreads) gp_predict_kernel(
    const double* __restrict__ Xtr, int ntr, int d, const double* __restrict__ alpha, const double* __restrict__ L,
    const double* __restrict__ amp, const double* __restrict__ len, const double* __restrict__ ymean,
    const double* __restrict__ ystd, const double* __restrict__ Xte, int P, int n_gp, double* __restrict__ mean,
    double* __restrict__ stdv) {
  extern __shared__ double sv[];                      // [ntr][kPredThreads]: column = this thread's k* / v
  const int tid = threadIdx.x;
  const int p = blockIdx.x * kPredThreads + tid;
  const int k = blockIdx.y;
  if (p >= P) return;
  const double c = amp[k], l = len[k];
  const double* __restrict__ xt = Xte + (size_t)p * d;
  const double* __restrict__ a = alpha + (size_t)k * ntr;
  const double* __restrict__ Lk = L + (size_t)k * ntr * ntr;
  double m = 0.0;
  for (int j = 0; j < ntr; ++j) {
    double d2 = 0.0;
    for (int i = 0; i < d; ++i) {
      const double diff = xt[i] / l - Xtr[(size_t)j * d + i] / l;       // cdist(X / l, Y / l, 'sqeuclidean')
      d2 = fma(diff, diff, d2);
    }
    const double ks = c * exp(-0.5 * d2);
    sv[j * kPredThreads + tid] = ks;
    m = fma(ks, a[j], m);
  }
  double q = 0.0;
  for (int i = 0; i < ntr; ++i) {                     // forward substitution L v = k*
    double v = sv[i * kPredThreads + tid];
    for (int j = 0; j < i; ++j) v = fma(-Lk[(size_t)i * ntr + j], sv[j * kPredThreads + tid], v);
    v /= Lk[(size_t)i * ntr + i];
    sv[i * kPredThreads + tid] = v;
    q = fma(v, v, q);
  }
  double var = c - q;
  if (var < 0.0) var = 0.0;
  const double ys = ystd ? ystd[k] : 1.0, ym = ymean ? ymean[k] : 0.0;
  mean[(size_t)p * n_gp + k] = ys * m + ym;
  stdv[(size_t)p * n_gp + k] = sqrt(var * ys * ys);
}

// ---- MT19937 -------------------------------------------------------------------------------------------------
constexpr int kMtN = 624, kMtM = 397, kMtLag = kMtN - kMtM;        // lag 227
constexpr int kMtStep = kMtN - 1;                                  // 623 new words per barrier
constexpr int kMtThreads = 256;                                    // 227 of them produce (up to three words each)
constexpr int kMtWinSteps = 12;                                    // steps before the window slides back
constexpr int kMtWin = kMtN + kMtWinSteps * kMtStep;               // shared-memory window, linear addressing

__device__ __forceinline__ unsigned int mt_twist(unsigned int a, unsigned int b) {
  const unsigned int y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ unsigned int mt_temper(unsigned int y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// X[0 .. 624) = the generator's current key block; X[i] = X[i - 227] ^ T(i - 624), T(i) = twist(X[i], X[i + 1])
// (the in-place block regeneration of mt19937_gen written as one linear recurrence over the untempered stream).
// The recurrence is XOR-linear, so thread t of a step that starts at word m0 chains three words off one another,
//   v0 = X[m]       = X[m - 227] ^ T(m - 624)          m = m0 + t, t < 227
//   v1 = X[m + 227] = v0 ^ T(m - 397)
//   v2 = X[m + 454] = v1 ^ T(m - 170)                  t < 169
// reading only words older than m0: 623 consecutive words per CTA barrier.  The kernel is bound by instruction issue
// and the barrier round trip of ONE CTA, so the loop is kept minimal: a linear shared-memory window (all seven loads
// and three stores at immediate offsets from one pointer) that slides back every 12 steps, no bounds tests (the
// stream buffer is padded to whole steps), global stores straight from the registers.
__global__ void __launch_bounds__(kMtThreads) mt_stream_kernel(const unsigned int* __restrict__ state,
                                                               unsigned int* __restrict__ X, long long n_steps) {
  __shared__ unsigned int win[kMtWin];
  const int tid = threadIdx.x;
  for (int i = tid; i < kMtN; i += kMtThreads) {
    const unsigned int v = state[i];
    win[i] = v;
    X[i] = v;
  }
  __syncthreads();
  const bool produce = tid < kMtLag, third = tid < kMtStep - 2 * kMtLag;
  unsigned int* xg = X + kMtN + tid;
  long long step = 0;
  while (step < n_steps) {
    unsigned int* p = win + kMtN + tid;
    const int burst = (int)((n_steps - step < kMtWinSteps) ? n_steps - step : kMtWinSteps);
    for (int b = 0; b < burst; ++b) {
      if (produce) {
        const unsigned int v0 = p[-227] ^ mt_twist(p[-624], p[-623]);
        const unsigned int v1 = v0 ^ mt_twist(p[-397], p[-396]);
        p[0] = v0;
        p[227] = v1;
        xg[0] = v0;
        xg[227] = v1;
        if (third) {
          const unsigned int v2 = v1 ^ mt_twist(p[-170], p[-169]);
          p[454] = v2;
          xg[454] = v2;
        }
      }
      p += kMtStep;
      xg += kMtStep;
      __syncthreads();
    }
    step += burst;
    if (step < n_steps) {                              // slide: the last 624 words become the head of the window
      unsigned int keep[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int i = tid + u * kMtThreads;
        keep[u] = i < kMtN ? win[burst * kMtStep + i] : 0u;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int i = tid + u * kMtThreads;
        if (i < kMtN) win[i] = keep[u];
      }
      __syncthreads();
    }
  }
}

struct Attempt {
  double x1, x2, r2;
  bool ok;
};
// one trial of legacy_gauss: four words -> two doubles in [0, 1) -> a point in the square; separate multiplies and
// adds (no FMA contraction), as the host library computes it
__device__ __forceinline__ Attempt attempt_at(const unsigned int* __restrict__ X, long long q) {
  const unsigned int a0 = mt_temper(X[q]) >> 5, b0 = mt_temper(X[q + 1]) >> 6;
  const unsigned int a1 = mt_temper(X[q + 2]) >> 5, b1 = mt_temper(X[q + 3]) >> 6;
  const double d0 = __dadd_rn(__dmul_rn((double)a0, 67108864.0), (double)b0) / 9007199254740992.0;
  const double d1 = __dadd_rn(__dmul_rn((double)a1, 67108864.0), (double)b1) / 9007199254740992.0;
  Attempt t;
  t.x1 = __dadd_rn(__dmul_rn(2.0, d0), -1.0);
  t.x2 = __dadd_rn(__dmul_rn(2.0, d1), -1.0);
  t.r2 = __dadd_rn(__dmul_rn(t.x1, t.x1), __dmul_rn(t.x2, t.x2));
  t.ok = !(t.r2 >= 1.0 || t.r2 == 0.0);
  return t;
}

constexpr int kGsThreads = 256;
constexpr int kGsPer = 4;                             // attempts per thread
constexpr int kGsBlock = kGsThreads * kGsPer;

__global__ void __launch_bounds__(kGsThreads) gauss_count_kernel(const unsigned int* __restrict__ X, long long q0,
                                                                 long long n_attempts,
                                                                 unsigned int* __restrict__ block_counts) {
  __shared__ unsigned int swarp[kGsThreads / 32];
  const long long a0 = (long long)blockIdx.x * kGsBlock + (long long)threadIdx.x * kGsPer;
  unsigned int cnt = 0;
#pragma unroll
  for (int u = 0; u < kGsPer; ++u)
    if (a0 + u < n_attempts) cnt += attempt_at(X, q0 + 4 * (a0 + u)).ok ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = 0;
    for (int w = 0; w < kGsThreads / 32; ++w) t += swarp[w];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive scan of block_counts (64-bit offsets), one CTA
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const unsigned int* __restrict__ counts, long long n,
                                                           unsigned long long* __restrict__ offsets) {
  __shared__ unsigned long long swarp[32];
  __shared__ unsigned long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + threadIdx.x;
    const unsigned long long v = i < n ? counts[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) swarp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = swarp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      swarp[lane] = w;                                // inclusive over warps
    }
    __syncthreads();
    const unsigned long long carry = carry_s;
    const unsigned long long before = carry + (warp ? swarp[warp - 1] : 0ull) + incl - v;
    if (i < n) offsets[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry_s;        // total number of accepted attempts
}

// info[0] = words consumed (64-bit), info[2] = has_gauss after the call, info[4..5] = cached gaussian bits
__global__ void __launch_bounds__(kGsThreads) gauss_scatter_kernel(
    const unsigned int* __restrict__ X, long long q0, long long n_attempts,
    const unsigned long long* __restrict__ offsets, unsigned int h0, double gauss0,
    const double* __restrict__ mean, const double* __restrict__ stdv, long long N, int S, int n,
    double* __restrict__ out, unsigned long long* __restrict__ info) {
  __shared__ unsigned int swarp[kGsThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long a0 = (long long)blockIdx.x * kGsBlock + (long long)threadIdx.x * kGsPer;
  const long long n_pairs = (N - h0 + 1) / 2;
  Attempt at[kGsPer];
  unsigned int cnt = 0;
#pragma unroll
  for (int u = 0; u < kGsPer; ++u) {
    at[u].ok = false;
    if (a0 + u < n_attempts) at[u] = attempt_at(X, q0 + 4 * (a0 + u));
    cnt += at[u].ok ? 1u : 0u;
  }
  unsigned int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) swarp[warp] = incl;
  __syncthreads();
  unsigned int wbefore = 0;
  for (int w = 0; w < warp; ++w) wbefore += swarp[w];
  long long j = (long long)offsets[blockIdx.x] + wbefore + (incl - cnt);     // pair index of this thread's first accept
  const long long SN = (long long)S * n;
  auto emit = [&](long long idx, double g) {
    const long long p = idx / SN;
    const int k = (int)(idx % n);
    const size_t mi = (size_t)p * n + k;
    out[idx] = __dadd_rn(mean[mi], __dmul_rn(stdv[mi], g));                   // loc + scale * gauss
  };
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (h0 && N > 0) emit(0, gauss0);                                         // the cached value goes first
    if (n_pairs == 0) {                                                       // nothing drawn from the stream
      info[0] = 0ull;
      info[1] = (N == 0 && h0) ? 1ull : 0ull;
      info[2] = (N == 0 && h0) ? (unsigned long long)__double_as_longlong(gauss0) : 0ull;
    }
  }
#pragma unroll
  for (int u = 0; u < kGsPer; ++u) {
    if (!at[u].ok) continue;
    if (j < n_pairs) {
      const double f = sqrt(__dmul_rn(-2.0, log(at[u].r2)) / at[u].r2);
      const double g_first = __dmul_rn(f, at[u].x2), g_second = __dmul_rn(f, at[u].x1);
      const long long idx = (long long)h0 + 2 * j;
      emit(idx, g_first);
      if (idx + 1 < N) emit(idx + 1, g_second);
      if (j == n_pairs - 1) {
        info[0] = (unsigned long long)(4 * (a0 + u + 1));                     // words consumed
        const bool left = idx + 1 >= N;                                       // second value stays cached
        info[1] = left ? 1ull : 0ull;
        info[2] = left ? (unsigned long long)__double_as_longlong(g_second) : 0ull;
      }
    }
    ++j;
  }
}

// new generator state from the number of words consumed: key = the block the read pointer sits in (numpy refills
// lazily: a pointer at the end of a block stays there with pos = 624).  state_out: key[624], pos, has_gauss,
// cached gaussian (2 words), shortfall flag.
__global__ void __launch_bounds__(kMtThreads) mt_finalize_kernel(const unsigned int* __restrict__ X,
                                                                 const unsigned long long* __restrict__ offsets,
                                                                 long long n_blocks, long long n_pairs, long long pos0,
                                                                 const unsigned long long* __restrict__ info,
                                                                 unsigned int* __restrict__ state_out) {
  const unsigned long long accepted = offsets[n_blocks];
  if (accepted < (unsigned long long)n_pairs) {       // not enough accepted attempts in the generated stream
    if (threadIdx.x == 0) state_out[kMtN + 4] = 1u;
    return;
  }
  const long long q = pos0 + (long long)info[0];
  const long long blk = q == 0 ? 0 : (q - 1) / kMtN;
  for (int i = threadIdx.x; i < kMtN; i += kMtThreads) state_out[i] = X[blk * kMtN + i];
  if (threadIdx.x == 0) {
    state_out[kMtN] = (unsigned int)(q - blk * kMtN);
    state_out[kMtN + 1] = (unsigned int)info[1];
    state_out[kMtN + 2] = (unsigned int)(info[2] & 0xffffffffull);
    state_out[kMtN + 3] = (unsigned int)(info[2] >> 32);
    state_out[kMtN + 4] = 0u;
  }
}

}  // namespace gp

int launch_gp_predict(glasdi_handle* h, const double* Xtr, int ntr, int d, int n_gp, const double* alpha,
                      const double* L, const double* amp, const double* len, const double* ymean, const double* ystd,
                      const double* Xte, int P, double* mean, double* stdv, cudaStream_t st) {
  const size_t smem = (size_t)ntr * gp::kPredThreads * sizeof(double);
  if (smem > 200 * 1024) return set_error(h, GLASDI_ERR_INVALID, "gp_predict: more than 400 training points");
  if (n_gp > 65535) return set_error(h, GLASDI_ERR_INVALID, "gp_predict: more than 65535 GPs");
  if (smem > 40 * 1024)
    GL_CUDA(h, cudaFuncSetAttribute(gp::gp_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((P + gp::kPredThreads - 1) / gp::kPredThreads, n_gp);
  gp::gp_predict_kernel<<<grid, gp::kPredThreads, smem, st>>>(Xtr, ntr, d, alpha, L, amp, len, ymean, ystd, Xte, P, n_gp,
                                                             mean, stdv);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "gp_predict launch");
}

int launch_normal_mt19937(glasdi_handle* h, glasdi_mt19937_state* state, const double* mean, const double* stdv, int P,
                          int S, int n, double* out, cudaStream_t st) {
  if (state->pos < 0 || state->pos > gp::kMtN)
    return set_error(h, GLASDI_ERR_INVALID, "normal_mt19937: state position out of range");
  const long long N = (long long)P * S * n;
  const unsigned int h0 = state->has_gauss ? 1u : 0u;
  const long long n_pairs = (N - h0 + 1) / 2;
  // attempts to generate: expected n_pairs * 4/pi, plus 1 % and a constant (the count of accepted attempts among A
  // trials has a standard deviation of 0.41 sqrt(A)); a shortfall is reported (GLASDI_ERR_NUMERIC), never ignored
  const long long n_attempts = (long long)std::ceil((double)n_pairs * 1.2732395447351628 * 1.01) + 4096;
  const long long n_blocks = (n_attempts + gp::kGsBlock - 1) / gp::kGsBlock;
  // words to generate: the key block, pos <= 624 words already consumed, 4 per attempt, one more key block (the
  // final state copies a whole one), in whole steps of 623 words (mt_stream_kernel has no bounds tests)
  const long long n_need = 2 * gp::kMtN + 4 * n_attempts + gp::kMtN;
  const long long n_steps = (n_need - gp::kMtN + gp::kMtStep - 1) / gp::kMtStep;
  const long long n_total = gp::kMtN + n_steps * gp::kMtStep;
  const size_t bytes_x = ((size_t)n_total * sizeof(unsigned int) + 255) / 256 * 256;
  const size_t bytes_cnt = ((size_t)n_blocks * sizeof(unsigned int) + 15) / 16 * 16;
  const size_t bytes_off = (size_t)(n_blocks + 1 + 4) * sizeof(unsigned long long);
  const size_t bytes_state = 640 * sizeof(unsigned int);
  if (bytes_x > ((size_t)16 << 30)) return set_error(h, GLASDI_ERR_OOM, "normal_mt19937: stream buffer above 16 GB");
  // the handle's grow-only workspace (shared with the greedy step; calls on one handle are stream-ordered)
  int rc = ensure_workspace(h, bytes_x + bytes_cnt + bytes_off + bytes_state);
  if (rc) return rc;
  char* scratch = reinterpret_cast<char*>(h->ws);
  unsigned int* X = reinterpret_cast<unsigned int*>(scratch);
  unsigned int* counts = reinterpret_cast<unsigned int*>(scratch + bytes_x);
  unsigned long long* offsets = reinterpret_cast<unsigned long long*>(scratch + bytes_x + bytes_cnt);
  unsigned long long* info = offsets + n_blocks + 1;
  unsigned int* dstate = reinterpret_cast<unsigned int*>(scratch + bytes_x + bytes_cnt + bytes_off);
  unsigned int hstate[632];
  rc = check_cuda(h, cudaMemcpyAsync(dstate, state->key, gp::kMtN * sizeof(unsigned int), cudaMemcpyHostToDevice, st),
                      "normal_mt19937 state upload");
  if (rc == GLASDI_OK) {
    gp::mt_stream_kernel<<<1, gp::kMtThreads, 0, st>>>(dstate, X, n_steps);
    gp::gauss_count_kernel<<<(unsigned)n_blocks, gp::kGsThreads, 0, st>>>(X, (long long)state->pos, n_attempts, counts);
    gp::scan_blocks_kernel<<<1, 1024, 0, st>>>(counts, n_blocks, offsets);
    gp::gauss_scatter_kernel<<<(unsigned)n_blocks, gp::kGsThreads, 0, st>>>(
        X, (long long)state->pos, n_attempts, offsets, h0, state->cached_gaussian, mean, stdv, N, S, n, out, info);
    gp::mt_finalize_kernel<<<1, gp::kMtThreads, 0, st>>>(X, offsets, n_blocks, n_pairs, (long long)state->pos, info,
                                                         dstate);
    h->launches += 5;
    rc = check_cuda(h, cudaGetLastError(), "normal_mt19937 launch");
  }
  if (rc == GLASDI_OK)
    rc = check_cuda(h, cudaMemcpyAsync(hstate, dstate, 632 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st),
                    "normal_mt19937 state download");
  if (rc == GLASDI_OK) rc = check_cuda(h, cudaStreamSynchronize(st), "normal_mt19937 sync");
  if (rc) return rc;
  if (hstate[gp::kMtN + 4]) return set_error(h, GLASDI_ERR_NUMERIC, "normal_mt19937: accepted-attempt shortfall (retry)");
  std::memcpy(state->key, hstate, gp::kMtN * sizeof(unsigned int));
  state->pos = (int)hstate[gp::kMtN];
  state->has_gauss = (int)hstate[gp::kMtN + 1];
  const unsigned long long bits = ((unsigned long long)hstate[gp::kMtN + 3] << 32) | hstate[gp::kMtN + 2];
  std::memcpy(&state->cached_gaussian, &bits, sizeof(double));
  return GLASDI_OK;
}

}  // namespace glasdi
