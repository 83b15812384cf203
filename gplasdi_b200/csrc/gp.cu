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
#include "mt_jump_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

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
// reading only words older than m0: 623 consecutive words per CTA barrier.  One CTA is bound by instruction issue and
// the barrier round trip (4.5 words / ns), so the STREAM IS CUT INTO SEGMENTS of J = 2^L words that different CTAs
// generate at the same time -- MT19937's jump-ahead:
//   every bit sequence of the generator obeys the recurrence of its characteristic polynomial phi (degree 19937), hence
//   with g_c(t) = t^(c J) mod phi the words c J places ahead are  X[i + c J] = XOR_{k : g_c[k] = 1} X[i + k]  (i >= 1; the
//   low 31 bits of X[0] are not part of the 19937-bit state).
// CTA c > 0 therefore (1) extends the key block to the first 21 183 words of the stream in shared memory (34 steps),
// (2) forms its window X[c J + 1 .. c J + 625) as 624 GF(2) dot products with g_c (thread = word: ~10^4 shared-memory
// loads + XORs each, uniform control flow), (3) streams words from there on with the chained recurrence, a linear
// shared-memory window (all seven loads and three stores at immediate offsets from one pointer) that slides back every
// 12 steps, no bounds tests, global stores straight from the registers.  A segment runs up to 622 words into the next
// one: both CTAs write the same values.  The polynomials g_c come from mt_poly_kernel (once per handle and segment size).
constexpr int kMtBaseSteps = 33;                                   // 624 + 33 * 623 = 21183 >= 19937 + 1 + 624 words
constexpr int kMtBase = kMtN + kMtBaseSteps * kMtStep;
static_assert(kMtBase >= 19937 + 1 + kMtN, "base sequence too short for the jump");
constexpr int kMtSegThreads = 640;                                 // 624 of them own a window word in the jump

__global__ void __launch_bounds__(kMtSegThreads) mt_stream_seg_kernel(const unsigned int* __restrict__ state,
                                                                      const unsigned int* __restrict__ polys,
                                                                      unsigned int* __restrict__ X, int seg_log2,
                                                                      long long n_words) {
  extern __shared__ unsigned int win[];                            // max(kMtBase, kMtWin) words
  const int tid = threadIdx.x;
  const long long c = blockIdx.x;
  const long long J = 1ll << seg_log2;
  for (int i = tid; i < kMtN; i += kMtSegThreads) win[i] = state[i];
  __syncthreads();
  const bool produce = tid < kMtLag, third = tid < kMtStep - 2 * kMtLag;
  long long w0 = 0;                                                // global index of win[0]
  if (c > 0) {
    // (1) the head of the stream
    {
      unsigned int* p = win + kMtN + tid;
      for (int b = 0; b < kMtBaseSteps; ++b) {
        if (produce) {
          const unsigned int v0 = p[-227] ^ mt_twist(p[-624], p[-623]);
          const unsigned int v1 = v0 ^ mt_twist(p[-397], p[-396]);
          p[0] = v0;
          p[227] = v1;
          if (third) p[454] = v1 ^ mt_twist(p[-170], p[-169]);
        }
        p += kMtStep;
        __syncthreads();
      }
    }
    // (2) window word i = X[c J + 1 + i] = XOR over the set bits k of g_c of X[k + 1 + i]
    unsigned int acc = 0;
    if (tid < kMtN) {
      const unsigned int* __restrict__ g = polys + (size_t)c * kMtN;
      const unsigned int* __restrict__ base = win + 1 + tid;
      for (int wq = 0; wq < kMtN; ++wq) {
        unsigned int pw = __ldg(g + wq);                           // uniform over the CTA
        const unsigned int* __restrict__ bq = base + 32 * wq;
        while (pw) {
          const int b = __ffs(pw) - 1;
          pw &= pw - 1;
          acc ^= bq[b];
        }
      }
    }
    __syncthreads();
    if (tid < kMtN) win[tid] = acc;
    __syncthreads();
    w0 = c * J + 1;
  }
  // (3) stream: words [w0 + 624, end) with end = the next segment's first window word (or the end of the buffer)
  const long long end = (c + 1 == gridDim.x) ? n_words : (c + 1) * J + 1;
  for (int i = tid; i < kMtN; i += kMtSegThreads) X[w0 + i] = win[i];
  const long long n_steps = (end - (w0 + kMtN) + kMtStep - 1) / kMtStep;
  unsigned int* xg = X + w0 + kMtN + tid;
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
      const int i = tid;
      const unsigned int keep = i < kMtN ? win[burst * kMtStep + i] : 0u;
      __syncthreads();
      if (i < kMtN) win[i] = keep;
      __syncthreads();
    }
  }
}

// g_c(t) = t^(c 2^L) mod phi for c = c0 + blockIdx.x, as (t^c)^(2^L): L squarings (bit spread) with a Barrett reduction
// each.  phi and mu = floor(t^(2n) / phi) are SPARSE (135 and 161 terms, mt_jump_tables.h), so both products of the
// reduction are a few hundred shifted XORs of a 312-word vector.  Polynomials as 64-bit words, thread = word.
constexpr int kPolyN = 19937;
constexpr int kPolyW = 312;                                        // 64-bit words of a reduced polynomial (19968 bits)
constexpr int kPolyThreads = 640;                                  // >= 624 words of a square

__device__ __forceinline__ unsigned long long spread32(unsigned int x) {
  unsigned long long v = x;
  v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
  v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
  v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
  v = (v | (v << 2)) & 0x3333333333333333ull;
  v = (v | (v << 1)) & 0x5555555555555555ull;
  return v;
}

// out[w] = (v * sum_j t^pos[j])[w] for a 312-word v stored at vpad[kPolyW .. 2 kPolyW) with zeros on both sides
__device__ __forceinline__ unsigned long long sparse_mul_word(const unsigned long long* vpad, const unsigned short* __restrict__ pos,
                                                              int npos, int w) {
  unsigned long long acc = 0;
  for (int j = 0; j < npos; ++j) {
    const int ps = __ldg(pos + j);
    const int idx = kPolyW + w - (ps >> 6), sh = ps & 63;
    if (idx < 1) continue;                                        // entirely below the operand
    const unsigned long long lo = vpad[idx], lo1 = vpad[idx - 1];
    acc ^= sh ? ((lo << sh) | (lo1 >> (64 - sh))) : lo;
  }
  return acc;
}

__global__ void __launch_bounds__(kPolyThreads) mt_poly_kernel(int c0, int seg_log2, const unsigned short* __restrict__ phi_pos,
                                                               int n_phi, const unsigned short* __restrict__ mu_pos, int n_mu,
                                                               unsigned int* __restrict__ polys) {
  __shared__ unsigned long long r[kPolyW];
  __shared__ unsigned long long sq[2 * kPolyW + 2];
  __shared__ unsigned long long m1[2 * kPolyW + 2];
  __shared__ unsigned long long vpad[3 * kPolyW + 2];              // operand of the sparse products, zero padded
  const int w = threadIdx.x;
  const int c = c0 + blockIdx.x;
  if (w < kPolyW) r[w] = (w == (c >> 6)) ? (1ull << (c & 63)) : 0ull;          // t^c, c < 19937
  for (int i = w; i < 3 * kPolyW + 2; i += kPolyThreads) vpad[i] = 0ull;
  if (w < 2) { sq[2 * kPolyW + w] = 0ull; m1[2 * kPolyW + w] = 0ull; }
  __syncthreads();
  constexpr int kWs = kPolyN >> 6, kBs = kPolyN & 63;              // 19937 = 311 * 64 + 33
  for (int it = 0; it < seg_log2; ++it) {
    if (w < 2 * kPolyW) sq[w] = spread32((unsigned int)(r[w >> 1] >> (32 * (w & 1))));
    __syncthreads();
    if (w < kPolyW) vpad[kPolyW + w] = (sq[w + kWs] >> kBs) | (sq[w + kWs + 1] << (64 - kBs));     // hi = sq >> n
    __syncthreads();
    if (w < 2 * kPolyW) m1[w] = sparse_mul_word(vpad, mu_pos, n_mu, w);                            // hi * mu
    __syncthreads();
    if (w < kPolyW) vpad[kPolyW + w] = (m1[w + kWs] >> kBs) | (m1[w + kWs + 1] << (64 - kBs));     // q = (hi mu) >> n
    __syncthreads();
    if (w < kPolyW) {
      unsigned long long v = sq[w] ^ sparse_mul_word(vpad, phi_pos, n_phi, w);                     // sq - q phi, low n bits
      if (w == kWs) v &= (1ull << kBs) - 1ull;
      r[w] = v;
    }
    __syncthreads();
  }
  if (w < kPolyW) {
    polys[(size_t)c * kMtN + 2 * w] = (unsigned int)r[w];
    polys[(size_t)c * kMtN + 2 * w + 1] = (unsigned int)(r[w] >> 32);
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

// g_c = t^(c 2^L) mod phi for c < n_seg, cached on the handle (recomputed when L changes or more segments are needed)
static int ensure_mt_polys(glasdi_handle* h, int seg_log2, int n_seg, cudaStream_t st) {
  if (!h->d_mt_pos) {
    std::vector<unsigned short> pos(gp::kMtPhiPosCount + gp::kMtMuPosCount);
    std::copy(gp::kMtPhiPos, gp::kMtPhiPos + gp::kMtPhiPosCount, pos.begin());
    std::copy(gp::kMtMuPos, gp::kMtMuPos + gp::kMtMuPosCount, pos.begin() + gp::kMtPhiPosCount);
    GL_CUDA(h, cudaMalloc(&h->d_mt_pos, pos.size() * sizeof(unsigned short)));
    GL_CUDA(h, cudaMemcpy(h->d_mt_pos, pos.data(), pos.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
  }
  if (h->mt_poly_log2 == seg_log2 && h->mt_poly_count >= n_seg) return GLASDI_OK;
  const int want = std::max(n_seg, 160);
  if (h->mt_poly_cap < want) {
    GL_CUDA(h, cudaStreamSynchronize(st));
    if (h->d_mt_polys) cudaFree(h->d_mt_polys);
    h->d_mt_polys = nullptr;
    h->mt_poly_cap = 0;
    GL_CUDA(h, cudaMalloc(&h->d_mt_polys, (size_t)want * gp::kMtN * sizeof(unsigned int)));
    h->mt_poly_cap = want;
    h->mt_poly_count = 0;
  }
  if (h->mt_poly_log2 != seg_log2) h->mt_poly_count = 0;
  const int c0 = h->mt_poly_count;
  gp::mt_poly_kernel<<<want - c0, gp::kPolyThreads, 0, st>>>(c0, seg_log2, h->d_mt_pos, gp::kMtPhiPosCount,
                                                            h->d_mt_pos + gp::kMtPhiPosCount, gp::kMtMuPosCount, h->d_mt_polys);
  h->launches++;
  int rc = check_cuda(h, cudaGetLastError(), "mt_poly_kernel launch");
  if (rc) return rc;
  if (seg_log2 == 18 && c0 <= 1) {                      // known answer of the squaring chain (tools/gen_mt_jump.py)
    unsigned int chk[4];
    GL_CUDA(h, cudaMemcpyAsync(chk, h->d_mt_polys + gp::kMtN, sizeof(chk), cudaMemcpyDeviceToHost, st));
    GL_CUDA(h, cudaStreamSynchronize(st));
    if (std::memcmp(chk, gp::kMtG18Check, sizeof(chk)) != 0)
      return set_error(h, GLASDI_ERR_NUMERIC, "MT19937 jump polynomial self-check failed");
  }
  h->mt_poly_log2 = seg_log2;
  h->mt_poly_count = want;
  return GLASDI_OK;
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
  // segments of 2^L words for mt_stream_seg_kernel; the polynomials t^(c 2^L), c < n_seg, need c < 19937: double L
  // until at most 4096 segments are left (L = 0: one CTA generates the whole stream)
  int seg_log2 = h->opt_mt_seg_log2;
  long long n_seg = 1;
  if (seg_log2 > 0) {
    while (((n_total - 1) >> seg_log2) + 1 > 4096) ++seg_log2;
    n_seg = std::max<long long>(1, ((n_total - 1) + (1ll << seg_log2) - 1) >> seg_log2);
  }
  const size_t bytes_x = ((size_t)(n_total + 2048) * sizeof(unsigned int) + 255) / 256 * 256;   // a segment runs < 623 words over
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
    if (n_seg > 1) rc = ensure_mt_polys(h, seg_log2, (int)n_seg, st);
  }
  if (rc == GLASDI_OK) {
    const size_t smem = (size_t)std::max(gp::kMtBase, gp::kMtWin) * sizeof(unsigned int);
    rc = check_cuda(h, cudaFuncSetAttribute(gp::mt_stream_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "mt_stream_seg_kernel shared memory");
    if (rc == GLASDI_OK)
      gp::mt_stream_seg_kernel<<<(unsigned)n_seg, gp::kMtSegThreads, smem, st>>>(dstate, h->d_mt_polys, X, seg_log2 > 0 ? seg_log2 : 62,
                                                                                n_total);
  }
  if (rc == GLASDI_OK) {
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
