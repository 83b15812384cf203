// simt_fallback.cu -- fp32 CUDA-core decode path for decoder shapes the fused tcgen05 kernel does
// not cover yet (last-layer input width > 112, e.g. the Vlasov / rising-bubble decoders with
// K_last = 1000).  Same semantics as decode_var2.cu (reference src/lasdi/gplasdi.py:52-74,
// src/lasdi/latent_space.py:157-164), plain tiled SGEMMs with fused epilogues; it is a GPU path
// (no CPU fallback exists), kept simple and exact in fp32, and is NOT the benchmarked kernel.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace glasdi {
namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;   // 16x16 threads, 4x4 micro tile

// Y[r][o] = act( sum_i X[r][i] W[o][i] + b[o] ),  X [R][in] row-major, W [out][in] row-major
__global__ void __launch_bounds__(256) linear_act_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                         const float* __restrict__ b, float* __restrict__ Y,
                                                         long long R, int in, int out, int act) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long r0 = (long long)blockIdx.x * BM;
  const int o0 = blockIdx.y * BN;
  float acc[TM][TN] = {};
  for (int k0 = 0; k0 < in; k0 += BK) {
    for (int e = threadIdx.x; e < BM * BK; e += 256) {
      const int m = e / BK, k = e % BK;
      const long long r = r0 + m;
      As[k][m] = (r < R && k0 + k < in) ? X[r * in + k0 + k] : 0.f;
    }
    for (int e = threadIdx.x; e < BN * BK; e += 256) {
      const int nn = e / BK, k = e % BK;
      const int o = o0 + nn;
      Bs[k][nn] = (o < out && k0 + k < in) ? W[(size_t)o * in + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long r = r0 + ty * TM + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int o = o0 + tx * TN + j;
      if (o < out) Y[r * out + o] = act_apply(act, acc[i][j] + b[o]);
    }
  }
}

// rows (g, s) of a slab of (p, t) groups: Xrow[(g*S + s)][i] = Zs[p][t][i][s]
__global__ void gather_rows_kernel(const float* __restrict__ Zs, int Spad, int S, int n, long long g0, long long ng,
                                   float* __restrict__ Xrow) {
  const long long total = ng * S;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < total;
       r += (long long)gridDim.x * blockDim.x) {
    const long long g = r / S;
    const int s = (int)(r - g * S);
    const float* src = Zs + ((g0 + g) * n) * Spad + s;
    for (int i = 0; i < n; ++i) Xrow[r * n + i] = src[(size_t)i * Spad];
  }
}

// last layer + variance: block = (DOF tile of 64, group g); loops over sample tiles of 64.
// Per DOF: shifted sums  s = sum(x - c), q = sum (x - c)^2  with c = x of sample 0.
__global__ void __launch_bounds__(256) last_var_kernel(const float* __restrict__ H, const float* __restrict__ W,
                                                       int S, int K, int D, int T, long long g0,
                                                       unsigned int* __restrict__ maxvar_bits) {
  __shared__ float As[BK][BM + 1];   // samples
  __shared__ float Bs[BK][BN + 1];   // dofs
  __shared__ float shift[BN];
  __shared__ float reds[16][BN], redq[16][BN];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long g = blockIdx.x;
  const int d0 = blockIdx.y * BN;
  const float* __restrict__ Hg = H + g * (long long)S * K;
  float ssum[TN] = {}, qsum[TN] = {};
  for (int s0 = 0; s0 < S; s0 += BM) {
    float acc[TM][TN] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
      for (int e = threadIdx.x; e < BM * BK; e += 256) {
        const int m = e / BK, k = e % BK;
        const int s = s0 + m;
        As[k][m] = (s < S && k0 + k < K) ? Hg[(size_t)s * K + k0 + k] : 0.f;
      }
      for (int e = threadIdx.x; e < BN * BK; e += 256) {
        const int nn = e / BK, k = e % BK;
        const int dd = d0 + nn;
        Bs[k][nn] = (dd < D && k0 + k < K) ? W[(size_t)dd * K + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], bb[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx * TN + j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
    if (s0 == 0) {
      if (ty == 0) {
#pragma unroll
        for (int j = 0; j < TN; ++j) shift[tx * TN + j] = acc[0][j];   // sample 0
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      if (s0 + ty * TM + i < S) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const float dlt = acc[i][j] - shift[tx * TN + j];
          ssum[j] += dlt;
          qsum[j] = fmaf(dlt, dlt, qsum[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) { reds[ty][tx * TN + j] = ssum[j]; redq[ty][tx * TN + j] = qsum[j]; }
  __syncthreads();
  float vmax = 0.f;
  if (threadIdx.x < BN) {
    float s = 0.f, q = 0.f;
    for (int r = 0; r < 16; ++r) { s += reds[r][threadIdx.x]; q += redq[r][threadIdx.x]; }
    const float inv = 1.0f / (float)S;
    const float mean = s * inv;
    if (d0 + (int)threadIdx.x < D) vmax = fmaxf(0.f, fmaf(-mean, mean, q * inv));
  }
  if (threadIdx.x < 64) {
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0 && vmax > 0.f) {
      const long long p = (g0 + g) / T;
      atomicMax(maxvar_bits + p, __float_as_uint(vmax));
    }
  }
}

// ---- 128 x 128 x 8 tiles, 8 x 8 outputs per thread, double-buffered shared memory ---------------------------------
// The 64 x 64 kernels above reach ~20 % of the fp32 FMA rate (scalar shared-memory reads, no prefetch); the layers
// that carry the FLOPs of the wide decoders (200 -> 1000, 1000 -> D) go through this main loop instead: float4 global
// loads prefetched one k-step ahead, operands stored k-major so that the inner loop reads two float4 per operand and
// issues 64 FMAs per 4 shared-memory loads.  Needs K % 4 == 0 (16-byte rows); other layers keep the small kernel.
constexpr int GM = 128, GN = 128, GK = 8, GP = 4;          // GP: row padding of the k-major tiles (keeps 16-B alignment)

struct Tiles {
  float A[2][GK][GM + GP];
  float B[2][GK][GN + GP];
};

// acc[i][j] += sum_k Arow[m(i)][k] * Brow[n(j)][k];  m(i) = (i < 4 ? ty*4 + i : 64 + ty*4 + i - 4), n(j) alike with tx
__device__ __forceinline__ void gemm128_mainloop(const float* __restrict__ A, long long m_valid, const float* __restrict__ B,
                                                 int n_valid, int K, Tiles& T, float (&acc)[8][8]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 1, lk = (tid & 1) * 4;             // this thread's row / k-quad of both operand tiles
  const bool a_ok = lr < m_valid, b_ok = lr < n_valid;
  const float* __restrict__ ap = A + (size_t)lr * K + lk;
  const float* __restrict__ bp = B + (size_t)lr * K + lk;
  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
  auto fetch = [&](int k0) {
    ra = (a_ok && k0 + lk < K) ? __ldg(reinterpret_cast<const float4*>(ap + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    rb = (b_ok && k0 + lk < K) ? __ldg(reinterpret_cast<const float4*>(bp + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto stash = [&](int buf) {
    T.A[buf][lk + 0][lr] = ra.x; T.A[buf][lk + 1][lr] = ra.y; T.A[buf][lk + 2][lr] = ra.z; T.A[buf][lk + 3][lr] = ra.w;
    T.B[buf][lk + 0][lr] = rb.x; T.B[buf][lk + 1][lr] = rb.y; T.B[buf][lk + 2][lr] = rb.z; T.B[buf][lk + 3][lr] = rb.w;
  };
  fetch(0);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += GK) {
    const bool more = k0 + GK < K;
    if (more) fetch(k0 + GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&T.A[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&T.A[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&T.B[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&T.B[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
}

__device__ __forceinline__ int sub_index(int t4, int i) { return i < 4 ? t4 * 4 + i : 64 + t4 * 4 + (i - 4); }

__global__ void __launch_bounds__(256, 2) linear_act_kernel128(const float* __restrict__ X, const float* __restrict__ W,
                                                            const float* __restrict__ b, float* __restrict__ Y,
                                                            long long R, int in, int out, int act) {
  __shared__ __align__(16) Tiles T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long r0 = (long long)blockIdx.x * GM;
  const int o0 = blockIdx.y * GN;
  float acc[8][8] = {};
  gemm128_mainloop(X + (size_t)r0 * in, R - r0, W + (size_t)o0 * in, out - o0, in, T, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + sub_index(ty, i);
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = o0 + sub_index(tx, j);
      if (o < out) Y[r * out + o] = act_apply(act, acc[i][j] + b[o]);
    }
  }
}

// last layer + variance, 128-dof tiles: block = (dof tile, group g); loops over sample tiles of 128.
__global__ void __launch_bounds__(256, 2) last_var_kernel128(const float* __restrict__ H, const float* __restrict__ W, int S,
                                                          int K, int D, int T_steps, long long g0,
                                                          unsigned int* __restrict__ maxvar_bits) {
  __shared__ __align__(16) Tiles T;
  __shared__ float shift[GN];
  __shared__ float reds[16][GN], redq[16][GN];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long g = blockIdx.x;
  const int d0 = blockIdx.y * GN;
  const float* __restrict__ Hg = H + g * (long long)S * K;
  float ssum[8] = {}, qsum[8] = {};
  for (int s0 = 0; s0 < S; s0 += GM) {
    float acc[8][8] = {};
    gemm128_mainloop(Hg + (size_t)s0 * K, S - s0, W + (size_t)d0 * K, D - d0, K, T, acc);
    if (s0 == 0) {
      if (ty == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) shift[sub_index(tx, j)] = acc[0][j];   // sample 0 = the pilot of the shifted sums
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (s0 + sub_index(ty, i) < S) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dlt = acc[i][j] - shift[sub_index(tx, j)];
          ssum[j] += dlt;
          qsum[j] = fmaf(dlt, dlt, qsum[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { reds[ty][sub_index(tx, j)] = ssum[j]; redq[ty][sub_index(tx, j)] = qsum[j]; }
  __syncthreads();
  float vmax = 0.f;
  if (threadIdx.x < GN) {
    float s = 0.f, q = 0.f;
    for (int r = 0; r < 16; ++r) { s += reds[r][threadIdx.x]; q += redq[r][threadIdx.x]; }
    const float inv = 1.0f / (float)S;
    const float mean = s * inv;
    if (d0 + (int)threadIdx.x < D) vmax = fmaxf(0.f, fmaf(-mean, mean, q * inv));
  }
  if (threadIdx.x < GN) {
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0 && vmax > 0.f) {
      const long long p = (g0 + g) / T_steps;
      atomicMax(maxvar_bits + p, __float_as_uint(vmax));
    }
  }
}

// layer launcher: the 128-tile kernel where it pays (16-byte rows, enough rows and outputs), else the small one
static inline void launch_linear(const float* X, const float* W, const float* b, float* Y, long long R, int in, int out,
                                 int act, cudaStream_t st) {
  if (in % 4 == 0 && in >= 32 && out >= 64 && R >= 256 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    dim3 grid((unsigned)((R + GM - 1) / GM), (out + GN - 1) / GN);
    linear_act_kernel128<<<grid, 256, 0, st>>>(X, W, b, Y, R, in, out, act);
  } else {
    dim3 grid((unsigned)((R + BM - 1) / BM), (out + BN - 1) / BN);
    linear_act_kernel<<<grid, 256, 0, st>>>(X, W, b, Y, R, in, out, act);
  }
}

// ---- exact re-evaluation of screened candidates, wide decoders (launch_refine_wide) ----------------------------------
struct WRow {
  int p, t;
  unsigned int first, total;        // candidates [first, first + total) of cand[] lie in this (p, t) row of the field
};

// rows (r, s) of a batch of listed (p, t) rows: Xrow[(r*S + s)][i] = Zs[p][t][i][s]
__global__ void gather_listed_rows_kernel(const float* __restrict__ Zs, int Spad, int S, int n, int T,
                                          const WRow* __restrict__ rows, long long nb, float* __restrict__ Xrow) {
  const long long total = nb * S;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < total;
       r += (long long)gridDim.x * blockDim.x) {
    const long long g = r / S;
    const int s = (int)(r - g * S);
    const float* src = Zs + (((long long)rows[g].p * T + rows[g].t) * n) * Spad + s;
    for (int i = 0; i < n; ++i) Xrow[r * n + i] = src[(size_t)i * Spad];
  }
}

constexpr int kWNC = 8;               // candidates per pass over the samples
constexpr int kWThreads = 256;

__device__ __forceinline__ double wr_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// CTA = (row of the batch, unit of <= 8 candidates; blockIdx.y strides the units).  H = exact fp32 activations of the last
// hidden layer of the row's S samples (the CUDA-core chain), thread <-> sample: the decoded value of a candidate dof is the
// fp32 FMA chain over k = 0 .. K-1 on the pilot-centred activations (sample 0 = the pilot) that refine_kernel evaluates for
// the narrow decoders, the sample sums are fp64 (warp shuffle, then a fixed-order sum over the warps).
__global__ void __launch_bounds__(kWThreads) wide_cand_kernel(const float* __restrict__ H, const float* __restrict__ Wlast,
                                                              const WRow* __restrict__ rows, const rf::Cand* __restrict__ cand,
                                                              int S, int K, unsigned int* __restrict__ exact_bits,
                                                              unsigned int* __restrict__ dev_bits, int* __restrict__ flags,
                                                              float dev_tol) {
  extern __shared__ __align__(16) float wsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = (K + 3) / 4 * 4;
  float* sW = wsm;                                   // [kWNC][Kp]
  float* sPil = sW + kWNC * Kp;                      // [Kp]
  float* sXabs = sPil + Kp;                          // [8]
  double* sRed = reinterpret_cast<double*>(sXabs + 8);   // [8 warps][kWNC][2]
  const WRow row = rows[blockIdx.x];
  const float* __restrict__ Hr = H + (size_t)blockIdx.x * S * K;
  const double inv_S = 1.0 / (double)S;
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(Hr) & 15) == 0);
  for (int k = tid; k < Kp; k += kWThreads) sPil[k] = k < K ? Hr[k] : 0.f;
  for (unsigned int unit = blockIdx.y; unit * kWNC < row.total; unit += gridDim.y) {
    const int nc = (int)min((unsigned int)kWNC, row.total - unit * kWNC);
    const rf::Cand* __restrict__ cd = cand + row.first + unit * kWNC;
    __syncthreads();                                 // the previous unit's rows are free (and sPil is there)
    for (int idx = tid; idx < kWNC * Kp; idx += kWThreads) {
      const int c = idx / Kp, k = idx - c * Kp;
      sW[idx] = (c < nc && k < K) ? __ldg(Wlast + (size_t)cd[c].d * K + k) : 0.f;
    }
    __syncthreads();
    if (warp < kWNC) {   // scale of the fp32 rounding noise of a decoded value, one candidate per warp
      float a = 0.f;
      for (int k = lane; k < Kp; k += 32) a = fmaf(fabsf(sW[warp * Kp + k]), fabsf(sPil[k]), a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) sXabs[warp] = a;
    }
    double s1[kWNC], s2[kWNC];
#pragma unroll
    for (int c = 0; c < kWNC; ++c) { s1[c] = 0.0; s2[c] = 0.0; }
    for (int s = tid; s < S; s += kWThreads) {
      const float* __restrict__ hs = Hr + (size_t)s * K;
      float dsum[kWNC];
#pragma unroll
      for (int c = 0; c < kWNC; ++c) dsum[c] = 0.f;
      if (vec) {
        for (int k = 0; k < K; k += 4) {
          const float4 hv = __ldg(reinterpret_cast<const float4*>(hs + k));
          const float4 pv = *reinterpret_cast<const float4*>(sPil + k);
          const float dh0 = hv.x - pv.x, dh1 = hv.y - pv.y, dh2 = hv.z - pv.z, dh3 = hv.w - pv.w;
#pragma unroll
          for (int c = 0; c < kWNC; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(sW + c * Kp + k);
            float d = dsum[c];                        // one chain over k = 0 .. K-1
            d = fmaf(w.x, dh0, d);
            d = fmaf(w.y, dh1, d);
            d = fmaf(w.z, dh2, d);
            d = fmaf(w.w, dh3, d);
            dsum[c] = d;
          }
        }
      } else {
        for (int k = 0; k < K; ++k) {
          const float dh = __ldg(hs + k) - sPil[k];
#pragma unroll
          for (int c = 0; c < kWNC; ++c) dsum[c] = fmaf(sW[c * Kp + k], dh, dsum[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < kWNC; ++c) {
        s1[c] += (double)dsum[c];
        s2[c] = fma((double)dsum[c], (double)dsum[c], s2[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < kWNC; ++c) {
      const double t1 = wr_warp_sum(s1[c]), t2 = wr_warp_sum(s2[c]);
      if (lane == 0) {
        sRed[(warp * kWNC + c) * 2 + 0] = t1;
        sRed[(warp * kWNC + c) * 2 + 1] = t2;
      }
    }
    __syncthreads();
    if (tid < nc) {
      const int c = tid;
      double t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int w = 0; w < kWThreads / 32; ++w) {
        t1 += sRed[(w * kWNC + c) * 2 + 0];
        t2 += sRed[(w * kWNC + c) * 2 + 1];
      }
      const double mean = t1 * inv_S;
      const float var = (float)fmax(t2 * inv_S - mean * mean, 0.0);
      atomicMax(exact_bits + row.p, __float_as_uint(var));
      // screening check, as in refine_kernel: deviations inside the fp32 rounding noise of the decoded value are not counted
      const float se = sqrtf(var), sa = sqrtf(fmaxf(cd[c].var, 0.f));
      const float excess = fabsf(sa - se) - 6.0e-8f * sXabs[c];
      const float dev = (se > 0.f && excess > 0.f) ? excess / se : 0.f;
      atomicMax(dev_bits, __float_as_uint(dev));
      if (dev > dev_tol) atomicOr(flags + 2, 1);
    }
  }
}

}  // namespace simt

static int run_chain(glasdi_handle* h, const float* X0, long long R, int upto_layer, float* bufA, float* bufB,
                     const float** out, cudaStream_t st) {
  // applies layers [0, upto_layer) with activation; returns pointer to the result
  const Decoder& d = h->dec;
  const float* cur = X0;
  float* bufs[2] = {bufA, bufB};
  for (int l = 0; l < upto_layer; ++l) {
    float* dst = bufs[l & 1];
    simt::launch_linear(cur, d.dW[l], d.db[l], dst, R, d.widths[l], d.widths[l + 1], d.act, st);
    h->launches++;
    int rc = check_cuda(h, cudaGetLastError(), "linear_act launch");
    if (rc) return rc;
    cur = dst;
  }
  *out = cur;
  return GLASDI_OK;
}

int launch_decode_var_simt(glasdi_handle* h, const float* Zs, int Spad, int P, int S, int T,
                           unsigned int* maxvar_bits, float* field, int field_ld, cudaStream_t st, int passes,
                           float* mean_field) {
  const Decoder& d = h->dec;
  const int n = d.widths[0];
  int maxw = n;
  for (int l = 1; l < d.n_linear; ++l) maxw = std::max(maxw, d.widths[l]);
  const long long groups = (long long)P * T;
  // slab of (p,t) groups: rows = ng * S, two ping-pong activations + gathered input
  const size_t row_bytes = (size_t)(2 * maxw + n) * sizeof(float);
  const bool wide = h->opt_wide && d.wide_packed != nullptr;        // last layer + variance on tcgen05 (decode_wide.cu)
  const bool wide_front = wide && h->opt_wide >= 2 && d.wide_front != nullptr;   // ... and the last front layer too (xact_kernel)
  if (field && !wide) return set_error(h, GLASDI_ERR_INVALID, "screening pass of a wide decoder needs GLASDI_OPT_WIDE_TENSOR >= 1");
  const int parts = passes > 0 ? passes : (field ? 1 : 3);          // screening: one fp16 pass; direct / field output: fp32-faithful split passes
  if (mean_field && !wide) return set_error(h, GLASDI_ERR_INVALID, "mean field of a wide decoder needs GLASDI_OPT_WIDE_TENSOR >= 1");
  const size_t xt_group = wide_front ? wide_x_bytes(d.wide_k1, S, 1, parts) : 0;
  const size_t bt_group = wide ? wide_b_bytes(d.K, S, 1, parts) + xt_group : 0;   // packed hidden layer (+ packed input) of one (p, t) group
  long long ng = std::max<long long>(
      1, std::min<long long>(groups, ((size_t)1 << 30) / (row_bytes * (size_t)S + bt_group)));
  ng = std::min<long long>(ng, 32768);
  // scratch lives behind the trajectory buffer in the workspace: allocate separately here
  float* scratch = nullptr;
  const size_t chain_bytes = ((size_t)ng * S * row_bytes + 1023) / 1024 * 1024;
  GL_CUDA(h, cudaMallocFromPoolAsync((void**)&scratch, chain_bytes + bt_group * ng, h->pool, st));
  float* Xrow = scratch;
  float* bufA = Xrow + (size_t)ng * S * n;
  float* bufB = bufA + (size_t)ng * S * maxw;
  void* bt_ws = reinterpret_cast<uint8_t*>(scratch) + chain_bytes;
  int rc = GLASDI_OK;
  for (long long g0 = 0; g0 < groups && rc == GLASDI_OK; g0 += ng) {
    const long long gn = std::min(ng, groups - g0);
    const long long R = gn * S;
    const float* Hlast = nullptr;
    if (wide_front && field && parts == 1 && wide_pre_supported(d)) {
      // screening pass: narrow front layers fused (wide_pre_kernel), then both wide layers on tcgen05, one fp16 pass each
      void* xt_ws = reinterpret_cast<uint8_t*>(bt_ws) + wide_b_bytes(d.K, S, ng, parts);
      rc = launch_wide_pre(h, Zs, Spad, g0, gn, S, xt_ws, st);
      if (rc) break;
      rc = launch_wide_front(h, nullptr, gn, S, xt_ws, bt_ws, parts, st);
      if (rc) break;
      rc = launch_wide_last_var(h, nullptr, g0, gn, S, T, bt_ws, maxvar_bits, parts, field, field_ld, st);
      continue;
    }
    simt::gather_rows_kernel<<<(unsigned)std::min<long long>((R + 255) / 256, 65535), 256, 0, st>>>(Zs, Spad, S, n, g0,
                                                                                                     gn, Xrow);
    h->launches++;
    if (wide_front) {
      // layers before the last front layer on the CUDA cores, then both wide layers on tcgen05
      void* xt_ws = reinterpret_cast<uint8_t*>(bt_ws) + wide_b_bytes(d.K, S, ng, parts);
      rc = run_chain(h, Xrow, R, d.n_linear - 2, bufA, bufB, &Hlast, st);
      if (rc) break;
      rc = launch_wide_front(h, Hlast, gn, S, xt_ws, bt_ws, parts, st);
      if (rc) break;
      rc = launch_wide_last_var(h, nullptr, g0, gn, S, T, bt_ws, maxvar_bits, parts, field, field_ld, st, mean_field);
      continue;
    }
    rc = run_chain(h, Xrow, R, d.n_linear - 1, bufA, bufB, &Hlast, st);
    if (rc) break;
    if (wide) {
      rc = launch_wide_last_var(h, Hlast, g0, gn, S, T, bt_ws, maxvar_bits, parts, field, field_ld, st, mean_field);
      continue;
    }
    if (d.K % 4 == 0 && d.K >= 32 && d.D >= 64 && S >= 64 && (reinterpret_cast<uintptr_t>(Hlast) & 15) == 0) {
      dim3 grid((unsigned)gn, (d.D + simt::GN - 1) / simt::GN);
      simt::last_var_kernel128<<<grid, 256, 0, st>>>(Hlast, d.dW[d.n_linear - 1], S, d.K, d.D, T, g0, maxvar_bits);
    } else {
      dim3 grid((unsigned)gn, (d.D + simt::BN - 1) / simt::BN);
      simt::last_var_kernel<<<grid, 256, 0, st>>>(Hlast, d.dW[d.n_linear - 1], S, d.K, d.D, T, g0, maxvar_bits);
    }
    h->launches++;
    rc = check_cuda(h, cudaGetLastError(), "last_var launch");
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

// The candidates of the slab (select_rows_kernel: units of <= 8 candidates, a row's units contiguous, the row total in the
// first unit) are re-evaluated row by row: the exact fp32 chain of ALL front layers for the S samples of each listed (p, t)
// row, then wide_cand_kernel.  The list is read back to size the batches: one stream synchronisation per slab, on a path
// whose slabs take hundreds of milliseconds.
int launch_refine_wide(glasdi_handle* h, const float* Zs, int Spad, int S, int T, unsigned int* exact_bits, float dev_tol,
                       cudaStream_t st) {
  const Decoder& d = h->dec;
  const int n = d.widths[0];
  unsigned int stats[4] = {0, 0, 0, 0};
  GL_CUDA(h, cudaMemcpyAsync(stats, h->d_cand_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
  GL_CUDA(h, cudaStreamSynchronize(st));
  const unsigned int nunits = std::min<unsigned int>(stats[3], (unsigned int)h->group_cap);
  if (nunits == 0 || stats[0] > h->cand_cap) return GLASDI_OK;     // nothing to do / overflow already flagged
  std::vector<rf::Group> groups(nunits);
  GL_CUDA(h, cudaMemcpyAsync(groups.data(), h->d_groups, (size_t)nunits * sizeof(rf::Group), cudaMemcpyDeviceToHost, st));
  GL_CUDA(h, cudaStreamSynchronize(st));
  std::vector<simt::WRow> rows;
  unsigned int max_total = 0;
  for (const rf::Group& g : groups) {
    const unsigned int total = g.count >> 8;
    if (total) {
      rows.push_back(simt::WRow{g.p, g.t, g.first, total});
      max_total = std::max(max_total, total);
    }
  }
  if (rows.empty()) return GLASDI_OK;
  int maxw = n;
  for (int l = 1; l < d.n_linear; ++l) maxw = std::max(maxw, d.widths[l]);
  const size_t row_bytes = (size_t)(2 * maxw + n) * sizeof(float) * (size_t)S;
  const long long nb_max = std::max<long long>(1, std::min<long long>((long long)rows.size(), ((size_t)1 << 30) / row_bytes));
  simt::WRow* d_rows = nullptr;
  float* scratch = nullptr;
  GL_CUDA(h, cudaMallocFromPoolAsync((void**)&d_rows, rows.size() * sizeof(simt::WRow), h->pool, st));
  GL_CUDA(h, cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(simt::WRow), cudaMemcpyHostToDevice, st));
  GL_CUDA(h, cudaMallocFromPoolAsync((void**)&scratch, (size_t)nb_max * row_bytes, h->pool, st));
  float* Xrow = scratch;
  float* bufA = Xrow + (size_t)nb_max * S * n;
  float* bufB = bufA + (size_t)nb_max * S * maxw;
  const int Kp = (d.K + 3) / 4 * 4;
  const size_t smem = ((size_t)(simt::kWNC + 1) * Kp + 8) * sizeof(float) + (size_t)(simt::kWThreads / 32) * simt::kWNC * 2 * sizeof(double);
  GL_CUDA(h, cudaFuncSetAttribute(simt::wide_cand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned int gy = std::min<unsigned int>((max_total + simt::kWNC - 1) / simt::kWNC, 64u);
  int rc = GLASDI_OK;
  for (size_t r0 = 0; r0 < rows.size() && rc == GLASDI_OK; r0 += (size_t)nb_max) {
    const long long nb = std::min<long long>(nb_max, (long long)(rows.size() - r0));
    const long long R = nb * S;
    simt::gather_listed_rows_kernel<<<(unsigned)std::min<long long>((R + 255) / 256, 65535), 256, 0, st>>>(
        Zs, Spad, S, n, T, d_rows + r0, nb, Xrow);
    h->launches++;
    const float* Hlast = nullptr;
    rc = run_chain(h, Xrow, R, d.n_linear - 1, bufA, bufB, &Hlast, st);
    if (rc) break;
    simt::wide_cand_kernel<<<dim3((unsigned)nb, gy), simt::kWThreads, smem, st>>>(
        Hlast, d.dW[d.n_linear - 1], d_rows + r0, reinterpret_cast<const rf::Cand*>(h->d_cand), S, d.K, exact_bits,
        h->d_cand_stats + 2, h->d_flags, dev_tol);
    h->launches++;
    rc = check_cuda(h, cudaGetLastError(), "wide_cand_kernel launch");
  }
  cudaFreeAsync(scratch, st);
  cudaFreeAsync(d_rows, st);
  return rc;
}

int launch_decode_store_simt(glasdi_handle* h, const float* Z, int64_t rows, float* X, cudaStream_t st) {
  const Decoder& d = h->dec;
  int maxw = 1;
  for (int l = 1; l < d.n_linear; ++l) maxw = std::max(maxw, d.widths[l]);
  const long long slab = std::max<long long>(1, std::min<long long>(rows, ((size_t)1 << 29) / ((size_t)maxw * 8)));
  float* scratch = nullptr;
  GL_CUDA(h, cudaMallocFromPoolAsync((void**)&scratch, (size_t)slab * maxw * 2 * sizeof(float), h->pool, st));
  float* bufA = scratch;
  float* bufB = scratch + (size_t)slab * maxw;
  int rc = GLASDI_OK;
  for (long long r0 = 0; r0 < rows && rc == GLASDI_OK; r0 += slab) {
    const long long R = std::min<long long>(slab, rows - r0);
    const float* Hlast = nullptr;
    rc = run_chain(h, Z + (size_t)r0 * d.widths[0], R, d.n_linear - 1, bufA, bufB, &Hlast, st);
    if (rc) break;
    const int l = d.n_linear - 1;
    simt::launch_linear(Hlast, d.dW[l], d.db[l], X + (size_t)r0 * d.D, R, d.K, d.D, GLASDI_ACT_IDENTITY, st);
    h->launches++;
    rc = check_cuda(h, cudaGetLastError(), "linear (last) launch");
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

// Encoder forward (reference latent_space.py:157-164 with reshape_index = 0): act(fc_i(x)) for the hidden layers, a
// linear last layer.  U fp32 [rows, D] -> Z fp32 [rows, n_z].  The first layer reads U once (HBM-bound at D inputs per
// row); the rest is a few KB per row.
int launch_encode_simt(glasdi_handle* h, const float* U, int64_t rows, float* Z, cudaStream_t st) {
  const Encoder& e = h->enc;
  int maxw = 1;
  for (int l = 1; l < e.n_linear; ++l) maxw = std::max(maxw, e.widths[l]);
  float* scratch = nullptr;
  if (e.n_linear > 1) GL_CUDA(h, cudaMallocFromPoolAsync((void**)&scratch, (size_t)rows * maxw * 2 * sizeof(float), h->pool, st));
  float* bufs[2] = {scratch, scratch ? scratch + (size_t)rows * maxw : nullptr};
  const float* cur = U;
  int rc = GLASDI_OK;
  for (int l = 0; l < e.n_linear && rc == GLASDI_OK; ++l) {
    const bool last = l == e.n_linear - 1;
    float* dst = last ? Z : bufs[l & 1];
    simt::launch_linear(cur, e.dW[l], e.db[l], dst, rows, e.widths[l], e.widths[l + 1], last ? GLASDI_ACT_IDENTITY : e.act,
                        st);
    h->launches++;
    rc = check_cuda(h, cudaGetLastError(), "encoder linear_act launch");
    cur = dst;
  }
  if (scratch) cudaFreeAsync(scratch, st);
  return rc;
}

}  // namespace glasdi
