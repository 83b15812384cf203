// simt_fallback.cu -- fp32 CUDA-core decode path for decoder shapes the fused tcgen05 kernel does
// not cover yet (last-layer input width > 112, e.g. the Vlasov / rising-bubble decoders with
// K_last = 1000).  Same semantics as decode_var.cu (reference src/lasdi/gplasdi.py:52-74,
// src/lasdi/latent_space.py:157-164), plain tiled SGEMMs with fused epilogues; it is a GPU path
// (no CPU fallback exists), kept simple and exact in fp32, and is NOT the benchmarked kernel.
#include "common.cuh"

#include <algorithm>

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

}  // namespace simt

static int run_chain(glasdi_handle* h, const float* X0, long long R, int upto_layer, float* bufA, float* bufB,
                     const float** out, cudaStream_t st) {
  // applies layers [0, upto_layer) with activation; returns pointer to the result
  const Decoder& d = h->dec;
  const float* cur = X0;
  float* bufs[2] = {bufA, bufB};
  for (int l = 0; l < upto_layer; ++l) {
    float* dst = bufs[l & 1];
    dim3 grid((unsigned)((R + simt::BM - 1) / simt::BM), (d.widths[l + 1] + simt::BN - 1) / simt::BN);
    simt::linear_act_kernel<<<grid, 256, 0, st>>>(cur, d.dW[l], d.db[l], dst, R, d.widths[l], d.widths[l + 1], d.act);
    h->launches++;
    int rc = check_cuda(h, cudaGetLastError(), "linear_act launch");
    if (rc) return rc;
    cur = dst;
  }
  *out = cur;
  return GLASDI_OK;
}

int launch_decode_var_simt(glasdi_handle* h, const float* Zs, int Spad, int P, int S, int T,
                           unsigned int* maxvar_bits, cudaStream_t st) {
  const Decoder& d = h->dec;
  const int n = d.widths[0];
  int maxw = n;
  for (int l = 1; l < d.n_linear; ++l) maxw = std::max(maxw, d.widths[l]);
  const long long groups = (long long)P * T;
  // slab of (p,t) groups: rows = ng * S, two ping-pong activations + gathered input
  const size_t row_bytes = (size_t)(2 * maxw + n) * sizeof(float);
  long long ng = std::max<long long>(1, std::min<long long>(groups, ((size_t)1 << 30) / (row_bytes * (size_t)S)));
  ng = std::min<long long>(ng, 32768);
  // scratch lives behind the trajectory buffer in the workspace: allocate separately here
  float* scratch = nullptr;
  GL_CUDA(h, cudaMallocAsync((void**)&scratch, (size_t)ng * S * row_bytes, st));
  float* Xrow = scratch;
  float* bufA = Xrow + (size_t)ng * S * n;
  float* bufB = bufA + (size_t)ng * S * maxw;
  int rc = GLASDI_OK;
  for (long long g0 = 0; g0 < groups && rc == GLASDI_OK; g0 += ng) {
    const long long gn = std::min(ng, groups - g0);
    const long long R = gn * S;
    simt::gather_rows_kernel<<<(unsigned)std::min<long long>((R + 255) / 256, 65535), 256, 0, st>>>(Zs, Spad, S, n, g0,
                                                                                                     gn, Xrow);
    h->launches++;
    const float* Hlast = nullptr;
    rc = run_chain(h, Xrow, R, d.n_linear - 1, bufA, bufB, &Hlast, st);
    if (rc) break;
    dim3 grid((unsigned)gn, (d.D + simt::BN - 1) / simt::BN);
    simt::last_var_kernel<<<grid, 256, 0, st>>>(Hlast, d.dW[d.n_linear - 1], S, d.K, d.D, T, g0, maxvar_bits);
    h->launches++;
    rc = check_cuda(h, cudaGetLastError(), "last_var launch");
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

int launch_decode_store_simt(glasdi_handle* h, const float* Z, int64_t rows, float* X, cudaStream_t st) {
  const Decoder& d = h->dec;
  int maxw = 1;
  for (int l = 1; l < d.n_linear; ++l) maxw = std::max(maxw, d.widths[l]);
  const long long slab = std::max<long long>(1, std::min<long long>(rows, ((size_t)1 << 29) / ((size_t)maxw * 8)));
  float* scratch = nullptr;
  GL_CUDA(h, cudaMallocAsync((void**)&scratch, (size_t)slab * maxw * 2 * sizeof(float), st));
  float* bufA = scratch;
  float* bufB = scratch + (size_t)slab * maxw;
  int rc = GLASDI_OK;
  for (long long r0 = 0; r0 < rows && rc == GLASDI_OK; r0 += slab) {
    const long long R = std::min<long long>(slab, rows - r0);
    const float* Hlast = nullptr;
    rc = run_chain(h, Z + (size_t)r0 * d.widths[0], R, d.n_linear - 1, bufA, bufB, &Hlast, st);
    if (rc) break;
    const int l = d.n_linear - 1;
    dim3 grid((unsigned)((R + simt::BM - 1) / simt::BM), (d.D + simt::BN - 1) / simt::BN);
    simt::linear_act_kernel<<<grid, 256, 0, st>>>(Hlast, d.dW[l], d.db[l], X + (size_t)r0 * d.D, R, d.K, d.D,
                                                  GLASDI_ACT_IDENTITY);
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
  if (e.n_linear > 1) GL_CUDA(h, cudaMallocAsync((void**)&scratch, (size_t)rows * maxw * 2 * sizeof(float), st));
  float* bufs[2] = {scratch, scratch ? scratch + (size_t)rows * maxw : nullptr};
  const float* cur = U;
  int rc = GLASDI_OK;
  for (int l = 0; l < e.n_linear && rc == GLASDI_OK; ++l) {
    const bool last = l == e.n_linear - 1;
    float* dst = last ? Z : bufs[l & 1];
    dim3 grid((unsigned)((rows + simt::BM - 1) / simt::BM), (e.widths[l + 1] + simt::BN - 1) / simt::BN);
    simt::linear_act_kernel<<<grid, 256, 0, st>>>(cur, e.dW[l], e.db[l], dst, rows, e.widths[l], e.widths[l + 1],
                                                  last ? GLASDI_ACT_IDENTITY : e.act);
    h->launches++;
    rc = check_cuda(h, cudaGetLastError(), "encoder linear_act launch");
    cur = dst;
  }
  if (scratch) cudaFreeAsync(scratch, st);
  return rc;
}

}  // namespace glasdi
