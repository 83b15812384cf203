// train_mlp.cu -- the autoencoder's training-mode forward / backward and the reconstruction loss.
//
// Replaces what torch autograd records and replays for the reference's training step (src/lasdi/gplasdi.py:180-182,197):
// `Z = encoder(X); X_pred = decoder(Z); loss_ae = MSE(X, X_pred)` through MultiLayerPerceptron.forward
// (src/lasdi/latent_space.py:157-164: Linear -> act -> ... -> Linear), then `loss.backward()`.
//
// The layers of the shipped autoencoder are [1001 -> 100 -> 5] and back on n_train * nt ~ 4e3 .. 2.5e4 rows: 0.4 .. 2.5 GFLOP
// per GEMM, launch- and latency-bound rather than FLOP-bound, and fp32 (the optimiser's master weights; the reference trains
// in fp32).  One tiled fp32 SIMT GEMM kernel (64 x 64 x 16 tiles, 4 x 4 outputs per thread, operand layouts as template
// flags) serves the three products of a layer -- Y = X W^T + b, dX = dA W, dW = dA^T X -- with the epilogues fused:
// bias + activation (pre- and post-activation saved for the backward), and the activation derivative applied to dY in the
// LOADER of the two backward products (dA is never materialised); the bias gradient is a column sum with a fixed
// summation order, the reconstruction loss and its gradient come out of one pass (fp64 partial sums, fixed order):
// results are bit-reproducible run to run.
#include "common.cuh"

#include <algorithm>

namespace glasdi {
namespace tmlp {

// d act(x) / dx for the activations of common.cuh::act_apply (torch's autograd formulas, default arguments)
__device__ __forceinline__ float act_grad(int act, float x) {
  switch (act) {
    case GLASDI_ACT_SIGMOID: { const float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f - s); }
    case GLASDI_ACT_TANH: { const float t = tanhf(x); return 1.0f - t * t; }
    case GLASDI_ACT_SOFTPLUS: return x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x));
    case GLASDI_ACT_RELU: return x > 0.0f ? 1.0f : 0.0f;
    case GLASDI_ACT_ELU: return x > 0.0f ? 1.0f : expf(x);
    case GLASDI_ACT_SILU: { const float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f + x * (1.0f - s)); }
    case GLASDI_ACT_GELU:
      return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
    case GLASDI_ACT_HARDSHRINK: return fabsf(x) > 0.5f ? 1.0f : 0.0f;
    case GLASDI_ACT_HARDSIGMOID: return (x > -3.0f && x < 3.0f) ? (1.0f / 6.0f) : 0.0f;
    case GLASDI_ACT_HARDTANH: return (x > -1.0f && x < 1.0f) ? 1.0f : 0.0f;
    case GLASDI_ACT_HARDSWISH: return x < -3.0f ? 0.0f : (x > 3.0f ? 1.0f : (2.0f * x + 3.0f) * (1.0f / 6.0f));
    case GLASDI_ACT_LEAKYRELU: return x > 0.0f ? 1.0f : 0.01f;
    case GLASDI_ACT_LOGSIGMOID: return 1.0f / (1.0f + expf(x));
    case GLASDI_ACT_RELU6: return (x > 0.0f && x < 6.0f) ? 1.0f : 0.0f;
    case GLASDI_ACT_SELU: return 1.0507009873554805f * (x > 0.0f ? 1.0f : 1.6732632423543772f * expf(x));
    case GLASDI_ACT_CELU: return x > 0.0f ? 1.0f : expf(x);
    case GLASDI_ACT_MISH: {
      const float sp = x > 20.0f ? x : log1pf(expf(x)), t = tanhf(sp), s = 1.0f / (1.0f + expf(-x));
      return t + x * (1.0f - t * t) * s;
    }
    case GLASDI_ACT_SOFTSHRINK: return fabsf(x) > 0.5f ? 1.0f : 0.0f;
    case GLASDI_ACT_TANHSHRINK: { const float t = tanhf(x); return t * t; }
    default: return 1.0f;
  }
}

constexpr int BM = 64, BN = 64, BK = 16, TT = 256;

// C[m, n] = sum_k A(m, k) B(k, n)  (+ bias[n], activation) with
//   TA = false: A(m, k) = A[m * lda + k]          TA = true: A(m, k) = A[k * lda + m]
//   TB = false: B(k, n) = B[k * ldb + n]          TB = true: B(k, n) = B[n * ldb + k]
// GA: the A operand is dY * act'(pre) formed on the fly (pre has A's layout); FWD: bias + activation epilogue, the
// pre-activation is stored as well (pre_out, same layout as C) when the layer has an activation.
template <bool TA, bool TB, bool GA, bool FWD>
__global__ void __launch_bounds__(TT) gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                  float* __restrict__ C, int ldc, int M, int N, int K,
                                                  const float* __restrict__ preA, int act, const float* __restrict__ bias,
                                                  float* __restrict__ pre_out, int apply_act) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid & 15, ty = tid >> 4;
  // split-K (gridDim.z > 1, only without an epilogue): slice z of the contraction goes to its own partial result
  // C + z * M * ldc; splitk_reduce_kernel adds the slices in a fixed order
  const int kper = ((K + gridDim.z - 1) / gridDim.z + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
  C += (size_t)blockIdx.z * M * ldc;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // 64 x 16 elements of each operand, four per thread; the fast index of the global layout runs over the threads
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * TT;
      int mm, kk;
      if (TA) { mm = idx & (BM - 1); kk = idx >> 6; } else { kk = idx & (BK - 1); mm = idx >> 4; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < kend) {
        const size_t o = TA ? (size_t)k * lda + m : (size_t)m * lda + k;
        v = A[o];
        if (GA) v *= act_grad(act, preA[o]);
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * TT;
      int nn, kk;
      if (TB) { kk = idx & (BK - 1); nn = idx >> 4; } else { nn = idx & (BN - 1); kk = idx >> 6; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < kend) v = TB ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (FWD) {
        v += bias[n];
        if (apply_act) {
          pre_out[(size_t)m * ldc + n] = v;
          v = act_apply(act, v);
        }
      }
      C[(size_t)m * ldc + n] = v;
    }
  }
}

// db[n] = sum_m dY[m, n] * act'(pre[m, n]) (pre == nullptr: plain column sum); one CTA per 32 columns, fixed order
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dY, const float* __restrict__ pre, int M,
                                                       int N, int act, float* __restrict__ db) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), r0 = threadIdx.x >> 5;
  float s = 0.f;
  if (c < N)
    for (int m = r0; m < M; m += 8) {
      const size_t o = (size_t)m * N + c;
      s += pre ? dY[o] * act_grad(act, pre[o]) : dY[o];
    }
  part[r0][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += part[r][threadIdx.x];
    db[c] = t;
  }
}

// partial[b] = sum over the block's elements of (a - b)^2 in fp64; dA = 2 (a - b) / n
constexpr int kMseBlocks = 256;
__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ A, const float* __restrict__ B, long long n,
                                                         float* __restrict__ dA, double* __restrict__ partial) {
  __shared__ double sw[8];
  const float g = 2.0f / (float)n;
  double s = 0.0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float d = A[i] - B[i];
    s += (double)d * (double)d;
    if (dA) dA[i] = g * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sw[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void mse_final_kernel(const double* __restrict__ partial, int nb, long long n, float* __restrict__ loss) {
  double t = 0.0;
  for (int b = 0; b < nb; ++b) t += partial[b];
  *loss = (float)(t / (double)n);
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t n,
                                                           float* __restrict__ C) {
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    float t = 0.f;
    for (int s = 0; s < splits; ++s) t += part[(size_t)s * n + i];
    C[i] = t;
  }
}

// splits > 1: C is a [splits][M][ldc] scratch, reduced by the caller (products whose contraction is the row count: dW)
template <bool TA, bool TB, bool GA, bool FWD>
static void launch_gemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                        const float* preA, int act, const float* bias, float* pre_out, int apply_act, cudaStream_t st,
                        int splits = 1) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  gemm_kernel<TA, TB, GA, FWD><<<grid, TT, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, preA, act, bias, pre_out, apply_act);
}

}  // namespace tmlp

static int check_mlp_args(glasdi_handle* h, int n_linear, const int* widths, const float* const* W, const float* const* b,
                          bool need_b, int act_id, int64_t rows) {
  if (n_linear < 1 || n_linear > 16) return set_error(h, GLASDI_ERR_INVALID, "mlp: 1..16 linear layers");
  if (!widths || !W || (need_b && !b)) return set_error(h, GLASDI_ERR_INVALID, "mlp: null layer arrays");
  if (act_id < 0 || act_id >= GLASDI_ACT_COUNT) return set_error(h, GLASDI_ERR_INVALID, "unknown activation id");
  if (rows < 1 || rows > 2000000000ll / 64) return set_error(h, GLASDI_ERR_INVALID, "mlp: bad row count");
  for (int l = 0; l <= n_linear; ++l)
    if (widths[l] < 1) return set_error(h, GLASDI_ERR_INVALID, "non-positive layer width");
  for (int l = 0; l < n_linear; ++l)
    if (!W[l] || (need_b && !b[l])) return set_error(h, GLASDI_ERR_INVALID, "mlp: null weight pointer");
  return GLASDI_OK;
}

}  // namespace glasdi

using namespace glasdi;

extern "C" {

int glasdi_mlp_forward_train(glasdi_handle* h, int n_linear, const int* widths, const float* const* W_dev,
                             const float* const* b_dev, int act_id, const float* X, int64_t rows, float* pre_out,
                             float* post_out, float* Y_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  int rc = check_mlp_args(h, n_linear, widths, W_dev, b_dev, true, act_id, rows);
  if (rc) return rc;
  if (!X || !Y_out || (n_linear > 1 && (!pre_out || !post_out))) return set_error(h, GLASDI_ERR_INVALID, "mlp forward: null pointer");
  DeviceGuard device_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const float* in = X;
  size_t off = 0;
  for (int l = 0; l < n_linear; ++l) {
    const bool last = l == n_linear - 1;
    const int wi = widths[l], wo = widths[l + 1];
    float* out = last ? Y_out : post_out + off;
    tmlp::launch_gemm<false, true, false, true>(in, wi, W_dev[l], wi, out, wo, (int)rows, wo, wi, nullptr, act_id, b_dev[l],
                                              last ? nullptr : pre_out + off, last ? 0 : 1, st);
    h->launches++;
    in = out;
    if (!last) off += (size_t)rows * wo;
  }
  return check_cuda(h, cudaGetLastError(), "mlp forward launch");
}

int glasdi_mlp_backward(glasdi_handle* h, int n_linear, const int* widths, const float* const* W_dev, int act_id,
                        const float* X, const float* pre, const float* post, const float* dY, int64_t rows, float* dX_out,
                        float* const* dW_out, float* const* db_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!dW_out || !db_out) return set_error(h, GLASDI_ERR_INVALID, "mlp backward: null gradient arrays");
  int rc = check_mlp_args(h, n_linear, widths, W_dev, nullptr, false, act_id, rows);
  if (rc) return rc;
  if (!X || !dY || (n_linear > 1 && (!pre || !post))) return set_error(h, GLASDI_ERR_INVALID, "mlp backward: null pointer");
  DeviceGuard device_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: two ping-pong gradient buffers of rows x (widest hidden layer)
  int maxw = 1;
  for (int l = 1; l < n_linear; ++l) maxw = std::max(maxw, widths[l]);
  float* scratch = nullptr;
  if (n_linear > 1) GL_CUDA(h, cudaMallocFromPoolAsync((void**)&scratch, (size_t)2 * rows * maxw * sizeof(float), h->pool, st));
  float* gbuf[2] = {scratch, scratch ? scratch + (size_t)rows * maxw : nullptr};
  std::vector<size_t> offs(n_linear, 0);                      // offset of layer l's OUTPUT in pre / post (l < n_linear - 1)
  {
    size_t off = 0;
    for (int l = 0; l + 1 < n_linear; ++l) { offs[l] = off; off += (size_t)rows * widths[l + 1]; }
  }
  const float* g = dY;                                         // gradient w.r.t. the layer's OUTPUT (post-activation)
  for (int l = n_linear - 1; l >= 0; --l) {
    const bool last = l == n_linear - 1;
    const int wi = widths[l], wo = widths[l + 1];
    const float* in = l == 0 ? X : post + offs[l - 1];         // the layer's input
    const float* pre_l = last ? nullptr : pre + offs[l];       // its pre-activation (hidden layers)
    if (!dW_out[l] || !db_out[l]) return set_error(h, GLASDI_ERR_INVALID, "mlp backward: null gradient pointer");
    // dW[o, i] = sum_r dA[r, o] in[r, i];  db[o] = sum_r dA[r, o];  dA = g * act'(pre)
    // the contraction runs over the rows (thousands) while dW has a handful of tiles: split it over enough CTAs for two
    // waves, partial products in a scratch, fixed-order reduction (no atomics)
    const int tiles = ((wo + tmlp::BM - 1) / tmlp::BM) * ((wi + tmlp::BN - 1) / tmlp::BN);
    int splits = std::max(1, std::min({64, (2 * h->num_sms + tiles - 1) / tiles, (int)((rows + 4 * tmlp::BK - 1) / (4 * tmlp::BK))}));
    float* part = dW_out[l];
    if (splits > 1) GL_CUDA(h, cudaMallocFromPoolAsync((void**)&part, (size_t)splits * wo * wi * sizeof(float), h->pool, st));
    if (last) tmlp::launch_gemm<true, false, false, false>(g, wo, in, wi, part, wi, wo, wi, (int)rows, nullptr, act_id, nullptr, nullptr, 0, st, splits);
    else tmlp::launch_gemm<true, false, true, false>(g, wo, in, wi, part, wi, wo, wi, (int)rows, pre_l, act_id, nullptr, nullptr, 0, st, splits);
    if (splits > 1) {
      const size_t nel = (size_t)wo * wi;
      tmlp::splitk_reduce_kernel<<<(unsigned)std::min<size_t>((nel + 255) / 256, 1024), 256, 0, st>>>(part, splits, nel, dW_out[l]);
      cudaFreeAsync(part, st);
      h->launches++;
    }
    tmlp::bias_grad_kernel<<<(wo + 31) / 32, 256, 0, st>>>(g, pre_l, (int)rows, wo, act_id, db_out[l]);
    h->launches += 2;
    float* gin = l == 0 ? dX_out : gbuf[l & 1];
    if (gin) {                                                  // dIn[r, i] = sum_o dA[r, o] W[o, i]
      if (last) tmlp::launch_gemm<false, false, false, false>(g, wo, W_dev[l], wi, gin, wi, (int)rows, wi, wo, nullptr, act_id, nullptr, nullptr, 0, st);
      else tmlp::launch_gemm<false, false, true, false>(g, wo, W_dev[l], wi, gin, wi, (int)rows, wi, wo, pre_l, act_id, nullptr, nullptr, 0, st);
      h->launches++;
    }
    g = gin;
  }
  if (scratch) cudaFreeAsync(scratch, st);
  return check_cuda(h, cudaGetLastError(), "mlp backward launch");
}

int glasdi_mse_loss_grad(glasdi_handle* h, const float* A, const float* B, int64_t n, float* loss_out, float* dA_out,
                         void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!A || !B || !loss_out || n < 1) return set_error(h, GLASDI_ERR_INVALID, "mse: bad arguments");
  DeviceGuard device_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = nullptr;
  GL_CUDA(h, cudaMallocFromPoolAsync((void**)&partial, tmlp::kMseBlocks * sizeof(double), h->pool, st));
  const int nb = (int)std::min<long long>(tmlp::kMseBlocks, (n + 255) / 256);
  tmlp::mse_partial_kernel<<<nb, 256, 0, st>>>(A, B, n, dA_out, partial);
  tmlp::mse_final_kernel<<<1, 1, 0, st>>>(partial, nb, n, loss_out);
  h->launches += 2;
  cudaFreeAsync(partial, st);
  return check_cuda(h, cudaGetLastError(), "mse launch");
}

}  // extern "C"
