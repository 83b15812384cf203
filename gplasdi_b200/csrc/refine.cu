// refine.cu -- second half of the SCREENED greedy step (GLASDI_OPT_MODE = 1).
//
// The tcgen05 kernel (decode_var2.cu) is run with ONE fp16 pass (W_hi * H_hi: ~2^-11 per operand) and
// writes the approximate variance of every (t, dof) next to the per-parameter maximum.  The maximum of
// the reference's fp32 result can only sit where the screened variance is within a factor (1 - delta)^2
// of the screened maximum as long as the screening error is below delta / 2 (measured: 7e-5 relative on
// sigma at the bench shape; delta = 2^-8 by default).  Those few candidates are re-evaluated here on the
// CUDA cores -- front MLP in fp32 exactly as the reference does (src/lasdi/latent_space.py:157-164),
// last layer as an fp32 FMA chain on the pilot-centred activations, sample sums in fp64 -- and the
// maximum over the candidates of a parameter is its max variance (reference src/lasdi/gplasdi.py:65-69).
// The deviation |sigma_screen - sigma_exact| / sigma_exact of every candidate is checked against
// delta / 4; a violation raises a device flag (GLASDI_ERR_NUMERIC from glasdi_check_numeric), it is never
// silently accepted.
#include "decode_common.cuh"

#include <algorithm>

namespace glasdi {
namespace rf {

using namespace dv;

struct Cand {
  int p, t, d;
  float var;       // screened variance (scaled units)
};

constexpr int kSelThreads = 256;
constexpr int kSelPerThread = 16;     // floats per thread (4 x float4)
constexpr int kRefWarps = 8;

// candidates of parameter p: every (t, dof) whose screened variance is >= keep * screened max of p
__global__ void __launch_bounds__(kSelThreads) select_kernel(const float* __restrict__ field, int field_ld,
                                                             const unsigned int* __restrict__ approx_bits, int T,
                                                             int D, float keep, Cand* __restrict__ cand,
                                                             unsigned int cap, unsigned int* __restrict__ stats,
                                                             int* __restrict__ flags) {
  const int p = blockIdx.y;
  const float vmax = __uint_as_float(approx_bits[p]);
  if (!(vmax > 0.f)) return;
  const float thr = vmax * keep;
  const long long n4 = (long long)T * field_ld / 4;                 // field_ld is a multiple of 256
  const float4* __restrict__ f4 = reinterpret_cast<const float4*>(field + (size_t)p * T * field_ld);
  const long long base = (long long)blockIdx.x * (kSelThreads * kSelPerThread / 4);
#pragma unroll
  for (int u = 0; u < kSelPerThread / 4; ++u) {
    const long long i4 = base + u * kSelThreads + threadIdx.x;
    if (i4 >= n4) break;
    const float4 v = __ldg(f4 + i4);
    const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (m >= thr) {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      const long long e0 = i4 * 4;
      const int t = (int)(e0 / field_ld);
      const int d0 = (int)(e0 - (long long)t * field_ld);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (vv[q] >= thr && d0 + q < D) {
          const unsigned int slot = atomicAdd(stats + 0, 1u);
          atomicAdd(stats + 1, 1u);
          if (slot < cap) cand[slot] = Cand{p, t, d0 + q, vv[q]};
          else atomicOr(flags + 1, 1);
        }
      }
    }
  }
}

struct RefParams {
  FrontMlp front;
  const float* Wlast;               // [D][K] fp32, nn.Linear layout
  const float* Zs;                  // [P][T][NZ][Spad]
  const Cand* cand;
  const unsigned int* stats;
  unsigned int cap;
  unsigned int* exact_bits;         // [P] float bits of the exact variance (true units)
  unsigned int* dev_bits;           // float bits of the max relative deviation screen vs exact
  int* flags;
  int K, Kpad, NZ, S, T, Spad;
  float unscale;                    // screened variance -> true units (2^(-2 (eW + eH)) for the x-space screening)
  float dev_tol;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per candidate; lanes stride the samples
template <int ACT, int INW>
__global__ void __launch_bounds__(kRefWarps * 32) refine_kernel(const __grid_constant__ RefParams R) {
  extern __shared__ __align__(16) float rsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int last_in_w = R.front.widths[R.front.n_front - 1];
  const int ld = wl_stride(last_in_w);
  float* sWl = rsm;                                    // [Kpad][ld] rows [w.., bias, 0..] of the last FRONT layer
  float* sWd = rsm + R.Kpad * ld + warp * 2 * R.Kpad;  // this warp's row of the LAST layer
  float* sPil = sWd + R.Kpad;                          // this warp's pilot activations
  {
    const int l = R.front.n_front - 1;
    const float* __restrict__ W = R.front.W[l];
    const float* __restrict__ b = R.front.b[l];
    for (int idx = threadIdx.x; idx < R.Kpad * ld; idx += blockDim.x) {
      const int k = idx / ld, i = idx - k * ld;
      float v = 0.f;
      if (k < R.K) {
        if (i < last_in_w) v = W[(size_t)k * last_in_w + i];
        else if (i == last_in_w) v = b[k];
      }
      sWl[idx] = v;
    }
  }
  __syncthreads();
  const unsigned int n = min(R.stats[0], R.cap);
  const int nkg = R.Kpad / 8;
  const double inv_S = 1.0 / (double)R.S;
  for (unsigned int c = blockIdx.x * kRefWarps + warp; c < n; c += gridDim.x * kRefWarps) {
    const Cand cd = R.cand[c];
    for (int k = lane; k < R.Kpad; k += 32) sWd[k] = k < R.K ? __ldg(R.Wlast + (size_t)cd.d * R.K + k) : 0.f;
    const float* __restrict__ zb = R.Zs + ((size_t)cd.p * R.T + cd.t) * R.NZ * R.Spad;
    __syncwarp();
    float xabs = 0.f;                 // sum_k |w_k| |h_k(pilot)|: scale of the fp32 rounding noise of a decoded value
    {
      float zz[kMaxLatent];
#pragma unroll
      for (int i = 0; i < kMaxLatent; ++i) zz[i] = (i < R.NZ) ? __ldg(zb + (size_t)i * R.Spad) : 0.f;   // sample 0
      float xp[INW > 0 ? 1 : kMaxHid];
      const float* xin = zz;
      if constexpr (INW == 0) {
        front_pre<ACT>(R.front, zz, xp);
        xin = xp;
      }
      for (int kg = lane; kg < nkg; kg += 32) {
        float hv[8];
        last_front_units<ACT, INW>(sWl, ld, last_in_w, R.front.act, kg * 8, xin, hv);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          sPil[kg * 8 + u] = hv[u];
          xabs = fmaf(fabsf(sWd[kg * 8 + u]), fabsf(hv[u]), xabs);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xabs += __shfl_xor_sync(0xffffffffu, xabs, o);
    __syncwarp();
    double s1 = 0.0, s2 = 0.0;
    for (int s = lane; s < R.S; s += 32) {
      float zz[kMaxLatent];
#pragma unroll
      for (int i = 0; i < kMaxLatent; ++i) zz[i] = (i < R.NZ) ? __ldg(zb + (size_t)i * R.Spad + s) : 0.f;
      float xp[INW > 0 ? 1 : kMaxHid];
      const float* xin = zz;
      if constexpr (INW == 0) {
        front_pre<ACT>(R.front, zz, xp);
        xin = xp;
      }
      float dsum = 0.f;
      for (int kg = 0; kg < nkg; ++kg) {
        float hv[8];
        last_front_units<ACT, INW>(sWl, ld, last_in_w, R.front.act, kg * 8, xin, hv);
        const float4 w0 = *reinterpret_cast<const float4*>(sWd + kg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sWd + kg * 8 + 4);
        const float4 p0 = *reinterpret_cast<const float4*>(sPil + kg * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(sPil + kg * 8 + 4);
        dsum = fmaf(w0.x, hv[0] - p0.x, dsum);
        dsum = fmaf(w0.y, hv[1] - p0.y, dsum);
        dsum = fmaf(w0.z, hv[2] - p0.z, dsum);
        dsum = fmaf(w0.w, hv[3] - p0.w, dsum);
        dsum = fmaf(w1.x, hv[4] - p1.x, dsum);
        dsum = fmaf(w1.y, hv[5] - p1.y, dsum);
        dsum = fmaf(w1.z, hv[6] - p1.z, dsum);
        dsum = fmaf(w1.w, hv[7] - p1.w, dsum);
      }
      s1 += (double)dsum;
      s2 = fma((double)dsum, (double)dsum, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      const double mean = s1 * inv_S;
      const float var = (float)fmax(s2 * inv_S - mean * mean, 0.0);
      atomicMax(R.exact_bits + cd.p, __float_as_uint(var));
      // Screening check.  Below ~2^-24 * sum|w||h| (half an ulp of the decoded value's terms) a spread is fp32
      // rounding noise in the reference as much as here (sigma of near-identical samples, e.g. at the GP's
      // training points, where the activations h(s) - h(pilot) themselves carry 1e-2 relative rounding error):
      // deviations inside that floor say nothing about the screening and are not counted.
      const float se = sqrtf(var), sa = sqrtf(fmaxf(cd.var, 0.f) * R.unscale);
      const float excess = fabsf(sa - se) - 6.0e-8f * xabs;
      const float dev = (se > 0.f && excess > 0.f) ? excess / se : 0.f;
      atomicMax(R.dev_bits, __float_as_uint(dev));
      if (dev > R.dev_tol) atomicOr(R.flags + 2, 1);
    }
    __syncwarp();
  }
}

}  // namespace rf

int launch_select_candidates(glasdi_handle* h, const float* var_field, int field_ld, const unsigned int* approx_bits,
                             int P, int T, int D, float keep, cudaStream_t st) {
  const long long n4 = (long long)T * field_ld / 4;
  const int per_block = rf::kSelThreads * rf::kSelPerThread / 4;
  dim3 grid((unsigned)((n4 + per_block - 1) / per_block), (unsigned)P);
  rf::select_kernel<<<grid, rf::kSelThreads, 0, st>>>(var_field, field_ld, approx_bits, T, D, keep,
                                                     reinterpret_cast<rf::Cand*>(h->d_cand), (unsigned)h->cand_cap,
                                                     h->d_cand_stats, h->d_flags);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "select_kernel launch");
}

template <int ACT, int INW>
static int launch_refine_inst(glasdi_handle* h, const rf::RefParams& R, size_t smem, cudaStream_t st) {
  GL_CUDA(h, cudaFuncSetAttribute(rf::refine_kernel<ACT, INW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rf::refine_kernel<ACT, INW><<<h->num_sms * 4, rf::kRefWarps * 32, smem, st>>>(R);
  return GLASDI_OK;
}

int launch_refine(glasdi_handle* h, const float* Zs, int Spad, int S, int T, unsigned int* exact_bits, float dev_tol,
                  float unscale, cudaStream_t st) {
  const Decoder& d = h->dec;
  rf::RefParams R{};
  R.front.n_front = d.n_linear - 1;
  R.front.act = d.act;
  for (int l = 0; l <= R.front.n_front; ++l) R.front.widths[l] = d.widths[l];
  for (int l = 0; l < R.front.n_front; ++l) { R.front.W[l] = d.dW[l]; R.front.b[l] = d.db[l]; }
  R.Wlast = d.dW[d.n_linear - 1];
  R.Zs = Zs;
  R.cand = reinterpret_cast<const rf::Cand*>(h->d_cand);
  R.stats = h->d_cand_stats;
  R.cap = (unsigned)h->cand_cap;
  R.exact_bits = exact_bits;
  R.dev_bits = h->d_cand_stats + 2;
  R.flags = h->d_flags;
  R.K = d.K; R.Kpad = d.Kpad; R.NZ = d.widths[0]; R.S = S; R.T = T; R.Spad = Spad;
  R.unscale = unscale;              // screened variance -> true units
  R.dev_tol = dev_tol;
  const int last_in_w = d.widths[d.n_linear - 2];
  const size_t smem = ((size_t)d.Kpad * dv::wl_stride(last_in_w) + (size_t)rf::kRefWarps * 2 * d.Kpad) * sizeof(float);
  int rc;
  // accurate libdevice activations (ACT >= 0 resolves the switch at compile time; -1 = runtime switch)
  if (R.front.n_front == 1 && R.NZ == 5) rc = launch_refine_inst<-1, 5>(h, R, smem, st);
  else rc = launch_refine_inst<-1, 0>(h, R, smem, st);
  if (rc) return rc;
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "refine_kernel launch");
}

}  // namespace glasdi
