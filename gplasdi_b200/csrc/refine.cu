// refine.cu -- second half of the SCREENED greedy step (GLASDI_MODE_SCREENED_GRAM / GLASDI_MODE_SCREENED).
//
// The tensor-core screening pass (decode_gram.cu: covariance form; decode_var2.cu: one fp16 pass over the decoded
// ensemble) writes the approximate variance of every (t, dof) next to the per-parameter maximum.  The maximum of
// the reference's fp32 result can only sit where the screened variance is within a factor (1 - delta)^2 of the
// screened maximum as long as the screening error is below delta / 2 (measured: 5e-5 relative on sigma for the
// random-weight bench decoder, 5e-4 for a smooth one; delta = 2^-8 by default).  Those candidates are re-evaluated
// here on the CUDA cores -- front MLP in fp32 exactly as the reference does (src/lasdi/latent_space.py:157-164),
// last layer as an fp32 FMA chain on the pilot-centred activations, sample sums in fp64 -- and the maximum over the
// candidates of a parameter is its max variance (reference src/lasdi/gplasdi.py:65-69).
// The deviation |sigma_screen - sigma_exact| / sigma_exact of every candidate is checked against delta / 4; a
// violation raises a device flag (GLASDI_ERR_NUMERIC from glasdi_check_numeric), it is never silently accepted.
//
// Candidates are collected PER (parameter, time step) ROW of the field: trained decoders are smooth in the dof
// index, so a row that holds one near-maximum dof holds dozens of them (72 candidates per parameter for a smooth
// last layer, 6 for the random-weight bench decoder).  All candidates of a row share the hidden activations of its
// S samples; one warp evaluates them once and dots them with up to 8 rows of the last layer at a time.
#include "decode_common.cuh"

#include <algorithm>

namespace glasdi {
namespace rf {

using namespace dv;

constexpr int kSelWarps = 8;          // rows per block of select_rows_kernel
constexpr int kRefWarps = 8;
constexpr int kNC = 8;                // candidates a warp evaluates per pass over the samples

// one warp per (p, t) row of the field: candidates = dofs whose screened variance is >= keep * screened max of p
__global__ void __launch_bounds__(kSelWarps * 32) select_rows_kernel(const float* __restrict__ field, int field_ld,
                                                                    const unsigned int* __restrict__ approx_bits,
                                                                    long long n_rows, int T, int D, float keep,
                                                                    const float* __restrict__ cnorm,
                                                                    const float* __restrict__ w4norm, float beta,
                                                                    float keep2,
                                                                    Cand* __restrict__ cand, unsigned int cap,
                                                                    Group* __restrict__ groups, unsigned int gcap,
                                                                    unsigned int* __restrict__ stats,
                                                                    int* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kSelWarps + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int p = (int)(row / T), t = (int)(row - (long long)p * T);
  const float vmax = __uint_as_float(approx_bits[p]);
  if (!(vmax > 0.f)) return;
  const float thr = vmax * keep;
  // A-priori bound of the screening's rounding error (covariance form): the fp16 roundings of the covariance entries and of
  // the products w_i w_j are independent, relative 2^-11 each, so the error of sigma^2 = sum_ij C_ij w_i w_j has a standard
  // deviation <= 2^-11 * 1.15 * sum_i C_ii w_i^2 <= 2^-11 * 1.15 * |diag C|_2 |w o w|_2 (|C_ij| <= sqrt(C_ii C_jj), Cauchy-
  // Schwarz).  A (t, dof) whose screened variance PLUS beta |diag C[p,t]|_2 |w_dof o w_dof|_2 (beta = the configured number
  // of sigmas) reaches (1 - delta/2)^2 of the screened maximum is a candidate too: a dof whose quadratic form cancels badly
  // (large weights nearly orthogonal to the covariance's dominant directions) is re-evaluated exactly instead of trusted.
  const float brow = (cnorm != nullptr) ? beta * __ldg(cnorm + row) : 0.f;
  const float thr2 = (cnorm != nullptr) ? vmax * keep2 : 3.0e38f;
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(w4norm);
  auto is_cand = [&](float v, float om) { return v >= thr || fmaf(brow, om, v) >= thr2; };
  const float4* __restrict__ f4 = reinterpret_cast<const float4*>(field + (size_t)row * field_ld);
  const int n4 = field_ld / 4;                                      // field_ld is a multiple of 256
  // pass 1: count
  int mine = 0;
  for (int i4 = lane; i4 < n4; i4 += 32) {
    const float4 v = __ldg(f4 + i4);
    const float4 om = (cnorm != nullptr) ? __ldg(w4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int d0 = i4 * 4;
    mine += (is_cand(v.x, om.x) && d0 < D) + (is_cand(v.y, om.y) && d0 + 1 < D) + (is_cand(v.z, om.z) && d0 + 2 < D) +
            (is_cand(v.w, om.w) && d0 + 3 < D);
  }
  int incl = mine;                                                  // inclusive prefix sum over the lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  // one work unit (= one warp of refine_kernel) per kNC candidates of the row
  const unsigned int nunits = ((unsigned int)total + kNC - 1) / kNC;
  unsigned int base = 0, gslot = 0;
  if (lane == 0) {
    base = atomicAdd(stats + 0, (unsigned int)total);
    atomicAdd(stats + 1, (unsigned int)total);
    gslot = atomicAdd(stats + 3, nunits);
    if (base + (unsigned int)total > cap || gslot + nunits > gcap) atomicOr(flags + 1, 1);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  gslot = __shfl_sync(0xffffffffu, gslot, 0);
  if (base + (unsigned int)total > cap || gslot + nunits > gcap) return;
  for (unsigned int u = lane; u < nunits; u += 32)
    groups[gslot + u] = Group{p, t, base + u * kNC,
                              min((unsigned int)kNC, (unsigned int)total - u * kNC) | (u == 0 ? (unsigned int)total << 8 : 0u)};
  // pass 2: write (the row is in L1/L2); lanes own interleaved float4s, so the order inside a row is by lane
  unsigned int w = base + (unsigned int)(incl - mine);
  for (int i4 = lane; i4 < n4; i4 += 32) {
    const float4 v = __ldg(f4 + i4);
    const float4 om = (cnorm != nullptr) ? __ldg(w4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const float oo[4] = {om.x, om.y, om.z, om.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (is_cand(vv[q], oo[q]) && i4 * 4 + q < D) {
        if (vv[q] < thr) atomicAdd(stats + 4, 1u);                  // admitted by the bound alone (diagnostic)
        cand[w++] = Cand{p, t, i4 * 4 + q, vv[q]};
      }
  }
}

struct RefParams {
  FrontMlp front;
  const float* Wlast;               // [D][K] fp32, nn.Linear layout
  const float* Zs;                  // [P][T][NZ][Spad]
  const Cand* cand;
  const Group* groups;
  const unsigned int* stats;
  unsigned int gcap;
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

// One CTA per work unit (<= NC candidates of one (p, t) row): thread <-> sample (stride 256), so a unit costs
// S / 256 activation sweeps of latency instead of S / 32 -- a slab with few candidates is no longer one warp's serial
// time.  The hidden activations of a sample are evaluated once and dotted with the NC rows of the last layer; the
// pilot activations are made by one thread per unit group, sample sums in fp64 (warp shuffle, then a fixed-order sum
// over the warps: bit-reproducible for any grid).  Two instantiations share the unit list: NC = 2 takes the units
// with one or two candidates (64 registers, four CTAs per SM -- the kernel is bound by the latency of the libdevice
// activation chains, so resident warps are what counts), NC = kNC the fuller ones.
template <int ACT, int INW, int NC>
__global__ void __launch_bounds__(kRefWarps * 32, NC <= 2 ? 4 : 2) refine_kernel(const __grid_constant__ RefParams R) {
  extern __shared__ __align__(16) float rsm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int last_in_w = R.front.widths[R.front.n_front - 1];
  const int ld = wl_stride(last_in_w);
  float* sWl = rsm;                                  // [Kpad][ld] rows [w.., bias, 0..] of the last FRONT layer
  float* sWd = sWl + R.Kpad * ld;                    // [NC][Kpad] rows of the LAST layer of the unit's candidates
  float* sPil = sWd + NC * R.Kpad;                   // [Kpad] pilot activations of the row
  float* sXabs = sPil + R.Kpad;                      // [8] sum_k |w_k| |h_k(pilot)| per candidate
  double* sRed = reinterpret_cast<double*>(sXabs + 8);   // [kRefWarps][NC][2]
  {
    const int l = R.front.n_front - 1;
    const float* __restrict__ W = R.front.W[l];
    const float* __restrict__ b = R.front.b[l];
    for (int idx = tid; idx < R.Kpad * ld; idx += blockDim.x) {
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
  const unsigned int nunits = min(R.stats[3], R.gcap);
  const int nkg = R.Kpad / 8;
  const double inv_S = 1.0 / (double)R.S;
  for (unsigned int gi = blockIdx.x; gi < nunits; gi += gridDim.x) {
    const Group gr = R.groups[gi];
    const int nc = (int)(gr.count & 0xffu);
    if ((nc <= 2) != (NC <= 2)) continue;              // uniform over the CTA
    const Cand* __restrict__ cd = R.cand + gr.first;
    const float* __restrict__ zb = R.Zs + ((size_t)gr.p * R.T + gr.t) * R.NZ * R.Spad;
    // rows of the last layer of the unit's candidates (unused slots: zero rows)
    for (int idx = tid; idx < NC * R.Kpad; idx += kRefWarps * 32) {
      const int c = idx / R.Kpad, k = idx - c * R.Kpad;
      sWd[idx] = (c < nc && k < R.K) ? __ldg(R.Wlast + (size_t)cd[c].d * R.K + k) : 0.f;
    }
    if (tid < nkg) {   // pilot activations of this row (sample 0), one unit group per thread
      float zz[kMaxLatent];
#pragma unroll
      for (int i = 0; i < kMaxLatent; ++i) zz[i] = (i < R.NZ) ? __ldg(zb + (size_t)i * R.Spad) : 0.f;
      float xp[INW > 0 ? 1 : kMaxHid];
      const float* xin = zz;
      if constexpr (INW == 0) {
        front_pre<ACT>(R.front, zz, xp);
        xin = xp;
      }
      float hv[8];
      last_front_units<ACT, INW>(sWl, ld, last_in_w, R.front.act, tid * 8, xin, hv);
#pragma unroll
      for (int u = 0; u < 8; ++u) sPil[tid * 8 + u] = hv[u];
    }
    __syncthreads();
    if (warp < NC) {   // scale of the fp32 rounding noise of a decoded value, one candidate per warp
      float a = 0.f;
      for (int k = lane; k < R.Kpad; k += 32) a = fmaf(fabsf(sWd[warp * R.Kpad + k]), fabsf(sPil[k]), a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) sXabs[warp] = a;
    }
    double s1[NC], s2[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { s1[c] = 0.0; s2[c] = 0.0; }
    for (int s = tid; s < R.S; s += kRefWarps * 32) {
      float zz[kMaxLatent];
#pragma unroll
      for (int i = 0; i < kMaxLatent; ++i) zz[i] = (i < R.NZ) ? __ldg(zb + (size_t)i * R.Spad + s) : 0.f;
      float xp[INW > 0 ? 1 : kMaxHid];
      const float* xin = zz;
      if constexpr (INW == 0) {
        front_pre<ACT>(R.front, zz, xp);
        xin = xp;
      }
      float dsum[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) dsum[c] = 0.f;
      for (int kg = 0; kg < nkg; ++kg) {
        float hv[8];
        last_front_units<ACT, INW>(sWl, ld, last_in_w, R.front.act, kg * 8, xin, hv);
        const float4 p0 = *reinterpret_cast<const float4*>(sPil + kg * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(sPil + kg * 8 + 4);
        const float dh[8] = {hv[0] - p0.x, hv[1] - p0.y, hv[2] - p0.z, hv[3] - p0.w,
                             hv[4] - p1.x, hv[5] - p1.y, hv[6] - p1.z, hv[7] - p1.w};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float4 w0 = *reinterpret_cast<const float4*>(sWd + c * R.Kpad + kg * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(sWd + c * R.Kpad + kg * 8 + 4);
          float d = dsum[c];                          // same FMA order as one chain over k = 0 .. K-1
          d = fmaf(w0.x, dh[0], d);
          d = fmaf(w0.y, dh[1], d);
          d = fmaf(w0.z, dh[2], d);
          d = fmaf(w0.w, dh[3], d);
          d = fmaf(w1.x, dh[4], d);
          d = fmaf(w1.y, dh[5], d);
          d = fmaf(w1.z, dh[6], d);
          d = fmaf(w1.w, dh[7], d);
          dsum[c] = d;
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        s1[c] += (double)dsum[c];
        s2[c] = fma((double)dsum[c], (double)dsum[c], s2[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double t1 = warp_sum(s1[c]), t2 = warp_sum(s2[c]);
      if (lane == 0) {
        sRed[(warp * NC + c) * 2 + 0] = t1;
        sRed[(warp * NC + c) * 2 + 1] = t2;
      }
    }
    __syncthreads();
    if (tid < nc) {
      const int c = tid;
      double t1 = 0.0, t2 = 0.0;
#pragma unroll
      for (int w = 0; w < kRefWarps; ++w) {
        t1 += sRed[(w * NC + c) * 2 + 0];
        t2 += sRed[(w * NC + c) * 2 + 1];
      }
      const double mean = t1 * inv_S;
      const float var = (float)fmax(t2 * inv_S - mean * mean, 0.0);
      atomicMax(R.exact_bits + gr.p, __float_as_uint(var));
      // Screening check.  Below ~2^-24 * sum|w||h| (half an ulp of the decoded value's terms) a spread is fp32
      // rounding noise in the reference as much as here (sigma of near-identical samples, e.g. at the GP's
      // training points, where the activations h(s) - h(pilot) themselves carry 1e-2 relative rounding error):
      // deviations inside that floor say nothing about the screening and are not counted.
      const float se = sqrtf(var), sa = sqrtf(fmaxf(cd[c].var, 0.f) * R.unscale);
      const float excess = fabsf(sa - se) - 6.0e-8f * sXabs[c];
      const float dev = (se > 0.f && excess > 0.f) ? excess / se : 0.f;
      atomicMax(R.dev_bits, __float_as_uint(dev));
      if (dev > R.dev_tol) atomicOr(R.flags + 2, 1);
    }
    __syncthreads();                                   // the unit's shared-memory rows are free again
  }
}

}  // namespace rf

int launch_select_candidates(glasdi_handle* h, const float* var_field, int field_ld, const unsigned int* approx_bits,
                             int P, int T, int D, float keep, const float* cnorm, const float* w4norm, float beta,
                             float keep2, cudaStream_t st) {
  if (!cnorm || !w4norm || !(beta > 0.f)) { cnorm = nullptr; w4norm = nullptr; }
  if (field_ld > ((D + 255) / 256) * 256) return set_error(h, GLASDI_ERR_INVALID, "select: field wider than the bound table");
  const long long n_rows = (long long)P * T;
  const unsigned int blocks = (unsigned int)((n_rows + rf::kSelWarps - 1) / rf::kSelWarps);
  rf::select_rows_kernel<<<blocks, rf::kSelWarps * 32, 0, st>>>(
      var_field, field_ld, approx_bits, n_rows, T, D, keep, cnorm, w4norm, beta, keep2, reinterpret_cast<rf::Cand*>(h->d_cand),
      (unsigned)h->cand_cap, reinterpret_cast<rf::Group*>(h->d_groups), (unsigned)h->group_cap, h->d_cand_stats,
      h->d_flags);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "select_rows_kernel launch");
}

template <int ACT, int INW>
static int launch_refine_inst(glasdi_handle* h, const rf::RefParams& R, size_t smem, cudaStream_t st) {
  GL_CUDA(h, cudaFuncSetAttribute(rf::refine_kernel<ACT, INW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GL_CUDA(h, cudaFuncSetAttribute(rf::refine_kernel<ACT, INW, rf::kNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rf::refine_kernel<ACT, INW, 2><<<h->num_sms * 4, rf::kRefWarps * 32, smem, st>>>(R);
  rf::refine_kernel<ACT, INW, rf::kNC><<<h->num_sms * 2, rf::kRefWarps * 32, smem, st>>>(R);
  h->launches++;
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
  R.groups = reinterpret_cast<const rf::Group*>(h->d_groups);
  R.stats = h->d_cand_stats;
  R.gcap = (unsigned)h->group_cap;
  R.exact_bits = exact_bits;
  R.dev_bits = h->d_cand_stats + 2;
  R.flags = h->d_flags;
  R.K = d.K; R.Kpad = d.Kpad; R.NZ = d.widths[0]; R.S = S; R.T = T; R.Spad = Spad;
  R.unscale = unscale;              // screened variance -> true units
  R.dev_tol = dev_tol;
  const int last_in_w = d.widths[d.n_linear - 2];
  const size_t smem = ((size_t)d.Kpad * dv::wl_stride(last_in_w) + (size_t)(rf::kNC + 1) * d.Kpad + 8) * sizeof(float) +
                      (size_t)rf::kRefWarps * rf::kNC * 2 * sizeof(double);
  int rc;
  // accurate libdevice activations (ACT >= 0 resolves the switch at compile time; -1 = runtime switch)
  if (R.front.n_front == 1 && R.NZ == 5) rc = launch_refine_inst<-1, 5>(h, R, smem, st);
  else rc = launch_refine_inst<-1, 0>(h, R, smem, st);
  if (rc) return rc;
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "refine_kernel launch");
}

}  // namespace glasdi
