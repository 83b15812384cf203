// reduce.cu -- finalisation of the greedy step: variance bits -> std, first-strict-maximum argmax,
// and the packed max-loc key for the multi-GPU all-reduce.
//
// Reference semantics (src/lasdi/gplasdi.py:62-72): `max_std = 0; for m: if max_std_m > max_std:
// m_index = m; max_std = max_std_m` -> the SMALLEST index attaining the maximum, provided it is > 0.
#include "common.cuh"

#include <algorithm>

namespace glasdi {

struct ValIdx {
  float v;
  long long i;
};
__device__ __forceinline__ ValIdx better(ValIdx a, ValIdx b) {
  // larger value wins; ties -> smaller index.  NaN never wins (comparisons false), like `>` in python.
  if (b.v > a.v || (b.v == a.v && b.i >= 0 && (a.i < 0 || b.i < a.i))) return b;
  return a;
}

__device__ void block_argmax_first(float myv, long long myi, int64_t* idx_out, float* val_out) {
  __shared__ float sv[32];
  __shared__ long long si[32];
  ValIdx me{myv, myi};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx ot;
    ot.v = __shfl_xor_sync(0xffffffffu, me.v, o);
    ot.i = __shfl_xor_sync(0xffffffffu, me.i, o);
    me = better(me, ot);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sv[warp] = me.v; si[warp] = me.i; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    me.v = lane < nw ? sv[lane] : 0.f;
    me.i = lane < nw ? si[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ValIdx ot;
      ot.v = __shfl_xor_sync(0xffffffffu, me.v, o);
      ot.i = __shfl_xor_sync(0xffffffffu, me.i, o);
      me = better(me, ot);
    }
    if (lane == 0) {
      if (idx_out) *idx_out = me.i;
      if (val_out) *val_out = me.i >= 0 ? me.v : 0.f;
    }
  }
}

// max_std[p] = sqrt(var_scaled) * 2^-e ; then argmax
__global__ void __launch_bounds__(1024) finalize_kernel(const unsigned int* __restrict__ bits, int P, int scale_exp2,
                                                        float* __restrict__ max_std, int64_t* argmax,
                                                        float* maxval) {
  float bv = 0.f;
  long long bi = -1;
  const float sc = ldexpf(1.0f, -scale_exp2);
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float sd = sqrtf(__uint_as_float(bits[p])) * sc;
    max_std[p] = sd;
    if (sd > bv) { bv = sd; bi = p; }   // strided scan keeps the smallest index per thread on ties
  }
  block_argmax_first(bv, bi, argmax, maxval);
}

__global__ void __launch_bounds__(1024) argmax_first_kernel(const float* __restrict__ vals, int P, int64_t* idx,
                                                            float* val) {
  float bv = 0.f;
  long long bi = -1;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float v = vals[p];
    if (v > bv) { bv = v; bi = p; }
  }
  block_argmax_first(bv, bi, idx, val);
}

__global__ void pack_maxloc_kernel(const float* val, const int64_t* idx, int64_t off, int64_t* key) {
  const long long li = *idx;
  if (li < 0) { *key = 0; return; }
  const unsigned long long vb = (unsigned long long)__float_as_uint(*val);      // non-negative float
  const unsigned long long gi = (unsigned long long)(li + off);
  *key = (int64_t)((vb << 32) | (0xFFFFFFFFull - (gi & 0xFFFFFFFFull)));
}

// pilot (sample 0) latent states of every (p, t): Zs [rows][n][Spad] -> Zrows [rows][n]
__global__ void gather_pilot_rows_kernel(const float* __restrict__ Zs, int Spad, long long rows, int n,
                                         float* __restrict__ Zrows) {
  const long long total = rows * n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    Zrows[e] = Zs[e * Spad];
}

// mean[r][d] = decode(pilot)[r][d] + mean_f 2^-e ; std[r][d] = sqrt(max(var_f, 0)) 2^-e
__global__ void finalize_fields_kernel(const float* __restrict__ mean_f, const float* __restrict__ var_f, int field_ld,
                                       long long rows, int D, float sc, float* __restrict__ mean_io,
                                       float* __restrict__ std_out) {
  const long long total = rows * D;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / D;
    const int d = (int)(e - r * D);
    const size_t fo = (size_t)r * field_ld + d;
    if (mean_io) mean_io[e] = fmaf(mean_f[fo], sc, mean_io[e]);
    if (std_out) std_out[e] = sqrtf(fmaxf(var_f[fo], 0.f)) * sc;
  }
}

int launch_gather_pilot_rows(glasdi_handle* h, const float* Zs, int Spad, long long rows, int n, float* Zrows,
                             cudaStream_t st) {
  const long long total = rows * n;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 16);
  gather_pilot_rows_kernel<<<blocks, 256, 0, st>>>(Zs, Spad, rows, n, Zrows);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "gather_pilot_rows launch");
}

int launch_finalize_fields(glasdi_handle* h, const float* mean_f, const float* var_f, int field_ld, long long rows, int D,
                           int scale_exp2, float* mean_io, float* std_out, cudaStream_t st) {
  const long long total = rows * D;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 32);
  finalize_fields_kernel<<<blocks, 256, 0, st>>>(mean_f, var_f, field_ld, rows, D, ldexpf(1.0f, -scale_exp2), mean_io,
                                                 std_out);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "finalize_fields launch");
}

int launch_finalize(glasdi_handle* h, const unsigned int* bits, int P, int scale_exp2, float* max_std,
                    int64_t* argmax, float* maxval, cudaStream_t st) {
  finalize_kernel<<<1, 1024, 0, st>>>(bits, P, scale_exp2, max_std, argmax, maxval);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "finalize launch");
}

int launch_argmax_first(glasdi_handle* h, const float* vals, int P, int64_t* idx, float* val, cudaStream_t st) {
  argmax_first_kernel<<<1, 1024, 0, st>>>(vals, P, idx, val);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "argmax launch");
}

int launch_pack_maxloc(glasdi_handle* h, const float* val, const int64_t* idx, int64_t off, int64_t* key,
                       cudaStream_t st) {
  pack_maxloc_kernel<<<1, 1, 0, st>>>(val, idx, off, key);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "pack_maxloc launch");
}

}  // namespace glasdi
