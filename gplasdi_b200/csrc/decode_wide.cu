// decode_wide.cu -- last decoder layer + variance on tcgen05 for WIDE last layers (K_last > 112: the Vlasov- and
// rising-bubble-shaped decoders with K_last = 1000), decoded-ensemble form.
//
// Reference semantics: std(0) over the samples of decoder(Z), max over (t, dof) (src/lasdi/gplasdi.py:52-74 with the
// decoder of src/lasdi/latent_space.py:157-164).  For K_last >= 256 the tcgen05 accumulator drain is no longer the
// bound of this form (K/256 >= 1), so the last layer is a plain K-looped GEMM per (parameter, time step):
//
//     X[dof, s] - X[dof, pilot] = sum_k W[dof, k] (h[s, k] - h[pilot, k])
//
// with a variance epilogue over the sample columns.  fp32-faithful products on the fp16 tensor cores as in
// decode_var2.cu: both operands are split x 2^e = hi + lo and W_hi H_hi + W_hi H_lo + W_lo H_hi is accumulated -- here
// by CONCATENATING the three passes along the contraction (A' = [W_hi | W_hi | W_lo], B' = [H_hi | H_lo | H_hi],
// K' = 3 K), so the kernel is an ordinary fp16 GEMM and shares its pipeline with quad_kernel (decode_gram.cu): one
// loader thread streaming pre-packed canonical K-major tiles with 1-D bulk copies through a 4-stage ring, one MMA warp,
// 128 x 256 tiles (M = dofs on the TMEM lanes, N = samples on the columns), double-buffered 256-column accumulators, four
// epilogue warps.  The hidden activations come from the exact fp32 CUDA-core chain (simt_fallback.cu) and are centred,
// scaled, split and packed by pack_hidden_kernel; padding samples are exact zeros.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace glasdi {
namespace wd {

constexpr int kStages = 4;
constexpr int kChunkGroups = 8;                         // 8 groups of 8 = 64 k' per stage
constexpr uint32_t kABytes = kChunkGroups * 128 * 16;   // 16 KB: 128 dofs x 64 k'
constexpr uint32_t kBBytes = kChunkGroups * 256 * 16;   // 32 KB: 256 samples x 64 k'
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;                           // 4 epilogue warps, loader, MMA
constexpr int kTmemCols = 512;

struct WideParams {
  const __half* A;        // [n_dofb][nkg][128][8]            W' = [W_hi | W_hi | W_lo] 2^eW
  const __half* Bt;       // [n_groups][n_sb][nkg][256][8]    H' = [H_hi | H_lo | H_hi], H = (h - h_pilot) 2^eH
  unsigned int* maxvar_bits;   // [P] float bits of the scaled variance maximum
  long long g0;           // first (p, t) group of this slab
  int n_groups, n_dofb, n_sb, nkg, T, S;
  float unscale;          // 2^(-2 (eW + eH)): scaled variance -> true units
};

struct Bars {
  uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
};

__global__ void __launch_bounds__(kThreads, 1) xvar_kernel(const __grid_constant__ WideParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  Bars* B = reinterpret_cast<Bars*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * kStageBytes + sizeof(Bars));
  const int nchunks = P.nkg / kChunkGroups;
  const long long n_units = (long long)P.n_groups * P.n_dofb;     // unit = (group, dof block); tiles = its sample blocks

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&B->full[i], 1);
      mbar_init(&B->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&B->acc_full[i], 1);
      mbar_init(&B->acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ---- loader ---------------------------------------------------------------------------------
    if (lane == 0) {
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        const long long g = u / P.n_dofb;
        const int dofb = (int)(u - g * P.n_dofb);
        const uint8_t* a0 = reinterpret_cast<const uint8_t*>(P.A) + (size_t)dofb * P.nkg * 128 * 16;
        for (int sb = 0; sb < P.n_sb; ++sb) {
          const uint8_t* b0 = reinterpret_cast<const uint8_t*>(P.Bt) + ((size_t)g * P.n_sb + sb) * P.nkg * 256 * 16;
          for (int kc = 0; kc < nchunks; ++kc) {
            const uint32_t s = it % kStages;
            mbar_wait_relaxed(&B->empty[s], ((it / kStages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&B->full[s], kStageBytes);
            uint8_t* dst = smem + s * kStageBytes;
            bulk_g2s(dst, a0 + (size_t)kc * kABytes, kABytes, &B->full[s]);
            bulk_g2s(dst + kABytes, b0 + (size_t)kc * kBBytes, kBBytes, &B->full[s]);
            ++it;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---- MMA issuer -------------------------------------------------------------------------------
    uint32_t it = 0, a_it = 0;
    const uint32_t idesc = umma_idesc_f16_m128(256u);
    const uint64_t adesc0 = umma_desc_kmajor(smem_u32(smem), 128u * 16u, 128u);
    const uint64_t bdesc0 = umma_desc_kmajor(smem_u32(smem + kABytes), 256u * 16u, 128u);
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      for (int sb = 0; sb < P.n_sb; ++sb) {
        const uint32_t ab = a_it & 1u;
        mbar_wait(&B->acc_empty[ab], ((a_it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * 256u;
        for (int kc = 0; kc < nchunks; ++kc) {
          const uint32_t s = it % kStages;
          mbar_wait(&B->full[s], (it / kStages) & 1u);
          tc_fence_after();
          const uint64_t a0 = adesc0 + (uint64_t)(s * (kStageBytes >> 4));
          const uint64_t b0 = bdesc0 + (uint64_t)(s * (kStageBytes >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kChunkGroups / 2; ++kk)
              tc_mma_f16(d_tmem, a0 + (uint64_t)(kk * ((2 * 128 * 16) >> 4)), b0 + (uint64_t)(kk * ((2 * 256 * 16) >> 4)),
                         idesc, (kc | kk) != 0 ? 1u : 0u);
            tc_commit(&B->empty[s]);
          }
          __syncwarp();
          ++it;
        }
        if (elect_one()) tc_commit(&B->acc_full[ab]);
        __syncwarp();
        ++a_it;
      }
    }
  } else {
    // ---- epilogue: thread <-> dof row; sum d, sum d^2 over the sample columns of all sample blocks of the unit -----
    uint32_t a_it = 0;
    const float inv_S = 1.0f / (float)P.S;
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      const long long g = u / P.n_dofb;
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};     // four independent chains
      for (int sb = 0; sb < P.n_sb; ++sb) {
        const uint32_t ab = a_it & 1u;
        mbar_wait_relaxed(&B->acc_full[ab], (a_it >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (((uint32_t)(warp * 32)) << 16) + ab * 256u;
        uint32_t ra[32], rb[32];
        auto fold = [&](const uint32_t (&r)[32]) {
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float d = __uint_as_float(r[q + c]);
              s1[c] += d;
              s2[c] = fmaf(d, d, s2[c]);
            }
          }
        };
        tmem_ld_32x32(taddr, ra);
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 64) {
          tmem_ld_wait();
          tmem_ld_32x32(taddr + c0 + 32, rb);
          fold(ra);
          tmem_ld_wait();
          if (c0 + 64 < 256) tmem_ld_32x32(taddr + c0 + 64, ra);
          fold(rb);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->acc_empty[ab]);
        ++a_it;
      }
      const float sum = (s1[0] + s1[1]) + (s1[2] + s1[3]);
      const float sq = (s2[0] + s2[1]) + (s2[2] + s2[3]);
      const float mean = sum * inv_S;
      const float var = fmaxf(0.f, fmaf(-mean, mean, sq * inv_S)) * P.unscale;  // rows >= D have zero weights: var = 0
      const float wm = warp_max(var);
      if (lane == 0 && wm > 0.f) atomicMax(P.maxvar_bits + (P.g0 + g) / P.T, __float_as_uint(wm));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

// xact_kernel: the LAST FRONT layer of a wide decoder on the same pipeline.  A' = [W1_hi | W1_hi | W1_lo] (128 output units
// per tile on the TMEM lanes), B' = [X_hi | X_lo | X_hi] (the layer's fp32 input from the CUDA-core chain, split, NOT
// centred; 256 samples per tile).  Epilogue: thread = output unit; a = acc 2^-(e) + bias, h = act(a) with the accurate
// activations of the reference path, pilot = h of sample 0, (h - pilot) 2^eH split hi + lo and written straight into the
// packed operand tiles of xvar_kernel ([hi | lo | hi] parts; padding samples and padding units are exact zeros) -- the
// hidden layer never exists in fp32.  The epilogue is ~100 instructions per value (libdevice activation, split, three
// stores) against 6 k clocks of MMA per tile: SIXTEEN epilogue warps (four per TMEM lane quadrant, 64 columns each;
// four warps were latency-bound and slower than the CUDA-core layer they replace).
constexpr int kActEpiWarps = 16;
constexpr int kActThreads = (kActEpiWarps + 2) * 32;     // + loader + MMA

struct ActParams {
  const __half* A;        // [n_ub][nkg][128][8]              W1' tiles (nkg = 3 K1p / 8)
  const __half* Xt;       // [n_groups][n_sb][nkg][256][8]    input tiles
  const float* bias;      // [K] bias of the layer
  __half* Bt;             // [n_groups][n_sb][3 Kp / 8][256][8]   output = xvar_kernel's B'
  int* flags;
  int n_groups, n_ub, n_sb, nkg, S, K, Kp, act;
  float unscale_in;       // 2^-(eW1 + eX)
  float scale_out;        // 2^eH
};

__global__ void __launch_bounds__(kActThreads, 1) xact_kernel(const __grid_constant__ ActParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  Bars* B = reinterpret_cast<Bars*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * kStageBytes + sizeof(Bars));
  const int nchunks = P.nkg / kChunkGroups;
  const long long n_units = (long long)P.n_groups * P.n_ub;       // unit = (group, block of 128 output units)

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&B->full[i], 1);
      mbar_init(&B->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&B->acc_full[i], 1);
      mbar_init(&B->acc_empty[i], kActEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kActEpiWarps + 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kActEpiWarps) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        const long long g = u / P.n_ub;
        const int ub = (int)(u - g * P.n_ub);
        const uint8_t* a0 = reinterpret_cast<const uint8_t*>(P.A) + (size_t)ub * P.nkg * 128 * 16;
        for (int sb = 0; sb < P.n_sb; ++sb) {
          const uint8_t* b0 = reinterpret_cast<const uint8_t*>(P.Xt) + ((size_t)g * P.n_sb + sb) * P.nkg * 256 * 16;
          for (int kc = 0; kc < nchunks; ++kc) {
            const uint32_t s = it % kStages;
            mbar_wait_relaxed(&B->empty[s], ((it / kStages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&B->full[s], kStageBytes);
            uint8_t* dst = smem + s * kStageBytes;
            bulk_g2s(dst, a0 + (size_t)kc * kABytes, kABytes, &B->full[s]);
            bulk_g2s(dst + kABytes, b0 + (size_t)kc * kBBytes, kBBytes, &B->full[s]);
            ++it;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kActEpiWarps + 1) {
    uint32_t it = 0, a_it = 0;
    const uint32_t idesc = umma_idesc_f16_m128(256u);
    const uint64_t adesc0 = umma_desc_kmajor(smem_u32(smem), 128u * 16u, 128u);
    const uint64_t bdesc0 = umma_desc_kmajor(smem_u32(smem + kABytes), 256u * 16u, 128u);
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      for (int sb = 0; sb < P.n_sb; ++sb) {
        const uint32_t ab = a_it & 1u;
        mbar_wait(&B->acc_empty[ab], ((a_it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * 256u;
        for (int kc = 0; kc < nchunks; ++kc) {
          const uint32_t s = it % kStages;
          mbar_wait(&B->full[s], (it / kStages) & 1u);
          tc_fence_after();
          const uint64_t a0 = adesc0 + (uint64_t)(s * (kStageBytes >> 4));
          const uint64_t b0 = bdesc0 + (uint64_t)(s * (kStageBytes >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kChunkGroups / 2; ++kk)
              tc_mma_f16(d_tmem, a0 + (uint64_t)(kk * ((2 * 128 * 16) >> 4)), b0 + (uint64_t)(kk * ((2 * 256 * 16) >> 4)),
                         idesc, (kc | kk) != 0 ? 1u : 0u);
            tc_commit(&B->empty[s]);
          }
          __syncwarp();
          ++it;
        }
        if (elect_one()) tc_commit(&B->acc_full[ab]);
        __syncwarp();
        ++a_it;
      }
    }
  } else {
    // ---- epilogue: thread <-> output unit (TMEM lane quadrant = warp % 4), columns [64 cq, 64 cq + 64) ---------------
    uint32_t a_it = 0;
    const int quad = warp & 3, cq = warp >> 2;
    const int nkgK = P.Kp / 8;
    const size_t part = (size_t)nkgK * 256 * 8;                      // halfs per [hi] / [lo] / [hi] part of one tile
    bool over = false;
    for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
      const long long g = u / P.n_ub;
      const int ub = (int)(u - g * P.n_ub);
      const int k2 = ub * 128 + quad * 32 + lane;
      const bool unit_ok = k2 < P.K, store_ok = k2 < P.Kp;
      const float bias = unit_ok ? __ldg(P.bias + k2) : 0.f;
      float pilot = 0.f;
      for (int sb = 0; sb < P.n_sb; ++sb) {
        const uint32_t ab = a_it & 1u;
        mbar_wait_relaxed(&B->acc_full[ab], (a_it >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (((uint32_t)(quad * 32)) << 16) + ab * 256u;
        __half* tile = P.Bt + ((size_t)g * P.n_sb + sb) * 3 * part + ((size_t)(k2 >> 3) * 256) * 8 + (k2 & 7);
        uint32_t ra[32], rb[32];
        if (sb == 0) {                                  // every warp makes its own copy of the pilot (sample 0 = column 0)
          tmem_ld_32x32(taddr, ra);
          tmem_ld_wait();
          pilot = act_apply(P.act, fmaf(__uint_as_float(ra[0]), P.unscale_in, bias));
        }
        auto emit = [&](const uint32_t (&r)[32], int c0) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int col = c0 + q;
            const int s = sb * 256 + col;
            const float a = fmaf(__uint_as_float(r[q]), P.unscale_in, bias);
            const float hval = act_apply(P.act, a);
            const float dv = (unit_ok && s < P.S) ? (hval - pilot) * P.scale_out : 0.f;
            over |= !(fabsf(dv) <= 65000.f);
            const __half vh = __float2half_rn(dv);
            const __half vl = __float2half_rn(dv - __half2float(vh));
            if (store_ok) {
              __half* dst = tile + (size_t)col * 8;
              dst[0] = vh;
              dst[part] = vl;
              dst[2 * part] = vh;
            }
          }
        };
        const int c_lo = cq * 64;
        tmem_ld_32x32(taddr + c_lo, ra);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + c_lo + 32, rb);
        emit(ra, c_lo);
        tmem_ld_wait();
        emit(rb, c_lo + 32);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->acc_empty[ab]);
        ++a_it;
      }
    }
    if (over) atomicOr(P.flags + 0, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kActEpiWarps + 1) tmem_dealloc(tmem_base, kTmemCols);
}

// H fp32 [n_groups * S][K] (row = group * S + sample) -> B' tiles: centred on the group's pilot (sample 0), scaled by 2^eH,
// split hi + lo, laid out [group][sample block of 256][kg over 3 Kp / 8][256][8] with kg blocks [hi | lo | hi].
// One thread per (group, sample column incl. padding, group of 8 k).
__global__ void __launch_bounds__(256) pack_hidden_kernel(const float* __restrict__ H, int n_groups, int S, int K, int Kp,
                                                          int n_sb, float scale, int centre, __half* __restrict__ Bt,
                                                          int* __restrict__ flags) {
  const int nkgK = Kp / 8;
  const long long total = (long long)n_groups * n_sb * 256 * nkgK;
  bool over = false;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % 256);
    long long r = idx / 256;
    const int kg = (int)(r % nkgK);
    r /= nkgK;
    const int sb = (int)(r % n_sb);
    const long long g = r / n_sb;
    const int s = sb * 256 + col;
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = kg * 8 + u;
      float v = 0.f;
      if (s < S && k < K) {
        const float* __restrict__ row0 = H + (size_t)g * S * K;
        v = (__ldg(row0 + (size_t)s * K + k) - (centre ? __ldg(row0 + k) : 0.f)) * scale;
      }
      over |= !(fabsf(v) <= 65000.f);
      hi[u] = __float2half_rn(v);
      lo[u] = __float2half_rn(v - __half2float(hi[u]));
    }
    __half* base = Bt + (((size_t)g * n_sb + sb) * (size_t)(3 * nkgK)) * 256 * 8;
    const size_t o = ((size_t)kg * 256 + col) * 8;
    const size_t part = (size_t)nkgK * 256 * 8;
    *reinterpret_cast<uint4*>(base + o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + part + o) = *reinterpret_cast<const uint4*>(lo);
    *reinterpret_cast<uint4*>(base + 2 * part + o) = *reinterpret_cast<const uint4*>(hi);
  }
  if (over) atomicOr(flags + 0, 1);                     // activation differences beyond the fp16 range of the split
}

}  // namespace wd

int wide_kp(int K) { return (K + 63) / 64 * 64; }

size_t wide_b_bytes(int K, int S, long long n_groups) {
  const size_t n_sb = (size_t)(S + 255) / 256;
  return (size_t)n_groups * n_sb * (size_t)(3 * wide_kp(K) / 8) * 256 * 16;
}

// [W_hi | W_hi | W_lo] 2^e of a layer W [out][in] as fp16 canonical K-major tiles [row block of 128][3 inp / 8][128][8]
static int pack_split_weights(glasdi_handle* h, const float* W, int out, int in, int e, __half** dev) {
  const int inp = wide_kp(in), nkg = inp / 8;
  const int n_rb = (out + 127) / 128;
  std::vector<__half> packed((size_t)n_rb * 3 * nkg * 128 * 8, __float2half_rn(0.f));
  for (int r = 0; r < out; ++r) {
    const int blk = r / 128, row = r % 128;
    __half* base = packed.data() + (size_t)blk * 3 * nkg * 128 * 8;
    const size_t part = (size_t)nkg * 128 * 8;
    for (int k = 0; k < in; ++k) {
      const float v = std::ldexp(W[(size_t)r * in + k], e);
      const __half vh = __float2half_rn(v);
      const __half vl = __float2half_rn(v - __half2float(vh));
      const size_t o = ((size_t)(k >> 3) * 128 + row) * 8 + (k & 7);
      base[o] = vh;
      base[part + o] = vh;
      base[2 * part + o] = vl;
    }
  }
  if (*dev) cudaFree(*dev);
  *dev = nullptr;
  GL_CUDA(h, cudaMalloc(dev, packed.size() * sizeof(__half)));
  GL_CUDA(h, cudaMemcpy(*dev, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
  return GLASDI_OK;
}

// host, once per glasdi_set_decoder: the last layer, and the last FRONT layer when it is wide enough to pay
int wide_pack_w(glasdi_handle* h, const float* WL /*[D][K] host*/, const float* W1 /*[K][K1] host or null*/, int K1) {
  Decoder& d = h->dec;
  int rc = pack_split_weights(h, WL, d.D, d.K, d.eW, &d.wide_packed);
  if (rc) return rc;
  if (d.wide_front) cudaFree(d.wide_front);
  d.wide_front = nullptr;
  d.wide_k1 = 0;
  if (W1 && K1 >= 16) {
    float wmax = 0.f;
    for (size_t i = 0; i < (size_t)d.K * K1; ++i) wmax = std::max(wmax, std::fabs(W1[i]));
    d.eW1 = (wmax > 0.f) ? 13 - std::ilogb(wmax) : 0;
    d.eW1 = std::max(-100, std::min(100, d.eW1));
    rc = pack_split_weights(h, W1, d.K, K1, d.eW1, &d.wide_front);
    if (rc) return rc;
    d.wide_k1 = K1;
  }
  return GLASDI_OK;
}

size_t wide_x_bytes(int K1, int S, long long n_groups) { return wide_b_bytes(K1, S, n_groups); }

// X fp32 [n_groups * S][K1] (input of the last front layer, from the CUDA-core chain) -> the packed, pilot-centred hidden
// layer tiles that xvar_kernel consumes (bt_ws); xt_ws = scratch for the packed input
int launch_wide_front(glasdi_handle* h, const float* X, long long n_groups, int S, void* xt_ws, void* bt_ws,
                      cudaStream_t st) {
  const Decoder& d = h->dec;
  const int K1 = d.wide_k1, K1p = wide_kp(K1), Kp = wide_kp(d.K);
  const int n_sb = (S + 255) / 256;
  const int eX = d.eH;                                   // same range argument as the hidden layer (unbounded act: 2^4)
  const long long total = n_groups * n_sb * 256 * (K1p / 8);
  const unsigned int blocks = (unsigned int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 32);
  wd::pack_hidden_kernel<<<blocks, 256, 0, st>>>(X, (int)n_groups, S, K1, K1p, n_sb, std::ldexp(1.0f, eX), 0,
                                                 reinterpret_cast<__half*>(xt_ws), h->d_flags);
  h->launches++;
  int rc = check_cuda(h, cudaGetLastError(), "pack_hidden_kernel (input) launch");
  if (rc) return rc;
  wd::ActParams A{};
  A.A = d.wide_front;
  A.Xt = reinterpret_cast<const __half*>(xt_ws);
  A.bias = d.db[d.n_linear - 2];
  A.Bt = reinterpret_cast<__half*>(bt_ws);
  A.flags = h->d_flags;
  A.n_groups = (int)n_groups;
  A.n_ub = (Kp + 127) / 128;
  A.n_sb = n_sb;
  A.nkg = 3 * K1p / 8;
  A.S = S;
  A.K = d.K;
  A.Kp = Kp;
  A.act = d.act;
  A.unscale_in = std::ldexp(1.0f, -(d.eW1 + eX));
  A.scale_out = std::ldexp(1.0f, d.eH);
  const size_t smem = wd::kStages * wd::kStageBytes + sizeof(wd::Bars) + 16;
  GL_CUDA(h, cudaFuncSetAttribute(wd::xact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long ctas = h->num_sms;
  if (h->opt_max_ctas > 0) ctas = std::max<long long>(1, std::min<long long>(ctas, h->opt_max_ctas));
  ctas = std::min<long long>(ctas, n_groups * A.n_ub);
  wd::xact_kernel<<<(unsigned)ctas, wd::kActThreads, smem, st>>>(A);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "xact_kernel launch");
}

// Hlast fp32 [n_groups * S][K] (the exact CUDA-core chain's output for a slab of (p, t) groups) -> scaled variance maxima
// (Hlast = nullptr: bt_ws already holds the packed hidden layer, written by launch_wide_front)
int launch_wide_last_var(glasdi_handle* h, const float* Hlast, long long g0, long long n_groups, int S, int T, void* bt_ws,
                         unsigned int* maxvar_bits, cudaStream_t st) {
  const Decoder& d = h->dec;
  const int Kp = wide_kp(d.K);
  const int n_sb = (S + 255) / 256;
  if (Hlast) {
    const long long total = n_groups * n_sb * 256 * (Kp / 8);
    const unsigned int blocks = (unsigned int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 32);
    wd::pack_hidden_kernel<<<blocks, 256, 0, st>>>(Hlast, (int)n_groups, S, d.K, Kp, n_sb, std::ldexp(1.0f, d.eH), 1,
                                                   reinterpret_cast<__half*>(bt_ws), h->d_flags);
    h->launches++;
    int rcp = check_cuda(h, cudaGetLastError(), "pack_hidden_kernel launch");
    if (rcp) return rcp;
  }
  wd::WideParams W{};
  W.A = d.wide_packed;
  W.Bt = reinterpret_cast<const __half*>(bt_ws);
  W.maxvar_bits = maxvar_bits;
  W.g0 = g0;
  W.n_groups = (int)n_groups;
  W.n_dofb = (d.D + 127) / 128;
  W.n_sb = n_sb;
  W.nkg = 3 * Kp / 8;
  W.T = T;
  W.S = S;
  W.unscale = std::ldexp(1.0f, -2 * (d.eW + d.eH));
  const size_t smem = wd::kStages * wd::kStageBytes + sizeof(wd::Bars) + 16;
  GL_CUDA(h, cudaFuncSetAttribute(wd::xvar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long ctas = h->num_sms;
  if (h->opt_max_ctas > 0) ctas = std::max<long long>(1, std::min<long long>(ctas, h->opt_max_ctas));
  ctas = std::min<long long>(ctas, n_groups * W.n_dofb);
  wd::xvar_kernel<<<(unsigned)ctas, wd::kThreads, smem, st>>>(W);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "xvar_kernel launch");
}

}  // namespace glasdi
