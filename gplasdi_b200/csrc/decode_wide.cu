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
#include "decode_common.cuh"

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
  int a_blk_kg;           // 16-byte k groups per dof block of A (3 Kp / 8); the first nkg of them are read: [W_hi] of a
                          // one-pass screening launch, [W_hi | W_hi | W_lo] of a three-pass one
  float* field;           // screening launch: [n_groups][field_ld] variance of every dof in true units (else null)
  int field_ld;
  float* mean_field;      // optional, same layout: mean_s (x(s) - x(pilot)) in true units
  float mean_unscale;
  float unscale;          // 2^(-2 (eW + eH)): scaled variance -> true units
};

struct Bars {
  uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
};

// Activations of the SCREENING pass: MUFU ex2 / lg2 / rcp forms (relative error ~2^-21) instead of the libdevice ones the
// fp32-faithful passes and the exact re-evaluation use -- ~6 instead of ~40 instructions per value in epilogues that
// are bound by exactly that.  The same function is applied to the pilot, so sample 0 still cancels to exactly 0.
__device__ __forceinline__ float act_fast(int act, float x) { return act_apply_t<-2>(act, x); }

// 32 accumulator columns of one unit row -> (act(a) - pilot) 2^eH as fp16 into the [hi] part of xvar_kernel's B tile
template <int ACT>
__device__ __forceinline__ void xact_emit_screen(int act_rt, const uint32_t (&r)[32], float unscale, float bias, float pilot_s, float scale,
                                                 int valid_cols, bool unit_ok, bool store_ok, __half* dst, float& amax,
                                                 float& nanacc) {
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const float a = fmaf(__uint_as_float(r[q]), unscale, bias);
    float dvv = fmaf(act_mufu<ACT>(a, act_rt), scale, -pilot_s);
    dvv = (unit_ok && q < valid_cols) ? dvv : 0.f;
    amax = fmaxf(amax, fabsf(dvv));
    nanacc = fmaf(dvv, 0.f, nanacc);                    // NaN / inf leave a NaN behind
    if (store_ok) dst[q * 8] = __float2half_rn(dvv);
  }
}
template <int ACT>
__device__ __forceinline__ float xact_pilot_screen(float a, int act_rt) { return act_mufu<ACT>(a, act_rt); }
#define GLASDI_ACT_DISPATCH(act, CALL)                                            \
  switch (act) {                                                                  \
    case GLASDI_ACT_SIGMOID: { constexpr int A_ = GLASDI_ACT_SIGMOID; CALL; } break;   \
    case GLASDI_ACT_TANH: { constexpr int A_ = GLASDI_ACT_TANH; CALL; } break;         \
    case GLASDI_ACT_SOFTPLUS: { constexpr int A_ = GLASDI_ACT_SOFTPLUS; CALL; } break; \
    case GLASDI_ACT_RELU: { constexpr int A_ = GLASDI_ACT_RELU; CALL; } break;         \
    case GLASDI_ACT_ELU: { constexpr int A_ = GLASDI_ACT_ELU; CALL; } break;           \
    case GLASDI_ACT_SILU: { constexpr int A_ = GLASDI_ACT_SILU; CALL; } break;         \
    case GLASDI_ACT_GELU: { constexpr int A_ = GLASDI_ACT_GELU; CALL; } break;         \
    case GLASDI_ACT_IDENTITY: { constexpr int A_ = GLASDI_ACT_IDENTITY; CALL; } break; \
    default: { constexpr int A_ = -1; CALL; } break;                                   \
  }

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
        const uint8_t* a0 = reinterpret_cast<const uint8_t*>(P.A) + (size_t)dofb * P.a_blk_kg * 128 * 16;
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
      const int dofb = (int)(u - g * P.n_dofb);
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
      if (P.field) P.field[(size_t)(P.g0 + g) * P.field_ld + dofb * 128 + warp * 32 + lane] = var;   // field_ld >= 128 n_dofb
      if (P.mean_field) P.mean_field[(size_t)(P.g0 + g) * P.field_ld + dofb * 128 + warp * 32 + lane] = mean * P.mean_unscale;
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
  int a_blk_kg;           // k groups per unit block of A (3 K1p / 8); the first nkg are read (one or three passes)
  int parts_out;          // 3: [hi | lo | hi] for a three-pass xvar_kernel, 1: [hi] for the screening pass
  float unscale_in;       // 2^-(eW1 + eX)
  float scale_out;        // 2^eH
};

template <bool SCREEN>
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
        const uint8_t* a0 = reinterpret_cast<const uint8_t*>(P.A) + (size_t)ub * P.a_blk_kg * 128 * 16;
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
    float amax = 0.f, nanacc = 0.f;
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
        __half* tile = P.Bt + ((size_t)g * P.n_sb + sb) * (SCREEN ? 1 : 3) * part + ((size_t)(k2 >> 3) * 256) * 8 + (k2 & 7);
        uint32_t ra[32], rb[32];
        if (sb == 0) {                                  // every warp makes its own copy of the pilot (sample 0 = column 0)
          tmem_ld_32x32(taddr, ra);
          tmem_ld_wait();
          const float ap = fmaf(__uint_as_float(ra[0]), P.unscale_in, bias);
          if constexpr (SCREEN) { GLASDI_ACT_DISPATCH(P.act, pilot = xact_pilot_screen<A_>(ap, P.act)) }
          else pilot = act_apply(P.act, ap);
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
            if (store_ok) {
              __half* dst = tile + (size_t)col * 8;
              dst[0] = vh;
              if constexpr (!SCREEN) {
                const __half vl = __float2half_rn(dv - __half2float(vh));
                dst[part] = vl;
                dst[2 * part] = vh;
              }
            }
          }
        };
        const int c_lo = cq * 64;
        tmem_ld_32x32(taddr + c_lo, ra);
        tmem_ld_wait();
        tmem_ld_32x32(taddr + c_lo + 32, rb);
        if constexpr (SCREEN) {
          const float pilot_s = pilot * P.scale_out;
          const int v0 = P.S - (sb * 256 + c_lo);          // valid columns from c_lo on
          GLASDI_ACT_DISPATCH(P.act, xact_emit_screen<A_>(P.act, ra, P.unscale_in, bias, pilot_s, P.scale_out, v0, unit_ok, store_ok,
                                                          tile + (size_t)c_lo * 8, amax, nanacc))
          tmem_ld_wait();
          GLASDI_ACT_DISPATCH(P.act, xact_emit_screen<A_>(P.act, rb, P.unscale_in, bias, pilot_s, P.scale_out, v0 - 32, unit_ok, store_ok,
                                                          tile + (size_t)(c_lo + 32) * 8, amax, nanacc))
        } else {
          emit(ra, c_lo);
          tmem_ld_wait();
          emit(rb, c_lo + 32);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->acc_empty[ab]);
        ++a_it;
      }
    }
    if (over || !(amax <= 65000.f) || nanacc != 0.f) atomicOr(P.flags + 0, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kActEpiWarps + 1) tmem_dealloc(tmem_base, kTmemCols);
}

// H fp32 [n_groups * S][K] (row = group * S + sample) -> B' tiles: centred on the group's pilot (sample 0), scaled by 2^eH,
// split hi + lo, laid out [group][sample block of 256][kg over 3 Kp / 8][256][8] with kg blocks [hi | lo | hi].
// One thread per (group, sample column incl. padding, group of 8 k).
__global__ void __launch_bounds__(256) pack_hidden_kernel(const float* __restrict__ H, int n_groups, int S, int K, int Kp,
                                                          int n_sb, float scale, int centre, int parts,
                                                          __half* __restrict__ Bt, int* __restrict__ flags) {
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
    __half* base = Bt + (((size_t)g * n_sb + sb) * (size_t)(parts * nkgK)) * 256 * 8;
    const size_t o = ((size_t)kg * 256 + col) * 8;
    const size_t part = (size_t)nkgK * 256 * 8;
    *reinterpret_cast<uint4*>(base + o) = *reinterpret_cast<const uint4*>(hi);
    if (parts == 3) {
      *reinterpret_cast<uint4*>(base + part + o) = *reinterpret_cast<const uint4*>(lo);
      *reinterpret_cast<uint4*>(base + 2 * part + o) = *reinterpret_cast<const uint4*>(hi);
    }
  }
  if (over) atomicOr(flags + 0, 1);                     // activation differences beyond the fp16 range of the split
}

// Screening pass: everything in front of xact_kernel in ONE kernel -- thread <-> sample of a (p, t) group: the latent state
// from the trajectory scratch, the narrow front layers in registers (front_pre), the layer that feeds xact_kernel from
// shared-memory weight rows, scaled fp16 written straight into xact_kernel's B tiles [group][sample block][kg][256][8]
// (16 bytes per thread and k group, coalesced over the sample columns).  Replaces gather_rows + one CUDA-core GEMM launch
// per layer (activations through HBM in fp32) + pack_hidden: 23 % of the Vlasov-shaped step.
struct PreParams {
  FrontMlp f;             // layers 0 .. n_front-1; the last one produces xact_kernel's input (width K1)
  const float* Zs;        // [P][T][n][Spad]
  __half* Xt;             // [n_groups][n_sb][K1p / 8][256][8]
  int* flags;
  long long g0;
  int n_groups, n_sb, S, Spad, K1, K1p;
  float scale;            // 2^eX
};

// shared-memory image of the pre-chain: narrow layers l < n_front - 1 as transposed, zero-padded [64 in][64 out] blocks + a
// bias row of 64, the last layer as [K1p / 8][64 in][8 units] + K1p biases.  Inputs and outputs of a layer live in REGISTERS
// (arrays indexed by fully unrolled loops; the first version kept them in local memory: 343 us per slab against 364 us for
// the five launches it replaced), widths are honoured in blocks of 8 inputs / 4 outputs.
constexpr int kPreW = 64;                              // = dv::kMaxHid
__host__ __device__ inline size_t wide_pre_smem_floats(int n_front, int K1p) {
  return (size_t)(n_front - 1) * (kPreW * kPreW + kPreW) + (size_t)K1p * kPreW + (size_t)K1p;
}

__global__ void __launch_bounds__(256) wide_pre_kernel(const __grid_constant__ PreParams P) {
  extern __shared__ __align__(16) float psm[];
  const int nf = P.f.n_front;
  float* sLast = psm + (size_t)(nf - 1) * (kPreW * kPreW + kPreW);       // [K1p / 8][64][8]
  float* sLastB = sLast + (size_t)P.K1p * kPreW;                          // [K1p]
  for (int l = 0; l + 1 < nf; ++l) {
    const int wi = P.f.widths[l], wo = P.f.widths[l + 1];
    float* Wt = psm + (size_t)l * (kPreW * kPreW + kPreW);
    for (int idx = threadIdx.x; idx < kPreW * kPreW; idx += blockDim.x) {
      const int i = idx / kPreW, o = idx - i * kPreW;
      Wt[idx] = (i < wi && o < wo) ? P.f.W[l][(size_t)o * wi + i] : 0.f;
    }
    for (int o = threadIdx.x; o < kPreW; o += blockDim.x) Wt[kPreW * kPreW + o] = o < wo ? P.f.b[l][o] : 0.f;
  }
  {
    const int l = nf - 1, wi = P.f.widths[l];
    for (int idx = threadIdx.x; idx < P.K1p * kPreW; idx += blockDim.x) {
      const int u = idx & 7, i = (idx >> 3) % kPreW, kg = idx / (8 * kPreW);
      const int k = kg * 8 + u;
      sLast[idx] = (k < P.K1 && i < wi) ? P.f.W[l][(size_t)k * wi + i] : 0.f;
    }
    for (int k = threadIdx.x; k < P.K1p; k += blockDim.x) sLastB[k] = k < P.K1 ? P.f.b[l][k] : 0.f;
  }
  __syncthreads();
  const int n = P.f.widths[0];
  const int nkg = P.K1p / 8;
  const int w_last_in = P.f.widths[nf - 1];
  float amax = 0.f, nanacc = 0.f;
  const long long total = (long long)P.n_groups * P.n_sb * 256;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx & 255);
    const long long r = idx >> 8;
    const int sb = (int)(r % P.n_sb);
    const long long g = r / P.n_sb;
    const int s = sb * 256 + col;
    uint4* dst = reinterpret_cast<uint4*>(P.Xt) + (((size_t)g * P.n_sb + sb) * nkg) * 256 + col;
    if (s >= P.S) {                                            // padding samples: exact zeros
      for (int kg = 0; kg < nkg; ++kg) dst[(size_t)kg * 256] = make_uint4(0, 0, 0, 0);
      continue;
    }
    float x[kPreW], y[kPreW];
    const float* __restrict__ zb = P.Zs + ((P.g0 + g) * n) * P.Spad + s;
#pragma unroll
    for (int i = 0; i < kPreW; ++i) x[i] = (i < kMaxLatent && i < n) ? __ldg(zb + (size_t)i * P.Spad) : 0.f;
    for (int l = 0; l + 1 < nf; ++l) {
      const int wi = P.f.widths[l], wo = P.f.widths[l + 1];
      const float* __restrict__ Wt = psm + (size_t)l * (kPreW * kPreW + kPreW);
#pragma unroll
      for (int o4 = 0; o4 < kPreW / 4; ++o4) {
        const float4 b4 = *reinterpret_cast<const float4*>(Wt + kPreW * kPreW + o4 * 4);
        y[o4 * 4 + 0] = b4.x; y[o4 * 4 + 1] = b4.y; y[o4 * 4 + 2] = b4.z; y[o4 * 4 + 3] = b4.w;
      }
#pragma unroll
      for (int i8 = 0; i8 < kPreW / 8; ++i8) {
        if (i8 * 8 < wi) {
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const float xi = x[i8 * 8 + ii];
#pragma unroll
            for (int o4 = 0; o4 < kPreW / 4; ++o4) {
              if (o4 * 4 < wo) {
                const float4 w4 = *reinterpret_cast<const float4*>(Wt + (i8 * 8 + ii) * kPreW + o4 * 4);
                y[o4 * 4 + 0] = fmaf(w4.x, xi, y[o4 * 4 + 0]);
                y[o4 * 4 + 1] = fmaf(w4.y, xi, y[o4 * 4 + 1]);
                y[o4 * 4 + 2] = fmaf(w4.z, xi, y[o4 * 4 + 2]);
                y[o4 * 4 + 3] = fmaf(w4.w, xi, y[o4 * 4 + 3]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int o = 0; o < kPreW; ++o) x[o] = (o < wo) ? act_fast(P.f.act, y[o]) : 0.f;
    }
    for (int kg = 0; kg < nkg; ++kg) {
      float acc[8];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(sLastB + kg * 8), b1 = *reinterpret_cast<const float4*>(sLastB + kg * 8 + 4);
        acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
      }
      const float* __restrict__ Wk = sLast + (size_t)kg * kPreW * 8;
#pragma unroll
      for (int i8 = 0; i8 < kPreW / 8; ++i8) {
        if (i8 * 8 < w_last_in) {
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const float xi = x[i8 * 8 + ii];
            const float4 w0 = *reinterpret_cast<const float4*>(Wk + (i8 * 8 + ii) * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(Wk + (i8 * 8 + ii) * 8 + 4);
            acc[0] = fmaf(w0.x, xi, acc[0]); acc[1] = fmaf(w0.y, xi, acc[1]); acc[2] = fmaf(w0.z, xi, acc[2]); acc[3] = fmaf(w0.w, xi, acc[3]);
            acc[4] = fmaf(w1.x, xi, acc[4]); acc[5] = fmaf(w1.y, xi, acc[5]); acc[6] = fmaf(w1.z, xi, acc[6]); acc[7] = fmaf(w1.w, xi, acc[7]);
          }
        }
      }
      uint32_t out[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v0 = (kg * 8 + 2 * u < P.K1) ? act_fast(P.f.act, acc[2 * u]) * P.scale : 0.f;
        const float v1 = (kg * 8 + 2 * u + 1 < P.K1) ? act_fast(P.f.act, acc[2 * u + 1]) * P.scale : 0.f;
        amax = fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1)));
        nanacc = fmaf(v0 + v1, 0.f, nanacc);
        out[u] = dv::h2_as_u32(__floats2half2_rn(v0, v1));
      }
      dst[(size_t)kg * 256] = make_uint4(out[0], out[1], out[2], out[3]);
    }
  }
  if (!(amax <= 65000.f) || nanacc != 0.f) atomicOr(P.flags + 0, 1);
}


// ---- the same pre-chain on the tensor cores (warp-level mma.sync m16n8k16, fp16 operands, fp32 accumulators) -----------
// wide_pre_kernel evaluates 30 k MACs per sample on the CUDA cores from broadcast shared-memory weights and took 41 % of the
// Vlasov-shaped step (as long as xvar_kernel's GEMM: profiles/r02_wide_launch_shares.txt).  Here a warp owns 16 samples (one
// m16 tile): the accumulator fragment of two neighbouring n8 tiles of a layer IS the A fragment of one k16 block of the next
// layer (rows g / g + 8, columns 2t, 2t + 1 | 2t + 8, 2t + 9), so the chain runs register to register.  Activations AND
// weights are split hi + lo in fp16 and every product is three MMAs (hi hi + lo hi + hi lo, dropped term 2^-22): the chain
// stays fp32-faithful like the CUDA-core version -- tcgen05 is not worth its pipeline for 0.4 % of the step's FLOP, and
// the legacy HMMA path is idle while the CTA-wide GEMM kernels are not running.
// Shared-memory weight image per layer: [hi | lo][out, padded to 8][in, padded to 16, row stride = in_pad + 8 halfs]:
// a B fragment (k = 2t, 2t + 1 | 2t + 8, 2t + 9 of column n = g) is two conflict-free 4-byte loads.
constexpr int kPmThreads = 512;

template <int ACT>      // compile-time activation (bare MUFU forms); < 0: runtime switch
__global__ void __launch_bounds__(kPmThreads, 1) wide_pre_mma_kernel(const __grid_constant__ PreParams P) {
  extern __shared__ __align__(16) uint8_t pmsm[];
  const int nf = P.f.n_front;
  __shared__ dv::PmLayer Ly[kMaxLinear];
  dv::pm_stage(pmsm, P.f, nf, threadIdx.x, blockDim.x, Ly);
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int n = P.f.widths[0];
  const int nkg = P.K1p / 8;
  const int act_rt = P.f.act;
  auto act = [&](float x) { return act_mufu<ACT>(x, act_rt); };
  float amax = 0.f, nanacc = 0.f;
  const long long n_tiles = (long long)P.n_groups * P.n_sb * 16;          // m16 tiles: 16 per 256-sample block
  const long long warp0 = (long long)blockIdx.x * (kPmThreads / 32) + (threadIdx.x >> 5);
  for (long long tile = warp0; tile < n_tiles; tile += (long long)gridDim.x * (kPmThreads / 32)) {
    const int mt = (int)(tile & 15);
    const long long r = tile >> 4;
    const int sb = (int)(r % P.n_sb);
    const long long grp = r / P.n_sb;
    const int col0 = mt * 16;                                               // first sample column of the tile in its block
    const int s_lo = sb * 256 + col0 + g, s_hi = s_lo + 8;
    __half* xt = P.Xt + (((size_t)grp * P.n_sb + sb) * nkg) * 256 * 8;
    // A fragments of the first layer: the latent state (k < n <= 8), rows g and g + 8; padding samples read sample 0
    uint32_t Ah[dv::kPmMaxKb][4], Al[dv::kPmMaxKb][4];
    {
      const float* __restrict__ zb = P.Zs + ((P.g0 + grp) * n) * P.Spad;
      const int sl = s_lo < P.S ? s_lo : 0, sh = s_hi < P.S ? s_hi : 0;
      const int k0 = 2 * t, k1 = 2 * t + 1;
      dv::pm_latent_frag(k0 < n ? __ldg(zb + (size_t)k0 * P.Spad + sl) : 0.f, k1 < n ? __ldg(zb + (size_t)k1 * P.Spad + sl) : 0.f,
                         k0 < n ? __ldg(zb + (size_t)k0 * P.Spad + sh) : 0.f, k1 < n ? __ldg(zb + (size_t)k1 * P.Spad + sh) : 0.f,
                         Ah, Al);
    }
    dv::pm_chain(Ly, nf - 1, g, t, Ah, Al, act);
    // the layer that feeds xact_kernel: (act(.) 2^eX) as fp16 into the B' tiles [kg][256 columns][8]
    const dv::PmLayer& LL = Ly[nf - 1];
    auto emit = [&](int j, bool real, const float (&acc)[4]) {
      uint32_t o_lo = 0u, o_hi = 0u;
      if (real) {
        const int k = 8 * j + 2 * t;
        const float v0 = (k < P.K1 && s_lo < P.S) ? act(acc[0]) * P.scale : 0.f;
        const float v1 = (k + 1 < P.K1 && s_lo < P.S) ? act(acc[1]) * P.scale : 0.f;
        const float v2 = (k < P.K1 && s_hi < P.S) ? act(acc[2]) * P.scale : 0.f;
        const float v3 = (k + 1 < P.K1 && s_hi < P.S) ? act(acc[3]) * P.scale : 0.f;
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v0), fabsf(v1)), fmaxf(fabsf(v2), fabsf(v3))));
        nanacc = fmaf((v0 + v1) + (v2 + v3), 0.f, nanacc);
        o_lo = dv::h2_as_u32(__floats2half2_rn(v0, v1));
        o_hi = dv::h2_as_u32(__floats2half2_rn(v2, v3));
      }
      uint32_t* dst = reinterpret_cast<uint32_t*>(xt + ((size_t)j * 256 + col0 + g) * 8 + 2 * t);
      dst[0] = o_lo;
      dst[8 * 8 / 2] = o_hi;                               // column + 8: 8 columns x 16 bytes further
    };
    const int ntl = LL.op / 8;
    for (int j = 0; j < nkg; j += 2) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
      if (j + 1 < ntl) dv::pm_tiles<dv::kPmMaxKb, true, true>(LL, j, j + 1, g, t, Ah, Al, acc, acc1);
      else if (j < ntl) dv::pm_tiles<dv::kPmMaxKb, true, false>(LL, j, j, g, t, Ah, Al, acc, acc1);
      emit(j, j < ntl, acc);
      if (j + 1 < nkg) emit(j + 1, j + 1 < ntl, acc1);
    }
  }
  if (!(amax <= 65000.f) || nanacc != 0.f) atomicOr(P.flags + 0, 1);
}

}  // namespace wd

int wide_kp(int K) { return (K + 63) / 64 * 64; }

size_t wide_b_bytes(int K, int S, long long n_groups, int parts) {
  const size_t n_sb = (size_t)(S + 255) / 256;
  return (size_t)n_groups * n_sb * (size_t)(parts * wide_kp(K) / 8) * 256 * 16;
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

size_t wide_x_bytes(int K1, int S, long long n_groups, int parts) { return wide_b_bytes(K1, S, n_groups, parts); }

// X fp32 [n_groups * S][K1] (input of the last front layer, from the CUDA-core chain) -> the packed, pilot-centred hidden
// layer tiles that xvar_kernel consumes (bt_ws); xt_ws = scratch for the packed input
int launch_wide_front(glasdi_handle* h, const float* X, long long n_groups, int S, void* xt_ws, void* bt_ws, int parts,
                      cudaStream_t st) {
  const Decoder& d = h->dec;
  const int K1 = d.wide_k1, K1p = wide_kp(K1), Kp = wide_kp(d.K);
  const int n_sb = (S + 255) / 256;
  const int eX = d.eH;                                   // same range argument as the hidden layer (unbounded act: 2^4)
  const long long total = n_groups * n_sb * 256 * (K1p / 8);
  const unsigned int blocks = (unsigned int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 32);
  int rc = GLASDI_OK;
  if (X) {                                               // X == nullptr: xt_ws was written by launch_wide_pre
    wd::pack_hidden_kernel<<<blocks, 256, 0, st>>>(X, (int)n_groups, S, K1, K1p, n_sb, std::ldexp(1.0f, eX), 0, parts,
                                                   reinterpret_cast<__half*>(xt_ws), h->d_flags);
    h->launches++;
    rc = check_cuda(h, cudaGetLastError(), "pack_hidden_kernel (input) launch");
    if (rc) return rc;
  }
  wd::ActParams A{};
  A.A = d.wide_front;
  A.Xt = reinterpret_cast<const __half*>(xt_ws);
  A.bias = d.db[d.n_linear - 2];
  A.Bt = reinterpret_cast<__half*>(bt_ws);
  A.flags = h->d_flags;
  A.n_groups = (int)n_groups;
  A.n_ub = (Kp + 127) / 128;
  A.n_sb = n_sb;
  A.nkg = parts * K1p / 8;
  A.a_blk_kg = 3 * K1p / 8;
  A.parts_out = parts;
  A.S = S;
  A.K = d.K;
  A.Kp = Kp;
  A.act = d.act;
  A.unscale_in = std::ldexp(1.0f, -(d.eW1 + eX));
  A.scale_out = std::ldexp(1.0f, d.eH);
  const size_t smem = wd::kStages * wd::kStageBytes + sizeof(wd::Bars) + 16;
  long long ctas = h->num_sms;
  if (h->opt_max_ctas > 0) ctas = std::max<long long>(1, std::min<long long>(ctas, h->opt_max_ctas));
  ctas = std::min<long long>(ctas, n_groups * A.n_ub);
  if (parts == 1) {
    GL_CUDA(h, cudaFuncSetAttribute(wd::xact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wd::xact_kernel<true><<<(unsigned)ctas, wd::kActThreads, smem, st>>>(A);
  } else {
    GL_CUDA(h, cudaFuncSetAttribute(wd::xact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wd::xact_kernel<false><<<(unsigned)ctas, wd::kActThreads, smem, st>>>(A);
  }
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "xact_kernel launch");
}

// can the layers in front of xact_kernel run in wide_pre_kernel?  (narrow front: registers of one thread per sample)
bool wide_pre_supported(const Decoder& d) {
  if (d.wide_front == nullptr || d.n_linear < 3 || d.widths[0] > kMaxLatent) return false;
  for (int l = 1; l + 2 < d.n_linear; ++l)
    if (d.widths[l] > dv::kMaxHid) return false;
  const int nf = d.n_linear - 2;                          // layers of the pre-chain
  return wd::wide_pre_smem_floats(nf, wide_kp(d.widths[nf])) * sizeof(float) <= 200 * 1024;
}

int launch_wide_pre(glasdi_handle* h, const float* Zs, int Spad, long long g0, long long n_groups, int S, void* xt_ws,
                    cudaStream_t st) {
  const Decoder& d = h->dec;
  wd::PreParams P{};
  const int nf = d.n_linear - 2;
  P.f.n_front = nf;
  P.f.act = d.act;
  for (int l = 0; l <= nf; ++l) P.f.widths[l] = d.widths[l];
  for (int l = 0; l < nf; ++l) { P.f.W[l] = d.dW[l]; P.f.b[l] = d.db[l]; }
  P.Zs = Zs;
  P.Xt = reinterpret_cast<__half*>(xt_ws);
  P.flags = h->d_flags;
  P.g0 = g0;
  P.n_groups = (int)n_groups;
  P.n_sb = (S + 255) / 256;
  P.S = S;
  P.Spad = Spad;
  P.K1 = d.wide_k1;
  P.K1p = wide_kp(d.wide_k1);
  P.scale = std::ldexp(1.0f, d.eH);
  if (h->opt_wide != 3) {                                 // default: the pre-chain on mma.sync (3 = the CUDA-core version)
    const size_t smem_mma = dv::pm_smem_bytes(P.f.widths, nf);
    if (smem_mma <= 200 * 1024) {
      const long long n_tiles = n_groups * P.n_sb * 16;
      const unsigned int blocks =
          (unsigned int)std::min<long long>((n_tiles + wd::kPmThreads / 32 - 1) / (wd::kPmThreads / 32), (long long)h->num_sms);
      auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
        if (e == cudaSuccess) kern<<<blocks, wd::kPmThreads, smem_mma, st>>>(P);
        return e;
      };
      cudaError_t e;
      switch (d.act) {
        case GLASDI_ACT_SIGMOID: e = go(wd::wide_pre_mma_kernel<GLASDI_ACT_SIGMOID>); break;
        case GLASDI_ACT_TANH: e = go(wd::wide_pre_mma_kernel<GLASDI_ACT_TANH>); break;
        case GLASDI_ACT_SOFTPLUS: e = go(wd::wide_pre_mma_kernel<GLASDI_ACT_SOFTPLUS>); break;
        case GLASDI_ACT_SILU: e = go(wd::wide_pre_mma_kernel<GLASDI_ACT_SILU>); break;
        case GLASDI_ACT_RELU: e = go(wd::wide_pre_mma_kernel<GLASDI_ACT_RELU>); break;
        default: e = go(wd::wide_pre_mma_kernel<-1>); break;
      }
      GL_CUDA(h, e);
      h->launches++;
      return check_cuda(h, cudaGetLastError(), "wide_pre_mma_kernel launch");
    }
  }
  const size_t smem = wd::wide_pre_smem_floats(nf, P.K1p) * sizeof(float);
  GL_CUDA(h, cudaFuncSetAttribute(wd::wide_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long total = n_groups * P.n_sb * 256;
  const unsigned int blocks = (unsigned int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 8);
  wd::wide_pre_kernel<<<blocks, 256, smem, st>>>(P);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "wide_pre_kernel launch");
}

// Hlast fp32 [n_groups * S][K] (the exact CUDA-core chain's output for a slab of (p, t) groups) -> scaled variance maxima
// (Hlast = nullptr: bt_ws already holds the packed hidden layer, written by launch_wide_front)
int launch_wide_last_var(glasdi_handle* h, const float* Hlast, long long g0, long long n_groups, int S, int T, void* bt_ws,
                         unsigned int* maxvar_bits, int parts, float* field, int field_ld, cudaStream_t st,
                         float* mean_field) {
  const Decoder& d = h->dec;
  const int Kp = wide_kp(d.K);
  const int n_sb = (S + 255) / 256;
  if (Hlast) {
    const long long total = n_groups * n_sb * 256 * (Kp / 8);
    const unsigned int blocks = (unsigned int)std::min<long long>((total + 255) / 256, (long long)h->num_sms * 32);
    wd::pack_hidden_kernel<<<blocks, 256, 0, st>>>(Hlast, (int)n_groups, S, d.K, Kp, n_sb, std::ldexp(1.0f, d.eH), 1, parts,
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
  W.nkg = parts * Kp / 8;
  W.a_blk_kg = 3 * Kp / 8;
  W.field = field;
  W.field_ld = field_ld;
  W.mean_field = mean_field;
  W.mean_unscale = std::ldexp(1.0f, -(d.eW + d.eH));
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
