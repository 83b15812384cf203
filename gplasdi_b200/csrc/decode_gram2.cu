// decode_gram2.cu -- gram2_kernel: the covariance GEMM of the screening pass (see decode_gram.cu) for decoders with ONE
// front layer and a smooth activation (sigmoid / tanh / softplus), restructured around two observations of the round-1
// kernel's profile (issue slots 55 %, MUFU 50 %, shared-memory data pipe 62 %: three CUDA-core limits at once, tensor 24 %):
//
// 1. What the Gram operand needs is the DIFFERENCE h(a_pilot + da) - h(a_pilot), and da = W0 (z - z_pilot) is small for a
//    GP ensemble (|da| < 0.4 at the bench shape).  A Taylor polynomial in da with per-(item, unit) coefficients
//    f^(k)(a_pilot) / k! is centred by construction, needs no MUFU at all, and at order 5 is closer to the exact difference
//    (<= 2e-3 of it at |da| = 1, < 1e-6 at the bench spread) than the fp16 rounding of the operand that follows.  It is still a
//    SCREENING value: refine.cu re-evaluates the near-maximum (t, dof) exactly and checks the deviation.  Sample blocks whose
//    bound |da| <= max_u |W0_u| * |z - z_pilot| exceeds the radius take the exact MUFU form, block by block, in the same kernel.
// 2. thread <-> hidden UNIT (TMEM lane), not sample: the pre-activation MMA is issued transposed (A = W0, B = dz tile), so a
//    thread reads 32 samples of ITS unit with one tcgen05.ld, keeps its six coefficients in registers for the whole item (no
//    per-group table loads: the shared-memory pipe carries only the operand stores) and writes 8 consecutive samples per
//    16-byte store into a K-major SWIZZLE_128B tile -- the layout both operands of G += H^T H read natively.
//    2.9 instead of 8 issue slots per (unit, sample).
//
#include "decode_gram.cuh"

#include <algorithm>
#include <cmath>

namespace glasdi {
namespace gr {

constexpr int kTabStride = 128;       // floats per table array (one per unit row)
constexpr int kTabArrays = 8;         // c0 .. c5, a(pilot), n(pilot)
constexpr int kFlagRing = 8;          // per stream element: may the four 32-sample blocks take the polynomial?

struct G2Smem {
  uint32_t h, a1, b1, bias, tab, flag, m, mx, c, bars, tmem, total;
};
__host__ __device__ inline G2Smem gcarve2(int nkgQ) {
  G2Smem s;
  uint32_t off = 0;
  s.h = off; off += kRing * kBufBytes;                          // Gram operand ring, 4 x 32 KB (1024-byte aligned tiles)
  s.a1 = off; off += kRing * kA1Bytes;                          // dz tiles (tf32 hi | lo)
  s.b1 = off; off += 8192;                                      // cf W0 (tf32 hi | remainder)
  s.bias = off; off += 128 * 4;
  s.tab = off; off += 4 * kTabArrays * kTabStride * 4;          // ring of 4 coefficient tables
  s.flag = off; off += kFlagRing * 4 * 4;
  s.m = off; off += 2 * 128 * 4u;
  s.mx = off; off += 2 * 16;
  off = (off + 15u) & ~15u;
  s.c = off; off += 2u * (uint32_t)nkgQ * 16u;
  s.bars = off; off += 32 * 8;
  s.tmem = off; off += 16;
  s.total = off;
  return s;
}

constexpr float kNegLn2 = -0.6931471805599453f;

// Taylor coefficients of f(a + x / cf) - f(a) in x, times 2^eH:  c[k] = hs f^(k)(a) / (k! cf^k), k = 1..5
template <int ACT>
__device__ __forceinline__ void taylor_coefs(float a_scaled, float hs, float (&c)[6]) {
  float d1, d2, d3, d4, d5, inv;
  if constexpr (ACT == GLASDI_ACT_SIGMOID) {
    inv = kNegLn2;                                              // a_scaled = -log2(e) a
    const float s = 1.0f / (1.0f + exp2f(a_scaled));
    d1 = s * (1.0f - s);
    d2 = d1 * (1.0f - 2.0f * s);
    d3 = d1 * (1.0f - 6.0f * d1);
    d4 = d2 * (1.0f - 12.0f * d1);
    d5 = d3 * (1.0f - 12.0f * d1) - 12.0f * d2 * d2;
  } else if constexpr (ACT == GLASDI_ACT_TANH) {
    inv = 1.0f;
    const float t = tanhf(a_scaled), g = 6.0f * t * t - 2.0f;
    d1 = 1.0f - t * t;
    d2 = -2.0f * t * d1;
    d3 = d1 * g;
    d4 = d2 * g + 12.0f * t * d1 * d1;
    d5 = d3 * g + 36.0f * t * d1 * d2 + 12.0f * d1 * d1 * d1;
  } else {                                                      // softplus: f' = sigmoid
    inv = 1.0f;
    const float s = 1.0f / (1.0f + expf(-a_scaled)), e1 = s * (1.0f - s);
    d1 = s;
    d2 = e1;
    d3 = e1 * (1.0f - 2.0f * s);
    d4 = e1 * (1.0f - 6.0f * e1);
    d5 = d3 * (1.0f - 12.0f * e1);
  }
  const float i2 = inv * inv;
  c[1] = hs * d1 * inv;
  c[2] = hs * d2 * (0.5f * i2);
  c[3] = hs * d3 * (i2 * inv * (1.0f / 6.0f));
  c[4] = hs * d4 * (i2 * i2 * (1.0f / 24.0f));
  c[5] = hs * d5 * (i2 * i2 * inv * (1.0f / 120.0f));
}

// exact form of one value in the pre-activation domain the MMA delivers (sigmoid: log2 domain)
template <int ACT>
__device__ __forceinline__ float act_scaled(int act_rt, float a_scaled) {
  if constexpr (ACT == GLASDI_ACT_SIGMOID) return rcp_ftz(1.0f + ex2_ftz(a_scaled));
  else return act_apply_t<ACT>(act_rt, a_scaled);
}

// Warp roles (24 warps; the register file gives 80 registers per thread to 21..24 warps alike):
//   warp 0          MMA issuer, an event loop over two independent queues: pre-activation MMAs (as soon as the dz tile is
//                   there and the TMEM buffer has been read) and Gram MMAs (as soon as an operand tile is full)
//   warps 1,2,3,20  epilogue of TMEM lane quadrants 1,2,3,0 (thread = row of the covariance) + the coefficient tables, two
//                   items ahead
//   warps 4..19     converters (phase B): thread = unit row, 32 samples per chunk
//   warps 21..23    phase A: latent differences dz of 32-sample blocks, dealt round-robin over the CTA's stream of chunks
// TMEM: two Gram accumulators (the epilogue of item n overlaps the MMAs of item n + 1) + two pre-activation buffers, each
// released as soon as the converters have LOADED it (not when they are done with it).  The first version (roles of
// gram_kernel, buffers released through the operand-full barrier) spent 2.4 k of its 2.5 k cycles per chunk in that
// dependency loop whatever the converters cost (profiles/r02_gram2_v1_*).
constexpr int kG2Threads = 24 * 32;
constexpr int kG2PhaseAWarp0 = 21, kG2PhaseAWarps = 3;
constexpr uint32_t kAccCol0 = 0, kD1Col0 = 256;

struct G2Bars {
  uint64_t h_full[kRing], h_empty[kRing], a1_full[kRing], a1_empty[kRing], d1_full[2], d1_empty[2], tab_full[4],
      acc_full[2], acc_empty[2];
};

template <int ACT, int INW>
__global__ void __launch_bounds__(kG2Threads, 1) gram2_kernel(const __grid_constant__ GramParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const G2Smem L = gcarve2(P.nkgQ);
  uint8_t* sH = smem + L.h;
  uint8_t* sA1 = smem + L.a1;
  uint8_t* sB1 = smem + L.b1;
  float* sBias = reinterpret_cast<float*>(smem + L.bias);
  float* sTab = reinterpret_cast<float*>(smem + L.tab);
  volatile int* sFlag = reinterpret_cast<volatile int*>(smem + L.flag);
  float* sM = reinterpret_cast<float*>(smem + L.m);
  float* sMax = reinterpret_cast<float*>(smem + L.mx);
  __half* sC = reinterpret_cast<__half*>(smem + L.c);
  G2Bars* B = reinterpret_cast<G2Bars*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);
  const int nchunks = (P.S + kCols - 1) / kCols;
  constexpr int NL = INW > 0 ? INW : kMaxLatent;
  const int NZ = P.NZ;
  const uint32_t n_my_items = (long long)blockIdx.x < P.n_items
                                  ? (uint32_t)((P.n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  const uint32_t n_elems = n_my_items * (uint32_t)nchunks;      // this CTA's stream of chunks

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&B->h_full[i], kGProdWarps);
      mbar_init(&B->h_empty[i], 1);
      mbar_init(&B->a1_full[i], 4);
      mbar_init(&B->a1_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&B->d1_full[i], 1);
      mbar_init(&B->d1_empty[i], kGProdWarps);
      mbar_init(&B->acc_full[i], 1);
      mbar_init(&B->acc_empty[i], kGEpiWarps);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&B->tab_full[i], kGEpiWarps);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, kGTmemColsFast);
    tmem_relinquish();
  }
  {
    const float* __restrict__ W = P.front.W[0];
    const float* __restrict__ b = P.front.b[0];
    // A operand of the pre-activation MMA: cf W0 as tf32 (hi | remainder), [unit][k < 8] K-major (k >= NZ and units >= K
    // zero); cf = -log2(e) for the sigmoid (2^x domain), 1 otherwise.  Biases (times cf) go to a table.
    const float cf = (ACT == GLASDI_ACT_SIGMOID) ? kNegLog2e : 1.0f;
    for (int idx = threadIdx.x; idx < 128 * 8; idx += blockDim.x) {
      const int u = idx >> 3, k = idx & 7;
      const float v = (u < P.K && k < P.NZ) ? cf * W[(size_t)u * P.NZ + k] : 0.f;
      const float vh = round_tf32(v);
      uint8_t* dst = sB1 + (k >> 2) * 2048 + (u >> 3) * 128 + (u & 7) * 16 + (k & 3) * 4;
      *reinterpret_cast<float*>(dst) = vh;                       // K chunks 0,1: tf32(W)
      *reinterpret_cast<float*>(dst + 4096) = v - vh;            // K chunks 2,3: the remainder
    }
    for (int k = threadIdx.x; k < 128; k += blockDim.x) sBias[k] = k < P.K ? cf * b[k] : 0.f;
    // operand tiles (unit rows > K stay zero for ever) and the packed-triangle staging (tail stays zero)
    uint4* z4 = reinterpret_cast<uint4*>(sH);
    for (uint32_t idx = threadIdx.x; idx < kRing * kBufBytes / 16; idx += blockDim.x) z4[idx] = make_uint4(0, 0, 0, 0);
    uint4* c4 = reinterpret_cast<uint4*>(sC);
    for (int idx = threadIdx.x; idx < 2 * P.nkgQ; idx += blockDim.x) c4[idx] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kG2PhaseAWarp0) {
    // ---- phase A: dz = z - z_pilot of 32-sample blocks as the tf32 B operand (hi in K 0..7, lo in K 8..15) of the
    // pre-activation MMA, and whether |da| of the block is inside the polynomial's radius for every unit (Cauchy-Schwarz
    // bound through the largest row norm of W0).  Block ids 4 e + b (stream element e, block b) are dealt round-robin to
    // the three warps; the latent states of a warp's next two blocks are in flight while it writes the current one. ----
    const uint32_t n_blocks = n_elems * 4u;
    const uint32_t w = (uint32_t)(warp - kG2PhaseAWarp0);
    float zq[2][NL], zp[2][NL];
    auto load_block = [&](uint32_t id, float (&z)[NL], float (&zpil)[NL]) {
      if (id >= n_blocks) return;
      const uint32_t e = id >> 2, bk = id & 3u;
      const long long it = (long long)blockIdx.x + (long long)(e / (uint32_t)nchunks) * gridDim.x;
      const int c = (int)(e % (uint32_t)nchunks);
      const float* __restrict__ base = P.Zs + (size_t)it * NZ * P.Spad;
      const float* __restrict__ src = base + c * kCols + bk * 32 + lane;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        z[i] = (INW > 0 || i < NZ) ? __ldg(src + (size_t)i * P.Spad) : 0.f;
        zpil[i] = (INW > 0 || i < NZ) ? __ldg(base + (size_t)i * P.Spad) : 0.f;      // sample 0 = the pilot
      }
    };
    load_block(w, zq[0], zp[0]);
    load_block(w + kG2PhaseAWarps, zq[1], zp[1]);
    uint32_t k = 0;
    for (uint32_t id = w; id < n_blocks; id += kG2PhaseAWarps, ++k) {
      const uint32_t e = id >> 2, bk = id & 3u;
      const uint32_t slot = e % kRing;
      float hi[NL], lo[NL];
      float n2 = 0.f;
      if (k & 1u) {
#pragma unroll
        for (int i = 0; i < NL; ++i) { const float dz = zq[1][i] - zp[1][i]; n2 = fmaf(dz, dz, n2); hi[i] = round_tf32(dz); lo[i] = dz - hi[i]; }
        load_block(id + 2 * kG2PhaseAWarps, zq[1], zp[1]);
      } else {
#pragma unroll
        for (int i = 0; i < NL; ++i) { const float dz = zq[0][i] - zp[0][i]; n2 = fmaf(dz, dz, n2); hi[i] = round_tf32(dz); lo[i] = dz - hi[i]; }
        load_block(id + 2 * kG2PhaseAWarps, zq[0], zp[0]);
      }
      const int all_ok = __all_sync(0xffffffffu, n2 <= P.poly_thr2);
      mbar_wait(&B->a1_empty[slot], ((e / kRing) & 1u) ^ 1u);
      const int col = (int)bk * 32 + lane;
      uint8_t* dst = sA1 + slot * kA1Bytes + (col >> 3) * 128 + (col & 7) * 16;
      auto comp = [&](const float (&v)[NL], int i) { return i < NL ? v[i < NL ? i : 0] : 0.f; };
      *reinterpret_cast<float4*>(dst) = make_float4(comp(hi, 0), comp(hi, 1), comp(hi, 2), comp(hi, 3));
      *reinterpret_cast<float4*>(dst + 2048) = make_float4(comp(hi, 4), comp(hi, 5), comp(hi, 6), comp(hi, 7));
      *reinterpret_cast<float4*>(dst + 4096) = make_float4(comp(lo, 0), comp(lo, 1), comp(lo, 2), comp(lo, 3));
      *reinterpret_cast<float4*>(dst + 6144) = make_float4(comp(lo, 4), comp(lo, 5), comp(lo, 6), comp(lo, 7));
      if (lane == 0) sFlag[(e & (kFlagRing - 1)) * 4 + bk] = all_ok;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->a1_full[slot]);
    }
  } else if (warp >= kGProdWarp0 && warp < kGProdWarp0 + kGProdWarps) {
    // ---- converters (phase B) --------------------------------------------------------------------------------
    const float hscale = P.h_scale;
    const int q = warp & 3;                                 // TMEM lane quadrant = unit rows 32 q .. 32 q + 31
    const int j = (warp - kGProdWarp0) >> 2;                // 32-sample block of the chunk this warp converts
    const int u = q * 32 + lane;                            // this thread's unit row
    const bool warp_active = q * 32 <= P.K;                 // the quadrant holds at least one real (or the constant) unit
    const bool row_active = u <= P.K;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    // K-major SWIZZLE_128B tile [128 unit rows x 128 samples] fp16: two 64-sample column blocks of 16 KB, rows of 128 B in
    // groups of eight (1 KB), 16-byte chunk index XOR (row mod 8).  This thread's four stores of a chunk:
    uint32_t st_off[4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
      st_off[m] = (uint32_t)(j >> 1) * 16384u + (uint32_t)(u >> 3) * 1024u + (uint32_t)(u & 7) * 128u +
                  ((((uint32_t)((j & 1) * 4 + m)) ^ (uint32_t)(u & 7)) << 4);
    uint32_t g = 0;                                         // chunk counter over the whole kernel
    float amax = 0.f;
    for (uint32_t n_it = 0; n_it < n_my_items; ++n_it) {
      mbar_wait(&B->tab_full[n_it & 3u], (n_it >> 2) & 1u);
      const float* T0 = sTab + (n_it & 3u) * (kTabArrays * kTabStride) + u;
      unsigned long long c2[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const float v = T0[k * kTabStride]; c2[k] = pack_f32x2(v, v); }
      const float a_p = T0[6 * kTabStride], n_p = T0[7 * kTabStride];
      for (int c = 0; c < nchunks; ++c, ++g) {
        const uint32_t db = g & 1u, hb = g % kRing;
        mbar_wait(&B->d1_full[db], (g >> 1) & 1u);
        tc_fence_after();
        const int poly = sFlag[(g & (kFlagRing - 1)) * 4 + j];
        uint32_t d[32];
        if (warp_active) {
          tmem_ld_32x32(tmem_base + lane_addr + kD1Col0 + db * 128u + (uint32_t)(j * 32), d);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->d1_empty[db]);       // the buffer is in registers: the next pre-activation MMA may overwrite it
        mbar_wait(&B->h_empty[hb], ((g / kRing) & 1u) ^ 1u);
        if (warp_active) {
          uint8_t* Hb = sH + hb * kBufBytes;
          if (poly) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              uint32_t out[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const unsigned long long x = pack_f32x2(__uint_as_float(d[8 * m + 2 * i]), __uint_as_float(d[8 * m + 2 * i + 1]));
                unsigned long long t = fma_f32x2(c2[5], x, c2[4]);
                t = fma_f32x2(t, x, c2[3]);
                t = fma_f32x2(t, x, c2[2]);
                t = fma_f32x2(t, x, c2[1]);
                t = fma_f32x2(t, x, c2[0]);
                float v0, v1;
                unpack_f32x2(t, v0, v1);
                out[i] = h2_as_u32(__floats2half2_rn(v0, v1));
              }
              if (row_active) *reinterpret_cast<uint4*>(Hb + st_off[m]) = make_uint4(out[0], out[1], out[2], out[3]);
            }
          } else {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              uint32_t out[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float v0 = fmaf(act_scaled<ACT>(P.front.act, a_p + __uint_as_float(d[8 * m + 2 * i])), hscale, n_p);
                const float v1 = fmaf(act_scaled<ACT>(P.front.act, a_p + __uint_as_float(d[8 * m + 2 * i + 1])), hscale, n_p);
                if constexpr (ACT != GLASDI_ACT_SIGMOID && ACT != GLASDI_ACT_TANH) {
                  if (P.check_overflow) amax = fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1)));
                }
                out[i] = h2_as_u32(__floats2half2_rn(v0, v1));
              }
              if (row_active) *reinterpret_cast<uint4*>(Hb + st_off[m]) = make_uint4(out[0], out[1], out[2], out[3]);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->h_full[hb]);
      }
    }
    if (P.check_overflow && !(amax < 60000.f)) atomicOr(P.flags, 1);
  } else if (warp == 0) {
    // ---- MMA issuer: whichever of the two queues can move -----------------------------------------------------------
    const uint32_t idesc = umma_idesc_f16_m128((uint32_t)P.Nmma);               // both operands K-major
    // K-major SWIZZLE_128B: LBO field = 1 (ignored), SBO = 1 KB between 8-row groups; a K = 16 slice = 32 bytes inside the
    // 128-byte swizzle row, the second 64-sample block 16 KB further
    const uint64_t desc0 = umma_desc_kmajor(smem_u32(sH), 16u, 1024u) | (2ull << 61);
    const uint32_t idesc1 = umma_idesc_tf32_m128(128u);
    const uint64_t wdesc1 = umma_desc_kmajor(smem_u32(sB1), 2048u, 128u);       // A: units x k
    const uint64_t zdesc1 = umma_desc_kmajor(smem_u32(sA1), 2048u, 128u);       // B: samples x k
    uint32_t g1 = 0;                                        // next pre-activation element
    uint32_t h_it = 0, n_it = 0;                            // next Gram element, its item
    int c = 0;                                              // ... and its chunk inside the item
    while (h_it < n_elems) {
      bool moved = false;
      if (g1 < n_elems && mbar_test(&B->a1_full[g1 % kRing], (g1 / kRing) & 1u) &&
          (g1 < 2u || mbar_test(&B->d1_empty[g1 & 1u], ((g1 >> 1) - 1u) & 1u))) {
        // D1[g1 & 1] = (cf W0) dz^T of stream element g1
        const uint32_t slot = g1 % kRing;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d1 = tmem_base + kD1Col0 + (g1 & 1u) * 128u;
          const uint64_t zd = zdesc1 + (uint64_t)(slot * (kA1Bytes >> 4));
          tc_mma_tf32(d1, wdesc1, zd, idesc1, 0u);                                         // W_hi dz_hi
          tc_mma_tf32(d1, wdesc1, zd + (uint64_t)(4096 >> 4), idesc1, 1u);                 // W_hi dz_lo
          tc_mma_tf32(d1, wdesc1 + (uint64_t)(4096 >> 4), zd, idesc1, 1u);                 // W_lo dz_hi
          tc_commit(&B->a1_empty[slot]);
          tc_commit(&B->d1_full[g1 & 1u]);
        }
        __syncwarp();
        ++g1;
        moved = true;
      }
      const uint32_t ab = n_it & 1u;
      if (mbar_test(&B->h_full[h_it % kRing], (h_it / kRing) & 1u) &&
          (c != 0 || n_it < 2u || mbar_test(&B->acc_empty[ab], ((n_it >> 1) - 1u) & 1u))) {
        const uint32_t buf = h_it % kRing;
        tc_fence_after();
        const uint32_t acc = tmem_base + kAccCol0 + ab * 128u;
        const int ncols = min(kCols, P.S - c * kCols);       // padding columns beyond hold exact zeros (ones in row K)
        const int nk = (ncols + 15) >> 4;
        const uint64_t d0 = desc0 + (uint64_t)(buf * (kBufBytes >> 4));
        if (elect_one()) {
#pragma unroll 1
          for (int k = 0; k < nk; ++k) {
            const uint64_t dk = d0 + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4);
            tc_mma_f16(acc, dk, dk, idesc, (c | k) != 0 ? 1u : 0u);
          }
          tc_commit(&B->h_empty[buf]);
          if (c == nchunks - 1) tc_commit(&B->acc_full[ab]);
        }
        __syncwarp();
        ++h_it;
        if (++c == nchunks) { c = 0; ++n_it; }
        moved = true;
      }
      if (!moved) __nanosleep(20);
    }
  } else {
    // ---- epilogue warps 1, 2, 3, 20: thread <-> row i of the Gram matrix (TMEM lane quadrant = warp mod 4); the same
    // thread makes row i of the coefficient tables, two items ahead ---------------------------------------------------
    const int qd = warp & 3;
    const int i = qd * 32 + lane;
    const int K = P.K;
    const float inv_S = P.inv_S;
    const float hscale = P.h_scale;
    const int off_i = (i * (i + 1)) / 2;                    // packed LOWER triangle, row-major: q(i, j <= i) = i (i + 1) / 2 + j
    // Coefficient table of an item, row i: c0 (1 for the constant unit K, else 0), c1..c5 of the polynomial in the MMA's
    // pre-activation difference, and for the exact form a(pilot) and n = c0 - h(pilot) 2^eH (same arithmetic as the samples:
    // padding columns, copies of sample 0, cancel to exactly 0 / 1 in either form).
    auto make_tables = [&](uint32_t nseq) {
      if (nseq >= n_my_items) return;
      if (nseq >= 4u) {
        // slot reuse: the converters loaded item nseq - 4's table at its start, and its epilogue (ours) is long done
      }
      const long long it = (long long)blockIdx.x + (long long)nseq * gridDim.x;
      const float* __restrict__ base = P.Zs + (size_t)it * NZ * P.Spad;
      float* T0 = sTab + (nseq & 3u) * (kTabArrays * kTabStride);
      float a = sBias[i];
      const uint8_t* w = sB1 + (i >> 3) * 128 + (i & 7) * 16;
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        const float zk = (INW > 0 || k < NZ) ? __ldg(base + (size_t)k * P.Spad) : 0.f;     // the pilot = sample 0
        const uint8_t* wi = w + (k >> 2) * 2048 + (k & 3) * 4;
        a = fmaf(*reinterpret_cast<const float*>(wi) + *reinterpret_cast<const float*>(wi + 4096), zk, a);
      }
      float cc[6];
      taylor_coefs<ACT>(a, hscale, cc);
      cc[0] = (i == K) ? 1.0f : 0.0f;
      const float hp = act_scaled<ACT>(P.front.act, a);
#pragma unroll
      for (int k = 0; k < 6; ++k) T0[k * kTabStride + i] = (i < K || k == 0) ? cc[k] : 0.f;
      T0[6 * kTabStride + i] = a;
      T0[7 * kTabStride + i] = cc[0] - hp * hscale;
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->tab_full[nseq & 3u]);
    };
    make_tables(0);
    make_tables(1);
    uint32_t n_it = 0;
    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x, ++n_it) {
      make_tables(n_it + 2u);
      const uint32_t ab = n_it & 1u;
      mbar_wait_sleep(&B->acc_full[ab], (n_it >> 1) & 1u, 256);      // an item takes several us
      tc_fence_after();
      const uint32_t taddr = tmem_base + kAccCol0 + ab * 128u + (((uint32_t)(qd * 32)) << 16);
      float* sMb = sM + (n_it & 1u) * 128;
      float* sMaxb = sMax + (n_it & 1u) * 4;
      __half* sCb = sC + (size_t)(n_it & 1u) * P.nkgQ * 8;
      uint32_t r[32];
      // step 1: own diagonal entry, own column sum, the row of column sums (from the constant unit's row)
      tmem_ld_32x32(taddr + qd * 32, r);
      uint32_t mi_bits;
      tmem_ld_32x1(taddr + K, mi_bits);
      tmem_ld_wait();
      float gii = 0.f;
#pragma unroll
      for (int qq = 0; qq < 32; ++qq) gii = (qq == lane) ? __uint_as_float(r[qq]) : gii;
      const float mi = __uint_as_float(mi_bits);
      const float cii = (i < K) ? (gii - mi * mi * inv_S) * inv_S : 0.f;
      const float wm = warp_max(fmaxf(cii, 0.f));
      if (lane == 0) sMaxb[qd] = wm;
      if (__any_sync(0xffffffffu, i < K && !(fabsf(gii) < 3.0e38f)) && lane == 0) atomicOr(P.flags, 1);   // NaN / inf activations
      if (qd == (K >> 5)) {
        for (int b = 0; b < 4; ++b) {
          tmem_ld_32x32(taddr + b * 32, r);
          tmem_ld_wait();
          if (lane == (K & 31)) {
#pragma unroll
            for (int qq = 0; qq < 32; ++qq) sMb[b * 32 + qq] = __uint_as_float(r[qq]);
          }
        }
      }
      asm volatile("bar.sync 3, 128;" ::: "memory");
      const float cmax = fmaxf(fmaxf(sMaxb[0], sMaxb[1]), fmaxf(sMaxb[2], sMaxb[3]));
      int e = 0;
      float f = 0.f;
      if (cmax > 0.f && cmax < 3.0e38f) {
        e = max(-60, min(60, 13 - ilogbf(cmax)));
        f = ldexpf(inv_S, e);
      }
      // step 2: lower-triangle entries j <= i of row i, scaled, fp16, packed at off(i) + j
      const float mis = mi * inv_S;
      for (int b = 0; b <= qd; ++b) {                   // blocks right of the diagonal block hold only j > i
        if (b * 32 >= K) break;
        tmem_ld_32x32(taddr + b * 32, r);
        tmem_ld_wait();
        if (i < K) {
          const float4* m4 = reinterpret_cast<const float4*>(sMb + b * 32);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 mj = m4[q4];
            const float mjv[4] = {mj.x, mj.y, mj.z, mj.w};
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const int jj = b * 32 + q4 * 4 + qq;
              if (jj <= i) {
                const float val = fmaf(-mis, mjv[qq], __uint_as_float(r[q4 * 4 + qq])) * f;
                sCb[off_i + jj] = __float2half_rn(val);
              }
            }
          }
        }
      }
      tc_fence_before();                                // this warp's TMEM reads of the item are done
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->acc_empty[ab]);    // the MMA warp may reuse this accumulator (item n_it + 2)
      asm volatile("bar.sync 4, 128;" ::: "memory");
      // copy-out: 16-byte pieces of the packed row into the quad kernel's A tiles
      const long long ib = item >> 7;
      const int row = (int)(item & 127);
      uint4* dst = reinterpret_cast<uint4*>(P.A) + ((size_t)ib * P.nkgQ) * 128 + row;
      const uint4* src = reinterpret_cast<const uint4*>(sCb);
      for (int kg = i; kg < P.nkgQ; kg += 128) dst[(size_t)kg * 128] = src[kg];
      if (i == 0) P.fscale[item] = (f > 0.f) ? ldexpf(1.0f, -(P.exp_base + e)) : 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kGTmemColsFast);
}

}  // namespace gr

size_t decode_gram2_smem_bytes(int K) { return gr::gcarve2(gram_nkgQ(K)).total; }

template <int ACT, int INW>
static int launch_gram2_inst(glasdi_handle* h, const gr::GramParams& G, int ctas, size_t smem, cudaStream_t st) {
  GL_CUDA(h, cudaFuncSetAttribute(gr::gram2_kernel<ACT, INW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gr::gram2_kernel<ACT, INW><<<ctas, gr::kG2Threads, smem, st>>>(G);
  return GLASDI_OK;
}

// one front layer + sigmoid / tanh / softplus: Taylor-polynomial producers (exact MUFU form outside the radius)
bool gram2_supported(const Decoder& d) {
  return d.n_linear == 2 && (d.act == GLASDI_ACT_SIGMOID || d.act == GLASDI_ACT_TANH || d.act == GLASDI_ACT_SOFTPLUS);
}

int launch_decode_gram2(glasdi_handle* h, gr::GramParams& G, int ctas, cudaStream_t st) {
  const Decoder& d = h->dec;
  // radius of the polynomial in |da| (natural pre-activation units), as a bound on |z - z_pilot|^2 through the largest row
  // norm of W0 (Cauchy-Schwarz); GLASDI_OPT_TAYLOR_RADIUS = 0 switches the polynomial off (exact form everywhere)
  const int NZ = d.widths[0];
  double w2max = 0.0;
  for (int u = 0; u < d.K; ++u) {
    double s = 0.0;
    for (int i = 0; i < NZ; ++i) s += (double)d.h_front_last_W[(size_t)u * NZ + i] * d.h_front_last_W[(size_t)u * NZ + i];
    w2max = std::max(w2max, s);
  }
  const double radius = h->opt_taylor_radius_q8 / 256.0;
  G.poly_thr2 = (h->opt_taylor_radius_q8 <= 0) ? -1.0f : (w2max > 0.0 ? (float)(radius * radius / w2max) : 3.0e38f);
  const size_t smem = gr::gcarve2(G.nkgQ).total;
  int rc;
  if (d.act == GLASDI_ACT_SIGMOID && G.NZ == 5) rc = launch_gram2_inst<GLASDI_ACT_SIGMOID, 5>(h, G, ctas, smem, st);
  else if (d.act == GLASDI_ACT_SIGMOID) rc = launch_gram2_inst<GLASDI_ACT_SIGMOID, 0>(h, G, ctas, smem, st);
  else if (d.act == GLASDI_ACT_TANH) rc = launch_gram2_inst<GLASDI_ACT_TANH, 0>(h, G, ctas, smem, st);
  else rc = launch_gram2_inst<GLASDI_ACT_SOFTPLUS, 0>(h, G, ctas, smem, st);
  if (rc) return rc;
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "gram2_kernel launch");
}

}  // namespace glasdi
