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
// Warp roles, barriers and the epilogue are those of gram_kernel's fast path (decode_gram.cu).
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

template <int ACT, int INW>
__global__ void __launch_bounds__(kGThreads, 1) gram2_kernel(const __grid_constant__ GramParams P) {
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
  GBars* B = reinterpret_cast<GBars*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);
  const int nchunks = (P.S + kCols - 1) / kCols;

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(&B->h_full[i], kGProdWarps);
      mbar_init(&B->h_empty[i], 1);
      mbar_init(&B->a1_full[i], 4);
      mbar_init(&B->a1_empty[i], 1);
    }
    mbar_init(&B->d1_full[0], 1);
    mbar_init(&B->d1_full[1], 1);
    mbar_init(&B->d1_full[2], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&B->tab_full[i], 4);
    mbar_init(&B->acc_full, 1);
    mbar_init(&B->acc_empty, kGEpiWarps - 1);
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

  if (warp >= kGProdWarp0) {
    // ---- producers ---------------------------------------------------------------------------------
    const float hscale = P.h_scale;
    const int q = warp & 3;                                 // TMEM lane quadrant = unit rows 32 q .. 32 q + 31
    const int j = (warp - kGProdWarp0) >> 2;                // 32-sample block of the chunk this warp converts
    const int u = q * 32 + lane;                            // this thread's unit row (phase B / tables)
    const int col = q * 32 + lane;                          // this thread's sample column (phase A)
    const bool phase_a = j == 0;                            // warps 4..7 also write the latent differences
    const bool tab_warp = j == 1;                           // warps 8..11 make the coefficient tables, one item ahead
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
    uint32_t g = 0, n_it = 0;                               // chunk counter over the whole kernel, item counter
    long long item = blockIdx.x;
    // Phase A (warps 4..7) walks the CTA's stream of chunks TWO AHEAD of phase B (see gram_kernel).
    constexpr int NL = INW > 0 ? INW : kMaxLatent;
    const int NZ = P.NZ;
    float zq[2][NL], zpa[NL], zpn[NL], zpt[NL];
    long long item_l = item;
    int c_l = 0;
    uint32_t ga = 0;
    int c_a = 0;
    long long item_a = item;
    const uint32_t n_elems = item < P.n_items
                                 ? (uint32_t)((P.n_items - item + gridDim.x - 1) / gridDim.x) * (uint32_t)nchunks : 0u;
    auto load_col = [&](long long it, int c, float (&z)[NL]) {
      const float* __restrict__ src = P.Zs + (size_t)it * NZ * P.Spad + c * kCols + col;
#pragma unroll
      for (int i = 0; i < NL; ++i) z[i] = (INW > 0 || i < NZ) ? __ldg(src + (size_t)i * P.Spad) : 0.f;
    };
    auto load_zp = [&](long long it, float (&z)[NL]) {     // sample 0 = the pilot
      const float* __restrict__ src = P.Zs + (size_t)it * NZ * P.Spad;
#pragma unroll
      for (int i = 0; i < NL; ++i) z[i] = (INW > 0 || i < NZ) ? __ldg(src + (size_t)i * P.Spad) : 0.f;
    };
    auto load_next = [&](float (&z)[NL]) {
      if (item_l < P.n_items) {
        load_col(item_l, c_l, z);
        if (++c_l == nchunks) { c_l = 0; item_l += gridDim.x; }
      }
    };
    if (phase_a && item < P.n_items) {
      load_next(zq[0]);
      load_next(zq[1]);
      load_zp(item, zpa);
      if (item + gridDim.x < P.n_items) load_zp(item + gridDim.x, zpn);
    }
    // Coefficient table of an item, row u: c0 (1 for the constant unit K, else 0), c1..c5 of the polynomial in the MMA's
    // pre-activation difference, and for the exact form a(pilot) and n = c0 - h(pilot) 2^eH (same arithmetic as the samples:
    // padding columns, copies of sample 0, cancel to exactly 0 / 1 in either form).
    auto make_tables = [&](uint32_t nseq) {
      float* T0 = sTab + (nseq & 3u) * (kTabArrays * kTabStride);
      float a = sBias[u];
      const uint8_t* w = sB1 + (u >> 3) * 128 + (u & 7) * 16;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const uint8_t* wi = w + (i >> 2) * 2048 + (i & 3) * 4;
        a = fmaf(*reinterpret_cast<const float*>(wi) + *reinterpret_cast<const float*>(wi + 4096), zpt[i], a);
      }
      float c[6];
      taylor_coefs<ACT>(a, hscale, c);
      c[0] = (u == P.K) ? 1.0f : 0.0f;
      const float hp = act_scaled<ACT>(P.front.act, a);
#pragma unroll
      for (int k = 0; k < 6; ++k) T0[k * kTabStride + u] = (u < P.K || k == 0) ? c[k] : 0.f;
      T0[6 * kTabStride + u] = a;
      T0[7 * kTabStride + u] = c[0] - hp * hscale;
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->tab_full[nseq & 3u]);
    };
    if (tab_warp && item < P.n_items) {
      load_zp(item, zpt);
      make_tables(0);
      if (item + gridDim.x < P.n_items) load_zp(item + gridDim.x, zpt);
    }
    // dz = z - z_pilot of this thread's sample as the tf32 B operand (hi in K 0..7, lo in K 8..15) of stream element ga,
    // and whether |da| of its 32-sample block is inside the polynomial's radius for every unit (Cauchy-Schwarz bound)
    auto phase_a_step = [&]() {
      const uint32_t slot = ga % kRing;
      float hi[NL], lo[NL];
      float n2 = 0.f;
      if (ga & 1u) {
#pragma unroll
        for (int i = 0; i < NL; ++i) { const float dz = zq[1][i] - zpa[i]; n2 = fmaf(dz, dz, n2); hi[i] = round_tf32(dz); lo[i] = dz - hi[i]; }
        load_next(zq[1]);
      } else {
#pragma unroll
        for (int i = 0; i < NL; ++i) { const float dz = zq[0][i] - zpa[i]; n2 = fmaf(dz, dz, n2); hi[i] = round_tf32(dz); lo[i] = dz - hi[i]; }
        load_next(zq[0]);
      }
      const int all_ok = __all_sync(0xffffffffu, n2 <= P.poly_thr2);
      mbar_wait_relaxed(&B->a1_empty[slot], ((ga / kRing) & 1u) ^ 1u);
      uint8_t* dst = sA1 + slot * kA1Bytes + (col >> 3) * 128 + (col & 7) * 16;
      auto comp = [&](const float (&v)[NL], int i) { return i < NL ? v[i < NL ? i : 0] : 0.f; };
      *reinterpret_cast<float4*>(dst) = make_float4(comp(hi, 0), comp(hi, 1), comp(hi, 2), comp(hi, 3));
      *reinterpret_cast<float4*>(dst + 2048) = make_float4(comp(hi, 4), comp(hi, 5), comp(hi, 6), comp(hi, 7));
      *reinterpret_cast<float4*>(dst + 4096) = make_float4(comp(lo, 0), comp(lo, 1), comp(lo, 2), comp(lo, 3));
      *reinterpret_cast<float4*>(dst + 6144) = make_float4(comp(lo, 4), comp(lo, 5), comp(lo, 6), comp(lo, 7));
      if (lane == 0) sFlag[(ga & (kFlagRing - 1)) * 4 + q] = all_ok;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->a1_full[slot]);
      ++ga;
      if (++c_a == nchunks) {
        c_a = 0;
        item_a += gridDim.x;
#pragma unroll
        for (int i = 0; i < NL; ++i) zpa[i] = zpn[i];
        if (item_a + gridDim.x < P.n_items) load_zp(item_a + gridDim.x, zpn);
      }
    };
    if (phase_a && n_elems > 0) phase_a_step();
    if (phase_a && n_elems > 1) phase_a_step();
    float amax = 0.f;
    for (; item < P.n_items; item += gridDim.x, ++n_it) {
      if (tab_warp && item + gridDim.x < P.n_items) {
        make_tables(n_it + 1u);
        if (item + 2 * (long long)gridDim.x < P.n_items) load_zp(item + 2 * (long long)gridDim.x, zpt);
      }
      while (!mbar_try_wait(&B->tab_full[n_it & 3u], (n_it >> 2) & 1u)) __nanosleep(32);
      const float* T0 = sTab + (n_it & 3u) * (kTabArrays * kTabStride) + u;
      unsigned long long c2[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { const float v = T0[k * kTabStride]; c2[k] = pack_f32x2(v, v); }
      const float a_p = T0[6 * kTabStride], n_p = T0[7 * kTabStride];
      for (int c = 0; c < nchunks; ++c, ++g) {
        if (phase_a && ga < n_elems) phase_a_step();        // stream element g + 2
        const uint32_t db = g % 3u, hb = g % kRing;
        while (!mbar_try_wait(&B->d1_full[db], (g / 3u) & 1u)) __nanosleep(32);
        tc_fence_after();
        mbar_wait_relaxed(&B->h_empty[hb], ((g / kRing) & 1u) ^ 1u);
        if (warp_active) {
          const int poly = sFlag[(g & (kFlagRing - 1)) * 4 + j];
          uint32_t d[32];
          tmem_ld_32x32(tmem_base + lane_addr + 128u + db * 128u + (uint32_t)(j * 32), d);
          tmem_ld_wait();
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
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->h_full[hb]);
      }
    }
    if (P.check_overflow && !(amax < 60000.f)) atomicOr(P.flags, 1);
  } else {
    // ---- consumer warps 0..3 (as gram_kernel's fast path; operands: see the header comment) --------------------
    const int i = warp * 32 + lane;
    const int K = P.K;
    const float inv_S = P.inv_S;
    const int off_i = (i * (i + 1)) / 2;
    const uint32_t idesc = umma_idesc_f16_m128((uint32_t)P.Nmma);               // both operands K-major
    // K-major SWIZZLE_128B: LBO field = 1 (ignored), SBO = 1 KB between 8-row groups; a K = 16 slice = 32 bytes inside the
    // 128-byte swizzle row, the second 64-sample block 16 KB further
    const uint64_t desc0 = umma_desc_kmajor(smem_u32(sH), 16u, 1024u) | (2ull << 61);
    const uint32_t taddr = tmem_base + (((uint32_t)(warp * 32)) << 16);
    uint32_t h_it = 0, n_it = 0, g1 = 0;
    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x, ++n_it) {
      if (warp == 0) {
        tc_fence_after();
        const uint32_t idesc1 = umma_idesc_tf32_m128(128u);
        const uint64_t wdesc1 = umma_desc_kmajor(smem_u32(sB1), 2048u, 128u);       // A: units x k
        const uint64_t zdesc1 = umma_desc_kmajor(smem_u32(sA1), 2048u, 128u);       // B: samples x k
        const uint32_t n_elems = (uint32_t)((P.n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) * (uint32_t)nchunks;
        auto issue_mma1 = [&]() {
          const uint32_t slot = g1 % kRing;
          while (!mbar_try_wait(&B->a1_full[slot], (g1 / kRing) & 1u)) __nanosleep(32);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d1 = tmem_base + 128u + (g1 % 3u) * 128u;
            const uint64_t zd = zdesc1 + (uint64_t)(slot * (kA1Bytes >> 4));
            tc_mma_tf32(d1, wdesc1, zd, idesc1, 0u);                                         // W_hi dz_hi
            tc_mma_tf32(d1, wdesc1, zd + (uint64_t)(4096 >> 4), idesc1, 1u);                 // W_hi dz_lo
            tc_mma_tf32(d1, wdesc1 + (uint64_t)(4096 >> 4), zd, idesc1, 1u);                 // W_lo dz_hi
            tc_commit(&B->a1_empty[slot]);
            tc_commit(&B->d1_full[g1 % 3u]);
          }
          __syncwarp();
          ++g1;
        };
        if (g1 == 0) {
          issue_mma1();
          if (g1 < n_elems) issue_mma1();
        }
        for (int c = 0; c < nchunks; ++c) {
          if (g1 < n_elems && g1 <= h_it + 2u) issue_mma1();
          const uint32_t buf = h_it % kRing;
          while (!mbar_try_wait(&B->h_full[buf], (h_it / kRing) & 1u)) __nanosleep(64);
          if (c == 0 && n_it > 0) mbar_wait(&B->acc_empty, (n_it - 1u) & 1u);
          tc_fence_after();
          const int ncols = min(kCols, P.S - c * kCols);       // padding columns beyond hold exact zeros (ones in row K)
          const int nk = (ncols + 15) >> 4;
          const uint64_t d0 = desc0 + (uint64_t)(buf * (kBufBytes >> 4));
          if (elect_one()) {
#pragma unroll 1
            for (int k = 0; k < nk; ++k) {
              const uint64_t dk = d0 + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4);
              tc_mma_f16(tmem_base, dk, dk, idesc, (c | k) != 0 ? 1u : 0u);
            }
            tc_commit(&B->h_empty[buf]);
          }
          __syncwarp();
          ++h_it;
        }
        if (elect_one()) tc_commit(&B->acc_full);
        __syncwarp();
        if (g1 < n_elems && g1 <= h_it + 2u) issue_mma1();
      }
      while (!mbar_try_wait(&B->acc_full, n_it & 1u)) __nanosleep(warp == 0 ? 64 : 1024);
      tc_fence_after();
      float* sMb = sM + (n_it & 1u) * 128;
      float* sMaxb = sMax + (n_it & 1u) * 4;
      __half* sCb = sC + (size_t)(n_it & 1u) * P.nkgQ * 8;
      uint32_t r[32];
      // step 1: own diagonal entry, own column sum, the row of column sums (from the constant unit's row)
      tmem_ld_32x32(taddr + warp * 32, r);
      uint32_t mi_bits;
      tmem_ld_32x1(taddr + K, mi_bits);
      tmem_ld_wait();
      float gii = 0.f;
#pragma unroll
      for (int qq = 0; qq < 32; ++qq) gii = (qq == lane) ? __uint_as_float(r[qq]) : gii;
      const float mi = __uint_as_float(mi_bits);
      const float cii = (i < K) ? (gii - mi * mi * inv_S) * inv_S : 0.f;
      const float wm = warp_max(fmaxf(cii, 0.f));
      if (lane == 0) sMaxb[warp] = wm;
      if (__any_sync(0xffffffffu, i < K && !(fabsf(gii) < 3.0e38f)) && lane == 0) atomicOr(P.flags, 1);
      if (warp == (K >> 5)) {
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
      for (int b = 0; b <= warp; ++b) {
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
      tc_fence_before();
      if (warp == 0) {
        asm volatile("bar.arrive 4, 128;" ::: "memory");
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->acc_empty);
        asm volatile("bar.sync 4, 128;" ::: "memory");
        const long long ib = item >> 7;
        const int row = (int)(item & 127);
        uint4* dst = reinterpret_cast<uint4*>(P.A) + ((size_t)ib * P.nkgQ) * 128 + row;
        const uint4* src = reinterpret_cast<const uint4*>(sCb);
        for (int kg = threadIdx.x - 32; kg < P.nkgQ; kg += (kGEpiWarps - 1) * 32) dst[(size_t)kg * 128] = src[kg];
        if (threadIdx.x == 32) P.fscale[item] = (f > 0.f) ? ldexpf(1.0f, -(P.exp_base + e)) : 0.f;
      }
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
  gr::gram2_kernel<ACT, INW><<<ctas, gr::kGThreads, smem, st>>>(G);
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
