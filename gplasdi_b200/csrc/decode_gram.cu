// decode_gram.cu -- SCREENING pass of the greedy step in covariance form (GLASDI_MODE_SCREENED_GRAM).
//
// Why: the decoded ensemble x[dof, sample] of one (parameter, time step) has D*S outputs but only
// K = 100 MACs per output; a tcgen05 accumulator leaves TMEM at 64 B/clk/SM (tcgen05.ld), so ANY kernel
// that materialises x in TMEM is drain-bound at K/256 = 44 % of the tensor peak (measured: 57 ms for the
// 44.5 TFLOP bench step, profiles/).  The variance over samples does not need x:
//
//     sigma^2[dof] = w_dof^T C w_dof,     C = cov_s(h)  (K x K),   h = hidden activations of the last layer,
//
// (reference semantics: std(0) of decoder(Z) over the samples, src/lasdi/gplasdi.py:65-67 with the decoder of
// src/lasdi/latent_space.py:157-164 whose last layer is linear).  Two tensor-core GEMMs whose outputs are tiny:
//
//   gram_kernel   per (p, t):  G = H^T H over the S samples (M = N = units, K = samples: MN-major operands,
//                 both read from the SAME shared-memory tile), an extra constant unit gives the column sums;
//                 epilogue: C = G/S - m m^T/S^2, per-item power-of-two scale, fp16, packed upper triangle
//                 written as the A operand of the second GEMM.
//   quad_kernel   var[(p,t), dof] = sum_q Ctilde[(p,t), q] V[dof, q],  V[dof, (i,j)] = (2 - delta_ij) w_i w_j,
//                 Q = K(K+1)/2 = 5050: a plain [P*T x Q] x [Q x D] GEMM; epilogue: un-scale, write the variance
//                 field, atomicMax per parameter.
//
// FLOP: K^2 S + Q D per (p, t) instead of K S D: 5x fewer than the direct product at the bench shape.
// Accuracy: one fp16 pass with cancellation in the quadratic form -> ~1e-4 relative on sigma; this is a
// SCREENING pass only -- refine.cu re-evaluates every near-maximum (t, dof) exactly and checks the deviation.
#include "decode_gram.cuh"

#include <type_traits>

#include <algorithm>
#include <cmath>

namespace glasdi {
namespace gr {

struct GSmem {
  uint32_t h, a1, b1, wl, pilot, m, mx, c, xb, bars, tmem, total;
  int wl_ld, kpu, xw;
};
__host__ __device__ inline GSmem gcarve(int nkg, int last_in_w, int nkgQ, int xw = 0) {
  GSmem s;
  uint32_t off = 0;
  s.kpu = nkg * 8;
  s.h = off; off += kRing * kBufBytes;
  // fast path: latent differences (tf32 A operand ring) + front-layer weights (tf32 hi | lo B operand);
  // generic path with deep fronts: two activation buffers x (128 columns + pilot) x widest hidden layer -- same bytes
  s.a1 = off;
  s.b1 = off + kRing * kA1Bytes;
  s.xb = off;
  s.xw = xw;
  {
    const uint32_t fast_bytes = kRing * kA1Bytes + 4u * 2048u, gen_bytes = 2u * 129u * (uint32_t)xw * 4u;
    off += fast_bytes > gen_bytes ? fast_bytes : ((gen_bytes + 15u) & ~15u);
  }
  s.wl_ld = wl_stride(last_in_w) > 12 ? wl_stride(last_in_w) : 12;   // fast path rows: 12 floats (duplicated pairs)
  s.wl = off; off += (uint32_t)s.kpu * (uint32_t)s.wl_ld * 4u;
  s.pilot = off; off += 4u * (uint32_t)s.kpu * 8u;                     // fast path: 4 buffers x (a_pilot, -h_pilot 2^eH) tables
  s.m = off; off += 2 * 128 * 4u;                                       // column sums, double buffered by item parity
  s.mx = off; off += 2 * 16;
  off = (off + 15u) & ~15u;
  s.c = off; off += 2u * (uint32_t)nkgQ * 16u;                          // packed triangle staging, double buffered
  s.bars = off; off += 32 * 8;
  s.tmem = off; off += 16;
  s.total = off;
  return s;
}

// Byte offset of the 16-byte vector (unit group kg = 8 consecutive units, sample column col) in an operand tile.
// Canonical MN-major SWIZZLE_128B layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in elements: a swizzle atom is
// 8 samples x 64 units (8 rows of 128 B), the 16-byte chunk index is XORed with the row (sample mod 8).
// The un-swizzled MN-major layout works too but every MMA then takes ~480 instead of ~60 cycles: for a fixed
// sample the 16 unit groups sit a multiple of 128 B apart, i.e. in the same four banks (16-way conflict in the
// operand fetch; measured: the kernel's time did not move when the producers' work was halved).
__device__ __forceinline__ uint32_t h_offset(int kg, int col) {
  const uint32_t r = (uint32_t)col & 7u;
  return ((uint32_t)kg >> 3) * 16384u + ((uint32_t)col >> 3) * 1024u + r * 128u + ((((uint32_t)kg & 7u) ^ r) << 4);
}

// Fast producer path (one front layer, latent dimension 5): TWO sample columns per thread as packed fp32 pairs
// (FFMA2 / FADD2), weights staged as duplicated pairs (w, w) so that one LDS.128 feeds two packed FMAs, the
// sigmoid's -log2(e) folded into the staged weights, MUFU ex2/rcp in their .ftz form (the non-ftz __expf costs
// three more instructions per value).  d2[u] = (h_A, h_B)[kg*8+u] * 2^eH - pilot: 8.5 issue slots per
// (unit, sample) instead of 19.6 (profiles/: gram_r1).
template <int ACT, int INW, bool FAST>
__global__ void __launch_bounds__(kGThreads, 1) gram_kernel(const __grid_constant__ GramParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int last_in_w = P.front.widths[P.front.n_front - 1];
  const GSmem L = gcarve(P.nkg, last_in_w, P.nkgQ, P.xw);
  uint8_t* sH = smem + L.h;
  float* sWl = reinterpret_cast<float*>(smem + L.wl);
  float* sPilot = reinterpret_cast<float*>(smem + L.pilot);
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
    tmem_alloc(tmem_slot, FAST ? kGTmemColsFast : kGTmemCols);
    tmem_relinquish();
  }
  uint8_t* sA1 = smem + L.a1;
  uint8_t* sB1 = smem + L.b1;
  {
    const int l = P.front.n_front - 1;
    const float* __restrict__ W = P.front.W[l];
    const float* __restrict__ b = P.front.b[l];
    const int ld = L.wl_ld;
    if constexpr (FAST) {
      // B operand of the pre-activation MMA: cf W0 as tf32 (hi | remainder), [unit][k < 8] K-major (k >= NZ and units
      // >= K zero); cf = -log2(e) for the sigmoid (2^x domain), 1 otherwise.  Biases (times cf) go to a table.
      const float cf = (ACT == GLASDI_ACT_SIGMOID) ? kNegLog2e : 1.0f;
      for (int idx = threadIdx.x; idx < 128 * 8; idx += blockDim.x) {
        const int u = idx >> 3, k = idx & 7;
        const float v = (u < P.K && k < P.NZ) ? cf * W[(size_t)u * P.NZ + k] : 0.f;
        const float vh = round_tf32(v);
        uint8_t* dst = sB1 + (k >> 2) * 2048 + (u >> 3) * 128 + (u & 7) * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(dst) = vh;                       // K chunks 0,1: tf32(W)
        *reinterpret_cast<float*>(dst + 4096) = v - vh;            // K chunks 2,3: the remainder
      }
      for (int k = threadIdx.x; k < L.kpu; k += blockDim.x) sWl[k] = k < P.K ? cf * b[k] : 0.f;
    } else {
      for (int idx = threadIdx.x; idx < L.kpu * ld; idx += blockDim.x) {
        const int k = idx / ld, i = idx - k * ld;
        float v = 0.f;
        if (k < P.K) {
          if (i < last_in_w) v = W[(size_t)k * last_in_w + i];
          else if (i == last_in_w) v = b[k];
        }
        sWl[idx] = v;
      }
    }
    // operand tiles (unit groups >= nkg stay zero for ever) and the packed-triangle staging (tail stays zero)
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
    // ---- producers: pilot-centred, scaled activations of 128 sample columns per chunk --------------
    const int ptid = threadIdx.x - kGProdWarp0 * 32;
    const float hscale = P.h_scale;
    const int kg_one = P.K >> 3, u_one = P.K & 7;
    uint32_t h_it = 0, n_it = 0;
    float amax = 0.f;
    if constexpr (FAST) {
      // Pre-activations come from the TENSOR CORE: a(s) = a(pilot) + (-log2 e) W0 (z(s) - z(pilot)), the second term
      // as one tf32 MMA pair per chunk (samples on the TMEM lanes, units on the columns; dz split hi + lo so that only
      // W0 is rounded to tf32: a 5e-5 systematic effect on sigma, inside the screening budget).  That removes the
      // 5 FMAs + 3 weight loads per (unit, sample) that dominated the instruction count (profiles/: gram_r1e: 60 k
      // warp-instructions per (p, t), 60 % issue utilisation) and gives the natural mapping
      //   thread <-> one sample (TMEM lane) x 8 consecutive units (one 16-byte operand vector).
      const int q = warp & 3;                               // TMEM lane quadrant = 32-sample block of the chunk
      const int jg = (warp - kGProdWarp0) >> 2;             // 0..3: unit groups jg, jg + 4, ... (+ rotating left-over)
      const int col = q * 32 + lane;
      const bool phase_a = jg == 0;                         // warps 4..7 also write the latent differences
      const int nkg4 = P.nkg & ~3;
      const unsigned long long hs2 = pack_f32x2(hscale, hscale);
      const unsigned long long one2 = pack_f32x2(1.0f, 1.0f);
      const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
      uint32_t g = 0;                                       // chunk counter over the whole kernel
      long long item = blockIdx.x;
      // Phase A (warps 4..7) walks the CTA's stream of chunks ONE AHEAD of phase B, so that the pre-activation MMA
      // of chunk c + 1 runs while chunk c is being processed.  Its latent states are prefetched two stream elements
      // ahead (zq[2], a two-entry queue), the pilot's one item ahead (zpa = item of the queue head, zpn = the next).
      constexpr int NL = INW > 0 ? INW : kMaxLatent;       // fast path: INW = compile-time latent dimension (5) or 0 = up to 8
      const int NZ = P.NZ;
      float zq[2][NL], zpa[NL], zpn[NL], zpt[NL];          // zpt: pilot of the item whose tables are made next (warps 8..11)
      long long item_l = item;                              // position of the next stream element to LOAD
      int c_l = 0;
      uint32_t ga = 0;                                      // stream index of the next phase-A element (= queue head)
      int c_a = 0;                                          // its chunk index inside its item
      long long item_a = item;                              // its item
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
      auto load_next = [&](float (&z)[NL]) {                 // next element of the stream, if any
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
      // Pilot tables a(pilot) (log2 domain) and -(h(pilot) 2^eH) of an item: made ONE ITEM AHEAD by warps 8..11 into
      // a ring of four buffers and published through an mbarrier, so that no CTA-wide barrier ties the producer
      // warps together at item boundaries (they drift up to two chunks apart).  Same pairwise arithmetic as the
      // samples: the padding columns (copies of sample 0, dz = 0) cancel to exactly 0.  The constant unit K (zero
      // weights, h = act(0)) gets 1 - h 2^eH: its "difference" is the constant 1.
      const bool tab_warp = jg == 1;
      auto make_tables = [&](uint32_t nseq) {
        float* tA = sPilot + (nseq & 3u) * 2 * L.kpu;
        float* tN = tA + L.kpu;
        const int k = ptid - 128;
        if (k < L.kpu) {
          // a(pilot) of unit k and of its pair partner from the staged weights (hi + remainder = the fp32 weight)
          auto preact = [&](int u) {
            float a = sWl[u];
            const uint8_t* w = sB1 + (u >> 3) * 128 + (u & 7) * 16;
#pragma unroll
            for (int i = 0; i < NL; ++i) {
              const uint8_t* wi = w + (i >> 2) * 2048 + (i & 3) * 4;
              a = fmaf(*reinterpret_cast<const float*>(wi) + *reinterpret_cast<const float*>(wi + 4096), zpt[i], a);
            }
            return a;
          };
          const float ak = preact(k);
          float hk;
          if constexpr (ACT == GLASDI_ACT_SIGMOID) {
            const float ap = preact(k ^ 1);
            const float tk = ex2_ftz(ak) + 1.0f, tp = ex2_ftz(ap) + 1.0f;
            const float te = (k & 1) ? tp : tk, to = (k & 1) ? tk : tp;
            const float r = rcp_ftz(te * to);
            hk = r * ((k & 1) ? te : to);
          } else {
            hk = act_apply_t<ACT>(P.front.act, ak);
          }
          tA[k] = ak;
          tN[k] = (k == P.K ? 1.0f : 0.0f) - hk * hscale;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->tab_full[nseq & 3u]);
      };
      if (tab_warp && item < P.n_items) {
        load_zp(item, zpt);
        make_tables(0);
        if (item + gridDim.x < P.n_items) load_zp(item + gridDim.x, zpt);
      }
      // dz = z - z_pilot of this thread's sample as the tf32 A operand (hi in K 0..7, lo in K 8..15) of stream element ga
      auto phase_a_step = [&]() {
        const uint32_t slot = ga % kRing;
        float hi[NL], lo[NL];
        if (ga & 1u) {
#pragma unroll
          for (int i = 0; i < NL; ++i) { const float dz = zq[1][i] - zpa[i]; hi[i] = round_tf32(dz); lo[i] = dz - hi[i]; }
          load_next(zq[1]);
        } else {
#pragma unroll
          for (int i = 0; i < NL; ++i) { const float dz = zq[0][i] - zpa[i]; hi[i] = round_tf32(dz); lo[i] = dz - hi[i]; }
          load_next(zq[0]);
        }
        mbar_wait_relaxed(&B->a1_empty[slot], ((ga / kRing) & 1u) ^ 1u);
        uint8_t* dst = sA1 + slot * kA1Bytes + (col >> 3) * 128 + (col & 7) * 16;
        auto comp = [&](const float (&v)[NL], int i) { return i < NL ? v[i < NL ? i : 0] : 0.f; };
        *reinterpret_cast<float4*>(dst) = make_float4(comp(hi, 0), comp(hi, 1), comp(hi, 2), comp(hi, 3));
        *reinterpret_cast<float4*>(dst + 2048) = make_float4(comp(hi, 4), comp(hi, 5), comp(hi, 6), comp(hi, 7));
        *reinterpret_cast<float4*>(dst + 4096) = make_float4(comp(lo, 0), comp(lo, 1), comp(lo, 2), comp(lo, 3));
        *reinterpret_cast<float4*>(dst + 6144) = make_float4(comp(lo, 4), comp(lo, 5), comp(lo, 6), comp(lo, 7));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->a1_full[slot]);
        ++ga;
        if (++c_a == nchunks) {                              // the queue head moves to the next item: its pilot
          c_a = 0;
          item_a += gridDim.x;
#pragma unroll
          for (int i = 0; i < NL; ++i) zpa[i] = zpn[i];
          if (item_a + gridDim.x < P.n_items) load_zp(item_a + gridDim.x, zpn);
        }
      };
      if (phase_a && n_elems > 0) phase_a_step();           // stream elements 0 and 1: phase A runs TWO ahead of phase B
      if (phase_a && n_elems > 1) phase_a_step();
      for (; item < P.n_items; item += gridDim.x, ++n_it) {
        if (tab_warp && item + gridDim.x < P.n_items) {       // tables of the NEXT item (zpt was prefetched an item ago)
          make_tables(n_it + 1u);
          if (item + 2 * (long long)gridDim.x < P.n_items) load_zp(item + 2 * (long long)gridDim.x, zpt);
        }
        const float* tabA = sPilot + (n_it & 3u) * 2 * L.kpu;
        const float* tabN = tabA + L.kpu;
        mbar_wait_sleep(&B->tab_full[n_it & 3u], (n_it >> 2) & 1u, 32);
        for (int c = 0; c < nchunks; ++c, ++g) {
          if (phase_a && ga < n_elems) phase_a_step();        // stream element g + 2
          // ---- phase B: sigmoid of (a_pilot + da), pilot-centred, scaled, fp16, into the Gram operand tile
          const uint32_t db = g % 3u, hb = g % kRing;
          mbar_wait_sleep(&B->d1_full[db], (g / 3u) & 1u, 32);
          tc_fence_after();
          mbar_wait_relaxed(&B->h_empty[hb], ((g / kRing) & 1u) ^ 1u);
          uint8_t* Hb = sH + hb * kBufBytes;
          const uint32_t taddr = tmem_base + lane_addr + 128u + db * 128u;   // D1[g mod 3]
          auto do_group = [&](int kg) {
            uint32_t d[8];
            tmem_ld_32x8(taddr + kg * 8, d);
            const float4 a0 = *reinterpret_cast<const float4*>(tabA + kg * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(tabA + kg * 8 + 4);
            tmem_ld_wait();
            unsigned long long acc[4];
            acc[0] = add_f32x2(pack_f32x2(a0.x, a0.y), pack_f32x2(__uint_as_float(d[0]), __uint_as_float(d[1])));
            acc[1] = add_f32x2(pack_f32x2(a0.z, a0.w), pack_f32x2(__uint_as_float(d[2]), __uint_as_float(d[3])));
            acc[2] = add_f32x2(pack_f32x2(a1.x, a1.y), pack_f32x2(__uint_as_float(d[4]), __uint_as_float(d[5])));
            acc[3] = add_f32x2(pack_f32x2(a1.z, a1.w), pack_f32x2(__uint_as_float(d[6]), __uint_as_float(d[7])));
            const float4 n0 = *reinterpret_cast<const float4*>(tabN + kg * 8);
            const float4 n1 = *reinterpret_cast<const float4*>(tabN + kg * 8 + 4);
            const unsigned long long np[4] = {pack_f32x2(n0.x, n0.y), pack_f32x2(n0.z, n0.w), pack_f32x2(n1.x, n1.y),
                                              pack_f32x2(n1.z, n1.w)};
            uint32_t out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float xe, xo;
              unpack_f32x2(acc[j], xe, xo);
              unsigned long long h2;
              if constexpr (ACT == GLASDI_ACT_SIGMOID) {
                float te, to;
                unpack_f32x2(add_f32x2(pack_f32x2(ex2_ftz(xe), ex2_ftz(xo)), one2), te, to);
                const float r = rcp_ftz(te * to);
                h2 = pack_f32x2(r * to, r * te);
              } else {
                h2 = pack_f32x2(act_apply_t<ACT>(P.front.act, xe), act_apply_t<ACT>(P.front.act, xo));
              }
              float de, dod;
              unpack_f32x2(fma_f32x2(h2, hs2, np[j]), de, dod);
              if constexpr (ACT != GLASDI_ACT_SIGMOID && ACT != GLASDI_ACT_TANH) {
                if (P.check_overflow) amax = fmaxf(amax, fmaxf(fabsf(de), fabsf(dod)));
              }
              out[j] = h2_as_u32(__floats2half2_rn(de, dod));
            }
            *reinterpret_cast<uint4*>(Hb + h_offset(kg, col)) = make_uint4(out[0], out[1], out[2], out[3]);
          };
          // The warps of group 0 also write dz (phase A, ~2/3 of a group's work): every 4th chunk their last fixed
          // group goes to one of the other warp groups instead, so that all four carry the same average load.
          const int give = nkg4 - 4;                                   // group 0's last fixed unit group
          const bool given = nkg4 >= 8 && (g & 3u) == 3u;
          const int taker = 1 + (int)((g >> 2) % 3u);
#pragma unroll 1
          for (int kg = jg; kg < nkg4; kg += 4) {
            if (given && kg == give) continue;
            do_group(kg);
          }
          if (given && jg == taker) do_group(give);
          if (jg > 0) {
            // left-over groups rotate over the warp groups 1..3; group 0 (the phase-A warps) has its extra work already.
            // Per chunk the warps are therefore NOT equally loaded: three pre-activation buffers (below) let them
            // drift up to two chunks apart, so that only the average matters.
            const int extra = nkg4 + (int)((uint32_t)(jg - 1 + 3 - (int)(g % 3u)) % 3u);
            if (extra < P.nkg) do_group(extra);
          }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&B->h_full[hb]);
        }
      }
    } else {
      // generic front MLP: thread <-> one sample column x a quarter of the unit groups.  With two or more front
      // layers the layers before the last one are evaluated COOPERATIVELY: the four threads of a column each compute
      // a quarter of every layer's units and exchange them through shared memory (one named barrier per layer).
      // Every thread evaluating the whole front on its own (and from local-memory arrays) cost 265 us per (p, t) at
      // the Burgers2D-shaped decoder [5,20,20,20,100,3600].
      const int col = ptid & 127;
      const int part = ptid >> 7;
      const int kg_begin = (P.nkg * part) / 4, kg_end = (P.nkg * (part + 1)) / 4;
      const int NZ = P.NZ;
      const int ld = L.wl_ld;
      const int n_pre = P.front.n_front - 1;               // layers before the last front layer
      float* xbuf = reinterpret_cast<float*>(smem + L.xb);
      const int xw = L.xw;
      // layers 0 .. n_pre-1 for column `cx` of the activation buffers (cx = 128: the pilot); returns the final buffer
      auto coop_layers = [&](const float (&z)[kMaxLatent], bool also_pilot, const float (&zpil)[kMaxLatent]) {
        asm volatile("bar.sync 2, %0;" ::"n"(kGProdThreads) : "memory");      // readers of the previous final buffer
        for (int l = 0; l < n_pre; ++l) {
          const int wi = P.front.widths[l], wo = P.front.widths[l + 1];
          const float* __restrict__ W = P.front.W[l];
          const float* __restrict__ bb = P.front.b[l];
          const int o_begin = (wo * part) / 4, o_end = (wo * (part + 1)) / 4;
          float* dst = xbuf + (size_t)(l & 1) * 129 * xw;
          const float* src = xbuf + (size_t)((l & 1) ^ 1) * 129 * xw;
          for (int pass = 0; pass < (also_pilot ? 2 : 1); ++pass) {
            const int cx = pass ? 128 : col;
            for (int o0 = o_begin; o0 < o_end; o0 += 4) {
              float acc[4];
              const float* wr[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int o = min(o0 + u, o_end - 1);
                acc[u] = __ldg(bb + o);
                wr[u] = W + (size_t)o * wi;
              }
              for (int i = 0; i < wi; ++i) {
                const float xi = (l == 0) ? (pass ? zpil[i] : z[i]) : src[cx * xw + i];
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] = fmaf(__ldg(wr[u] + i), xi, acc[u]);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (o0 + u < o_end) dst[cx * xw + o0 + u] = act_apply_t<ACT>(P.front.act, acc[u]);
            }
          }
          asm volatile("bar.sync 2, %0;" ::"n"(kGProdThreads) : "memory");
        }
      };
      const float* xfinal = xbuf + (size_t)((n_pre - 1) & 1) * 129 * xw;       // valid when n_pre >= 1
      // narrow fronts (every layer before the last front layer at most 24 wide): each thread evaluates them for its own
      // sample in registers from a transposed shared-memory image of the weights (decode_common.cuh) -- no barrier, no
      // global-memory weight loads; the four threads of a column repeat the 900 FMAs of [5,20,20,20] rather than share them
      const bool regs_path = INW == 0 && n_pre > 0 && xw <= kNarrowW && NZ <= kMaxLatent &&
                             (size_t)n_pre * kNarrowLayerFloats * 4 <= (size_t)(kRing * kA1Bytes + 4u * 2048u);
      // Deep narrow fronts on the TENSOR CORES (warp-level mma.sync chain, decode_common.cuh): a warp owns 16 sample
      // columns of the chunk and half of the last front layer's unit groups; the narrow layers chain register to
      // register with split fp16 operands (fp32-faithful), the pilot goes through the same arithmetic (16 copies of
      // sample 0), and (act(.) - pilot) 2^eH lands in the operand tile as 4-byte stores.  The per-thread CUDA-core
      // version below ran the Burgers2D-shaped item in 333 k clocks, 15x its FMA bound.
      bool mma_path = INW == 0 && n_pre > 0 && NZ <= kMaxLatent && pm_out_pad(P.K, true) / 8 <= P.nkg;
      {
        size_t need = 0;
        for (int l = 0; l < P.front.n_front; ++l) {
          const bool last = l == P.front.n_front - 1;
          if (P.front.widths[l] > kMaxHid) mma_path = false;
          need += (pm_layer_halfs(P.front.widths[l], P.front.widths[l + 1], last) * 2 + 15) / 16 * 16;
          need += ((size_t)pm_out_pad(P.front.widths[l + 1], last) * 4 + 15) / 16 * 16;
        }
        if (need > (size_t)(L.wl - L.xb)) mma_path = false;
      }
      if (mma_path) {
        const int nf = P.front.n_front;
        __shared__ PmLayer sLy[kMaxLinear];
        PmLayer* Ly = sLy;
        pm_stage(smem + L.xb, P.front, nf, ptid, kGProdThreads, Ly);
        asm volatile("bar.sync 2, %0;" ::"n"(kGProdThreads) : "memory");
        const int pw = ptid >> 5;                            // producer warp 0..15
        // warps 0..7 take the even chunks of the CTA's stream, warps 8..15 the odd ones (the operand ring is four deep), each
        // warp one m16 tile with ALL unit groups: the narrow chain is evaluated once per sample (two warps sharing a tile
        // by halves of the unit groups repeated it: 204 instead of 138 MMA triples per pair of chunks)
        const int mt = pw & 7, grp = pw >> 3;
        const int g = lane >> 2, t = lane & 3;
        const int ntl = Ly[nf - 1].op / 8;                   // n8 tiles the last front layer really has (units < K)
        const int j_begin = 0, j_end = P.nkg;
        const int act_rt = P.front.act;
        // bare MUFU forms (6 instead of ~12 instructions per softplus; the pilot goes through the same function)
        auto act = [&](float x) { return act_mufu<ACT>(x, act_rt); };
        const int k0 = 2 * t, k1 = 2 * t + 1;
        bool narrow32 = true;                                // every layer input <= 32 wide: two k16 blocks of fragments
        for (int l = 0; l < nf; ++l) narrow32 = narrow32 && P.front.widths[l] <= 32;
        auto run = [&](auto kbm_tag) {
        constexpr int KBM = decltype(kbm_tag)::value;
        uint32_t Ah[KBM][4], Al[KBM][4];
        for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
          const float* __restrict__ zb = P.Zs + (size_t)item * NZ * P.Spad;
          {
            // pilot table: the chain on 16 copies of sample 0; row 0 (lanes 0..3) holds units 8 j + 2 t, + 1.  EVERY warp
            // runs the narrow chain and takes one of the last layer's tiles (one warp doing all of them kept the other
            // fifteen at the barrier below for a whole chunk time: 13 % of the stall samples)
            const float zp0 = k0 < NZ ? __ldg(zb + (size_t)k0 * P.Spad) : 0.f, zp1 = k1 < NZ ? __ldg(zb + (size_t)k1 * P.Spad) : 0.f;
            pm_latent_frag(zp0, zp1, zp0, zp1, Ah, Al);
            pm_chain(Ly, nf - 1, g, t, Ah, Al, act);
            for (int j = pw; j < P.nkg; j += kGProdWarps) {
              float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4];
              if (j < ntl) pm_tiles<KBM, false, false>(Ly[nf - 1], j, j, g, t, Ah, Al, acc, acc1);
              if (g == 0) {
                sPilot[8 * j + k0] = (8 * j + k0 < P.K) ? act(acc[0]) * hscale : 0.f;
                sPilot[8 * j + k1] = (8 * j + k1 < P.K) ? act(acc[1]) * hscale : 0.f;
              }
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kGProdThreads) : "memory");
          for (int c = 0; c < nchunks; ++c) {
            if ((h_it & 1u) != (uint32_t)grp) { ++h_it; continue; }       // the other warp group's chunk
            const int col_lo = mt * 16 + g, col_hi = col_lo + 8;
            const int s_lo = c * kCols + col_lo, s_hi = s_lo + 8;
            const int sl = s_lo < P.S ? s_lo : 0, sh = s_hi < P.S ? s_hi : 0;
            pm_latent_frag(k0 < NZ ? __ldg(zb + (size_t)k0 * P.Spad + sl) : 0.f, k1 < NZ ? __ldg(zb + (size_t)k1 * P.Spad + sl) : 0.f,
                           k0 < NZ ? __ldg(zb + (size_t)k0 * P.Spad + sh) : 0.f, k1 < NZ ? __ldg(zb + (size_t)k1 * P.Spad + sh) : 0.f,
                           Ah, Al);
            pm_chain(Ly, nf - 1, g, t, Ah, Al, act);
            const uint32_t buf = h_it % kRing;
            uint8_t* Hb = sH + buf * kBufBytes;
            mbar_wait_relaxed(&B->h_empty[buf], ((h_it / kRing) & 1u) ^ 1u);
            auto emit = [&](int j, const float (&acc)[4]) {
              const float2 pv = *reinterpret_cast<const float2*>(sPilot + 8 * j + k0);
              const int k = 8 * j + k0;
              float v0 = fmaf(act(acc[0]), hscale, -pv.x), v1 = fmaf(act(acc[1]), hscale, -pv.y);
              float v2 = fmaf(act(acc[2]), hscale, -pv.x), v3 = fmaf(act(acc[3]), hscale, -pv.y);
              if (k >= P.K) { v0 = (k == P.K) ? 1.0f : 0.f; v2 = v0; }            // the constant unit, then padding
              if (k + 1 >= P.K) { v1 = (k + 1 == P.K) ? 1.0f : 0.f; v3 = v1; }
              if (s_lo >= P.S) { v0 = 0.f; v1 = 0.f; }
              if (s_hi >= P.S) { v2 = 0.f; v3 = 0.f; }
              if (P.check_overflow) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v0), fabsf(v1)), fmaxf(fabsf(v2), fabsf(v3))));
              *reinterpret_cast<uint32_t*>(Hb + h_offset(j, col_lo) + 4 * t) = h2_as_u32(__floats2half2_rn(v0, v1));
              *reinterpret_cast<uint32_t*>(Hb + h_offset(j, col_hi) + 4 * t) = h2_as_u32(__floats2half2_rn(v2, v3));
            };
            for (int j = j_begin; j < j_end; j += 2) {
              float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
              if (j + 1 < ntl) pm_tiles<KBM, false, true>(Ly[nf - 1], j, j + 1, g, t, Ah, Al, acc, acc1);
              else if (j < ntl) pm_tiles<KBM, false, false>(Ly[nf - 1], j, j, g, t, Ah, Al, acc, acc1);
              emit(j, acc);
              if (j + 1 < j_end) emit(j + 1, acc1);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane < 2) mbar_arrive(&B->h_full[buf]);                    // eight warps fill a tile: two arrivals each (count 16)
            ++h_it;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kGProdThreads) : "memory");     // pilot is rewritten next item
        }
        };
        if (narrow32) run(std::integral_constant<int, 2>{});
        else run(std::integral_constant<int, kPmMaxKb>{});
      } else {
      if (regs_path) {
        narrow_layers_stage(xbuf, P.front, n_pre, ptid, kGProdThreads);
        asm volatile("bar.sync 2, %0;" ::"n"(kGProdThreads) : "memory");
      }
      for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const float* __restrict__ zb = P.Zs + (size_t)item * NZ * P.Spad;
        float znext[kMaxLatent], zpil[kMaxLatent];
#pragma unroll
        for (int i = 0; i < kMaxLatent; ++i) {
          znext[i] = (col < P.S && i < NZ) ? __ldg(zb + (size_t)i * P.Spad + col) : 0.f;
          zpil[i] = (i < NZ) ? __ldg(zb + (size_t)i * P.Spad) : 0.f;          // sample 0 = the pilot
        }
        for (int c = 0; c < nchunks; ++c) {
          const int s = c * kCols + col;
          const bool valid = s < P.S;
          float zin[kMaxLatent];
#pragma unroll
          for (int i = 0; i < kMaxLatent; ++i) zin[i] = znext[i];
          if (c + 1 < nchunks) {               // next chunk's latent state: in flight while this chunk is computed
#pragma unroll
            for (int i = 0; i < kMaxLatent; ++i)
              znext[i] = (s + kCols < P.S && i < NZ) ? __ldg(zb + (size_t)i * P.Spad + s + kCols) : 0.f;
          }
          const float* x = zin;
          if constexpr (INW == 0) {
            if (!regs_path && n_pre > 0) {
              coop_layers(zin, c == 0 && col == 0, zpil);
              x = xfinal + col * xw;
            }
          }
          if (c == 0) {
            // pilot = activations of sample 0 (its front layers were evaluated with chunk 0's, column 128)
            float xp[kNarrowW];
            if (INW == 0 && regs_path && ptid < ((P.nkg + 31) & ~31)) {           // the warps that make the pilot table
#pragma unroll
              for (int i = 0; i < kNarrowW; ++i) xp[i] = (i < kMaxLatent) ? zpil[i < kMaxLatent ? i : 0] : 0.f;
              narrow_layers_regs<ACT>(xbuf, n_pre, P.front.widths, P.front.act, xp);
            }
            for (int kg = ptid; kg < P.nkg; kg += kGProdThreads) {
              float hv[8];
              if (INW == 0 && regs_path)
                last_front_units_regs<ACT>(sWl, ld, last_in_w, P.front.act, kg * 8, xp, hv);
              else if (INW == 0 && n_pre > 0)
                last_front_units<ACT, 0>(sWl, ld, last_in_w, P.front.act, kg * 8, xfinal + 128 * xw, hv);
              else
                last_front_units<ACT, INW>(sWl, ld, last_in_w, P.front.act, kg * 8, zpil, hv);
#pragma unroll
              for (int u = 0; u < 8; ++u) sPilot[kg * 8 + u] = hv[u] * hscale;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kGProdThreads) : "memory");
          }
          float xr[kNarrowW];                               // after the pilot block: its registers are free again
          if constexpr (INW == 0) {
            if (regs_path) {
#pragma unroll
              for (int i = 0; i < kNarrowW; ++i) xr[i] = (i < kMaxLatent) ? zin[i < kMaxLatent ? i : 0] : 0.f;
              narrow_layers_regs<ACT>(xbuf, n_pre, P.front.widths, P.front.act, xr);
            }
          }
          const uint32_t buf = h_it % kRing;
          uint8_t* Hb = sH + buf * kBufBytes;
          mbar_wait_relaxed(&B->h_empty[buf], ((h_it / kRing) & 1u) ^ 1u);
          for (int kg = kg_begin; kg < kg_end; ++kg) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (valid) {
              float hv[8];
              if (INW == 0 && regs_path) last_front_units_regs<ACT>(sWl, ld, last_in_w, P.front.act, kg * 8, xr, hv);
              else last_front_units<ACT, INW>(sWl, ld, last_in_w, P.front.act, kg * 8, x, hv);
              const float4 p0 = *reinterpret_cast<const float4*>(sPilot + kg * 8);
              const float4 p1 = *reinterpret_cast<const float4*>(sPilot + kg * 8 + 4);
              const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
              float dv8[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                dv8[u] = fmaf(hv[u], hscale, -pv[u]);
                if (P.check_overflow) amax = fmaxf(amax, fabsf(dv8[u]));
              }
              if (kg == kg_one) {
#pragma unroll
                for (int u = 0; u < 8; ++u) dv8[u] = (u == u_one) ? 1.0f : dv8[u];      // the constant unit
              }
              v.x = h2_as_u32(__floats2half2_rn(dv8[0], dv8[1]));
              v.y = h2_as_u32(__floats2half2_rn(dv8[2], dv8[3]));
              v.z = h2_as_u32(__floats2half2_rn(dv8[4], dv8[5]));
              v.w = h2_as_u32(__floats2half2_rn(dv8[6], dv8[7]));
            }
            *reinterpret_cast<uint4*>(Hb + h_offset(kg, col)) = v;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&B->h_full[buf]);
          ++h_it;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kGProdThreads) : "memory");     // pilot is rewritten next item
      }
      }   // !mma_path
    }
    if (P.check_overflow && !(amax < 60000.f)) atomicOr(P.flags, 1);
  } else {
    // ---- consumer warps 0..3.  Warp 0 first issues the item's MMAs (G += H_chunk^T H_chunk as the producers
    // deliver the chunks); then thread <-> unit row i of the Gram matrix for the epilogue.  The next item's
    // MMAs start after this item's epilogue: the 4-deep chunk ring keeps the producers busy meanwhile. --------
    const int i = warp * 32 + lane;
    const int K = P.K;
    const float inv_S = P.inv_S;
    const int off_i = (i * (i + 1)) / 2;                    // packed LOWER triangle, row-major: q(i, j <= i) = i (i + 1) / 2 + j
    const uint32_t idesc = umma_idesc_f16_m128_mn((uint32_t)P.Nmma);
    // MN-major, 128-byte swizzle (see h_offset): LBO = step between 64-unit blocks (16 KB), SBO = step between
    // 8-sample groups (1 KB)
    const uint64_t desc0 = umma_desc_kmajor(smem_u32(sH), 16384u, 1024u) | (2ull << 61);
    const uint32_t taddr = tmem_base + (((uint32_t)(warp * 32)) << 16);
    uint32_t h_it = 0, n_it = 0, g1 = 0;
    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x, ++n_it) {
      if (warp == 0 && FAST) {
        // The pre-activation MMAs run ONE stream element (chunk) ahead of the Gram MMAs over the whole kernel:
        // D1[g & 1] = dz (hi, lo) x (-log2 e W0)^T (hi, lo), then G += H^T H.  D1[(g + 1) & 1] is free when element
        // g + 1 is issued: its last readers (phase B of element g - 1) arrived on h_full before the previous Gram MMAs.
        tc_fence_after();                               // the other warps' TMEM reads of the previous item (bar 4)
        const uint32_t idesc1 = umma_idesc_tf32_m128((uint32_t)P.Nmma);
        const uint64_t adesc1 = umma_desc_kmajor(smem_u32(sA1), 2048u, 128u);
        const uint64_t bdesc1 = umma_desc_kmajor(smem_u32(sB1), 2048u, 128u);
        const uint32_t n_elems = (uint32_t)((P.n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) * (uint32_t)nchunks;
        auto issue_mma1 = [&]() {
          const uint32_t slot = g1 % kRing;
          mbar_wait_sleep(&B->a1_full[slot], (g1 / kRing) & 1u, 32);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d1 = tmem_base + 128u + (g1 % 3u) * 128u;
            const uint64_t a = adesc1 + (uint64_t)(slot * (kA1Bytes >> 4));
            tc_mma_tf32(d1, a, bdesc1, idesc1, 0u);                                          // dz_hi W_hi
            tc_mma_tf32(d1, a + (uint64_t)(4096 >> 4), bdesc1, idesc1, 1u);                  // dz_lo W_hi
            tc_mma_tf32(d1, a, bdesc1 + (uint64_t)(4096 >> 4), idesc1, 1u);                  // dz_hi W_lo
            tc_commit(&B->a1_empty[slot]);
            tc_commit(&B->d1_full[g1 % 3u]);
          }
          __syncwarp();
          ++g1;
        };
        if (g1 == 0) {                                  // stream elements 0 and 1
          issue_mma1();
          if (g1 < n_elems) issue_mma1();
        }
        for (int c = 0; c < nchunks; ++c) {
          if (g1 < n_elems && g1 <= h_it + 2u) issue_mma1();      // element h_it + 2 (its buffer: element h_it - 1, drained)
          const uint32_t buf = h_it % kRing;
          mbar_wait_sleep(&B->h_full[buf], (h_it / kRing) & 1u, 64);   // a chunk takes ~1 us
          if (c == 0 && n_it > 0) mbar_wait(&B->acc_empty, (n_it - 1u) & 1u);             // previous accumulator drained
          tc_fence_after();
          const int ncols = min(kCols, P.S - c * kCols);       // padding columns beyond hold exact zeros
          const int nk = (ncols + 15) >> 4;
          const uint64_t d0 = desc0 + (uint64_t)(buf * (kBufBytes >> 4));
          if (elect_one()) {
#pragma unroll 1
            for (int k = 0; k < nk; ++k) {
              const uint64_t dk = d0 + (uint64_t)(k * 128);             // 16 samples = two 1 KB swizzle atoms per K slice
              tc_mma_f16(tmem_base, dk, dk, idesc, (c | k) != 0 ? 1u : 0u);
            }
            tc_commit(&B->h_empty[buf]);
          }
          __syncwarp();
          ++h_it;
        }
        if (elect_one()) tc_commit(&B->acc_full);
        __syncwarp();
        // the epilogue below keeps this warp away for ~2 chunk times: put a second pre-activation MMA in flight
        // (both D1 buffers are free: every phase B of this item has arrived) so that the producers do not starve
        if (g1 < n_elems && g1 <= h_it + 2u) issue_mma1();
      } else if (warp == 0) {
        tc_fence_after();                               // the other warps' TMEM reads of the previous item (bar 4)
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t buf = h_it % kRing;
          mbar_wait_sleep(&B->h_full[buf], (h_it / kRing) & 1u, 64);   // a chunk takes ~1.5 us
          if (c == 0 && n_it > 0) mbar_wait(&B->acc_empty, (n_it - 1u) & 1u);             // previous accumulator drained
          tc_fence_after();
          const int ncols = min(kCols, P.S - c * kCols);     // fast path: padding columns beyond hold exact zeros
          const int nk = (ncols + 15) >> 4;
          const uint64_t d0 = desc0 + (uint64_t)(buf * (kBufBytes >> 4));
          if (elect_one()) {
#pragma unroll 1
            for (int k = 0; k < nk; ++k) {
              const uint64_t dk = d0 + (uint64_t)(k * 128);           // 16 samples = two 1 KB swizzle atoms per K slice
              tc_mma_f16(tmem_base, dk, dk, idesc, (c | k) != 0 ? 1u : 0u);
            }
            tc_commit(&B->h_empty[buf]);
          }
          __syncwarp();
          ++h_it;
        }
        if (elect_one()) tc_commit(&B->acc_full);
        __syncwarp();
      }
      mbar_wait_sleep(&B->acc_full, n_it & 1u, warp == 0 ? 64 : 1024);   // an item takes ~10 us
      tc_fence_after();
      // Warp 0 also issues the MMAs, so its share of the epilogue is kept minimal: row i packs the LOWER triangle
      // j <= i (rows 0..31: one 32-column block), it only ARRIVES on the barrier that separates packing from the
      // copy-out, and warps 1..3 do the copy-out.  Everything per-item in shared memory is double buffered.
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
      for (int q = 0; q < 32; ++q) gii = (q == lane) ? __uint_as_float(r[q]) : gii;
      const float mi = __uint_as_float(mi_bits);
      const float cii = (i < K) ? (gii - mi * mi * inv_S) * inv_S : 0.f;
      const float wm = warp_max(fmaxf(cii, 0.f));
      if (lane == 0) sMaxb[warp] = wm;
      if (__any_sync(0xffffffffu, i < K && !(fabsf(gii) < 3.0e38f)) && lane == 0) atomicOr(P.flags, 1);   // NaN / inf activations
      if (warp == (K >> 5)) {
        for (int b = 0; b < 4; ++b) {
          tmem_ld_32x32(taddr + b * 32, r);
          tmem_ld_wait();
          if (lane == (K & 31)) {
#pragma unroll
            for (int q = 0; q < 32; ++q) sMb[b * 32 + q] = __uint_as_float(r[q]);
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
      for (int b = 0; b <= warp; ++b) {                 // blocks right of the diagonal block hold only j > i
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
              const int j = b * 32 + q4 * 4 + qq;
              if (j <= i) {
                const float val = fmaf(-mis, mjv[qq], __uint_as_float(r[q4 * 4 + qq])) * f;
                sCb[off_i + j] = __float2half_rn(val);
              }
            }
          }
        }
      }
      tc_fence_before();                                // this warp's TMEM reads of the item are done
      if (warp == 0) {
        asm volatile("bar.arrive 4, 128;" ::: "memory");                 // packed its rows; back to issuing MMAs
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&B->acc_empty);                         // warp 0 may overwrite the accumulator
        asm volatile("bar.sync 4, 128;" ::: "memory");
        // copy-out: 16-byte pieces of the packed row into the quad kernel's A tiles
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
  if (warp == 0) tmem_dealloc(tmem_base, FAST ? kGTmemColsFast : kGTmemCols);
}

// ================================================================================================
// quad_kernel: var[item, dof] = sum_q A[item, q] V[dof, q]
constexpr int kQStages = 4;
constexpr int kQChunkGroups = 8;                        // 8 groups of 8 = 64 q per stage
constexpr uint32_t kQABytes = kQChunkGroups * 128 * 16; // 16 KB
constexpr uint32_t kQBBytes = kQChunkGroups * 256 * 16; // 32 KB
constexpr uint32_t kQStageBytes = kQABytes + kQBBytes;
constexpr int kQThreads = 192;                          // 4 epilogue warps, loader, MMA
constexpr int kQTmemCols = 512;

struct QuadParams {
  const __half* A;        // [n_ib][nkgQ][128][8]
  const __half* V;        // [n_db][nkgQ][256][8]
  const float* fscale;    // [n_items]
  float* field;           // [n_items][field_ld]
  unsigned int* maxvar_bits;   // [P]
  long long n_items;
  int T, D, field_ld, nkgQ, n_ib, n_db;
};

struct QBars {
  uint64_t full[kQStages], empty[kQStages], acc_full[2], acc_empty[2];
};

__global__ void __launch_bounds__(kQThreads, 1) quad_kernel(const __grid_constant__ QuadParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  QBars* B = reinterpret_cast<QBars*>(smem + kQStages * kQStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kQStages * kQStageBytes + sizeof(QBars));
  const int nchunks = P.nkgQ / kQChunkGroups;
  const long long n_tiles = (long long)P.n_ib * P.n_db;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kQStages; ++i) {
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
    tmem_alloc(tmem_slot, kQTmemCols);
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
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long ib = tile / P.n_db;
        const int db = (int)(tile - ib * P.n_db);
        const uint8_t* a0 = reinterpret_cast<const uint8_t*>(P.A) + (size_t)ib * P.nkgQ * 128 * 16;
        const uint8_t* b0 = reinterpret_cast<const uint8_t*>(P.V) + (size_t)db * P.nkgQ * 256 * 16;
        for (int kc = 0; kc < nchunks; ++kc) {
          const uint32_t s = it % kQStages;
          mbar_wait_relaxed(&B->empty[s], ((it / kQStages) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&B->full[s], kQStageBytes);
          uint8_t* dst = smem + s * kQStageBytes;
          bulk_g2s(dst, a0 + (size_t)kc * kQABytes, kQABytes, &B->full[s]);
          bulk_g2s(dst + kQABytes, b0 + (size_t)kc * kQBBytes, kQBBytes, &B->full[s]);
          ++it;
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---- MMA issuer -------------------------------------------------------------------------------
    uint32_t it = 0, a_it = 0;
    const uint32_t idesc = umma_idesc_f16_m128(256u);
    const uint64_t adesc0 = umma_desc_kmajor(smem_u32(smem), 128u * 16u, 128u);
    const uint64_t bdesc0 = umma_desc_kmajor(smem_u32(smem + kQABytes), 256u * 16u, 128u);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const uint32_t ab = a_it & 1u;
      mbar_wait(&B->acc_empty[ab], ((a_it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * 256u;
      for (int kc = 0; kc < nchunks; ++kc) {
        const uint32_t s = it % kQStages;
        mbar_wait(&B->full[s], (it / kQStages) & 1u);
        tc_fence_after();
        const uint64_t a0 = adesc0 + (uint64_t)(s * (kQStageBytes >> 4));
        const uint64_t b0 = bdesc0 + (uint64_t)(s * (kQStageBytes >> 4));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kQChunkGroups / 2; ++kk)
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
  } else {
    // ---- epilogue: thread <-> item row; un-scale, write the field, per-parameter maximum -----------------
    uint32_t a_it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long ib = tile / P.n_db;
      const int db = (int)(tile - ib * P.n_db);
      const long long item = ib * 128 + warp * 32 + lane;
      const bool valid = item < P.n_items;
      const float fs = valid ? __ldg(P.fscale + item) : 0.f;
      const uint32_t ab = a_it & 1u;
      mbar_wait_relaxed(&B->acc_full[ab], (a_it >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (((uint32_t)(warp * 32)) << 16) + ab * 256u;
      float* frow = P.field + (size_t)(valid ? item : 0) * P.field_ld + db * 256;
      float vmax = 0.f;
      uint32_t ra[32], rb[32];
      auto emit = [&](const uint32_t (&r)[32], int c0) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float4 o;
          o.x = __uint_as_float(r[q]) * fs;
          o.y = __uint_as_float(r[q + 1]) * fs;
          o.z = __uint_as_float(r[q + 2]) * fs;
          o.w = __uint_as_float(r[q + 3]) * fs;
          vmax = fmaxf(vmax, fmaxf(fmaxf(o.x, o.y), fmaxf(o.z, o.w)));
          if (valid) *reinterpret_cast<float4*>(frow + c0 + q) = o;
        }
      };
      tmem_ld_32x32(taddr, ra);
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 64) {
        tmem_ld_wait();
        tmem_ld_32x32(taddr + c0 + 32, rb);
        emit(ra, c0);
        tmem_ld_wait();
        if (c0 + 64 < 256) tmem_ld_32x32(taddr + c0 + 64, ra);
        emit(rb, c0 + 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B->acc_empty[ab]);
      {
        // one atomic per warp when its 32 items belong to one parameter (T >= 32: at most two per warp)
        const int p = valid ? (int)(item / P.T) : -1;
        const int p0 = __shfl_sync(0xffffffffu, p, 0);
        if (__all_sync(0xffffffffu, p == p0)) {
          const float wm = warp_max(vmax);
          if (lane == 0 && p0 >= 0 && wm > 0.f) atomicMax(P.maxvar_bits + p0, __float_as_uint(wm));
        } else if (valid && vmax > 0.f) {
          atomicMax(P.maxvar_bits + p, __float_as_uint(vmax));
        }
      }
      ++a_it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kQTmemCols);
}

}  // namespace gr

// ---- host side ------------------------------------------------------------------------------------
int gram_nkgQ(int K) {
  const int Q = K * (K + 1) / 2;
  const int g = (Q + 7) / 8;
  return (g + gr::kQChunkGroups - 1) / gr::kQChunkGroups * gr::kQChunkGroups;
}
size_t gram_a_bytes(int K, long long n_items) {
  const long long n_ib = (n_items + 127) / 128;
  return (size_t)n_ib * gram_nkgQ(K) * 128 * 16;
}
size_t gram_fscale_bytes(long long n_items) { return (size_t)((n_items + 127) / 128 * 128) * sizeof(float); }

// V[dof, q(i,j)] = (2 - delta_ij) w_i w_j 2^eV in fp16, laid out [dof block of 256][nkgQ][256][8]
int gram_pack_v(glasdi_handle* h, const float* WL /*[D][K] host*/) {
  Decoder& d = h->dec;
  const int K = d.K, D = d.D;
  const int nkgQ = gram_nkgQ(K);
  const int n_db = (D + 255) / 256;
  float wmax = 0.f;
  for (size_t i = 0; i < (size_t)D * K; ++i) wmax = std::max(wmax, std::fabs(WL[i]));
  d.eV = (wmax > 0.f) ? 13 - std::ilogb(2.0f * wmax * wmax) : 0;
  d.eV = std::max(-100, std::min(100, d.eV));
  std::vector<__half> packed((size_t)n_db * nkgQ * 256 * 8, __float2half_rn(0.f));
  for (int dof = 0; dof < D; ++dof) {
    const float* w = WL + (size_t)dof * K;
    const int db = dof / 256, row = dof % 256;
    __half* base = packed.data() + (size_t)db * nkgQ * 256 * 8;
    int q = 0;
    for (int i = 0; i < K; ++i)
      for (int j = 0; j <= i; ++j, ++q) {                      // same order as the epilogue of gram_kernel
        const float v = std::ldexp((i == j ? 1.0f : 2.0f) * w[i] * w[j], d.eV);
        base[((size_t)(q >> 3) * 256 + row) * 8 + (q & 7)] = __float2half_rn(v);
      }
  }
  if (d.v_packed) cudaFree(d.v_packed);
  d.v_packed = nullptr;
  GL_CUDA(h, cudaMalloc(&d.v_packed, packed.size() * sizeof(__half)));
  GL_CUDA(h, cudaMemcpy(d.v_packed, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
  // omega[dof] = |w_dof o w_dof|_2, the per-dof factor of the a-priori screening bound (refine.cu)
  std::vector<float> w4((size_t)(D + 255) / 256 * 256, 0.f);
  for (int dof = 0; dof < D; ++dof) {
    double a = 0.0;
    for (int i = 0; i < K; ++i) { const double w2 = (double)WL[(size_t)dof * K + i] * WL[(size_t)dof * K + i]; a += w2 * w2; }
    w4[dof] = (float)std::sqrt(a);
  }
  if (d.w4norm) cudaFree(d.w4norm);
  d.w4norm = nullptr;
  GL_CUDA(h, cudaMalloc(&d.w4norm, w4.size() * sizeof(float)));
  GL_CUDA(h, cudaMemcpy(d.w4norm, w4.data(), w4.size() * sizeof(float), cudaMemcpyHostToDevice));
  return GLASDI_OK;
}

// cnorm[item] = |diag C[item]|_2 in true units, from the packed covariances gram_kernel wrote: entry q(i, i) = i (i + 3) / 2
// of the item's row, times fscale[item] 2^eV (the scale quad_kernel's accumulator carries on top of the covariance's).
// Thread = item: a warp reads 32 neighbouring 16-byte groups per diagonal entry.
__global__ void gram_diag_norm_kernel(const __half* __restrict__ A, const float* __restrict__ fscale, long long n_items,
                                      int K, int nkgQ, float scale_v, float* __restrict__ cnorm) {
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  const __half* base = A + ((size_t)(item >> 7) * nkgQ * 128 + (size_t)(item & 127)) * 8;
  float acc = 0.f;
  for (int i = 0; i < K; ++i) {
    const int q = (i * (i + 3)) >> 1;
    const float v = __half2float(base[(size_t)(q >> 3) * 128 * 8 + (q & 7)]);
    acc = fmaf(v, v, acc);
  }
  cnorm[item] = sqrtf(acc) * fscale[item] * scale_v;
}

int launch_gram_diag_norm(glasdi_handle* h, const void* a_ws, const float* fscale_ws, long long n_items, float* cnorm,
                          cudaStream_t st) {
  const Decoder& d = h->dec;
  const unsigned int blocks = (unsigned int)((n_items + 255) / 256);
  gram_diag_norm_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __half*>(a_ws), fscale_ws, n_items, d.K,
                                                gram_nkgQ(d.K), ldexpf(1.0f, d.eV), cnorm);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "gram_diag_norm_kernel launch");
}

namespace gr { struct GramParams; }
int launch_decode_gram2(glasdi_handle* h, gr::GramParams& G, int ctas, cudaStream_t st);

static bool unbounded_act_g(int act) {
  return !(act == GLASDI_ACT_SIGMOID || act == GLASDI_ACT_TANH || act == GLASDI_ACT_HARDSIGMOID || act == GLASDI_ACT_HARDTANH);
}

template <int ACT, int INW, bool FAST>
static int launch_gram_inst(glasdi_handle* h, const gr::GramParams& G, int ctas, size_t smem, cudaStream_t st) {
  GL_CUDA(h, cudaFuncSetAttribute(gr::gram_kernel<ACT, INW, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gr::gram_kernel<ACT, INW, FAST><<<ctas, gr::kGThreads, smem, st>>>(G);
  return GLASDI_OK;
}

// first GEMM of the screening pass in covariance form: packed, scaled fp16 covariances of every (p, t) -> a_ws
int launch_decode_gram(glasdi_handle* h, const float* Zs, int Spad, int P_, int S, int T, void* a_ws, float* fscale_ws,
                       cudaStream_t st) {
  const Decoder& d = h->dec;
  if (!d.ready || !d.gram_ok) return set_error(h, GLASDI_ERR_INVALID, "decoder shape not supported by the Gram kernels");
  gr::GramParams G{};
  G.P = P_; G.S = S; G.T = T; G.Spad = Spad; G.NZ = d.widths[0];
  G.K = d.K;
  G.nkg = (d.K + 1 + 7) / 8;
  G.Nmma = (d.K + 1 + 15) / 16 * 16;
  G.nkgQ = gram_nkgQ(d.K);
  G.Q = d.K * (d.K + 1) / 2;
  G.h_scale = ldexpf(1.0f, d.eH);
  G.inv_S = 1.0f / (float)S;
  G.exp_base = 2 * d.eH + d.eV;
  G.check_overflow = unbounded_act_g(d.act) ? 1 : 0;
  G.n_items = (long long)P_ * T;
  G.Zs = Zs;
  G.A = reinterpret_cast<__half*>(a_ws);
  G.fscale = fscale_ws;
  G.flags = h->d_flags;
  G.front.n_front = d.n_linear - 1;
  G.front.act = d.act;
  for (int l = 0; l <= G.front.n_front; ++l) G.front.widths[l] = d.widths[l];
  for (int l = 0; l < G.front.n_front; ++l) { G.front.W[l] = d.dW[l]; G.front.b[l] = d.db[l]; }
  const int last_in_w = d.widths[d.n_linear - 2];
  const bool fast = G.front.n_front == 1;            // one front layer: pre-activations on the tensor core
  G.xw = 0;
  if (!fast)
    for (int l = 1; l < G.front.n_front; ++l) G.xw = std::max(G.xw, (d.widths[l] + 3) / 4 * 4);
  const size_t smem = gr::gcarve(G.nkg, last_in_w, G.nkgQ, G.xw).total;
  int ctas = h->num_sms;
  if (h->opt_max_ctas > 0) ctas = std::max(1, std::min(ctas, h->opt_max_ctas));
  if ((long long)ctas > G.n_items) ctas = (int)G.n_items;
  if (fast && h->opt_gram_kernel == 1 && gram2_supported(d)) return launch_decode_gram2(h, G, ctas, st);
  int rc;
  if (fast && d.act == GLASDI_ACT_SIGMOID && G.NZ == 5)
    rc = launch_gram_inst<GLASDI_ACT_SIGMOID, 5, true>(h, G, ctas, smem, st);     // the shipped Burgers decoder
  else if (fast && d.act == GLASDI_ACT_SIGMOID)
    rc = launch_gram_inst<GLASDI_ACT_SIGMOID, 0, true>(h, G, ctas, smem, st);
  else if (fast && d.act == GLASDI_ACT_SOFTPLUS)
    rc = launch_gram_inst<GLASDI_ACT_SOFTPLUS, 0, true>(h, G, ctas, smem, st);
  else if (fast)
    rc = launch_gram_inst<-1, 0, true>(h, G, ctas, smem, st);
  else if (d.act == GLASDI_ACT_SOFTPLUS)
    rc = launch_gram_inst<GLASDI_ACT_SOFTPLUS, 0, false>(h, G, ctas, smem, st);
  else
    rc = launch_gram_inst<-1, 0, false>(h, G, ctas, smem, st);
  if (rc) return rc;
  h->launches++;
  if ((rc = check_cuda(h, cudaGetLastError(), "gram_kernel launch"))) return rc;

  return GLASDI_OK;
}

// second GEMM of the screening pass: var[(p,t), dof] = Ctilde[(p,t), :] . V[dof, :], field + per-parameter maximum
int launch_decode_quad(glasdi_handle* h, int P_, int T, unsigned int* maxvar_bits, float* var_field, const void* a_ws,
                       const float* fscale_ws, cudaStream_t st) {
  const Decoder& d = h->dec;
  gr::QuadParams Q{};
  Q.A = reinterpret_cast<const __half*>(a_ws); Q.V = d.v_packed; Q.fscale = fscale_ws; Q.field = var_field;
  Q.maxvar_bits = maxvar_bits;
  Q.n_items = (long long)P_ * T; Q.T = T; Q.D = d.D; Q.field_ld = decode_field_ld(d.n_tiles); Q.nkgQ = gram_nkgQ(d.K);
  Q.n_ib = (int)((Q.n_items + 127) / 128);
  Q.n_db = (d.D + 255) / 256;
  const size_t qsmem = gr::kQStages * gr::kQStageBytes + sizeof(gr::QBars) + 16;
  GL_CUDA(h, cudaFuncSetAttribute(gr::quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem));
  int qctas = h->num_sms;
  if (h->opt_max_ctas > 0) qctas = std::max(1, std::min(qctas, h->opt_max_ctas));
  const long long n_tiles = (long long)Q.n_ib * Q.n_db;
  if ((long long)qctas > n_tiles) qctas = (int)n_tiles;
  gr::quad_kernel<<<qctas, gr::kQThreads, qsmem, st>>>(Q);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "quad_kernel launch");
}

size_t decode_gram_smem_bytes(int K, int last_in_w, int xw) { return gr::gcarve((K + 1 + 7) / 8, last_in_w, gram_nkgQ(K), xw).total; }

}  // namespace glasdi
