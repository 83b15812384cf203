// decode_var2.cu -- fused decode + variance kernel of the decoded-ensemble form (GLASDI_MODE_DIRECT / GLASDI_MODE_SCREENED,
// glasdi_decode, glasdi_decode_stat_fields), on CTA pairs (cta_group::2).
//
// Warp-specialised: producers evaluate the front MLP and write the fp16 hi/lo split of the pilot-centred hidden activations
// as the B operand, a loader streams the pre-packed last-layer tiles with bulk copies, one thread issues the MMAs, the
// epilogue forms sum d / sum d^2 per dof from TMEM.  Two CTAs on the two SMs of a
// TPC cooperate on every MMA: M = 256 (each CTA owns one 128-DOF tile: its A operand and its TMEM
// accumulator rows) x N = 256 columns (each CTA produces and holds 128 of them: half of the B
// operand), K = 16.  Per SM and per MMA this halves the shared-memory operand reads (64 B/clk
// instead of the 128 B/clk = 100 % of the SMEM pipe that M=128 x N=128 needs) and the W-tile bytes
// streamed per FLOP, which is what bounded the one-CTA-per-SM kernel of round 1 (retired; profiles/r01_*).
//
// Cross-CTA protocol (rank 0 = leader issues every tcgen05.mma for the pair):
//   h_full[2]    leader   count 16  <- producer warps of BOTH CTAs (remote mbarrier arrive, release.cluster)
//   w_full[3]    each CTA count 1+tx <- its own bulk-async W copies
//   w_peer[3]    leader   count 1   <- relay thread of the peer CTA once ITS copy of the slot landed
//   acc_empty[2] leader   count 8   <- epilogue warps of BOTH CTAs
//   h_empty[2], w_empty[3], acc_full[2]  each CTA, count 1 <- tcgen05.commit multicast to both CTAs
#include "decode_common.cuh"

namespace glasdi {
namespace dv2 {

using namespace dv;

constexpr int kPairN = 256;          // columns per chunk for the pair (128 per CTA)
constexpr int kTmemCols2 = 512;      // 2 accumulators x 256 fp32 columns
constexpr int kPairTiles = 2 * kNtMax;   // DOF tiles per item group (8 per CTA)

struct Bars {
  uint64_t h_full[2], h_empty[2], w_full[kWSlots], w_peer[kWSlots], w_empty[kWSlots], acc_full[2], acc_empty[2];
};

struct Item {
  int p, tg, grp, tile0, nsteps;
};
template <int EPI>
__device__ __forceinline__ Item decode_item(const Params& P, long long item) {
  Item it;
  it.grp = (int)(item % P.n_groups);
  long long r = item / P.n_groups;
  if (EPI != kStore) {
    it.tg = (int)(r % P.n_tg);
    it.p = (int)(r / P.n_tg);
  } else {
    it.tg = 0;
    it.p = (int)r;
  }
  it.tile0 = it.grp * kPairTiles;
  it.nsteps = (min(kPairTiles, P.n_tiles - it.tile0) + 1) / 2;
  return it;
}
template <int EPI>
__device__ __forceinline__ int chunk_cols(const Params& P, const Item& it, int c) {
  if (EPI == kVarT1) return min(kPairN, P.S - c * kPairN);
  if (EPI == kVarSeg) return min(P.TPC, P.T - it.tg * P.TPC) * P.S;
  const long long r0 = (long long)it.p * kPairN;
  return (int)min((long long)kPairN, P.rows - r0);
}
__device__ __forceinline__ int cols32(int cols) { return max(32, (cols + 31) & ~31); }

template <int EPI, int ACT, int INW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
decode_pair_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const long long pair = cluster_id_x();
  const long long n_pairs = cluster_nctaid_x();
  const int last_in_w = P.front.widths[P.front.n_front - 1];
  const Smem L = carve(P.Kpad, last_in_w);
  const uint32_t part_bytes = (uint32_t)P.Kpad * 256u;
  const int nchunks = (EPI == kVarT1) ? P.n_chunks : 1;

  uint8_t* sW = smem + L.w;
  uint8_t* sH = smem + L.h;
  float* sWl = reinterpret_cast<float*>(smem + L.wl);
  float* sPilot = reinterpret_cast<float*>(smem + L.pilot);
  float* sAcc = reinterpret_cast<float*>(smem + L.acc);
  Bars* B = reinterpret_cast<Bars*>(smem + L.bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&B->h_full[i], 2 * kProdWarps);
      mbar_init(&B->h_empty[i], 1);
      mbar_init(&B->acc_full[i], 1);
      mbar_init(&B->acc_empty[i], 2 * kEpiWarps);
    }
    for (int i = 0; i < kWSlots; ++i) {
      mbar_init(&B->w_full[i], 1);
      mbar_init(&B->w_peer[i], 1);
      mbar_init(&B->w_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_2cta(tmem_slot, kTmemCols2);
    tmem_relinquish_2cta();
  }
  {
    const int l = P.front.n_front - 1;
    const float* __restrict__ W = P.front.W[l];
    const float* __restrict__ b = P.front.b[l];
    const int ld = L.wl_ld;
    for (int idx = threadIdx.x; idx < P.Kpad * ld; idx += blockDim.x) {
      const int k = idx / ld, i = idx - k * ld;
      float v = 0.f;
      if (k < P.K) {
        if (i < last_in_w) v = W[(size_t)k * last_in_w + i];
        else if (i == last_in_w) v = b[k];
      }
      sWl[idx] = v;
    }
    for (int idx = threadIdx.x; idx < kMaxSeg * P.Kpad; idx += blockDim.x) sPilot[idx] = 0.f;
  }
  tc_fence_before();
  cluster_sync_all();            // barriers of both CTAs initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // addresses of the leader's barriers in the cluster window (valid from either CTA)
  const uint32_t ldr_h_full0 = mapa_u32(smem_u32(&B->h_full[0]), 0);
  const uint32_t ldr_acc_empty0 = mapa_u32(smem_u32(&B->acc_empty[0]), 0);
  const uint32_t ldr_w_peer0 = mapa_u32(smem_u32(&B->w_peer[0]), 0);
  const int nparts = (P.passes >= 2) ? 2 : 1;

  // =============================================================================================
  if (warp == kLoaderWarp) {
    // ---- W loader (both CTAs): this CTA's DOF tile of every pair step ---------------------------
    if (lane == 0) {
      uint32_t w_it = 0;
      for (long long item = pair; item < P.n_items; item += n_pairs) {
        const Item it = decode_item<EPI>(P, item);
        for (int c = 0; c < nchunks; ++c) {
          for (int j = 0; j < it.nsteps; ++j) {
            const int tile = it.tile0 + 2 * j + (int)rank;        // < n_tiles rounded up to even (zero padded)
            for (int part = 0; part < nparts; ++part) {
              const uint32_t slot = w_it % kWSlots;
              mbar_wait_relaxed(&B->w_empty[slot], ((w_it / kWSlots) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&B->w_full[slot], part_bytes);
              const uint8_t* src = reinterpret_cast<const uint8_t*>(P.w_packed) + ((size_t)tile * 2 + part) * part_bytes;
              bulk_g2s(sW + slot * part_bytes, src, part_bytes, &B->w_full[slot]);
              ++w_it;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    if (leader) {
      // ---- MMA issuer for the pair.  The WHOLE warp walks the loop (warp-uniform control flow, so the
      // descriptors stay in uniform registers); one elected lane issues each tcgen05.mma / commit. ----
      uint32_t w_it = 0, h_it = 0, a_it = 0;
      const int nk16 = P.Kpad / 16;
      const uint32_t idesc = umma_idesc_f16_m256((uint32_t)kPairN);
      constexpr uint32_t kDescStep = (2 * kLBO) >> 4;      // start-address units (16 B) per K=16 slice
      const uint64_t wdesc0 = umma_desc_kmajor(smem_u32(sW), kLBO, kSBO);
      const uint32_t wslot_step = part_bytes >> 4;
      const uint64_t hdesc0 = umma_desc_kmajor(smem_u32(sH), kLBO, kSBO);
      for (long long item = pair; item < P.n_items; item += n_pairs) {
        const Item it = decode_item<EPI>(P, item);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t hb = h_it & 1u;
          mbar_wait_cluster(&B->h_full[hb], (h_it >> 1) & 1u);
          tc_fence_after();
          const uint64_t bhi0 = hdesc0 + (uint64_t)((hb * 2 + 0) * wslot_step);
          const uint64_t blo0 = hdesc0 + (uint64_t)((hb * 2 + 1) * wslot_step);
          for (int j = 0; j < it.nsteps; ++j) {
            const uint32_t ab = a_it & 1u;
            mbar_wait_cluster(&B->acc_empty[ab], ((a_it >> 1) & 1u) ^ 1u);
            const uint32_t d_tmem = tmem_base + ab * (uint32_t)kPairN;
            uint32_t slot = w_it % kWSlots;
            mbar_wait(&B->w_full[slot], (w_it / kWSlots) & 1u);
            mbar_wait_cluster(&B->w_peer[slot], (w_it / kWSlots) & 1u);
            tc_fence_after();
            uint64_t a0 = wdesc0 + (uint64_t)(slot * wslot_step);
            if (elect_one()) {
#pragma unroll 1
              for (int k = 0; k < nk16; ++k)
                tc_mma_f16_2cta(d_tmem, a0 + (uint64_t)(k * kDescStep), bhi0 + (uint64_t)(k * kDescStep), idesc,
                                k > 0 ? 1u : 0u);
              if (P.passes >= 3) {
#pragma unroll 1
                for (int k = 0; k < nk16; ++k)
                  tc_mma_f16_2cta(d_tmem, a0 + (uint64_t)(k * kDescStep), blo0 + (uint64_t)(k * kDescStep), idesc, 1u);
              }
              tc_commit_2cta(&B->w_empty[slot], 3);
            }
            __syncwarp();
            ++w_it;
            if (P.passes >= 2) {
              slot = w_it % kWSlots;
              mbar_wait(&B->w_full[slot], (w_it / kWSlots) & 1u);
              mbar_wait_cluster(&B->w_peer[slot], (w_it / kWSlots) & 1u);
              tc_fence_after();
              a0 = wdesc0 + (uint64_t)(slot * wslot_step);
              if (elect_one()) {
#pragma unroll 1
                for (int k = 0; k < nk16; ++k)
                  tc_mma_f16_2cta(d_tmem, a0 + (uint64_t)(k * kDescStep), bhi0 + (uint64_t)(k * kDescStep), idesc, 1u);
                tc_commit_2cta(&B->w_empty[slot], 3);
              }
              __syncwarp();
              ++w_it;
            }
            if (elect_one()) tc_commit_2cta(&B->acc_full[ab], 3);
            __syncwarp();
            ++a_it;
          }
          if (elect_one()) tc_commit_2cta(&B->h_empty[hb], 3);
          __syncwarp();
          ++h_it;
        }
      }
    } else if (lane == 0) {
      // ---- relay (peer CTA): tell the leader when THIS CTA's copy of each W slot has landed ----
      uint32_t w_it = 0;
      for (long long item = pair; item < P.n_items; item += n_pairs) {
        const Item it = decode_item<EPI>(P, item);
        const int total = nchunks * it.nsteps * nparts;
        for (int q = 0; q < total; ++q) {
          const uint32_t slot = w_it % kWSlots;
          mbar_wait_relaxed(&B->w_full[slot], (w_it / kWSlots) & 1u);
          mbar_arrive_cluster(ldr_w_peer0 + slot * 8u);
          ++w_it;
        }
      }
    }
    __syncwarp();
  } else if (warp >= kProdWarp0) {
    // ---- producers: this CTA's 128 columns of every 256-column chunk -----------------------------
    const int ptid = threadIdx.x - kProdWarp0 * 32;
    const int col = ptid & 127;
    const int gcol = (int)rank * 128 + col;            // column inside the pair's chunk
    const int half = ptid >> 7;
    const int nkg = P.Kpad / 8;
    const int kg_begin = half == 0 ? 0 : (nkg + 1) / 2;
    const int kg_end = half == 0 ? (nkg + 1) / 2 : nkg;
    const int NZ = P.NZ;
    const int ld = L.wl_ld;
    const float hscale = P.h_scale;
    uint32_t h_it = 0;
    float amax = 0.f;

    auto load_z = [&](const Item& it, int c, float (&zin)[kMaxLatent], int& seg, bool& valid) {
      const int ncols = chunk_cols<EPI>(P, it, c);
      valid = gcol < ncols;
      seg = 0;
      const float* zsrc = nullptr;
      size_t zstride = 0;
      if (EPI != kStore) {
        int s;
        if (EPI == kVarT1) {
          s = c * kPairN + gcol;
        } else {
          seg = gcol / P.S;
          s = gcol - seg * P.S;
        }
        const int t = it.tg * P.TPC + seg;
        if (valid) zsrc = P.Zs + ((size_t)it.p * P.T + t) * NZ * P.Spad + s;
        zstride = P.Spad;
      } else {
        const long long row = (long long)it.p * kPairN + gcol;
        if (valid) zsrc = P.Zrows + (size_t)row * NZ;
        zstride = 1;
      }
#pragma unroll
      for (int i = 0; i < kMaxLatent; ++i) zin[i] = (valid && i < NZ) ? __ldg(zsrc + i * zstride) : 0.f;
      if (!valid) seg = 0;
    };

    for (long long item = pair; item < P.n_items; item += n_pairs) {
      const Item it = decode_item<EPI>(P, item);
      float zin[kMaxLatent];
      int seg;
      bool valid;
      load_z(it, 0, zin, seg, valid);
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t hb = h_it & 1u;
        uint8_t* Hhi = sH + (hb * 2 + 0) * part_bytes;
        uint8_t* Hlo = sH + (hb * 2 + 1) * part_bytes;

        if (EPI != kStore && c == 0) {
          const int nseg = (EPI == kVarT1) ? 1 : min(P.TPC, P.T - it.tg * P.TPC);
          for (int task = ptid; task < nseg * nkg; task += kProdThreads) {
            const int sg = task / nkg, kg = task - sg * nkg;
            const int t = it.tg * P.TPC + sg;
            const float* zp = P.Zs + ((size_t)it.p * P.T + t) * NZ * P.Spad;   // sample 0
            float zz[kMaxLatent];
#pragma unroll
            for (int i = 0; i < kMaxLatent; ++i) zz[i] = (i < NZ) ? __ldg(zp + (size_t)i * P.Spad) : 0.f;
            float hv[8];
            if constexpr (INW > 0) {
              last_front_units<ACT, INW>(sWl, ld, last_in_w, P.front.act, kg * 8, zz, hv);
            } else {
              float x[kMaxHid];
              front_pre<ACT>(P.front, zz, x);
              last_front_units<ACT, 0>(sWl, ld, last_in_w, P.front.act, kg * 8, x, hv);
            }
            float4* dst = reinterpret_cast<float4*>(sPilot + sg * P.Kpad + kg * 8);
            dst[0] = make_float4(hv[0] * hscale, hv[1] * hscale, hv[2] * hscale, hv[3] * hscale);
            dst[1] = make_float4(hv[4] * hscale, hv[5] * hscale, hv[6] * hscale, hv[7] * hscale);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads) : "memory");
        }

        float xloc[INW > 0 ? 1 : kMaxHid];
        const float* x = zin;
        if constexpr (INW == 0) {
          if (valid) front_pre<ACT>(P.front, zin, xloc);
          x = xloc;
        }
        float znext[kMaxLatent];
        int seg_next = 0;
        bool valid_next = false;
        if (c + 1 < nchunks) load_z(it, c + 1, znext, seg_next, valid_next);

        mbar_wait_relaxed(&B->h_empty[hb], ((h_it >> 1) & 1u) ^ 1u);

        const float* pil = sPilot + seg * P.Kpad;
        for (int kg = kg_begin; kg < kg_end; ++kg) {
          uint4 vhi = make_uint4(0, 0, 0, 0), vlo = make_uint4(0, 0, 0, 0);
          if (valid) {
            float hv[8];
            last_front_units<ACT, INW>(sWl, ld, last_in_w, P.front.act, kg * 8, x, hv);
            const float4 p0 = *reinterpret_cast<const float4*>(pil + kg * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(pil + kg * 8 + 4);
            const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            uint32_t uh[4], ul[4];
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
              const float v0 = fmaf(hv[u], hscale, -pv[u]);
              const float v1 = fmaf(hv[u + 1], hscale, -pv[u + 1]);
              if (P.check_overflow) amax = fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1)));
              const __half2 hh = __floats2half2_rn(v0, v1);
              uh[u >> 1] = h2_as_u32(hh);
              if (nparts == 2) {
                const float2 bk = __half22float2(hh);
                ul[u >> 1] = h2_as_u32(__floats2half2_rn(v0 - bk.x, v1 - bk.y));
              } else {
                ul[u >> 1] = 0u;
              }
            }
            vhi = make_uint4(uh[0], uh[1], uh[2], uh[3]);
            vlo = make_uint4(ul[0], ul[1], ul[2], ul[3]);
          }
          const uint32_t off = ((uint32_t)kg * 128u + (uint32_t)col) * 16u;
          *reinterpret_cast<uint4*>(Hhi + off) = vhi;
          if (nparts == 2) *reinterpret_cast<uint4*>(Hlo + off) = vlo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ldr_h_full0 + hb * 8u);
        ++h_it;
        if (EPI != kStore && c == nchunks - 1) {
          asm volatile("bar.sync 2, %0;" ::"n"(kProdThreads) : "memory");
        }
#pragma unroll
        for (int i = 0; i < kMaxLatent; ++i) zin[i] = znext[i];
        seg = seg_next;
        valid = valid_next;
      }
    }
    if (P.check_overflow && !(amax < 60000.f)) atomicOr(P.flags, 1);
  } else {
    // ---- epilogue warps 0..3 : this CTA's DOF tile (128 TMEM lanes) x 256 columns ------------------
    uint32_t a_it = 0;
    const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
    float* myAcc = sAcc + threadIdx.x;
    for (long long item = pair; item < P.n_items; item += n_pairs) {
      const Item it = decode_item<EPI>(P, item);
      float vmax = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        const int ncols = chunk_cols<EPI>(P, it, c);
        const int nread = cols32(ncols);                 // columns beyond are exact zeros: skip them
        for (int j = 0; j < it.nsteps; ++j) {
          const uint32_t ab = a_it & 1u;
          mbar_wait_relaxed(&B->acc_full[ab], (a_it >> 1) & 1u);
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + ab * (uint32_t)kPairN;
          if constexpr (EPI == kVarT1) {
            unsigned long long s2a = 0ull, s2b = 0ull, q2a = 0ull, q2b = 0ull;
            auto accumulate = [&](const uint32_t (&r)[32]) {
#pragma unroll
              for (int q = 0; q < 32; q += 4) {
                unsigned long long d0, d1;
                asm("mov.b64 %0, {%1, %2};" : "=l"(d0) : "r"(r[q]), "r"(r[q + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(d1) : "r"(r[q + 2]), "r"(r[q + 3]));
                s2a = add_f32x2(s2a, d0);
                s2b = add_f32x2(s2b, d1);
                q2a = fma_f32x2(d0, d0, q2a);
                q2b = fma_f32x2(d1, d1, q2b);
              }
            };
            // two register buffers: the TMEM load of the next 32 columns is in flight while the sums of the
            // current 32 are formed (the un-pipelined loop stalled on every LDTM, profiles/: decode_r1_screen)
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(taddr, ra);
            for (int c0 = 0; c0 < nread; c0 += 64) {
              tmem_ld_wait();
              const bool more = c0 + 32 < nread;
              if (more) tmem_ld_32x32(taddr + c0 + 32, rb);
              accumulate(ra);
              if (more) {
                tmem_ld_wait();
                if (c0 + 64 < nread) tmem_ld_32x32(taddr + c0 + 64, ra);
                accumulate(rb);
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(ldr_acc_empty0 + ab * 8u);
            float sa, sb, qa, qb;
            unpack_f32x2(add_f32x2(s2a, s2b), sa, sb);
            unpack_f32x2(add_f32x2(q2a, q2b), qa, qb);
            float ssum = sa + sb, qsum = qa + qb;
            if (c > 0) {
              ssum += myAcc[(j * 2 + 0) * 128];
              qsum += myAcc[(j * 2 + 1) * 128];
            }
            if (c + 1 < nchunks) {
              myAcc[(j * 2 + 0) * 128] = ssum;
              myAcc[(j * 2 + 1) * 128] = qsum;
            } else {
              const float mean = ssum * P.inv_S;
              const float var = fmaf(-mean, mean, qsum * P.inv_S);
              vmax = fmaxf(vmax, var);
              if (P.var_field) {
                const size_t fo = ((size_t)it.p * P.T + it.tg) * P.field_ld + (it.tile0 + 2 * j + (int)rank) * kTileM +
                                  warp * 32 + lane;
                P.var_field[fo] = var;
                if (P.mean_field) P.mean_field[fo] = mean;
              }
            }
          } else if constexpr (EPI == kVarSeg) {
            float ss = 0.f, qq = 0.f;
            int cnt = 0, sg = 0;
            for (int c0 = 0; c0 < nread; c0 += 32) {
              uint32_t r[32];
              tmem_ld_32x32(taddr + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                const float d = __uint_as_float(r[q]);
                ss += d;
                qq = fmaf(d, d, qq);
                if (++cnt == P.S) {
                  const float mean = ss * P.inv_S;
                  const float var = fmaf(-mean, mean, qq * P.inv_S);
                  vmax = fmaxf(vmax, var);
                  if (P.var_field && it.tg * P.TPC + sg < P.T) {    // zero-padded columns complete phantom segments
                    const size_t fo = ((size_t)it.p * P.T + it.tg * P.TPC + sg) * P.field_ld +
                                      (it.tile0 + 2 * j + (int)rank) * kTileM + warp * 32 + lane;
                    P.var_field[fo] = var;
                    if (P.mean_field) P.mean_field[fo] = mean;
                  }
                  ss = 0.f; qq = 0.f; cnt = 0; ++sg;
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(ldr_acc_empty0 + ab * 8u);
          } else {
            const int dof = (it.tile0 + 2 * j + (int)rank) * kTileM + warp * 32 + lane;
            const float bias = dof < P.D ? __ldg(P.b_last + dof) : 0.f;
            const long long row0 = (long long)it.p * kPairN;
            for (int c0 = 0; c0 < nread; c0 += 32) {
              uint32_t r[32];
              tmem_ld_32x32(taddr + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 32; ++q) {
                const long long row = row0 + c0 + q;
                if (dof < P.D && c0 + q < ncols)
                  P.X[(size_t)row * P.D + dof] = fmaf(__uint_as_float(r[q]), P.out_scale, bias);
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(ldr_acc_empty0 + ab * 8u);
          }
          ++a_it;
        }
      }
      if constexpr (EPI != kStore) {
        vmax = warp_max(fmaxf(vmax, 0.f));
        if (lane == 0 && vmax > 0.f) atomicMax(P.maxvar_bits + it.p, __float_as_uint(vmax));
      }
    }
  }

  // ---- teardown: nobody leaves while the peer may still touch its shared memory / barriers --------
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc_2cta(tmem_base, kTmemCols2);
}

}  // namespace dv2

size_t decode_var_smem_bytes(int Kpad, int last_in_w) { return dv::carve(Kpad, last_in_w).total; }

static bool unbounded_act2(int act) {
  return !(act == GLASDI_ACT_SIGMOID || act == GLASDI_ACT_TANH || act == GLASDI_ACT_HARDSIGMOID || act == GLASDI_ACT_HARDTANH);
}

static int fill_common2(glasdi_handle* h, dv::Params& P) {
  const Decoder& d = h->dec;
  if (!d.ready) return set_error(h, GLASDI_ERR_NOT_READY, "decoder weights not set (glasdi_set_decoder)");
  if (!d.tc_ok) return set_error(h, GLASDI_ERR_INVALID, "decoder shape not supported by the tcgen05 kernel");
  P.NZ = d.widths[0];
  P.D = d.D; P.K = d.K; P.Kpad = d.Kpad; P.n_tiles = d.n_tiles;
  P.n_groups = (d.n_tiles + dv2::kPairTiles - 1) / dv2::kPairTiles;
  P.passes = h->opt_passes;
  P.h_scale = ldexpf(1.0f, d.eH);
  P.check_overflow = unbounded_act2(d.act) ? 1 : 0;
  P.w_packed = d.w_packed;
  P.flags = h->d_flags;
  P.front.n_front = d.n_linear - 1;
  P.front.act = d.act;
  for (int l = 0; l <= P.front.n_front; ++l) P.front.widths[l] = d.widths[l];
  for (int l = 0; l < P.front.n_front; ++l) { P.front.W[l] = d.dW[l]; P.front.b[l] = d.db[l]; }
  return GLASDI_OK;
}

template <int EPI, int ACT, int INW>
static int launch_inst2(glasdi_handle* h, const dv::Params& P, int pairs, size_t smem, cudaStream_t st) {
  GL_CUDA(h, cudaFuncSetAttribute(dv2::decode_pair_kernel<EPI, ACT, INW>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dv2::decode_pair_kernel<EPI, ACT, INW><<<2 * pairs, dv::kThreads, smem, st>>>(P);
  return GLASDI_OK;
}

template <int EPI>
static int launch_epi2(glasdi_handle* h, const dv::Params& P, cudaStream_t st) {
  const int last_in_w = P.front.widths[P.front.n_front - 1];
  const size_t smem = decode_var_smem_bytes(P.Kpad, last_in_w);
  int pairs = h->num_sms / 2;
  if (h->opt_max_ctas > 0) pairs = std::max(1, std::min(pairs, h->opt_max_ctas / 2));
  if ((long long)pairs > P.n_items) pairs = (int)P.n_items;
  if (pairs <= 0) return GLASDI_OK;
  int rc;
  if (P.front.n_front == 1 && P.NZ == 5 && P.front.act == GLASDI_ACT_SIGMOID)
    rc = launch_inst2<EPI, GLASDI_ACT_SIGMOID, 5>(h, P, pairs, smem, st);
  else if (P.front.act == GLASDI_ACT_SOFTPLUS)
    rc = launch_inst2<EPI, GLASDI_ACT_SOFTPLUS, 0>(h, P, pairs, smem, st);
  else
    rc = launch_inst2<EPI, -1, 0>(h, P, pairs, smem, st);
  if (rc) return rc;
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "decode_pair_kernel launch");
}

int launch_decode_var_pair(glasdi_handle* h, const float* Zs, int Spad, int P_, int S, int T,
                           unsigned int* maxvar_bits, float* var_field, int passes, cudaStream_t st,
                           float* mean_field) {
  dv::Params P{};
  int rc = fill_common2(h, P);
  if (rc) return rc;
  P.P = P_; P.S = S; P.T = T; P.Spad = Spad;
  P.Zs = Zs;
  P.maxvar_bits = maxvar_bits;
  P.var_field = var_field;
  P.mean_field = mean_field;
  P.field_ld = decode_field_ld(h->dec.n_tiles);
  if (passes > 0) P.passes = passes;
  P.inv_S = 1.0f / (float)S;
  const bool seg = S < dv2::kPairN / 2;          // at least two time steps fit one 256-column chunk
  if (!seg) {
    P.TPC = 1;
    P.n_chunks = (S + dv2::kPairN - 1) / dv2::kPairN;
  } else {
    P.TPC = std::min(dv::kMaxSeg, dv2::kPairN / S);
    P.n_chunks = 1;
  }
  P.n_tg = (T + P.TPC - 1) / P.TPC;
  P.n_items = (long long)P_ * P.n_tg * P.n_groups;
  return seg ? launch_epi2<dv::kVarSeg>(h, P, st) : launch_epi2<dv::kVarT1>(h, P, st);
}

int launch_decode_store_pair(glasdi_handle* h, const float* Z, int64_t rows, float* X, cudaStream_t st) {
  dv::Params P{};
  int rc = fill_common2(h, P);
  if (rc) return rc;
  const Decoder& d = h->dec;
  P.Zrows = Z; P.rows = rows; P.X = X;
  P.b_last = d.db[d.n_linear - 1];
  P.out_scale = ldexpf(1.0f, -(d.eW + d.eH));
  P.TPC = 1; P.n_tg = 1; P.n_chunks = 1; P.S = 1; P.T = 1; P.Spad = 1; P.inv_S = 1.f;
  P.n_items = ((rows + dv2::kPairN - 1) / dv2::kPairN) * P.n_groups;
  return launch_epi2<dv::kStore>(h, P, st);
}

}  // namespace glasdi
