// decode_var.cu -- fused MLP decode + sample-variance + max kernel on tcgen05 tensor cores.
//
// Replaces get_fom_max_std (reference src/lasdi/gplasdi.py:52-74): per test parameter p,
//   X[s,t,:] = decoder(Z[p,s,t,:])   (MultiLayerPerceptron.forward, src/lasdi/latent_space.py:157-164)
//   max_std[p] = max_{t,d} std_s X[s,t,d]      (numpy std, ddof = 0)
// The decoded ensemble (S x T x D per parameter) is never materialised: the last (large) linear
// layer runs on the 5th-gen tensor cores with DOFs on the MMA M axis (TMEM lanes) and samples on
// the N axis, so every epilogue thread owns one DOF row and reduces over samples in registers.
//
// Persistent warp-specialised kernel, one CTA per SM, 14 warps:
//   warps 0-3  epilogue : tcgen05.ld accumulator -> per-DOF sum / sum-of-squares over samples
//                         (packed f32x2 FADD2/FFMA2) -> variance -> running max -> warp-shuffle max
//                         -> atomicMax(maxvar[p])
//   warp  4    loader   : bulk-async (TMA engine) copies of pre-packed fp16 W tiles, 3-slot ring
//   warp  5    MMA      : one thread issues tcgen05.mma (M=128, N<=128, K=16, kind::f16)
//   warps 6-13 producers: front MLP layers on CUDA cores from the rollout's Zs, subtract the
//                         per-time-step pilot activation, split to fp16 hi/lo and write the B
//                         operand straight into UMMA's canonical K-major layout in shared memory
//
// fp32-faithful products on fp16 tensor cores: W = (Whi + Wlo) 2^-eW and delta = (Hhi + Hlo) 2^-eH
// with hi = fp16(x), lo = fp16(x - hi); passes Whi*Hhi + Whi*Hlo + Wlo*Hhi accumulate in the same
// fp32 TMEM accumulator (dropped term Wlo*Hlo ~ 2^-22).  `passes` = 3 (default), 2 (no Whi*Hlo)
// or 1 (plain fp16).  Variance is shift invariant, so the producers feed
// delta = h(s) - h(pilot sample) and the GEMM yields x(s) - x(pilot) directly: no bias, no
// cancellation in sum-of-squares, and no extra epilogue work.
//
// Template parameters: EPI (epilogue/work decomposition), ACT (compile-time activation, -1 = runtime
// switch), INW (compile-time input width of the last front layer for the single-hidden-layer
// decoders of the Burgers configs, 0 = generic).
#include "decode_common.cuh"

namespace glasdi {
namespace dv {

// Work decomposition shared by every role so that all of them walk the identical sequence.
struct Item {
  int p, tg, grp, tile0, nt;
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
    it.p = (int)r;   // row-chunk index
  }
  it.tile0 = it.grp * kNtMax;
  it.nt = min(kNtMax, P.n_tiles - it.tile0);
  return it;
}
// number of valid columns of chunk `c` of an item
template <int EPI>
__device__ __forceinline__ int chunk_cols(const Params& P, const Item& it, int c) {
  if (EPI == kVarT1) return min(kChunkN, P.S - c * kChunkN);
  if (EPI == kVarSeg) return min(P.TPC, P.T - it.tg * P.TPC) * P.S;
  const long long r0 = (long long)it.p * kChunkN;
  return (int)min((long long)kChunkN, P.rows - r0);
}
__device__ __forceinline__ int mma_n(int cols) { return max(32, (cols + 31) & ~31); }

template <int EPI, int ACT, int INW>
__global__ void __launch_bounds__(kThreads, 1) decode_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int last_in_w = P.front.widths[P.front.n_front - 1];
  const Smem L = carve(P.Kpad, last_in_w);
  const uint32_t part_bytes = (uint32_t)P.Kpad * 256u;
  const int nchunks = (EPI == kVarT1) ? P.n_chunks : 1;

  uint8_t* sW = smem + L.w;
  uint8_t* sH = smem + L.h;
  float* sWl = reinterpret_cast<float*>(smem + L.wl);
  float* sPilot = reinterpret_cast<float*>(smem + L.pilot);
  float* sAcc = reinterpret_cast<float*>(smem + L.acc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* h_full = bars + 0;     // [2] producers -> MMA
  uint64_t* h_empty = bars + 2;    // [2] MMA -> producers
  uint64_t* w_full = bars + 4;     // [3] loader -> MMA
  uint64_t* w_empty = bars + 7;    // [3] MMA -> loader
  uint64_t* acc_full = bars + 10;  // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 12; // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);

  // ---- one-time setup -------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], kProdWarps);
      mbar_init(&h_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    for (int i = 0; i < kWSlots; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  // last front layer weights (+bias) -> shared memory rows of L.wl_ld floats; rows >= K are zero
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // =============================================================================================
  if (warp == kLoaderWarp) {
    // ---- W loader: one thread, bulk async copies into the 3-slot ring --------------------------
    if (lane == 0) {
      uint32_t w_it = 0;
      const int nparts = (P.passes >= 2) ? 2 : 1;
      for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode_item<EPI>(P, item);
        for (int c = 0; c < nchunks; ++c) {
          for (int m = 0; m < it.nt; ++m) {
            for (int part = 0; part < nparts; ++part) {
              const uint32_t slot = w_it % kWSlots;
              mbar_wait_relaxed(&w_empty[slot], ((w_it / kWSlots) & 1u) ^ 1u);
              mbar_arrive_expect_tx(&w_full[slot], part_bytes);
              const uint8_t* src = reinterpret_cast<const uint8_t*>(P.w_packed) +
                                   ((size_t)(it.tile0 + m) * 2 + part) * part_bytes;
              bulk_g2s(sW + slot * part_bytes, src, part_bytes, &w_full[slot]);
              ++w_it;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ---- MMA issuer: the whole warp walks the loop (warp-uniform control flow keeps the descriptors in
    // uniform registers); one elected lane issues each tcgen05.mma / commit -------------------------
    {
      uint32_t w_it = 0, h_it = 0, a_it = 0;
      const int nk16 = P.Kpad / 16;
      constexpr uint32_t kDescStep = (2 * kLBO) >> 4;
      const uint64_t wdesc0 = umma_desc_kmajor(smem_u32(sW), kLBO, kSBO);
      const uint64_t hdesc0 = umma_desc_kmajor(smem_u32(sH), kLBO, kSBO);
      const uint32_t slot_step = part_bytes >> 4;
      for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode_item<EPI>(P, item);
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t idesc = umma_idesc_f16_m128((uint32_t)mma_n(chunk_cols<EPI>(P, it, c)));
          const uint32_t hb = h_it & 1u;
          mbar_wait(&h_full[hb], (h_it >> 1) & 1u);
          tc_fence_after();
          const uint64_t bhi0 = hdesc0 + (uint64_t)((hb * 2 + 0) * slot_step);
          const uint64_t blo0 = hdesc0 + (uint64_t)((hb * 2 + 1) * slot_step);
          for (int m = 0; m < it.nt; ++m) {
            const uint32_t ab = a_it & 1u;
            mbar_wait(&acc_empty[ab], ((a_it >> 1) & 1u) ^ 1u);
            const uint32_t d_tmem = tmem_base + ab * (uint32_t)kChunkN;
            uint32_t slot = w_it % kWSlots;
            mbar_wait(&w_full[slot], (w_it / kWSlots) & 1u);
            tc_fence_after();
            uint64_t a0 = wdesc0 + (uint64_t)(slot * slot_step);
            if (elect_one()) {
#pragma unroll 1
              for (int k = 0; k < nk16; ++k)
                tc_mma_f16(d_tmem, a0 + (uint64_t)(k * kDescStep), bhi0 + (uint64_t)(k * kDescStep), idesc,
                           k > 0 ? 1u : 0u);
              if (P.passes >= 3) {
#pragma unroll 1
                for (int k = 0; k < nk16; ++k)
                  tc_mma_f16(d_tmem, a0 + (uint64_t)(k * kDescStep), blo0 + (uint64_t)(k * kDescStep), idesc, 1u);
              }
              tc_commit(&w_empty[slot]);
            }
            __syncwarp();
            ++w_it;
            if (P.passes >= 2) {
              slot = w_it % kWSlots;
              mbar_wait(&w_full[slot], (w_it / kWSlots) & 1u);
              tc_fence_after();
              a0 = wdesc0 + (uint64_t)(slot * slot_step);
              if (elect_one()) {
#pragma unroll 1
                for (int k = 0; k < nk16; ++k)
                  tc_mma_f16(d_tmem, a0 + (uint64_t)(k * kDescStep), bhi0 + (uint64_t)(k * kDescStep), idesc, 1u);
                tc_commit(&w_empty[slot]);
              }
              __syncwarp();
              ++w_it;
            }
            if (elect_one()) tc_commit(&acc_full[ab]);
            __syncwarp();
            ++a_it;
          }
          if (elect_one()) tc_commit(&h_empty[hb]);
          __syncwarp();
          ++h_it;
        }
      }
    }
    __syncwarp();
  } else if (warp >= kProdWarp0) {
    // ---- producers: front MLP -> delta -> fp16 hi/lo -> UMMA B operand in shared memory ---------
    const int ptid = threadIdx.x - kProdWarp0 * 32;
    const int col = ptid & (kChunkN - 1);
    const int half = ptid >> 7;
    const int nkg = P.Kpad / 8;
    const int kg_begin = half == 0 ? 0 : (nkg + 1) / 2;
    const int kg_end = half == 0 ? (nkg + 1) / 2 : nkg;
    const int NZ = P.NZ;
    const int ld = L.wl_ld;
    const float hscale = P.h_scale;
    uint32_t h_it = 0;
    float amax = 0.f;

    // z of (item, chunk c) for my column; zero when the column is padding
    auto load_z = [&](const Item& it, int c, float (&zin)[kMaxLatent], int& seg, bool& valid) {
      const int ncols = chunk_cols<EPI>(P, it, c);
      valid = col < ncols;
      seg = 0;
      const float* zsrc = nullptr;
      size_t zstride = 0;
      if (EPI != kStore) {
        int s;
        if (EPI == kVarT1) {
          s = c * kChunkN + col;
        } else {
          seg = col / P.S;
          s = col - seg * P.S;
        }
        const int t = it.tg * P.TPC + seg;
        if (valid) zsrc = P.Zs + ((size_t)it.p * P.T + t) * NZ * P.Spad + s;
        zstride = P.Spad;
      } else {
        const long long row = (long long)it.p * kChunkN + col;
        if (valid) zsrc = P.Zrows + (size_t)row * NZ;
        zstride = 1;
      }
#pragma unroll
      for (int i = 0; i < kMaxLatent; ++i) zin[i] = (valid && i < NZ) ? __ldg(zsrc + i * zstride) : 0.f;
    };

    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode_item<EPI>(P, item);
      float zin[kMaxLatent];
      int seg;
      bool valid;
      load_z(it, 0, zin, seg, valid);
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t hb = h_it & 1u;
        uint8_t* Hhi = sH + (hb * 2 + 0) * part_bytes;
        uint8_t* Hlo = sH + (hb * 2 + 1) * part_bytes;

        // ---- pilot activations (pre-scaled by 2^eH): first sample of every time step -------------
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
        // prefetch the next chunk's z while this one is being produced
        float znext[kMaxLatent];
        int seg_next = 0;
        bool valid_next = false;
        if (c + 1 < nchunks) load_z(it, c + 1, znext, seg_next, valid_next);

        // ---- wait until the MMA warp has drained this H buffer --------------------------------
        mbar_wait_relaxed(&h_empty[hb], ((h_it >> 1) & 1u) ^ 1u);

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
              const float2 bk = __half22float2(hh);
              const __half2 hl = __floats2half2_rn(v0 - bk.x, v1 - bk.y);
              uh[u >> 1] = h2_as_u32(hh);
              ul[u >> 1] = h2_as_u32(hl);
            }
            vhi = make_uint4(uh[0], uh[1], uh[2], uh[3]);
            vlo = make_uint4(ul[0], ul[1], ul[2], ul[3]);
          }
          const uint32_t off = ((uint32_t)kg * kChunkN + (uint32_t)col) * 16u;
          *reinterpret_cast<uint4*>(Hhi + off) = vhi;
          *reinterpret_cast<uint4*>(Hlo + off) = vlo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_full[hb]);
        ++h_it;
        if (EPI != kStore && c == nchunks - 1) {
          // nobody may overwrite the pilot before every producer finished reading it
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
    // ---- epilogue warps 0..3 : lane quadrant = warp -----------------------------------------------
    uint32_t a_it = 0;
    const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
    float* myAcc = sAcc + threadIdx.x;          // [m][2][128] floats, thread-contiguous => conflict free
    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode_item<EPI>(P, item);
      float vmax = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        const int ncols = chunk_cols<EPI>(P, it, c);
        const int nmma = mma_n(ncols);
        for (int m = 0; m < it.nt; ++m) {
          const uint32_t ab = a_it & 1u;
          mbar_wait_relaxed(&acc_full[ab], (a_it >> 1) & 1u);
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + ab * (uint32_t)kChunkN;
          if constexpr (EPI == kVarT1) {
            unsigned long long s2a = 0ull, s2b = 0ull, q2a = 0ull, q2b = 0ull;   // packed (0.f, 0.f)
            for (int c0 = 0; c0 < nmma; c0 += 32) {
              uint32_t r[32];
              tmem_ld_32x32(taddr + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                unsigned long long d0, d1;
                asm("mov.b64 %0, {%1, %2};" : "=l"(d0) : "r"(r[j]), "r"(r[j + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(d1) : "r"(r[j + 2]), "r"(r[j + 3]));
                s2a = add_f32x2(s2a, d0);
                s2b = add_f32x2(s2b, d1);
                q2a = fma_f32x2(d0, d0, q2a);
                q2b = fma_f32x2(d1, d1, q2b);
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
            float sa, sb, qa, qb;
            unpack_f32x2(add_f32x2(s2a, s2b), sa, sb);
            unpack_f32x2(add_f32x2(q2a, q2b), qa, qb);
            float ssum = sa + sb, qsum = qa + qb;
            if (c > 0) {
              ssum += myAcc[(m * 2 + 0) * 128];
              qsum += myAcc[(m * 2 + 1) * 128];
            }
            if (c + 1 < nchunks) {
              myAcc[(m * 2 + 0) * 128] = ssum;
              myAcc[(m * 2 + 1) * 128] = qsum;
            } else {
              const float mean = ssum * P.inv_S;
              vmax = fmaxf(vmax, fmaf(-mean, mean, qsum * P.inv_S));
            }
          } else if constexpr (EPI == kVarSeg) {
            // several time steps in one chunk: segmented reduction, S columns per segment
            float ss = 0.f, qq = 0.f;
            int cnt = 0;
            for (int c0 = 0; c0 < nmma; c0 += 32) {
              uint32_t r[32];
              tmem_ld_32x32(taddr + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float d = __uint_as_float(r[j]);
                ss += d;
                qq = fmaf(d, d, qq);
                if (++cnt == P.S) {
                  const float mean = ss * P.inv_S;
                  vmax = fmaxf(vmax, fmaf(-mean, mean, qq * P.inv_S));
                  ss = 0.f; qq = 0.f; cnt = 0;
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
          } else {
            // STORE: X[row][dof] = acc * 2^-(eW+eH) + b[dof]
            const int dof = (it.tile0 + m) * kTileM + warp * 32 + lane;
            const float bias = dof < P.D ? __ldg(P.b_last + dof) : 0.f;
            const long long row0 = (long long)it.p * kChunkN;
            for (int c0 = 0; c0 < nmma; c0 += 32) {
              uint32_t r[32];
              tmem_ld_32x32(taddr + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const long long row = row0 + c0 + j;
                if (dof < P.D && c0 + j < ncols)
                  P.X[(size_t)row * P.D + dof] = fmaf(__uint_as_float(r[j]), P.out_scale, bias);
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
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

  // ---- teardown --------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace dv

size_t decode_var_smem_bytes(int Kpad, int last_in_w) { return dv::carve(Kpad, last_in_w).total; }

static bool unbounded_act(int act) {
  return !(act == GLASDI_ACT_SIGMOID || act == GLASDI_ACT_TANH || act == GLASDI_ACT_HARDSIGMOID || act == GLASDI_ACT_HARDTANH);
}

static int fill_common(glasdi_handle* h, dv::Params& P) {
  const Decoder& d = h->dec;
  if (!d.ready) return set_error(h, GLASDI_ERR_NOT_READY, "decoder weights not set (glasdi_set_decoder)");
  if (!d.tc_ok) return set_error(h, GLASDI_ERR_INVALID, "decoder shape not supported by the tcgen05 kernel");
  P.NZ = d.widths[0];
  P.D = d.D; P.K = d.K; P.Kpad = d.Kpad; P.n_tiles = d.n_tiles;
  P.n_groups = (d.n_tiles + dv::kNtMax - 1) / dv::kNtMax;
  P.passes = h->opt_passes;
  P.h_scale = ldexpf(1.0f, d.eH);
  P.check_overflow = unbounded_act(d.act) ? 1 : 0;
  P.w_packed = d.w_packed;
  P.flags = h->d_flags;
  P.front.n_front = d.n_linear - 1;
  P.front.act = d.act;
  for (int l = 0; l <= P.front.n_front; ++l) P.front.widths[l] = d.widths[l];
  for (int l = 0; l < P.front.n_front; ++l) { P.front.W[l] = d.dW[l]; P.front.b[l] = d.db[l]; }
  return GLASDI_OK;
}

template <int EPI, int ACT, int INW>
static int launch_inst(glasdi_handle* h, const dv::Params& P, int grid, size_t smem, cudaStream_t st) {
  GL_CUDA(h, cudaFuncSetAttribute(dv::decode_kernel<EPI, ACT, INW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  dv::decode_kernel<EPI, ACT, INW><<<grid, dv::kThreads, smem, st>>>(P);
  return GLASDI_OK;
}

template <int EPI>
static int launch_epi(glasdi_handle* h, const dv::Params& P, cudaStream_t st) {
  const int last_in_w = P.front.widths[P.front.n_front - 1];
  const size_t smem = decode_var_smem_bytes(P.Kpad, last_in_w);
  int grid = h->num_sms;
  if (h->opt_max_ctas > 0) grid = std::min(grid, h->opt_max_ctas);
  if ((long long)grid > P.n_items) grid = (int)P.n_items;
  if (grid <= 0) return GLASDI_OK;
  int rc;
  // compile-time specialisations for the decoders of the named configurations; the rest goes
  // through the runtime-activation / generic-width instantiation (same kernel, same numerics)
  if (P.front.n_front == 1 && P.NZ == 5 && P.front.act == GLASDI_ACT_SIGMOID)
    rc = launch_inst<EPI, GLASDI_ACT_SIGMOID, 5>(h, P, grid, smem, st);
  else if (P.front.act == GLASDI_ACT_SOFTPLUS)
    rc = launch_inst<EPI, GLASDI_ACT_SOFTPLUS, 0>(h, P, grid, smem, st);
  else
    rc = launch_inst<EPI, -1, 0>(h, P, grid, smem, st);
  if (rc) return rc;
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "decode_kernel launch");
}

int launch_decode_var(glasdi_handle* h, const float* Zs, int Spad, int P_, int S, int T, unsigned int* maxvar_bits,
                      cudaStream_t st) {
  dv::Params P{};
  int rc = fill_common(h, P);
  if (rc) return rc;
  P.P = P_; P.S = S; P.T = T; P.Spad = Spad;
  P.Zs = Zs;
  P.maxvar_bits = maxvar_bits;
  P.inv_S = 1.0f / (float)S;
  const bool seg = S < dv::kChunkN;
  if (!seg) {
    P.TPC = 1;
    P.n_chunks = (S + dv::kChunkN - 1) / dv::kChunkN;
  } else {
    P.TPC = std::min(dv::kMaxSeg, dv::kChunkN / S);
    P.n_chunks = 1;
  }
  P.n_tg = (T + P.TPC - 1) / P.TPC;
  P.n_items = (long long)P_ * P.n_tg * P.n_groups;
  return seg ? launch_epi<dv::kVarSeg>(h, P, st) : launch_epi<dv::kVarT1>(h, P, st);
}

int launch_decode_store(glasdi_handle* h, const float* Z, int64_t rows, float* X, cudaStream_t st) {
  dv::Params P{};
  int rc = fill_common(h, P);
  if (rc) return rc;
  const Decoder& d = h->dec;
  P.Zrows = Z; P.rows = rows; P.X = X;
  P.b_last = d.db[d.n_linear - 1];
  P.out_scale = ldexpf(1.0f, -(d.eW + d.eH));
  P.TPC = 1; P.n_tg = 1; P.n_chunks = 1; P.S = 1; P.T = 1; P.Spad = 1; P.inv_S = 1.f;
  P.n_items = ((rows + dv::kChunkN - 1) / dv::kChunkN) * P.n_groups;
  return launch_epi<dv::kStore>(h, P, st);
}

}  // namespace glasdi
