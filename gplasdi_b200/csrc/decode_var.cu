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
//                         -> variance -> running max -> warp-shuffle max -> atomicMax(maxvar[p])
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
#include "common.cuh"

namespace glasdi {
namespace dv {

constexpr int kTileM = 128;       // DOFs per MMA tile (TMEM lanes)
constexpr int kChunkN = 128;      // columns (samples or time-step x sample) per chunk
constexpr int kNtMax = 8;         // DOF tiles per group (register accumulators per epilogue thread)
constexpr int kWSlots = 3;        // ring of W part-tiles (hi or lo) in shared memory
constexpr int kMaxSeg = 16;       // max time steps packed in one chunk when S < kChunkN
constexpr int kMaxHid = 64;       // max width of intermediate front layers
constexpr int kEpiWarps = 4;
constexpr int kLoaderWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kProdWarp0 = 6;
constexpr int kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = (kProdWarp0 + kProdWarps) * 32;   // 448
constexpr int kTmemCols = 256;    // 2 accumulators x 128 fp32 columns
constexpr uint32_t kLBO = kChunkN * 16;   // bytes between the two K halves (k-groups of 8 halfs)
constexpr uint32_t kSBO = 128;            // bytes between 8-row groups

enum Mode { kVar = 0, kStore = 1 };

struct Params {
  int mode;
  // VAR mode geometry
  int P, S, T, Spad;
  int TPC, n_tg, n_chunks;
  float inv_S;
  const float* Zs;                  // [P][T][NZ][Spad]
  unsigned int* maxvar_bits;        // [P]
  // STORE mode
  const float* Zrows;               // [rows][NZ]
  long long rows;
  float* X;                         // [rows][D]
  const float* b_last;              // [D]
  float out_scale;                  // 2^-(eW+eH)
  // common
  int NZ, D, K, Kpad, n_tiles, n_groups, passes;
  long long n_items;
  float h_scale;                    // 2^eH
  const __half* w_packed;
  int* flags;
  FrontMlp front;
};

struct Smem {
  // offsets into dynamic shared memory
  uint32_t w, h, wl, pilot, bars, tmem;
  uint32_t total;
};

__host__ __device__ inline Smem carve(int Kpad, int last_in_w) {
  Smem s;
  const uint32_t part = (uint32_t)Kpad * 256u;          // one 128-row x Kpad fp16 operand tile
  uint32_t off = 0;
  s.w = off; off += kWSlots * part;
  s.h = off; off += 2 * 2 * part;                      // 2 buffers x (hi, lo)
  s.wl = off; off += (uint32_t)Kpad * (uint32_t)(last_in_w + 1) * 4u;
  off = (off + 15u) & ~15u;
  s.pilot = off; off += kMaxSeg * (uint32_t)Kpad * 4u;
  off = (off + 15u) & ~15u;
  s.bars = off; off += 16 * 8;
  s.tmem = off; off += 16;
  s.total = off;
  return s;
}

// ---- front MLP -------------------------------------------------------------------------------
// layers 0 .. n_front-2 (small), fp32, weights via the read-only path; result = input of the last
// front layer.  For a single front layer this is the identity on z.
__device__ __forceinline__ int front_pre(const FrontMlp& f, const float* zin, float* out) {
  int w = f.widths[0];
  for (int i = 0; i < w; ++i) out[i] = zin[i];
  float tmp[kMaxHid];
  for (int l = 0; l + 1 < f.n_front; ++l) {
    const int wo = f.widths[l + 1];
    const float* __restrict__ W = f.W[l];
    const float* __restrict__ b = f.b[l];
    for (int o = 0; o < wo; ++o) {
      float acc = __ldg(b + o);
      for (int i = 0; i < w; ++i) acc = fmaf(__ldg(W + (size_t)o * w + i), out[i], acc);
      tmp[o] = act_apply(f.act, acc);
    }
    for (int o = 0; o < wo; ++o) out[o] = tmp[o];
    w = wo;
  }
  return w;
}

// 8 consecutive units of the last front layer from shared-memory weights wl[k][inw+1] (bias last)
template <int INW>
__device__ __forceinline__ void last_front_units(const float* __restrict__ wl, int inw_rt, int act, int K, int k0,
                                                 const float* x, float (&hv)[8]) {
  const int inw = INW > 0 ? INW : inw_rt;
  const int ld = inw + 1;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int k = k0 + u;
    float v = 0.f;
    if (k < K) {
      const float* wr = wl + k * ld;
      float acc = wr[inw];
      if (INW > 0) {
#pragma unroll
        for (int i = 0; i < (INW > 0 ? INW : 1); ++i) acc = fmaf(wr[i], x[i], acc);
      } else {
        for (int i = 0; i < inw; ++i) acc = fmaf(wr[i], x[i], acc);
      }
      v = act_apply(act, acc);
    }
    hv[u] = v;
  }
}

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// Work decomposition shared by every role so that all of them walk the identical sequence.
struct Item {
  int p, tg, grp, tile0, nt;
};
__device__ __forceinline__ Item decode_item(const Params& P, long long item) {
  Item it;
  it.grp = (int)(item % P.n_groups);
  long long r = item / P.n_groups;
  if (P.mode == kVar) {
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
__device__ __forceinline__ int chunk_cols(const Params& P, const Item& it, int c) {
  if (P.mode == kVar) {
    if (P.TPC == 1) return min(kChunkN, P.S - c * kChunkN);
    const int nseg = min(P.TPC, P.T - it.tg * P.TPC);
    return nseg * P.S;
  }
  const long long r0 = (long long)it.p * kChunkN;
  return (int)min((long long)kChunkN, P.rows - r0);
}
__device__ __forceinline__ int mma_n(int cols) { return max(32, (cols + 31) & ~31); }

template <int INW>
__global__ void __launch_bounds__(kThreads, 1) decode_kernel(const Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int last_in_w = P.front.widths[P.front.n_front - 1];
  const Smem L = carve(P.Kpad, last_in_w);
  const uint32_t part_bytes = (uint32_t)P.Kpad * 256u;

  uint8_t* sW = smem + L.w;
  uint8_t* sH = smem + L.h;
  float* sWl = reinterpret_cast<float*>(smem + L.wl);
  float* sPilot = reinterpret_cast<float*>(smem + L.pilot);
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
  // last front layer weights (+bias) -> shared memory, [k][inw+1]
  {
    const int l = P.front.n_front - 1;
    const float* __restrict__ W = P.front.W[l];
    const float* __restrict__ b = P.front.b[l];
    const int ld = last_in_w + 1;
    for (int idx = threadIdx.x; idx < P.K * ld; idx += blockDim.x) {
      const int k = idx / ld, i = idx - k * ld;
      sWl[idx] = (i < last_in_w) ? W[(size_t)k * last_in_w + i] : b[k];
    }
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
      for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode_item(P, item);
        const int nchunks = (P.mode == kVar) ? P.n_chunks : 1;
        for (int c = 0; c < nchunks; ++c) {
          for (int m = 0; m < it.nt; ++m) {
            const int nparts = (P.passes >= 2) ? 2 : 1;
            for (int part = 0; part < nparts; ++part) {
              const uint32_t slot = w_it % kWSlots;
              mbar_wait(&w_empty[slot], ((w_it / kWSlots) & 1u) ^ 1u);
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
    // ---- MMA issuer: one thread -----------------------------------------------------------------
    if (lane == 0) {
      uint32_t w_it = 0, h_it = 0, a_it = 0;
      const int nk16 = P.Kpad / 16;
      for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = decode_item(P, item);
        const int nchunks = (P.mode == kVar) ? P.n_chunks : 1;
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t idesc = umma_idesc_f16_m128((uint32_t)mma_n(chunk_cols(P, it, c)));
          const uint32_t hb = h_it & 1u;
          mbar_wait(&h_full[hb], (h_it >> 1) & 1u);
          tc_fence_after();
          const uint32_t hhi = smem_u32(sH + (hb * 2 + 0) * part_bytes);
          const uint32_t hlo = smem_u32(sH + (hb * 2 + 1) * part_bytes);
          for (int m = 0; m < it.nt; ++m) {
            const uint32_t ab = a_it & 1u;
            mbar_wait(&acc_empty[ab], ((a_it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + ab * (uint32_t)kChunkN;
            // W hi part
            uint32_t slot = w_it % kWSlots;
            mbar_wait(&w_full[slot], (w_it / kWSlots) & 1u);
            tc_fence_after();
            uint32_t wa = smem_u32(sW + slot * part_bytes);
            for (int k = 0; k < nk16; ++k)
              tc_mma_f16(d_tmem, umma_desc_kmajor(wa + k * 2 * kLBO, kLBO, kSBO),
                         umma_desc_kmajor(hhi + k * 2 * kLBO, kLBO, kSBO), idesc, k > 0 ? 1u : 0u);
            if (P.passes >= 3)
              for (int k = 0; k < nk16; ++k)
                tc_mma_f16(d_tmem, umma_desc_kmajor(wa + k * 2 * kLBO, kLBO, kSBO),
                           umma_desc_kmajor(hlo + k * 2 * kLBO, kLBO, kSBO), idesc, 1u);
            tc_commit(&w_empty[slot]);
            ++w_it;
            if (P.passes >= 2) {
              slot = w_it % kWSlots;
              mbar_wait(&w_full[slot], (w_it / kWSlots) & 1u);
              tc_fence_after();
              wa = smem_u32(sW + slot * part_bytes);
              for (int k = 0; k < nk16; ++k)
                tc_mma_f16(d_tmem, umma_desc_kmajor(wa + k * 2 * kLBO, kLBO, kSBO),
                           umma_desc_kmajor(hhi + k * 2 * kLBO, kLBO, kSBO), idesc, 1u);
              tc_commit(&w_empty[slot]);
              ++w_it;
            }
            tc_commit(&acc_full[ab]);
            ++a_it;
          }
          tc_commit(&h_empty[hb]);
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
    uint32_t h_it = 0;
    bool overflow = false;
    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const int nchunks = (P.mode == kVar) ? P.n_chunks : 1;
      for (int c = 0; c < nchunks; ++c) {
        const int ncols = chunk_cols(P, it, c);
        const uint32_t hb = h_it & 1u;
        uint8_t* Hhi = sH + (hb * 2 + 0) * part_bytes;
        uint8_t* Hlo = sH + (hb * 2 + 1) * part_bytes;

        // which (time step, sample) / row does my column hold?
        int seg = 0;
        const float* zsrc = nullptr;
        size_t zstride = 0;
        const bool valid = col < ncols;
        int nseg = 1;
        if (P.mode == kVar) {
          int s;
          if (P.TPC == 1) {
            s = c * kChunkN + col;
          } else {
            seg = col / P.S;
            s = col - seg * P.S;
            nseg = min(P.TPC, P.T - it.tg * P.TPC);
          }
          const int t = it.tg * P.TPC + seg;
          if (valid) zsrc = P.Zs + ((size_t)it.p * P.T + t) * NZ * P.Spad + s;
          zstride = P.Spad;
        } else {
          const long long row = (long long)it.p * kChunkN + col;
          if (valid) zsrc = P.Zrows + (size_t)row * NZ;
          zstride = 1;
        }
        float zin[kMaxLatent];
#pragma unroll
        for (int i = 0; i < kMaxLatent; ++i) zin[i] = (valid && i < NZ) ? __ldg(zsrc + i * zstride) : 0.f;

        // ---- pilot activations: first sample of every time step in this chunk (VAR mode) -------
        if (P.mode == kVar) {
          if (c == 0 || P.TPC > 1) {
            // (segment, k-group) tasks
            for (int task = ptid; task < nseg * nkg; task += kProdThreads) {
              const int sg = task / nkg, kg = task - sg * nkg;
              const int t = it.tg * P.TPC + sg;
              const float* zp = P.Zs + ((size_t)it.p * P.T + t) * NZ * P.Spad;   // sample 0
              float zz[kMaxLatent];
#pragma unroll
              for (int i = 0; i < kMaxLatent; ++i) zz[i] = (i < NZ) ? __ldg(zp + (size_t)i * P.Spad) : 0.f;
              float hv[8];
              if (INW > 0) {
                last_front_units<INW>(sWl, last_in_w, P.front.act, P.K, kg * 8, zz, hv);
              } else {
                float x[kMaxHid];
                front_pre(P.front, zz, x);
                last_front_units<0>(sWl, last_in_w, P.front.act, P.K, kg * 8, x, hv);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) sPilot[sg * P.Kpad + kg * 8 + u] = hv[u];
            }
            // the pilot of the previous chunk/item may still be read by slow producer threads:
            // both sides of the exchange are fenced by the two named barriers below.
            asm volatile("bar.sync 1, %0;" ::"n"(kProdThreads) : "memory");
          }
        }

        // ---- wait until the MMA warp has drained this H buffer --------------------------------
        mbar_wait(&h_empty[hb], ((h_it >> 1) & 1u) ^ 1u);

        float xloc[INW > 0 ? 1 : kMaxHid];
        const float* x = zin;
        if (INW == 0) {
          if (valid) front_pre(P.front, zin, xloc);
          x = xloc;
        }
        const float* pil = sPilot + seg * P.Kpad;
        for (int kg = kg_begin; kg < kg_end; ++kg) {
          uint4 vhi = make_uint4(0, 0, 0, 0), vlo = make_uint4(0, 0, 0, 0);
          if (valid) {
            float hv[8];
            last_front_units<INW>(sWl, last_in_w, P.front.act, P.K, kg * 8, x, hv);
            __half hh[8], hl[8];
            float amax = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float d = hv[u];
              if (P.mode == kVar) d -= pil[kg * 8 + u];
              const float v = d * P.h_scale;
              amax = fmaxf(amax, fabsf(v));
              hh[u] = __float2half_rn(v);
              hl[u] = __float2half_rn(v - __half2float(hh[u]));
            }
            overflow |= !(amax < 60000.f);
            vhi = make_uint4(pack_half2(hh[0], hh[1]), pack_half2(hh[2], hh[3]), pack_half2(hh[4], hh[5]),
                             pack_half2(hh[6], hh[7]));
            vlo = make_uint4(pack_half2(hl[0], hl[1]), pack_half2(hl[2], hl[3]), pack_half2(hl[4], hl[5]),
                             pack_half2(hl[6], hl[7]));
          }
          const uint32_t off = ((uint32_t)kg * kChunkN + (uint32_t)col) * 16u;
          *reinterpret_cast<uint4*>(Hhi + off) = vhi;
          *reinterpret_cast<uint4*>(Hlo + off) = vlo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_full[hb]);
        ++h_it;
        if (P.mode == kVar && (c == nchunks - 1 || P.TPC > 1)) {
          // nobody may overwrite the pilot before every producer finished reading it
          asm volatile("bar.sync 2, %0;" ::"n"(kProdThreads) : "memory");
        }
      }
    }
    if (overflow) atomicOr(P.flags, 1);
  } else {
    // ---- epilogue warps 0..3 : lane quadrant = warp -----------------------------------------------
    uint32_t a_it = 0;
    const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const int nchunks = (P.mode == kVar) ? P.n_chunks : 1;
      float accs[kNtMax], accq[kNtMax];
#pragma unroll
      for (int m = 0; m < kNtMax; ++m) { accs[m] = 0.f; accq[m] = 0.f; }
      float vmax = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        const int ncols = chunk_cols(P, it, c);
        const int nmma = mma_n(ncols);
#pragma unroll
        for (int m = 0; m < kNtMax; ++m) {
          if (m < it.nt) {
            const uint32_t ab = a_it & 1u;
            mbar_wait(&acc_full[ab], (a_it >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_addr + ab * (uint32_t)kChunkN;
            if (P.mode == kVar) {
              if (P.TPC == 1) {
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                for (int c0 = 0; c0 < nmma; c0 += 32) {
                  uint32_t r[32];
                  tmem_ld_32x32(taddr + c0, r);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 32; j += 2) {
                    const float d0 = __uint_as_float(r[j]), d1 = __uint_as_float(r[j + 1]);
                    s0 += d0; s1 += d1;
                    q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
                  }
                }
                accs[m] += s0 + s1;
                accq[m] += q0 + q1;
              } else {
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
              }
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
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
            ++a_it;
          }
        }
      }
      if (P.mode == kVar) {
        if (P.TPC == 1) {
#pragma unroll
          for (int m = 0; m < kNtMax; ++m) {
            if (m < it.nt) {
              const float mean = accs[m] * P.inv_S;
              vmax = fmaxf(vmax, fmaf(-mean, mean, accq[m] * P.inv_S));
            }
          }
        }
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

static int fill_common(glasdi_handle* h, dv::Params& P) {
  const Decoder& d = h->dec;
  if (!d.ready) return set_error(h, GLASDI_ERR_NOT_READY, "decoder weights not set (glasdi_set_decoder)");
  if (!d.tc_ok) return set_error(h, GLASDI_ERR_INVALID, "decoder shape not supported by the tcgen05 kernel");
  P.NZ = d.widths[0];
  P.D = d.D; P.K = d.K; P.Kpad = d.Kpad; P.n_tiles = d.n_tiles;
  P.n_groups = (d.n_tiles + dv::kNtMax - 1) / dv::kNtMax;
  P.passes = h->opt_passes;
  P.h_scale = ldexpf(1.0f, d.eH);
  P.w_packed = d.w_packed;
  P.flags = h->d_flags;
  P.front.n_front = d.n_linear - 1;
  P.front.act = d.act;
  for (int l = 0; l <= P.front.n_front; ++l) P.front.widths[l] = d.widths[l];
  for (int l = 0; l < P.front.n_front; ++l) { P.front.W[l] = d.dW[l]; P.front.b[l] = d.db[l]; }
  return GLASDI_OK;
}

static int launch(glasdi_handle* h, const dv::Params& P, cudaStream_t st) {
  const int last_in_w = P.front.widths[P.front.n_front - 1];
  const size_t smem = decode_var_smem_bytes(P.Kpad, last_in_w);
  int grid = h->num_sms;
  if (h->opt_max_ctas > 0) grid = std::min(grid, h->opt_max_ctas);
  if ((long long)grid > P.n_items) grid = (int)P.n_items;
  if (grid <= 0) return GLASDI_OK;
  const bool fast5 = (P.front.n_front == 1 && P.NZ == 5);
  if (fast5) {
    GL_CUDA(h, cudaFuncSetAttribute(dv::decode_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dv::decode_kernel<5><<<grid, dv::kThreads, smem, st>>>(P);
  } else {
    GL_CUDA(h, cudaFuncSetAttribute(dv::decode_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dv::decode_kernel<0><<<grid, dv::kThreads, smem, st>>>(P);
  }
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "decode_kernel launch");
}

int launch_decode_var(glasdi_handle* h, const float* Zs, int Spad, int P_, int S, int T, unsigned int* maxvar_bits,
                      cudaStream_t st) {
  dv::Params P{};
  int rc = fill_common(h, P);
  if (rc) return rc;
  P.mode = dv::kVar;
  P.P = P_; P.S = S; P.T = T; P.Spad = Spad;
  P.Zs = Zs;
  P.maxvar_bits = maxvar_bits;
  P.inv_S = 1.0f / (float)S;
  if (S >= dv::kChunkN) {
    P.TPC = 1;
    P.n_chunks = (S + dv::kChunkN - 1) / dv::kChunkN;
  } else {
    P.TPC = std::min(dv::kMaxSeg, dv::kChunkN / S);
    P.n_chunks = 1;
  }
  P.n_tg = (T + P.TPC - 1) / P.TPC;
  P.n_items = (long long)P_ * P.n_tg * P.n_groups;
  return launch(h, P, st);
}

int launch_decode_store(glasdi_handle* h, const float* Z, int64_t rows, float* X, cudaStream_t st) {
  dv::Params P{};
  int rc = fill_common(h, P);
  if (rc) return rc;
  const Decoder& d = h->dec;
  P.mode = dv::kStore;
  P.Zrows = Z; P.rows = rows; P.X = X;
  P.b_last = d.db[d.n_linear - 1];
  P.out_scale = ldexpf(1.0f, -(d.eW + d.eH));
  P.TPC = 1; P.n_tg = 1; P.n_chunks = 1; P.S = 1; P.T = 1; P.Spad = 1; P.inv_S = 1.f;
  P.n_items = ((rows + dv::kChunkN - 1) / dv::kChunkN) * P.n_groups;
  return launch(h, P, st);
}

}  // namespace glasdi
