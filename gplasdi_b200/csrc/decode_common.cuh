// decode_common.cuh -- pieces shared by the tcgen05 decode kernels (decode_var2.cu: CTA pairs with cta_group::2;
// decode_gram.cu, refine.cu): parameters, shared-memory carve-up, front-MLP producers.
#pragma once
#include "common.cuh"

namespace glasdi {
namespace dv {

constexpr int kTileM = 128;       // DOFs per MMA tile (TMEM lanes)
constexpr int kChunkN = 128;      // columns (samples or time-step x sample) per chunk
constexpr int kNtMax = 8;         // DOF tiles per group (accumulator slots per epilogue thread)
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

enum Epi { kVarT1 = 0, kVarSeg = 1, kStore = 2 };

struct Params {
  // VAR geometry
  int P, S, T, Spad;
  int TPC, n_tg, n_chunks;
  float inv_S;
  const float* Zs;                  // [P][T][NZ][Spad]
  unsigned int* maxvar_bits;        // [P]
  float* var_field;                 // optional [P][T][field_ld]: scaled variance of every (t, dof) (screening / std field)
  float* mean_field;                // optional, same layout: scaled mean over the samples of x(s) - x(pilot)
  int field_ld;                     // n_tiles rounded up to even * 128
  // STORE
  const float* Zrows;               // [rows][NZ]
  long long rows;
  float* X;                         // [rows][D]
  const float* b_last;              // [D]
  float out_scale;                  // 2^-(eW+eH)
  // common
  int NZ, D, K, Kpad, n_tiles, n_groups, passes;
  long long n_items;
  float h_scale;                    // 2^eH
  int check_overflow;               // unbounded activation: watch the fp16 split range
  const __half* w_packed;
  int* flags;
  FrontMlp front;
};

struct Smem {
  uint32_t w, h, wl, pilot, acc, bars, tmem, total;
  int wl_ld;
};

// row stride (floats) of the staged last-front-layer weights: [w_0 .. w_{inw-1}, bias, 0-pad] -> 16-byte rows
__host__ __device__ inline int wl_stride(int last_in_w) { return (last_in_w + 1 + 3) & ~3; }

__host__ __device__ inline Smem carve(int Kpad, int last_in_w) {
  Smem s;
  const uint32_t part = (uint32_t)Kpad * 256u;          // one 128-row x Kpad fp16 operand tile
  uint32_t off = 0;
  s.w = off; off += kWSlots * part;
  s.h = off; off += 2 * 2 * part;                      // 2 buffers x (hi, lo)
  s.wl_ld = wl_stride(last_in_w);
  s.wl = off; off += (uint32_t)Kpad * (uint32_t)s.wl_ld * 4u;
  s.pilot = off; off += kMaxSeg * (uint32_t)Kpad * 4u;
  s.acc = off; off += kNtMax * 2 * 128 * 4u;           // per-thread (sum, sumsq) for 8 tiles
  s.bars = off; off += 32 * 8;
  s.tmem = off; off += 16;
  s.total = off;
  return s;
}

// ---- front MLP -------------------------------------------------------------------------------
// layers 0 .. n_front-2 (small), fp32, weights via the read-only path; result = input of the last
// front layer.  For a single front layer this is the identity on z.
template <int ACT>
__device__ __forceinline__ void front_pre(const FrontMlp& f, const float* zin, float* out) {
  int w = f.widths[0];
  for (int i = 0; i < w; ++i) out[i] = zin[i];
  float tmp[kMaxHid];
  for (int l = 0; l + 1 < f.n_front; ++l) {
    const int wo = f.widths[l + 1];
    const float* __restrict__ W = f.W[l];
    const float* __restrict__ b = f.b[l];
    // four output units at a time, inputs in the outer loop: four independent FMA chains and one read of the
    // (local-memory) input per step instead of one serial chain per unit
    for (int o0 = 0; o0 < wo; o0 += 4) {
      float acc[4];
      const float* wr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int o = min(o0 + u, wo - 1);
        acc[u] = __ldg(b + o);
        wr[u] = W + (size_t)o * w;
      }
      for (int i = 0; i < w; ++i) {
        const float xi = out[i];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(__ldg(wr[u] + i), xi, acc[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (o0 + u < wo) tmp[o0 + u] = act_apply_t<ACT>(f.act, acc[u]);
    }
    for (int o = 0; o < wo; ++o) out[o] = tmp[o];
    w = wo;
  }
}

// 8 consecutive units k0..k0+7 of the last front layer.  Weights come from shared memory rows
// [w_0..w_{inw-1}, bias, 0...] read as 16-byte vectors (warp-uniform address -> one broadcast
// wavefront each).  Rows k >= K are all zero, so padded units evaluate to act(0) for every column
// AND for the pilot: their difference is exactly 0 and needs no predicate.
template <int ACT, int INW>
__device__ __forceinline__ void last_front_units(const float* __restrict__ wl, int ld, int inw_rt, int act_rt,
                                                 int k0, const float* x, float (&hv)[8]) {
  if constexpr (INW == 5) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4* wr = reinterpret_cast<const float4*>(wl + (k0 + u) * ld);
      const float4 a = wr[0], b = wr[1];
      float acc = b.y;
      acc = fmaf(a.x, x[0], acc);
      acc = fmaf(a.y, x[1], acc);
      acc = fmaf(a.z, x[2], acc);
      acc = fmaf(a.w, x[3], acc);
      acc = fmaf(b.x, x[4], acc);
      hv[u] = act_apply_t<ACT>(act_rt, acc);
    }
  } else {
    // inputs in the outer loop (four at a time), the eight units as independent chains
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = wl[(k0 + u) * ld + inw_rt];
    int i = 0;
    for (; i + 4 <= inw_rt; i += 4) {
      const float x0 = x[i], x1 = x[i + 1], x2 = x[i + 2], x3 = x[i + 3];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 w4 = *reinterpret_cast<const float4*>(wl + (k0 + u) * ld + i);     // ld is a multiple of 4
        acc[u] = fmaf(w4.x, x0, acc[u]);
        acc[u] = fmaf(w4.y, x1, acc[u]);
        acc[u] = fmaf(w4.z, x2, acc[u]);
        acc[u] = fmaf(w4.w, x3, acc[u]);
      }
    }
    for (; i < inw_rt; ++i) {
      const float xi = x[i];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = fmaf(wl[(k0 + u) * ld + i], xi, acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) hv[u] = act_apply_t<ACT>(act_rt, acc[u]);
  }
}

// ---- narrow front layers in REGISTERS (widths <= 24) --------------------------------------------------------------
// Shared-memory image per layer: transposed, zero-padded weights [32 in][32 out] followed by 32 biases.  A thread keeps the
// activations of its sample in registers (arrays indexed by fully unrolled loops), reads the weights as warp-uniform 16-byte
// vectors (one broadcast wavefront per 4 FMAs) and honours the widths in blocks of 8 inputs / 4 outputs.  Replaces the
// cooperative evaluation through shared memory with one named barrier per layer and weights from global memory (87 k
// instructions per 128-sample chunk at the Burgers2D-shaped decoder, 15 k of them FMAs: profiles/r02_gram_generic_*).
constexpr int kNarrowW = 24;
constexpr int kNarrowLayerFloats = kNarrowW * kNarrowW + kNarrowW;

template <int ACT>
__device__ __forceinline__ void narrow_layers_regs(const float* __restrict__ sNW, int n_pre, const int* __restrict__ widths,
                                                   int act_rt, float (&x)[kNarrowW]) {
  for (int l = 0; l < n_pre; ++l) {
    const int wi = widths[l], wo = widths[l + 1];
    const float* __restrict__ Wt = sNW + (size_t)l * kNarrowLayerFloats;
    float y[kNarrowW];
#pragma unroll
    for (int o4 = 0; o4 < kNarrowW / 4; ++o4) {
      const float4 b4 = *reinterpret_cast<const float4*>(Wt + kNarrowW * kNarrowW + o4 * 4);
      y[o4 * 4 + 0] = b4.x; y[o4 * 4 + 1] = b4.y; y[o4 * 4 + 2] = b4.z; y[o4 * 4 + 3] = b4.w;
    }
#pragma unroll
    for (int i8 = 0; i8 < kNarrowW / 8; ++i8) {
      if (i8 * 8 < wi) {
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
          const float xi = x[i8 * 8 + ii];
#pragma unroll
          for (int o4 = 0; o4 < kNarrowW / 4; ++o4) {
            if (o4 * 4 < wo) {
              const float4 w4 = *reinterpret_cast<const float4*>(Wt + (i8 * 8 + ii) * kNarrowW + o4 * 4);
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
    for (int o = 0; o < kNarrowW; ++o) x[o] = (o < wo) ? act_apply_t<ACT>(act_rt, y[o]) : 0.f;
  }
}

// writes the shared-memory image above (all `nthreads` threads of the caller's group; the caller synchronises)
__device__ __forceinline__ void narrow_layers_stage(float* sNW, const FrontMlp& f, int n_pre, int tid, int nthreads) {
  for (int l = 0; l < n_pre; ++l) {
    const int wi = f.widths[l], wo = f.widths[l + 1];
    float* Wt = sNW + (size_t)l * kNarrowLayerFloats;
    const float* __restrict__ W = f.W[l];
    const float* __restrict__ b = f.b[l];
    for (int idx = tid; idx < kNarrowW * kNarrowW; idx += nthreads) {
      const int i = idx / kNarrowW, o = idx - i * kNarrowW;
      Wt[idx] = (i < wi && o < wo) ? W[(size_t)o * wi + i] : 0.f;
    }
    for (int o = tid; o < kNarrowW; o += nthreads) Wt[kNarrowW * kNarrowW + o] = o < wo ? b[o] : 0.f;
  }
}

// last_front_units for an input that lives in registers (x zero beyond the layer's input width)
template <int ACT>
__device__ __forceinline__ void last_front_units_regs(const float* __restrict__ wl, int ld, int inw_rt, int act_rt, int k0,
                                                      const float (&x)[kNarrowW], float (&hv)[8]) {
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = wl[(k0 + u) * ld + inw_rt];
#pragma unroll
  for (int i4 = 0; i4 < kNarrowW / 4; ++i4) {
    if (i4 * 4 < inw_rt) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 w4 = *reinterpret_cast<const float4*>(wl + (k0 + u) * ld + i4 * 4);     // ld is a multiple of 4
        float a = acc[u];
        a = fmaf(w4.x, x[i4 * 4 + 0], a);
        a = fmaf(w4.y, x[i4 * 4 + 1], a);
        a = fmaf(w4.z, x[i4 * 4 + 2], a);
        a = fmaf(w4.w, x[i4 * 4 + 3], a);
        acc[u] = a;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) hv[u] = act_apply_t<ACT>(act_rt, acc[u]);
}

__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// ---- warp-level tensor-core MLP chain (mma.sync m16n8k16, fp16 split operands, fp32 accumulators) ----------------------
// Shared by the screening pass of the wide decoders (decode_wide.cu, wide_pre_mma_kernel) and the generic producers of
// gram_kernel (decode_gram.cu): a warp owns 16 samples; the accumulator fragment of two neighbouring n8 tiles of a layer
// IS the A fragment of one k16 block of the next layer, so narrow front layers chain register to register.
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_h2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  hi = h2_as_u32(h);
  lo = h2_as_u32(__floats2half2_rn(x - hf.x, y - hf.y));
}

constexpr int kPmMaxKb = kMaxHid / 16;              // k16 blocks of a narrow layer's input (<= 64 wide)
__host__ __device__ inline int pm_in_pad(int w) { return (w + 15) / 16 * 16; }
__host__ __device__ inline int pm_out_pad(int w, bool last) { return last ? (w + 7) / 8 * 8 : (w + 15) / 16 * 16; }
__host__ __device__ inline size_t pm_layer_halfs(int wi, int wo, bool last) {
  return 2 * (size_t)pm_out_pad(wo, last) * (pm_in_pad(wi) + 8);
}
// + biases (fp32) behind each layer's weights
__host__ inline size_t pm_smem_bytes(const int* widths, int nf) {
  size_t b = 0;
  for (int l = 0; l < nf; ++l) {
    b += pm_layer_halfs(widths[l], widths[l + 1], l == nf - 1) * sizeof(__half);
    b = (b + 15) / 16 * 16;
    b += (size_t)pm_out_pad(widths[l + 1], l == nf - 1) * sizeof(float);
    b = (b + 15) / 16 * 16;
  }
  return b;
}


// Writes the image of layers 0 .. nf-1 at `base` (every thread of the calling group, stride nthreads; the caller
// synchronises): per layer [hi | lo] fp16 weights [out_pad][in_pad + 8] + fp32 biases [out_pad].  `last_exact8`: the
// last layer's outputs are padded to 8 (its tiles are written out), the others to 16 (they feed the next layer's k blocks).
// Layer descriptor: 32-bit SHARED-space addresses (explicit ld.shared: generic loads of the weight image were long-scoreboard
// stalls, 19 % of the samples of the first version) -- the callers keep the descriptor array itself in shared memory too.
struct PmLayer { uint32_t wh, wl, bias; int nkb, op, ld, nt; };   // nt: n8 tiles with real units; ld in halfs
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void pm_stage(uint8_t* base, const FrontMlp& f, int nf, int tid, int nthreads, PmLayer* out) {
  size_t off = 0;
  for (int l = 0; l < nf; ++l) {
    const bool last = l == nf - 1;
    const int wi = f.widths[l], wo = f.widths[l + 1];
    const int ip = pm_in_pad(wi), op = pm_out_pad(wo, last), ld = ip + 8;
    __half* w = reinterpret_cast<__half*>(base + off);
    for (int idx = tid; idx < op * ip; idx += nthreads) {
      const int o = idx / ip, i = idx - o * ip;
      const float v = (o < wo && i < wi) ? f.W[l][(size_t)o * wi + i] : 0.f;
      const __half vh = __float2half_rn(v);
      w[(size_t)o * ld + i] = vh;
      w[(size_t)op * ld + (size_t)o * ld + i] = __float2half_rn(v - __half2float(vh));
    }
    off += pm_layer_halfs(wi, wo, last) * sizeof(__half);
    off = (off + 15) / 16 * 16;
    float* b = reinterpret_cast<float*>(base + off);
    for (int o = tid; o < op; o += nthreads) b[o] = o < wo ? f.b[l][o] : 0.f;
    if (tid == 0)
      out[l] = PmLayer{smem_u32(w), smem_u32(w + (size_t)op * ld), smem_u32(b), ip / 16, op, ld, (wo + 7) / 8};
    off += (size_t)op * sizeof(float);
    off = (off + 15) / 16 * 16;
  }
}
// n8 tiles j0 (and j1 if TWO) of a layer for the warp's 16 samples: acc = bias + A W^T, three split MMAs per k block (two
// without the weights' fp16 remainder, WLO = false: where the layer's OUTPUT is rounded to fp16 anyway and the rounding is
// common to the samples and the pilot).  The two tiles' accumulation chains are independent and issued interleaved; the
// number of k blocks is a template argument behind a warp-uniform switch (a predicated-off HMMA still takes its issue slot).
template <int NKB, int KBM, bool WLO, bool TWO>
__device__ __forceinline__ void pm_tiles_n(const PmLayer& Ly, int j0, int j1, int g, int t, const uint32_t (&Ah)[KBM][4],
                                           const uint32_t (&Al)[KBM][4], float (&a0)[4], float (&a1)[4]) {
  const uint32_t row = (uint32_t)(g * Ly.ld + 2 * t) * 2u, tile = (uint32_t)(8 * Ly.ld) * 2u;
  const uint32_t h0 = Ly.wh + row + (uint32_t)j0 * tile, l0 = Ly.wl + row + (uint32_t)j0 * tile;
  const uint32_t h1 = Ly.wh + row + (uint32_t)j1 * tile, l1 = Ly.wl + row + (uint32_t)j1 * tile;
  {
    const float2 b0 = lds_f2(Ly.bias + (uint32_t)(8 * j0 + 2 * t) * 4u);
    a0[0] = b0.x; a0[1] = b0.y; a0[2] = b0.x; a0[3] = b0.y;
    if constexpr (TWO) {
      const float2 b1 = lds_f2(Ly.bias + (uint32_t)(8 * j1 + 2 * t) * 4u);
      a1[0] = b1.x; a1[1] = b1.y; a1[2] = b1.x; a1[3] = b1.y;
    }
  }
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb) {
    const uint32_t x0 = lds_u32(h0 + 32u * kb), x1 = lds_u32(h0 + 32u * kb + 16u);
    uint32_t y0 = 0, y1 = 0;
    if constexpr (TWO) { y0 = lds_u32(h1 + 32u * kb); y1 = lds_u32(h1 + 32u * kb + 16u); }
    mma16816(a0, Ah[kb], x0, x1);
    if constexpr (TWO) mma16816(a1, Ah[kb], y0, y1);
    mma16816(a0, Al[kb], x0, x1);
    if constexpr (TWO) mma16816(a1, Al[kb], y0, y1);
    if constexpr (WLO) {
      const uint32_t u0 = lds_u32(l0 + 32u * kb), u1 = lds_u32(l0 + 32u * kb + 16u);
      mma16816(a0, Ah[kb], u0, u1);
      if constexpr (TWO) {
        const uint32_t v0 = lds_u32(l1 + 32u * kb), v1 = lds_u32(l1 + 32u * kb + 16u);
        mma16816(a1, Ah[kb], v0, v1);
      }
    }
  }
}
// KBM = k16 blocks the caller's fragment arrays hold (2 when every layer input is <= 32 wide: half the registers)
template <int KBM, bool WLO, bool TWO>
__device__ __forceinline__ void pm_tiles(const PmLayer& Ly, int j0, int j1, int g, int t, const uint32_t (&Ah)[KBM][4],
                                         const uint32_t (&Al)[KBM][4], float (&a0)[4], float (&a1)[4]) {
  if constexpr (KBM <= 2) {
    if (Ly.nkb == 1) pm_tiles_n<1, KBM, WLO, TWO>(Ly, j0, j1, g, t, Ah, Al, a0, a1);
    else pm_tiles_n<2, KBM, WLO, TWO>(Ly, j0, j1, g, t, Ah, Al, a0, a1);
  } else {
    switch (Ly.nkb) {
      case 1: pm_tiles_n<1, KBM, WLO, TWO>(Ly, j0, j1, g, t, Ah, Al, a0, a1); break;
      case 2: pm_tiles_n<2, KBM, WLO, TWO>(Ly, j0, j1, g, t, Ah, Al, a0, a1); break;
      case 3: pm_tiles_n<3, KBM, WLO, TWO>(Ly, j0, j1, g, t, Ah, Al, a0, a1); break;
      default: pm_tiles_n<4, KBM, WLO, TWO>(Ly, j0, j1, g, t, Ah, Al, a0, a1); break;
    }
  }
}
template <int KBM, bool WLO = true>
__device__ __forceinline__ void pm_tile(const PmLayer& Ly, int j, int g, int t, const uint32_t (&Ah)[KBM][4],
                                        const uint32_t (&Al)[KBM][4], float (&acc)[4]) {
  float dummy[4];
  pm_tiles<KBM, WLO, false>(Ly, j, j, g, t, Ah, Al, acc, dummy);
}
// layers 0 .. n_layers-1 (none of them the written-out one): A fragments of layer 0 in, A fragments of layer n_layers out
template <int KBM, typename ActF>
__device__ __forceinline__ void pm_chain(const PmLayer* Ly, int n_layers, int g, int t, uint32_t (&Ah)[KBM][4],
                                         uint32_t (&Al)[KBM][4], ActF act) {
  for (int l = 0; l < n_layers; ++l) {
    uint32_t Nh[KBM][4], Nl[KBM][4];
#pragma unroll
    for (int jp = 0; jp < KBM; ++jp) {
      if (2 * jp < Ly[l].nt) {
        float a0[4], a1[4] = {0.f, 0.f, 0.f, 0.f};
        if (2 * jp + 1 < Ly[l].nt) pm_tiles<KBM, true, true>(Ly[l], 2 * jp, 2 * jp + 1, g, t, Ah, Al, a0, a1);
        else pm_tiles<KBM, true, false>(Ly[l], 2 * jp, 2 * jp, g, t, Ah, Al, a0, a1);    // tile 2 jp + 1 is pure padding
#pragma unroll
        for (int q = 0; q < 4; ++q) { a0[q] = act(a0[q]); a1[q] = act(a1[q]); }
        split_h2(a0[0], a0[1], Nh[jp][0], Nl[jp][0]);      // row g,     k = 16 jp + 2t, + 1
        split_h2(a0[2], a0[3], Nh[jp][1], Nl[jp][1]);      // row g + 8
        split_h2(a1[0], a1[1], Nh[jp][2], Nl[jp][2]);      // row g,     k = 16 jp + 8 + 2t, + 1
        split_h2(a1[2], a1[3], Nh[jp][3], Nl[jp][3]);      // row g + 8
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) { Nh[jp][q] = 0u; Nl[jp][q] = 0u; }
      }
    }
#pragma unroll
    for (int kb = 0; kb < KBM; ++kb)
#pragma unroll
      for (int q = 0; q < 4; ++q) { Ah[kb][q] = Nh[kb][q]; Al[kb][q] = Nl[kb][q]; }
  }
}
// A fragments of layer 0 from the latent states of rows g and g + 8: this lane's components 2t, 2t + 1 (0 beyond n <= 8)
template <int KBM>
__device__ __forceinline__ void pm_latent_frag(float zlo0, float zlo1, float zhi0, float zhi1, uint32_t (&Ah)[KBM][4],
                                               uint32_t (&Al)[KBM][4]) {
#pragma unroll
  for (int kb = 0; kb < KBM; ++kb)
#pragma unroll
    for (int q = 0; q < 4; ++q) { Ah[kb][q] = 0u; Al[kb][q] = 0u; }
  split_h2(zlo0, zlo1, Ah[0][0], Al[0][0]);
  split_h2(zhi0, zhi1, Ah[0][1], Al[0][1]);
}

}  // namespace dv
}  // namespace glasdi
