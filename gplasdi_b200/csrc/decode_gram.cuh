// decode_gram.cuh -- declarations shared by the covariance-form screening kernels (decode_gram.cu: gram_kernel + quad_kernel,
// decode_gram2.cu: gram2_kernel).
#pragma once
#include "decode_common.cuh"

namespace glasdi {
namespace gr {

using namespace dv;

constexpr int kCols = 128;                  // sample columns per chunk
constexpr int kRing = 4;                    // chunk ring
constexpr int kGroupsM = 16;                // unit groups of 8 in a tile (M = 128)
constexpr uint32_t kBufBytes = kGroupsM * kCols * 16;   // 32 KB
// Roles: warps 0..3 = consumers (warp 0 also issues the MMAs, then all four run the epilogue of the item),
// warps 4..19 = producers.  20 warps = 640 threads leave 96 registers per thread: the producers are bound by
// instruction latency (MUFU, shared-memory loads), which four producer warps per SM sub-partition hide where
// two could not (profiles/: gram_r1d: 48 % issue, 32 % XU with two).
constexpr int kGEpiWarps = 4;
constexpr int kGProdWarp0 = 4;
constexpr int kGProdWarps = 16;
constexpr int kGProdThreads = kGProdWarps * 32;
constexpr int kGThreads = (kGProdWarp0 + kGProdWarps) * 32;
constexpr int kGParts = kGProdThreads / 64;   // thread groups a chunk's tasks are dealt to (fast path)
constexpr int kGTmemCols = 128;             // generic path: one 128 x 128 fp32 accumulator
constexpr int kGTmemColsFast = 512;         // fast path: + three 128-column pre-activation buffers
constexpr uint32_t kA1Bytes = 4 * 2048;     // dz tile: 128 samples x 16 tf32 (hi | lo), K-major, 4 K chunks of 16 B

struct GramParams {
  int P, S, T, Spad, NZ;
  int K;                 // real hidden units (input width of the last layer); unit index K is the constant unit
  int nkg;               // unit groups the producers fill = ceil((K + 1) / 8)
  int Nmma;              // N of the MMA = ceil((K + 1) / 16) * 16 <= 128
  int nkgQ;              // 16-byte groups of the packed triangle, padded to the quad kernel's K chunk
  int Q;                 // K (K + 1) / 2
  float h_scale, inv_S;
  float poly_thr2;       // gram2_kernel: |z - z_pilot|^2 up to which a 32-sample block takes the Taylor polynomial (< 0: never)
  int exp_base;          // 2 eH + eV : exponent of the fixed part of the scale
  int check_overflow;
  int xw;                // generic path with >= 2 front layers: widest intermediate layer, rounded up to 4 (else 0)
  long long n_items;     // P * T
  const float* Zs;       // [P][T][NZ][Spad]
  __half* A;             // [item block of 128][nkgQ][128 rows][8]  : A operand of the quad kernel
  float* fscale;         // [n_items] : variance = accumulator * fscale
  int* flags;
  FrontMlp front;
};

struct GBars {
  uint64_t h_full[kRing], h_empty[kRing], a1_full[kRing], a1_empty[kRing], d1_full[3], tab_full[4], acc_full, acc_empty;
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x2(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32 (fp32 containers, 10-bit mantissa, fp32 accumulate), K = 8
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32 instruction descriptor: tf32 A/B (K-major both), fp32 accumulator, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32_m128(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;              // c_format = F32
  d |= 2u << 7;              // a_format = TF32
  d |= 2u << 10;             // b_format = TF32
  d |= (n >> 3) << 17;
  d |= (128u >> 4) << 24;
  return d;
}

__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

// kind::f16, fp32 accumulate, M = 128, both operands MN-major (the contraction index -- samples -- is the
// strided one: canonical layout ((8,1,m),(8,k)) : ((1,8,SBO),(8,LBO)) in elements)
__device__ __forceinline__ uint32_t umma_idesc_f16_m128_mn(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;              // c_format = F32
  d |= 1u << 15;             // a_major = MN
  d |= 1u << 16;             // b_major = MN
  d |= (n >> 3) << 17;
  d |= (128u >> 4) << 24;
  return d;
}

__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// volatile: a plain asm is a pure function to the compiler, which then hoists the pilot's loads above their
// `ptid < nkg` guard and executes them for every producer thread -- out of the shared-memory window
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned long long lds64_volatile(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}

constexpr float kNegLog2e = -1.4426950408889634f;


}  // namespace gr
}  // namespace glasdi
