// common.cuh -- shared declarations for the glasdi_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/glasdi_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "glasdi_b200 targets sm_100a (Blackwell B200) only"
#endif

namespace glasdi {

constexpr int kMaxLatent = 8;     // n_z <= 8
constexpr int kMaxLinear = 8;     // decoder linear layers <= 8

// Front (all-but-last) decoder layers, fp32, device pointers.  Passed by value to kernels.
struct FrontMlp {
  int n_front;                      // number of front linear layers (>= 1)
  int widths[kMaxLinear + 1];       // widths[0] = n_z ... widths[n_front] = K (input width of last layer)
  const float* W[kMaxLinear];       // [widths[l+1]][widths[l]] row-major
  const float* b[kMaxLinear];       // [widths[l+1]]
  int act;                          // enum glasdi_act
};

struct Decoder {
  bool ready = false;
  int n_linear = 0;
  int widths[kMaxLinear + 1] = {0};
  int act = 0;
  std::vector<float*> dW, db;       // device fp32 copies of every layer
  // last layer packed for tcgen05: [tile][part(hi,lo)][kg][128 rows][8 halfs]
  __half* w_packed = nullptr;
  int K = 0, Kpad = 0, D = 0, n_tiles = 0;
  int eW = 0;                       // power-of-two scale applied to W before the fp16 split
  int eH = 0;                       // power-of-two scale applied to the hidden activations
  bool tc_ok = false;               // tcgen05 path supports this shape
  // covariance-form screening (decode_gram.cu): V[dof, (i<=j)] = (2 - delta_ij) w_i w_j 2^eV, fp16 UMMA tiles
  __half* v_packed = nullptr;
  float* w4norm = nullptr;            // [ceil256(D)] |w_dof o w_dof|_2: per-dof factor of the a-priori screening bound
  int eV = 0;
  bool gram_ok = false;
  __half* wide_packed = nullptr;    // decode_wide.cu: [W_hi | W_hi | W_lo] tiles of the last layer (decoders with !tc_ok)
  __half* wide_front = nullptr;     // ... of the last FRONT layer (input width wide_k1 >= 16), scale 2^eW1
  int wide_k1 = 0, eW1 = 0;
  std::vector<float> h_front_last_W, h_front_last_b;   // host copy of the last FRONT layer (kernel-parameter weights)
};

// Encoder MLP (reference Autoencoder.encoder, latent_space.py:223): plain fp32 device copies, evaluated layer by layer.
struct Encoder {
  bool ready = false;
  int n_linear = 0;
  int act = 0;
  std::vector<int> widths;          // widths[0] = number of dofs ... widths[n_linear] = n_z
  std::vector<float*> dW, db;
};

}  // namespace glasdi

struct glasdi_handle {
  int device = 0;
  int num_sms = 0;
  std::string err;
  glasdi::Decoder dec;
  glasdi::Encoder enc;
  int opt_passes = 3;
  int opt_integrator = GLASDI_INT_EXPM;
  int opt_substeps = 1;
  int opt_library = GLASDI_LIB_POLY1;  // GLASDI_OPT_SINDY_LIBRARY
  int opt_mt_seg_log2 = 18;            // GLASDI_OPT_MT_SEGMENT_LOG2: words per MT19937 stream segment (0: one CTA)
  unsigned int* d_mt_polys = nullptr;  // [mt_poly_cap][624] jump polynomials t^(c 2^L) mod phi
  unsigned short* d_mt_pos = nullptr;  // exponents of phi, then of mu
  int mt_poly_cap = 0, mt_poly_count = 0, mt_poly_log2 = -1;
  int opt_screen_bound = 6;            // GLASDI_OPT_SCREEN_BOUND: sigmas of the a-priori rounding-error bound (0 = off)
  int opt_max_ctas = 0;
  int opt_timing = 0;
  int64_t launches = 0;
  // stage timing (GLASDI_OPT_STAGE_TIMING): event pairs of the last greedy call
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_rollout, ev_decode;   // indices of (start, stop) pairs into ev_pool
  // workspace (grown on demand)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  unsigned int* d_maxvar = nullptr;   // [P] float bits
  size_t maxvar_cap = 0;
  cudaMemPool_t pool = nullptr;       // stream-ordered scratch (release threshold = keep everything)
  int* d_flags = nullptr;             // [4] device error flags: [0] fp16 range, [1] candidate overflow, [2] screening deviation, [3] rollout step bound
  // screened mode (GLASDI_OPT_MODE = 1): candidates of the last greedy call
  int opt_mode = 2;                   // 0: direct (split passes), 1: x-space screening + refinement, 2: covariance-form screening + refinement
  int opt_wide = 2;                   // wide decoders: 0 CUDA cores only, 1 last layer on tcgen05 (decode_wide.cu), 2 + last front layer
  int opt_gram_kernel = 1;            // covariance GEMM of one-front-layer decoders: 1 gram2_kernel (Taylor producers), 0 gram_kernel
  int opt_taylor_radius_q8 = 256;     // gram2_kernel: radius of the polynomial in |da|, in 1/256 (0: exact form everywhere)
  int opt_margin_log2 = 8;            // candidates = variance within (1 - 2^-margin)^2 of the screened maximum
  void* d_cand = nullptr;             // [cand_cap] {p, t, d, approx variance}
  size_t cand_cap = 0;
  void* d_groups = nullptr;           // [group_cap] {p, t, first, count}: the candidates of one (p, t) row of the field
  size_t group_cap = 0;
  unsigned int* d_cand_stats = nullptr;   // [0] count (this slab), [1] total count, [2] float bits of max relative deviation, [3] row groups (this slab)
  unsigned int* d_exact = nullptr;    // [P] float bits of the exact variance
  size_t exact_cap = 0;
  std::vector<int> ev_refine, ev_quad;
};

namespace glasdi {

int set_error(glasdi_handle* h, int code, const std::string& msg);
int check_cuda(glasdi_handle* h, cudaError_t e, const char* what);
int ensure_workspace(glasdi_handle* h, size_t bytes);
int ensure_maxvar(glasdi_handle* h, size_t n);

// Entry points run on the handle's device and put the caller's current device back when they return.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) prev = cur;
    cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define GL_CUDA(h, expr)                                                   \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return glasdi::check_cuda((h), _e, #expr);      \
  } while (0)

// ---- kernel launchers (implemented in the .cu files) -----------------------------------------
// rollout.cu
int launch_rollout(glasdi_handle* h, const double* coefs, const float* z0, int z0_per_sample,
                   int P, int S, int n, int T, double dt,
                   float* Zs /*[P][T][n][Spad] or null*/, int Spad, double* Zref /*[P,S,T,n] or null*/,
                   cudaStream_t st, const double* dts = nullptr /* device [T-1] step sizes of a non-uniform grid */);
int library_terms(int lib, int n);   // |Theta| of enum glasdi_library, -1 if unknown
int launch_zis_to_soa(glasdi_handle* h, const double* Zis, int P, int S, int T, int n, float* Zs, int Spad,
                      cudaStream_t st);
int launch_pad_columns(glasdi_handle* h, float* Zs, long long rows, int S, int Spad, cudaStream_t st);
// decode_var2.cu
size_t decode_var_smem_bytes(int Kpad, int last_in_w);
// decode_var2.cu : CTA-pair (cta_group::2) variant of the same kernel
// var_field (optional): [P][T][decode_field_ld] scaled variance of every (t, dof); passes > 0 overrides the handle option
int launch_decode_var_pair(glasdi_handle* h, const float* Zs, int Spad, int P, int S, int T,
                           unsigned int* maxvar_bits, float* var_field, int passes, cudaStream_t st,
                           float* mean_field = nullptr);
inline int decode_field_ld(int n_tiles) { return (n_tiles + 1) / 2 * 2 * 128; }
// decode_gram.cu : covariance-form screening pass (gram_kernel + quad_kernel)
int gram_nkgQ(int K);
size_t gram_a_bytes(int K, long long n_items);
size_t gram_fscale_bytes(long long n_items);
size_t decode_gram_smem_bytes(int K, int last_in_w, int xw /* widest intermediate front layer, rounded up to 4; 0 if none */);
int gram_pack_v(glasdi_handle* h, const float* WL_host);
size_t decode_gram2_smem_bytes(int K);
bool gram2_supported(const Decoder& d);
int launch_decode_gram(glasdi_handle* h, const float* Zs, int Spad, int P, int S, int T, void* a_ws, float* fscale_ws,
                       cudaStream_t st);
int launch_decode_quad(glasdi_handle* h, int P, int T, unsigned int* maxvar_bits, float* var_field, const void* a_ws,
                       const float* fscale_ws, cudaStream_t st);
// refine.cu : candidate selection on the screened variance field + exact fp32/fp64 re-evaluation
int launch_gram_diag_norm(glasdi_handle* h, const void* a_ws, const float* fscale_ws, long long n_items, float* cnorm,
                          cudaStream_t st);
// cnorm / w4norm / beta: a-priori rounding-error bound of the covariance-form screening (null / 0: off); keep2 = the
// fraction of the screened maximum a (t, dof) must reach WITH its bound added to become a candidate as well
int launch_select_candidates(glasdi_handle* h, const float* var_field, int field_ld, const unsigned int* approx_bits,
                             int P, int T, int D, float keep, const float* cnorm, const float* w4norm, float beta,
                             float keep2, cudaStream_t st);
int launch_refine(glasdi_handle* h, const float* Zs, int Spad, int S, int T, unsigned int* exact_bits,
                  float dev_tol, float unscale, cudaStream_t st);
int launch_decode_store_pair(glasdi_handle* h, const float* Z, int64_t rows, float* X, cudaStream_t st);
// simt_fallback.cu : fp32 CUDA-core path for decoder shapes the tcgen05 kernel does not cover
// field != nullptr: screening launch (ONE fp16 pass on the tensor cores, variance field [P * T][field_ld] in true units)
// passes > 0 overrides the number of fp16 passes (3 + field = fp32-faithful FIELD output, exact activations); mean_field
// (optional, same layout): mean over the samples of x(s) - x(pilot) in true units
int launch_decode_var_simt(glasdi_handle* h, const float* Zs, int Spad, int P, int S, int T,
                           unsigned int* maxvar_bits, float* field, int field_ld, cudaStream_t st, int passes = 0,
                           float* mean_field = nullptr);
// exact re-evaluation of the screened candidates of a wide decoder: fp32 CUDA-core chain + fp64 sample sums
int launch_refine_wide(glasdi_handle* h, const float* Zs, int Spad, int S, int T, unsigned int* exact_bits, float dev_tol,
                       cudaStream_t st);
int launch_decode_store_simt(glasdi_handle* h, const float* Z, int64_t rows, float* X, cudaStream_t st);
int launch_encode_simt(glasdi_handle* h, const float* U, int64_t rows, float* Z, cudaStream_t st);
// decode_wide.cu : last layer + variance of wide decoders (K_last > 112) as a K-looped three-pass tcgen05 GEMM
int wide_pack_w(glasdi_handle* h, const float* WL_host, const float* W1_host, int K1);
// parts = 3: fp32-faithful [hi | lo | hi] split passes, parts = 1: one fp16 screening pass
size_t wide_b_bytes(int K, int S, long long n_groups, int parts);
size_t wide_x_bytes(int K1, int S, long long n_groups, int parts);
int launch_wide_front(glasdi_handle* h, const float* X, long long n_groups, int S, void* xt_ws, void* bt_ws, int parts,
                      cudaStream_t st);
bool wide_pre_supported(const Decoder& d);
int launch_wide_pre(glasdi_handle* h, const float* Zs, int Spad, long long g0, long long n_groups, int S, void* xt_ws,
                    cudaStream_t st);
int launch_wide_last_var(glasdi_handle* h, const float* Hlast, long long g0, long long n_groups, int S, int T, void* bt_ws,
                         unsigned int* maxvar_bits, int parts, float* field, int field_ld, cudaStream_t st,
                         float* mean_field = nullptr);
// candidate records of the screened modes (refine.cu, simt_fallback.cu)
namespace rf {
struct Cand {
  int p, t, d;
  float var;       // screened variance (field units)
};
struct Group {     // work unit: up to 8 candidates of one (p, t) row
  int p, t;
  unsigned int first;      // candidates [first, first + (count & 0xff)) of cand[]
  unsigned int count;      // low byte: candidates of this unit; bits 8..31: candidates of the whole row in the row's FIRST unit, else 0
};
}  // namespace rf
// reduce.cu
int launch_finalize(glasdi_handle* h, const unsigned int* maxvar_bits, int P, int scale_exp2,
                    float* max_std, int64_t* argmax, float* maxval, cudaStream_t st);
int launch_argmax_first(glasdi_handle* h, const float* vals, int P, int64_t* idx, float* val, cudaStream_t st);
int launch_pack_maxloc(glasdi_handle* h, const float* val, const int64_t* idx, int64_t off, int64_t* key,
                       cudaStream_t st);
// sample-statistics fields (reduce.cu): pilot rows of Zs -> [rows][n]; scaled (mean, var) fields -> mean / std [rows][D]
int launch_gather_pilot_rows(glasdi_handle* h, const float* Zs, int Spad, long long rows, int n, float* Zrows,
                             cudaStream_t st);
int launch_finalize_fields(glasdi_handle* h, const float* mean_f, const float* var_f, int field_ld, long long rows, int D,
                           int scale_exp2, float* mean_io, float* std_out, cudaStream_t st);
// calibrate.cu
int launch_fd_dzdt(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt, float* dZ,
                   cudaStream_t st);
int launch_sindy_fit(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt, int norm_order,
                     float* coefs, float* loss_sindy, float* loss_coef, int32_t* info, cudaStream_t st);
int launch_sindy_fit_backward(glasdi_handle* h, const float* Z, const float* coefs, const float* g_sindy,
                              const float* g_coef, int B, int T, int n, int fd_id, double dt, int norm_order,
                              float* Zbar, cudaStream_t st);

// gp.cu : GP posterior moments and numpy-legacy (MT19937 + polar Box-Muller) normal draws
int launch_gp_predict(glasdi_handle* h, const double* Xtr, int ntr, int d, int n_gp, const double* alpha,
                      const double* L, const double* amp, const double* len, const double* ymean, const double* ystd,
                      const double* Xte, int P, double* mean, double* stdv, cudaStream_t st);
int launch_normal_mt19937(glasdi_handle* h, glasdi_mt19937_state* state, const double* mean, const double* stdv, int P,
                          int S, int n, double* out, cudaStream_t st);

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float act_apply(int act, float x) {
  // torch-default activations (reference src/lasdi/latent_space.py:5-28)
  switch (act) {
    case GLASDI_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case GLASDI_ACT_TANH: return tanhf(x);
    case GLASDI_ACT_SOFTPLUS: return x > 20.0f ? x : log1pf(expf(x));   // beta=1, threshold=20
    case GLASDI_ACT_RELU: return fmaxf(x, 0.0f);
    case GLASDI_ACT_ELU: return x > 0.0f ? x : expm1f(x);              // alpha=1
    case GLASDI_ACT_SILU: return x / (1.0f + expf(-x));
    case GLASDI_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case GLASDI_ACT_HARDSHRINK: return fabsf(x) > 0.5f ? x : 0.0f;
    case GLASDI_ACT_HARDSIGMOID: return fminf(fmaxf(x * (1.0f / 6.0f) + 0.5f, 0.0f), 1.0f);
    case GLASDI_ACT_HARDTANH: return fminf(fmaxf(x, -1.0f), 1.0f);
    case GLASDI_ACT_HARDSWISH: return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    case GLASDI_ACT_LEAKYRELU: return x >= 0.0f ? x : 0.01f * x;
    case GLASDI_ACT_LOGSIGMOID: return fminf(x, 0.0f) - log1pf(expf(-fabsf(x)));
    case GLASDI_ACT_RELU6: return fminf(fmaxf(x, 0.0f), 6.0f);
    case GLASDI_ACT_SELU: return 1.0507009873554805f * (x > 0.0f ? x : 1.6732632423543772f * expm1f(x));
    case GLASDI_ACT_CELU: return x > 0.0f ? x : expm1f(x);             // alpha=1
    case GLASDI_ACT_MISH: return x * tanhf(x > 20.0f ? x : log1pf(expf(x)));
    case GLASDI_ACT_SOFTSHRINK: return x > 0.5f ? x - 0.5f : (x < -0.5f ? x + 0.5f : 0.0f);
    case GLASDI_ACT_TANHSHRINK: return x - tanhf(x);
    default: return x;
  }
}

// Compile-time activation for the hot producers.  ACT < 0 selects the runtime switch above.
// sigmoid / softplus / SiLU use the MUFU ex2/rcp/lg2 units (max rel. error ~2^-21, i.e. ~4 fp32 ulp:
// far inside the 1e-5 parity budget); everything else is the accurate libdevice form.
template <int ACT>
__device__ __forceinline__ float act_apply_t(int act_rt, float x) {
  if constexpr (ACT == GLASDI_ACT_SIGMOID) {
    return __fdividef(1.0f, 1.0f + __expf(-x));
  } else if constexpr (ACT == GLASDI_ACT_SOFTPLUS) {
    return x > 20.0f ? x : __logf(1.0f + __expf(x));
  } else if constexpr (ACT == GLASDI_ACT_SILU) {
    return __fdividef(x, 1.0f + __expf(-x));
  } else if constexpr (ACT >= 0) {
    return act_apply(ACT, x);
  } else if constexpr (ACT == -2) {      // runtime activation, MUFU forms where they exist (screening passes)
    switch (act_rt) {
      case GLASDI_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-x));
      case GLASDI_ACT_SOFTPLUS: return x > 20.0f ? x : __logf(1.0f + __expf(x));
      case GLASDI_ACT_SILU: return __fdividef(x, 1.0f + __expf(-x));
      case GLASDI_ACT_TANH: return fmaf(2.0f, __fdividef(1.0f, 1.0f + __expf(-2.0f * x)), -1.0f);
      default: return act_apply(act_rt, x);
    }
  } else {
    return act_apply(act_rt, x);
  }
}

// ---- screening-pass activations: bare MUFU forms, activation resolved at compile time (ACT < 0: runtime switch) ----
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// bare MUFU forms, activation resolved at compile time: 3-6 instructions per value (the runtime switch + __expf / __logf
// range handling of act_fast cost ~25 in xact_kernel's epilogue: 40 instructions per value, profiles/r02_xact_*)
template <int ACT>
__device__ __forceinline__ float act_mufu(float a, int act_rt) {   // ACT < 0: runtime switch (the rarely used activations)
  constexpr float kL2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  if constexpr (ACT == GLASDI_ACT_SIGMOID) return mufu_rcp(1.0f + mufu_ex2(-kL2e * a));
  else if constexpr (ACT == GLASDI_ACT_TANH) return fmaf(2.0f, mufu_rcp(1.0f + mufu_ex2(-2.0f * kL2e * a)), -1.0f);
  else if constexpr (ACT == GLASDI_ACT_SOFTPLUS) return fmaf(kLn2, mufu_lg2(1.0f + mufu_ex2(-kL2e * fabsf(a))), fmaxf(a, 0.0f));
  else if constexpr (ACT == GLASDI_ACT_SILU) return a * mufu_rcp(1.0f + mufu_ex2(-kL2e * a));
  else if constexpr (ACT == GLASDI_ACT_RELU) return fmaxf(a, 0.0f);
  else if constexpr (ACT == GLASDI_ACT_IDENTITY) return a;
  else if constexpr (ACT < 0) return act_apply(act_rt, a);
  else return act_apply(ACT, a);
}

// packed fp32 pairs (Blackwell FFMA2 / FADD2): one instruction for two lanes of the epilogue sums
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b,
                                                        unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in HW, wake on phase flip
      : "memory");
  return ok != 0;
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// waiters that run far ahead of the tensor pipe (producers, epilogue, loader, relay): back off instead
// of polling -- 2.1e9 polling retries per launch showed up in the first profiles and cost power
// (mbarrier.try_wait with a suspend-time hint compiles to its own loop of SYNCS.PHASECHK.TRYWAIT + NANOSLEEP.SYNCS that comes
// back every ~40 clocks until the hint expires -- a warp "suspended" in it issued 28-36 % of all instructions of the Gram
// kernels (profiles/r02_gram_generic_*): long waits poll with the non-blocking test_wait and a real sleep instead)
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) __nanosleep(128);
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  while (!mbar_test(bar, parity)) __nanosleep(ns);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- bulk async copy (TMA engine, 1-D) --------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- tcgen05 ----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16 inputs, fp32 accumulate), single CTA.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread l <-> lane base+l)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// K-major, no-swizzle ("interleaved") UMMA shared-memory descriptor.
// Canonical layout in 16-byte units: ((8, n), 2) : ((1, SBO), LBO) -- a core matrix is 8 rows x 16 B
// stored contiguously; SBO = byte step between 8-row groups, LBO = byte step between the two
// 16-byte K halves of one K=16 slice.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  // base_offset = 0, lbo_mode = 0, layout_type (bits 61..63) = 0 : SWIZZLE_NONE
  return d;
}

// kind::f16 instruction descriptor: fp16 A/B (K-major both), fp32 accumulator, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_f16_m128(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;              // c_format = F32
  // a_format = b_format = 0 (F16); no negate; a_major = b_major = 0 (K-major)
  d |= (n >> 3) << 17;       // n_dim
  d |= (128u >> 4) << 24;    // m_dim
  return d;
}

// ---- thread-block clusters / CTA pairs (cta_group::2) -----------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctaid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// arrive on an mbarrier given by its shared::cluster address (may be in the peer CTA).  Default semantics
// (release at CTA scope) compile to the bare SYNCS.ARRIVE; `.release.cluster` costs MEMBAR.ALL.GPU + ERRBAR +
// CGAERRBAR per arrive, which was 27 % of the epilogue warps' time (profiles/: decode_r1_screen).  What the
// arrivals publish is either consumed through the async proxy after an explicit fence.proxy.async (operand
// tiles) or is the completion of tcgen05.ld reads (accumulator release): neither needs the GPU-scope fence.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) {
  }
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit all prior MMAs of this thread; arrive on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// CTA-pair MMA: D[tmem, 256 x N split 128 rows per CTA] (+)= A[256 x 16] * B[N x 16]^T, issued by the
// leader CTA; each CTA supplies its own 128 rows of A and N/2 rows of B from its shared memory.
__device__ __forceinline__ void tc_mma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 instruction descriptor for the pair: M = 256, N = n
__device__ __forceinline__ uint32_t umma_idesc_f16_m256(uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;              // c_format = F32
  d |= (n >> 3) << 17;       // n_dim
  d |= (256u >> 4) << 24;    // m_dim
  return d;
}

// one lane of a fully converged warp (warp-uniform control flow keeps MMA operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

}  // namespace glasdi
