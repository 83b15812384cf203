// api.cu -- extern "C" entry points of libglasdi_b200.so (see include/glasdi_b200.h).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace glasdi {

int set_error(glasdi_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

int check_cuda(glasdi_handle* h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return GLASDI_OK;
  return set_error(h, e == cudaErrorMemoryAllocation ? GLASDI_ERR_OOM : GLASDI_ERR_CUDA,
                   std::string(what) + ": " + cudaGetErrorString(e));
}

int ensure_workspace(glasdi_handle* h, size_t bytes) {
  if (bytes <= h->ws_bytes) return GLASDI_OK;
  if (h->ws) cudaFree(h->ws);
  h->ws = nullptr;
  h->ws_bytes = 0;
  GL_CUDA(h, cudaMalloc(&h->ws, bytes));
  h->ws_bytes = bytes;
  return GLASDI_OK;
}

int ensure_maxvar(glasdi_handle* h, size_t n) {
  if (n <= h->maxvar_cap) return GLASDI_OK;
  if (h->d_maxvar) cudaFree(h->d_maxvar);
  h->d_maxvar = nullptr;
  h->maxvar_cap = 0;
  GL_CUDA(h, cudaMalloc(&h->d_maxvar, n * sizeof(unsigned int)));
  h->maxvar_cap = n;
  return GLASDI_OK;
}

// candidate list + exact-variance bits of the screened mode
static int ensure_screen_buffers(glasdi_handle* h, size_t P) {
  if (!h->d_cand) {
    const size_t cap = (size_t)1 << 20;                        // 16 MB of {p, t, d, var}
    GL_CUDA(h, cudaMalloc(&h->d_cand, cap * 16));
    h->cand_cap = cap;
  }
  if (!h->d_groups) {
    const size_t gcap = (size_t)1 << 19;                       // 8 MB of {p, t, first, count}
    GL_CUDA(h, cudaMalloc(&h->d_groups, gcap * 16));
    h->group_cap = gcap;
  }
  if (!h->d_cand_stats) GL_CUDA(h, cudaMalloc(&h->d_cand_stats, 8 * sizeof(unsigned int)));
  if (P > h->exact_cap) {
    if (h->d_exact) cudaFree(h->d_exact);
    h->d_exact = nullptr;
    h->exact_cap = 0;
    GL_CUDA(h, cudaMalloc(&h->d_exact, P * sizeof(unsigned int)));
    h->exact_cap = P;
  }
  return GLASDI_OK;
}

static void free_decoder(Decoder& d) {
  for (float* p : d.dW) if (p) cudaFree(p);
  for (float* p : d.db) if (p) cudaFree(p);
  d.dW.clear();
  d.db.clear();
  if (d.w_packed) cudaFree(d.w_packed);
  d.w_packed = nullptr;
  if (d.v_packed) cudaFree(d.v_packed);
  d.v_packed = nullptr;
  if (d.w4norm) cudaFree(d.w4norm);
  d.w4norm = nullptr;
  if (d.wide_packed) cudaFree(d.wide_packed);
  d.wide_packed = nullptr;
  if (d.wide_front) cudaFree(d.wide_front);
  d.wide_front = nullptr;
  d.wide_k1 = 0;
  d.gram_ok = false;
  d.ready = false;
  d.tc_ok = false;
}

static void free_encoder(Encoder& e) {
  for (float* p : e.dW) if (p) cudaFree(p);
  for (float* p : e.db) if (p) cudaFree(p);
  e.dW.clear();
  e.db.clear();
  e.ready = false;
}

// activations with values in [-1, 1]: the fp16 operands of the hidden layer may be scaled by 2^12 instead of 2^4
static bool bounded_act(int act) {
  return act == GLASDI_ACT_SIGMOID || act == GLASDI_ACT_TANH || act == GLASDI_ACT_HARDSIGMOID || act == GLASDI_ACT_HARDTANH;
}

// check-and-clear the device-side numeric flags (activation overflow of the fp16 split)
static int check_flags(glasdi_handle* h, cudaStream_t st) {
  int flags[4] = {0, 0, 0, 0};
  GL_CUDA(h, cudaMemcpyAsync(flags, h->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
  GL_CUDA(h, cudaStreamSynchronize(st));
  if (flags[0] || flags[1] || flags[2] || flags[3]) {
    GL_CUDA(h, cudaMemsetAsync(h->d_flags, 0, sizeof(flags), st));
    if (flags[3])
      return set_error(h, GLASDI_ERR_NUMERIC,
                       "latent rollout: a trajectory of the polynomial / sine library needs more than 16384 RK4 sub-steps "
                       "per output step or turned non-finite (finite-time blow-up)");
    if (flags[0])
      return set_error(h, GLASDI_ERR_NUMERIC,
                       "hidden activations overflowed the fp16 split range (|h - h_pilot| * 2^eH > 6e4)");
    if (flags[1])
      return set_error(h, GLASDI_ERR_NUMERIC,
                       "screened greedy step: more near-maximum (t, dof) candidates than the candidate buffer holds; "
                       "use GLASDI_OPT_MODE = 0 (direct) or a larger GLASDI_OPT_SCREEN_MARGIN");
    return set_error(h, GLASDI_ERR_NUMERIC,
                     "screened greedy step: single-pass screening deviates from the exact re-evaluation by more than "
                     "a quarter of the candidate margin; use GLASDI_OPT_MODE = 0 (direct) or a smaller "
                     "GLASDI_OPT_SCREEN_MARGIN");
  }
  return GLASDI_OK;
}

}  // namespace glasdi

using namespace glasdi;

extern "C" {

int glasdi_abi_version(void) { return GLASDI_ABI_VERSION; }

#ifndef GLASDI_SOURCE_HASH
#define GLASDI_SOURCE_HASH "unknown"
#endif
const char* glasdi_source_hash(void) { return GLASDI_SOURCE_HASH; }

int glasdi_create(glasdi_handle** out, int device) {
  if (!out) return GLASDI_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return GLASDI_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return GLASDI_ERR_INVALID;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GLASDI_ERR_CUDA;
  if (prop.major != 10) return GLASDI_ERR_NO_DEVICE;   // tcgen05 / TMEM kernels: sm_100 only, no fallback
  glasdi_handle* h = new glasdi_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaError_t e = cudaMalloc(&h->d_flags, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(h->d_flags, 0, 4 * sizeof(int));
  if (e == cudaSuccess) {
    // stream-ordered scratch of the CUDA-core / wide-decoder paths comes from the handle's own pool, which keeps its memory
    // across synchronisations (the default pool returns it to the driver at every sync: 10-40 ms per GB-sized slab)
    cudaMemPoolProps pp{};
    pp.allocType = cudaMemAllocationTypePinned;
    pp.handleTypes = cudaMemHandleTypeNone;
    pp.location.type = cudaMemLocationTypeDevice;
    pp.location.id = device;
    e = cudaMemPoolCreate(&h->pool, &pp);
    if (e == cudaSuccess) {
      unsigned long long keep = ~0ull;
      e = cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaSetDevice(prev);
  if (e != cudaSuccess) { delete h; return GLASDI_ERR_CUDA; }
  *out = h;
  return GLASDI_OK;
}

int glasdi_destroy(glasdi_handle* h) {
  if (!h) return GLASDI_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  free_decoder(h->dec);
  free_encoder(h->enc);
  if (h->ws) cudaFree(h->ws);
  if (h->d_maxvar) cudaFree(h->d_maxvar);
  if (h->d_flags) cudaFree(h->d_flags);
  if (h->d_mt_polys) cudaFree(h->d_mt_polys);
  if (h->d_mt_pos) cudaFree(h->d_mt_pos);
  if (h->d_cand) cudaFree(h->d_cand);
  if (h->d_groups) cudaFree(h->d_groups);
  if (h->d_cand_stats) cudaFree(h->d_cand_stats);
  if (h->d_exact) cudaFree(h->d_exact);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  if (h->pool) {
    cudaDeviceSynchronize();
    cudaMemPoolDestroy(h->pool);
  }
  cudaSetDevice(prev);
  delete h;
  return GLASDI_OK;
}

const char* glasdi_last_error(const glasdi_handle* h) { return h ? h->err.c_str() : "null handle"; }

int64_t glasdi_launch_count(const glasdi_handle* h) { return h ? h->launches : 0; }

int glasdi_set_option(glasdi_handle* h, int option, int value) {
  if (!h) return GLASDI_ERR_INVALID;
  switch (option) {
    case GLASDI_OPT_SPLIT_PASSES:
      if (value < 1 || value > 3) return set_error(h, GLASDI_ERR_INVALID, "split passes must be 1, 2 or 3");
      h->opt_passes = value;
      return GLASDI_OK;
    case GLASDI_OPT_INTEGRATOR:
      if (value != GLASDI_INT_EXPM && value != GLASDI_INT_RK4) return set_error(h, GLASDI_ERR_INVALID, "unknown integrator");
      h->opt_integrator = value;
      return GLASDI_OK;
    case GLASDI_OPT_RK4_SUBSTEPS:
      if (value < 1) return set_error(h, GLASDI_ERR_INVALID, "substeps must be >= 1");
      h->opt_substeps = value;
      return GLASDI_OK;
    case GLASDI_OPT_MAX_CTAS:
      if (value < 0) return set_error(h, GLASDI_ERR_INVALID, "max ctas must be >= 0");
      h->opt_max_ctas = value;
      return GLASDI_OK;
    case GLASDI_OPT_STAGE_TIMING:
      h->opt_timing = value ? 1 : 0;
      return GLASDI_OK;
    case GLASDI_OPT_DECODE_KERNEL:
      if (value != 1)
        return set_error(h, GLASDI_ERR_INVALID, "decode kernel variant must be 1 (the one-CTA-per-SM kernel of round 1 was retired)");
      return GLASDI_OK;
    case GLASDI_OPT_MODE:
      if (value != GLASDI_MODE_DIRECT && value != GLASDI_MODE_SCREENED && value != GLASDI_MODE_SCREENED_GRAM)
        return set_error(h, GLASDI_ERR_INVALID, "mode must be GLASDI_MODE_DIRECT, _SCREENED or _SCREENED_GRAM");
      h->opt_mode = value;
      return GLASDI_OK;
    case GLASDI_OPT_SCREEN_MARGIN:
      if (value < 4 || value > 16) return set_error(h, GLASDI_ERR_INVALID, "screen margin must be 4..16 (delta = 2^-value)");
      h->opt_margin_log2 = value;
      return GLASDI_OK;
    case GLASDI_OPT_WIDE_TENSOR:
      if (value < 0 || value > 3) return set_error(h, GLASDI_ERR_INVALID, "wide-tensor option must be 0..3");
      h->opt_wide = value;
      return GLASDI_OK;
    case GLASDI_OPT_GRAM_KERNEL:
      if (value != 0 && value != 1) return set_error(h, GLASDI_ERR_INVALID, "gram kernel variant must be 0 or 1");
      h->opt_gram_kernel = value;
      return GLASDI_OK;
    case GLASDI_OPT_TAYLOR_RADIUS:
      if (value < 0 || value > 1024) return set_error(h, GLASDI_ERR_INVALID, "Taylor radius must be 0..1024 (units of 1/256)");
      h->opt_taylor_radius_q8 = value;
      return GLASDI_OK;
    case GLASDI_OPT_MT_SEGMENT_LOG2:
      if (value != 0 && (value < 11 || value > 24))
        return set_error(h, GLASDI_ERR_INVALID, "MT19937 segment size must be 0 (one CTA) or 2^11..2^24 words");
      h->opt_mt_seg_log2 = value;
      return GLASDI_OK;
    case GLASDI_OPT_SCREEN_BOUND:
      if (value < 0 || value > 64) return set_error(h, GLASDI_ERR_INVALID, "screening bound must be 0..64 sigmas");
      h->opt_screen_bound = value;
      return GLASDI_OK;
    case GLASDI_OPT_SINDY_LIBRARY:
      if (library_terms(value, 1) < 0) return set_error(h, GLASDI_ERR_INVALID, "unknown SINDy library id");
      h->opt_library = value;
      return GLASDI_OK;
    default: return set_error(h, GLASDI_ERR_INVALID, "unknown option");
  }
}

int glasdi_library_terms(int library, int n) { return library_terms(library, n); }

int glasdi_get_option(const glasdi_handle* h, int option, int* value) {
  if (!h || !value) return GLASDI_ERR_INVALID;
  switch (option) {
    case GLASDI_OPT_SPLIT_PASSES: *value = h->opt_passes; return GLASDI_OK;
    case GLASDI_OPT_INTEGRATOR: *value = h->opt_integrator; return GLASDI_OK;
    case GLASDI_OPT_RK4_SUBSTEPS: *value = h->opt_substeps; return GLASDI_OK;
    case GLASDI_OPT_MAX_CTAS: *value = h->opt_max_ctas; return GLASDI_OK;
    case GLASDI_OPT_STAGE_TIMING: *value = h->opt_timing; return GLASDI_OK;
    case GLASDI_OPT_DECODE_KERNEL: *value = 1; return GLASDI_OK;
    case GLASDI_OPT_MODE: *value = h->opt_mode; return GLASDI_OK;
    case GLASDI_OPT_SCREEN_MARGIN: *value = h->opt_margin_log2; return GLASDI_OK;
    case GLASDI_OPT_WIDE_TENSOR: *value = h->opt_wide; return GLASDI_OK;
    case GLASDI_OPT_GRAM_KERNEL: *value = h->opt_gram_kernel; return GLASDI_OK;
    case GLASDI_OPT_TAYLOR_RADIUS: *value = h->opt_taylor_radius_q8; return GLASDI_OK;
    case GLASDI_OPT_SINDY_LIBRARY: *value = h->opt_library; return GLASDI_OK;
    case GLASDI_OPT_SCREEN_BOUND: *value = h->opt_screen_bound; return GLASDI_OK;
    case GLASDI_OPT_MT_SEGMENT_LOG2: *value = h->opt_mt_seg_log2; return GLASDI_OK;
    default: return GLASDI_ERR_INVALID;
  }
}

int glasdi_set_decoder(glasdi_handle* h, int n_linear, const int* widths, const float* const* W_host,
                       const float* const* b_host, int act_id) {
  if (!h) return GLASDI_ERR_INVALID;
  if (n_linear < 2 || n_linear > kMaxLinear) return set_error(h, GLASDI_ERR_INVALID, "decoder needs 2..8 linear layers");
  if (!widths || !W_host || !b_host) return set_error(h, GLASDI_ERR_INVALID, "null decoder arrays");
  if (act_id < 0 || act_id >= GLASDI_ACT_COUNT) return set_error(h, GLASDI_ERR_INVALID, "unknown activation id");
  if (widths[0] < 1 || widths[0] > kMaxLatent) return set_error(h, GLASDI_ERR_INVALID, "latent dimension must be 1..8");
  for (int l = 0; l <= n_linear; ++l)
    if (widths[l] < 1) return set_error(h, GLASDI_ERR_INVALID, "non-positive layer width");
  DeviceGuard device_guard(h->device);
  Decoder& d = h->dec;
  free_decoder(d);
  d.n_linear = n_linear;
  d.act = act_id;
  for (int l = 0; l <= n_linear; ++l) d.widths[l] = widths[l];
  d.dW.assign(n_linear, nullptr);
  d.db.assign(n_linear, nullptr);
  for (int l = 0; l < n_linear; ++l) {
    const size_t nw = (size_t)widths[l] * widths[l + 1];
    GL_CUDA(h, cudaMalloc(&d.dW[l], nw * sizeof(float)));
    GL_CUDA(h, cudaMalloc(&d.db[l], (size_t)widths[l + 1] * sizeof(float)));
    GL_CUDA(h, cudaMemcpy(d.dW[l], W_host[l], nw * sizeof(float), cudaMemcpyHostToDevice));
    GL_CUDA(h, cudaMemcpy(d.db[l], b_host[l], (size_t)widths[l + 1] * sizeof(float), cudaMemcpyHostToDevice));
  }
  d.K = widths[n_linear - 1];
  d.D = widths[n_linear];
  d.Kpad = (d.K + 15) / 16 * 16;
  d.n_tiles = (d.D + 127) / 128;

  // tcgen05 path: K of the last layer must fit the resident operand tiles, intermediate front
  // layers must fit the producers' local arrays.
  bool ok = d.Kpad <= 112;
  for (int l = 1; l < n_linear - 1; ++l) ok = ok && widths[l] <= 64;
  const int last_in_w = widths[n_linear - 2];
  ok = ok && decode_var_smem_bytes(d.Kpad, last_in_w) <= 227 * 1024;
  d.tc_ok = ok;

  // ---- pack the last layer: fp16 hi/lo split of W * 2^eW in UMMA canonical K-major tiles -------
  const float* WL = W_host[n_linear - 1];
  float wmax = 0.f;
  for (size_t i = 0; i < (size_t)d.D * d.K; ++i) wmax = std::max(wmax, std::fabs(WL[i]));
  if (!std::isfinite(wmax)) return set_error(h, GLASDI_ERR_NUMERIC, "non-finite decoder weights");
  d.eW = (wmax > 0.f) ? 13 - std::ilogb(wmax) : 0;
  d.eW = std::max(-100, std::min(100, d.eW));
  d.eH = bounded_act(act_id) ? 12 : 4;
  if (ok) {
    const size_t part_halfs = (size_t)d.Kpad * 128;     // [kg][128 rows][8]
    const int n_tiles_pad = (d.n_tiles + 1) / 2 * 2;   // CTA-pair kernel consumes tiles two at a time
    std::vector<__half> packed((size_t)n_tiles_pad * 2 * part_halfs, __float2half_rn(0.f));
    for (int tile = 0; tile < d.n_tiles; ++tile) {
      __half* hi = packed.data() + ((size_t)tile * 2 + 0) * part_halfs;
      __half* lo = packed.data() + ((size_t)tile * 2 + 1) * part_halfs;
      for (int kg = 0; kg < d.Kpad / 8; ++kg)
        for (int r = 0; r < 128; ++r)
          for (int u = 0; u < 8; ++u) {
            const int dof = tile * 128 + r, k = kg * 8 + u;
            float v = 0.f;
            if (dof < d.D && k < d.K) v = std::ldexp(WL[(size_t)dof * d.K + k], d.eW);
            const __half vh = __float2half_rn(v);
            const __half vl = __float2half_rn(v - __half2float(vh));
            const size_t o = ((size_t)kg * 128 + r) * 8 + u;
            hi[o] = vh;
            lo[o] = vl;
          }
    }
    GL_CUDA(h, cudaMalloc(&d.w_packed, packed.size() * sizeof(__half)));
    GL_CUDA(h, cudaMemcpy(d.w_packed, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  if (!ok) {   // wide decoders: the last layer goes through decode_wide.cu
    int rcw = wide_pack_w(h, WL, n_linear >= 3 ? W_host[n_linear - 2] : nullptr, n_linear >= 3 ? widths[n_linear - 2] : 0);
    if (rcw) return rcw;
  }
  // covariance-form screening: the constant unit must fit the 128-row tile, shared memory must hold the ring
  int xw = 0;
  for (int l = 1; l < n_linear - 1; ++l) xw = std::max(xw, (widths[l] + 3) / 4 * 4);
  d.gram_ok = ok && d.K + 1 <= 128 && decode_gram_smem_bytes(d.K, last_in_w, xw) <= 227 * 1024;
  d.h_front_last_W.assign(W_host[n_linear - 2], W_host[n_linear - 2] + (size_t)d.K * last_in_w);
  d.h_front_last_b.assign(b_host[n_linear - 2], b_host[n_linear - 2] + d.K);
  if (d.gram_ok) {
    int rc = gram_pack_v(h, WL);
    if (rc) return rc;
  }
  d.ready = true;
  return GLASDI_OK;
}

int glasdi_rollout(glasdi_handle* h, const double* coefs, const float* z0, int z0_per_sample, int P, int S, int n,
                   int T, double dt, double* Z_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!coefs || !z0 || !Z_out) return set_error(h, GLASDI_ERR_INVALID, "rollout: null pointer");
  DeviceGuard device_guard(h->device);
  // P may exceed the grid.y limit: slab it
  for (int p0 = 0; p0 < P; p0 += 32768) {
    const int pn = std::min(32768, P - p0);
    int rc = launch_rollout(h, coefs + (size_t)p0 * S * n * library_terms(h->opt_library, n),
                            z0 + (size_t)p0 * (z0_per_sample ? (size_t)S * n : (size_t)n), z0_per_sample, pn, S, n, T,
                            dt, nullptr, 0, Z_out + (size_t)p0 * S * T * n, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return GLASDI_OK;
}

int glasdi_rollout_grid(glasdi_handle* h, const double* coefs, const float* z0, int z0_per_sample, int P, int S, int n,
                        int T, const double* t_grid_host, double* Z_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!coefs || !z0 || !Z_out || !t_grid_host) return set_error(h, GLASDI_ERR_INVALID, "rollout_grid: null pointer");
  if (T < 1) return set_error(h, GLASDI_ERR_INVALID, "rollout_grid: empty time grid");
  DeviceGuard device_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<double> dts((size_t)std::max(1, T - 1), 0.0);
  for (int k = 1; k < T; ++k) {
    dts[k - 1] = t_grid_host[k] - t_grid_host[k - 1];
    if (!(std::fabs(dts[k - 1]) < 1e300)) return set_error(h, GLASDI_ERR_INVALID, "rollout_grid: non-finite time grid");
  }
  double* d_dts = nullptr;
  GL_CUDA(h, cudaMallocFromPoolAsync((void**)&d_dts, dts.size() * sizeof(double), h->pool, st));
  GL_CUDA(h, cudaMemcpyAsync(d_dts, dts.data(), dts.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  GL_CUDA(h, cudaStreamSynchronize(st));                    // `dts` is a host temporary
  int rc = GLASDI_OK;
  const size_t cstride = (size_t)S * n * library_terms(h->opt_library, n);
  for (int p0 = 0; p0 < P && rc == GLASDI_OK; p0 += 32768) {
    const int pn = std::min(32768, P - p0);
    rc = launch_rollout(h, coefs + (size_t)p0 * cstride, z0 + (size_t)p0 * (z0_per_sample ? (size_t)S * n : (size_t)n),
                        z0_per_sample, pn, S, n, T, 0.0, nullptr, 0, Z_out + (size_t)p0 * S * T * n, st, d_dts);
  }
  cudaFreeAsync(d_dts, st);
  return rc;
}

// event pair helpers for GLASDI_OPT_STAGE_TIMING
static int ev_get(glasdi_handle* h, size_t idx, cudaEvent_t* out) {
  while (h->ev_pool.size() <= idx) {
    cudaEvent_t e;
    GL_CUDA(h, cudaEventCreate(&e));
    h->ev_pool.push_back(e);
  }
  *out = h->ev_pool[idx];
  return GLASDI_OK;
}
struct StageTimer {
  glasdi_handle* h; cudaStream_t st; std::vector<int>* list; bool on; size_t base;
  StageTimer(glasdi_handle* h_, cudaStream_t st_, std::vector<int>* l, size_t& next)
      : h(h_), st(st_), list(l), on(h_->opt_timing != 0), base(next) { if (on) next += 2; }
  int start() { if (!on) return 0; cudaEvent_t e; int rc = ev_get(h, base, &e); if (rc) return rc; return check_cuda(h, cudaEventRecord(e, st), "event record"); }
  int stop() { if (!on) return 0; cudaEvent_t e; int rc = ev_get(h, base + 1, &e); if (rc) return rc; list->push_back((int)base); return check_cuda(h, cudaEventRecord(e, st), "event record"); }
};

// shared tail of the two max-std entry points: run the decode+variance kernel over parameter slabs
// whose fp32 trajectory scratch fits the workspace budget
static int greedy_core(glasdi_handle* h, const double* coefs, const float* z0, const double* Zis, int P, int S, int n,
                       int T, double dt, float* max_std_out, int64_t* argmax_out, float* maxval_out,
                       cudaStream_t st) {
  const Decoder& d = h->dec;
  if (!d.ready) return set_error(h, GLASDI_ERR_NOT_READY, "decoder weights not set (glasdi_set_decoder)");
  if (n != d.widths[0]) return set_error(h, GLASDI_ERR_INVALID, "latent dimension does not match the decoder");
  if (P <= 0 || S <= 0 || T <= 0) return set_error(h, GLASDI_ERR_INVALID, "empty ensemble");
  if (!max_std_out) return set_error(h, GLASDI_ERR_INVALID, "null output");
  DeviceGuard device_guard(h->device);
  const int Spad = (S + 127) / 128 * 128;        // whole 128-sample chunks (covariance-form producers)
  // screened mode: one fp16 pass on the tensor cores writes the variance field, refine.cu re-evaluates
  // the near-maximum candidates exactly.  Only the CTA-pair tcgen05 kernel has the field epilogue.
  const bool gram = h->opt_mode == GLASDI_MODE_SCREENED_GRAM && d.gram_ok;
  // wide decoders (K_last > 112, decode_wide.cu): one fp16 screening pass of the K-looped GEMMs + the exact CUDA-core chain
  // at the candidates' rows, instead of three fp32-faithful split passes
  const bool wide_screen = !d.tc_ok && d.wide_packed != nullptr && h->opt_wide >= 1 && h->opt_mode != GLASDI_MODE_DIRECT;
  const bool screened = gram || wide_screen || (h->opt_mode != GLASDI_MODE_DIRECT && d.tc_ok);
  const int field_ld = wide_screen ? (d.D + 255) / 256 * 256 : decode_field_ld(d.n_tiles);
  const size_t traj_bytes = (size_t)T * n * Spad * sizeof(float);
  const size_t field_bytes = screened ? (size_t)T * field_ld * sizeof(float) : 0;
  const size_t gram_bytes = gram ? (size_t)T * (gram_nkgQ(d.K) * 16 + sizeof(float)) : 0;   // packed covariances + scales
  const size_t per_param = traj_bytes + field_bytes + gram_bytes;
  const size_t budget = (size_t)12 << 30;                      // trajectory (+ field, + covariances) scratch per slab
  int slab = (int)std::max<size_t>(1, std::min<size_t>((size_t)P, budget / per_param));
  slab = std::min(slab, 32768);
  const size_t a_off = (traj_bytes + field_bytes) * slab;
  const size_t a_bytes = gram ? gram_a_bytes(d.K, (long long)slab * T) : 0;
  const size_t fs_bytes = gram ? gram_fscale_bytes((long long)slab * T) : 0;
  const bool bound = gram && h->opt_screen_bound > 0 && d.w4norm != nullptr;
  int rc = ensure_workspace(h, a_off + a_bytes + fs_bytes * (bound ? 2 : 1));
  if (rc) return rc;
  float* cnorm = bound ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h->ws) + a_off + a_bytes + fs_bytes) : nullptr;
  // sigmas * 1.15 * 2^-11: standard deviation of the fp16 rounding error of the quadratic form (refine.cu, select_rows_kernel)
  const float beta = (float)h->opt_screen_bound * 1.15f * ldexpf(1.0f, -11);
  rc = ensure_maxvar(h, (size_t)P);
  if (rc) return rc;
  GL_CUDA(h, cudaMemsetAsync(h->d_maxvar, 0, (size_t)P * sizeof(unsigned int), st));
  if (screened) {
    if ((rc = ensure_screen_buffers(h, (size_t)P))) return rc;
    GL_CUDA(h, cudaMemsetAsync(h->d_exact, 0, (size_t)P * sizeof(unsigned int), st));
    GL_CUDA(h, cudaMemsetAsync(h->d_cand_stats, 0, 8 * sizeof(unsigned int), st));
  }
  float* Zs = reinterpret_cast<float*>(h->ws);
  float* field = screened ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h->ws) + traj_bytes * slab) : nullptr;
  // the one-pass x-space screening of the wide decoders carries the fp16 rounding of the last layer's WEIGHTS, an error
  // common to all samples of a dof that does not average out over the ensemble (measured 1e-3 on sigma at S = 24, 3.5e-4 at
  // S = 300, against 5e-5 for the covariance form): its candidate margin is one bit wider
  const float delta = ldexpf(1.0f, -(h->opt_margin_log2 - (wide_screen ? 1 : 0)));
  h->ev_rollout.clear();
  h->ev_decode.clear();
  h->ev_refine.clear();
  h->ev_quad.clear();
  size_t ev_next = 0;
  for (int p0 = 0; p0 < P; p0 += slab) {
    const int pn = std::min(slab, P - p0);
    StageTimer t_roll(h, st, &h->ev_rollout, ev_next), t_dec(h, st, &h->ev_decode, ev_next),
        t_ref(h, st, &h->ev_refine, ev_next), t_quad(h, st, &h->ev_quad, ev_next);
    if ((rc = t_roll.start())) return rc;
    if (coefs) {
      rc = launch_rollout(h, coefs + (size_t)p0 * S * n * library_terms(h->opt_library, n), z0 + (size_t)p0 * n, 0, pn, S, n, T, dt, Zs, Spad,
                          nullptr, st);
    } else {
      rc = launch_zis_to_soa(h, Zis + (size_t)p0 * S * T * n, pn, S, T, n, Zs, Spad, st);
    }
    if (rc) return rc;
    if (gram && (rc = launch_pad_columns(h, Zs, (long long)pn * T * n, S, Spad, st))) return rc;
    if ((rc = t_roll.stop())) return rc;
    if ((rc = t_dec.start())) return rc;
    if (gram) rc = launch_decode_gram(h, Zs, Spad, pn, S, T, reinterpret_cast<uint8_t*>(h->ws) + a_off,
                                      reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h->ws) + a_off + a_bytes), st);
    else if (wide_screen) rc = launch_decode_var_simt(h, Zs, Spad, pn, S, T, h->d_maxvar + p0, field, field_ld, st);
    else if (screened) rc = launch_decode_var_pair(h, Zs, Spad, pn, S, T, h->d_maxvar + p0, field, 1, st);
    else if (d.tc_ok) rc = launch_decode_var_pair(h, Zs, Spad, pn, S, T, h->d_maxvar + p0, nullptr, 0, st);
    else rc = launch_decode_var_simt(h, Zs, Spad, pn, S, T, h->d_maxvar + p0, nullptr, 0, st);
    if (rc) return rc;
    if ((rc = t_dec.stop())) return rc;
    if (gram) {
      if ((rc = t_quad.start())) return rc;
      rc = launch_decode_quad(h, pn, T, h->d_maxvar + p0, field, reinterpret_cast<uint8_t*>(h->ws) + a_off,
                              reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h->ws) + a_off + a_bytes), st);
      if (rc) return rc;
      if ((rc = t_quad.stop())) return rc;
    }
    if (screened) {
      if ((rc = t_ref.start())) return rc;
      GL_CUDA(h, cudaMemsetAsync(h->d_cand_stats, 0, sizeof(unsigned int), st));      // per-slab candidate count
      GL_CUDA(h, cudaMemsetAsync(h->d_cand_stats + 3, 0, sizeof(unsigned int), st));  // per-slab row-group count
      if (bound && (rc = launch_gram_diag_norm(h, reinterpret_cast<uint8_t*>(h->ws) + a_off,
                                               reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h->ws) + a_off + a_bytes),
                                               (long long)pn * T, cnorm, st)))
        return rc;
      rc = launch_select_candidates(h, field, field_ld, h->d_maxvar + p0, pn, T, d.D, (1.f - delta) * (1.f - delta), cnorm,
                                    d.w4norm, beta, (1.f - 0.5f * delta) * (1.f - 0.5f * delta), st);
      if (rc) return rc;
      if (wide_screen) rc = launch_refine_wide(h, Zs, Spad, S, T, h->d_exact + p0, 0.25f * delta, st);
      else rc = launch_refine(h, Zs, Spad, S, T, h->d_exact + p0, 0.25f * delta,
                              gram ? 1.0f : ldexpf(1.0f, -2 * (d.eW + d.eH)), st);
      if (rc) return rc;
      if ((rc = t_ref.stop())) return rc;
    }
  }
  if (screened) return launch_finalize(h, h->d_exact, P, 0, max_std_out, argmax_out, maxval_out, st);
  const int scale = d.tc_ok ? (d.eW + d.eH) : 0;
  return launch_finalize(h, h->d_maxvar, P, scale, max_std_out, argmax_out, maxval_out, st);
}

int glasdi_greedy_step(glasdi_handle* h, const double* coefs, const float* z0, int P, int S, int n, int T, double dt,
                       float* max_std_out, int64_t* argmax_out, float* maxval_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!coefs || !z0) return set_error(h, GLASDI_ERR_INVALID, "greedy_step: null input");
  return greedy_core(h, coefs, z0, nullptr, P, S, n, T, dt, max_std_out, argmax_out, maxval_out, (cudaStream_t)stream);
}

int glasdi_decode_max_std(glasdi_handle* h, const double* Zis, int P, int S, int T, int n, float* max_std_out,
                          int64_t* argmax_out, float* maxval_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!Zis) return set_error(h, GLASDI_ERR_INVALID, "decode_max_std: null input");
  return greedy_core(h, nullptr, nullptr, Zis, P, S, n, T, 0.0, max_std_out, argmax_out, maxval_out,
                     (cudaStream_t)stream);
}

int glasdi_decode_stat_fields(glasdi_handle* h, const double* Zis, int P, int S, int T, int n, float* mean_out,
                              float* std_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  const Decoder& d = h->dec;
  if (!d.ready) return set_error(h, GLASDI_ERR_NOT_READY, "decoder weights not set (glasdi_set_decoder)");
  const bool wide = !d.tc_ok;
  if (wide && (d.wide_packed == nullptr || h->opt_wide < 1))
    return set_error(h, GLASDI_ERR_INVALID, "stat fields of a wide decoder need GLASDI_OPT_WIDE_TENSOR >= 1");
  if (!Zis || (!mean_out && !std_out)) return set_error(h, GLASDI_ERR_INVALID, "stat fields: null pointer");
  if (n != d.widths[0]) return set_error(h, GLASDI_ERR_INVALID, "latent dimension does not match the decoder");
  if (P <= 0 || S <= 0 || T <= 0) return set_error(h, GLASDI_ERR_INVALID, "empty ensemble");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard device_guard(h->device);
  const int Spad = (S + 127) / 128 * 128;
  const int field_ld = wide ? (d.D + 255) / 256 * 256 : decode_field_ld(d.n_tiles);
  const size_t traj_bytes = (size_t)T * n * Spad * sizeof(float);
  const size_t field_bytes = (size_t)T * field_ld * sizeof(float);
  const size_t rows_bytes = ((size_t)T * n * sizeof(float) + 255) / 256 * 256;
  const size_t per_param = traj_bytes + 2 * field_bytes + rows_bytes;
  const size_t budget = (size_t)12 << 30;
  int slab = (int)std::max<size_t>(1, std::min<size_t>((size_t)P, budget / per_param));
  slab = std::min(slab, 32768);
  int rc = ensure_workspace(h, per_param * slab);
  if (rc) return rc;
  rc = ensure_maxvar(h, (size_t)P);
  if (rc) return rc;
  uint8_t* ws = reinterpret_cast<uint8_t*>(h->ws);
  float* Zs = reinterpret_cast<float*>(ws);
  float* varf = reinterpret_cast<float*>(ws + traj_bytes * slab);
  float* meanf = reinterpret_cast<float*>(ws + (traj_bytes + field_bytes) * slab);
  float* zrows = reinterpret_cast<float*>(ws + (traj_bytes + 2 * field_bytes) * slab);
  for (int p0 = 0; p0 < P; p0 += slab) {
    const int pn = std::min(slab, P - p0);
    const long long rows = (long long)pn * T;
    if ((rc = launch_zis_to_soa(h, Zis + (size_t)p0 * S * T * n, pn, S, T, n, Zs, Spad, st))) return rc;
    GL_CUDA(h, cudaMemsetAsync(h->d_maxvar, 0, (size_t)pn * sizeof(unsigned int), st));
    // three split passes: the fields are fp32-faithful (this is an output, not a screening pass)
    // wide decoders (K_last > 112): exact CUDA-core front layers, xact_kernel / xvar_kernel with three split passes and
    // the field epilogue (variance and mean of x(s) - x(pilot) in true units)
    if (wide) rc = launch_decode_var_simt(h, Zs, Spad, pn, S, T, h->d_maxvar, varf, field_ld, st, 3, meanf);
    else rc = launch_decode_var_pair(h, Zs, Spad, pn, S, T, h->d_maxvar, varf, 3, st, meanf);
    if (rc) return rc;
    float* mean_p = mean_out ? mean_out + (size_t)p0 * T * d.D : nullptr;
    if (mean_p) {
      if ((rc = launch_gather_pilot_rows(h, Zs, Spad, rows, n, zrows, st))) return rc;
      if (wide) rc = launch_decode_store_simt(h, zrows, rows, mean_p, st);
      else rc = launch_decode_store_pair(h, zrows, rows, mean_p, st);                 // decode(pilot), exact fp32
      if (rc) return rc;
    }
    if ((rc = launch_finalize_fields(h, meanf, varf, field_ld, rows, d.D, wide ? 0 : d.eW + d.eH, mean_p,
                                     std_out ? std_out + (size_t)p0 * T * d.D : nullptr, st)))
      return rc;
  }
  return GLASDI_OK;
}

int glasdi_decode(glasdi_handle* h, const float* Z, int64_t rows, float* X_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!h->dec.ready) return set_error(h, GLASDI_ERR_NOT_READY, "decoder weights not set (glasdi_set_decoder)");
  if (rows < 0 || (rows > 0 && (!Z || !X_out))) return set_error(h, GLASDI_ERR_INVALID, "decode: bad arguments");
  if (rows == 0) return GLASDI_OK;
  DeviceGuard device_guard(h->device);
  if (h->dec.tc_ok) return launch_decode_store_pair(h, Z, rows, X_out, (cudaStream_t)stream);
  return launch_decode_store_simt(h, Z, rows, X_out, (cudaStream_t)stream);
}

int glasdi_stage_times(glasdi_handle* h, double* rollout_ms, double* decode_ms, int* decode_launches) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!h->opt_timing) return set_error(h, GLASDI_ERR_NOT_READY, "stage timing is off (GLASDI_OPT_STAGE_TIMING)");
  double acc[2] = {0.0, 0.0};
  const std::vector<int>* lists[2] = {&h->ev_rollout, &h->ev_decode};
  for (int k = 0; k < 2; ++k)
    for (int base : *lists[k]) {
      GL_CUDA(h, cudaEventSynchronize(h->ev_pool[base + 1]));
      float ms = 0.f;
      GL_CUDA(h, cudaEventElapsedTime(&ms, h->ev_pool[base], h->ev_pool[base + 1]));
      acc[k] += ms;
    }
  if (rollout_ms) *rollout_ms = acc[0];
  if (decode_ms) *decode_ms = acc[1];
  if (decode_launches) *decode_launches = (int)h->ev_decode.size();
  return GLASDI_OK;
}

int glasdi_screen_stats(glasdi_handle* h, double* refine_ms, double* quad_ms, int64_t* n_candidates, float* max_rel_dev,
                        void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  DeviceGuard device_guard(h->device);
  double* outs[2] = {refine_ms, quad_ms};
  const std::vector<int>* lists[2] = {&h->ev_refine, &h->ev_quad};
  for (int k = 0; k < 2; ++k) {
    if (!outs[k]) continue;
    *outs[k] = 0.0;
    if (h->opt_timing)
      for (int base : *lists[k]) {
        GL_CUDA(h, cudaEventSynchronize(h->ev_pool[base + 1]));
        float ms = 0.f;
        GL_CUDA(h, cudaEventElapsedTime(&ms, h->ev_pool[base], h->ev_pool[base + 1]));
        *outs[k] += ms;
      }
  }
  unsigned int st[4] = {0, 0, 0, 0};
  if (h->d_cand_stats) {
    GL_CUDA(h, cudaMemcpyAsync(st, h->d_cand_stats, sizeof(st), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GL_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  }
  if (n_candidates) *n_candidates = (int64_t)st[1];
  if (max_rel_dev) std::memcpy(max_rel_dev, &st[2], sizeof(float));
  return GLASDI_OK;
}

int glasdi_screen_bound_count(glasdi_handle* h, int64_t* n_bound_candidates, void* stream) {
  if (!h || !n_bound_candidates) return GLASDI_ERR_INVALID;
  DeviceGuard device_guard(h->device);
  unsigned int v = 0;
  if (h->d_cand_stats) {
    GL_CUDA(h, cudaMemcpyAsync(&v, h->d_cand_stats + 4, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GL_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  }
  *n_bound_candidates = (int64_t)v;
  return GLASDI_OK;
}

// host-visible check of the numeric flags raised by the last asynchronous calls (synchronises `stream`)
int glasdi_check_numeric(glasdi_handle* h, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  DeviceGuard device_guard(h->device);
  return check_flags(h, (cudaStream_t)stream);
}

int glasdi_argmax_first(glasdi_handle* h, const float* vals, int P, int64_t* idx_out, float* val_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!vals || P <= 0 || !idx_out) return set_error(h, GLASDI_ERR_INVALID, "argmax: bad arguments");
  DeviceGuard device_guard(h->device);
  return launch_argmax_first(h, vals, P, idx_out, val_out, (cudaStream_t)stream);
}

int glasdi_pack_maxloc(glasdi_handle* h, const float* val, const int64_t* local_idx, int64_t index_offset,
                       int64_t* key_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!val || !local_idx || !key_out) return set_error(h, GLASDI_ERR_INVALID, "pack_maxloc: null pointer");
  DeviceGuard device_guard(h->device);
  return launch_pack_maxloc(h, val, local_idx, index_offset, key_out, (cudaStream_t)stream);
}

int glasdi_fd_dzdt(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt, float* dZ_out,
                   void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!Z || !dZ_out) return set_error(h, GLASDI_ERR_INVALID, "fd_dzdt: null pointer");
  DeviceGuard device_guard(h->device);
  return launch_fd_dzdt(h, Z, B, T, n, fd_id, dt, dZ_out, (cudaStream_t)stream);
}

int glasdi_sindy_fit(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt, int norm_order,
                     float* coefs_out, float* loss_sindy_out, float* loss_coef_out, int32_t* info_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!Z || !coefs_out) return set_error(h, GLASDI_ERR_INVALID, "sindy_fit: null pointer");
  DeviceGuard device_guard(h->device);
  return launch_sindy_fit(h, Z, B, T, n, fd_id, dt, norm_order, coefs_out, loss_sindy_out, loss_coef_out, info_out,
                          (cudaStream_t)stream);
}

int glasdi_gp_predict(glasdi_handle* h, const double* X_train, int n_train, int n_param, int n_gp, const double* alpha,
                      const double* L, const double* amplitude, const double* length_scale, const double* y_mean,
                      const double* y_std, const double* X_test, int P, double* mean_out, double* std_out,
                      void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!X_train || !alpha || !L || !amplitude || !length_scale || !X_test || !mean_out || !std_out)
    return set_error(h, GLASDI_ERR_INVALID, "gp_predict: null pointer");
  if (n_train < 1 || n_param < 1 || n_gp < 1 || P < 1) return set_error(h, GLASDI_ERR_INVALID, "gp_predict: empty input");
  DeviceGuard device_guard(h->device);
  return launch_gp_predict(h, X_train, n_train, n_param, n_gp, alpha, L, amplitude, length_scale, y_mean, y_std, X_test,
                           P, mean_out, std_out, (cudaStream_t)stream);
}

int glasdi_normal_mt19937(glasdi_handle* h, glasdi_mt19937_state* state, const double* mean, const double* std, int P,
                          int S, int n, double* out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!state || P < 0 || S < 0 || n < 0) return set_error(h, GLASDI_ERR_INVALID, "normal_mt19937: bad arguments");
  if ((long long)P * S * n > 0 && (!mean || !std || !out))
    return set_error(h, GLASDI_ERR_INVALID, "normal_mt19937: null pointer");
  DeviceGuard device_guard(h->device);
  return launch_normal_mt19937(h, state, mean, std, P, S, n, out, (cudaStream_t)stream);
}

int glasdi_set_encoder(glasdi_handle* h, int n_linear, const int* widths, const float* const* W_host,
                       const float* const* b_host, int act_id) {
  if (!h) return GLASDI_ERR_INVALID;
  if (n_linear < 1 || n_linear > 16) return set_error(h, GLASDI_ERR_INVALID, "encoder needs 1..16 linear layers");
  if (!widths || !W_host || !b_host) return set_error(h, GLASDI_ERR_INVALID, "null encoder arrays");
  if (act_id < 0 || act_id >= GLASDI_ACT_COUNT) return set_error(h, GLASDI_ERR_INVALID, "unknown activation id");
  for (int l = 0; l <= n_linear; ++l)
    if (widths[l] < 1) return set_error(h, GLASDI_ERR_INVALID, "non-positive layer width");
  DeviceGuard device_guard(h->device);
  Encoder& e = h->enc;
  free_encoder(e);
  e.n_linear = n_linear;
  e.act = act_id;
  e.widths.assign(widths, widths + n_linear + 1);
  e.dW.assign(n_linear, nullptr);
  e.db.assign(n_linear, nullptr);
  for (int l = 0; l < n_linear; ++l) {
    const size_t nw = (size_t)widths[l] * widths[l + 1];
    GL_CUDA(h, cudaMalloc(&e.dW[l], nw * sizeof(float)));
    GL_CUDA(h, cudaMalloc(&e.db[l], (size_t)widths[l + 1] * sizeof(float)));
    GL_CUDA(h, cudaMemcpy(e.dW[l], W_host[l], nw * sizeof(float), cudaMemcpyHostToDevice));
    GL_CUDA(h, cudaMemcpy(e.db[l], b_host[l], (size_t)widths[l + 1] * sizeof(float), cudaMemcpyHostToDevice));
  }
  e.ready = true;
  return GLASDI_OK;
}

int glasdi_encode(glasdi_handle* h, const float* U, int64_t rows, float* Z_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!h->enc.ready) return set_error(h, GLASDI_ERR_NOT_READY, "encoder weights not set (glasdi_set_encoder)");
  if (rows < 0 || (rows > 0 && (!U || !Z_out))) return set_error(h, GLASDI_ERR_INVALID, "encode: bad arguments");
  if (rows == 0) return GLASDI_OK;
  DeviceGuard device_guard(h->device);
  return launch_encode_simt(h, U, rows, Z_out, (cudaStream_t)stream);
}

int glasdi_sindy_fit_backward(glasdi_handle* h, const float* Z, const float* coefs, const float* grad_loss_sindy,
                              const float* grad_loss_coef, int B, int T, int n, int fd_id, double dt, int norm_order,
                              float* grad_Z_out, void* stream) {
  if (!h) return GLASDI_ERR_INVALID;
  if (!Z || !coefs || !grad_Z_out) return set_error(h, GLASDI_ERR_INVALID, "sindy_fit_backward: null pointer");
  DeviceGuard device_guard(h->device);
  return launch_sindy_fit_backward(h, Z, coefs, grad_loss_sindy, grad_loss_coef, B, T, n, fd_id, dt, norm_order,
                                   grad_Z_out, (cudaStream_t)stream);
}

}  // extern "C"
