/*
 * glasdi_b200.h -- C ABI of the B200-native (sm_100a) GPLaSDI greedy-sampling / SINDy-calibration
 * hot path.  This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 *
 * Conventions
 *   - every entry point returns an int status (GLASDI_OK == 0); no exceptions cross the boundary;
 *     glasdi_last_error() gives a human-readable message for the last non-zero status of a handle.
 *   - unless marked HOST, pointers are caller-owned DEVICE pointers on the handle's device
 *     (e.g. torch CUDA tensors' data_ptr()); the library never frees or keeps them.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *     re-entrant per handle (one handle = one device + one workspace; do not share a handle
 *     between host threads concurrently).
 *   - there is NO CPU fallback: if no sm_100 device is present glasdi_create fails.
 *
 * Each entry point names the reference interface (LLNL/GPLaSDI, file:line) it replaces.
 * INTEGRATION.md shows the ctypes stubs a reference maintainer would add.
 */
#ifndef GLASDI_B200_H
#define GLASDI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLASDI_ABI_VERSION 1

typedef struct glasdi_handle glasdi_handle;

enum glasdi_status {
  GLASDI_OK = 0,
  GLASDI_ERR_INVALID = 1,      /* bad argument / unsupported shape */
  GLASDI_ERR_CUDA = 2,         /* CUDA runtime error (message has the cudaError string) */
  GLASDI_ERR_NO_DEVICE = 3,    /* no sm_100 device: there is no CPU fallback */
  GLASDI_ERR_NOT_READY = 4,    /* e.g. greedy step before glasdi_set_decoder */
  GLASDI_ERR_NUMERIC = 5,      /* fp16 split overflow of activations, all-zero std, singular Gram */
  GLASDI_ERR_OOM = 6
};

/* activation ids = names of reference src/lasdi/latent_space.py:5-28 (`act_dict`), torch defaults */
enum glasdi_act {
  GLASDI_ACT_SIGMOID = 0, GLASDI_ACT_TANH = 1, GLASDI_ACT_SOFTPLUS = 2, GLASDI_ACT_RELU = 3,
  GLASDI_ACT_ELU = 4, GLASDI_ACT_SILU = 5, GLASDI_ACT_GELU = 6, GLASDI_ACT_IDENTITY = 7,
  /* the remaining parameter-free entries of act_dict (torch default arguments): Hardshrink(0.5), Hardsigmoid, Hardtanh(-1, 1),
     Hardswish, LeakyReLU(0.01), LogSigmoid, ReLU6, SELU, CELU(1), Mish, Softshrink(0.5), Tanhshrink.  Not covered: 'threshold'
     (two parameters), 'PReLU' (learnable), 'RReLU' (random in training mode), 'multihead' (not an activation) */
  GLASDI_ACT_HARDSHRINK = 8, GLASDI_ACT_HARDSIGMOID = 9, GLASDI_ACT_HARDTANH = 10, GLASDI_ACT_HARDSWISH = 11,
  GLASDI_ACT_LEAKYRELU = 12, GLASDI_ACT_LOGSIGMOID = 13, GLASDI_ACT_RELU6 = 14, GLASDI_ACT_SELU = 15,
  GLASDI_ACT_CELU = 16, GLASDI_ACT_MISH = 17, GLASDI_ACT_SOFTSHRINK = 18, GLASDI_ACT_TANHSHRINK = 19,
  GLASDI_ACT_COUNT = 20
};

/* finite-difference scheme ids = keys of reference src/lasdi/fd.py:219-222 (`FDdict`) */
enum glasdi_fd { GLASDI_FD_SBP12 = 0, GLASDI_FD_SBP24 = 1, GLASDI_FD_SBP36 = 2, GLASDI_FD_SBP48 = 3 };

/* latent ODE integrators for the affine SINDy system (reference: scipy odeint/LSODA, sindy.py:115) */
enum glasdi_integrator {
  GLASDI_INT_EXPM = 0,   /* fp64 exact affine propagator expm([[A,b],[0,0]] dt), stepped in fp64 */
  GLASDI_INT_RK4 = 1     /* fp64 classical RK4 with `substeps` sub-steps per output step */
};

/* options for glasdi_set_option */
enum glasdi_option {
  GLASDI_OPT_SPLIT_PASSES = 0,  /* 1, 2 or 3 fp16 split-product passes of the last decoder layer
                                   (3 = fp32-faithful default; see DESIGN.md "precision") */
  GLASDI_OPT_INTEGRATOR = 1,    /* enum glasdi_integrator (default EXPM) */
  GLASDI_OPT_RK4_SUBSTEPS = 2,  /* default 1 */
  GLASDI_OPT_MAX_CTAS = 3,      /* cap on persistent CTAs (0 = one per SM); for tests */
  GLASDI_OPT_DECODE_KERNEL = 5, /* only 1 is accepted: the CTA-pair tcgen05 kernel (cta_group::2, 256x256 tiles); the
                                   one-CTA-per-SM kernel of round 1 (value 0) was retired */
  GLASDI_OPT_STAGE_TIMING = 4,  /* 1: bracket the rollout and decode kernels of the greedy step with
                                   CUDA events on the caller's stream (read with glasdi_stage_times) */
  GLASDI_OPT_MODE = 6,          /* enum glasdi_mode (default SCREENED_GRAM) */
  GLASDI_OPT_SCREEN_MARGIN = 7, /* m in 4..16 (default 8): screened mode re-evaluates every (t, dof) whose
                                   single-pass std is within delta = 2^-m of the parameter's screened maximum */
  GLASDI_OPT_WIDE_TENSOR = 8,   /* decoders outside the fused tcgen05 kernels (last-layer input width > 112): 2 (default)
                                   last front layer AND last layer + variance as tcgen05 GEMMs (decode_wide.cu: one fp16
                                   screening pass in the screened modes, three fp16-split passes in the direct mode and
                                   for field outputs), the narrow layers in front of them as a register-to-register
                                   mma.sync chain with split operands (screening pass) or fp32 CUDA-core GEMMs (exact
                                   passes); 3: as 2 with the screening pass's narrow layers on the CUDA cores; 1: only the
                                   last layer on tcgen05; 0: fp32 CUDA-core GEMMs throughout.  Same results. */
  GLASDI_OPT_GRAM_KERNEL = 9,   /* covariance GEMM of decoders with ONE front layer and sigmoid / tanh / softplus: 1 (default)
                                   gram2_kernel -- thread = hidden unit, centred activations h(a_pilot + da) - h(a_pilot) from a
                                   5th-order Taylor polynomial in da (no MUFU), exact form for sample blocks outside the radius;
                                   0: gram_kernel (round-1 producers).  Screening only: results are the exact re-evaluation's */
  GLASDI_OPT_TAYLOR_RADIUS = 10, /* radius of that polynomial: bound on |da| in units of 1/256 (default 256 = 1.0; 0 = never) */
  GLASDI_OPT_SCREEN_BOUND = 12, /* covariance-form screening: number of standard deviations (default 6; 0 = off) of the A-PRIORI
                                   bound on the fp16 rounding error of sigma^2 = sum_ij C_ij w_i w_j that is added to the
                                   screened variance of every (t, dof): beta |diag C[p,t]|_2 |w_dof o w_dof|_2, beta =
                                   sigmas * 1.15 * 2^-11.  A (t, dof) that reaches (1 - delta/2)^2 of its parameter's
                                   screened maximum only WITH that bound is re-evaluated exactly as well: a dof whose
                                   quadratic form cancels badly cannot be screened out unseen. */
  GLASDI_OPT_MT_SEGMENT_LOG2 = 13, /* glasdi_normal_mt19937: the MT19937 word stream is cut into segments of 2^value words
                                   (default 18; 11..24) that different CTAs generate at the same time, each from a state
                                   reached by jump-ahead (t^(c 2^value) mod the generator's characteristic polynomial,
                                   computed once per handle); 0 = the whole stream from one CTA.  Same bits either way. */
  GLASDI_OPT_SINDY_LIBRARY = 11 /* enum glasdi_library (default POLY1 = the affine library of src/lasdi): which candidate
                                   functions the coefficient rows of glasdi_rollout / glasdi_greedy_step multiply */
};

/* how glasdi_greedy_step / glasdi_decode_max_std obtain max_std[p] (same result within the fp32 tolerance) */
enum glasdi_mode {
  GLASDI_MODE_DIRECT = 0,    /* GLASDI_OPT_SPLIT_PASSES fp16 split-product passes over the whole ensemble */
  GLASDI_MODE_SCREENED_GRAM = 2, /* default.  Screening in covariance form: per (parameter, time step) the K x K
                                covariance of the last hidden layer over the samples (one tcgen05 GEMM whose
                                contraction runs over the samples), then var[dof] = w_dof^T C w_dof for every dof
                                as a second GEMM against the packed products w_i w_j.  5x fewer FLOP than decoding
                                the ensemble and no accumulator drain; approximate (fp16, ~1e-4), so the same
                                exact re-evaluation + deviation check as GLASDI_MODE_SCREENED follows. */
  GLASDI_MODE_SCREENED = 1   /* ONE fp16 pass over the whole ensemble writes an approximate variance field;
                                the (t, dof) within the margin of each parameter's maximum are then re-evaluated
                                on the CUDA cores exactly as the reference computes them (fp32 decode, fp64
                                sample sums).  A screening deviation above delta/4 at any candidate, or a
                                candidate-buffer overflow, is reported as GLASDI_ERR_NUMERIC by
                                glasdi_check_numeric -- never silently accepted. */
};

/* SINDy candidate library Theta(z) of the latent ODE dz/dt = Theta(z) R, R = coefficients [|Theta|, n] row-major.
 * Replaces: the four right-hand sides of `simulate_sindy` (reference legacy/Vlasov1D1V/utils/utils_sindy.py:91-170;
 * library size :12-27, construction order :31-64): Theta = [1, z_0..z_{n-1}, z_i z_j (i <= j, i slowest) if poly_order = 2,
 * sin z_0..sin z_{n-1} if sine_term].  POLY1 is the library of src/lasdi/latent_dynamics/sindy.py:104-117 (affine: exact
 * propagator or RK4 by GLASDI_OPT_INTEGRATOR); the other three are always integrated by fp64 RK4 whose sub-step count is
 * re-chosen at every output step from a row-sum bound L of the Jacobian at the current state (h L <= 1/16, at least
 * GLASDI_OPT_RK4_SUBSTEPS); a trajectory that needs more than 16384 sub-steps per output step (finite-time blow-up of a
 * quadratic system) or turns non-finite raises GLASDI_ERR_NUMERIC at the next glasdi_check_numeric. */
enum glasdi_library {
  GLASDI_LIB_POLY1 = 0, GLASDI_LIB_POLY1_SINE = 1, GLASDI_LIB_POLY2 = 2, GLASDI_LIB_POLY2_SINE = 3
};
/* |Theta| of a library for latent dimension n (= 1 + n + `lib_size(n, 2)` + n sines); -1 for an unknown id */
int glasdi_library_terms(int library, int n);

int glasdi_abi_version(void);
/* sha1 of the CUDA sources + this header the library was built from (gplasdi_b200/build.py::source_hash); the Python
 * loader compares it with the sources beside it and rebuilds, or refuses, a stale library */
const char* glasdi_source_hash(void);

/* lifetime ------------------------------------------------------------------------------------ */
int glasdi_create(glasdi_handle** out, int device);
int glasdi_destroy(glasdi_handle* h);
const char* glasdi_last_error(const glasdi_handle* h);
int glasdi_set_option(glasdi_handle* h, int option, int value);
int glasdi_get_option(const glasdi_handle* h, int option, int* value);
/* number of kernels this handle has launched so far (bench.py's `gpu_launches`) */
int64_t glasdi_launch_count(const glasdi_handle* h);

/* Device time of the stages of the LAST greedy step / decode_max_std call, from CUDA events recorded
 * on the caller's stream (requires GLASDI_OPT_STAGE_TIMING = 1; synchronises those events).
 * rollout_ms: latent rollout (or layout conversion) kernels; decode_ms: decode+variance kernels
 * (summed over parameter slabs); decode_launches: number of decode+variance launches. */
int glasdi_stage_times(glasdi_handle* h, double* rollout_ms, double* decode_ms, int* decode_launches);

/* Screened mode diagnostics of the LAST greedy call (synchronises `stream`): device time of candidate
 * selection + exact re-evaluation and of the second screening GEMM (quad_kernel; covariance form only) --
 * both 0 unless GLASDI_OPT_STAGE_TIMING; glasdi_stage_times' decode_ms is then the first screening kernel
 * alone --, number of re-evaluated (t, dof) candidates, and the largest relative deviation
 * |std_screen - std_exact| / std_exact seen at them. */
int glasdi_screen_stats(glasdi_handle* h, double* refine_ms, double* quad_ms, int64_t* n_candidates,
                        float* max_rel_dev, void* stream);

/* Number of (t, dof) of the LAST greedy call that became candidates through the a-priori bound alone
 * (GLASDI_OPT_SCREEN_BOUND); 0 for the other modes.  Synchronises `stream`. */
int glasdi_screen_bound_count(glasdi_handle* h, int64_t* n_bound_candidates, void* stream);

/* decoder weights ------------------------------------------------------------------------------
 * Replaces: the `autoencoder.decoder` nn.Module state consumed by get_fom_max_std
 *   (reference src/lasdi/gplasdi.py:66; MultiLayerPerceptron, src/lasdi/latent_space.py:65-171;
 *   re-read from checkpoint.pt at gplasdi.py:254 -> call again after every load_state_dict).
 * n_linear linear layers; widths[0..n_linear] = [n_z, hidden..., D]; W[l] HOST fp32
 * [widths[l+1], widths[l]] row-major (nn.Linear layout), b[l] HOST fp32 [widths[l+1]].
 * Packs the last layer into fp16 hi/lo UMMA tiles and uploads everything (synchronous). */
int glasdi_set_decoder(glasdi_handle* h, int n_linear, const int* widths,
                       const float* const* W_host, const float* const* b_host, int act_id);

/* latent rollout ------------------------------------------------------------------------------
 * Replaces: SINDy.simulate in the loops of gplasdi.py:20-21, :45-48, :262-266 and
 *   LatentDynamics.sample (latent_dynamics/__init__.py:39-54).
 * coefs fp64 [P, S, n(n+1)] (row-major flatten of the (n+1, n) lstsq solution, row 0 = constant,
 * sindy.py:72,80,112-113; [P, S, |Theta| n] under GLASDI_OPT_SINDY_LIBRARY, here and in glasdi_greedy_step); z0 fp32 [P, n] (one initial condition per parameter, shared by its S
 * samples; pass z0_per_sample != 0 for z0 [P, S, n]); uniform grid t_k = k*dt, k < T.
 * Z_out fp64 [P, S, T, n] in the reference's layout. */
int glasdi_rollout(glasdi_handle* h, const double* coefs, const float* z0, int z0_per_sample,
                   int P, int S, int n, int T, double dt, double* Z_out, void* stream);

/* The same rollout on an ARBITRARY increasing (or decreasing) time grid, as scipy's odeint accepts it in SINDy.simulate
 * (sindy.py:104-117): t_grid HOST fp64 [T], Z_out[.., k, :] = z(t_grid[k]), z(t_grid[0]) = z0.  Every library, the affine one
 * included, is integrated by fp64 RK4 whose sub-step count is re-chosen at every output step from the Jacobian bound
 * (h L <= 1/16); synchronises `stream` once (the grid is copied from a host temporary). */
int glasdi_rollout_grid(glasdi_handle* h, const double* coefs, const float* z0, int z0_per_sample,
                        int P, int S, int n, int T, const double* t_grid_host, double* Z_out, void* stream);

/* fused greedy step ---------------------------------------------------------------------------
 * Replaces: gplasdi.py:262-268 = the simulate double loop + get_fom_max_std(ae, Zis).
 * Rollout (fp64) -> MLP decode -> population std over the S samples (ddof = 0) per (t, dof)
 * -> max over (t, dof) -> max_std[p]; then the first strict maximum over p, starting from 0
 * (gplasdi.py:62-72).  Nothing but max_std[P] and the (index, value) pair is written.
 * max_std_out fp32 [P]; argmax_out int64[1] (LOCAL index in [0,P), -1 if no std exceeds 0 --
 * the reference raises UnboundLocalError there); maxval_out fp32[1]. */
int glasdi_greedy_step(glasdi_handle* h, const double* coefs, const float* z0,
                       int P, int S, int n, int T, double dt,
                       float* max_std_out, int64_t* argmax_out, float* maxval_out, void* stream);

/* Replaces: get_fom_max_std(autoencoder, Zis) (gplasdi.py:52-74) for caller-supplied
 * trajectories Zis fp64 [P, S, T, n] (cast to fp32 first, as `torch.Tensor(Zi)` does). */
int glasdi_decode_max_std(glasdi_handle* h, const double* Zis, int P, int S, int T, int n,
                          float* max_std_out, int64_t* argmax_out, float* maxval_out, void* stream);

/* Sample statistics of the decoded ensemble as FIELDS.
 * Replaces: `pred = autoencoder.decoder(torch.Tensor(Z)); pred.mean(0); pred.std(0)` of plot_prediction
 *   (src/lasdi/postprocess.py:32-34; examples/burgers1d.ipynb) for every parameter of Zis fp64 [P, S, T, n].
 * mean_out / std_out fp32 [P, T, D] (either may be NULL).  Three fp16 split passes on the tensor cores (fp32-faithful),
 * mean = decode(sample 0) + mean_s(x(s) - x(0)); the decoded ensemble is never written.  Wide decoders (last-layer input
 * width > 112) take the K-looped GEMMs of decode_wide.cu with the same three passes and a field epilogue. */
int glasdi_decode_stat_fields(glasdi_handle* h, const double* Zis, int P, int S, int T, int n, float* mean_out,
                              float* std_out, void* stream);

/* Replaces: autoencoder.decoder(Z) forward (latent_space.py:135-171; call sites gplasdi.py:66,182,
 * postprocess.py:32).  Z fp32 [rows, n_z] -> X fp32 [rows, D]. */
int glasdi_decode(glasdi_handle* h, const float* Z, int64_t rows, float* X_out, void* stream);

/* Synchronises `stream` and reports (then clears) numeric flags raised on the device by earlier
 * asynchronous calls: GLASDI_ERR_NUMERIC if hidden activations overflowed the fp16 split range. */
int glasdi_check_numeric(glasdi_handle* h, void* stream);

/* first strict maximum of vals[0..P) starting from 0 (gplasdi.py:62-72); idx_out = -1 if none */
int glasdi_argmax_first(glasdi_handle* h, const float* vals, int P, int64_t* idx_out, float* val_out,
                        void* stream);

/* multi-GPU max-loc key: (float_bits(val) << 32) | (0xFFFFFFFF - global_index); an integer MAX
 * all-reduce over ranks then selects the largest value and, on ties, the smallest global index
 * (= the reference's first-strict-maximum rule).  key_out int64[1]; a local idx of -1 packs to 0. */
int glasdi_pack_maxloc(glasdi_handle* h, const float* val, const int64_t* local_idx, int64_t index_offset,
                       int64_t* key_out, void* stream);

/* SINDy calibration ---------------------------------------------------------------------------
 * Replaces: SINDy.compute_time_derivative (sindy.py:89-102) with the SBP operators of fd.py:14-217.
 * Z fp32 [B, T, n] -> dZ fp32 [B, T, n] = (1/dt) * D_fd Z. */
int glasdi_fd_dzdt(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt,
                   float* dZ_out, void* stream);

/* Replaces: SINDy.calibrate (sindy.py:41-87): per series dZ/dt by the stencil, library [1 | Z],
 * least squares with LAPACK xGELSY's semantics (what torch.linalg.lstsq runs at sindy.py:72): column-pivoted
 * factorisation, effective rank from the incremental condition estimate against rcond = eps_fp32 * max(T, n + 1),
 * minimum-norm solution of the rank-truncated problem -- computed on the fp64 normal equations --, MSE(dZdt, [1|Z] R)
 * and ||R||_p (norm_order 1 or 2; 2 = 'fro').  coefs_out fp32 [B, n(n+1)] (row-major (n+1, n));
 * loss_sindy_out / loss_coef_out fp32 [B] (per series; the reference SUMS them over the batch);
 * info_out int32 [B]: number of truncated directions -- 0 = full rank, 1..n = rank-deficient series (finite
 * minimum-norm coefficients, as the reference returns), n + 1 = non-finite series (NaN coefficients). */
int glasdi_sindy_fit(glasdi_handle* h, const float* Z, int B, int T, int n, int fd_id, double dt,
                     int norm_order, float* coefs_out, float* loss_sindy_out, float* loss_coef_out,
                     int32_t* info_out, void* stream);

/* GP posterior and coefficient draws ---------------------------------------
 * Replaces: eval_gp (src/lasdi/gp.py:36-53) = GaussianProcessRegressor.predict(X, return_std=True) of every
 * coefficient's GP, for the kernel fit_gps builds (ConstantKernel * RBF, gp.py:28): k(x, x') = amplitude *
 * exp(-|x - x'|^2 / (2 length_scale^2)).  All arrays fp64 device pointers: X_train [n_train, n_param]; per GP k:
 * alpha [n_gp, n_train] (gp.alpha_), L [n_gp, n_train, n_train] (gp.L_, lower Cholesky factor, row-major),
 * amplitude / length_scale [n_gp], y_mean / y_std [n_gp] (gp._y_train_mean/_std; NULL = 0 / 1);
 * X_test [P, n_param] -> mean_out, std_out [P, n_gp].  n_train <= 400. */
int glasdi_gp_predict(glasdi_handle* h, const double* X_train, int n_train, int n_param, int n_gp, const double* alpha,
                      const double* L, const double* amplitude, const double* length_scale, const double* y_mean,
                      const double* y_std, const double* X_test, int P, double* mean_out, double* std_out,
                      void* stream);

/* State of numpy's legacy global generator, np.random.get_state(): ('MT19937', key, pos, has_gauss, cached_gaussian). */
typedef struct glasdi_mt19937_state {
  uint32_t key[624];
  int32_t pos;
  int32_t has_gauss;
  double cached_gaussian;
} glasdi_mt19937_state;

/* Replaces: the double loop `coef_samples[s, k] = np.random.normal(pred_mean[k], pred_std[k])` of sample_coefs
 * (gp.py:76-78) run for every test parameter in turn (gplasdi.py:260), with the SAME stream: out[p, s, k] =
 * mean[p, k] + std[p, k] * g, g = the next value of numpy's legacy_gauss (MT19937 -> 53-bit doubles -> polar
 * Box-Muller with rejection and a cached second value), in (p, s, k) order.  `state` (HOST, in/out) is the generator
 * state before / after the call, so a seeded run continues on the host exactly where the reference would.
 * mean, std fp64 device [P, n]; out fp64 device [P, S, n].  Values match numpy up to the rounding of log() (<= 1 ulp).
 * Synchronises `stream` before returning (the new state is a host result). */
int glasdi_normal_mt19937(glasdi_handle* h, glasdi_mt19937_state* state, const double* mean, const double* std, int P,
                          int S, int n, double* out, void* stream);

/* Encoder ----------------------------------------------------------------------------------------
 * Replaces: autoencoder.encoder(X) forward (latent_space.py:135-171 with reshape_index = 0, built at :220-222), as
 * called once per test parameter by initial_condition_latent (latent_space.py:30-63) -- here all
 * parameters in one batch.  Weights in nn.Linear layout like glasdi_set_decoder; widths[0] = number of dofs,
 * widths[n_linear] = n_z; any width.  U fp32 [rows, widths[0]] -> Z fp32 [rows, n_z]. */
int glasdi_set_encoder(glasdi_handle* h, int n_linear, const int* widths, const float* const* W_host,
                       const float* const* b_host, int act_id);
int glasdi_encode(glasdi_handle* h, const float* U, int64_t rows, float* Z_out, void* stream);

/* Backward of glasdi_sindy_fit.  Replaces: what torch autograd records through sparse.mm, linalg.lstsq, MSELoss and
 * norm when the reference's training loop differentiates the two losses (gplasdi.py:186-197 -> sindy.py:66-77; the
 * coefficients themselves are returned detached, sindy.py:80).
 * Z [B, T, n] and coefs [B, n(n+1)] as passed to / returned by the forward; grad_loss_sindy / grad_loss_coef fp32 [B]
 * upstream gradients of the per-series losses (NULL = 0); grad_Z_out fp32 [B, T, n] (rank-truncated series: the
 * derivative of the pseudo-inverse at constant rank, as torch's lstsq backward; NaN for a non-finite series).  Needs 2 * T * n * 4 bytes <= 200 KB of shared memory. */
int glasdi_sindy_fit_backward(glasdi_handle* h, const float* Z, const float* coefs, const float* grad_loss_sindy,
                              const float* grad_loss_coef, int B, int T, int n, int fd_id, double dt, int norm_order,
                              float* grad_Z_out, void* stream);

/* Training-mode autoencoder ---------------------------------------------------
 * Replaces: the graph torch autograd records through MultiLayerPerceptron.forward (src/lasdi/latent_space.py:157-164) in
 * the training step `Z = encoder(X); X_pred = decoder(Z)` (src/lasdi/gplasdi.py:180-181) and replays in `loss.backward()`
 * (:197).  Weights are DEVICE fp32 pointers (the optimiser's own parameters, nn.Linear layout): W[l] [widths[l+1],
 * widths[l]], b[l] [widths[l+1]]; the activation follows every layer but the last.
 * forward: X [rows, widths[0]] -> Y [rows, widths[n_linear]]; pre_out / post_out [rows, widths[1] + .. + widths[n_linear-1]]
 * (layer after layer, each [rows, width]) keep the hidden layers' pre- and post-activations for the backward. */
int glasdi_mlp_forward_train(glasdi_handle* h, int n_linear, const int* widths, const float* const* W_dev,
                             const float* const* b_dev, int act_id, const float* X, int64_t rows, float* pre_out,
                             float* post_out, float* Y_out, void* stream);
/* backward: dY [rows, widths[n_linear]] -> dX_out [rows, widths[0]] (NULL: not needed), dW_out[l], db_out[l] (device,
 * overwritten).  The activation derivative is applied to the incoming gradient in the loaders of the two products of a layer
 * (dW = dA^T in, dIn = dA W); fixed summation orders: bit-reproducible. */
int glasdi_mlp_backward(glasdi_handle* h, int n_linear, const int* widths, const float* const* W_dev, int act_id,
                        const float* X, const float* pre, const float* post, const float* dY, int64_t rows, float* dX_out,
                        float* const* dW_out, float* const* db_out, void* stream);
/* Replaces: `loss_ae = MSELoss()(X, X_pred)` (gplasdi.py:157,182) and its gradient: loss_out[0] = mean((A - B)^2) over n
 * elements (fp64 partial sums, fixed order) and, in the same pass, dA_out = 2 (A - B) / n (NULL: loss only). */
int glasdi_mse_loss_grad(glasdi_handle* h, const float* A, const float* B, int64_t n, float* loss_out, float* dA_out,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLASDI_B200_H */
