/* crnn_pfr.h -- C ABI of the B200-native batched CRNN plug-flow-reactor surrogate.
 *
 * The reference (CHOIHSpotato/n-hexane-pyrolysis-surrogate-reactor-model) has no FFI: its surrogate path
 * is a set of Python call seams.  Each entry point below names the seam it replaces (paths relative to
 * the reference checkout).  Conventions:
 *   - every data pointer is DEVICE memory owned by the caller; the library allocates only what its own
 *     handles hold (uploaded parameters); nothing here synchronises the device unless stated;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *   - per-condition arrays are SoA with the condition index fastest: x[n], grid[801][n], y[9][n];
 *   - return value: PFR_OK or a negative PFR_E* code; per-condition solver outcomes go to status[n].
 * One host thread per GPU.  Library state is kept per device (tables, work-queue counters, auxiliary streams); handles that own
 * device memory (pfr_mlp_t, pfr_mlp_trainer_t, pfr_sweep_t) are bound to the device current at creation, and a call that arrives
 * with another device current returns PFR_EINVAL.  A crnn_model_t holds host memory only and may be used on any device.
 */
#ifndef CRNN_PFR_H
#define CRNN_PFR_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFR_OK 0
#define PFR_EINVAL (-1)     /* bad argument */
#define PFR_ECUDA (-2)      /* CUDA runtime error (pfr_last_cuda_error() has the text) */
#define PFR_EWORKSPACE (-3) /* workspace too small */

#define PFR_NS 9      /* species: H2 CH4 C2H4 C2H6 C3H6 C4H8-1 NC6H14 C4H10 C5H10-1 */
#define PFR_NR 9      /* pseudo-reactions */
#define PFR_NKNOTS 801

/* per-condition solver status (mirrors torchdiffeq's asserts) */
#define PFR_ST_OK 0
#define PFR_ST_MAXSTEPS 1
#define PFR_ST_NONFINITE 2
#define PFR_ST_UNDERFLOW 3
#define PFR_ST_STIFF 4      /* explicit fast paths only (PFR_METHOD_BS23, _BS23_WARP, _DP54, _DP54_WARP): the explicit fast path met a stiff knot interval; integrate this condition with
                            * PFR_METHOD_ROS3 / PFR_METHOD_RODAS4 (Surrogate does so automatically) */

#define PFR_FLAG_DENSE_RAW 1

#define PFR_METHOD_RODAS4 0     /* adaptive Rosenbrock, knot-aware (the product integrator): 3 lanes per condition */
#define PFR_METHOD_DOPRI5 1     /* torchdiffeq-semantics dopri5 (reference-behaviour mode) */
#define PFR_METHOD_RODAS4_TPC 2 /* the same Rosenbrock method, one thread per condition (LU parked in shared memory) */
#define PFR_METHOD_ROS3 3       /* 3-stage L-stable Rosenbrock of order 3(2), 2 right-hand sides per step, same kernel structure
                                 * as PFR_METHOD_RODAS4: cheaper per knot-limited step, needs a ~10x tighter tolerance for the
                                 * same accuracy (work-precision table) */
#define PFR_METHOD_BS23 4       /* explicit Bogacki-Shampine 3(2), one thread per condition, knot-limited (tgrid required): the
                                 * fast path for the non-stiff knot intervals of the coupled (Eon) path; a condition that turns
                                 * out stiff stops with PFR_ST_STIFF */
#define PFR_METHOD_DP54 5       /* explicit Dormand-Prince 5(4), one thread per condition, free stepping to t_end at T = T0 (tgrid,
                                 * Tprof, idx_end, y_dense must be NULL): the fast path of the isothermal (Eoff) sweep; our own
                                 * step-size controller, NOT torchdiffeq's (that is PFR_METHOD_DOPRI5); PFR_ST_STIFF as above */

#define PFR_METHOD_BS23_WARP 6  /* PFR_METHOD_BS23 with ONE CONDITION PER WARP (lane = species = reaction; float64, tgrid required): the
                                 * latency-oriented mapping for batches of a few hundred conditions with dense output -- the forward
                                 * pass of the training step; same method and controller, dot products summed in another order */
#define PFR_METHOD_DP54_WARP 8  /* Dormand-Prince 5(4) with ONE CONDITION PER WARP, free stepping along tgrid's time span at T = T0 (tgrid required,
                                 * Tprof must be NULL, float64), y_dense from the method's 4th-order continuous extension at the 801 knots -- what
                                 * torchdiffeq's dopri5 does for the reference's isothermal trainers (...WIDE_Eoff_surrogate_model_training.py:383):
                                 * the forward pass of the isothermal training step, ~30-60 steps instead of 800; PFR_ST_STIFF as for DP54 */
#define PFR_METHOD_TAYLOR4 7    /* explicit Taylor-series method of order 4 (3), one thread per condition, knot-limited (tgrid required): the time
                                 * derivatives of f = W exp(kT + nu^T ln y) come from the series recurrences of log and exp, so a step costs ONE
                                 * set of logarithms / exponentials and eight 9x9 mat-vecs instead of BS23's three right-hand sides; steps are cut
                                 * where a species crosses the lower state clamp; PFR_ST_STIFF as for BS23.  stats[2] counts coefficient
                                 * evaluations (one per attempt) */

typedef struct crnn_model* crnn_model_t;
typedef struct pfr_mlp* pfr_mlp_t;
typedef struct pfr_mlp_trainer* pfr_mlp_trainer_t;
typedef struct pfr_sweep* pfr_sweep_t;

int pfr_version(void);
const char* pfr_status_string(int code);
const char* pfr_last_cuda_error(void);
/* number of kernels this library has launched in this process (monotone counter) */
unsigned long long pfr_launch_count(void);

/* CRNN parameters: `parameters[-1]` of a training history
 *   load_npz_parameters                      SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:223-230
 * w_in[11][9] (rows 0-8 reaction orders, 9 Ea [kcal/mol], 10 b), w_b[9] = ln A, w_out[9][9], HOST pointers.
 * clamps = {lb, ub, zlo, zhi, dulo, duhi} or NULL for the inference defaults {1e-6, 60, -30, 30, -1e5, 1e5}
 *   (...Eoff_single_model.py:123-129; the WIDE trainer uses zlo/zhi = -/+10). */
int crnn_model_create(const float* w_in, const float* w_b, const float* w_out, const double* clamps, crnn_model_t* out);
int crnn_model_destroy(crnn_model_t m);
/* New parameters for an existing handle (the training loop: one call per optimisation step, ParameterConverter(p) of
 * SURROGATE_MODEL_TRAINING/WIDE_Eoff_surrogate_model_training.py:194-228 evaluated on the host).  Parameters reach the kernels by
 * value at launch time: launches already enqueued keep the old ones, later launches see the new ones; clamps are unchanged. */
int crnn_model_update(crnn_model_t m, const float* w_in, const float* w_b, const float* w_out);

/* 512-wide predictor MLP (nn.Linear layout [out][in], HOST pointers) + its .pkl output scaler
 *   MultiLayerPerceptron_time / MLP_Time / MLP_Temp   ...Eoff_single_model.py:192-208, ...Eon_single_model.py:94-128
 * in_dim 4: (T, P, L, u0) time predictor; in_dim 2: (T, P) temperature predictor.
 * in_lo/in_hi [in_dim]: input scaling bounds (...Eoff_single_model.py:282-283). */
int pfr_mlp_create(int in_dim, const float* const weights[4], const float* const biases[4], double out_min,
                   double out_max, const double* in_lo, const double* in_hi, pfr_mlp_t* out);
int pfr_mlp_destroy(pfr_mlp_t mlp);
/* Arithmetic of the three 512-wide layers: PFR_MLP_FP32 = FP32 FFMA, one accumulator per output, k ascending (default);
 * PFR_MLP_TF32X3 = tcgen05 tensor cores with an error-compensated 3xTF32 split (float32 accumulation in TMEM);
 * PFR_MLP_F16X3 = the same three-product split with FLOAT16 operand pairs (same 11-bit significands, half the bytes and half
 * the instructions; the low part is scaled by 2^11 so that it stays a normal float16).  Needs every fc2..fc4 weight inside
 * float16's range, else pfr_mlp_set_mode returns PFR_EINVAL; activations beyond 65 504 become NaN in the grid. */
#define PFR_MLP_FP32 0
#define PFR_MLP_TF32X3 1
#define PFR_MLP_F16X3 2
int pfr_mlp_set_mode(pfr_mlp_t mlp, int mode);
/* scratch needed by the calls below for batches processed `chunk` conditions at a time (chunk<=0: default) */
size_t pfr_mlp_workspace_bytes(int n, int chunk);

/* calculate_spec_conc_0_list / build_spec_conc_0_list   ...Eoff_single_model.py:45-55, ...Eon_single_model.py:41-50
 * c0[n]: inlet n-hexane concentration (mol/m3); P in Pa. */
int pfr_inlet_concentration(const float* T, const float* P, int n, float* c0, void* stream);

/* model_time(x) -> un-scale -> cat(t0, .) -> enforce_strict   ...Eoff_single_model.py:296-318
 * predict_time_profile(T,P,L,u)                                ...Eon_single_model.py:265-273
 * L/u0 NULL: the full-length grid at L = 1.0 m, u0 = 2.5 m/s (...Eon_single_model.py:309).
 * tgrid [801][n] and/or t_end [n] (= tgrid[800]); either may be NULL.
 * raw != 0: skip un-scaling and enforce_strict, rows 1..800 receive the bare network output (row 0 = 0). */
int pfr_time_grid(pfr_mlp_t mlp, const float* T, const float* P, const float* L, const float* u0, int n, float* tgrid,
                  float* t_end, int raw, void* workspace, size_t workspace_bytes, int chunk, void* stream);

/* predict_temp_profile(T,P) = [T0, mlp*(max-min)+min]   ...Eon_single_model.py:257-263 ; Tprof [801][n] */
int pfr_temp_profile(pfr_mlp_t mlp, const float* T, const float* P, int n, float* Tprof, int raw, void* workspace,
                     size_t workspace_bytes, int chunk, void* stream);

/* idx_cut = argmin |t_full - end_time|   ...Eon_single_model.py:348-350 ; idx [n] */
int pfr_idx_cut(const float* t_full, const float* t_end, int n, int* idx, void* stream);

/* CRNNFunc.forward at given temperatures (first-kernel parity hook)   ...Eoff_single_model.py:135-153
 * precision 64: T[n], u[9][n], du[9][n] double; precision 32: float. */
int pfr_rhs(crnn_model_t m, int n, const void* T, const void* u, void* du, int precision, void* stream);

/* Trainer.predict_n_ode / crnn_predict = clamp(odeint(CRNNFunc, u0, t, ...))   ...Eoff_single_model.py:175-186,
 *                                                                             ...Eon_single_model.py:153-156
 * T0[n], c0[n]; tgrid[801][n] or NULL (then t_end[n] gives the final time and Tprof must be NULL);
 * Tprof[801][n] or NULL (isothermal, T = T0); idx_end[n] or NULL (= 800): knot whose state is the outlet.
 * y_out[9][n] (double for precision 64, float for 32), clamped to [lb, ub]; y_dense[801][9][n] or NULL
 * (requires tgrid); status[n]; stats[3][n] = accepted steps, rejected steps, RHS evaluations (or NULL).
 * perm[n] or NULL: thread j integrates condition perm[j] -- a cost-sorted order keeps the lanes of a warp in
 * step with each other; every array is still indexed by the condition, so outputs need no un-permuting.
 * flags: PFR_FLAG_DENSE_RAW leaves y_dense unclamped (the training step needs the raw knot states).
 * method: PFR_METHOD_* above.  The explicit fast paths (BS23: tgrid required; DP54: t_end only, no Tprof / idx_end / y_dense)
 * stop a condition that turns out stiff with status PFR_ST_STIFF; the caller integrates those conditions again with
 * PFR_METHOD_ROS3 / PFR_METHOD_RODAS4 (the host mirror does: Surrogate.integrate, stiff_fallback).  stats of BS23 count one
 * zero-length entry step per condition: rhs = 3 (attempts + 1); DP54: rhs = 6 attempts + 1. */
int pfr_integrate(crnn_model_t m, int method, int precision, int n, const float* T0, const float* c0,
                  const float* tgrid, const float* Tprof, const float* t_end, const int* idx_end, const int* perm,
                  double rtol, double atol, int max_steps, int flags, void* y_out, void* y_dense, int* status,
                  int* stats, void* stream);

/* Device-side hand-over of the explicit fast paths: collects the conditions that pfr_integrate(PFR_METHOD_BS23 | PFR_METHOD_DP54) left
 * with status PFR_ST_STIFF into a list on the device and integrates them again with `method` (PFR_METHOD_ROS3 | PFR_METHOD_RODAS4),
 * overwriting their y_out / y_dense / status / stats.  Same arrays and meaning as pfr_integrate; nothing waits for the host.
 * scratch: n + 1 ints of device memory, scratch[n] = number of conditions handed over. */
int pfr_stiff_fallback(crnn_model_t m, int method, int precision, int n, const float* T0, const float* c0, const float* tgrid,
                       const float* Tprof, const float* t_end, const int* idx_end, double rtol, double atol, int max_steps, int flags,
                       void* y_out, void* y_dense, int* status, int* stats, int* scratch, void* stream);

/* ---- the whole sweep as one call ----------------------------------------------------------------------------------------
 * main() of SURROGATE_MODEL/surrogate_model_Eoff_single_model.py:259-369 (temp_mlp NULL: isothermal sweep, T = T0, outlet at the
 * last knot of the (T, P, L, u0) time grid) and of ...Eon_single_model.py:279-354 (coupled sweep: temperature profile and
 * full-length grid at (T, P, 1.0 m, 2.5 m/s), outlet = state at knot idx_cut = argmin |t_full - t_short[-1]|; L = u0 = NULL:
 * the outlet is the last knot of the full-length grid, as for sampling_case_2D.csv).
 * The handle owns every intermediate buffer for up to n_max conditions (pfr_sweep_device_bytes): inputs in visiting order,
 * c0, the two [801][n] grids, outlet knots / times, three MLP workspaces, two side streams for the independent MLP passes.
 * pfr_sweep_run is asynchronous on `stream` and never waits for the host:
 *   visiting order (counting sort by a cost proxy known before the MLPs run: L / u0 resp. T0, expensive first; the grids are
 *   then WRITTEN in the order the integrator reads them) -> pfr_inlet_concentration -> pfr_temp_profile / pfr_time_grid x2 ->
 *   pfr_idx_cut -> integrator -> Rosenbrock fallback over the device-side list of conditions the explicit fast path flagged
 *   stiff -> results scattered to the caller's order.
 * T, P, L, u0 [n], y_out [9][n] (precision 64: double, 32: float), status [n], stats [3][n] or NULL (a condition that went
 * through the fallback reports the fallback's counters), idx_cut_out [n] / t_end_out [n] or NULL -- all in the caller's order;
 * stiff_count_out: one int on the device or NULL.
 * method: coupled PFR_METHOD_BS23 | ROS3 | RODAS4; isothermal PFR_METHOD_DP54 | ROS3 | RODAS4. */
#define PFR_SWEEP_NO_ORDER 1     /* flags: visit the conditions in the caller's order */
#define PFR_SWEEP_NO_FALLBACK 2  /* flags: leave conditions flagged PFR_ST_STIFF as they are */
int pfr_sweep_create(crnn_model_t crnn, pfr_mlp_t time_mlp, pfr_mlp_t temp_mlp, int n_max, pfr_sweep_t* out);
int pfr_sweep_destroy(pfr_sweep_t s);
size_t pfr_sweep_device_bytes(int n_max, int energy_on);
int pfr_sweep_run(pfr_sweep_t s, const float* T, const float* P, const float* L, const float* u0, int n, int method, int precision,
                  double rtol, double atol, int max_steps, int flags, void* y_out, int* status, int* stats, int* idx_cut_out,
                  float* t_end_out, int* stiff_count_out, void* stream);
/* conditions the last run handed to the Rosenbrock fallback: written to stiff_count_out (one int on the device, caller-owned) if
 * given, else kept by the handle and read with this call (which synchronises the device) */
int pfr_sweep_stiff_count(pfr_sweep_t s, int* count);
/* mean device time of the integrator launch over the runs since the previous call of this function (at most the last 16; the
 * last run again if there was none), between event pairs the handle records around it; waits for those launches to finish */
int pfr_sweep_integrator_ms(pfr_sweep_t s, float* ms);

/* Loss and gradient of one training step for a batch of conditions
 *   loss = Trainer.loss_n_ode(p, i_exp); loss.backward()   SURROGATE_MODEL_TRAINING/WIDE_Eoff_surrogate_model_training.py:387-396,414-416
 * y_knots[801][9][n]: raw knot states of the forward pass (pfr_integrate with y_dense and PFR_FLAG_DENSE_RAW, double);
 * ref[801][7][n] labels (mol/m3), yscale[7][n] (...:105).  loss[n]: per-condition MSE over 7 x 801 points;
 * grad[189][n]: per-condition d loss / d (w_in[11][9] | w_b[9] | w_out[9][9]) by the continuous adjoint, RK4 with
 * |substeps| steps per knot interval; substeps > 0: one condition per warp (the product kernel), < 0: one per thread
 * (cross-check).  The model's clamps are those given to crnn_model_create. */
int pfr_loss_grad(crnn_model_t m, int n, const float* T0, const float* tgrid, const float* Tprof, const double* y_knots,
                  const float* ref, const float* yscale, int substeps, double* loss, double* grad, void* stream);
/* The same loss and gradient (RK4 adjoint, `substeps` > 0 per knot interval) in three kernels: node quantities and the 9 x 9
 * Jacobian transposes of all conditions x intervals at once, the sequential adjoint walk (one mat-vec per stage, one warp per
 * condition), the parameter-gradient quadrature in parallel.  4-5x faster than pfr_loss_grad for a training batch of a few
 * hundred conditions; needs pfr_loss_grad_workspace_bytes(n, substeps) of device memory (3.8 MB per condition at 2 sub-steps). */
size_t pfr_loss_grad_workspace_bytes(int n, int substeps);
int pfr_loss_grad_staged(crnn_model_t m, int n, const float* T0, const float* tgrid, const float* Tprof, const double* y_knots,
                         const float* ref, const float* yscale, int substeps, double* loss, double* grad, void* workspace,
                         size_t workspace_bytes, void* stream);
/* out[r] = sum_i x[r][i] with a fixed summation tree (deterministic reduction of per-condition gradients) */
int pfr_reduce_rows(const double* x, int rows, int n, double* out, void* stream);
/* the same over the columns with status[i] == 0 only (conditions whose forward integration succeeded); out[rows] = their number.
 * The training step reduces [grad(189) | loss] this way, so a failed trajectory enters neither the gradient nor the mean. */
int pfr_reduce_rows_ok(const double* x, int rows, int n, const int* status, double* out, void* stream);

/* Accuracy numbers of the reference's per-case / per-species CSV, all conditions at once
 * (...Eoff_single_model.py:384-480, ...Eon_single_model.py:381-463).
 * dense [801][9][n] (precision 64: double, 32: float) as pfr_integrate writes it; label [801][7][n] float at the same
 * knots (mol/m3); idx_end [n] or NULL: last knot used (Eon trim, knots 1..idx_end; NULL = 800); abs_den != 0 divides
 * the relative errors by |ref| + 1e-5 (Eon script) instead of ref + 1e-5 (Eoff script).
 * out [8][7][n] double: RMSE_final, NRMSE_final, RelError_final %, RMSE_time_avg, NRMSE_time_avg, RelError_time_avg %,
 * FCD, Max_Norm. */
int pfr_accuracy(const void* dense, int precision, const float* label, const int* idx_end, int n, int abs_den, double* out,
                 void* stream);

/* ---- predictor-MLP training ----------------------------------------------------------------------
 * One optimisation step of  TEMP_PRED_MODEL_TRAINING/temp_profile_model_training_2D.py:145-160  /
 * TIME_PRED_MODEL_TRAINING/time_profile_model_training_4D.py:174-190 :
 *   outputs = model(images); loss = nn.MSELoss()(outputs, labels); optimizer.zero_grad(); loss.backward(); optimizer.step()
 * with torch.optim.Adam(betas, eps; no weight decay) on the 2|4 -> 512 -> 512 -> 512 -> 800 ReLU network (nn.Linear layout:
 * weights[l] is [out][in] row-major, biases[l] [out], host pointers at creation).  The trainer owns parameters, Adam moments
 * and activations on the device. */
int pfr_mlp_trainer_create(int in_dim, const float* const weights[4], const float* const biases[4], pfr_mlp_trainer_t* out);
int pfr_mlp_trainer_destroy(pfr_mlp_trainer_t t);
/* x [B][in_dim], y [B][800] (already scaled as the Dataset classes do), B <= 32 (the scripts' batch size), device pointers;
 * loss: one float on the device (MSE before the update).  lr is this step's learning rate (StepLR lives on the host). */
int pfr_mlp_trainer_step(pfr_mlp_trainer_t t, const float* x, const float* y, int B, double lr, double beta1, double beta2,
                         double eps, float* loss, void* stream);
/* model.eval() forward: x [n][in_dim] -> out [n][800]; and the MSE of one batch without an update (validation loop) */
int pfr_mlp_trainer_forward(pfr_mlp_trainer_t t, const float* x, int n, float* out, void* stream);
int pfr_mlp_trainer_loss(pfr_mlp_trainer_t t, const float* x, const float* y, int B, float* loss, void* stream);
/* model.state_dict(): copies the current parameters to host arrays of the creation shapes (synchronises the device) */
int pfr_mlp_trainer_read(pfr_mlp_trainer_t t, float* const weights[4], float* const biases[4]);

/* Parity hook for the table-driven double-precision log (kind 0, x positive normal) / exp (kind 1, |x| < 700)
 * used inside the integrators in place of torch.log / torch.exp of CRNNFunc.forward (...Eoff_single_model.py:139,151).
 * kind 2 = the exponential as the explicit integrators evaluate it (exponents carried in units of ln2 / 256; the hook scales x),
 * kind 3 / 4 = the latency-oriented log / exp of the warp-per-condition kernels.  x[n] -> y[n], device pointers. */
int pfr_fastmath(int kind, int n, const double* x, double* y, void* stream);

/* Pipe micro-benchmarks used as roofline denominators (synchronous; a few ms each).
 * out[0] FP32 FFMA flop/s, out[1] FP64 DFMA flop/s, out[2] MUFU.EX2 op/s, out[3] SM clock seen (Hz, from clock64). */
int pfr_measure_peaks(double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* CRNN_PFR_H */
