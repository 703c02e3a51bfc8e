/* speedy-b200: C ABI of the B200-native SPEEDY hot path (libspeedy_b200.so).
 *
 * Every entry point replaces one procedure of the reference's f2py-facing Fortran module `speedy_driver`
 * (generated from registry/templates/speedy_driver.f90.j2, result speedy.f90/speedy_driver.f90); the template
 * line of the procedure each one stands in for is cited.  Handles are opaque 64-bit integers exactly like the
 * reference's `integer(8)` containers (.j2:8-14,38-39).  Plain pointers and sizes only; no torch types.
 * Arrays cross the boundary in Fortran order with the registry dtype (complex128 / float64 / float32 / int32),
 * copies in both directions (docs/user_guide.rst:86-93).  Error codes: 0 ok, -1 state not initialised,
 * -2 diagnostics out of range (error_codes.f90:7-9).  INTEGRATION.md shows the Fortran iso_c_binding and the
 * Python ctypes bindings.
 */
#ifndef SPEEDY_B200_H
#define SPEEDY_B200_H
#include <stddef.h>
#include <stdint.h>

#include "spdy_registry.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- model state (.j2:216-258) ---------------------------------------------------------------------- */
int64_t spdy_modelstate_init(void);                              /* modelstate_init          .j2:216-223 */
void spdy_modelstate_init_sst_anom(int64_t state, int n_months); /* modelstate_init_sst_anom .j2:225-238 */
void spdy_modelstate_close(int64_t state);                       /* modelstate_close         .j2:240-258 */

/* ---- datetime and control containers (.j2:131-210) ---------------------------------------------------- */
int64_t spdy_create_datetime(int year, int month, int day, int hour, int minute); /* create_datetime .j2:163-186 */
void spdy_get_datetime(int64_t dt, int *ymdhm /* [5] */);                         /* get_datetime    .j2:188-202 */
void spdy_close_datetime(int64_t dt);                                             /* close_datetime  .j2:204-210 */
int64_t spdy_controlparams_init(int64_t start_dt, int64_t end_dt);                /* controlparams_init  .j2:131-148 */
void spdy_controlparams_close(int64_t control);                                   /* controlparams_close .j2:150-157 */

/* ---- model (.j2:29-125) --------------------------------------------------------------------------------- */
int spdy_init(int64_t state, int64_t control);                   /* init  .j2:29-41  -> initialization.f90:13-91 */
int spdy_step(int64_t state, int64_t control);                   /* step  .j2:43-55  -> speedy.f90:20-74          */
void spdy_parallel_step(const int64_t *states, const int64_t *controls, int *error_codes,
                        int n_members);                          /* parallel_step .j2:58-79                       */
int spdy_check(int64_t state);                                   /* check .j2:81-91  (time level 1)               */
void spdy_transform_spectral2grid(int64_t state);                /* .j2:94-103  -> prognostics.f90:125-154        */
void spdy_transform_grid2spectral(int64_t state);                /* .j2:105-114 -> prognostics.f90:157-176        */
void spdy_apply_grid_filter(int64_t state);                      /* .j2:116-125 -> prognostics.f90:180-219        */

/* ---- registry accessors: get_<v> / set_<v> / get_<v>_shape / is_array_<v> (.j2:264-334), var = enum spdy_var */
int spdy_get(int64_t state, int var, void *dst, size_t bytes);
int spdy_set(int64_t state, int var, const void *src, size_t bytes);
int spdy_shape(int64_t state, int var, int *dims /* [5] */, int *ndim);

/* ---- ensemble extensions (no reference counterpart; same semantics as repeated parallel_step) ------------- */
/* Advance all listed members `nsteps` steps without a host round trip per step; error_codes receives the first
 * non-zero code of each member (0 if none; a failing member is frozen at its failing step).  Returns the number of
 * failed members.  When the call returns, every registry variable holds what `nsteps` spdy_parallel_step calls leave,
 * bit for bit; in between -- where nothing can observe the state -- the column physics does not store the outputs that no
 * kernel reads and the next step overwrites, and the spectral steps skip coefficients outside the triangular truncation
 * on which the time filter is the identity (SPDY_LAZY_DIAG=0 / SPDY_SCAN_OUTER=0 switch both off). */
int spdy_run_steps(const int64_t *states, const int64_t *controls, int n_members, int nsteps, int *error_codes);
int spdy_reserve(int n_members);            /* pre-size the device arenas */
/* SPPT, stochastically perturbed parametrisation tendencies (sppt.f90:40-146, physics.f90:233-248; the reference's
 * compile-time switch params.f90:44, off by default here as there): per-member AR(1) pattern, counter-based generator */
int spdy_set_sppt(int on, unsigned long long seed);
int spdy_debug_get_sppt(int64_t state, double *spec /* (31,32,8) complex */, double *grid /* (96,48,8) */, long long *calls);
int spdy_set_device(int ordinal);           /* select the GPU (before the first state is created) */
int spdy_device_count(void);                /* CUDA devices visible to this process (0 without a driver) */
int spdy_synchronize(void);
/* elapsed device time (ms) of the last spdy_run_steps / spdy_parallel_step call, CUDA events on the launch stream */
float spdy_last_elapsed_ms(void);
long long spdy_kernel_launches(void);       /* kernels launched by this library so far */
/* host wall time (us) of the phases of the last step / parallel_step / run_steps call: prologue (entry to first launch),
 * kernel launches, waiting for the GPU, epilogue */
void spdy_last_call_host_us(double *out4);
/* cudaProfilerStart / cudaProfilerStop, so that `ncu --profile-from-start off` sees only a bracketed region */
int spdy_profiler_start(void);
int spdy_profiler_stop(void);
/* The member's model date as the device calendar holds it (control_params%model_datetime of the bound control,
   model_control.f90:113-163); out = {year, month, day, hour, minute}.  Returns 0, or -1 for an unknown handle. */
int spdy_get_model_datetime(int64_t state, int *out);
/* partial sums for ensemble mean / spread of ONE registry array variable over the listed members:
 * sum[i] = sum_m x_m[i], sumsq[i] = sum_m (x_m[i]-shift[i])^2, reduced on the device, summed over all ranks when a
 * communicator is up (spdy_comm_init), copied to the host */
int spdy_ensemble_sums(const int64_t *states, int n_members, int var, const double *shift, double *sum, double *sumsq);
/* same for this rank only, leaving the results on the device: returns a device pointer (2 * nelem doubles) */
int spdy_ensemble_sums_device(const int64_t *states, int n_members, int var, const double *shift_dev, void **sum_sumsq_dev,
                              size_t *nelem);
/* transform_spectral2grid (.j2:94-103) for the listed members and, in the epilogue of the same pass, ensemble mean and
 * spread (std, ddof 0: examples/Ensemble_forecast.ipynb cells 12, 16) of the six default outputs u, v, t, q, phi, ps on
 * the grid over all n_total members of all ranks: one in-stream ncclAllReduce(sum, double), one device-to-host copy, one
 * synchronisation.  out = 2 x 188,928 doubles (mean, then spread; variables one after another in Fortran order), or NULL
 * to leave them on the device (spdy_ensemble_stats_device_ptr). */
int spdy_ensemble_mean_spread(const int64_t *states, int n_members, long long n_total, double *out);
void *spdy_ensemble_stats_device_ptr(void);
/* get_<v> (.j2:264-290) for every listed member in one call: dst[member][Fortran-order array], float64 or (as_f32) cast
 * to float32 on the device as pyspeedy/speedy.py:443 does on the host */
int spdy_ensemble_get(const int64_t *states, int n_members, int var, void *dst, size_t bytes, int as_f32);
/* check (.j2:81-91) for every listed member in one call */
int spdy_batch_check(const int64_t *states, int n_members, int *error_codes);
/* datetime containers updated in place; the batched form stores one date in every listed container */
int spdy_set_datetime(int64_t dt, int year, int month, int day, int hour, int minute);
int spdy_set_datetimes(const int64_t *dts, int n, const int *ymdhm /* [5] */);

/* ---- multi-GPU: one process per GPU, members sharded across ranks, no communication inside a time step (.j2:58-79:
 * members are independent); NCCL (loaded with dlopen) only for the ensemble diagnostics above.  The 128-byte id made by
 * rank 0 reaches the other ranks through the launcher (pyspeedy_b200/distributed.py: file rendezvous under torchrun). */
int spdy_comm_unique_id(void *id128);
int spdy_comm_init(int rank, int world, const void *id128); /* after spdy_set_device */
int spdy_comm_rank(void);
int spdy_comm_world(void);
int spdy_comm_allreduce(double *host_inout, int n /* <= 64 */, int op /* 0 sum, 1 max */);
int spdy_comm_barrier(void);
int spdy_comm_destroy(void);

/* copy the complete device state of `src` into every member of `dst` (ensemble set-up from one initialised member) */
int spdy_clone_state(int64_t src, const int64_t *dst, int n);
/* t_grid += N(0,sigma) per grid point, then grid2spectral of T (examples/Ensemble_forecast.ipynb cell 8) for all listed
 * members, on the device, with a counter-based generator seeded per (member, point, level) */
int spdy_perturb_temperature(const int64_t *states, int n, unsigned long long seed, double sigma);
int spdy_batch_spectral2grid(const int64_t *states, int n);      /* transform_spectral2grid for a member list */
/* one step with CUDA events between kernel classes: ms[10] = forcing, pre-ops, legendre_inv, fft_inv, grid_dyn,
 * physics, fft_fwd, legendre_dir, spec_step, post (first chunk of 16 tiles is instrumented) */
int spdy_profile_step(const int64_t *states, const int64_t *controls, int n, float *ms, int *error_codes);
/* on != 0: spdy_profile_step times the step as an INTERMEDIATE step of a multi-step call runs it (the column physics does
 * not store the outputs that nothing reads before the next step overwrites them; extension, see spdy_run_steps) */
int spdy_profile_intermediate(int on);

/* ---- stage-level entry points (parity tests, microbenchmarks); host buffers, one field after another -------- */
int spdy_table(const char *name, double *dst, int cap);
int spdy_batch_spec2grid(const double *spec /* n x (31,32) complex */, double *grid /* n x (96,48) */, int kcos, int n);
int spdy_batch_grid2spec(const double *grid, double *spec, int n);
int spdy_batch_legendre_inv(const double *spec, double *four /* n x (62,48) */, int n);
int spdy_batch_legendre_dir(const double *four, double *spec, int n);
int spdy_batch_fourier_inv(const double *four, double *grid, int kcos, int n);
int spdy_batch_fourier_dir(const double *grid, double *four, int n);
/* spectral operators on batches of (31,32) complex fields, through the kernels of the model step (operator-level parity):
 * vort2vel = uvspec (spectral.f90:190-214), vel2vort = vdspec (:160-186), gradient (:275-296), laplacian / laplacian_inv
 * (:140-155), grid_vel2vort (:218-248; kcos = 2: cosgr, else cosgr2) */
int spdy_batch_vort2vel(const double *vor, const double *div, double *ucos, double *vcos, int n);
int spdy_batch_vel2vort(const double *ucos, const double *vcos, double *vor, double *div, int n);
int spdy_batch_gradient(const double *psi, double *dx, double *dy, int n);
int spdy_batch_laplacian(const double *in, double *out, int inverse, int n);
int spdy_batch_grid_vel2vort(const double *ug, const double *vg, double *vor, double *div, int kcos, int n);
/* BASELINE config 4: npairs synthetic (vor, div) pairs resident in HBM; one rep = vort2vel -> spec2grid(kcos 2) of the
 * 2 npairs wind fields -> grid2spec (cos-latitude loader) -> vel2vort -> gradient; ms[6] = mean device ms per rep, then per stage */
int spdy_bench_spectral_chain(const double *vor, const double *div, int npairs, int reps, float *ms);
/* BASELINE config 5: the column physics alone on synthetic columns; sets = nsets x 45 x (96,48) doubles (ug8, vg8, pslg,
 * utend8, vtend8, then tg, qg, phig, ttend, qtend with 8 levels each); member i works on set i % nsets; ms[2] = mean device ms
 * of a short-wave step and of a long-wave-only step */
int spdy_bench_physics(const int64_t *states, int n, const double *sets, int nsets, int reps, float *ms);
/* resident round trip for the spectral microbench: n synthetic fields stay in HBM, `reps` x (spec2grid, grid2spec);
 * returns average device ms per rep; per_kernel_ms[4] = legendre_inv, fft_inv, fft_fwd, legendre_dir */
int spdy_bench_roundtrip(const double *spec, double *spec_out, int n, int reps, float *ms_per_rep, float *per_kernel_ms);
/* column physics of one member on explicit grid inputs (physics.f90:103-231): (96,48,8) fields, tendencies in/out */
int spdy_debug_physics(int64_t state, const double *ug8, const double *vg8, const double *tg, const double *qg,
                       const double *phig, const double *pslg, double *utend8, double *vtend8, double *ttend,
                       double *qtend, int *dbg /* 3 x (96,48): itop, icnv, icltop */);
/* one raw leapfrog step(j1, j2, dt) of time_stepping.f90:38-147 without calendar/coupler; dt_kind 0: delt/2, 1: delt, 2: 2*delt */
int spdy_debug_raw_step(int64_t state, int j1, int j2, int dt_kind);
int spdy_debug_get_corh(int64_t state, double *tcorh, double *qcorh);
/* tendencies as returned by get_tendencies(state, ..., j2) (tendencies.f90:11-39); the prognostics are not advanced */
int spdy_debug_tendencies(int64_t state, int j2, double *vordt, double *divdt, double *tdt, double *psdt, double *trdt);
/* stage 1: divdt, tdt, psdt as they enter implicit_terms (time_stepping.f90:71-75; vordt, trdt as in stage 2); stage 2: same as above */
int spdy_debug_tendencies_stage(int64_t state, int j2, int stage, double *vordt, double *divdt, double *tdt, double *psdt,
                                double *trdt);

#ifdef __cplusplus
}
#endif
#endif
