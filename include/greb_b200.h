/*
 * greb_b200.h — C ABI of the B200-native GREB time-stepping core.
 *
 * This is the drop-in boundary for the reference's hot path.  The
 * reference (sieste/greb-climate-model, Fortran 90) has no FFI layer: its boundary is a set of
 * external subroutines over `real(xdim,ydim)` arrays plus module globals.  Each entry point
 * below names the reference interface it replaces (file:line under the reference root); the
 * Fortran ISO_C_BINDING stub a maintainer would add is in INTEGRATION.md and
 * greb-climate-model_b200/fortran/greb_b200_host.f90.
 *
 * Conventions
 *   - Plain C: opaque handle, pointers, sizes.  No C++/torch types.
 *   - Every function returns 0 on success, a negative GREB_E_* code on failure; the message is
 *     available from greb_b200_last_error().  Nothing calls exit().
 *   - Arrays are the reference's own memory layout, no transposes: Fortran X(i,j[,n]) ==
 *     C X[n][j][i], i = longitude fastest (96), j = latitude (48), n = step of year (730);
 *     sw_solar(j,n) == [n][j].  All data are IEEE fp32.
 *   - Host pointers unless the name says _device.  The library owns all device memory behind the
 *     handle and keeps no caller pointer after a call returns.
 *   - One handle drives one GPU.  Calls on one handle are not thread-safe; distinct handles are.
 *   - There is no CPU fallback: every compute entry point fails with GREB_E_NO_DEVICE when no
 *     sm_100 device is usable.
 */
#ifndef GREB_B200_H
#define GREB_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GREB_XDIM 96
#define GREB_YDIM 48
#define GREB_NSTEP_YR 730
#define GREB_NCELL (GREB_XDIM * GREB_YDIM)
#define GREB_NVAR_OUT 5 /* Tsurf, Tair, Tocean, q, albedo: reference src/greb.f90:978-982 */

enum {
  GREB_OK = 0,
  GREB_E_INVALID = -1,   /* bad argument / call order */
  GREB_E_NO_DEVICE = -2, /* no usable sm_100 GPU (there is no CPU fallback) */
  GREB_E_CUDA = -3,      /* CUDA runtime error, see greb_b200_last_error */
  GREB_E_NOMEM = -4,
  GREB_E_NONFINITE = -5  /* a member produced a non-finite state */
};

/* The reference's namelist physics_par (src/greb.f90:68-104,128-132) plus co2_flux of
 * namelist co2_par (:104,134).  One block per ensemble member.  greb_b200_physics_defaults fills
 * the reference defaults; greb_b200_physics_original fills greb.original.model.f90:63-101 with
 * CO2_ctrl=340 (:178). */
typedef struct greb_physics_par {
  float pi, sig, rho_ocean, rho_land, rho_air, cp_ocean, cp_land, cp_air, eps;
  float d_ocean, d_land, d_air, ct_sens, da_ice, a_no_ice, a_cloud;
  float Tl_ice1, Tl_ice2, To_ice1, To_ice2, co_turb, kappa, ce, cq_latent, cq_rain;
  float z_air, z_vapor, r_qviwv;
  float p_emi[10];
  float co2_flux;
} greb_physics_par;

typedef struct greb_b200_handle_s* greb_b200_t;

/* ---- lifecycle ----------------------------------------------------------------------------- */

void greb_b200_physics_defaults(greb_physics_par* p);
void greb_b200_physics_original(greb_physics_par* p);

/* Creates a context for `n_members` independent ensemble members on CUDA device `device`.
 * Replaces: one `./greb <namelist>` OS process per member (src/greb.f90:1030-1038, 1063-1068). */
int greb_b200_create(greb_b200_t* h, int n_members, int device);
int greb_b200_destroy(greb_b200_t h);
const char* greb_b200_last_error(greb_b200_t h); /* h may be NULL: last create() error */
int greb_b200_n_members(greb_b200_t h);

/* ---- inputs -------------------------------------------------------------------------------- */

/* The ten input fields PROGRAM greb_run reads (src/greb.f90:1073-1085) — shared by all members.
 * Toclim is derived inside exactly as src/greb.f90:1087-1094.  Copies to the device once. */
/* Arithmetic of the circulation kernels.  GREB_ARITH_EXACT (default): IEEE fp32 in the reference's
 * expression order, no FMA contraction — diffusion/advection/circulation are bit-identical to the
 * reference built with its own flags (Makefile:5-13).  GREB_ARITH_FAST: the same stencils
 * algebraically factored with FMA contraction, 2.4x fewer instructions; results
 * agree with the reference within the tolerances of BASELINE.json (per-cell monthly T <= 0.01 K,
 * q <= 1e-6, global mean <= 1e-3 K over the 50-year run) instead of bit for bit.  May be changed
 * between launches. */
enum { GREB_ARITH_EXACT = 0, GREB_ARITH_FAST = 1 };
int greb_b200_set_arithmetic(greb_b200_t h, int mode);

int greb_b200_set_forcing(greb_b200_t h, const float* z_topo /*[48][96]*/, const float* glacier /*[48][96]*/,
                          const float* sw_solar /*[730][48]*/, const float* tclim /*[730][48][96]*/,
                          const float* qclim, const float* swetclim, const float* uclim, const float* vclim,
                          const float* mldclim, const float* cldclim);

/* Per-member namelist values: physics_par (+co2_flux) and the CO2 path of co2_par already
 * padded to n_years entries the way src/greb.f90:1053-1061 pads it (greb_b200_pad_co2 does it).
 * Members with identical physics share one flux-correction spin-up. */
int greb_b200_set_member(greb_b200_t h, int member, const greb_physics_par* p, const float* co2_ppm, int n_years,
                         int year0);
/* Process switches of one member: the well-defined sensitivity experiments of
 * src/greb.original.model.f90 (`log_exp`, :60) as a bit mask; 0 (default) is the full model =
 * log_exp 10 = src/greb.f90.  INTEGRATION.md lists the mask + input changes of every log_exp.
 *   NO_ICE_ALBEDO      a_surf = a_no_ice (:394) and cap_surf without the sea-ice ramp (:492-495)
 *   NO_HYDRO           hydro returns zeros (:452-453)
 *   NO_DEEP_OCEAN      deep_ocean returns zeros (:513-515)
 *   VAPOR_DIFFUSION_ONLY  circulation of q without advection (:560-564)
 *   LINEAR_VAPOR_EMISSIVITY  e_vapor from qclim + linear term in em (:423, :430)
 *   NO_HEAT_CIRCULATION / NO_VAPOR_CIRCULATION   `circulation` of Ta / of q returns at once (:553-555).  The
 *                      reference leaves its intent(out) result unassigned there; this library DEFINES it
 *                      as dX_crcl = 0 (what the reference computes when its local arrays are zero-initialised
 *                      static storage, e.g. gfortran -fno-automatic, and what the translated reference behind the
 *                      golden vectors does)
 *   SST_PLUS_1K        scenario only: Ts1 = Tclim(:,:,ityr) + 1 where z_topo < 0 before every step (:226;
 *                      ityr there still is the PREVIOUS step's, :248 updates it afterwards)
 * Members with different masks never share a spin-up.  All bits but SST_PLUS_1K must be set
 * before greb_b200_init; SST_PLUS_1K may be toggled between runs (control run without, scenario
 * with). */
enum {
  GREB_SW_NO_ICE_ALBEDO = 1,
  GREB_SW_NO_HYDRO = 2,
  GREB_SW_NO_DEEP_OCEAN = 4,
  GREB_SW_VAPOR_DIFFUSION_ONLY = 8,
  GREB_SW_LINEAR_VAPOR_EMISSIVITY = 16,
  GREB_SW_SST_PLUS_1K = 32,
  GREB_SW_NO_HEAT_CIRCULATION = 64,
  GREB_SW_NO_VAPOR_CIRCULATION = 128,
  GREB_SW_ALL = 255
};
int greb_b200_set_switches(greb_b200_t h, int member, unsigned mask);
/* wz = exp(-z_topo / h_scale) (src/greb.f90:201-202: wz_air with z_air, wz_vapor with z_vapor) for n cells,
 * evaluated on the host with the libm the rest of the set-up uses (hosts of grids other than 96x48). */
void greb_b200_wz(const float* z_topo, float h_scale, float* out, long n);
/* src/greb.f90:1047-1061: `n_given` values followed by padding to n_years (first<0 -> 680). */
void greb_b200_pad_co2(const float* given, int n_given, float* co2_ppm, int n_years);

/* ---- model run ----------------------------------------------------------------------------- */

/* greb_model preamble (src/greb.f90:176-216): derived fields, heat capacities, initial state
 * (step 730 of the climatology), wz_*, geometry of the circulation sub-steps.  Call after
 * set_forcing and all set_member calls. */
int greb_b200_init(greb_b200_t h);

/* qflux_correction (src/greb.f90:311-364): `years` years at each member's co2_flux; leaves the
 * flux corrections on the device and the members in the spin-up end state (:361). */
int greb_b200_spinup(greb_b200_t h, int years);

/* Scenario loop (src/greb.f90:226-234), `years` more years for every member, continuing the
 * calendar (call greb_b200_reset_scenario first for the :227 reset).  For each simulated year
 * the 12x5 monthly-mean records of every member are produced on the device; if `out` is not
 * NULL they are copied to host as out[member][year][month][var][lat][lon] (the reference's
 * record stream of unit 22, :978-982, per member); `out_members` (may be NULL = all) lists the
 * `n_out` members to copy.  gmean (may be NULL) receives [member][year] the console value
 * sum(tsmn)/(xdim*ydim)-273.15 (:954); gmean_coslat (may be NULL) the cos-lat weighted annual
 * mean Tsurf in deg C (README.md:36-37). */
int greb_b200_reset_scenario(greb_b200_t h);
int greb_b200_run(greb_b200_t h, int years, float* out, const int* out_members, int n_out, float* gmean,
                  float* gmean_coslat);

/* The same loop without blocking the host: greb_b200_run_async enqueues the years (kernels on the
 * handle's compute stream, the device->host copies of each year's records on its copy stream) and
 * returns; greb_b200_wait blocks until everything enqueued so far is complete and only then fills
 * gmean / gmean_coslat (from a pinned mirror: no host synchronisation per simulated year).  Calls may
 * be chained: a second run_async issued before the wait starts its kernels while the previous call's
 * records are still being copied — `out` of a call must stay valid (and should be pinned) until the
 * next greb_b200_wait.  greb_b200_run == run_async + wait.  Every other entry point that touches the
 * device state waits implicitly first. */
int greb_b200_run_async(greb_b200_t h, int years, float* out, const int* out_members, int n_out, float* gmean,
                        float* gmean_coslat);
int greb_b200_wait(greb_b200_t h);
/* The last completed year's records -> host (pinned) on the copy stream, ordered behind whatever is enqueued
 * on the compute stream at the time of the call; completed by greb_b200_wait.  With run_async(out = NULL)
 * it lets a host put a small transfer (the end state) ahead of the gigabyte of records on the DMA engine. */
int greb_b200_fetch_monthly_async(greb_b200_t h, float* out, const int* out_members, int n_out);

/* One time_loop call (src/greb.f90:239-274) for every member with step counter `it` (1-based,
 * as the reference's `it`).  Test entry: lets a host drive the loop step by step.
 * greb_b200_time_steps: `nsteps` (<= 730) consecutive calls it0, it0+1, ... in one launch — how a
 * host continues from a mid-year checkpoint to the next year boundary (greb_b200_run needs one). */
int greb_b200_time_loop(greb_b200_t h, int it);
int greb_b200_time_steps(greb_b200_t h, int it0, int nsteps);

/* ---- state / results access ---------------------------------------------------------------- */

enum { GREB_TS = 0, GREB_TA = 1, GREB_TO = 2, GREB_Q = 3, GREB_CAP = 4 };
int greb_b200_get_state(greb_b200_t h, int member, int which, float* out /*[48][96]*/);
int greb_b200_set_state(greb_b200_t h, int member, int which, const float* in);
/* bulk variants: all members at once, [n_members][5][48][96] in the order of the enum above */
int greb_b200_get_states(greb_b200_t h, float* out);
int greb_b200_set_states(greb_b200_t h, const float* in);
/* Pipelined host loops: the same transfers enqueued on the handle's compute stream without waiting
 * (ordered with the kernels of run_async); greb_b200_sync_compute waits for that stream only, so
 * record copies of earlier years keep running on the copy stream.  Host buffers should be pinned. */
int greb_b200_set_states_async(greb_b200_t h, const float* in);
int greb_b200_get_states_async(greb_b200_t h, float* out);
int greb_b200_sync_compute(greb_b200_t h);
/* Checkpoint / resume of a scenario.  The reference keeps its loop state in the module arrays
 * Ts1,Ta1,To1,q1 (+cap_surf), the counters it/year/mon/irec — all functions of `it`
 * (src/greb.f90:226-234, 241-252, 975-985) — and the accumulators Tmm,Tamm,Tomm,qmm,apmm (:149) and
 * tsmn (:145).  get/set_states + get/set_calendar (`it_next` = the `it` of the next step, 1-based)
 * + get/set_accumulators ([n_members][6][48][96] in that order) save and restore all of it; with the
 * flux corrections (get/set_fluxcorr) a fresh handle continues a run bit for bit. */
int greb_b200_get_calendar(greb_b200_t h, int* it_next);
int greb_b200_set_calendar(greb_b200_t h, int it_next);
int greb_b200_get_accumulators(greb_b200_t h, float* out);
int greb_b200_set_accumulators(greb_b200_t h, const float* in);
/* which: 0 = TF_correct, 1 = qF_correct, 2 = ToF_correct (src/greb.f90:110), out [730][48][96] */
int greb_b200_get_fluxcorr(greb_b200_t h, int member, int which, float* out);
/* Restores flux corrections saved with greb_b200_get_fluxcorr (a spin-up cache: together with
 * get/set_state it lets a later run skip qflux_correction, src/greb.f90:311-364).  The corrections
 * belong to the member's physics group, i.e. to every member with identical physics_par. */
int greb_b200_set_fluxcorr(greb_b200_t h, int member, int which, const float* in);
/* last completed year's monthly means of one member: out[12][5][48][96] */
int greb_b200_get_monthly(greb_b200_t h, int member, float* out);
/* device-resident per-member diagnostics of the last completed year, 2 floats per member
 * {unweighted mean, cos-lat mean} in deg C — the vector a multi-GPU driver all-reduces. */
int greb_b200_diag_device(greb_b200_t h, const float** dev_ptr, int* n_floats);
/* Ensemble statistics of the last completed year's monthly means, computed on the device: for every
 * element of the record block [12][5][48][96] the sum and the sum of squares over the handle's members
 * (float64, member order).  A 65,536-member campaign keeps these instead of 72 GB of records per year
 * (SURVEY.md 8d, config 4); the _device variant returns the device buffers (valid until the next call) —
 * the vectors a multi-GPU driver reduces with ncclReduce / all-reduce (greb_b200/sharding.py). */
int greb_b200_ensemble_moments(greb_b200_t h, double* sum /*[12][5][48][96]*/, double* sumsq);
int greb_b200_ensemble_moments_device(greb_b200_t h, const double** dev_sum, const double** dev_sumsq, int* n_elements);
/* per-member non-finite flags (1 = a non-finite value was seen) */
int greb_b200_get_flags(greb_b200_t h, int* flags /*[n_members]*/);

/* ---- kernel-level entries (parity tests against the reference subroutines) ------------------ */

/* circulation(X_in, dX_crcl, h_scl, wz) (src/greb.f90:528-553) for `n` independent fields, using
 * member `member`'s geometry (pi, kappa) and the wind climatology of step `ityr` (1..730). */
int greb_b200_circulation(greb_b200_t h, int member, int ityr, const float* X_in, const float* wz, float* dX_crcl,
                          int n);

/* The exact mode's expf (which = 0) / logf (which = 1) on n arguments.  The reference evaluates exp and
 * log (src/greb.f90:422-424, 457) with glibc's libm; the exact mode restates glibc's algorithm on the
 * device so that whole runs, not only the circulation, are bit-identical (greb_simt.h). */
int greb_b200_device_libm(greb_b200_t h, int which, const float* x, float* y, int n);

/* Column physics of one 12-h step on TILES — the cell-local part of `tendencies` + the `time_loop` updates
 * (src/greb.f90:277-308, 258-272: SWradiation, LWradiation, sensible heat, hydro, deep_ocean, seaice) for a
 * grid of any size (BASELINE.json configs[4]); the circulations come from include/greb_grid.h.  A band of the
 * grid is cut into tiles of 4,608 consecutive cells laid out like an ensemble member; all pointers are
 * DEVICE pointers: forc [ntiles][10][4608] (u, v, cld, dTrad, swet, abswind, mld, dmld, dmld/(z_ocean-mld),
 * dmld/mld of the step), sw_solar [ntiles][48] (one value per 96-cell segment), mask [ntiles][4608] (bit 0:
 * z_topo >= 0, 1: < 0, 2: glacier, 3: > 0), z_ocean, wz [ntiles][2][4608] (air, vapour), corr [ntiles][3][4608]
 * (TF, ToF, qF of the step), state [ntiles][5][4608], acc [ntiles][6][4608], stash [ntiles][2][4608] (carries
 * the tendencies from phase 0 to phases 1, 2), X [ntiles][4608] = the circulated field.  phase 0 = before the
 * circulations, 1 = after circulation(Ta), 2 = after circulation(q).  Runs on the legacy default stream and
 * returns when done.  The same device code as the member kernel: at 96x48 a step is bit-identical to it. */
int greb_b200_tile_phase(int device, int arith, int phase, int ntiles, const greb_physics_par* p, float co2,
                         const float* forc, const float* sw_solar, const int* mask, const float* z_ocean,
                         const float* wz, float* corr, float* state, float* acc, float* stash, const float* X);

/* ---- timing of the last spinup/run call (CUDA events on the launch stream, ms) -------------- */
int greb_b200_last_kernel_ms(greb_b200_t h, float* ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* GREB_B200_H */
