/*
 * greb_oracle.h — CPU restatement of the reference GREB time-stepping core.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle and the CPU baseline; nothing in the
 * product path (greb-climate-model_b200/, include/) may include, link or call it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * HOW IT IS PINNED: sieste/greb-climate-model ships no tests, golden vectors or stored outputs,
 * there is no Fortran compiler in this image, and 7 of the 10 input files are absent (synthetic
 * inputs in the reference's binary format are used, greb_b200/synth.py).  The reference is
 * nevertheless executed here: oracle/f90_to_cpp.py translates its Fortran modules and
 * subroutines mechanically into C++ from the source tree where it lies, oracle/ref.py drives the
 * compiled result (oracle/_ref/) like PROGRAM greb_run does, and this restatement is compared
 * with it BIT FOR BIT — every kernel routine, the default 3+50-year run (3,000 records), a
 * perturbed member, the greb-original control + scenario run (tests/test_ref_pin.py; golden
 * vectors generated from it: tests/golden/, tests/test_golden.py).  Independently:
 *   (a) a second NumPy-float32 transcription (tests/np_greb.py),
 *   (b) the property / known-answer / bug-compatibility tests listed in SURVEY.md section 4.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference
 * root; "greb.f90" = src/greb.f90).  Arithmetic contract (reference Makefile:5-13 = gfortran
 * -O3, no -ffast-math, no -march): IEEE fp32 everywhere, no FMA contraction, no
 * reassociation, left-to-right evaluation with the parentheses as written.  Build this file
 * with `gcc -O3 -ffp-contract=off` and nothing that relaxes IEEE semantics.
 */
#ifndef GREB_ORACLE_H
#define GREB_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define GO_XDIM 96
#define GO_YDIM 48
#define GO_NSTEP_YR 730
#define GO_NCELL (GO_XDIM * GO_YDIM)

/* namelist physics_par + co2_flux; defaults = greb.f90:68-104 */
typedef struct {
  float pi, sig, rho_ocean, rho_land, rho_air, cp_ocean, cp_land, cp_air, eps;
  float d_ocean, d_land, d_air, ct_sens, da_ice, a_no_ice, a_cloud;
  float Tl_ice1, Tl_ice2, To_ice1, To_ice2, co_turb, kappa, ce, cq_latent, cq_rain;
  float z_air, z_vapor, r_qviwv;
  float p_emi[10];
  float co2_flux;
} go_physics;

typedef struct go_model go_model;

void go_physics_defaults(go_physics *p);              /* greb.f90:68-104 */
void go_physics_original(go_physics *p);              /* src/greb.original.model.f90:63-101 */

go_model *go_create(void);
void go_destroy(go_model *m);
void go_set_physics(go_model *m, const go_physics *p);
void go_get_physics(const go_model *m, go_physics *p);

/* Host part of PROGRAM greb_run (greb.f90:1073-1094): copies the inputs (C order
 * [time][lat][lon], sw_solar [time][lat]) and derives Toclim. */
void go_set_forcing(go_model *m, const float *z_topo, const float *glacier, const float *sw_solar,
                    const float *tclim, const float *qclim, const float *swetclim, const float *uclim,
                    const float *vclim, const float *mldclim, const float *cldclim);

/* Preamble of greb_model (greb.f90:176-216): dTrad, z_ocean, heat capacities, cap_surf,
 * initial state, wz_*, wind sign split.  Must be called after set_physics + set_forcing. */
void go_setup(go_model *m);

/* qflux_correction (greb.f90:311-364) on the model's current state, `years` years at `co2`. */
void go_qflux_correction(go_model *m, int years, float co2);

/* Scenario loop (greb.f90:226-234).  co2_ppm has `years` entries (already padded).  Resets
 * mon/irec/year and Tmm,Tamm,qmm,apmm (NOT Tomm) like greb.f90:227.  `out` (may be NULL)
 * receives years*12*5 records [month][var][lat][lon]; `gmean` (may be NULL) receives per year
 * the reference console value sum(tsmn)/(xdim*ydim)-273.15 (greb.f90:954).  year0 is the
 * namelist start year.  If `continue_run` is non-zero the calendar/accumulators are not reset
 * (used to chain calls). */
void go_run_scenario(go_model *m, int years, const float *co2_ppm, int year0, float *out, float *gmean,
                     int continue_run);

/* One time_loop call (greb.f90:239-274) on the model state with step counter `it` (1-based)
 * and CO2.  If out5 != NULL and a month ends at this step, the 5 records are written there and 1
 * is returned. */
int go_time_loop(go_model *m, int it, float co2, float *out5);

/* state access: fields are [48][96] */
enum { GO_TS = 0, GO_TA = 1, GO_TO = 2, GO_Q = 3, GO_CAP = 4 };
void go_get_state(const go_model *m, int which, float *out);
void go_set_state(go_model *m, int which, const float *in);
/* flux corrections [730][48][96]: 0=TF 1=qF 2=ToF */
const float *go_fluxcorr(const go_model *m, int which);
/* derived fields: 0=wz_air 1=wz_vapor 2=z_ocean 3=Toclim(step 1) */
void go_get_derived(const go_model *m, int which, float *out);
void go_set_ityr(go_model *m, int ityr); /* module variable ityr (1..730) for kernel-level calls */

/* kernel-level entries (same argument meaning as the Fortran subroutines) */
void go_diffusion(const go_model *m, const float *T1, float *dX, const float *wz);          /* :556-723 */
void go_advection(const go_model *m, const float *T1, float *dX, const float *wz);          /* :726-915 */
void go_circulation(const go_model *m, const float *X_in, float *dX_crcl, const float *wz); /* :528-553 */
void go_SWradiation(const go_model *m, const float *Tsurf, float *sw, float *albedo);       /* :367-403 */
void go_LWradiation(const go_model *m, const float *Tsurf, const float *Tair, const float *q, float CO2,
                    float *LWsurf, float *LWair_up, float *LWair_down, float *em);          /* :407-434 */
void go_hydro(const go_model *m, const float *Tsurf, const float *q, float *Qlat, float *Qlat_air,
              float *dq_eva, float *dq_rain);                                               /* :438-469 */
void go_seaice(go_model *m, const float *Tsurf);                                            /* :472-492 */
void go_deep_ocean(const go_model *m, const float *Ts, const float *To, float *dT_ocean, float *dTo); /* :495-525 */

/* geometry of diffusion/advection for row k (1..48): fills the per-row constants the
 * reference recomputes on every call (greb.f90:578-582, 652-654, 749-753, 838-840). */
typedef struct {
  float deg, dyy, ccy_diff, ccy_adv;
  float dxlat[GO_YDIM], ccx_diff[GO_YDIM], ccx_adv[GO_YDIM];
  float ccx2_diff[GO_YDIM], ccx2_adv[GO_YDIM];
  int polar[GO_YDIM], time2_diff[GO_YDIM], time2_adv[GO_YDIM];
} go_geometry;
void go_geometry_compute(float pi, float kappa, go_geometry *g);
void go_libm_array(int which, const float *x, float *y, int n);

#ifdef __cplusplus
}
#endif
#endif
