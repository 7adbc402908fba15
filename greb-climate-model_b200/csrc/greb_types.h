// greb_types.h — plain-old-data shared by the host runtime and the kernels.
#pragma once

#define GX 96           // longitudes (reference src/greb.f90:36 xdim)
#define GY 48           // latitudes  (ydim)
#define GNC (GX * GY)   // cells per field
#define GNT 730         // steps per year (nstep_yr, :41)
#define GSUB 24         // circulation sub-steps per step: nint(43200/1800) (:543)

// CTA shape: 12 "main" warps, each lane group of 8 lanes owns one latitude row (12 consecutive
// longitudes per lane, 4 rows per warp), plus GREB_NHELP helper warps that run the polar
// sub-sub-steps of the rows whose diffusion needs more than one (time2_diff > 1, f:652-717).
#define GREB_NMAIN 12
#define GREB_NHELP 2
#define GREB_NWARP (GREB_NMAIN + GREB_NHELP)
// Warp placement.  The kernels are launched with 16 warp slots (512 threads); a warp's SM sub-partition
// is (hardware warp id % 4).  Which slot plays which role is a run-time table (GrebKernelArgs::warp_map,
// 4 bits per hardware warp id: logical warp 0..11 = main, 12..13 = helper, 15 = empty slot, exits at once),
// so that the placement can differ between the arithmetic modes (greb_b200.cu, greb_layouts): the pole-row
// chain of a helper warp is a latency-bound serial chain that sets the length of a sub-step, and sharing
// a sub-partition with three main warps made it 2.6x slower than alone.
#define GREB_NSLOTS 16
#define GREB_NTHREADS (GREB_NSLOTS * 32)
#define GREB_SLOT_EMPTY 15
// Stagger (GrebKernelArgs::late_mask, one bit per logical main warp): a "late" warp does the x-direction
// part of a sub-step AFTER the barrier instead of before it, so that the shared-memory-bound y parts of
// one half of the warps overlap the arithmetic-bound x parts of the other half (after a barrier all warps
// used to hit the LDS pipe at once).  Same arithmetic, same results.
// GREB_PACKED_Y (exact mode): the y part of a sub-step on pairs of cells with FADD2/FMUL2/FFMA2
// (greb_core.h substep_y_packed): same IEEE operations — 16 perturbed members x 53 years stay bit-identical
// on the B200 — and 20 % fewer issued instructions per sub-step (114 packed instructions replace 228), yet
// the throughput does not move: 1,653.9 vs 1,654.2 member-years/s.  The circulation is bound by the
// latency of each warp's dependent chain (SHFL -> stencil -> barrier -> LDS -> update) at 3.5 warps per
// scheduler, not by issue slots.  Off by default (no gain, and .ftz on the packed adds); kept as a build
// option and as evidence.  A variant that kept the wind branch in a bit mask register instead of
// re-deriving it from V cost 5.5 % — one more live register at the 128-register limit spills in the loop.
#ifndef GREB_PACKED_Y
#define GREB_PACKED_Y 0
#endif
// GREB_YCOEF (fast arithmetic mode only): the latitudinal coefficients of a step are folded into three
// per-cell factors CA, CB, CF once per step (instead of V, wz(k-1), wz(k+1), WFY), and the x part hands
// over one array wz*dTx + aTx instead of two.
#ifndef GREB_YCOEF
#define GREB_YCOEF 1
#endif
#define GREB_CPT 12     // cells per thread (96 / 8)
#define GREB_MAXH 4     // max helper-owned rows (2 per helper warp)

// shared memory layout of a member CTA, in floats
#define GSM_HB 0                        // [2][GNC]   double-buffered copy of the circulating field
#define GSM_STASH (2 * GNC)             // [2][GNC]   tendA, tq between the column phase and the circulations
#define GSM_SYNC (4 * GNC)              // SplitBar (32 floats reserved)
// per-thread private constants of the y-direction part, kept out of the register file:
// [PRIV_*][chunk 0..2][main thread 0..383][4]  -> consecutive threads read consecutive 16 bytes
#define GSM_PRIV (GSM_SYNC + 32)
enum { PRIV_V = 0, PRIV_WM1, PRIV_WP1, PRIV_WFY, PRIV_COUNT };
// Flux corrections of a scenario step (55 KB per step and physics group, the one per-member HBM stream).
// Default (GREB_TMA_CORR = 0): one thread asks the bulk-copy engine to pull the NEXT step's slice into L2
// (cp.async.bulk.prefetch.L2) while the circulations run; phases A and C then read it with ordinary loads.
// GREB_TMA_CORR = 1: the slices are streamed into SHARED memory by the TMA engine (cp.async.bulk +
// transaction barriers, member_run_main).  Built, bit-identical, and measured SLOWER on B200 (exact mode
// 1,552 -> 1,420, fast mode 2,673 -> 2,474 member-years/s): the extra 54 KB of shared memory shrink the
// L1 that serves the five-fold re-reads of the wz rows, and the corrections were never the latency that
// bounds phase A (3 of 27 field reads per step).  Kept as a build option.
#ifndef GREB_TMA_CORR
#define GREB_TMA_CORR 0
#endif
#define GSM_CORR (GSM_PRIV + PRIV_COUNT * 3 * GREB_NMAIN * 32 * 4)
#if GREB_TMA_CORR
#define GSM_FLOATS (GSM_CORR + 3 * GNC)
#else
#define GSM_FLOATS GSM_CORR
#endif
#define GSM_TMA_BAR_A (GSM_SYNC + 4)   // floats: 8-byte aligned (GSM_SYNC is)
#define GSM_TMA_BAR_Q (GSM_SYNC + 6)

// per-step shared forcing record: forc[ityr][GF_*][GNC]
// GF_RDEEP = dmld / (z_ocean - mld) and GF_RMIX = dmld / mld are the two member-independent quotients of
// deep_ocean (f:511-514), evaluated once on the host with the same IEEE fp32 operations
enum { GF_U = 0, GF_V, GF_CLD, GF_DTRAD, GF_SWET, GF_ABSWIND, GF_MLD, GF_DMLD, GF_RDEEP, GF_RMIX, GF_COUNT };
// flux corrections: corr[group][ityr][GC_*][GNC]   (src/greb.f90:110)
enum { GC_TF = 0, GC_TOF, GC_QF, GC_COUNT };
// state[member][GS_*][GNC]
enum { GS_TS = 0, GS_TA, GS_TO, GS_Q, GS_CAP, GS_COUNT };
// acc[member][GA_*][GNC]: the five monthly accumulators (:149) + annual Tsurf accumulator tsmn (:145)
enum { GA_TMM = 0, GA_TAMM, GA_TOMM, GA_QMM, GA_APMM, GA_TSMN, GA_COUNT };
// static mask bits (the reference's three different land/ocean predicates)
enum { GM_TOPO_GE0 = 1, GM_TOPO_LT0 = 2, GM_GLACIER = 4, GM_TOPO_GT0 = 8 };

struct GrebMemberConst {
  // double-precision reciprocals of the member constants the column physics divides by (exact mode):
  // RN_f32(x / c) == RN_f32(RN_f64(x * RN_f64(1/c))) for all fp32 x, c (greb_core.h v_divc)
  double rc_alb_land, rc_alb_ocean;  // 1 / fl(Tl_ice2 - Tl_ice1), 1 / fl(To_ice2 - To_ice1)   f:384-392, f:486
  double rc_pe8, rc_cq_latent, rc_r_qviwv, rc_cap_air;
  // physics scalars used on the device (namelist physics_par, src/greb.f90:68-104)
  float sig, ct_sens, da_ice, a_no_ice, a_cloud, Tl_ice1, Tl_ice2, To_ice1, To_ice2;
  float co_turb, ce, cq_latent, cq_rain, rho_air, r_qviwv;
  float p_emi[10];
  float cap_ocean, cap_land, cap_air;  // :186-188
  float co2_flux;
  // circulation geometry (:578-582, 652-654, 749-753, 838-840), computed on the host with the
  // same libm the reference links, so the device never evaluates cos()
  float ccy_diff, ccy_adv;
  float ccx_diff[GY], ccx_adv[GY], ccx2_diff[GY], ccx2_adv[GY];
  int polar[GY], time2_diff[GY], time2_adv[GY];
  // row ownership: main lane group g = 4*warp + (lane>>3) owns latitude row row_of_group[g]
  int row_of_group[GY];
  // rows whose polar diffusion needs several sub-sub-steps go to the helper warps:
  // hslot_of_row[k] = slot or -1; helper warp h serves slots h, h+GREB_NHELP, ...
  int hslot_of_row[GY];
  int helper_row[GREB_MAXH];
  int n_hslots;
  int group;  // physics group (shares wz fields and flux corrections)
  int switches;  // GREB_SW_* process switches (include/greb_b200.h), 0 = the full model
  int pad_[1];
};

struct GrebKernelArgs {
  const GrebMemberConst* mc;  // [n_members]
  const int* member_ids;      // [gridDim.x] member handled by each CTA
  const float* forc;          // [730][GF_COUNT][GNC]
  const float* sw_solar;      // [730][48]
  const int* mask;            // [GNC]
  const float* z_ocean;       // [GNC]
  const float* toclim;        // [GNC]
  const float* tclim;         // [730][GNC] (spin-up target)
  const float* qclim;         // [730][GNC]
  const float* wz;            // [n_groups][2][GNC]: wz_air, wz_vapor
  float* corr;                // [n_groups][730][GC_COUNT][GNC]
  float* state;               // [n_members][GS_COUNT][GNC]
  float* acc;                 // [n_members][GA_COUNT][GNC]
  float* out;                 // [n_members][out_months][5][GNC] monthly means of this launch (may be null)
  const float* co2;           // [n_members][co2_stride] annual CO2 path
  float* diag;                // [n_members][2]: {unweighted, cos-lat} annual-mean Tsurf [deg C]
  const float* coslat_w;      // [48] normalised cos-lat weights
  int* flags;                 // [n_members] non-finite flag
  int co2_stride;
  int out_months;             // capacity of `out` in months per member
  int it0;                    // first step counter `it` (1-based) of this launch
  int nsteps;
  int spinup;                 // 1 = qflux_correction step (:325-362), 0 = time_loop (:239-274)
  unsigned late_mask;         // stagger: bit w set = logical main warp w does its x part after the barrier
  unsigned long long warp_map;  // 4 bits per hardware warp id -> logical warp (GREB_SLOT_EMPTY = unused slot)
};

// kernel-level circulation entry
struct GrebCirculationArgs {
  const GrebMemberConst* mc;  // one member's constants
  const float* uv;            // [2][GNC] u, v of the requested step
  const float* X_in;          // [n][GNC]
  const float* wz;            // [n][GNC]
  float* dX;                  // [n][GNC]
  unsigned late_mask;
  unsigned long long warp_map;
};
