// greb_setup.h — host-side preparation shared by the CUDA runtime (greb_b200.cu) and the
// test-only lane emulator (tests/emu): everything the reference computes once before the time
// loops (PROGRAM greb_run / greb_model preamble) plus the geometry the reference recomputes in
// every diffusion/advection call.  Pure C++ on host memory; no CUDA.
#pragma once

#include <vector>

#include "../../include/greb_b200.h"
#include "greb_types.h"

struct GrebHostForcing {
  // shared by all members
  std::vector<float> forc;      // [730][GF_COUNT][GNC]
  std::vector<float> sw_solar;  // [730][48]
  std::vector<int> mask;        // [GNC]
  std::vector<float> z_topo;    // [GNC]
  std::vector<float> z_ocean;   // [GNC]
  std::vector<float> toclim;    // [GNC]
  std::vector<float> tclim;     // [730][GNC]
  std::vector<float> qclim;     // [730][GNC]
  std::vector<float> mld0;      // [GNC] mldclim(:,:,1) for the cap_surf initialisation
  std::vector<float> coslat_w;  // [48]
};

// PROGRAM greb_run f:1073-1094 + the time-invariant parts of greb_model f:176-183 and of
// hydro f:452-454 / deep_ocean f:507-508.
void greb_build_forcing(GrebHostForcing& F, const float* z_topo, const float* glacier, const float* sw_solar,
                        const float* tclim, const float* qclim, const float* swetclim, const float* uclim,
                        const float* vclim, const float* mldclim, const float* cldclim);

// physics scalars, heat capacities (f:186-188), circulation geometry and the row assignment.
// Returns 0, or <0 if the geometry needs more helper rows than the kernel supports (kappa far
// outside the reference's range) or sub-stepped polar advection (impossible at 96x48).
int greb_build_member_const(GrebMemberConst& mc, const greb_physics_par& p, int group,
                            const int* warp_order = nullptr);

// wz_air / wz_vapor of one physics group (f:201-202): out[2][GNC]
void greb_build_wz(float* out, const GrebHostForcing& F, const greb_physics_par& p);

// initial state of one member (f:190-197): out[GS_COUNT][GNC]
void greb_build_initial_state(float* out, const GrebHostForcing& F, const GrebMemberConst& mc);

// lane group -> latitude row table and helper-warp slots
int greb_assign_rows(const int* polar, const int* time2_diff, const int* time2_adv, int* row_of_group,
                     int* hslot_of_row, int* helper_row, int* n_hslots, const int* warp_order = nullptr);

bool greb_physics_equal(const greb_physics_par& a, const greb_physics_par& b);
