// greb_emu.cpp — TEST-ONLY lane emulator for the warp-level kernel source.
//
// Compiles greb-climate-model_b200/csrc/greb_core.h with -DGREB_EMU (vf = 32-lane array, one
// pthread per warp, pthread barrier for the CTA barrier) so the exact kernel logic — row
// ownership, SHFL wrap, halo exchange, polar sub-sub-steps, the f:881 index bug — can be
// compared bit-for-bit with the oracle on a machine without a GPU.  It is NOT a CPU fallback:
// nothing in the product library links or calls it.
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include <stdint.h>

#include <vector>

#include "greb_core.h"
#include "greb_setup.h"

namespace {

struct Emu {
  GrebHostForcing F;
  greb_physics_par phys;
  GrebMemberConst mc;
  std::vector<float> wz, corr, state, acc, out, co2, diag;
  std::vector<int> flags;
  int member_id = 0;
};

struct WarpTask {
  SimtCtx ctx;
  const GrebKernelArgs* ka;
  const GrebCirculationArgs* ca;
  const GrebMemberConst* mc;
  int n_idx;
};

#define EMU_UNITS (GY + GREB_NHELP)  // 48 row groups (8 lanes each) + the helper warps

void* warp_main(void* p) {
  WarpTask* t = (WarpTask*)p;
  if (t->ka) {
    if (t->mc->switches) member_run<0, 1>(t->ctx, *t->ka, *t->mc, 0);
    else member_run<0, 0>(t->ctx, *t->ka, *t->mc, 0);
  } else {
    const GrebCirculationArgs& a = *t->ca;
    const GrebMemberConst& mc = *t->mc;
    float* smem = t->ctx.smem;
    SyncState ss;
    ss.bar = reinterpret_cast<SplitBar*>(smem + GSM_SYNC);
    ss.hb = smem + GSM_HB;
    ss.smem = smem;
    ss.phase = 0;
    const size_t off = (size_t)t->n_idx * GNC;
    if (!ctx_is_helper(t->ctx)) {
      const RowGeom g = row_geom(t->ctx, mc);
      Tile tile;
      tile_load_uv(tile, g, a.uv, a.uv + GNC, smem);
      tile_load_wz(tile, g, a.wz + off, smem);
      tile_load_field(tile, g, a.X_in + off);
      circulation_main(t->ctx, tile, g, mc, ss);
      for (int c = 0; c < GREB_CPT; ++c) {
        const vi idx = g.k * GX + g.col + c;
        v_st(a.dX + off, idx, tile.T[c] - v_ld(a.X_in + off, idx));
      }
    } else {
      const HelperGeom hg = helper_geom(t->ctx, mc);
      HelperRow hr[GREB_HROWS];
      helper_load_uv(hr, hg, a.uv, a.uv + GNC);
      helper_load_wz(hr, hg, a.wz + off);
      circulation_helper(t->ctx, hr, hg, mc, a.X_in + off, ss);
    }
  }
  return nullptr;
}

void run_cta(const GrebKernelArgs* ka, const GrebCirculationArgs* ca, const GrebMemberConst* mc, int n_idx) {
  std::vector<float> smem(GSM_FLOATS + 64, 0.f);
  // SplitBar needs pointer alignment
  float* base = smem.data();
  while (((uintptr_t)(base + GSM_SYNC)) % 16) ++base;
  sb_init(reinterpret_cast<SplitBar*>(base + GSM_SYNC), EMU_UNITS);
  static_assert(sizeof(SplitBar) <= 16 * sizeof(float), "SplitBar fits its shared-memory slot");
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, EMU_UNITS);
  std::vector<WarpTask> tasks(EMU_UNITS);
  std::vector<pthread_t> th(EMU_UNITS);
  for (int w = 0; w < EMU_UNITS; ++w) {
    tasks[w].ctx.warp = w;
    for (int l = 0; l < 32; ++l) tasks[w].ctx.lane_v.v[l] = l;
    tasks[w].ctx.smem = base;
    tasks[w].ctx.late = (w >> 2) & 1;   // stagger (greb_types.h): every other group of four rows
    tasks[w].ctx.bar = &bar;
    tasks[w].ka = ka;
    tasks[w].ca = ca;
    tasks[w].mc = mc;
    tasks[w].n_idx = n_idx;
    pthread_create(&th[w], nullptr, warp_main, &tasks[w]);
  }
  for (int w = 0; w < EMU_UNITS; ++w) pthread_join(th[w], nullptr);
  pthread_barrier_destroy(&bar);
  sb_destroy(reinterpret_cast<SplitBar*>(base + GSM_SYNC));
}

}  // namespace

extern "C" {

// circulation(X_in, dX, h_scl, wz) for one field with member geometry from `p`
int emu_circulation(const greb_physics_par* p, const float* u, const float* v, const float* X_in,
                    const float* wz, float* dX) {
  GrebMemberConst mc;
  greb_build_member_const(mc, *p, 0);
  std::vector<float> uv(2 * GNC);
  memcpy(uv.data(), u, GNC * sizeof(float));
  memcpy(uv.data() + GNC, v, GNC * sizeof(float));
  GrebCirculationArgs ca;
  ca.mc = &mc;
  ca.uv = uv.data();
  ca.X_in = X_in;
  ca.wz = wz;
  ca.dX = dX;
  run_cta(nullptr, &ca, &mc, 0);
  return 0;
}

// row_of_group[48], hslot_of_row[48]; returns greb_build_member_const's status
int emu_row_tables(const greb_physics_par* p, int* row_of_group, int* hslot_of_row) {
  GrebMemberConst mc;
  const int rc = greb_build_member_const(mc, *p, 0);
  memcpy(row_of_group, mc.row_of_group, sizeof mc.row_of_group);
  memcpy(hslot_of_row, mc.hslot_of_row, sizeof mc.hslot_of_row);
  return rc;
}

void* emu_create(const float* z_topo, const float* glacier, const float* sw_solar, const float* tclim,
                 const float* qclim, const float* swetclim, const float* uclim, const float* vclim,
                 const float* mldclim, const float* cldclim, const greb_physics_par* p, const float* co2,
                 int n_years) {
  Emu* e = new Emu;
  greb_build_forcing(e->F, z_topo, glacier, sw_solar, tclim, qclim, swetclim, uclim, vclim, mldclim, cldclim);
  e->phys = *p;
  greb_build_member_const(e->mc, *p, 0);
  e->wz.resize(2 * GNC);
  greb_build_wz(e->wz.data(), e->F, *p);
  e->corr.assign((size_t)GNT * GC_COUNT * GNC, 0.f);
  e->state.resize(GS_COUNT * GNC);
  greb_build_initial_state(e->state.data(), e->F, e->mc);
  e->acc.assign(GA_COUNT * GNC, 0.f);
  e->co2.assign(co2, co2 + n_years);
  e->diag.assign(2, 0.f);
  e->flags.assign(1, 0);
  return e;
}

void emu_destroy(void* h) { delete (Emu*)h; }

// runs steps it0 .. it0+nsteps-1; out (may be NULL) [out_months][5][GNC]
int emu_steps(void* h, int it0, int nsteps, int spinup, float* out, int out_months) {
  Emu* e = (Emu*)h;
  GrebKernelArgs a;
  memset(&a, 0, sizeof a);
  a.mc = &e->mc;
  a.member_ids = &e->member_id;
  a.forc = e->F.forc.data();
  a.sw_solar = e->F.sw_solar.data();
  a.mask = e->F.mask.data();
  a.z_ocean = e->F.z_ocean.data();
  a.toclim = e->F.toclim.data();
  a.tclim = e->F.tclim.data();
  a.qclim = e->F.qclim.data();
  a.wz = e->wz.data();
  a.corr = e->corr.data();
  a.state = e->state.data();
  a.acc = e->acc.data();
  a.out = out;
  a.co2 = e->co2.data();
  a.diag = e->diag.data();
  a.coslat_w = e->F.coslat_w.data();
  a.flags = e->flags.data();
  a.co2_stride = (int)e->co2.size();
  a.out_months = out_months;
  a.it0 = it0;
  a.nsteps = nsteps;
  a.spinup = spinup;
  run_cta(&a, nullptr, &e->mc, 0);
  return 0;
}

void emu_set_switches(void* h, unsigned mask) { ((Emu*)h)->mc.switches = (int)mask; }
void emu_reset_scenario(void* h) {
  Emu* e = (Emu*)h;
  std::fill(e->acc.begin(), e->acc.end(), 0.f);
}
void emu_get_state(void* h, int which, float* out) { memcpy(out, &((Emu*)h)->state[(size_t)which * GNC], GNC * 4); }
void emu_set_state(void* h, int which, const float* in) { memcpy(&((Emu*)h)->state[(size_t)which * GNC], in, GNC * 4); }
void emu_get_corr(void* h, float* out) { memcpy(out, ((Emu*)h)->corr.data(), ((Emu*)h)->corr.size() * 4); }
void emu_get_diag(void* h, float* out) { memcpy(out, ((Emu*)h)->diag.data(), 8); }
}
