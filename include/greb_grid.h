/* greb_grid.h — C ABI of the big-grid circulation path (BASELINE.json configs[4]: one member on a
 * grid too large for one SM, decomposed in latitude bands over several GPUs).
 *
 * Replaces, for an arbitrary xdim x ydim grid: circulation / diffusion / advection of the reference
 * (src/greb.f90:528-553, 556-723, 726-915), i.e. the part of a 12-hour step that needs neighbour
 * data.  One handle = one latitude band [k0, k1) of the global grid on one GPU, stored with
 * `halo_rows` extra rows on each inner side.  A sub-step needs rows k-2..k+2 (f:587-590, 771-780),
 * so after an exchange of 2*s halo rows the band can advance s sub-steps without communication
 * (the updated range shrinks by 2 rows per sub-step; the redundant halo rows are recomputed).
 *
 * The reference formulas cannot run at 0.25 degrees as written.  Declared rules,
 * both inactive on the reference's 96x48 grid (there the results are bit-identical to the
 * reference arithmetic):  R1  dt_crcl = 1800*(48/ydim)^2 s;  R2  the latitude entering dxlat is
 * clamped to +-88.125 degrees and dtdff2 is floored at 1 s.  (The test suite carries a CPU restatement of both.)
 *
 * All functions return 0 or a negative error code; greb_grid_last_error gives the message.
 * No CPU fallback: greb_grid_create fails without an sm_100 device. */
#ifndef GREB_GRID_H
#define GREB_GRID_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct greb_grid_handle_s* greb_grid_t;

int greb_grid_create(greb_grid_t* h, int nx, int ny, int k0, int k1, int halo_rows, int device);
int greb_grid_destroy(greb_grid_t h);
const char* greb_grid_last_error(greb_grid_t h);

/* geometry of f:543, 578-582, 652-654, 749-753, 838-840 under R1/R2; *nsub = sub-steps per 12-h step */
int greb_grid_set_geometry(greb_grid_t h, float pi, float kappa, int* nsub, float* dt_crcl);

/* FULL global host fields [ydim][xdim] (the band and its halos are cut out inside):
 * X = the circulating field (Ta or q), wz = wz_air or wz_vapor, u/v = the wind climatology of the step */
int greb_grid_set_fields(greb_grid_t h, const float* X, const float* wz, const float* u, const float* v);

/* the wind climatology of another step (FULL global host fields); the circulating field, its level and the
 * halo bookkeeping are untouched — how a host steps through the calendar (src/greb.f90:251-252) */
int greb_grid_set_winds(greb_grid_t h, const float* u, const float* v);

/* n more sub-steps X = (X + dx_diffuse) + dx_advec (f:546-549) without communication; fails if the
 * halo is too thin for n (exchange first). */
int greb_grid_substeps(greb_grid_t h, int n);

/* the same without waiting: the sub-steps are queued on the handle's stream, so the host can
 * exchange another field's halos meanwhile; greb_grid_sync waits for them (and must be called
 * before the band's rows are read or its halos rewritten) */
int greb_grid_substeps_async(greb_grid_t h, int n);
int greb_grid_sync(greb_grid_t h);

/* device view for the halo exchange: pointer to local row 0 of the CURRENT field buffer, the
 * global index of that row, the number of local rows, and the global row range that is valid */
int greb_grid_view(greb_grid_t h, float** dev_rows, int* kbase, int* nrows, int* valid_lo, int* valid_hi);
/* call after the neighbours' rows were written into the halo rows of the current buffer */
int greb_grid_halo_refreshed(greb_grid_t h);

/* ---- persistent path: no kernel launch per sub-step, no host in the exchange ---------------------------
 * greb_grid_run_persistent advances a GROUP of one or two handles of the same band (the two independent
 * fields of a step, air temperature and humidity, src/greb.f90:299-304) by n sub-steps in ONE cooperative
 * launch: work item = (field, row), a grid barrier between sub-steps, and the halo exchange done by the
 * kernel itself — the CTA that computes one of the band's two outermost rows stores the new row also into
 * the neighbour band's halo row (peer memory over NVLink) and releases a per-row level flag there; the CTA
 * that computes a row next to a halo acquires those flags first.  Needs halo_rows >= 2 (exactly the reach
 * of a sub-step, f:587-590, 771-780: nothing is recomputed redundantly).
 * The neighbours' buffers are mapped with CUDA IPC: every rank exports a blob (greb_grid_ipc_bytes bytes)
 * per handle, the host side swaps the blobs (torch.distributed all_gather_object in greb_b200/bigrid.py)
 * and imports its south (side 0) and north (side 1) neighbour's.  All ranks must call set_fields, then
 * synchronise (a host barrier), then call run_persistent with the same n; a rank that never arrives makes
 * the others' waits time out (error return) instead of hanging the GPU. */
int greb_grid_ipc_bytes(void);
int greb_grid_ipc_export(greb_grid_t h, void* blob);
int greb_grid_ipc_import(greb_grid_t h, int side, const void* neighbour_blob);
int greb_grid_run_persistent(greb_grid_t* handles, int nfields, int n);

/* the band's own rows [k0, k1) -> host [k1-k0][xdim] */
int greb_grid_get(greb_grid_t h, float* out);
/* CUDA-event time of the last greb_grid_substeps call and its kernel launches */
int greb_grid_last_ms(greb_grid_t h, float* ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif
