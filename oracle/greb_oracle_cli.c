/*
 * greb_oracle_cli.c — `./greb [namelist]` work-alike around the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY (CPU baseline: "one ./greb process per host core").
 * Follows PROGRAM greb_run, /root/reference/src/greb.f90:996-1098: opens input/<files>
 * relative to the working directory, reads the four namelist groups, pads co2_ppm
 * (f:1053-1061), builds output_file[_ens_id] (f:1064-1068), derives Toclim, runs greb_model
 * and writes the raw fp32 record stream the reference writes on unit 22 (f:978-982).
 *
 * Extra, non-reference options (after the namelist path):
 *   --input DIR     directory holding the ten input files (default "input")
 *   --no-output     do not write the output file (timing runs)
 *   --time          print "oracle_seconds spinup=<s> scenario=<s>" at the end
 */
#define _POSIX_C_SOURCE 200809L
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>

#include "greb_oracle.h"

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static float *read_file(const char *dir, const char *name, size_t nfloat) {
  char path[1024];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE *f = fopen(path, "rb");
  if (!f) {
    fprintf(stderr, "greb_oracle: cannot open %s\n", path);
    exit(2);
  }
  float *buf = (float *)malloc(nfloat * sizeof(float));
  if (fread(buf, sizeof(float), nfloat, f) != nfloat) {
    fprintf(stderr, "greb_oracle: short read on %s\n", path);
    exit(2);
  }
  fclose(f);
  return buf;
}

/* ---- minimal Fortran namelist reader ------------------------------------------------------ */

typedef struct {
  go_physics phys;
  int ipx, ipy, time_flux, time_scnr, year0;
  char output_file[121], ens_id[11];
  float co2_ppm[4096];
  int n_co2;
} config;

static void set_value(config *c, const char *group, const char *name, const char *vals) {
  /* vals: comma/space separated list */
  float fv[64];
  int nf = 0;
  char sv[256] = "";
  {
    const char *p = vals;
    while (*p && isspace((unsigned char)*p)) ++p;
    if (*p == '"' || *p == '\'') {
      char q = *p++;
      size_t n = 0;
      while (*p && *p != q && n < sizeof sv - 1) sv[n++] = *p++;
      sv[n] = 0;
    } else {
      char tmp[4096];
      strncpy(tmp, vals, sizeof tmp - 1);
      tmp[sizeof tmp - 1] = 0;
      for (char *t = tmp; *t; ++t)
        if (*t == ',' || *t == '(' || *t == ')' || *t == '/') *t = ' ';
      char *save = NULL;
      for (char *tok = strtok_r(tmp, " \t\r\n", &save); tok && nf < 64; tok = strtok_r(NULL, " \t\r\n", &save)) {
        for (char *t = tok; *t; ++t)
          if (*t == 'd' || *t == 'D') *t = 'e';
        fv[nf++] = strtof(tok, NULL);
      }
    }
  }
#define PF(x) if (!strcasecmp(name, #x)) { if (nf) c->phys.x = fv[0]; return; }
  if (!strcasecmp(group, "physics_par")) {
    PF(pi) PF(sig) PF(rho_ocean) PF(rho_land) PF(rho_air) PF(cp_ocean) PF(cp_land) PF(cp_air) PF(eps)
    PF(d_ocean) PF(d_land) PF(d_air) PF(ct_sens) PF(da_ice) PF(a_no_ice) PF(a_cloud) PF(Tl_ice1)
    PF(Tl_ice2) PF(To_ice1) PF(To_ice2) PF(co_turb) PF(kappa) PF(ce) PF(cq_latent) PF(cq_rain)
    PF(z_air) PF(z_vapor) PF(r_qviwv)
    if (!strcasecmp(name, "p_emi")) {
      for (int i = 0; i < nf && i < 10; ++i) c->phys.p_emi[i] = fv[i];
      return;
    }
  } else if (!strcasecmp(group, "numerics_par")) {
#define PI_(x) if (!strcasecmp(name, #x)) { if (nf) c->x = (int)fv[0]; return; }
    PI_(ipx) PI_(ipy) PI_(time_flux) PI_(time_scnr) PI_(year0)
  } else if (!strcasecmp(group, "diagnostics_par")) {
    if (!strcasecmp(name, "output_file")) { strncpy(c->output_file, sv, 120); return; }
    if (!strcasecmp(name, "ens_id")) { strncpy(c->ens_id, sv, 10); return; }
  } else if (!strcasecmp(group, "co2_par")) {
    if (!strcasecmp(name, "co2_flux")) { if (nf) c->phys.co2_flux = fv[0]; return; }
    if (!strcasecmp(name, "co2_ppm")) {
      for (int i = 0; i < nf && i < 4096; ++i) c->co2_ppm[i] = fv[i];
      c->n_co2 = nf;
      return;
    }
  }
  fprintf(stderr, "greb_oracle: unknown namelist entry %s / %s\n", group, name);
  exit(2);
}

static void read_namelist(const char *path, config *c) {
  FILE *f = fopen(path, "r");
  if (!f) {
    fprintf(stderr, "greb_oracle: cannot open namelist %s\n", path);
    exit(2);
  }
  char line[8192], group[64] = "";
  char name[64] = "", vals[8192] = "";
  while (fgets(line, sizeof line, f)) {
    /* strip comments outside quotes */
    int inq = 0;
    for (char *p = line; *p; ++p) {
      if (*p == '"' || *p == '\'') inq = !inq;
      if (*p == '!' && !inq) { *p = 0; break; }
    }
    char *p = line;
    while (*p && isspace((unsigned char)*p)) ++p;
    if (!*p) continue;
    if (*p == '&') {
      sscanf(p + 1, "%63s", group);
      continue;
    }
    if (*p == '/') {
      if (name[0]) set_value(c, group, name, vals);
      name[0] = 0; vals[0] = 0; group[0] = 0;
      continue;
    }
    char *eq = strchr(p, '=');
    if (eq) {
      if (name[0]) set_value(c, group, name, vals);
      *eq = 0;
      sscanf(p, "%63s", name);
      strncpy(vals, eq + 1, sizeof vals - 1);
    } else {
      strncat(vals, " ", sizeof vals - strlen(vals) - 1);
      strncat(vals, p, sizeof vals - strlen(vals) - 1); /* continuation of an array value */
    }
  }
  if (name[0]) set_value(c, group, name, vals);
  fclose(f);
}

int main(int argc, char **argv) {
  config c;
  memset(&c, 0, sizeof c);
  go_physics_defaults(&c.phys);
  c.ipx = 1; c.ipy = 1; c.time_flux = 0; c.time_scnr = 0; c.year0 = 1940; /* f:49-53 */
  strcpy(c.output_file, "output/scenario");                              /* f:152 */
  const char *nml = "namelist"; /* f:1033-1034 */
  const char *indir = "input";
  int no_output = 0, timing = 0;
  for (int a = 1; a < argc; ++a) {
    if (!strcmp(argv[a], "--input") && a + 1 < argc) indir = argv[++a];
    else if (!strcmp(argv[a], "--no-output")) no_output = 1;
    else if (!strcmp(argv[a], "--time")) timing = 1;
    else nml = argv[a];
  }
  read_namelist(nml, &c);

  /* f:1047-1061 co2 padding */
  int nyr = c.time_scnr;
  float *co2 = (float *)malloc(sizeof(float) * (nyr > 0 ? nyr : 1));
  for (int i = 0; i < nyr; ++i) co2[i] = (i < c.n_co2) ? c.co2_ppm[i] : -1.f;
  if (nyr > 0 && co2[0] == -1.f) co2[0] = 680.f;
  for (int i = 1; i < nyr; ++i)
    if (co2[i] < 0.f) {
      for (int j = i; j < nyr; ++j) co2[j] = co2[i - 1];
      break;
    }

  char outpath[160];
  if (strlen(c.ens_id) == 0) snprintf(outpath, sizeof outpath, "%s", c.output_file);
  else snprintf(outpath, sizeof outpath, "%s_%s", c.output_file, c.ens_id); /* f:1064-1068 */

  printf(" %% diagonstic point lat/lon:  %g %g\n", 3.75 * c.ipy - 90, 3.75 * c.ipx); /* f:1070 */

  const size_t F = GO_NCELL, N = GO_NSTEP_YR;
  float *z_topo = read_file(indir, "topography", F);
  float *sw_solar = read_file(indir, "solar.radiation", (size_t)GO_YDIM * N);
  float *glacier = read_file(indir, "glacier.masks", F);
  float *tclim = read_file(indir, "tsurf", F * N);
  float *qclim = read_file(indir, "vapor", F * N);
  float *swet = read_file(indir, "soil.moisture", F * N);
  float *ucl = read_file(indir, "zonal.wind", F * N);
  float *vcl = read_file(indir, "meridional.wind", F * N);
  float *mld = read_file(indir, "ocean.mld", F * N);
  float *cld = read_file(indir, "cloud.cover", F * N);

  go_model *m = go_create();
  go_set_physics(m, &c.phys);
  go_set_forcing(m, z_topo, glacier, sw_solar, tclim, qclim, swet, ucl, vcl, mld, cld);
  go_setup(m);

  printf(" %% FLUX CORRECTION RUN; years = %d  co2 = %g\n", c.time_flux, c.phys.co2_flux); /* f:219 */
  double t0 = now();
  go_qflux_correction(m, c.time_flux, c.phys.co2_flux);
  double t1 = now();
  printf(" %% MODEL RUN; years = %d\n %% saving output in file %s\n", c.time_scnr, outpath); /* f:224-225 */
  printf(" console output: year, co2, global avg temp, avg temp for ipx/ipy\n");          /* f:941 */

  float *out = NULL, *gmean = (float *)calloc(nyr > 0 ? nyr : 1, sizeof(float));
  if (!no_output) out = (float *)malloc(sizeof(float) * F * 60 * (size_t)(nyr > 0 ? nyr : 1));
  go_run_scenario(m, nyr, co2, c.year0, out, gmean, 0);
  double t2 = now();
  for (int y = 0; y < nyr; ++y) printf(" %12.4f %12.4f %12.5f\n", (double)(c.year0 + y), co2[y], gmean[y]); /* f:954 */
  if (out) {
    FILE *f = fopen(outpath, "r+b"); /* the reference opens without status='replace' (f:174) */
    if (!f) f = fopen(outpath, "wb");
    if (!f) {
      fprintf(stderr, "greb_oracle: cannot write %s\n", outpath);
      return 2;
    }
    fwrite(out, sizeof(float), F * 60 * (size_t)nyr, f);
    fclose(f);
  }
  if (timing) printf("oracle_seconds spinup=%.3f scenario=%.3f\n", t1 - t0, t2 - t1);
  go_destroy(m);
  return 0;
}
