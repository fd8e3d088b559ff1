/* A compiled, non-Python consumer of the C ABI (include/fv3lm_b200.h): the closest stand-in available here for the Fortran
 * ISO_C_BINDING shim (fv3-jedi-linearmodel_b200/fortran/fv3jedi_lm_dynamics_mod.F90, reference
 * src/dynamics/fv3jedi_lm_dynamics_mod.F90:69-689), which cannot be compiled in this image (no Fortran compiler).
 *
 * It replays   create -> set_metric* -> set_phis -> traj_set -> step_nl -> traj_get -> step_tl -> step_ad -> destroy
 * on a file of inputs written by tests/test_c_driver.py and writes the results to a second file; the test compares them with what
 * the ctypes binding returns for the same inputs.  TEST INFRASTRUCTURE: links the library under test, nothing else.
 *
 * input file (little endian):  "FV3LMIN1", sizeof(fv3lm_config), the struct bytes, then records until EOF:
 *   char name[32]; int32 kind; int64 n; double data[n]
 *   kind 0 = 2-D metric, 1 = 1-D metric, 2 = metric scalar, 3 = ak, 4 = bk, 5 = phis, 6 = trajectory field, 7 = increment field,
 *        8 = adjoint test vector field
 * output file: records  char name[32]; int64 n; double data[n]  named nl.<f>, tl.<f>, ad.<f>
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/fv3lm_b200.h"

#define NF 10
static const char* fnames[NF] = {"u", "v", "t", "delp", "qv", "ql", "qi", "o3", "w", "delz"};

static double** slot_of(fv3lm_fields* f, int n) {
  switch (n) {
    case 0: return &f->u; case 1: return &f->v; case 2: return &f->t; case 3: return &f->delp; case 4: return &f->qv;
    case 5: return &f->ql; case 6: return &f->qi; case 7: return &f->o3; case 8: return &f->w; default: return &f->delz;
  }
}
static int field_index(const char* nm) { for (int n = 0; n < NF; n++) if (!strcmp(nm, fnames[n])) return n; return -1; }

static void die(fv3lm_handle* h, const char* what) {
  fprintf(stderr, "c_driver: %s failed: %s\n", what, fv3lm_last_error(h));
  exit(2);
}
static void put(FILE* fo, const char* pre, const char* nm, const double* a, int64_t n) {
  char name[32]; memset(name, 0, sizeof name); snprintf(name, sizeof name, "%s.%s", pre, nm);
  fwrite(name, 1, 32, fo); fwrite(&n, 8, 1, fo); fwrite(a, 8, (size_t)n, fo);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: c_driver input.bin output.bin\n"); return 1; }
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) { perror("input"); return 1; }
  char magic[8]; int32_t csize = 0;
  if (fread(magic, 1, 8, fi) != 8 || memcmp(magic, "FV3LMIN1", 8)) { fprintf(stderr, "bad magic\n"); return 1; }
  if (fread(&csize, 4, 1, fi) != 1 || csize != (int32_t)sizeof(fv3lm_config)) {
    fprintf(stderr, "config size mismatch: file %d, header %zu\n", csize, sizeof(fv3lm_config)); return 1;
  }
  fv3lm_config cfg;
  if (fread(&cfg, sizeof cfg, 1, fi) != 1) return 1;
  /* first pass: read every record into memory (create needs ak / bk before anything else can be uploaded) */
  enum { MAXREC = 256 };
  static char rname[MAXREC][32]; static int32_t rkind[MAXREC]; static int64_t rn[MAXREC]; static double* rdata[MAXREC];
  int nrec = 0;
  while (nrec < MAXREC && fread(rname[nrec], 1, 32, fi) == 32) {
    if (fread(&rkind[nrec], 4, 1, fi) != 1 || fread(&rn[nrec], 8, 1, fi) != 1) return 1;
    rdata[nrec] = (double*)malloc((size_t)rn[nrec] * 8);
    if (fread(rdata[nrec], 8, (size_t)rn[nrec], fi) != (size_t)rn[nrec]) { fprintf(stderr, "short record %s\n", rname[nrec]); return 1; }
    nrec++;
  }
  fclose(fi);
  const double *ak = NULL, *bk = NULL;
  for (int r = 0; r < nrec; r++) { if (rkind[r] == 3) ak = rdata[r]; if (rkind[r] == 4) bk = rdata[r]; }
  fv3lm_handle* h = NULL;
  if (fv3lm_create(&cfg, ak, bk, &h)) die(NULL, "fv3lm_create");
  fv3lm_fields traj, pert, yvec; memset(&traj, 0, sizeof traj); memset(&pert, 0, sizeof pert); memset(&yvec, 0, sizeof yvec);
  int64_t nfield = 0;
  for (int r = 0; r < nrec; r++) {
    if (rkind[r] == 0 || rkind[r] == 1) { if (fv3lm_set_metric(h, rname[r], rdata[r], rkind[r])) die(h, rname[r]); }
    else if (rkind[r] == 2) { if (fv3lm_set_metric_scalar(h, rname[r], rdata[r][0])) die(h, rname[r]); }
    else if (rkind[r] >= 6) {
      int f = field_index(rname[r]);
      if (f < 0) { fprintf(stderr, "unknown field %s\n", rname[r]); return 1; }
      *slot_of(rkind[r] == 6 ? &traj : rkind[r] == 7 ? &pert : &yvec, f) = rdata[r];
      nfield = rn[r];
    }
  }
  for (int r = 0; r < nrec; r++) if (rkind[r] == 5 && fv3lm_set_phis(h, rdata[r])) die(h, "fv3lm_set_phis");
  if (fv3lm_traj_set(h, 0, &traj)) die(h, "fv3lm_traj_set");
  FILE* fo = fopen(argv[2], "wb");
  if (!fo) { perror("output"); return 1; }
  const int nf = cfg.hydrostatic ? 8 : 10;
  /* step_nl: slot 0 -> slot 1 */
  if (fv3lm_step_nl(h, 0, 1)) die(h, "fv3lm_step_nl");
  fv3lm_fields out; memset(&out, 0, sizeof out);
  for (int f = 0; f < nf; f++) *slot_of(&out, f) = (double*)malloc((size_t)nfield * 8);
  if (fv3lm_traj_get(h, 1, &out)) die(h, "fv3lm_traj_get");
  for (int f = 0; f < nf; f++) put(fo, "nl", fnames[f], *slot_of(&out, f), nfield);
  /* step_tl / step_ad work in place on the caller's arrays, like the reference's pert members */
  if (fv3lm_step_tl(h, 0, &pert)) die(h, "fv3lm_step_tl");
  for (int f = 0; f < nf; f++) put(fo, "tl", fnames[f], *slot_of(&pert, f), nfield);
  if (fv3lm_step_ad(h, 0, &yvec)) die(h, "fv3lm_step_ad");
  for (int f = 0; f < nf; f++) put(fo, "ad", fnames[f], *slot_of(&yvec, f), nfield);
  fclose(fo);
  /* error convention: a bad call returns non-zero and leaves a message, it does not abort */
  if (fv3lm_traj_set(h, -1, &traj) == 0) { fprintf(stderr, "negative slot was accepted\n"); return 3; }
  if (!fv3lm_last_error(h) || !fv3lm_last_error(h)[0]) { fprintf(stderr, "no error message\n"); return 3; }
  if (fv3lm_destroy(h)) { fprintf(stderr, "destroy failed\n"); return 3; }
  printf("c_driver ok: %d fields of %lld doubles through create / set_metric / traj_set / step_nl / step_tl / step_ad / destroy\n", nf, (long long)nfield);
  return 0;
}
