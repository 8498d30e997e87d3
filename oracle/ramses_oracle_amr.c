/*
 * ramses_oracle_amr.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * The per-cell passes of the reference's mesh-adaptation control flow (amr/flag_utils.f90, hydro/hydro_flag.f90,
 * amr/physical_boundaries.f90, the scan loops of amr/refine_utils.f90:332-585) on the oracle's tree arrays.  They are
 * the same passes as the pure-Python ones of oracle/amr.py (AmrRun) -- which stay as the readable statement and as a
 * cross-check: tests/test_oracle_golden.py runs both on the sod-tube golden case and requires identical results --
 * moved to C so that the long 2-D golden run (tests/hydro/implosion, ~1e4 level steps on ~4000 octs) finishes in
 * minutes.  Only integer maps are written here; all floating-point evolution stays in ramses_oracle.c.
 *
 * Citations "file:line" are relative to the reference tree.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ramses_oracle.h"

static inline int amr_cell(const orc_mesh* m, int ind, int ig) { return m->ncoarse + ind * m->ngridmax + ig; }
static inline int amr_nbor(const orc_mesh* m, int j, int ig) { return m->nbor[(size_t)j * (m->ngridmax + 1) + ig]; }

/* getnborgrids amr/nbors_utils.f90:530: gn[0]=ig, gn[j+1]=son(nbor(ig,j)) */
static void amr_nbor_grids(const orc_mesh* m, int ig, int* gn) {
  gn[0] = ig;
  for (int j = 0; j < 2 * m->ndim; j++) {
    int c = amr_nbor(m, j, ig);
    gn[j + 1] = c > 0 ? m->son[c] : 0;
  }
}

/* getnborcells amr/nbors_utils.f90:363 (the ggg/hhh tables written as bit operations): neighbour CELL of cell `ind`
 * of oct gn[0] in direction (d, s); 0 when the neighbouring oct does not exist                                       */
static void amr_nbor_cells(const orc_mesh* m, const int* gn, int ind, int* out) {
  for (int d = 0; d < m->ndim; d++)
    for (int s = 0; s < 2; s++) {
      int bit = (ind >> d) & 1;
      int ind2 = ind ^ (1 << d);
      int g = bit != s ? gn[0] : gn[2 * d + s + 1];
      out[2 * d + s] = g > 0 ? amr_cell(m, ind2, g) : 0;
    }
}

/* make_boundary_flag amr/physical_boundaries.f90:259 */
void orc_amr_make_boundary_flag(const orc_mesh* m, int l, int* flag1) {
  static const int inbor_of[7] = {0, 2, 1, 4, 3, 6, 5};
  int T = 1 << m->ndim;
  for (int b = 0; b < m->nboundary; b++) {
    int bt = m->boundary_type[b];
    int bdir = bt - 10 * (bt / 10);
    int inbor = inbor_of[bdir];
    int d = (bdir - 1) / 2;
    for (int i = 0; i < m->nbound[b][l]; i++) {
      int ig = m->bound[b][l][i];
      int cn = amr_nbor(m, inbor - 1, ig);
      int gref = cn > 0 ? m->son[cn] : 0;
      for (int ind = 0; ind < T; ind++) {
        int indr;
        if (bt / 10 == 0) indr = ind ^ (1 << d);                                  /* reflexive: mirror cell        */
        else indr = (ind & ~(1 << d)) | ((bdir % 2 == 1 ? 0 : 1) << d);           /* zero gradient: adjacent cell  */
        flag1[amr_cell(m, ind, ig)] = gref > 0 ? flag1[amr_cell(m, indr, gref)] : 0;
      }
    }
  }
}

/* init_flag amr/flag_utils.f90:109-195 (+ test_flag) */
void orc_amr_init_flag(const orc_mesh* m, int l, int levelmin, int* flag1) {
  int T = 1 << m->ndim;
  int n = m->nactive[l];
  const int* act = m->active[l];
  for (int i = 0; i < n; i++)
    for (int ind = 0; ind < T; ind++) flag1[amr_cell(m, ind, act[i])] = 0;
  if (l >= levelmin) {
    for (int ind = 0; ind < T; ind++)
      for (int i = 0; i < n; i++) {
        int c = amr_cell(m, ind, act[i]);
        int gs = m->son[c];
        int ok = 0;
        if (gs > 0)
          for (int inds = 0; inds < T; inds++) {
            int cs = amr_cell(m, inds, gs);
            ok = ok || m->son[cs] > 0 || flag1[cs] == 1;
          }
        if (ok) flag1[c] = 1;
      }
  } else {
    for (int i = 0; i < n; i++)
      for (int ind = 0; ind < T; ind++) flag1[amr_cell(m, ind, act[i])] = 1;
  }
  orc_amr_make_boundary_flag(m, l, flag1);
}

/* smooth_fine amr/flag_utils.f90:556-632 (the caller tests ilevel==nlevelmax / numbtot==0) */
void orc_amr_smooth_fine(const orc_mesh* m, int l, int* flag1, int* flag2) {
  static const int n_nbor[3] = {1, 2, 2};
  int T = 1 << m->ndim, tw = 2 * m->ndim;
  int n = m->nactive[l];
  const int* act = m->active[l];
  int gn[7], nc[6];
  flag1[0] = 0;
  for (int ismooth = 0; ismooth < m->ndim; ismooth++) {
    for (int i = 0; i < n; i++)
      for (int ind = 0; ind < T; ind++) flag2[amr_cell(m, ind, act[i])] = 0;
    for (int i = 0; i < n; i++) {
      amr_nbor_grids(m, act[i], gn);
      for (int ind = 0; ind < T; ind++) {
        amr_nbor_cells(m, gn, ind, nc);
        int cnt = 0;
        for (int j = 0; j < tw; j++) cnt += flag1[nc[j]];
        if (cnt >= n_nbor[ismooth]) flag2[amr_cell(m, ind, act[i])] = 1;
      }
    }
    for (int i = 0; i < n; i++)
      for (int ind = 0; ind < T; ind++) {
        int c = amr_cell(m, ind, act[i]);
        if (flag1[c] == 1) flag2[c] = 0;
        if (flag2[c] == 1) flag1[c] = 1;
      }
    orc_amr_make_boundary_flag(m, l, flag1);
  }
}

/* hydro_refine hydro/godunov_utils.f90:125-263 for one (left, centre, right) triple; u are conservative, nvar=ndim+2 */
static int amr_hydro_refine(const orc_params* p, int ndim, const double* ug_, const double* um_, const double* ud_,
                            const double err[3], const double flo[3]) {
  double u[3][8];
  const double* src[3] = {ug_, um_, ud_};
  int ip = ndim + 1;
  for (int s = 0; s < 3; s++) {
    for (int v = 0; v < ndim + 2; v++) u[s][v] = src[s][v];
    u[s][0] = u[s][0] > p->smallr ? u[s][0] : p->smallr;
    for (int d = 0; d < ndim; d++) u[s][d + 1] = u[s][d + 1] / u[s][0];
    double ek = 0.0;
    for (int d = 0; d < ndim; d++) ek = ek + 0.5 * u[s][0] * (u[s][d + 1] * u[s][d + 1]);
    u[s][ip] = (p->gamma - 1.0) * (u[s][ip] - ek);
  }
  const double *g = u[0], *c = u[1], *d_ = u[2];
  int ok = 0;
  if (err[0] >= 0.0) {
    double a = fabs((d_[0] - c[0]) / (d_[0] + c[0] + flo[0])), b = fabs((c[0] - g[0]) / (c[0] + g[0] + flo[0]));
    double e = 2.0 * (a > b ? a : b);
    ok = ok || e > err[0];
  }
  if (err[2] >= 0.0) {
    double a = fabs((d_[ip] - c[ip]) / (d_[ip] + c[ip] + flo[2])), b = fabs((c[ip] - g[ip]) / (c[ip] + g[ip] + flo[2]));
    double e = 2.0 * (a > b ? a : b);
    ok = ok || e > err[2];
  }
  if (err[1] >= 0.0) {
    double f2 = flo[1] * flo[1];
    for (int k = 0; k < ndim; k++) {
      double vg = g[k + 1], vm = c[k + 1], vd = d_[k + 1];
      double tg = p->gamma * g[ip] / g[0], tm = p->gamma * c[ip] / c[0], td = p->gamma * d_[ip] / d_[0];
      double cg = sqrt(tg > f2 ? tg : f2), cm = sqrt(tm > f2 ? tm : f2), cd = sqrt(td > f2 ? td : f2);
      double a = fabs((vd - vm) / (cd + cm + fabs(vd) + fabs(vm) + flo[1]));
      double b = fabs((vm - vg) / (cm + cg + fabs(vm) + fabs(vg) + flo[1]));
      double e = 2.0 * (a > b ? a : b);
      ok = ok || e > err[1];
    }
  }
  return ok;
}

/* hydro_flag hydro/hydro_flag.f90:1 -- err = (err_grad_d, err_grad_u, err_grad_p), flo = (floor_d, floor_u, floor_p) */
void orc_amr_hydro_flag(const orc_params* p, const orc_mesh* m, int l, const double* uold, int* flag1,
                        const double err[3], const double flo[3]) {
  int ndim = m->ndim, T = 1 << ndim, tw = 2 * ndim, nvar = ndim + 2;
  int n = m->nactive[l];
  const int* act = m->active[l];
  int gn[7], nc[6];
  double ug[8], um[8], ud[8];
  size_t ncell = (size_t)m->ncell;
  for (int i = 0; i < n; i++) {
    int ig = act[i];
    amr_nbor_grids(m, ig, gn);
    for (int ind = 0; ind < T; ind++) {
      int c = amr_cell(m, ind, ig);
      amr_nbor_cells(m, gn, ind, nc);
      for (int j = 0; j < tw; j++)
        if (nc[j] == 0) nc[j] = amr_nbor(m, j, ig);                 /* coarser neighbour: its father cell          */
      int ok = 0;
      for (int d = 0; d < ndim; d++) {
        for (int v = 0; v < nvar; v++) {
          ug[v] = uold[v * ncell + nc[2 * d] - 1];
          um[v] = uold[v * ncell + c - 1];
          ud[v] = uold[v * ncell + nc[2 * d + 1] - 1];
        }
        ok = ok || amr_hydro_refine(p, ndim, ug, um, ud, err, flo);
      }
      if (ok) flag1[c] = 1;
    }
  }
}

/* ensure_ref_rules amr/flag_utils.f90:197-255 */
void orc_amr_ensure_ref_rules(const orc_mesh* m, int l, int* flag1) {
  int T = 1 << m->ndim;
  int n3 = 1;
  for (int d = 0; d < m->ndim; d++) n3 *= 3;
  int nfc[27], nfg[8];
  for (int i = 0; i < m->nactive[l]; i++) {
    int ig = m->active[l][i];
    orc_get3cubefather(m, m->father[ig], l, nfc, nfg);
    int ok = 1;
    for (int j = 0; j < n3; j++)
      if (nfc[j] == 0 || m->son[nfc[j]] == 0) ok = 0;
    if (!ok)
      for (int ind = 0; ind < T; ind++) flag1[amr_cell(m, ind, ig)] = 0;
  }
  orc_amr_make_boundary_flag(m, l, flag1);
}

/* authorize_fine amr/virtual_boundaries.f90:34 (serial) + init_boundary_fine amr/physical_boundaries.f90:106 */
void orc_amr_authorize_fine(const orc_mesh* m, int l, int* flag1, int* flag2) {
  static const int n_nbor[3] = {1, 2, 3};
  int T = 1 << m->ndim, tw = 2 * m->ndim;
  int gn[7], nc[6];
  for (int i = 0; i < m->nactive[l]; i++)
    for (int ind = 0; ind < T; ind++) flag2[amr_cell(m, ind, m->active[l][i])] = 1;
  flag2[0] = 0;
  for (int b = 0; b < m->nboundary; b++)
    for (int i = 0; i < m->nbound[b][l]; i++)
      for (int ind = 0; ind < T; ind++) flag2[amr_cell(m, ind, m->bound[b][l][i])] = 0;
  if (m->nactive[l] > 0) {
    for (int ismooth = 0; ismooth < m->ndim; ismooth++) {
      for (int b = 0; b < m->nboundary; b++)
        for (int i = 0; i < m->nbound[b][l]; i++)
          for (int ind = 0; ind < T; ind++) flag1[amr_cell(m, ind, m->bound[b][l][i])] = 0;
      for (int b = 0; b < m->nboundary; b++)
        for (int i = 0; i < m->nbound[b][l]; i++) {
          int ig = m->bound[b][l][i];
          amr_nbor_grids(m, ig, gn);
          for (int ind = 0; ind < T; ind++) {
            amr_nbor_cells(m, gn, ind, nc);
            int cnt = 0;
            for (int j = 0; j < tw; j++) cnt += flag2[nc[j]];
            if (cnt >= n_nbor[ismooth]) flag1[amr_cell(m, ind, ig)] = 1;
          }
        }
      for (int b = 0; b < m->nboundary; b++)
        for (int i = 0; i < m->nbound[b][l]; i++)
          for (int ind = 0; ind < T; ind++) {
            int c = amr_cell(m, ind, m->bound[b][l][i]);
            if (flag1[c] == 1) flag2[c] = 1;
          }
    }
  }
  orc_amr_make_boundary_flag(m, l, flag1);
}

/* the two scan loops of refine_fine amr/refine_utils.f90:332-585 over ONE list (kind 0 active, 1 boundary region b):
 * mode 0: cells to refine (flag2==1, flag1==1, son==0), mode 1: cells to de-refine (flag1==0, son>0), reported in the
 * reference's visiting order (batches of nvector octs, cell position outermost inside a batch) as pairs (igrid, ind).
 * Returns the number of pairs (at most cap are stored).                                                              */
int orc_amr_refine_scan(const orc_mesh* m, int l, int kind, int b, int nvector, const int* flag1, const int* flag2,
                        int mode, int* out, int cap) {
  int T = 1 << m->ndim;
  int n = kind == 0 ? m->nactive[l] : m->nbound[b][l];
  const int* lst = kind == 0 ? m->active[l] : m->bound[b][l];
  int cnt = 0;
  for (int i0 = 0; i0 < n; i0 += nvector) {
    int i1 = i0 + nvector < n ? i0 + nvector : n;
    for (int ind = 0; ind < T; ind++)
      for (int i = i0; i < i1; i++) {
        int c = amr_cell(m, ind, lst[i]);
        int hit = mode == 0 ? (flag2[c] == 1 && flag1[c] == 1 && m->son[c] == 0) : (flag1[c] == 0 && m->son[c] > 0);
        if (hit) {
          if (cnt < cap) { out[2 * cnt] = lst[i]; out[2 * cnt + 1] = ind; }
          cnt++;
        }
      }
  }
  return cnt;
}
