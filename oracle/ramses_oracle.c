/*
 * ramses_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the RAMSES (tatary/ramses) per-level hydro sweep in
 * the reference's own shape: per-oct 6^ndim patches gathered from the oct
 * tree, processed in batches of `nvector` octs, with the reference pass
 * structure courant_fine -> set_unew -> godunov_fine -> set_uold.  Every
 * floating-point expression keeps the Fortran evaluation order (left to
 * right, `**2` = x*x, `**3` = x*x*x, sign(one,x) = copysign, MAX/MIN keep the
 * first argument on ties); build with  -O3 -ffp-contract=off  (no FMA, no
 * fast-math) to mirror `gfortran -O3` on baseline x86-64 (bin/Makefile:99-106).
 *
 * PARITY PINNING STATUS ("how do we know this restatement is right?"):
 *   The reference is Fortran 90; neither gfortran nor MPI exists in the build
 *   image, so the reference itself cannot be compiled or run here
 *   (oracle/_ref is therefore absent).  What pins this oracle:
 *    1. PINNED (tolerance 3e-13, the reference's own): tests/golden/sod_tube_ref.json
 *       <- tests/hydro/sod-tube/sod-tube-ref.dat.  tests/test_oracle_golden.py runs
 *       the 1-D AMR Sod tube (levelmin 3, levelmax 10, hllc, moncen, sub-cycling,
 *       interpol_type 2) through oracle/amr.py, which calls THIS library for every
 *       floating-point routine, and reproduces all golden sums (density and pressure
 *       log-sums to 16 digits, velocity to 3e-16, time to 2e-15, ncells/level/x
 *       exactly) plus the mesh structure and step counts of doc/wiki/Start.md:158-187.
 *    1b. PINNED (tolerance 3e-13): tests/golden/implosion_ref.json
 *       <- tests/hydro/implosion/implosion-ref.dat.  The 2-D AMR implosion (levels 5-8,
 *       hllc, moncen, four reflexive walls with corner octs, nsubcycle 2, t=5: 1049
 *       coarse / 8392 fine steps, 15 349 leaf cells) through oracle/amr.py::FastAmrRun:
 *       ncells, level, dx, x, y exact; log-sums of density 4e-14, pressure 2.0e-13,
 *       sum|v_x| 2.2e-13, sum|v_y| 1.8e-13, time 4e-15 -- all inside the reference's
 *       tolerance; the golden file's own x<->y asymmetry is 1.5e-13, i.e. the residual is
 *       the round-off amplification of that flow (it moves by 1e-14 with nvector).
 *    2. tests/golden/sod_tube_ana.json  <- tests/hydro/sod-tube/sod-tube-ana.dat
 *       (exact Sod solution, 1024 points): discretisation-level check of every
 *       Riemann solver on a uniform grid.
 *    3. tests/golden/indices3cube.json <- amr/nbors_utils.f90:305-358 (the
 *       lll/mmm neighbour tables) against the generated tables below.
 *    4. invariants: conservation to round-off on periodic runs, x<->y<->z
 *       permutation symmetry, dt parity.
 *    5. NDIM=3 (what the GPU kernels are compared with): level steps of a 2-D run and of the
 *       z-invariant 3-D run are bit-identical for every solver (tests/test_oracle.py), so the
 *       NDIM=3 branches are tied to the golden-pinned NDIM=2 branches.
 *   NOT covered by a golden file (none exists in the reference): the bodies of
 *   riemann='exact'/'acoustic'/'hll'/'llf' and slope types other than 2 -- held by 2-5.
 *   PARITY UNPINNED for two options: poisson (gravity predictor, gloc gather, add_gravity_source_terms,
 *   gravity term of cmpdt; orc_set_gravity) and pressure_fix (divu / enew, add_pdv_source_terms, the
 *   energy switch; orc_set_pressure_fix).  Every reference test that enables them also needs cooling,
 *   sinks, RT or a patched condinit, which are outside this path, so no golden vector exists for them;
 *   they are restated from the reference text and held by the properties of
 *   tests/test_oracle.py::test_gravity_restatement / test_pressure_fix_restatement only.
 *
 * Citations are reference file:line.
 */
#include "ramses_oracle.h"
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_abi_version(void) { return 1; }

/* host threads for the per-level passes (the reference spreads every pass over its MPI ranks) */
static int g_nthreads = 1;
void orc_set_threads(int n) { g_nthreads = n < 1 ? 1 : n; }

/* module const (hydro/hydro_commons.f90:14-27) */
static const double zero = 0.0, one = 1.0, two = 2.0, half = 0.5;

/* Fortran MAX/MIN as gfortran emits them: keep first argument on ties */
static inline double FMAX(double a, double b) { return (b > a) ? b : a; }
static inline double FMIN(double a, double b) { return (b < a) ? b : a; }
static inline double FSIGN(double a, double b) { return copysign(a, b); } /* sign(a,b), F2003 signed zero */
static inline int IMIN(int a, int b) { return a < b ? a : b; }
static inline int IMAX(int a, int b) { return a > b ? a : b; }
static inline int ipow2(int n) { return 1 << n; }

/* ------------------------------------------------------------------------- */
/*                       Riemann solvers (godunov_utils.f90)                  */
/* ------------------------------------------------------------------------- */
#define QL(l, n) ql[(l) + (size_t)nv * ((n)-1)]
#define QR(l, n) qr[(l) + (size_t)nv * ((n)-1)]
#define FG(l, n) fg[(l) + (size_t)nv * ((n)-1)]

/* hydro/godunov_utils.f90:660-820 */
void orc_riemann_llf(const orc_params* p, const double* ql, const double* qr, double* fg, int ngrid) {
  const int nv = p->nvector, nvar = p->nvar, ndim = p->ndim;
  const double gamma = p->gamma, smallr = p->smallr, smallc = p->smallc;
  const double smallp = smallc * smallc / gamma;
  const double entho = one / (gamma - one);
  double ul_[16], ur_[16], fl_[16], fr_[16];
  for (int i = 0; i < ngrid; i++) {
    /* maximum wave speed :685-712 */
    double rl = FMAX(QL(i, 1), smallr);
    double ul = QL(i, 2);
    double pl = FMAX(QL(i, 3), rl * smallp);
    double cl = gamma * pl;
    cl = sqrt(cl / rl);
    double rr = FMAX(QR(i, 1), smallr);
    double ur = QR(i, 2);
    double pr = FMAX(QR(i, 3), rr * smallp);
    double cr = gamma * pr;
    cr = sqrt(cr / rr);
    double cmax = FMAX(fabs(ul) + cl, fabs(ur) + cr);
    /* conservative variables :717-768 */
    ul_[1] = QL(i, 1);
    ur_[1] = QR(i, 1);
    ul_[2] = QL(i, 1) * QL(i, 2);
    ur_[2] = QR(i, 1) * QR(i, 2);
    ul_[3] = QL(i, 3) * entho + half * QL(i, 1) * (QL(i, 2) * QL(i, 2));
    ur_[3] = QR(i, 3) * entho + half * QR(i, 1) * (QR(i, 2) * QR(i, 2));
    if (ndim > 1) {
      ul_[3] = ul_[3] + half * QL(i, 1) * (QL(i, 4) * QL(i, 4));
      ur_[3] = ur_[3] + half * QR(i, 1) * (QR(i, 4) * QR(i, 4));
    }
    if (ndim > 2) {
      ul_[3] = ul_[3] + half * QL(i, 1) * (QL(i, 5) * QL(i, 5));
      ur_[3] = ur_[3] + half * QR(i, 1) * (QR(i, 5) * QR(i, 5));
    }
    for (int n = 4; n <= ndim + 2; n++) {
      ul_[n] = QL(i, 1) * QL(i, n);
      ur_[n] = QR(i, 1) * QR(i, n);
    }
    for (int n = 3 + ndim; n <= nvar; n++) {
      ul_[n] = QL(i, 1) * QL(i, n);
      ur_[n] = QR(i, 1) * QR(i, n);
    }
    ul_[nvar + 1] = QL(i, 3) * entho;
    ur_[nvar + 1] = QR(i, 3) * entho;
    /* left and right fluxes :773-802 */
    fl_[1] = QL(i, 2) * ul_[1];
    fr_[1] = QR(i, 2) * ur_[1];
    fl_[2] = QL(i, 2) * ul_[2] + QL(i, 3);
    fr_[2] = QR(i, 2) * ur_[2] + QR(i, 3);
    fl_[3] = QL(i, 2) * (ul_[3] + QL(i, 3));
    fr_[3] = QR(i, 2) * (ur_[3] + QR(i, 3));
    for (int n = 4; n <= nvar + 1; n++) {
      fl_[n] = QL(i, 2) * ul_[n];
      fr_[n] = QR(i, 2) * ur_[n];
    }
    /* Lax-Friedrich :807-811 */
    for (int n = 1; n <= nvar + 1; n++)
      FG(i, n) = half * (fl_[n] + fr_[n] - cmax * (ur_[n] - ul_[n]));
  }
}

/* hydro/godunov_utils.f90:825-983 */
void orc_riemann_hll(const orc_params* p, const double* ql, const double* qr, double* fg, int ngrid) {
  const int nv = p->nvector, nvar = p->nvar, ndim = p->ndim;
  const double gamma = p->gamma, smallr = p->smallr, smallc = p->smallc;
  const double smallp = smallc * smallc / gamma;
  const double entho = one / (gamma - one);
  double ul_[16], ur_[16], fl_[16], fr_[16];
  for (int i = 0; i < ngrid; i++) {
    double rl = FMAX(QL(i, 1), smallr);
    double ul = QL(i, 2);
    double pl = FMAX(QL(i, 3), rl * smallp);
    double cl = gamma * pl;
    cl = sqrt(cl / rl);
    double rr = FMAX(QR(i, 1), smallr);
    double ur = QR(i, 2);
    double pr = FMAX(QR(i, 3), rr * smallp);
    double cr = gamma * pr;
    cr = sqrt(cr / rr);
    double SL = FMIN(FMIN(ul, ur) - FMAX(cl, cr), zero); /* :873 */
    double SR = FMAX(FMAX(ul, ur) + FMAX(cl, cr), zero); /* :874 */
    ul_[1] = QL(i, 1);
    ur_[1] = QR(i, 1);
    ul_[2] = QL(i, 1) * QL(i, 2);
    ur_[2] = QR(i, 1) * QR(i, 2);
    ul_[3] = QL(i, 3) * entho + half * QL(i, 1) * (QL(i, 2) * QL(i, 2));
    ur_[3] = QR(i, 3) * entho + half * QR(i, 1) * (QR(i, 2) * QR(i, 2));
    if (ndim > 1) {
      ul_[3] = ul_[3] + half * QL(i, 1) * (QL(i, 4) * QL(i, 4));
      ur_[3] = ur_[3] + half * QR(i, 1) * (QR(i, 4) * QR(i, 4));
    }
    if (ndim > 2) {
      ul_[3] = ul_[3] + half * QL(i, 1) * (QL(i, 5) * QL(i, 5));
      ur_[3] = ur_[3] + half * QR(i, 1) * (QR(i, 5) * QR(i, 5));
    }
    for (int n = 4; n <= ndim + 2; n++) {
      ul_[n] = QL(i, 1) * QL(i, n);
      ur_[n] = QR(i, 1) * QR(i, n);
    }
    for (int n = 3 + ndim; n <= nvar; n++) {
      ul_[n] = QL(i, 1) * QL(i, n);
      ur_[n] = QR(i, 1) * QR(i, n);
    }
    ul_[nvar + 1] = QL(i, 3) * entho;
    ur_[nvar + 1] = QR(i, 3) * entho;
    /* :939-966 */
    fl_[1] = ul_[2];
    fr_[1] = ur_[2];
    fl_[2] = QL(i, 3) + ul_[2] * QL(i, 2);
    fr_[2] = QR(i, 3) + ur_[2] * QR(i, 2);
    fl_[3] = QL(i, 2) * (ul_[3] + QL(i, 3));
    fr_[3] = QR(i, 2) * (ur_[3] + QR(i, 3));
    for (int n = 4; n <= nvar + 1; n++) {
      fl_[n] = QL(i, 2) * ul_[n];
      fr_[n] = QR(i, 2) * ur_[n];
    }
    /* :971-976 */
    for (int n = 1; n <= nvar + 1; n++)
      FG(i, n) = (SR * fl_[n] - SL * fr_[n] + SR * SL * (ur_[n] - ul_[n])) / (SR - SL);
  }
}

/* hydro/godunov_utils.f90:988-1209 */
void orc_riemann_hllc(const orc_params* p, const double* ql, const double* qr, double* fg, int ngrid) {
  const int nv = p->nvector, nvar = p->nvar, ndim = p->ndim;
  const double gamma = p->gamma, smallr = p->smallr, smallc = p->smallc;
  const double smallp = smallc * smallc / gamma;
  const double entho = one / (gamma - one);
  for (int i = 0; i < ngrid; i++) {
    /* left :1017-1042 */
    double rl = FMAX(QL(i, 1), smallr);
    double Pl = FMAX(QL(i, 3), rl * smallp);
    double ul = QL(i, 2);
    double el = Pl * entho;
    double ecinl = half * rl * ul * ul;
    if (ndim > 1) ecinl = ecinl + half * rl * (QL(i, 4) * QL(i, 4));
    if (ndim > 2) ecinl = ecinl + half * rl * (QL(i, 5) * QL(i, 5));
    double etotl = el + ecinl;
    double Ptotl = Pl;
    /* right :1044-1069 */
    double rr = FMAX(QR(i, 1), smallr);
    double Pr = FMAX(QR(i, 3), rr * smallp);
    double ur = QR(i, 2);
    double er = Pr * entho;
    double ecinr = half * rr * ur * ur;
    if (ndim > 1) ecinr = ecinr + half * rr * (QR(i, 4) * QR(i, 4));
    if (ndim > 2) ecinr = ecinr + half * rr * (QR(i, 5) * QR(i, 5));
    double etotr = er + ecinr;
    double Ptotr = Pr;
    /* fast speeds :1072-1086 */
    double cfastl = gamma * Pl;
    cfastl = sqrt(FMAX(cfastl / rl, smallc * smallc));
    double cfastr = gamma * Pr;
    cfastr = sqrt(FMAX(cfastr / rr, smallc * smallc));
    /* HLL wave speeds :1089-1090 */
    double SL = FMIN(ul, ur) - FMAX(cfastl, cfastr);
    double SR = FMAX(ul, ur) + FMAX(cfastl, cfastr);
    /* lagrangian sound speed :1093-1094 */
    double rcl = rl * (ul - SL);
    double rcr = rr * (SR - ur);
    /* acoustic star state :1097-1098 */
    double ustar = (rcr * ur + rcl * ul + (Ptotl - Ptotr)) / (rcr + rcl);
    double Ptotstar = (rcr * Ptotl + rcl * Ptotr + rcl * rcr * (ul - ur)) / (rcr + rcl);
    /* left star :1101-1103 */
    double rstarl = rl * (SL - ul) / (SL - ustar);
    double etotstarl = ((SL - ul) * etotl - Ptotl * ul + Ptotstar * ustar) / (SL - ustar);
    double estarl = el * (SL - ul) / (SL - ustar);
    /* right star :1111-1113 */
    double rstarr = rr * (SR - ur) / (SR - ustar);
    double etotstarr = ((SR - ur) * etotr - Ptotr * ur + Ptotstar * ustar) / (SR - ustar);
    double estarr = er * (SR - ur) / (SR - ustar);
    /* sample :1121-1170 */
    double ro, uo, Ptoto, etoto, eo;
    if (SL > 0.0) {
      ro = rl; uo = ul; Ptoto = Ptotl; etoto = etotl; eo = el;
    } else if (ustar > 0.0) {
      ro = rstarl; uo = ustar; Ptoto = Ptotstar; etoto = etotstarl; eo = estarl;
    } else if (SR > 0.0) {
      ro = rstarr; uo = ustar; Ptoto = Ptotstar; etoto = etotstarr; eo = estarr;
    } else {
      ro = rr; uo = ur; Ptoto = Ptotr; etoto = etotr; eo = er;
    }
    /* Godunov flux :1175-1205 */
    FG(i, 1) = ro * uo;
    FG(i, 2) = ro * uo * uo + Ptoto;
    FG(i, 3) = (etoto + Ptoto) * uo;
    for (int ivar = 4; ivar <= ndim + 2; ivar++) {
      if (ustar > 0) FG(i, ivar) = ro * uo * QL(i, ivar);
      else           FG(i, ivar) = ro * uo * QR(i, ivar);
    }
    for (int ivar = 3 + ndim; ivar <= nvar; ivar++) {
      if (ustar > 0) FG(i, ivar) = ro * uo * QL(i, ivar);
      else           FG(i, ivar) = ro * uo * QR(i, ivar);
    }
    FG(i, nvar + 1) = uo * eo;
  }
}

/* common tail of riemann_approx / riemann_acoustic: fluxes from qgdnv
 * hydro/godunov_utils.f90:474-493 and :634-652 (identical text)               */
static void flux_from_qgdnv(const orc_params* p, const double* qg /*[nvar+2], 1-based*/, double* fg, int nv, int i) {
  const int nvar = p->nvar, ndim = p->ndim;
  const double entho = one / (p->gamma - one);
  FG(i, 1) = qg[1] * qg[2];
  FG(i, 2) = qg[3] + qg[1] * (qg[2] * qg[2]);
  double etot = qg[3] * entho + half * qg[1] * (qg[2] * qg[2]);
  if (ndim > 1) etot = etot + half * qg[1] * (qg[4] * qg[4]);
  if (ndim > 2) etot = etot + half * qg[1] * (qg[5] * qg[5]);
  FG(i, 3) = qg[2] * (etot + qg[3]);
  for (int n = 4; n <= nvar + 1; n++) FG(i, n) = FG(i, 1) * qg[n];
}

/* hydro/godunov_utils.f90:500-655 */
void orc_riemann_acoustic(const orc_params* p, const double* ql, const double* qr, double* fg, int ngrid) {
  const int nv = p->nvector, nvar = p->nvar;
  const double gamma = p->gamma, smallr = p->smallr, smallc = p->smallc;
  const double smallp = smallc * smallc / gamma;
  const double entho = one / (gamma - one);
  double qg[18];
  for (int i = 0; i < ngrid; i++) {
    double rl = FMAX(QL(i, 1), smallr);
    double ul = QL(i, 2);
    double pl = FMAX(QL(i, 3), rl * smallp);
    double rr = FMAX(QR(i, 1), smallr);
    double ur = QR(i, 2);
    double pr = FMAX(QR(i, 3), rr * smallp);
    /* acoustic star state :541-551 */
    double cl = sqrt(gamma * pl / rl);
    double cr = sqrt(gamma * pr / rr);
    double wl = cl * rl;
    double wr = cr * rr;
    double pstar = ((wr * pl + wl * pr) + wl * wr * (ul - ur)) / (wl + wr);
    double ustar = ((wr * ur + wl * ul) + (pl - pr)) / (wl + wr);
    double sgnm = FSIGN(one, ustar);
    double ro, uo, po, wo, co;
    if (sgnm == one) { ro = rl; uo = ul; po = pl; wo = wl; co = cl; }
    else             { ro = rr; uo = ur; po = pr; wo = wr; co = cr; }
    (void)wo;
    /* star region density and sound speed :576-581 */
    double rstar = ro + (pstar - po) / (co * co);
    rstar = FMAX(rstar, smallr);
    double cstar = sqrt(fabs(gamma * pstar / rstar));
    cstar = FMAX(cstar, smallc);
    /* head and tail speeds :584-587 */
    double spout = co - sgnm * uo;
    double spin = cstar - sgnm * ustar;
    /* shock speed :590-599 */
    double ushock = half * (spin + spout);
    ushock = FMAX(ushock, -sgnm * ustar);
    if (pstar >= po) { spout = ushock; spin = spout; }
    /* sample :602-617 */
    if (spout < zero) {
      qg[1] = ro; qg[2] = uo; qg[3] = po;
    } else if (spin >= zero) {
      qg[1] = rstar; qg[2] = ustar; qg[3] = pstar;
    } else {
      double frac = spout / (spout - spin);
      qg[1] = frac * rstar + (one - frac) * ro;
      qg[2] = frac * ustar + (one - frac) * uo;
      qg[3] = frac * pstar + (one - frac) * po;
    }
    for (int n = 4; n <= nvar; n++) qg[n] = (sgnm == one) ? QL(i, n) : QR(i, n);
    qg[nvar + 1] = po / ro * entho;
    flux_from_qgdnv(p, qg, fg, nv, i);
  }
}

/* hydro/godunov_utils.f90:268-495 ('exact' = two-shock Newton iteration).
 * The reference compacts unconverged lanes (:330-366); per lane this is
 * "iterate until |delp/(pold+smallpp)| <= 1e-6, at most niter_riemann times". */
void orc_riemann_approx(const orc_params* p, const double* ql, const double* qr, double* fg, int ngrid) {
  const int nv = p->nvector, nvar = p->nvar;
  const double gamma = p->gamma, smallr = p->smallr, smallc = p->smallc;
  const double smallp = smallc * smallc / gamma;
  const double smallpp = smallr * smallp;
  const double gamma6 = (gamma + one) / (two * gamma);
  const double entho = one / (gamma - one);
  double qg[18];
  for (int i = 0; i < ngrid; i++) {
    double rl = FMAX(QL(i, 1), smallr);
    double ul = QL(i, 2);
    double pl = FMAX(QL(i, 3), rl * smallp);
    double rr = FMAX(QR(i, 1), smallr);
    double ur = QR(i, 2);
    double pr = FMAX(QR(i, 3), rr * smallp);
    /* lagrangian sound speed :306-309 */
    double cl = gamma * pl * rl;
    double cr = gamma * pr * rr;
    /* first guess :312-317 */
    double wl = sqrt(cl), wr = sqrt(cr);
    double pstar = ((wr * pl + wl * pr) + wl * wr * (ul - ur)) / (wl + wr);
    pstar = FMAX(pstar, 0.0);
    double pold = pstar;
    /* Newton-Raphson :325-366 */
    for (int iter = 1; iter <= p->niter_riemann; iter++) {
      double wwl = sqrt(cl * (one + gamma6 * (pold - pl) / pl));
      double wwr = sqrt(cr * (one + gamma6 * (pold - pr) / pr));
      double qql = two * (wwl * wwl * wwl) / (wwl * wwl + cl);
      double qqr = two * (wwr * wwr * wwr) / (wwr * wwr + cr);
      double usl = ul - (pold - pl) / wwl;
      double usr = ur + (pold - pr) / wwr;
      double delp = FMAX(qqr * qql / (qqr + qql) * (usl - usr), -pold);
      pold = pold + delp;
      double uo_ = fabs(delp / (pold + smallpp));
      if (!(uo_ > 1e-06)) break;
    }
    /* star region :370-382 */
    pstar = pold;
    wl = sqrt(cl * (one + gamma6 * (pstar - pl) / pl));
    wr = sqrt(cr * (one + gamma6 * (pstar - pr) / pr));
    double ustar = half * (ul + (pl - pstar) / wl + ur - (pr - pstar) / wr);
    double sgnm = FSIGN(one, ustar);
    double ro, uo, po, wo;
    if (sgnm == one) { ro = rl; uo = ul; po = pl; wo = wl; }
    else             { ro = rr; uo = ur; po = pr; wo = wr; }
    double co = FMAX(smallc, sqrt(fabs(gamma * po / ro)));
    /* star region density :409-417 */
    double rstar;
    if (pstar >= po) rstar = ro / (one + ro * (po - pstar) / (wo * wo));
    else             rstar = ro * pow(pstar / po, one / gamma);
    rstar = FMAX(rstar, smallr);
    double cstar = sqrt(fabs(gamma * pstar / rstar));
    cstar = FMAX(cstar, smallc);
    double spout = co - sgnm * uo;
    double spin = cstar - sgnm * ustar;
    double ushock = wo / ro - sgnm * uo;
    if (pstar >= po) { spout = ushock; spin = spout; }
    /* sample :440-456 */
    if (spout <= zero) {
      qg[1] = ro; qg[2] = uo; qg[3] = po;
    } else if (spin >= zero) {
      qg[1] = rstar; qg[2] = ustar; qg[3] = pstar;
    } else {
      double frac = spout / (spout - spin);
      qg[2] = frac * ustar + (one - frac) * uo;
      qg[3] = frac * pstar + (one - frac) * po;
      qg[1] = ro * pow(qg[3] / po, one / gamma);
    }
    for (int n = 4; n <= nvar; n++) qg[n] = (sgnm == one) ? QL(i, n) : QR(i, n);
    qg[nvar + 1] = po / ro * entho;
    flux_from_qgdnv(p, qg, fg, nv, i);
  }
}
#undef QL
#undef QR
#undef FG

/* ------------------------------------------------------------------------- */
/*                     unsplit and friends (hydro/umuscl.f90)                 */
/* ------------------------------------------------------------------------- */
struct orc_work {
  int nv, ndim, nvar;
  int nj, nk, j0, k0;    /* patch extents: i -1..4 ; j j0..j0+nj-1 ; k ...    */
  int nfj, nfk;          /* flux extents 1..3 / 1..nfj / 1..nfk               */
  size_t np;             /* patch scalar-field size = nv*6*nj*nk              */
  size_t nfp;            /* flux scalar-field size  = nv*3*nfj*nfk            */
  double *qin, *cin, *dq, *qm, *qp, *fx, *tx;
  double *qleft, *qright, *fgdnv;
  /* godfine1 buffers */
  double *uloc, *gloc, *flux, *tmp;
  unsigned char* ok;
};

orc_work* orc_work_new(const orc_params* p) {
  orc_work* w = (orc_work*)calloc(1, sizeof(orc_work));
  w->nv = p->nvector; w->ndim = p->ndim; w->nvar = p->nvar;
  w->nj = p->ndim > 1 ? 6 : 1; w->nk = p->ndim > 2 ? 6 : 1;
  w->j0 = p->ndim > 1 ? -1 : 1; w->k0 = p->ndim > 2 ? -1 : 1; /* hydro_parameters.f90:20-25 */
  w->nfj = p->ndim > 1 ? 3 : 1; w->nfk = p->ndim > 2 ? 3 : 1; /* :26-31 */
  w->np = (size_t)w->nv * 6 * w->nj * w->nk;
  w->nfp = (size_t)w->nv * 3 * w->nfj * w->nfk;
  size_t np = w->np, nvar = p->nvar, ndim = p->ndim;
  w->qin = (double*)calloc(np * nvar, 8);
  w->cin = (double*)calloc(np, 8);
  w->dq = (double*)calloc(np * nvar * ndim, 8);
  w->qm = (double*)calloc(np * nvar * ndim, 8);
  w->qp = (double*)calloc(np * nvar * ndim, 8);
  w->fx = (double*)calloc(np * nvar, 8);
  w->tx = (double*)calloc(np * 2, 8);
  w->qleft = (double*)calloc((size_t)w->nv * nvar, 8);
  w->qright = (double*)calloc((size_t)w->nv * nvar, 8);
  w->fgdnv = (double*)calloc((size_t)w->nv * (nvar + 1), 8);
  w->uloc = (double*)calloc(np * nvar, 8);
  w->gloc = (double*)calloc(np * ndim, 8);
  w->flux = (double*)calloc(w->nfp * nvar * ndim, 8);
  w->tmp = (double*)calloc(w->nfp * 2 * ndim, 8);
  w->ok = (unsigned char*)calloc(np, 1);
  return w;
}
void orc_work_free(orc_work* w) {
  if (!w) return;
  free(w->qin); free(w->cin); free(w->dq); free(w->qm); free(w->qp); free(w->fx); free(w->tx);
  free(w->qleft); free(w->qright); free(w->fgdnv); free(w->uloc); free(w->gloc); free(w->flux);
  free(w->tmp); free(w->ok); free(w);
}

/* patch index of (l,i,j,k): Fortran (1:nvector,iu1:iu2,ju1:ju2,ku1:ku2) */
#define PIX(w, l, i, j, k) ((size_t)(l) + (size_t)(w)->nv * (((i) + 1) + 6 * (((j) - (w)->j0) + (size_t)(w)->nj * ((k) - (w)->k0))))
/* flux index (1:nvector,if1:if2,jf1:jf2,kf1:kf2) */
#define FIX(w, l, i, j, k) ((size_t)(l) + (size_t)(w)->nv * (((i)-1) + 3 * (((j)-1) + (size_t)(w)->nfj * ((k)-1))))

/* ctoprim hydro/umuscl.f90:861-965 */
static void ctoprim(const orc_params* p, orc_work* w, const double* uin, const double* gravin, double dt, int ngrid) {
  const int ndim = p->ndim, nvar = p->nvar;
  const size_t np = w->np;
  const double gamma = p->gamma, smallr = p->smallr, smallc = p->smallc;
  const double smalle = smallc * smallc / gamma / (gamma - one); /* :884 */
  const double dtxhalf = dt * half;
  double* q = w->qin;
  double* c = w->cin;
  for (int k = w->k0; k < w->k0 + w->nk; k++)
    for (int j = w->j0; j < w->j0 + w->nj; j++)
      for (int i = -1; i <= 4; i++) {
        size_t b = PIX(w, 0, i, j, k);
        for (int l = 0; l < ngrid; l++) {
          size_t x = b + l;
          q[x] = FMAX(uin[x], smallr);
          double oneoverrho = one / q[x];
          q[np + x] = uin[np + x] * oneoverrho;
          if (ndim > 1) q[2 * np + x] = uin[2 * np + x] * oneoverrho;
          if (ndim > 2) q[3 * np + x] = uin[3 * np + x] * oneoverrho;
          double eken = half * q[np + x] * q[np + x];
          if (ndim > 1) eken = eken + half * q[2 * np + x] * q[2 * np + x];
          if (ndim > 2) eken = eken + half * q[3 * np + x] * q[3 * np + x];
          double erad = zero;
          double eint = FMAX(uin[(ndim + 1) * np + x] * oneoverrho - eken - erad, smalle);
          q[(ndim + 1) * np + x] = (gamma - one) * q[x] * eint;
          c[x] = gamma * q[(ndim + 1) * np + x];
          c[x] = sqrt(c[x] * oneoverrho);
          /* gravity predictor :932-938 */
          if (gravin) {
            q[np + x] = q[np + x] + gravin[x] * dtxhalf;
            if (ndim > 1) q[2 * np + x] = q[2 * np + x] + gravin[np + x] * dtxhalf;
            if (ndim > 2) q[3 * np + x] = q[3 * np + x] + gravin[2 * np + x] * dtxhalf;
          } else {
            q[np + x] = q[np + x] + 0.0 * dtxhalf;
            if (ndim > 1) q[2 * np + x] = q[2 * np + x] + 0.0 * dtxhalf;
            if (ndim > 2) q[3 * np + x] = q[3 * np + x] + 0.0 * dtxhalf;
          }
        }
      }
  /* passive scalars :948-961 */
  for (int n = ndim + 3; n <= nvar; n++)
    for (int k = w->k0; k < w->k0 + w->nk; k++)
      for (int j = w->j0; j < w->j0 + w->nj; j++)
        for (int i = -1; i <= 4; i++) {
          size_t b = PIX(w, 0, i, j, k);
          for (int l = 0; l < ngrid; l++) {
            double oneoverrho = one / q[b + l];
            q[(n - 1) * np + b + l] = uin[(n - 1) * np + b + l] * oneoverrho;
          }
        }
}

/* one limited slope from left/centre/right values; formulas of uslope
 * hydro/umuscl.f90:970-1480, which differ in form between NDIM builds          */
static inline double slope_lcr(const orc_params* p, double qlft, double qcen, double qrgt) {
  const int st = p->slope_type, ndim = p->ndim;
  double dlft, drgt, dcen, dsgn, slop, dlim;
  if ((ndim == 1 && (st == 1 || st == 2 || st == 3)) || (ndim == 2 && (st == 1 || st == 2)) || (ndim == 3 && st == 2)) {
    /* :1009-1019 (1-D uses MIN(slope_type,2)), :1079-1099, :1286-1326 */
    double f = (double)IMIN(st, 2);
    dlft = f * (qcen - qlft);
    drgt = f * (qrgt - qcen);
    dcen = half * (dlft + drgt) / f;
    dsgn = FSIGN(one, dcen);
    slop = FMIN(fabs(dlft), fabs(drgt));
    dlim = slop;
    if ((dlft * drgt) <= zero) dlim = zero;
    return dsgn * FMIN(dlim, fabs(dcen));
  }
  if (ndim == 3 && st == 1) { /* :1241-1284 */
    dlft = qcen - qlft;
    drgt = qrgt - qcen;
    if ((dlft * drgt) <= zero) return zero;
    else if (dlft > 0) return FMIN(dlft, drgt);
    else return FMAX(dlft, drgt);
  }
  if (st == 7) { /* van Leer :1060-1069, :1146-1171, :1393-1431 */
    dlft = (qcen - qlft);
    drgt = (qrgt - qcen);
    if ((dlft * drgt) <= zero) return zero;
    return (2 * dlft * drgt / (dlft + drgt));
  }
  if (st == 8) { /* generalised moncen :1070-1081, :1172-1201, :1432-1473 */
    dlft = (qcen - qlft);
    drgt = (qrgt - qcen);
    dcen = half * (dlft + drgt);
    dsgn = FSIGN(one, dcen);
    slop = FMIN(p->slope_theta * fabs(dlft), p->slope_theta * fabs(drgt));
    dlim = slop;
    if ((dlft * drgt) <= zero) dlim = zero;
    return dsgn * FMIN(dlim, fabs(dcen));
  }
  fprintf(stderr, "orc: Unknown slope type %d for ndim %d\n", st, ndim);
  abort();
}

/* uslope hydro/umuscl.f90:970 */
static void uslope(const orc_params* p, orc_work* w, double dx, double dt, int ngrid) {
  const int ndim = p->ndim, nvar = p->nvar, st = p->slope_type;
  const size_t np = w->np;
  const double* q = w->qin;
  double* dq = w->dq;
  const int ilo = 0, ihi = 3; /* MIN(1,iu1+1)=0, MAX(1,iu2-1)=3 :998-1000 */
  const int jlo = ndim > 1 ? 0 : 1, jhi = ndim > 1 ? 3 : 1;
  const int klo = ndim > 2 ? 0 : 1, khi = ndim > 2 ? 3 : 1;
  if (st == 0) { /* :1002-1005 */
    memset(dq, 0, sizeof(double) * np * nvar * ndim);
    return;
  }
  const size_t si = w->nv, sj = (size_t)w->nv * 6, sk = (size_t)w->nv * 6 * w->nj;
  for (int n = 0; n < nvar; n++)
    for (int k = klo; k <= khi; k++)
      for (int j = jlo; j <= jhi; j++)
        for (int i = ilo; i <= ihi; i++) {
          size_t b = PIX(w, 0, i, j, k);
          const double* qn = q + n * np;
          for (int l = 0; l < ngrid; l++) {
            size_t x = b + l;
            if (ndim == 1 && (st == 4 || st == 5 || st == 6)) {
              double dlft, drgt, dcen, dsgn, slop, dlim, r;
              if (st == 4) { /* superbee :1021-1031 */
                dcen = q[np + x] * dt / dx;
                dlft = two / (one + dcen) * (qn[x] - qn[x - si]);
                drgt = two / (one - dcen) * (qn[x + si] - qn[x]);
                dcen = half * (qn[x + si] - qn[x - si]);
                dsgn = FSIGN(one, dlft);
                slop = FMIN(fabs(dlft), fabs(drgt));
                dlim = slop;
                if ((dlft * drgt) <= zero) dlim = zero;
                r = dsgn * dlim;
              } else if (st == 5) { /* ultrabee :1032-1054 */
                if (n == 0) {
                  dcen = q[np + x] * dt / dx;
                  if (dcen >= 0) {
                    dlft = two / (zero + dcen + 1e-10) * (qn[x] - qn[x - si]);
                    drgt = two / (one - dcen) * (qn[x + si] - qn[x]);
                  } else {
                    dlft = two / (one + dcen) * (qn[x] - qn[x - si]);
                    drgt = two / (zero - dcen + 1e-10) * (qn[x + si] - qn[x]);
                  }
                  dsgn = FSIGN(one, dlft);
                  slop = FMIN(fabs(dlft), fabs(drgt));
                  dlim = slop;
                  if ((dlft * drgt) <= zero) dlim = zero;
                  r = dsgn * dlim;
                } else r = 0;
              } else { /* unstable :1055-1068 */
                if (n == 0) {
                  dlft = (qn[x] - qn[x - si]);
                  drgt = (qn[x + si] - qn[x]);
                  slop = 0.5 * (dlft + drgt);
                  r = slop;
                } else r = 0;
              }
              dq[(n + 0 * nvar) * np + x] = r;
              continue;
            }
            if (st == 3 && ndim == 2) { /* positivity preserving 2d :1101-1144 */
              double c0 = qn[x];
              double vmin = 0, vmax = 0;
              int first = 1;
              for (int a = -1; a <= 1; a++)       /* i offset outer: dfll,dflm,dflr,dfml,... */
                for (int bb = -1; bb <= 1; bb++) { /* j offset inner */
                  double d = qn[x + a * (ptrdiff_t)si + bb * (ptrdiff_t)sj] - c0;
                  if (first) { vmin = d; vmax = d; first = 0; }
                  else { vmin = FMIN(vmin, d); vmax = FMAX(vmax, d); }
                }
              double dfx = half * (qn[x + si] - qn[x - si]);
              double dfy = half * (qn[x + sj] - qn[x - sj]);
              double dff = half * (fabs(dfx) + fabs(dfy));
              double slop;
              if (dff > zero) slop = FMIN(one, FMIN(fabs(vmin), fabs(vmax)) / dff);
              else slop = one;
              dq[(n + 0 * nvar) * np + x] = slop * dfx;
              dq[(n + 1 * nvar) * np + x] = slop * dfy;
              continue;
            }
            if (st == 3 && ndim == 3) { /* positivity preserving 3d :1328-1391 */
              double c0 = qn[x];
              double vmin = 0, vmax = 0;
              int first = 1;
              /* argument order of the min/max lists: k outer (l,m,r), then i, then j */
              for (int cc = -1; cc <= 1; cc++)
                for (int a = -1; a <= 1; a++)
                  for (int bb = -1; bb <= 1; bb++) {
                    double d = qn[x + a * (ptrdiff_t)si + bb * (ptrdiff_t)sj + cc * (ptrdiff_t)sk] - c0;
                    if (first) { vmin = d; vmax = d; first = 0; }
                    else { vmin = FMIN(vmin, d); vmax = FMAX(vmax, d); }
                  }
              double dfx = half * (qn[x + si] - qn[x - si]);
              double dfy = half * (qn[x + sj] - qn[x - sj]);
              double dfz = half * (qn[x + sk] - qn[x - sk]);
              double dff = half * (fabs(dfx) + fabs(dfy) + fabs(dfz));
              double slop;
              if (dff > zero) slop = FMIN(one, FMIN(fabs(vmin), fabs(vmax)) / dff);
              else slop = one;
              dq[(n + 0 * nvar) * np + x] = slop * dfx;
              dq[(n + 1 * nvar) * np + x] = slop * dfy;
              dq[(n + 2 * nvar) * np + x] = slop * dfz;
              continue;
            }
            dq[(n + 0 * nvar) * np + x] = slope_lcr(p, qn[x - si], qn[x], qn[x + si]);
            if (ndim > 1) dq[(n + 1 * nvar) * np + x] = slope_lcr(p, qn[x - sj], qn[x], qn[x + sj]);
            if (ndim > 2) dq[(n + 2 * nvar) * np + x] = slope_lcr(p, qn[x - sk], qn[x], qn[x + sk]);
          }
        }
}

/* trace1d/2d/3d hydro/umuscl.f90:176,305,483 (one body: the three Fortran
 * routines differ only by the terms present; evaluation order is the same)     */
static void trace(const orc_params* p, orc_work* w, double dx, double dy, double dz, double dt, int ngrid) {
  const int ndim = p->ndim, nvar = p->nvar;
  const size_t np = w->np;
  const double gamma = p->gamma, smallr = p->smallr;
  const double* q = w->qin;
  const double* dq = w->dq;
  double *qm = w->qm, *qp = w->qp;
  const double dtd[3] = {dt / dx, dt / dy, dt / dz};
  const int ilo = 0, ihi = 3;
  const int jlo = ndim > 1 ? 0 : 1, jhi = ndim > 1 ? 3 : 1;
  const int klo = ndim > 2 ? 0 : 1, khi = ndim > 2 ? 3 : 1;
  const int ir = 0, iu = 1, iv = 2, iw = 3, ip = ndim + 1;
#define DQ(n, d) dq[((n) + (size_t)(d)*nvar) * np + x]
#define QM(n, d) qm[((n) + (size_t)(d)*nvar) * np + x]
#define QP(n, d) qp[((n) + (size_t)(d)*nvar) * np + x]
  for (int k = klo; k <= khi; k++)
    for (int j = jlo; j <= jhi; j++)
      for (int i = ilo; i <= ihi; i++) {
        size_t b = PIX(w, 0, i, j, k);
        for (int l = 0; l < ngrid; l++) {
          size_t x = b + l;
          double r = q[ir * np + x], u = q[iu * np + x], pp = q[ip * np + x];
          double v = ndim > 1 ? q[iv * np + x] : 0, ww = ndim > 2 ? q[iw * np + x] : 0;
          double drx = DQ(ir, 0), dpx = DQ(ip, 0), dux = DQ(iu, 0);
          double dvx = ndim > 1 ? DQ(iv, 0) : 0, dwx = ndim > 2 ? DQ(iw, 0) : 0;
          double dry = 0, dpy = 0, duy = 0, dvy = 0, dwy = 0, drz = 0, dpz = 0, duz = 0, dvz = 0, dwz = 0;
          if (ndim > 1) { dry = DQ(ir, 1); dpy = DQ(ip, 1); duy = DQ(iu, 1); dvy = DQ(iv, 1); if (ndim > 2) dwy = DQ(iw, 1); }
          if (ndim > 2) { drz = DQ(ir, 2); dpz = DQ(ip, 2); duz = DQ(iu, 2); dvz = DQ(iv, 2); dwz = DQ(iw, 2); }
          double sr0, sp0, su0, sv0 = 0, sw0 = 0;
          if (ndim == 1) { /* :236-238 */
            sr0 = -u * drx - (dux)*r;
            sp0 = -u * dpx - (dux)*gamma * pp;
            su0 = -u * dux - (dpx) / r;
          } else if (ndim == 2) { /* :383-386 */
            sr0 = -u * drx - v * dry - (dux + dvy) * r;
            sp0 = -u * dpx - v * dpy - (dux + dvy) * gamma * pp;
            su0 = -u * dux - v * duy - (dpx) / r;
            sv0 = -u * dvx - v * dvy - (dpy) / r;
          } else { /* :576-580 */
            sr0 = -u * drx - v * dry - ww * drz - (dux + dvy + dwz) * r;
            sp0 = -u * dpx - v * dpy - ww * dpz - (dux + dvy + dwz) * gamma * pp;
            su0 = -u * dux - v * duy - ww * duz - (dpx) / r;
            sv0 = -u * dvx - v * dvy - ww * dvz - (dpy) / r;
            sw0 = -u * dwx - v * dwy - ww * dwz - (dpz) / r;
          }
          const double dr_[3] = {drx, dry, drz}, dp_[3] = {dpx, dpy, dpz}, du_[3] = {dux, duy, duz};
          const double dv_[3] = {dvx, dvy, dvz}, dw_[3] = {dwx, dwy, dwz};
          for (int d = 0; d < ndim; d++) {
            double dtdx = dtd[d];
            /* right state at left interface :592-600 */
            QP(ir, d) = r - half * dr_[d] + sr0 * dtdx * half;
            QP(ip, d) = pp - half * dp_[d] + sp0 * dtdx * half;
            QP(iu, d) = u - half * du_[d] + su0 * dtdx * half;
            if (ndim > 1) QP(iv, d) = v - half * dv_[d] + sv0 * dtdx * half;
            if (ndim > 2) QP(iw, d) = ww - half * dw_[d] + sw0 * dtdx * half;
            if (QP(ir, d) < smallr) QP(ir, d) = r;
            /* left state at right interface :608-616 */
            QM(ir, d) = r + half * dr_[d] + sr0 * dtdx * half;
            QM(ip, d) = pp + half * dp_[d] + sp0 * dtdx * half;
            QM(iu, d) = u + half * du_[d] + su0 * dtdx * half;
            if (ndim > 1) QM(iv, d) = v + half * dv_[d] + sv0 * dtdx * half;
            if (ndim > 2) QM(iw, d) = ww + half * dw_[d] + sw0 * dtdx * half;
            if (QM(ir, d) < smallr) QM(ir, d) = r;
          }
          /* passive scalars :681-704 */
          for (int n = ndim + 2; n < nvar; n++) {
            double a = q[n * np + x];
            double dax = DQ(n, 0), day = ndim > 1 ? DQ(n, 1) : 0, daz = ndim > 2 ? DQ(n, 2) : 0;
            double sa0;
            if (ndim == 1) sa0 = -u * dax;
            else if (ndim == 2) sa0 = -u * dax - v * day;
            else sa0 = -u * dax - v * day - ww * daz;
            const double da_[3] = {dax, day, daz};
            for (int d = 0; d < ndim; d++) {
              QP(n, d) = a - half * da_[d] + sa0 * dtd[d] * half;
              QM(n, d) = a + half * da_[d] + sa0 * dtd[d] * half;
            }
          }
        }
      }
#undef DQ
#undef QM
#undef QP
}

/* cmpflxm hydro/umuscl.f90:714-856.  (si,sj,sk) is the index shift applied to
 * qm by passing it with bounds iu1+1:iu2+1 etc. (:97,:120,:144): face (i,j,k)
 * takes qleft = qm(i-si,j-sj,k-sk), qright = qp(i,j,k).                       */
static void cmpflxm(const orc_params* p, orc_work* w, int si, int sj, int sk, int ilo, int ihi, int jlo, int jhi,
                    int klo, int khi, int ln, int lt1, int lt2, int ngrid) {
  const int ndim = p->ndim, nvar = p->nvar, nv = w->nv;
  const size_t np = w->np;
  const int xdim = ln - 1; /* 1-based dimension */
  double *qleft = w->qleft, *qright = w->qright, *fgdnv = w->fgdnv;
  const double *qm = w->qm, *qp = w->qp;
  double *flx = w->fx, *tmp = w->tx;
#define QMv(n, x) qm[(((n)-1) + (size_t)(xdim - 1) * nvar) * np + (x)]
#define QPv(n, x) qp[(((n)-1) + (size_t)(xdim - 1) * nvar) * np + (x)]
  for (int k = klo; k <= khi; k++)
    for (int j = jlo; j <= jhi; j++)
      for (int i = ilo; i <= ihi; i++) {
        size_t xl = PIX(w, 0, i - si, j - sj, k - sk);
        size_t xr = PIX(w, 0, i, j, k);
        for (int l = 0; l < ngrid; l++) {
          qleft[l + nv * 0] = QMv(1, xl + l);  qright[l + nv * 0] = QPv(1, xr + l);
          qleft[l + nv * 1] = QMv(ln, xl + l); qright[l + nv * 1] = QPv(ln, xr + l);
          qleft[l + nv * 2] = QMv(ndim + 2, xl + l); qright[l + nv * 2] = QPv(ndim + 2, xr + l);
          if (ndim > 1) { qleft[l + nv * 3] = QMv(lt1, xl + l); qright[l + nv * 3] = QPv(lt1, xr + l); }
          if (ndim > 2) { qleft[l + nv * 4] = QMv(lt2, xl + l); qright[l + nv * 4] = QPv(lt2, xr + l); }
          for (int n = ndim + 3; n <= nvar; n++) { qleft[l + nv * (n - 1)] = QMv(n, xl + l); qright[l + nv * (n - 1)] = QPv(n, xr + l); }
        }
        switch (p->riemann) { /* :791-804 */
          case ORC_RIEMANN_ACOUSTIC: orc_riemann_acoustic(p, qleft, qright, fgdnv, ngrid); break;
          case ORC_RIEMANN_EXACT:    orc_riemann_approx(p, qleft, qright, fgdnv, ngrid); break;
          case ORC_RIEMANN_LLF:      orc_riemann_llf(p, qleft, qright, fgdnv, ngrid); break;
          case ORC_RIEMANN_HLLC:     orc_riemann_hllc(p, qleft, qright, fgdnv, ngrid); break;
          case ORC_RIEMANN_HLL:      orc_riemann_hll(p, qleft, qright, fgdnv, ngrid); break;
          default: fprintf(stderr, "unknown Riemann solver\n"); abort();
        }
        for (int l = 0; l < ngrid; l++) { /* :808-850 */
          flx[(1 - 1) * np + xr + l] = fgdnv[l + nv * 0];
          flx[(ln - 1) * np + xr + l] = fgdnv[l + nv * 1];
          if (ndim > 1) flx[(lt1 - 1) * np + xr + l] = fgdnv[l + nv * 3];
          if (ndim > 2) flx[(lt2 - 1) * np + xr + l] = fgdnv[l + nv * 4];
          flx[(ndim + 2 - 1) * np + xr + l] = fgdnv[l + nv * 2];
          for (int n = ndim + 3; n <= nvar; n++) flx[(n - 1) * np + xr + l] = fgdnv[l + nv * (n - 1)];
          tmp[0 * np + xr + l] = half * (qleft[l + nv * 1] + qright[l + nv * 1]);
          tmp[1 * np + xr + l] = fgdnv[l + nv * nvar];
        }
      }
#undef QMv
#undef QPv
}

/* cmpdivu hydro/uplmde.f90:702-764 and consup :769-869 (difmag>0 only) */
static void cmpdivu_consup(const orc_params* p, orc_work* w, const double* uin, double* flux, double dx, double dy,
                           double dz, double dt, int ngrid);

/* unsplit hydro/umuscl.f90:22-171 */
void orc_unsplit(const orc_params* p, orc_work* w, const double* uin, const double* gravin, double* flux, double* tmp,
                 double dx, double dy, double dz, double dt, int ngrid) {
  const int ndim = p->ndim, nvar = p->nvar;
  const size_t np = w->np, nfp = w->nfp;
  /* ilo=MIN(1,iu1+2)=1; ihi=MAX(1,iu2-2)=2 (:62-64) */
  const int ilo = 1, ihi = 2;
  const int jlo = 1, jhi = ndim > 1 ? 2 : 1;
  const int klo = 1, khi = ndim > 2 ? 2 : 1;
  const int if1 = 1, if2 = 3, jf1 = 1, jf2 = ndim > 1 ? 3 : 1, kf1 = 1, kf2 = ndim > 2 ? 3 : 1;
  if (p->scheme != ORC_SCHEME_MUSCL) { fprintf(stderr, "orc: scheme plmde not restated\n"); abort(); }
  ctoprim(p, w, uin, gravin, dt, ngrid);
  uslope(p, w, dx, dt, ngrid);
  trace(p, w, dx, dy, dz, dt, ngrid);
  /* X :97-116 */
  cmpflxm(p, w, 1, 0, 0, if1, if2, jlo, jhi, klo, khi, 2, 3, 4, ngrid);
  for (int i = if1; i <= if2; i++)
    for (int j = jlo; j <= jhi; j++)
      for (int k = klo; k <= khi; k++) {
        size_t xp = PIX(w, 0, i, j, k), xf = FIX(w, 0, i, j, k);
        for (int ivar = 0; ivar < nvar; ivar++)
          for (int l = 0; l < ngrid; l++) flux[(ivar + (size_t)0 * nvar) * nfp + xf + l] = w->fx[ivar * np + xp + l] * dt / dx;
        for (int ivar = 0; ivar < 2; ivar++)
          for (int l = 0; l < ngrid; l++) tmp[(ivar + (size_t)0 * 2) * nfp + xf + l] = w->tx[ivar * np + xp + l] * dt / dx;
      }
  /* Y :120-139 */
  if (ndim > 1) {
    cmpflxm(p, w, 0, 1, 0, ilo, ihi, jf1, jf2, klo, khi, 3, 2, 4, ngrid);
    for (int i = ilo; i <= ihi; i++)
      for (int j = jf1; j <= jf2; j++)
        for (int k = klo; k <= khi; k++) {
          size_t xp = PIX(w, 0, i, j, k), xf = FIX(w, 0, i, j, k);
          for (int ivar = 0; ivar < nvar; ivar++)
            for (int l = 0; l < ngrid; l++) flux[(ivar + (size_t)1 * nvar) * nfp + xf + l] = w->fx[ivar * np + xp + l] * dt / dy;
          for (int ivar = 0; ivar < 2; ivar++)
            for (int l = 0; l < ngrid; l++) tmp[(ivar + (size_t)1 * 2) * nfp + xf + l] = w->tx[ivar * np + xp + l] * dt / dy;
        }
  }
  /* Z :144-163 */
  if (ndim > 2) {
    cmpflxm(p, w, 0, 0, 1, ilo, ihi, jlo, jhi, kf1, kf2, 4, 2, 3, ngrid);
    for (int i = ilo; i <= ihi; i++)
      for (int j = jlo; j <= jhi; j++)
        for (int k = kf1; k <= kf2; k++) {
          size_t xp = PIX(w, 0, i, j, k), xf = FIX(w, 0, i, j, k);
          for (int ivar = 0; ivar < nvar; ivar++)
            for (int l = 0; l < ngrid; l++) flux[(ivar + (size_t)2 * nvar) * nfp + xf + l] = w->fx[ivar * np + xp + l] * dt / dz;
          for (int ivar = 0; ivar < 2; ivar++)
            for (int l = 0; l < ngrid; l++) tmp[(ivar + (size_t)2 * 2) * nfp + xf + l] = w->tx[ivar * np + xp + l] * dt / dz;
        }
  }
  if (p->difmag > 0.0) cmpdivu_consup(p, w, uin, flux, dx, dy, dz, dt, ngrid); /* :166-169 */
}

/* cmpdt hydro/godunov_utils.f90:5-120.  uu(nvector,nvar), gg(nvector,ndim)     */
void orc_cmpdt(const orc_params* p, double* uu, const double* gg, double dx, double* dt_out, int ncell) {
  const int nv = p->nvector, ndim = p->ndim;
  const double gamma = p->gamma, smallr = p->smallr, smallc = p->smallc;
  const double smallp = smallc * smallc / gamma;
#define UU(k, n) uu[(k) + (size_t)nv * ((n)-1)]
#define GG(k, n) gg[(k) + (size_t)nv * ((n)-1)]
  for (int k = 0; k < ncell; k++) UU(k, 1) = FMAX(UU(k, 1), smallr);
  for (int idim = 1; idim <= ndim; idim++)
    for (int k = 0; k < ncell; k++) UU(k, idim + 1) = UU(k, idim + 1) / UU(k, 1);
  for (int idim = 1; idim <= ndim; idim++)
    for (int k = 0; k < ncell; k++) UU(k, ndim + 2) = UU(k, ndim + 2) - half * UU(k, 1) * (UU(k, idim + 1) * UU(k, idim + 1));
  for (int k = 0; k < ncell; k++) UU(k, ndim + 2) = FMAX((gamma - one) * UU(k, ndim + 2), UU(k, 1) * smallp);
  for (int k = 0; k < ncell; k++) UU(k, ndim + 2) = gamma * UU(k, ndim + 2);
  for (int k = 0; k < ncell; k++) UU(k, ndim + 2) = sqrt(UU(k, ndim + 2) / UU(k, 1));
  for (int k = 0; k < ncell; k++) UU(k, ndim + 2) = (double)ndim * UU(k, ndim + 2);
  for (int idim = 1; idim <= ndim; idim++)
    for (int k = 0; k < ncell; k++) UU(k, ndim + 2) = UU(k, ndim + 2) + fabs(UU(k, idim + 1));
  for (int k = 0; k < ncell; k++) UU(k, 1) = zero;
  for (int idim = 1; idim <= ndim; idim++)
    for (int k = 0; k < ncell; k++) UU(k, 1) = UU(k, 1) + fabs(gg ? GG(k, idim) : 0.0);
  for (int k = 0; k < ncell; k++) {
    UU(k, 1) = UU(k, 1) * dx / (UU(k, ndim + 2) * UU(k, ndim + 2));
    UU(k, 1) = FMAX(UU(k, 1), 0.0001);
  }
  double dt = p->courant_factor * dx / smallc;
  for (int k = 0; k < ncell; k++) {
    double dtcell = dx / UU(k, ndim + 2) * (sqrt(one + two * p->courant_factor * UU(k, 1)) - one) / UU(k, 1);
    dt = FMIN(dt, dtcell);
  }
  *dt_out = dt;
#undef UU
#undef GG
}

/* cmpdivu hydro/uplmde.f90:702-764 + consup :769-869 (only when difmag>0)   */
static void cmpdivu_consup(const orc_params* p, orc_work* w, const double* uin, double* flux, double dx, double dy,
                           double dz, double dt, int ngrid) {
  const int ndim = p->ndim, nvar = p->nvar;
  const size_t np = w->np, nfp = w->nfp;
  const double* q = w->qin;
  double* div = (double*)calloc(nfp, 8);
  double hp = one; /* half**(ndim-1) */
  for (int d = 1; d < ndim; d++) hp = hp * half;
  const double factorx = hp / dx, factory = hp / dy, factorz = hp / dz;
  const int kf2 = ndim > 2 ? 3 : 1, jf2 = ndim > 1 ? 3 : 1;
#define Q(l, i, j, k, n) q[((n)-1) * np + PIX(w, l, i, j, k)]
#define U(l, i, j, k, n) uin[((n)-1) * np + PIX(w, l, i, j, k)]
#define DIV(l, i, j, k) div[FIX(w, l, i, j, k)]
  for (int k = 1; k <= kf2; k++)
    for (int j = 1; j <= jf2; j++)
      for (int i = 1; i <= 3; i++)
        for (int l = 0; l < ngrid; l++) {
          double ux = zero, vy = zero, wz = zero;
          ux = ux + factorx * (Q(l, i, j, k, 2) - Q(l, i - 1, j, k, 2));
          if (ndim > 1) {
            ux = ux + factorx * (Q(l, i, j - 1, k, 2) - Q(l, i - 1, j - 1, k, 2));
            vy = vy + factory * (Q(l, i, j, k, 3) - Q(l, i, j - 1, k, 3) + Q(l, i - 1, j, k, 3) - Q(l, i - 1, j - 1, k, 3));
          }
          if (ndim > 2) {
            ux = ux + factorx * (Q(l, i, j, k - 1, 2) - Q(l, i - 1, j, k - 1, 2) + Q(l, i, j - 1, k - 1, 2) - Q(l, i - 1, j - 1, k - 1, 2));
            vy = vy + factory * (Q(l, i, j, k - 1, 3) - Q(l, i, j - 1, k - 1, 3) + Q(l, i - 1, j, k - 1, 3) - Q(l, i - 1, j - 1, k - 1, 3));
            wz = wz + factorz * (Q(l, i, j, k, 4) - Q(l, i, j, k - 1, 4) + Q(l, i, j - 1, k, 4) - Q(l, i, j - 1, k - 1, 4) +
                                 Q(l, i - 1, j, k, 4) - Q(l, i - 1, j, k - 1, 4) + Q(l, i - 1, j - 1, k, 4) - Q(l, i - 1, j - 1, k - 1, 4));
          }
          DIV(l, i, j, k) = ux + vy + wz;
        }
  const double factor = hp, difmag = p->difmag;
  const int kmax = ndim > 2 ? 2 : 1, jmax = ndim > 1 ? 2 : 1; /* MAX(kf1,ku2-2), MAX(jf1,ju2-2) */
  for (int n = 1; n <= nvar; n++) {
    for (int k = 1; k <= kmax; k++)
      for (int j = 1; j <= jmax; j++)
        for (int i = 1; i <= 3; i++)
          for (int l = 0; l < ngrid; l++) {
            double div1 = factor * DIV(l, i, j, k);
            if (ndim > 1) div1 = div1 + factor * DIV(l, i, j + 1, k);
            if (ndim > 2) div1 = div1 + factor * (DIV(l, i, j, k + 1) + DIV(l, i, j + 1, k + 1));
            div1 = difmag * FMIN(zero, div1);
            double* f = &flux[((n - 1) + (size_t)0 * nvar) * nfp + FIX(w, l, i, j, k)];
            *f = *f + dt * div1 * (U(l, i, j, k, n) - U(l, i - 1, j, k, n));
          }
    if (ndim > 1)
      for (int k = 1; k <= kmax; k++)
        for (int j = 1; j <= 3; j++)
          for (int i = 1; i <= 2; i++)
            for (int l = 0; l < ngrid; l++) {
              double div1 = zero;
              div1 = div1 + factor * (DIV(l, i, j, k) + DIV(l, i + 1, j, k));
              if (ndim > 2) div1 = div1 + factor * (DIV(l, i, j, k + 1) + DIV(l, i + 1, j, k + 1));
              div1 = difmag * FMIN(zero, div1);
              double* f = &flux[((n - 1) + (size_t)1 * nvar) * nfp + FIX(w, l, i, j, k)];
              *f = *f + dt * div1 * (U(l, i, j, k, n) - U(l, i, j - 1, k, n));
            }
    if (ndim > 2)
      for (int k = 1; k <= 3; k++)
        for (int j = 1; j <= 2; j++)
          for (int i = 1; i <= 2; i++)
            for (int l = 0; l < ngrid; l++) {
              double div1 = factor * (DIV(l, i, j, k) + DIV(l, i + 1, j, k) + DIV(l, i, j + 1, k) + DIV(l, i + 1, j + 1, k));
              div1 = difmag * FMIN(zero, div1);
              double* f = &flux[((n - 1) + (size_t)2 * nvar) * nfp + FIX(w, l, i, j, k)];
              *f = *f + dt * div1 * (U(l, i, j, k, n) - U(l, i, j, k - 1, n));
            }
  }
#undef Q
#undef U
#undef DIV
  free(div);
}

/* ------------------------------------------------------------------------- */
/*                                 mesh                                       */
/* ------------------------------------------------------------------------- */
#define NBOR(m, ig, j) (m)->nbor[(size_t)((j)-1) * ((m)->ngridmax + 1) + (ig)]
#define XG(m, ig, d) (m)->xg[(size_t)(d) * ((m)->ngridmax + 1) + (ig)]

/* getindices3cube amr/nbors_utils.f90:305-358, GENERATED from geometry instead
 * of tabulated: for the father cell sitting at position `ind` of its oct, the
 * j-th of the 3^ndim neighbouring father cells (j = 1+i1+3*j1+9*k1, offsets
 * i1-1 etc.) lives in neighbour-oct slot lll (1+ii+2*jj+4*kk, ii=1 if the
 * offset leaves the oct along x) at cell position mmm.                         */
void orc_getindices3cube(int ndim, int ind, int lll[27], int mmm[27]) {
  int c[3] = {(ind - 1) & 1, ((ind - 1) >> 1) & 1, ((ind - 1) >> 2) & 1};
  int n3 = 1;
  for (int d = 0; d < ndim; d++) n3 *= 3;
  for (int j = 0; j < 27; j++) { lll[j] = 0; mmm[j] = 0; }
  for (int j = 0; j < n3; j++) {
    int o[3] = {j % 3 - 1, (j / 3) % 3 - 1, (j / 9) % 3 - 1};
    int g = 0, cc = 0;
    for (int d = 0; d < ndim; d++) {
      int t = c[d] + o[d];
      int out = (t < 0 || t > 1);
      g += out << d;
      cc += (((t % 2) + 2) % 2) << d;
    }
    lll[j] = 1 + g;
    mmm[j] = 1 + cc;
  }
}

/* get3cubefather amr/nbors_utils.f90:5-194 + get3cubepos :199-300 */
void orc_get3cubefather(const orc_mesh* m, int icf, int ilevel, int* nfc, int* nfg) {
  const int ndim = m->ndim, nx = m->nx, ny = m->ny, nz = m->nz, nxny = nx * ny;
  const int twotondim = ipow2(ndim);
  int threetondim = 1;
  for (int d = 0; d < ndim; d++) threetondim *= 3;
  if (ilevel == 1) {
    int iz = (icf - 1) / nxny;
    int iy = (icf - 1 - iz * nxny) / nx;
    int ix = (icf - 1 - iy * nx - iz * nxny);
    for (int k1 = 0; k1 <= (ndim > 2 ? 2 : 0); k1++) {
      int iiz = iz;
      if (ndim > 2) { iiz = iz + k1 - 1; if (iiz < 0) iiz = nz - 1; if (iiz > nz - 1) iiz = 0; }
      for (int j1 = 0; j1 <= (ndim > 1 ? 2 : 0); j1++) {
        int iiy = iy;
        if (ndim > 1) { iiy = iy + j1 - 1; if (iiy < 0) iiy = ny - 1; if (iiy > ny - 1) iiy = 0; }
        for (int i1 = 0; i1 <= 2; i1++) {
          int iix = ix + i1 - 1;
          if (iix < 0) iix = nx - 1;
          if (iix > nx - 1) iix = 0;
          nfc[i1 + 3 * j1 + 9 * k1] = 1 + iix + iiy * nx + iiz * nxny;
        }
      }
    }
    if (nfg) for (int j = 0; j < twotondim; j++) nfg[j] = 0; /* level-1 father grids unused by the hydro path */
    return;
  }
  int pos = (icf - m->ncoarse - 1) / m->ngridmax + 1;
  int igf = icf - m->ncoarse - (pos - 1) * m->ngridmax;
  /* get3cubepos :228-283: neighbour octs of the father oct towards the side the cell sits on */
  static const int iii[8] = {1, 2, 1, 2, 1, 2, 1, 2}, jjj[8] = {3, 3, 4, 4, 3, 3, 4, 4}, kkk[8] = {5, 5, 5, 5, 6, 6, 6, 6};
  int ng[8];
  for (int kk = 0; kk <= (ndim > 2 ? 1 : 0); kk++) {
    int g1 = igf;
    if (kk > 0 && igf > 0) g1 = m->son[NBOR(m, igf, kkk[pos - 1])];
    for (int jj = 0; jj <= (ndim > 1 ? 1 : 0); jj++) {
      int g2 = g1;
      if (jj > 0 && g1 > 0) g2 = m->son[NBOR(m, g1, jjj[pos - 1])];
      for (int ii = 0; ii <= 1; ii++) {
        int g3 = g2;
        if (ii > 0 && g2 > 0) g3 = m->son[NBOR(m, g2, iii[pos - 1])];
        ng[ii + 2 * jj + 4 * kk] = g3;
      }
    }
  }
  if (nfg) for (int j = 0; j < twotondim; j++) nfg[j] = ng[j];
  int lll[27], mmm[27];
  orc_getindices3cube(ndim, pos, lll, mmm);
  for (int j = 0; j < threetondim; j++) {
    int ig = ng[lll[j] - 1];
    nfc[j] = ig > 0 ? m->ncoarse + (mmm[j] - 1) * m->ngridmax + ig : 0;
  }
}

double orc_dx(const orc_params* p, const orc_mesh* m, int ilevel) {
  /* godunov_fine.f90:532-534 */
  int nx_loc = m->icoarse_max - m->icoarse_min + 1;
  double scale = p->boxlen / (double)nx_loc;
  return pow(0.5, ilevel) * scale;
}

/* integer position of an oct in units of its own size (walk father chain) */
void orc_mesh_oct_pos(const orc_mesh* m, int ilevel, int igrid, int pos[3]) {
  int ic = m->father[igrid];
  if (ilevel == 1) {
    int nxny = m->nx * m->ny;
    pos[2] = (ic - 1) / nxny;
    pos[1] = (ic - 1 - pos[2] * nxny) / m->nx;
    pos[0] = ic - 1 - pos[1] * m->nx - pos[2] * nxny;
    return;
  }
  int ind = (ic - m->ncoarse - 1) / m->ngridmax;
  int pg = ic - m->ncoarse - ind * m->ngridmax;
  int pp[3];
  orc_mesh_oct_pos(m, ilevel - 1, pg, pp);
  pos[0] = 2 * pp[0] + (ind & 1);
  pos[1] = 2 * pp[1] + ((ind >> 1) & 1);
  pos[2] = 2 * pp[2] + ((ind >> 2) & 1);
}

typedef struct { int pos[3]; int region; } octrec; /* region 0 domain, b>0 boundary b */

static unsigned lcg(unsigned* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

orc_mesh* orc_mesh_build_uniform(int ndim, int levelmax, const int bt[6], int order, unsigned seed) {
  orc_mesh* m = (orc_mesh*)calloc(1, sizeof(orc_mesh));
  m->ndim = ndim; m->nlevelmax = levelmax;
  const int twotondim = ipow2(ndim), twondim = 2 * ndim;
  /* coarse grid geometry: hydro/read_hydro_params.f90:316-350 */
  int nn[3] = {1, 1, 1}, cmin[3] = {0, 0, 0}, cmax[3] = {0, 0, 0};
  m->nboundary = 0;
  int bface[ORC_MAXBOUND]; /* which face each boundary region covers: 2*d+s */
  for (int d = 0; d < ndim; d++)
    for (int s = 0; s < 2; s++)
      if (bt[2 * d + s] > 0) {
        nn[d]++;
        if (s == 0) { cmin[d]++; cmax[d]++; }
        bface[m->nboundary] = 2 * d + s;
        m->boundary_type[m->nboundary] = (bt[2 * d + s] - 1) * 10 + (2 * d + s + 1); /* 1->1..6, 2->11..16 */
        m->nboundary++;
      }
  m->nx = nn[0]; m->ny = nn[1]; m->nz = nn[2];
  m->icoarse_min = cmin[0]; m->icoarse_max = cmax[0];
  m->jcoarse_min = cmin[1]; m->jcoarse_max = cmax[1];
  m->kcoarse_min = cmin[2]; m->kcoarse_max = cmax[2];
  m->ncoarse = nn[0] * nn[1] * nn[2];
  /* count octs: level l has the domain octs 2^(ndim*(l-1)) plus one layer of boundary octs */
  long total = 0;
  for (int l = 1; l <= levelmax; l++) {
    long dom = 1, all = 1;
    for (int d = 0; d < ndim; d++) {
      long n = 1L << (l - 1);
      dom *= n;
      all *= n + (nn[d] - 1);
    }
    total += (l == 1) ? m->ncoarse : all;
    (void)dom;
  }
  m->ngridmax = (int)total + 8;
  m->ncell = m->ncoarse + twotondim * m->ngridmax;
  m->son = (int*)calloc((size_t)m->ncell + 1, sizeof(int));
  m->cpu_map = (int*)calloc((size_t)m->ncell + 1, sizeof(int));
  m->father = (int*)calloc((size_t)m->ngridmax + 1, sizeof(int));
  m->nbor = (int*)calloc((size_t)twondim * (m->ngridmax + 1), sizeof(int));
  m->xg = (double*)calloc((size_t)ndim * (m->ngridmax + 1), sizeof(double));
  m->nactive = (int*)calloc(levelmax + 2, sizeof(int));
  m->active = (int**)calloc(levelmax + 2, sizeof(int*));
  m->nrecv = (int*)calloc(levelmax + 2, sizeof(int));
  m->recv = (int**)calloc(levelmax + 2, sizeof(int*));
  for (int b = 0; b < ORC_MAXBOUND; b++) {
    m->nbound[b] = (int*)calloc(levelmax + 2, sizeof(int));
    m->bound[b] = (int**)calloc(levelmax + 2, sizeof(int*));
  }
  /* region of a cell at integer position c[] at level l (cells per coarse cell = 2^l) */
  int next = 1;                 /* next free igrid */
  int nprev = 0; int* prev = NULL; /* octs of the previous level, creation order */
  octrec* rec = (octrec*)calloc((size_t)m->ngridmax + 1, sizeof(octrec));
  for (int l = 1; l <= levelmax; l++) {
    /* candidate father cells, in the chosen creation order */
    int ncand = 0;
    int* cand = (int*)malloc(sizeof(int) * (l == 1 ? (size_t)m->ncoarse : (size_t)nprev * twotondim));
    if (l == 1) {
      for (int ic = 1; ic <= m->ncoarse; ic++) cand[ncand++] = ic;
    } else {
      /* reference creation order: cell position outermost, parents in list order */
      for (int ind = 0; ind < twotondim; ind++)
        for (int a = 0; a < nprev; a++) cand[ncand++] = m->ncoarse + ind * m->ngridmax + prev[a];
    }
    /* keep cells that must be refined: domain cells, and boundary cells within one cell of the domain */
    int nkeep = 0;
    int* keep = (int*)malloc(sizeof(int) * (size_t)ncand);
    octrec* krec = (octrec*)malloc(sizeof(octrec) * (size_t)ncand);
    for (int a = 0; a < ncand; a++) {
      int ic = cand[a];
      int c[3] = {0, 0, 0};
      if (l == 1) {
        int nxny = m->nx * m->ny;
        c[2] = (ic - 1) / nxny; c[1] = (ic - 1 - c[2] * nxny) / m->nx; c[0] = ic - 1 - c[1] * m->nx - c[2] * nxny;
      } else {
        int ind = (ic - m->ncoarse - 1) / m->ngridmax;
        int pg = ic - m->ncoarse - ind * m->ngridmax;
        c[0] = 2 * rec[pg].pos[0] + (ind & 1);
        c[1] = 2 * rec[pg].pos[1] + ((ind >> 1) & 1);
        c[2] = 2 * rec[pg].pos[2] + ((ind >> 2) & 1);
      }
      /* cell c at level l-1: per coarse cell 2^(l-1) cells */
      int n1 = 1 << (l - 1);
      int region = 0, near = 1;
      for (int d = 0; d < ndim; d++) {
        int lo = cmin[d] * n1, hi = (cmax[d] + 1) * n1 - 1;
        if (c[d] < lo) {
          if (region == 0) for (int b = 0; b < m->nboundary; b++) if (bface[b] == 2 * d) region = b + 1;
          if (c[d] < lo - 1) near = 0;
        } else if (c[d] > hi) {
          if (region == 0) for (int b = 0; b < m->nboundary; b++) if (bface[b] == 2 * d + 1) region = b + 1;
          if (c[d] > hi + 1) near = 0;
        }
      }
      if (region == 0 || near) {
        keep[nkeep] = ic;
        krec[nkeep].pos[0] = c[0]; krec[nkeep].pos[1] = c[1]; krec[nkeep].pos[2] = c[2];
        krec[nkeep].region = region;
        nkeep++;
      }
    }
    /* optional re-ordering of creation */
    int* perm = (int*)malloc(sizeof(int) * (size_t)nkeep);
    for (int a = 0; a < nkeep; a++) perm[a] = a;
    if (order == 1) {
      /* lattice row-major by position: counting sort via keys (positions are small) */
      long ext[3] = {1, 1, 1};
      for (int d = 0; d < ndim; d++) ext[d] = (long)nn[d] << (l - 1);
      long* key = (long*)malloc(sizeof(long) * (size_t)nkeep);
      for (int a = 0; a < nkeep; a++) key[a] = krec[a].pos[0] + ext[0] * (krec[a].pos[1] + ext[1] * (long)krec[a].pos[2]);
      /* simple stable merge-free approach: build inverse map (keys unique) */
      long nkeys = ext[0] * ext[1] * ext[2];
      int* inv = (int*)malloc(sizeof(int) * (size_t)nkeys);
      for (long q = 0; q < nkeys; q++) inv[q] = -1;
      for (int a = 0; a < nkeep; a++) inv[key[a]] = a;
      int c2 = 0;
      for (long q = 0; q < nkeys; q++) if (inv[q] >= 0) perm[c2++] = inv[q];
      free(inv); free(key);
    } else if (order == 2) {
      unsigned s = seed + 977u * (unsigned)l;
      for (int a = nkeep - 1; a > 0; a--) { int b = (int)(lcg(&s) % (unsigned)(a + 1)); int t = perm[a]; perm[a] = perm[b]; perm[b] = t; }
    }
    /* create the octs */
    int* cur = (int*)malloc(sizeof(int) * (size_t)nkeep);
    int nact = 0, nb[ORC_MAXBOUND] = {0, 0, 0, 0, 0, 0};
    for (int a = 0; a < nkeep; a++) { if (krec[a].region == 0) nact++; else nb[krec[a].region - 1]++; }
    m->active[l] = (int*)malloc(sizeof(int) * (size_t)(nact + 1));
    for (int b = 0; b < m->nboundary; b++) m->bound[b][l] = (int*)malloc(sizeof(int) * (size_t)(nb[b] + 1));
    m->recv[l] = (int*)malloc(sizeof(int));
    for (int a = 0; a < nkeep; a++) {
      int k = perm[a];
      int ig = next++;
      if (ig > m->ngridmax) { fprintf(stderr, "orc mesh: ngridmax overflow\n"); abort(); }
      cur[a] = ig;
      m->father[ig] = keep[k];
      m->son[keep[k]] = ig;
      rec[ig] = krec[k];
      double dxl = pow(0.5, l);
      for (int d = 0; d < ndim; d++) XG(m, ig, d) = (rec[ig].pos[d] + 0.5) * 2.0 * dxl; /* oct centre, coarse-cell units */
      if (rec[ig].region == 0) m->active[l][m->nactive[l]++] = ig;
      else { int b = rec[ig].region - 1; m->bound[b][l][m->nbound[b][l]++] = ig; }
      for (int ind = 0; ind < twotondim; ind++) m->cpu_map[m->ncoarse + ind * m->ngridmax + ig] = rec[ig].region == 0 ? 1 : 0;
    }
    /* neighbours: nbor(igrid,j) = neighbouring father cell (refine_utils.f90:213-263 level 1,
     * :685-691 deeper levels through getnborfather amr/nbors_utils.f90:404)            */
    for (int a = 0; a < nkeep; a++) {
      int ig = cur[a];
      int ic = m->father[ig];
      if (l == 1) {
        int c[3] = {rec[ig].pos[0], rec[ig].pos[1], rec[ig].pos[2]};
        int str[3] = {1, m->nx, m->nx * m->ny};
        for (int d = 0; d < ndim; d++) {
          int cm = c[d] > 0 ? ic - str[d] : ic + (nn[d] - 1) * str[d];
          int cp = c[d] < nn[d] - 1 ? ic + str[d] : ic - (nn[d] - 1) * str[d];
          NBOR(m, ig, 2 * d + 1) = cm;
          NBOR(m, ig, 2 * d + 2) = cp;
        }
      } else {
        int ind = (ic - m->ncoarse - 1) / m->ngridmax;
        int pg = ic - m->ncoarse - ind * m->ngridmax;
        for (int d = 0; d < ndim; d++)
          for (int s = 0; s < 2; s++) {
            int bit = (ind >> d) & 1;
            int ind2 = ind ^ (1 << d);
            int res;
            if (bit != s) res = m->ncoarse + ind2 * m->ngridmax + pg; /* inside the same father oct */
            else {
              int ng = m->son[NBOR(m, pg, 2 * d + s + 1)];
              res = ng > 0 ? m->ncoarse + ind2 * m->ngridmax + ng : NBOR(m, pg, 2 * d + s + 1);
            }
            NBOR(m, ig, 2 * d + s + 1) = res;
          }
      }
    }
    free(prev); prev = cur; nprev = nkeep;
    free(cand); free(keep); free(krec); free(perm);
  }
  free(prev); free(rec);
  m->ngrid_used = next - 1;
  return m;
}

void orc_mesh_free(orc_mesh* m) {
  if (!m) return;
  for (int l = 0; l <= m->nlevelmax + 1; l++) {
    if (m->active) free(m->active[l]);
    if (m->recv) free(m->recv[l]);
    for (int b = 0; b < ORC_MAXBOUND; b++) if (m->bound[b]) free(m->bound[b][l]);
  }
  free(m->active); free(m->nactive); free(m->recv); free(m->nrecv);
  for (int b = 0; b < ORC_MAXBOUND; b++) { free(m->bound[b]); free(m->nbound[b]); }
  free(m->son); free(m->cpu_map); free(m->father); free(m->nbor); free(m->xg); free(m);
}

/* ------------------------------------------------------------------------- */
/*                             per-level passes                               */
/* ------------------------------------------------------------------------- */
#define UO(ic, iv) uold[(size_t)((iv)-1) * m->ncell + (ic)-1]
#define UN(ic, iv) unew[(size_t)((iv)-1) * m->ncell + (ic)-1]

/* condinit hydro/condinit.f90:5-73 + region_condinit hydro/init_flow_fine.f90:475-596;
 * cell centres as init_flow_fine.f90:74-91                                       */
void orc_condinit_regions(const orc_params* p, const orc_mesh* m, int ilevel, double* uold, int nregion,
                          const int* region_type, const double* x_center, const double* y_center,
                          const double* z_center, const double* length_x, const double* length_y,
                          const double* length_z, const double* exp_region, const double* d_region,
                          const double* u_region, const double* v_region, const double* w_region,
                          const double* p_region) {
  const int ndim = p->ndim, nvar = p->nvar, twotondim = ipow2(ndim);
  const double gamma = p->gamma;
  int nx_loc = m->icoarse_max - m->icoarse_min + 1;
  double scale = p->boxlen / (double)nx_loc;
  double dx = pow(0.5, ilevel);
  double dx_loc = dx * scale;
  double skip_loc[3] = {(double)m->icoarse_min, (double)m->jcoarse_min, (double)m->kcoarse_min};
  /* all octs of the level, domain and boundary alike (init_flow_fine loops over active grids only;
   * boundary octs are then filled by make_boundary_hydro) */
  for (int a = 0; a < m->nactive[ilevel]; a++) {
    int ig = m->active[ilevel][a];
    for (int ind = 0; ind < twotondim; ind++) {
      double xc[3] = {((ind & 1) - 0.5) * dx, (((ind >> 1) & 1) - 0.5) * dx, (((ind >> 2) & 1) - 0.5) * dx};
      double x[3] = {0, 0, 0};
      for (int d = 0; d < ndim; d++) x[d] = (XG(m, ig, d) + xc[d] - skip_loc[d]) * scale;
      double q[16];
      q[1] = p->smallr; q[2] = 0; q[3] = 0; q[4] = 0;
      q[ndim + 2] = p->smallr * p->smallc * p->smallc / gamma;
      for (int iv = ndim + 3; iv <= nvar; iv++) q[iv] = 0;
      for (int k = 0; k < nregion; k++) {
        if (region_type[k] == 0) { /* square */
          double en = exp_region[k];
          double xn = 0, yn = 0, zn = 0, r;
          xn = 2.0 * fabs(x[0] - x_center[k]) / length_x[k];
          if (ndim > 1) yn = 2.0 * fabs(x[1] - y_center[k]) / length_y[k];
          if (ndim > 2) zn = 2.0 * fabs(x[2] - z_center[k]) / length_z[k];
          if (exp_region[k] < 10) r = pow(pow(xn, en) + pow(yn, en) + pow(zn, en), 1.0 / en);
          else r = FMAX(FMAX(xn, yn), zn);
          if (r < 1.0) {
            q[1] = d_region[k]; q[2] = u_region[k];
            if (ndim > 1) q[3] = v_region[k];
            if (ndim > 2) q[4] = w_region[k];
            q[ndim + 2] = p_region[k];
          }
        } else { /* point */
          double vol = 1;
          for (int d = 0; d < ndim; d++) vol *= dx_loc; /* dx**ndim */
          double xn = 1, yn = 1, zn = 1;
          xn = FMAX(1.0 - fabs(x[0] - x_center[k]) / dx_loc, 0.0);
          if (ndim > 1) yn = FMAX(1.0 - fabs(x[1] - y_center[k]) / dx_loc, 0.0);
          if (ndim > 2) zn = FMAX(1.0 - fabs(x[2] - z_center[k]) / dx_loc, 0.0);
          double r = xn * yn * zn;
          q[1] = q[1] + d_region[k] * r / vol;
          q[2] = q[2] + u_region[k] * r;
          if (ndim > 1) q[3] = q[3] + v_region[k] * r;
          if (ndim > 2) q[4] = q[4] + w_region[k] * r;
          q[ndim + 2] = q[ndim + 2] + p_region[k] * r / vol;
        }
      }
      /* primitive -> conservative condinit.f90:33-66 */
      int ic = m->ncoarse + ind * m->ngridmax + ig;
      UO(ic, 1) = q[1];
      UO(ic, 2) = q[1] * q[2];
      if (ndim > 1) UO(ic, 3) = q[1] * q[3];
      if (ndim > 2) UO(ic, 4) = q[1] * q[4];
      double e = 0.0;
      e = e + 0.5 * q[1] * (q[2] * q[2]);
      if (ndim > 1) e = e + 0.5 * q[1] * (q[3] * q[3]);
      if (ndim > 2) e = e + 0.5 * q[1] * (q[4] * q[4]);
      e = e + q[ndim + 2] / (gamma - 1.0);
      UO(ic, ndim + 2) = e;
      for (int iv = ndim + 3; iv <= nvar; iv++) UO(ic, iv) = q[1] * q[iv];
    }
  }
}

/* set_unew hydro/godunov_fine.f90:40-130 */
/* pressure_fix (amr_parameters.f90:167,197): the two extra cell arrays of hydro_commons -- divu (= -div(u)*dt accumulated from
 * the face velocities tmp(:,1)) and enew (internal energy advanced with the flux tmp(:,2) and the -pdV source) -- are owned by
 * the caller; NULL switches the option off.  hexp = 0 (no cosmology).                                                       */
static double *g_divu = NULL, *g_enew = NULL;
static double g_beta_fix = 0.0, g_dt_level[64];
/* poisson=.true.: the acceleration f(1:ncell,1:ndim) of poisson_commons (amr/init_poisson.f90), owned by the caller; NULL = off.
 * The Poisson solver itself is out of scope: f is an input (an analytic field, gravity_type>0, or whatever the host computed). */
static const double* g_force = NULL;
void orc_set_gravity(const double* f) { g_force = f; }
#define FO(ic, idim) g_force[(size_t)((idim)-1) * m->ncell + (ic)-1]
void orc_set_pressure_fix(double* divu, double* enew, double beta_fix) { g_divu = divu; g_enew = enew; g_beta_fix = beta_fix; }

void orc_set_unew(const orc_params* p, const orc_mesh* m, int ilevel, const double* uold, double* unew) {
  const int twotondim = ipow2(p->ndim);
  const int na = m->nactive[ilevel];
  const int* act = m->active[ilevel];
#ifdef _OPENMP
#pragma omp parallel num_threads(g_nthreads) if (g_nthreads > 1)
#endif
  for (int ind = 0; ind < twotondim; ind++) {
    const int iskip = m->ncoarse + ind * m->ngridmax;
    for (int iv = 1; iv <= p->nvar; iv++) {
#ifdef _OPENMP
#pragma omp for schedule(static) nowait
#endif
      for (int a = 0; a < na; a++) UN(act[a] + iskip, iv) = UO(act[a] + iskip, iv);
    }
  }
  if (g_divu)                                   /* pressure_fix :71-90 */
    for (int ind = 0; ind < twotondim; ind++) {
      const int iskip = m->ncoarse + ind * m->ngridmax, ndim = p->ndim;
      for (int a = 0; a < na; a++) g_divu[act[a] + iskip - 1] = 0;
      for (int a = 0; a < na; a++) {
        const int ic = act[a] + iskip;
        double d = FMAX(UO(ic, 1), p->smallr), u = 0, v = 0, w = 0;
        if (ndim > 0) u = UO(ic, 2) / d;
        if (ndim > 1) v = UO(ic, 3) / d;
        if (ndim > 2) w = UO(ic, 4) / d;
        g_enew[ic - 1] = UO(ic, ndim + 2) - 0.5 * d * (u * u + v * v + w * w);
      }
    }
  for (int ind = 0; ind < twotondim; ind++) { /* :93-126 */
    int iskip = m->ncoarse + ind * m->ngridmax;
    for (int iv = 1; iv <= p->nvar; iv++)
      for (int a = 0; a < m->nrecv[ilevel]; a++) UN(m->recv[ilevel][a] + iskip, iv) = 0;
    if (g_divu)
      for (int a = 0; a < m->nrecv[ilevel]; a++) { g_divu[m->recv[ilevel][a] + iskip - 1] = 0; g_enew[m->recv[ilevel][a] + iskip - 1] = 0; }
  }
}

/* add_pdv_source_terms hydro/godunov_fine.f90:294-437 (pressure_fix part): enew -= (gamma-1) e_old div(u) dt with the
 * velocity divergence from the face-neighbour cells of uold (the coarser father cell at 1.5 dx where no neighbour oct exists) */
static void add_pdv_source_terms(const orc_params* p, const orc_mesh* m, int ilevel, const double* uold) {
  static const int iii[3][2][8] = {{{1, 0, 1, 0, 1, 0, 1, 0}, {0, 2, 0, 2, 0, 2, 0, 2}},
                                   {{3, 3, 0, 0, 3, 3, 0, 0}, {0, 0, 4, 4, 0, 0, 4, 4}},
                                   {{5, 5, 5, 5, 0, 0, 0, 0}, {0, 0, 0, 0, 6, 6, 6, 6}}};
  static const int jjj[3][2][8] = {{{2, 1, 4, 3, 6, 5, 8, 7}, {2, 1, 4, 3, 6, 5, 8, 7}},
                                   {{3, 4, 1, 2, 7, 8, 5, 6}, {3, 4, 1, 2, 7, 8, 5, 6}},
                                   {{5, 6, 7, 8, 1, 2, 3, 4}, {5, 6, 7, 8, 1, 2, 3, 4}}};
  const int ndim = p->ndim, twotondim = ipow2(ndim);
  const double dx_loc = orc_dx(p, m, ilevel);
  const double dt = g_dt_level[ilevel];
  for (int a = 0; a < m->nactive[ilevel]; a++) {
    const int ig = m->active[ilevel][a];
    int igridn[7], ind_left[3], ind_right[3];
    igridn[0] = ig;
    for (int d = 0; d < ndim; d++) {
      ind_left[d] = NBOR(m, ig, 2 * d + 1); ind_right[d] = NBOR(m, ig, 2 * d + 2);
      igridn[2 * d + 1] = m->son[ind_left[d]]; igridn[2 * d + 2] = m->son[ind_right[d]];
    }
    for (int ind = 0; ind < twotondim; ind++) {
      const int ic = m->ncoarse + ind * m->ngridmax + ig;
      double divu_loc = 0.0;
      for (int d = 0; d < ndim; d++) {
        double velg, veld, dx_g, dx_d;
        const int g1 = igridn[iii[d][0][ind]], c1 = g1 > 0 ? g1 + m->ncoarse + (jjj[d][0][ind] - 1) * m->ngridmax : ind_left[d];
        velg = UO(c1, d + 2) / FMAX(UO(c1, 1), p->smallr);
        dx_g = g1 > 0 ? dx_loc : dx_loc * 1.5;
        const int g2 = igridn[iii[d][1][ind]], c2 = g2 > 0 ? g2 + m->ncoarse + (jjj[d][1][ind] - 1) * m->ngridmax : ind_right[d];
        veld = UO(c2, d + 2) / FMAX(UO(c2, 1), p->smallr);
        dx_d = g2 > 0 ? dx_loc : dx_loc * 1.5;
        divu_loc = divu_loc + (veld - velg) / (dx_g + dx_d);
      }
      double dd = FMAX(UO(ic, 1), p->smallr), u = 0, v = 0, w = 0;
      if (ndim > 0) u = UO(ic, 2) / dd;
      if (ndim > 1) v = UO(ic, 3) / dd;
      if (ndim > 2) w = UO(ic, 4) / dd;
      const double eold = UO(ic, ndim + 2) - 0.5 * dd * (u * u + v * v + w * w);
      g_enew[ic - 1] = g_enew[ic - 1] - (p->gamma - 1.0) * eold * divu_loc * dt;
    }
  }
}

/* add_gravity_source_terms hydro/godunov_fine.f90:237-289: momentum and total energy of unew get the half-step kick
 * f*dt/2 weighted by the OLD density (strict_equilibrium = 0)                                                          */
static void add_gravity_source_terms(const orc_params* p, const orc_mesh* m, int ilevel, const double* uold, double* unew) {
  const int ndim = p->ndim, twotondim = ipow2(ndim);
  const double req = 0.0;
  for (int ind = 0; ind < twotondim; ind++) {
    const int iskip = m->ncoarse + ind * m->ngridmax;
    for (int a = 0; a < m->nactive[ilevel]; a++) {
      const int ic = m->active[ilevel][a] + iskip;
      double d = FMAX(UN(ic, 1), p->smallr), u = 0, v = 0, w = 0;
      if (ndim > 0) u = UN(ic, 2) / d;
      if (ndim > 1) v = UN(ic, 3) / d;
      if (ndim > 2) w = UN(ic, 4) / d;
      double e_kin = 0.5 * d * (u * u + v * v + w * w);
      const double e_prim = UN(ic, ndim + 2) - e_kin;
      const double d_old = FMAX(UO(ic, 1), p->smallr);
      const double fact = (d_old - req) / d * 0.5 * g_dt_level[ilevel];
      if (ndim > 0) { u = u + FO(ic, 1) * fact; UN(ic, 2) = d * u; }
      if (ndim > 1) { v = v + FO(ic, 2) * fact; UN(ic, 3) = d * v; }
      if (ndim > 2) { w = w + FO(ic, 3) * fact; UN(ic, 4) = d * w; }
      e_kin = 0.5 * d * (u * u + v * v + w * w);
      UN(ic, ndim + 2) = e_prim + e_kin;
    }
  }
}

/* set_uold hydro/godunov_fine.f90:135-232 */
void orc_set_uold(const orc_params* p, const orc_mesh* m, int ilevel, double* uold, const double* unew) {
  const int twotondim = ipow2(p->ndim), ndim = p->ndim, nvar = p->nvar;
  const double smallr = p->smallr;
  if (g_force) add_gravity_source_terms(p, m, ilevel, uold, (double*)unew);   /* :159-161 */
  if (g_divu) add_pdv_source_terms(p, m, ilevel, uold);             /* :164-168 */
  for (int ind = 0; ind < twotondim; ind++) {
    int iskip = m->ncoarse + ind * m->ngridmax;
    if (nvar > ndim + 2) { /* :176-190 passive-scalar floor fix */
      for (int a = 0; a < m->nactive[ilevel]; a++) {
        int ic = m->active[ilevel][a] + iskip;
        if (UO(ic, 1) < smallr && UN(ic, 1) > UO(ic, 1)) {
          for (int iv = ndim + 3; iv <= nvar; iv++)
            ((double*)unew)[(size_t)(iv - 1) * m->ncell + ic - 1] = UO(ic, iv) * FMAX(UN(ic, 1), smallr) / smallr;
        } else if (UN(ic, 1) < smallr && UO(ic, 1) > UN(ic, 1)) {
          for (int iv = ndim + 3; iv <= nvar; iv++)
            ((double*)unew)[(size_t)(iv - 1) * m->ncell + ic - 1] = UO(ic, iv) * smallr / FMAX(UO(ic, 1), smallr);
        }
      }
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(g_nthreads) if (g_nthreads > 1)
#endif
    for (int iv = 1; iv <= nvar; iv++) {
#ifdef _OPENMP
#pragma omp for schedule(static) nowait
#endif
      for (int a = 0; a < m->nactive[ilevel]; a++) UO(m->active[ilevel][a] + iskip, iv) = UN(m->active[ilevel][a] + iskip, iv);
    }
    if (g_divu) {                                                     /* correct total energy if internal energy is too small :203-227 */
      const double dx = orc_dx(p, m, ilevel), hexp = 0.0;
      for (int a = 0; a < m->nactive[ilevel]; a++) {
        const int ic = m->active[ilevel][a] + iskip;
        double d = FMAX(UO(ic, 1), smallr), u = 0, v = 0, w = 0;
        if (ndim > 0) u = UO(ic, 2) / d;
        if (ndim > 1) v = UO(ic, 3) / d;
        if (ndim > 2) w = UO(ic, 4) / d;
        const double e_kin = 0.5 * d * (u * u + v * v + w * w);
        const double e_cons = UO(ic, ndim + 2) - e_kin;
        const double e_prim = g_enew[ic - 1];
        const double div = fabs(g_divu[ic - 1]) * dx / g_dt_level[ilevel];     /* divu = -div.u*dt */
        const double mx = FMAX(div, 3.0 * hexp * dx);
        const double e_trunc = g_beta_fix * d * (mx * mx);
        if (e_cons < e_trunc) UO(ic, ndim + 2) = e_prim + e_kin;
      }
    }
  }
}


/* ------------------------------------------------------------------------- */
/*                AMR pieces: getnborfather, interpol_hydro, upl              */
/* ------------------------------------------------------------------------- */
static int g_interpol_type = 1, g_interpol_var = 0;   /* hydro_parameters.f90:88-89 */
void orc_set_interpol(int type, int var) { g_interpol_type = type; g_interpol_var = var; }

/* getnborfather amr/nbors_utils.f90:404-525 for ONE cell of level ilevel-1 (ilevel==1: a coarse cell):
 * ind_father[0] = the cell, ind_father[1..2*ndim] = its neighbours at the same level, or the coarser
 * neighbouring father cell where that neighbour oct does not exist.                                   */
void orc_getnborfather(const orc_mesh* m, int ind_cell, int ilevel, int* ind_father) {
  const int ndim = m->ndim, nx = m->nx, ny = m->ny, nz = m->nz, nxny = nx * ny;
  ind_father[0] = ind_cell;
  if (ilevel == 1) {
    const int ibound[3] = {nx - 1, ny - 1, nz - 1};
    const int iskip1[3] = {1, nx, nxny}, iskip2[3] = {nx - 1, (ny - 1) * nx, (nz - 1) * nxny};
    int ix[3];
    ix[2] = (ind_cell - 1) / nxny;
    ix[1] = (ind_cell - 1 - ix[2] * nxny) / nx;
    ix[0] = (ind_cell - 1 - ix[1] * nx - ix[2] * nxny);
    for (int d = 0; d < ndim; d++) {
      ind_father[2 * d + 1] = ix[d] > 0 ? ind_cell - iskip1[d] : ind_cell + iskip2[d];
      ind_father[2 * d + 2] = ix[d] < ibound[d] ? ind_cell + iskip1[d] : ind_cell - iskip2[d];
    }
    return;
  }
  const int pos = (ind_cell - m->ncoarse - 1) / m->ngridmax;             /* 0-based cell position */
  const int gf = ind_cell - m->ncoarse - pos * m->ngridmax;
  for (int d = 0; d < ndim; d++)
    for (int s = 0; s < 2; s++) {
      const int j = 2 * d + s + 1;
      const int bit = (pos >> d) & 1, pos2 = pos ^ (1 << d);
      int g;                                                               /* getnborgrids / getnborcells :363,:530 */
      if (bit != s) g = gf; else g = m->son[NBOR(m, gf, j)];
      ind_father[j] = g > 0 ? m->ncoarse + pos2 * m->ngridmax + g : NBOR(m, gf, j);
    }
}

/* the three slope routines of hydro/interpol_hydro.f90 for one variable: a[0..2*ndim] -> w[ndim] */
static void interpol_slopes(int ndim, int type, const double* a, double* w, const double xc[8][3]) {
  const int twotondim = ipow2(ndim), twondim = 2 * ndim;
  for (int d = 0; d < ndim; d++) w[d] = 0.0;
  if (type == 1) { /* compute_limiter_minmod :449-470 */
    for (int d = 0; d < ndim; d++) {
      double dl = 0.5 * (a[2 * d + 2] - a[0]), dr = 0.5 * (a[0] - a[2 * d + 1]), mm;
      if (dl * dr <= 0.0) mm = 0; else mm = FMIN(fabs(dl), fabs(dr)) * dl / fabs(dl);
      w[d] = mm;
    }
  } else if (type == 2) { /* compute_limiter_central :481-613 */
    double ac[8];
    for (int d = 0; d < ndim; d++) w[d] = 0.25 * (a[2 * d + 2] - a[2 * d + 1]);
    for (int ind = 0; ind < twotondim; ind++) ac[ind] = a[0];
    for (int d = 0; d < ndim; d++)
      for (int ind = 0; ind < twotondim; ind++) ac[ind] = ac[ind] + 2.0 * w[d] * xc[ind][d];
    double corner = ac[0], kernel = a[1];
    for (int j = 1; j < twotondim; j++) corner = FMAX(corner, ac[j]);
    for (int j = 2; j <= twondim; j++) kernel = FMAX(kernel, a[j]);
    double dk = a[0] - kernel, dc = a[0] - corner, maxl = 0.0, minl = 0.0;
    if (dk * dc > 0.0) maxl = FMIN(1.0, dk / dc);
    corner = ac[0]; kernel = a[1];
    for (int j = 1; j < twotondim; j++) corner = FMIN(corner, ac[j]);
    for (int j = 2; j <= twondim; j++) kernel = FMIN(kernel, a[j]);
    dk = a[0] - kernel; dc = a[0] - corner;
    if (dk * dc > 0.0) minl = FMIN(1.0, dk / dc);
    const double lim = FMIN(minl, maxl);
    for (int d = 0; d < ndim; d++) w[d] = w[d] * lim;
  } else if (type == 3) { /* compute_central :618-637 */
    for (int d = 0; d < ndim; d++) w[d] = 0.25 * (a[2 * d + 2] - a[2 * d + 1]);
  }
}

/* interpol_hydro hydro/interpol_hydro.f90:268-444 for one father cell: u1[(2*ndim+1)][nvar] -> u2[2^ndim][nvar].
 * interpol_var 0 (rho, rho u, E), 1 (rho, rho u, rho eps), 2 (rho, u, rho eps + momentum correction :393-415);
 * interpol_type 0,1,2,3 and 4 (type 3 for the velocities, type 2 for the rest; needs interpol_var=2 :357-366).           */
void orc_interpol_hydro(const orc_params* p, const double* u1_in, double* u2) {
  const int ndim = p->ndim, nvar = p->nvar, twotondim = ipow2(ndim), twondim = 2 * ndim;
  const int var = g_interpol_var, type = g_interpol_type;
  if (var < 0 || var > 2 || type < 0 || type > 4 || (type == 4 && var != 2)) {
    fprintf(stderr, "orc: interpol_var=%d interpol_type=%d not valid (type 4 is designed for interpol_var=2)\n", var, type);
    abort();
  }
  const double oneover_twotondim = 1.0 / (double)twotondim;
  double xc[8][3];
  for (int ind = 0; ind < twotondim; ind++) {
    xc[ind][0] = (double)(ind & 1) - 0.5; xc[ind][1] = (double)((ind >> 1) & 1) - 0.5; xc[ind][2] = (double)((ind >> 2) & 1) - 0.5;
  }
  double u1[7 * 16];
  for (int j = 0; j <= twondim; j++)
    for (int iv = 0; iv < nvar; iv++) u1[j * nvar + iv] = u1_in[j * nvar + iv];
#define U1(j, iv) u1[(j)*nvar + (iv)-1]
#define U2(ind, iv) u2[(ind)*nvar + (iv)-1]
  if (var == 1 || var == 2) { /* father total energy -> internal energy :318-345, momenta -> velocities for var 2 */
    for (int j = 0; j <= twondim; j++) {
      double ekin = 0.0;
      for (int d = 1; d <= ndim; d++) ekin = ekin + 0.5 * (U1(j, d + 1) * U1(j, d + 1)) / FMAX(U1(j, 1), p->smallr);
      const double erad = 0.0;
      U1(j, ndim + 2) = U1(j, ndim + 2) - ekin - erad;
      if (var == 2)
        for (int d = 1; d <= ndim; d++) U1(j, d + 1) = U1(j, d + 1) / FMAX(U1(j, 1), p->smallr);
    }
  }
  for (int iv = 1; iv <= nvar; iv++) {
    double a[7], w[3];
    for (int j = 0; j <= twondim; j++) a[j] = U1(j, iv);
    int t = type;
    if (type == 4) t = (iv > 1 && iv <= 1 + ndim) ? 3 : 2;
    interpol_slopes(ndim, t, a, w, xc);
    for (int ind = 0; ind < twotondim; ind++) { /* :372-379 */
      double v = a[0];
      for (int d = 0; d < ndim; d++) v = v + w[d] * xc[ind][d];
      U2(ind, iv) = v;
    }
  }
  if (var == 1 || var == 2) {
    if (var == 2) {
      for (int ind = 0; ind < twotondim; ind++)
        for (int d = 1; d <= ndim; d++) U2(ind, d + 1) = U2(ind, d + 1) * U2(ind, 1);
      for (int d = 1; d <= ndim; d++) { /* correct the total momentum keeping the slope fixed :399-413 */
        double mom = 0.0;
        for (int ind = 0; ind < twotondim; ind++) mom = mom + U2(ind, d + 1) * oneover_twotondim;
        mom = mom - U1(0, d + 1) * U1(0, 1);
        for (int ind = 0; ind < twotondim; ind++) U2(ind, d + 1) = U2(ind, d + 1) - mom;
      }
    }
    for (int ind = 0; ind < twotondim; ind++) { /* children internal energy -> total energy :418-440 */
      double ekin = 0.0;
      for (int d = 1; d <= ndim; d++) ekin = ekin + 0.5 * (U2(ind, d + 1) * U2(ind, d + 1)) / FMAX(U2(ind, 1), p->smallr);
      const double erad = 0.0;
      U2(ind, ndim + 2) = U2(ind, ndim + 2) + ekin + erad;
    }
  }
#undef U1
#undef U2
}

/* prolongation of one father cell from the coarse state: what godfine1 (:583-593) and make_grid_fine
 * (amr/refine_utils.f90:136-167) do for a missing / newly created oct                                            */
void orc_interpol_cell(const orc_params* p, const orc_mesh* m, int ind_cell, int ilevel, const double* uold, double* u2) {
  int fa[7];
  double u1[7 * 16];
  orc_getnborfather(m, ind_cell, ilevel, fa);
  for (int j = 0; j <= 2 * p->ndim; j++)
    for (int iv = 1; iv <= p->nvar; iv++) u1[j * p->nvar + iv - 1] = uold[(size_t)(iv - 1) * m->ncell + fa[j] - 1];
  orc_interpol_hydro(p, u1, u2);
}

/* godfine1 hydro/godunov_fine.f90:486-911 for one batch of octs */
/* phase 0: the whole routine; phase 1: gather + fluxes + flux reset only (reads uold, writes the work arrays);
 * phase 2: the updates of unew only, from the work arrays left by phase 1                                        */
static void godfine1(const orc_params* p, const orc_mesh* m, orc_work* w, const int* ind_grid, int ncache, int ilevel,
                     double dt, const double* uold, double* unew, int phase) {
  const int ndim = p->ndim, nvar = p->nvar, twotondim = ipow2(ndim);
  const size_t np = w->np, nfp = w->nfp;
  const double oneontwotondim = 1.0 / (double)twotondim;
  const double dx = orc_dx(p, m, ilevel);
  const int i1max = 2, j1max = ndim > 1 ? 2 : 0, k1max = ndim > 2 ? 2 : 0;
  const int i2max = 1, j2max = ndim > 1 ? 1 : 0, k2max = ndim > 2 ? 1 : 0;
  const int i3min = 1, i3max = 2, j3min = 1, j3max = ndim > 1 ? 2 : 1, k3min = 1, k3max = ndim > 2 ? 2 : 1;
  if (phase != 2) {
  /* gather 3^ndim neighbouring father cells :553-556 and the 6^ndim stencil :562-675 */
  for (int i = 0; i < ncache; i++) {
    int nfc[27];
    orc_get3cubefather(m, m->father[ind_grid[i]], ilevel, nfc, NULL);
    for (int k1 = 0; k1 <= k1max; k1++)
      for (int j1 = 0; j1 <= j1max; j1++)
        for (int i1 = 0; i1 <= i1max; i1++) {
          int ind_father = i1 + 3 * j1 + 9 * k1;
          int igrid_nbor = m->son[nfc[ind_father]];
          double u2[8 * 16];
          if (igrid_nbor <= 0) orc_interpol_cell(p, m, nfc[ind_father], ilevel, uold, u2);   /* :583-593 */
          for (int k2 = 0; k2 <= k2max; k2++)
            for (int j2 = 0; j2 <= j2max; j2++)
              for (int i2 = 0; i2 <= i2max; i2++) {
                int ind_son = i2 + 2 * j2 + 4 * k2;
                int ic = m->ncoarse + ind_son * m->ngridmax + igrid_nbor;
                int i3 = 1 + 2 * (i1 - 1) + i2, j3 = 1, k3 = 1;
                if (ndim > 1) j3 = 1 + 2 * (j1 - 1) + j2;
                if (ndim > 2) k3 = 1 + 2 * (k1 - 1) + k2;
                size_t x = PIX(w, i, i3, j3, k3);
                if (igrid_nbor > 0) {
                  for (int iv = 1; iv <= nvar; iv++) w->uloc[(iv - 1) * np + x] = UO(ic, iv);
                  w->ok[x] = m->son[ic] > 0; /* :661-663 */
                } else {
                  for (int iv = 1; iv <= nvar; iv++) w->uloc[(iv - 1) * np + x] = u2[ind_son * nvar + iv - 1];
                  w->ok[x] = 0;               /* :664-666 */
                }
                if (g_force)                  /* :637-647: straight injection of the father's f for buffer cells */
                  for (int idim = 1; idim <= ndim; idim++)
                    w->gloc[(idim - 1) * np + x] = igrid_nbor > 0 ? FO(ic, idim) : FO(nfc[ind_father], idim);
              }
        }
  }
  /* fluxes :681 */
  orc_unsplit(p, w, w->uloc, g_force ? w->gloc : NULL, w->flux, w->tmp, dx, dx, dx, dt, ncache);
  /* reset flux along direction at refined interface :720-747 */
  for (int idim = 0; idim < ndim; idim++) {
    int i0 = idim == 0, j0 = idim == 1, k0 = idim == 2;
    for (int k3 = k3min; k3 <= k3max + k0; k3++)
      for (int j3 = j3min; j3 <= j3max + j0; j3++)
        for (int i3 = i3min; i3 <= i3max + i0; i3++)
          for (int iv = 0; iv < nvar; iv++)
            for (int i = 0; i < ncache; i++)
              if (w->ok[PIX(w, i, i3 - i0, j3 - j0, k3 - k0)] || w->ok[PIX(w, i, i3, j3, k3)])
                w->flux[(iv + (size_t)idim * nvar) * nfp + FIX(w, i, i3, j3, k3)] = 0.0;
    if (g_divu)                                                     /* pressure_fix :737-745 */
      for (int k3 = k3min; k3 <= k3max + k0; k3++)
        for (int j3 = j3min; j3 <= j3max + j0; j3++)
          for (int i3 = i3min; i3 <= i3max + i0; i3++)
            for (int iv = 0; iv < 2; iv++)
              for (int i = 0; i < ncache; i++)
                if (w->ok[PIX(w, i, i3 - i0, j3 - j0, k3 - k0)] || w->ok[PIX(w, i, i3, j3, k3)])
                  w->tmp[(iv + (size_t)idim * 2) * nfp + FIX(w, i, i3, j3, k3)] = 0.0;
  }
  }
  if (phase == 1) return;
  /* conservative update at level ilevel :751-792 */
  for (int idim = 0; idim < ndim; idim++) {
    int i0 = idim == 0, j0 = idim == 1, k0 = idim == 2;
    for (int k2 = 0; k2 <= k2max; k2++)
      for (int j2 = 0; j2 <= j2max; j2++)
        for (int i2 = 0; i2 <= i2max; i2++) {
          int ind_son = i2 + 2 * j2 + 4 * k2;
          int iskip = m->ncoarse + ind_son * m->ngridmax;
          int i3 = 1 + i2, j3 = 1 + j2, k3 = 1 + k2;
          for (int iv = 1; iv <= nvar; iv++)
            for (int i = 0; i < ncache; i++) {
              int ic = iskip + ind_grid[i];
              const double* f = &w->flux[((iv - 1) + (size_t)idim * nvar) * nfp];
              UN(ic, iv) = UN(ic, iv) + (f[FIX(w, i, i3, j3, k3)] - f[FIX(w, i, i3 + i0, j3 + j0, k3 + k0)]);
            }
          if (g_divu) {                                             /* :773-786 */
            const double* t1 = &w->tmp[(0 + (size_t)idim * 2) * nfp];
            const double* t2 = &w->tmp[(1 + (size_t)idim * 2) * nfp];
            for (int i = 0; i < ncache; i++) {
              int ic = iskip + ind_grid[i];
              g_divu[ic - 1] = g_divu[ic - 1] + (t1[FIX(w, i, i3, j3, k3)] - t1[FIX(w, i, i3 + i0, j3 + j0, k3 + k0)]);
            }
            for (int i = 0; i < ncache; i++) {
              int ic = iskip + ind_grid[i];
              g_enew[ic - 1] = g_enew[ic - 1] + (t2[FIX(w, i, i3, j3, k3)] - t2[FIX(w, i, i3 + i0, j3 + j0, k3 + k0)]);
            }
          }
        }
  }
  /* conservative update at level ilevel-1 :798-908.  Loop order of the reference: variable, then face, then the
   * octs of the batch that sit at a coarse-fine boundary (several octs may reflux into the same coarse cell). */
  for (int idim = 0; idim < ndim; idim++) {
    int i0 = idim == 0, j0 = idim == 1, k0 = idim == 2;
    int nb_noneigh = 0;
    int* ind_buffer = (int*)malloc(sizeof(int) * (size_t)(ncache + 1));
    int* ind_cell = (int*)malloc(sizeof(int) * (size_t)(ncache + 1));
    for (int i = 0; i < ncache; i++) { /* left :811-817 */
      int nb = NBOR(m, ind_grid[i], 2 * idim + 1);
      if (m->son[nb] == 0) { ind_buffer[nb_noneigh] = nb; ind_cell[nb_noneigh] = i; nb_noneigh++; }
    }
    for (int iv = 1; iv <= nvar; iv++)
      for (int k3 = k3min; k3 <= k3max - k0; k3++)
        for (int j3 = j3min; j3 <= j3max - j0; j3++)
          for (int i3 = i3min; i3 <= i3max - i0; i3++)
            for (int i = 0; i < nb_noneigh; i++)
              UN(ind_buffer[i], iv) = UN(ind_buffer[i], iv) -
                  w->flux[((iv - 1) + (size_t)idim * nvar) * nfp + FIX(w, ind_cell[i], i3, j3, k3)] * oneontwotondim;
    if (g_divu)                                                     /* :830-851 */
      for (int tv = 0; tv < 2; tv++) {
        double* dst = tv == 0 ? g_divu : g_enew;
        for (int k3 = k3min; k3 <= k3max - k0; k3++)
          for (int j3 = j3min; j3 <= j3max - j0; j3++)
            for (int i3 = i3min; i3 <= i3max - i0; i3++)
              for (int i = 0; i < nb_noneigh; i++)
                dst[ind_buffer[i] - 1] = dst[ind_buffer[i] - 1] -
                    w->tmp[(tv + (size_t)idim * 2) * nfp + FIX(w, ind_cell[i], i3, j3, k3)] * oneontwotondim;
      }
    nb_noneigh = 0;
    for (int i = 0; i < ncache; i++) { /* right :863-869 */
      int nb = NBOR(m, ind_grid[i], 2 * idim + 2);
      if (m->son[nb] == 0) { ind_buffer[nb_noneigh] = nb; ind_cell[nb_noneigh] = i; nb_noneigh++; }
    }
    for (int iv = 1; iv <= nvar; iv++)
      for (int k3 = k3min + k0; k3 <= k3max; k3++)
        for (int j3 = j3min + j0; j3 <= j3max; j3++)
          for (int i3 = i3min + i0; i3 <= i3max; i3++)
            for (int i = 0; i < nb_noneigh; i++)
              UN(ind_buffer[i], iv) = UN(ind_buffer[i], iv) +
                  w->flux[((iv - 1) + (size_t)idim * nvar) * nfp + FIX(w, ind_cell[i], i3 + i0, j3 + j0, k3 + k0)] * oneontwotondim;
    if (g_divu)                                                     /* :882-903 */
      for (int tv = 0; tv < 2; tv++) {
        double* dst = tv == 0 ? g_divu : g_enew;
        for (int k3 = k3min + k0; k3 <= k3max; k3++)
          for (int j3 = j3min + j0; j3 <= j3max; j3++)
            for (int i3 = i3min + i0; i3 <= i3max; i3++)
              for (int i = 0; i < nb_noneigh; i++)
                dst[ind_buffer[i] - 1] = dst[ind_buffer[i] - 1] +
                    w->tmp[(tv + (size_t)idim * 2) * nfp + FIX(w, ind_cell[i], i3 + i0, j3 + j0, k3 + k0)] * oneontwotondim;
      }
    free(ind_buffer); free(ind_cell);
  }
}

/* godunov_fine hydro/godunov_fine.f90:5-35.  nthreads>1 splits the batches of a
 * level over OpenMP threads (the reference splits them over MPI ranks).          */
static int g_amr_threads = 1;
/* threads for the flux phase of orc_godunov_fine when it is called with nthreads=1 (the AMR drivers): batches are taken
 * g_amr_threads at a time, their fluxes computed concurrently (phase 1 only reads uold), and their updates of unew --
 * including the coarse refluxes -- applied serially in batch order, so the result is bit-identical to the serial routine */
void orc_set_amr_threads(int n) { g_amr_threads = n < 1 ? 1 : n; }

void orc_godunov_fine(const orc_params* p, const orc_mesh* m, int ilevel, double dt, const double* uold, double* unew,
                      int nthreads) {
  const int ncache = m->nactive[ilevel], nv = p->nvector;
  if (ilevel >= 0 && ilevel < 64) g_dt_level[ilevel] = dt;          /* dtnew(ilevel) for set_uold (pressure_fix) */
  if (ncache == 0) return;
  int nbatch = (ncache + nv - 1) / nv;
  if (nthreads < 1) nthreads = 1;
  if (g_divu && nthreads > 1) nthreads = 1;                          /* divu/enew refluxes need the serial update order */
  if (nthreads == 1 && g_amr_threads > 1 && nbatch > 1) {
    const int nt = IMIN(g_amr_threads, 64);
    static orc_work* ws[64];                       /* kept between calls (thousands of level steps per run) */
    static int ws_n = 0, ws_key[3] = {0, 0, 0};
    if (ws_n != nt || ws_key[0] != p->ndim || ws_key[1] != p->nvar || ws_key[2] != p->nvector) {
      for (int t = 0; t < ws_n; t++) orc_work_free(ws[t]);
      for (int t = 0; t < nt; t++) ws[t] = orc_work_new(p);
      ws_n = nt; ws_key[0] = p->ndim; ws_key[1] = p->nvar; ws_key[2] = p->nvector;
    }
    for (int b0 = 0; b0 < nbatch; b0 += nt) {
      const int nb = IMIN(nt, nbatch - b0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(nt)
#endif
      for (int t = 0; t < nb; t++) {
        const int ig = (b0 + t) * nv;
        godfine1(p, m, ws[t], m->active[ilevel] + ig, IMIN(nv, ncache - ig), ilevel, dt, uold, unew, 1);
      }
      for (int t = 0; t < nb; t++) {
        const int ig = (b0 + t) * nv;
        godfine1(p, m, ws[t], m->active[ilevel] + ig, IMIN(nv, ncache - ig), ilevel, dt, uold, unew, 2);
      }
    }
    return;
  }
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
  {
    orc_work* w = orc_work_new(p);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int b = 0; b < nbatch; b++) {
      int ig = b * nv;
      int ngrid = IMIN(nv, ncache - ig);
      godfine1(p, m, w, m->active[ilevel] + ig, ngrid, ilevel, dt, uold, unew, 0);
    }
    orc_work_free(w);
  }
}

/* courant_fine hydro/courant_fine.f90:1-159 (serial build: WITHOUTMPI) */
double orc_courant_fine(const orc_params* p, const orc_mesh* m, int ilevel, double dt_in, const double* uold, double sums[3]) {
  const int ndim = p->ndim, nvar = p->nvar, nv = p->nvector, twotondim = ipow2(ndim);
  const double dx = orc_dx(p, m, ilevel);
  double vol = 1;
  for (int d = 0; d < ndim; d++) vol *= dx; /* dx**ndim */
  double mass_loc = 0, ekin_loc = 0, eint_loc = 0, dt_loc = dt_in;
  const int ncache = m->nactive[ilevel];
  const int nbatch = (ncache + nv - 1) / nv;
#ifdef _OPENMP
#pragma omp parallel num_threads(g_nthreads) reduction(+ : mass_loc, ekin_loc, eint_loc) reduction(min : dt_loc) if (g_nthreads > 1)
#endif
  {
    double* uu = (double*)calloc((size_t)nv * nvar, 8);
    double* gg = (double*)calloc((size_t)nv * 3, 8);
    int* ind_leaf = (int*)calloc(nv, sizeof(int));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int b = 0; b < nbatch; b++) {
      const int igrid = b * nv;
      int ngrid = IMIN(nv, ncache - igrid);
      for (int ind = 0; ind < twotondim; ind++) {
        int iskip = m->ncoarse + ind * m->ngridmax;
        int nleaf = 0;
        for (int i = 0; i < ngrid; i++) {
          int ic = m->active[ilevel][igrid + i] + iskip;
          if (m->son[ic] == 0) ind_leaf[nleaf++] = ic;
        }
        for (int iv = 1; iv <= nvar; iv++)
          for (int i = 0; i < nleaf; i++) uu[i + (size_t)nv * (iv - 1)] = UO(ind_leaf[i], iv);
        for (int i = 0; i < nleaf; i++) mass_loc = mass_loc + uu[i] * vol;
        for (int i = 0; i < nleaf; i++) ekin_loc = ekin_loc + uu[i + (size_t)nv * (ndim + 1)] * vol;
        for (int i = 0; i < nleaf; i++) eint_loc = eint_loc + uu[i + (size_t)nv * (ndim + 1)] * vol;
        for (int iv = 1; iv <= ndim; iv++)
          for (int i = 0; i < nleaf; i++) {
            double mo = uu[i + (size_t)nv * iv];
            eint_loc = eint_loc - 0.5 * (mo * mo) / FMAX(uu[i], p->smallr) * vol;
          }
        if (nleaf > 0) {
          double dt_lev;
          if (g_force)                         /* courant_fine.f90:75-83 */
            for (int idim = 1; idim <= ndim; idim++)
              for (int i = 0; i < nleaf; i++) gg[i + (size_t)nv * (idim - 1)] = FO(ind_leaf[i], idim);
          orc_cmpdt(p, uu, g_force ? gg : NULL, dx, &dt_lev, nleaf);
          dt_loc = FMIN(dt_loc, dt_lev);
        }
      }
    }
    free(uu); free(gg); free(ind_leaf);
  }
  if (sums) { sums[0] += mass_loc; sums[1] += ekin_loc; sums[2] += eint_loc; }
  return FMIN(dt_in, dt_loc);
}

/* make_boundary_hydro hydro/hydro_boundary.f90:5-269 (reflexive and outflow) */
/* imposed boundaries (bound_type=3): boundary_var(ibound,1:nvar), the conservative state hydro/read_hydro_params.f90:440-468
 * builds from d_bound, u_bound, ... and the default boundana (hydro/boundana.f90) copies into every boundary cell          */
static double g_boundary_var[ORC_MAXBOUND][16];
void orc_set_boundary_var(int ibound, const double* var, int nvar) {
  for (int iv = 0; iv < nvar && iv < 16; iv++) g_boundary_var[ibound][iv] = var[iv];
}

void orc_make_boundary_hydro(const orc_params* p, const orc_mesh* m, int ilevel, double* uold) {
  const int ndim = p->ndim, nvar = p->nvar, twotondim = ipow2(ndim);
  static const int ref_x[8] = {2, 1, 4, 3, 6, 5, 8, 7}, ref_y[8] = {3, 4, 1, 2, 7, 8, 5, 6}, ref_z[8] = {5, 6, 7, 8, 1, 2, 3, 4};
  static const int free_[6][8] = {{1, 1, 3, 3, 5, 5, 7, 7}, {2, 2, 4, 4, 6, 6, 8, 8}, {1, 2, 1, 2, 5, 6, 5, 6},
                                  {3, 4, 3, 4, 7, 8, 7, 8}, {1, 2, 3, 4, 1, 2, 3, 4}, {5, 6, 7, 8, 5, 6, 7, 8}};
  for (int ib = 0; ib < m->nboundary; ib++) {
    int bt = m->boundary_type[ib];
    int boundary_dir = bt - 10 * (bt / 10);
    static const int inb[7] = {0, 2, 1, 4, 3, 6, 5};
    int inbor = inb[boundary_dir];
    const int* ind_ref;
    if (bt / 10 == 0) ind_ref = (boundary_dir <= 2) ? ref_x : (boundary_dir <= 4) ? ref_y : ref_z;
    else ind_ref = free_[boundary_dir - 1];
    double gs[3] = {1, 1, 1};
    if (bt == 1 || bt == 2) gs[0] = -1;
    if (bt == 3 || bt == 4) gs[1] = -1;
    if (bt == 5 || bt == 6) gs[2] = -1;
    for (int a = 0; a < m->nbound[ib][ilevel]; a++) {
      int ig = m->bound[ib][ilevel][a];
      int igr = m->son[NBOR(m, ig, inbor)];
      for (int ind = 0; ind < twotondim; ind++) {
        int ic = m->ncoarse + ind * m->ngridmax + ig;
        int icr = m->ncoarse + (ind_ref[ind] - 1) * m->ngridmax + igr;
        double uu[16];
        for (int iv = 1; iv <= nvar; iv++) uu[iv] = UO(icr, iv);
        if (bt / 10 == 0) { /* wall :141-157 */
          for (int iv = 1; iv <= nvar; iv++) {
            double sw = 1;
            if (iv > 1 && iv < ndim + 2) sw = gs[iv - 2];
            UO(ic, iv) = uu[iv] * sw;
          }
        } else if (bt / 10 == 1) { /* free :160-211, no_inflow=.false. */
          double ekin = 0.0, d = FMAX(uu[1], p->smallr);
          for (int idim = 1; idim <= ndim; idim++) { double v = uu[idim + 1] / d; ekin = ekin + 0.5 * d * (v * v); }
          uu[ndim + 2] = uu[ndim + 2] - ekin;
          for (int iv = 1; iv <= nvar; iv++) UO(ic, iv) = uu[iv];
          ekin = 0.0; d = FMAX(UO(ic, 1), p->smallr);
          for (int idim = 1; idim <= ndim; idim++) { double v = UO(ic, idim + 1) / d; ekin = ekin + 0.5 * d * (v * v); }
          UO(ic, ndim + 2) = UO(ic, ndim + 2) + ekin;
        } else { /* imposed :229-252 with the default boundana */
          for (int iv = 1; iv <= nvar; iv++) UO(ic, iv) = g_boundary_var[ib][iv - 1];
        }
      }
    }
  }
}

/* amr_step order for a single fully refined level (amr/amr_step.f90:326 newdt_fine ->
 * courant_fine, :333 set_unew, :388 godunov_fine, :423 set_uold, :514 make_boundary_hydro;
 * dtnew starts at boxlen/smallc, pm/newdt_fine.f90:47-51)                           */
void orc_run_uniform(const orc_params* p, const orc_mesh* m, int ilevel, int nstep, double* uold, double* unew,
                     double* dt_hist, double* t_io, int nthreads) {
  double t = t_io ? *t_io : 0.0;
  orc_set_threads(nthreads);
  orc_make_boundary_hydro(p, m, ilevel, uold);
  for (int s = 0; s < nstep; s++) {
    double sums[3] = {0, 0, 0};
    double dt = orc_courant_fine(p, m, ilevel, p->boxlen / p->smallc, uold, sums);
    orc_set_unew(p, m, ilevel, uold, unew);
    orc_godunov_fine(p, m, ilevel, dt, uold, unew, nthreads);
    orc_set_uold(p, m, ilevel, uold, unew);
    orc_make_boundary_hydro(p, m, ilevel, uold);
    t = t + dt;
    if (dt_hist) dt_hist[s] = dt;
  }
  if (t_io) *t_io = t;
}

/* upload_fine hydro/interpol_hydro.f90:5-68 + upl :73-263 (interpol_var=0) */
void orc_upload_fine(const orc_params* p, const orc_mesh* m, int ilevel, double* uold) {
  const int twotondim = ipow2(p->ndim), nvar = p->nvar;
  if (ilevel == m->nlevelmax) return;
  for (int a = 0; a < m->nactive[ilevel]; a++)
    for (int ind = 0; ind < twotondim; ind++) {
      const int ic = m->ncoarse + ind * m->ngridmax + m->active[ilevel][a];
      const int gs = m->son[ic];
      if (gs <= 0) continue;
      double getx = 0.0;
      for (int is = 0; is < twotondim; is++) getx = getx + FMAX(UO(m->ncoarse + is * m->ngridmax + gs, 1), p->smallr);
      UO(ic, 1) = getx / (double)twotondim;
      for (int iv = 2; iv <= nvar; iv++) {
        getx = 0.0;
        for (int is = 0; is < twotondim; is++) getx = getx + UO(m->ncoarse + is * m->ngridmax + gs, iv);
        UO(ic, iv) = getx / (double)twotondim;
      }
      if (g_interpol_var == 1 || g_interpol_var == 2) { /* average internal energy instead of total energy :204-261 */
        const int ndim = p->ndim;
        getx = 0.0;
        for (int is = 0; is < twotondim; is++) {
          const int cs = m->ncoarse + is * m->ngridmax + gs;
          double ekin = 0.0;
          for (int d = 1; d <= ndim; d++) ekin = ekin + 0.5 * (UO(cs, 1 + d) * UO(cs, 1 + d)) / FMAX(UO(cs, 1), p->smallr);
          getx = getx + UO(cs, ndim + 2) - ekin - 0.0;
        }
        double ekin = 0.0;
        for (int d = 1; d <= ndim; d++) ekin = ekin + 0.5 * (UO(ic, 1 + d) * UO(ic, 1 + d)) / FMAX(UO(ic, 1), p->smallr);
        UO(ic, ndim + 2) = getx / (double)twotondim + ekin + 0.0;
      }
    }
}

/* an empty mesh (coarse grid only) whose tree arrays the AMR driver (oracle/amr.py) fills in like
 * amr/refine_utils.f90 does; lists are pushed with orc_mesh_set_list                                 */
orc_mesh* orc_mesh_new(int ndim, const int bt[6], int ngridmax, int nlevelmax) {
  orc_mesh* m = (orc_mesh*)calloc(1, sizeof(orc_mesh));
  m->ndim = ndim; m->nlevelmax = nlevelmax;
  int nn[3] = {1, 1, 1}, cmin[3] = {0, 0, 0}, cmax[3] = {0, 0, 0};
  for (int d = 0; d < ndim; d++)
    for (int s2 = 0; s2 < 2; s2++)
      if (bt[2 * d + s2] > 0) {
        nn[d]++;
        if (s2 == 0) { cmin[d]++; cmax[d]++; }
        m->boundary_type[m->nboundary] = (bt[2 * d + s2] - 1) * 10 + (2 * d + s2 + 1);
        m->nboundary++;
      }
  m->nx = nn[0]; m->ny = nn[1]; m->nz = nn[2];
  m->icoarse_min = cmin[0]; m->icoarse_max = cmax[0]; m->jcoarse_min = cmin[1]; m->jcoarse_max = cmax[1];
  m->kcoarse_min = cmin[2]; m->kcoarse_max = cmax[2];
  m->ncoarse = nn[0] * nn[1] * nn[2];
  m->ngridmax = ngridmax;
  m->ncell = m->ncoarse + ipow2(ndim) * ngridmax;
  m->son = (int*)calloc((size_t)m->ncell + 1, sizeof(int));
  m->cpu_map = (int*)calloc((size_t)m->ncell + 1, sizeof(int));
  m->father = (int*)calloc((size_t)ngridmax + 1, sizeof(int));
  m->nbor = (int*)calloc((size_t)2 * ndim * (ngridmax + 1), sizeof(int));
  m->xg = (double*)calloc((size_t)ndim * (ngridmax + 1), sizeof(double));
  m->nactive = (int*)calloc(nlevelmax + 2, sizeof(int));
  m->active = (int**)calloc(nlevelmax + 2, sizeof(int*));
  m->nrecv = (int*)calloc(nlevelmax + 2, sizeof(int));
  m->recv = (int**)calloc(nlevelmax + 2, sizeof(int*));
  for (int b = 0; b < ORC_MAXBOUND; b++) {
    m->nbound[b] = (int*)calloc(nlevelmax + 2, sizeof(int));
    m->bound[b] = (int**)calloc(nlevelmax + 2, sizeof(int*));
  }
  return m;
}
/* kind 0: active(ilevel), kind 1: boundary(b,ilevel) */
void orc_mesh_set_list(orc_mesh* m, int kind, int b, int ilevel, int n, const int* igrid) {
  int** slot = kind == 0 ? &m->active[ilevel] : &m->bound[b][ilevel];
  int* cnt = kind == 0 ? &m->nactive[ilevel] : &m->nbound[b][ilevel];
  free(*slot);
  *slot = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  memcpy(*slot, igrid, sizeof(int) * (size_t)n);
  *cnt = n;
}
