// mhd_device.cuh -- per-face / per-edge device numerics of the ideal-MHD variant of the Godunov sweep
// (constrained transport, staggered face fields), FP64 on sm_100a.
//
// Semantics follow the reference routines cited at each function (mhd/godunov_utils.f90, mhd/umuscl.f90); every scalar
// expression keeps the reference's evaluation order, so that results are bit-identical to an IEEE, non-FMA evaluation:
// compile with -fmad=false.  State order handed to the 1-D solvers (mhd/umuscl.f90:1350-1367):
//   q[0]=rho q[1]=P q[2]=v_n q[3]=B_n q[4]=v_t1 q[5]=B_t1 q[6]=v_t2 q[7]=B_t2 ; f[0..8] (f[8] = internal-energy flux)
#pragma once
#include <cuda_runtime.h>
#include "hydro_device.cuh"
#include "real64.cuh"

namespace rgpu {

// iriemann / iriemann2d (hydro/read_hydro_params.f90:190-220)
enum { MHD_LLF = 0, MHD_ROE = 1, MHD_HLL = 2, MHD_HLLD = 3, MHD_UPWIND = 4, MHD_HYDRO = 5 };
enum { MHD2D_LLF = 0, MHD2D_ROE = 1, MHD2D_UPWIND = 2, MHD2D_HLL = 3, MHD2D_HLLA = 4, MHD2D_HLLD = 5 };

struct MPhys {
  double gamma, smallr, smallc, slope_theta, courant_factor;
  double smallp;      // smallr*smallc**2/gamma   mhd/umuscl.f90:2054
  int slope_type, slope_mag_type;
};

using real = R64;   // see real64.cuh: IEEE-identical, branch-free `/` and sqrt
template <class T> __device__ __forceinline__ T fmx4(T a, T b, T c, T d) { return fmx(fmx(fmx(a, b), c), d); }
template <class T> __device__ __forceinline__ T fmn4(T a, T b, T c, T d) { return fmn(fmn(fmn(a, b), c), d); }
template <class T> __device__ __forceinline__ T SQ(T a) { return a * a; }
#define zero 0.0
#define one 1.0
#define two 2.0
#define half 0.5
#define forth 0.25

__device__ __forceinline__ void find_mhd_flux(const MPhys& M, const real* q, real* c, real* ff) { /* :704 */
  real entho = one / (M.gamma - one);
  real d = q[0], P = q[1], u = q[2], A = q[3], v = q[4], B = q[5], w = q[6], C = q[7];
  real ecin = half * (u * u + v * v + w * w) * d;
  real emag = half * (A * A + B * B + C * C);
  real etot = P * entho + ecin + emag;
  real Ptot = P + emag;
  c[0] = d; c[1] = etot; c[2] = d * u; c[3] = A; c[4] = d * v; c[5] = B; c[6] = d * w; c[7] = C;
  c[8] = P * entho;
  ff[0] = d * u;
  ff[1] = (etot + Ptot) * u - A * (A * u + B * v + C * w);
  ff[2] = d * u * u + Ptot - A * A;
  ff[3] = zero;
  ff[4] = d * u * v - A * B;
  ff[5] = B * u - A * v;
  ff[6] = d * u * w - A * C;
  ff[7] = C * u - A * w;
  ff[8] = P * entho * u;
}

__device__ __forceinline__ real find_speed_fast(const MPhys& M, const real* q) { /* :823 */
  real d = q[0], P = q[1], A = q[3], B = q[5], C = q[7];
  const RcpD rd(d);
  real B2 = A * A + B * B + C * C;
  real c2 = M.gamma * P / rd;
  real d2 = half * (B2 / rd + c2);
  return rsqrt64(d2 + rsqrt64(d2 * d2 - c2 * A * A / rd));
}
__device__ __forceinline__ real find_speed_info(const MPhys& M, const real* q) { return find_speed_fast(M, q) + rabs(q[2]); } /* :787 */
__device__ __forceinline__ real find_speed_alfven(const real* q) { return rsqrt64(q[3] * q[3] / q[0]); }                                /* :857 */

__device__ __forceinline__ void mean_bn(real* ql, real* qr) { real bx = half * (ql[3] + qr[3]); ql[3] = bx; qr[3] = bx; }

__device__ __forceinline__ void lax_friedrich(const MPhys& M, real* ql, real* qr, real* fg, real zero_flux) { /* :352 */
  real ul[9], ur[9], fl[9], fr[9];
  mean_bn(ql, qr);
  find_mhd_flux(M, ql, ul, fl);
  find_mhd_flux(M, qr, ur, fr);
  real vl = find_speed_info(M, ql), vr = find_speed_info(M, qr);
  real vm = fmx(vl, vr);
  for (int n = 0; n < 9; n++) {
    real fmean = half * (fr[n] + fl[n]) * zero_flux;
    real udiff = half * (ur[n] - ul[n]);
    fg[n] = fmean - vm * udiff;
  }
}

__device__ __forceinline__ void upwind(const MPhys& M, real* ql, real* qr, real* fg, real zero_flux) { /* :313 */
  real ul[9], ur[9], fl[9], fr[9];
  mean_bn(ql, qr);
  find_mhd_flux(M, ql, ul, fl);
  find_mhd_flux(M, qr, ur, fr);
  real vleft = half * (ql[2] + qr[2]);
  for (int n = 0; n < 9; n++) {
    real fmean = half * (fr[n] + fl[n]) * zero_flux;
    real udiff = half * (ur[n] - ul[n]);
    fg[n] = fmean - rabs(vleft) * udiff;
  }
}

__device__ __forceinline__ void hll(const MPhys& M, real* ql, real* qr, real* fg) { /* :391 */
  real ul[9], ur[9], fl[9], fr[9];
  mean_bn(ql, qr);
  find_mhd_flux(M, ql, ul, fl);
  find_mhd_flux(M, qr, ur, fr);
  real cfl = find_speed_fast(M, ql), cfr = find_speed_fast(M, qr);
  real vl = ql[2], vr = qr[2];
  real SL = fmn(fmn(vl, vr) - fmx(cfl, cfr), zero);
  real SR = fmx(fmx(vl, vr) + fmx(cfl, cfr), zero);
  const RcpD rs(SR - SL);
#pragma unroll
  for (int n = 0; n < 9; n++) fg[n] = (SR * fl[n] - SL * fr[n] + SR * SL * (ur[n] - ul[n])) / rs;
}

__device__ __forceinline__ void hlld(const MPhys& M, real* ql, real* qr, real* fg) { /* :426 */
  real entho = one / (M.gamma - one);
  real A = half * (ql[3] + qr[3]);
  real sgnm = rsign(1.0, A);
  ql[3] = A; qr[3] = A;
  real rl = ql[0], Pl = ql[1], ul = ql[2], vl = ql[4], Bl = ql[5], wl = ql[6], Cl = ql[7];
  real ecinl = half * (ul * ul + vl * vl + wl * wl) * rl;
  real emagl = half * (A * A + Bl * Bl + Cl * Cl);
  real etotl = Pl * entho + ecinl + emagl;
  real Ptotl = Pl + emagl;
  real vdotBl = ul * A + vl * Bl + wl * Cl;
  real eintl = Pl * entho;
  real rr = qr[0], Pr = qr[1], ur = qr[2], vr = qr[4], Br = qr[5], wr = qr[6], Cr = qr[7];
  real ecinr = half * (ur * ur + vr * vr + wr * wr) * rr;
  real emagr = half * (A * A + Br * Br + Cr * Cr);
  real etotr = Pr * entho + ecinr + emagr;
  real Ptotr = Pr + emagr;
  real vdotBr = ur * A + vr * Br + wr * Cr;
  real eintr = Pr * entho;
  real cfastl = find_speed_fast(M, ql), cfastr = find_speed_fast(M, qr);
  real SL = fmn(ul, ur) - fmx(cfastl, cfastr);
  real SR = fmx(ul, ur) + fmx(cfastl, cfastr);
  real rcl = rl * (ul - SL), rcr = rr * (SR - ur);
  const RcpD rrc(rcr + rcl);
  real ustar = (rcr * ur + rcl * ul + (Ptotl - Ptotr)) / rrc;
  real Ptotstar = (rcr * Ptotl + rcl * Ptotr + rcl * rcr * (ul - ur)) / rrc;
  const RcpD rsl(SL - ustar), rsr(SR - ustar);
  /* left star region */
  real rstarl = rl * (SL - ul) / rsl;
  real estar = rl * (SL - ul) * (SL - ustar) - A * A;
  real el = rl * (SL - ul) * (SL - ul) - A * A;
  real eintstarl = eintl * (SL - ul) / rsl;
  real vstarl, Bstarl, wstarl, Cstarl;
  if (rabs(estar) < (real)1e-4f * (A * A)) /* `1e-4` is a default-real literal */ { vstarl = vl; Bstarl = Bl; wstarl = wl; Cstarl = Cl; }
  else {
    const RcpD re(estar);
    vstarl = vl - A * Bl * (ustar - ul) / re;
    Bstarl = Bl * el / re;
    wstarl = wl - A * Cl * (ustar - ul) / re;
    Cstarl = Cl * el / re;
  }
  real vdotBstarl = ustar * A + vstarl * Bstarl + wstarl * Cstarl;
  real etotstarl = ((SL - ul) * etotl - Ptotl * ul + Ptotstar * ustar + A * (vdotBl - vdotBstarl)) / rsl;
  real sqrrstarl = rsqrt64(rstarl);
  real calfvenl = rabs(A) / sqrrstarl;
  real SAL = ustar - calfvenl;
  /* right star region */
  real rstarr = rr * (SR - ur) / rsr;
  estar = rr * (SR - ur) * (SR - ustar) - A * A;
  real er = rr * (SR - ur) * (SR - ur) - A * A;
  real eintstarr = eintr * (SR - ur) / rsr;
  real vstarr, Bstarr, wstarr, Cstarr;
  if (rabs(estar) < (real)1e-4f * (A * A)) /* `1e-4` is a default-real literal */ { vstarr = vr; Bstarr = Br; wstarr = wr; Cstarr = Cr; }
  else {
    const RcpD re(estar);
    vstarr = vr - A * Br * (ustar - ur) / re;
    Bstarr = Br * er / re;
    wstarr = wr - A * Cr * (ustar - ur) / re;
    Cstarr = Cr * er / re;
  }
  real vdotBstarr = ustar * A + vstarr * Bstarr + wstarr * Cstarr;
  real etotstarr = ((SR - ur) * etotr - Ptotr * ur + Ptotstar * ustar + A * (vdotBr - vdotBstarr)) / rsr;
  real sqrrstarr = rsqrt64(rstarr);
  real calfvenr = rabs(A) / sqrrstarr;
  real SAR = ustar + calfvenr;
  /* real star region */
  const RcpD den(sqrrstarl + sqrrstarr);
  real vstarstar = (sqrrstarl * vstarl + sqrrstarr * vstarr + sgnm * (Bstarr - Bstarl)) / den;
  real wstarstar = (sqrrstarl * wstarl + sqrrstarr * wstarr + sgnm * (Cstarr - Cstarl)) / den;
  real Bstarstar = (sqrrstarl * Bstarr + sqrrstarr * Bstarl + sgnm * sqrrstarl * sqrrstarr * (vstarr - vstarl)) / den;
  real Cstarstar = (sqrrstarl * Cstarr + sqrrstarr * Cstarl + sgnm * sqrrstarl * sqrrstarr * (wstarr - wstarl)) / den;
  real vdotBstarstar = ustar * A + vstarstar * Bstarstar + wstarstar * Cstarstar;
  real etotstarstarl = etotstarl - sgnm * sqrrstarl * (vdotBstarl - vdotBstarstar);
  real etotstarstarr = etotstarr + sgnm * sqrrstarr * (vdotBstarr - vdotBstarstar);
  real ro, uo, vo, wo, Bo, Co, Ptoto, etoto, vdotBo, einto;
  if (SL > 0.0) { ro = rl; uo = ul; vo = vl; wo = wl; Bo = Bl; Co = Cl; Ptoto = Ptotl; etoto = etotl; vdotBo = vdotBl; einto = eintl; }
  else if (SAL > 0.0) { ro = rstarl; uo = ustar; vo = vstarl; wo = wstarl; Bo = Bstarl; Co = Cstarl; Ptoto = Ptotstar; etoto = etotstarl; vdotBo = vdotBstarl; einto = eintstarl; }
  else if (ustar > 0.0) { ro = rstarl; uo = ustar; vo = vstarstar; wo = wstarstar; Bo = Bstarstar; Co = Cstarstar; Ptoto = Ptotstar; etoto = etotstarstarl; vdotBo = vdotBstarstar; einto = eintstarl; }
  else if (SAR > 0.0) { ro = rstarr; uo = ustar; vo = vstarstar; wo = wstarstar; Bo = Bstarstar; Co = Cstarstar; Ptoto = Ptotstar; etoto = etotstarstarr; vdotBo = vdotBstarstar; einto = eintstarr; }
  else if (SR > 0.0) { ro = rstarr; uo = ustar; vo = vstarr; wo = wstarr; Bo = Bstarr; Co = Cstarr; Ptoto = Ptotstar; etoto = etotstarr; vdotBo = vdotBstarr; einto = eintstarr; }
  else { ro = rr; uo = ur; vo = vr; wo = wr; Bo = Br; Co = Cr; Ptoto = Ptotr; etoto = etotr; vdotBo = vdotBr; einto = eintr; }
  fg[0] = ro * uo;
  fg[1] = (etoto + Ptoto) * uo - A * vdotBo;
  fg[2] = ro * uo * uo + Ptoto - A * A;
  fg[3] = zero;
  fg[4] = ro * uo * vo - A * Bo;
  fg[5] = Bo * uo - A * vo;
  fg[6] = ro * uo * wo - A * Co;
  fg[7] = Co * uo - A * wo;
  fg[8] = uo * einto;
}

__device__ __forceinline__ void hydro_acoustic(const MPhys& M, real* ql, real* qr, real* fg) { /* :1092 */
  real smallp = M.smallr * (M.smallc * M.smallc);
  mean_bn(ql, qr);
  real rl = fmx(ql[0], M.smallr), rr = fmx(qr[0], M.smallr);
  real pl = fmx(ql[1], smallp), pr = fmx(qr[1], smallp);
  real ul = ql[2], ur = qr[2];
  real cl = rsqrt64(M.gamma * pl / rl), cr = rsqrt64(M.gamma * pr / rr);
  real wl = cl * rl, wr = cr * rr;
  real pstar = ((wr * pl + wl * pr) + wl * wr * (ul - ur)) / (wl + wr);
  real ustar = ((wr * ur + wl * ul) + (pl - pr)) / (wl + wr);
  real sgnm = rsign(1.0, ustar);
  real ro, uo, po, co;
  if (sgnm == one) { ro = rl; uo = ul; po = pl; co = cl; } else { ro = rr; uo = ur; po = pr; co = cr; }
  real rstar = ro + (pstar - po) / (co * co);
  rstar = fmx(rstar, M.smallr);
  real cstar = rsqrt64(rabs(M.gamma * pstar / rstar));
  cstar = fmx(cstar, M.smallc);
  real spout = co - sgnm * uo;
  real spin = cstar - sgnm * ustar;
  real ushock = half * (spin + spout);
  ushock = fmx(ushock, -sgnm * ustar);
  if (pstar >= po) { spout = ushock; spin = spout; }
  real qg[8];
  if (spout < zero) { qg[0] = ro; qg[1] = po; qg[2] = uo; }
  else if (spin >= zero) { qg[0] = rstar; qg[1] = pstar; qg[2] = ustar; }
  else {
    real frac = spout / (spout - spin);
    qg[0] = frac * rstar + (one - frac) * ro;
    qg[1] = frac * pstar + (one - frac) * po;
    qg[2] = frac * ustar + (one - frac) * uo;
  }
  for (int n = 3; n < 8; n++) qg[n] = (sgnm == one) ? ql[n] : qr[n];
  real ug[9];
  find_mhd_flux(M, qg, ug, fg);
}

__device__ __forceinline__ void eigenvalues(const MPhys& M, real d, real vx, real pr, real bx, real by, real bz, real* lambda) { /* :1207 */
  const RcpD rd(d);
  real btsq = by * by + bz * bz;
  real vaxsq = bx * bx / rd;
  real vax = rsqrt64(vaxsq);
  real asq = M.gamma * pr / rd;
  asq = fmx(asq, real(M.smallc * M.smallc));
  real astarsq = asq + vaxsq + btsq / rd;
  real disc = rsqrt64(astarsq * astarsq - 4.0 * asq * vaxsq);
  real cfsq = .5 * (astarsq + disc);
  real cfast = rsqrt64(cfsq);
  real cssq = .5 * (astarsq - disc);
  if (cssq <= 0.) cssq = 0.;
  real cslow = rsqrt64(cssq);
  lambda[0] = vx - cfast; lambda[1] = vx - vax; lambda[2] = vx - cslow; lambda[3] = vx;
  lambda[4] = vx + cslow; lambda[5] = vx + vax; lambda[6] = vx + cfast;
}

/* athena_roe :878 with eigen_cons :1266 written out in place.  Every entry of the eigenvector matrices keeps the reference's
 * expression; only the ORDER in which independent entries are evaluated differs: the left eigenvectors are formed column
 * by column and consumed at once by the wave strengths a(n) (:985-994), before the right eigenvectors are formed, which
 * halves the live register set (the 7x7 matrices never coexist).  Quotients with a common divisor share one reciprocal
 * (RcpD, real64.cuh; same bits).  fm[8] (the internal-energy flux) is never set by the reference on the Roe branch
 * (godunov_utils.f90:1069-1088); it is only read when pressure_fix=.true., which is out of scope: 0 here. */
__device__ __forceinline__ void athena_roe(const MPhys& M, real* ql, real* qr, real* fm, real zero_flux) {
  real ul_[9], ur_[9], fl[9], fr[9];
  real lambda[7], lambdal[7], lambdar[7], a[7];
  mean_bn(ql, qr);
  find_mhd_flux(M, ql, ul_, fl);
  find_mhd_flux(M, qr, ur_, fr);
  real dl = ql[0], dr = qr[0], pl = ql[1], pr = qr[1], vxl = ql[2], vxr = qr[2], vyl = ql[4], vyr = qr[4];
  real byl = ql[5], byr = qr[5], vzl = ql[6], vzr = qr[6], bzl = ql[7], bzr = qr[7];
  real bx = 0.5 * (ql[3] + qr[3]);
  real el = ul_[1], er = ur_[1], mxl = ul_[2], mxr = ur_[2], myl = ul_[4], myr = ur_[4], mzl = ul_[6], mzr = ur_[6];
  real pbl = half * (bx * bx + byl * byl + bzl * bzl);
  real pbr = half * (bx * bx + byr * byr + bzr * bzr);
  real hl = (el + pl + pbl) / dl;
  real hr = (er + pr + pbr) / dr;
  real sqrtdl = rsqrt64(dl), sqrtdr = rsqrt64(dr);
  real droe = sqrtdl * sqrtdr;
  const RcpD ssum(sqrtdl + sqrtdr);
  real vx = (sqrtdl * vxl + sqrtdr * vxr) / ssum;
  real vy = (sqrtdl * vyl + sqrtdr * vyr) / ssum;
  real vz = (sqrtdl * vzl + sqrtdr * vzr) / ssum;
  real by = (sqrtdr * byl + sqrtdl * byr) / ssum;
  real bz = (sqrtdr * bzl + sqrtdl * bzr) / ssum;
  real h = (sqrtdl * hl + sqrtdr * hr) / ssum;
  const RcpD droe2(2 * droe);
  real Xfac = ((by * by - byl * byr) + (bz * bz - bzl * bzr)) / droe2;
  real Yfac = (dl + dr) / droe2;
  real d = droe, Bx = bx;
  /* ---- eigen_cons(droe,vxroe,vyroe,vzroe,hroe,bx,byroe,bzroe,Xfactor,Yfactor,...) :1266-1330 ---- */
  const real gamma = M.gamma;
  const real hsq2 = 0.707106769084930419921875 /* the default-real (single precision) expression .5*sqrt(2.) */;
  const RcpD rd(d);
  real vsq = vx * vx + vy * vy + vz * vz;
  real btsq = by * by + bz * bz;
  real bt_starsq = (gamma - 1. - (gamma - 2.) * Yfac) * btsq;
  real bt = rsqrt64(btsq);
  real bt_star = rsqrt64(bt_starsq);
  real vaxsq = Bx * Bx / rd;
  real vax = rsqrt64(vaxsq);
  real hp = h - (vaxsq + btsq / rd);
  real twid_asq = ((gamma - 1.) * (hp - .5 * vsq) - (gamma - 2.) * Xfac);
  twid_asq = fmx(twid_asq, real(M.smallc * M.smallc));
  real q_starsq = twid_asq + (vaxsq + bt_starsq / rd);
  real disc = rsqrt64(q_starsq * q_starsq - 4.0 * twid_asq * vaxsq);
  real cfsq = .5 * (q_starsq + disc);
  real cfast = rsqrt64(cfsq);
  real cssq = .5 * (q_starsq - disc);
  if (cssq <= 0.) cssq = 0.;
  real cslow = rsqrt64(cssq);
  real beta_y, beta_z, beta_ystar, beta_zstar;
  if (bt == 0) { beta_y = hsq2; beta_z = hsq2; beta_ystar = hsq2; beta_zstar = hsq2; }
  else {
    const RcpD rbt(bt), rbts(bt_star);
    beta_y = by / rbt; beta_z = bz / rbt; beta_ystar = by / rbts; beta_zstar = bz / rbts;
  }
  real beta_starsq = beta_ystar * beta_ystar + beta_zstar * beta_zstar;
  real vbeta = vy * beta_ystar + vz * beta_zstar;
  real alpha_f, alpha_s;
  if ((cfsq - cssq) == 0.) { alpha_f = 1.0; alpha_s = 0.0; }
  else if ((twid_asq - cssq) <= 0.) { alpha_f = 0.0; alpha_s = 1.0; }
  else if ((cfsq - twid_asq) <= 0.) { alpha_f = 1.0; alpha_s = 0.0; }
  else {
    const RcpD rcc(cfsq - cssq);
    alpha_f = rsqrt64((twid_asq - cssq) / rcc); alpha_s = rsqrt64((cfsq - twid_asq) / rcc);
  }
  real droot = rsqrt64(d);
  const RcpD rdroot(droot);
  real s = rsign(1.0, Bx);
  real twid_a = rsqrt64(twid_asq);
  real Qfast = s * cfast * alpha_f;
  real Qslow = s * cslow * alpha_s;
  real af_prime = twid_a * alpha_f / rdroot;
  real as_prime = twid_a * alpha_s / rdroot;
  real Afpbb = af_prime * bt_star * beta_starsq;
  real Aspbb = as_prime * bt_star * beta_starsq;
  lambda[0] = vx - cfast; lambda[1] = vx - vax; lambda[2] = vx - cslow; lambda[3] = vx;
  lambda[4] = vx + cslow; lambda[5] = vx + vax; lambda[6] = vx + cfast;
  /* ---- left eigenvectors :1390-1470, column by column, and the wave strengths a(n) :985-994 ---- */
  {
    real na = 0.5 / twid_asq;
    real cff = na * alpha_f * cfast;
    real css = na * alpha_s * cslow;
    real Qfl = Qfast * na;
    real Qsl = Qslow * na;
    real af = na * af_prime * d;
    real as = na * as_prime * d;
    real Afpb = na * af_prime * bt_star;
    real Aspb = na * as_prime * bt_star;
    real alf = (gamma - 1.) * na * alpha_f;
    real als = (gamma - 1.) * na * alpha_s;
    const RcpD rbss(beta_starsq);
    real Q_ystar = beta_ystar / rbss;
    real Q_zstar = beta_zstar / rbss;
    real vqstr = (vy * Q_ystar + vz * Q_zstar);
    real norm = (gamma - 1.) * 2. * na;
    const real du0 = dr - dl, du1 = mxr - mxl, du2 = myr - myl, du3 = mzr - mzl, du4 = er - el, du5 = byr - byl, du6 = bzr - bzl;
#define WAVE(n, L0, L1, L2, L3, L4, L5, L6)                                                                  \
  a[n] = 0.0; a[n] = a[n] + du0 * (L0); a[n] = a[n] + du1 * (L1); a[n] = a[n] + du2 * (L2);                  \
  a[n] = a[n] + du3 * (L3); a[n] = a[n] + du4 * (L4); a[n] = a[n] + du5 * (L5); a[n] = a[n] + du6 * (L6);
    const real lem61 = as * Q_ystar - alf * by, lem71 = as * Q_zstar - alf * bz;
    WAVE(0, alf * (vsq - hp) + cff * (cfast + vx) - Qsl * vqstr - Aspb, -alf * vx - cff, -alf * vy + Qsl * Q_ystar, -alf * vz + Qsl * Q_zstar,
         alf, lem61, lem71)
    const real lem12 = 0.5 * (vy * beta_z - vz * beta_y), lem32 = -0.5 * beta_z, lem42 = 0.5 * beta_y;
    const real lem62 = -0.5 * droot * beta_z * s, lem72 = 0.5 * droot * beta_y * s;
    WAVE(1, lem12, real(0.), lem32, lem42, real(0.), lem62, lem72)
    const real lem63 = -af * Q_ystar - als * by, lem73 = -af * Q_zstar - als * bz;
    WAVE(2, als * (vsq - hp) + css * (cslow + vx) + Qfl * vqstr + Afpb, -als * vx - css, -als * vy - Qfl * Q_ystar, -als * vz - Qfl * Q_zstar,
         als, lem63, lem73)
    WAVE(3, 1. - norm * (.5 * vsq - (gamma - 2.) * Xfac / (gamma - 1.)), norm * vx, norm * vy, norm * vz, -norm, norm * by, norm * bz)
    WAVE(4, als * (vsq - hp) + css * (cslow - vx) - Qfl * vqstr + Afpb, -als * vx + css, -als * vy + Qfl * Q_ystar, -als * vz + Qfl * Q_zstar,
         als, lem63, lem73)
    WAVE(5, -lem12, real(0.), -lem32, -lem42, real(0.), lem62, lem72)
    WAVE(6, alf * (vsq - hp) + cff * (cfast - vx) + Qsl * vqstr - Aspb, -alf * vx + cff, -alf * vy - Qsl * Q_ystar, -alf * vz - Qsl * Q_zstar,
         alf, lem61, lem71)
#undef WAVE
  }
  /* ---- right eigenvectors :1332-1388, ROW BY ROW: row n is formed, used by the intermediate-state check (:996-1033) and by the
   * flux assembly (:1046-1066), and dropped -- the 7x7 matrix never exists in registers.  The two accumulations keep the
   * reference's order (n = 1..7 inside each sum); the flux is assembled speculatively and replaced by the LLF flux when one of
   * the intermediate states is unphysical, exactly like the early return of the reference.  The physical fluxes fl, fr and the
   * conservative states are recomputed here (find_mhd_flux is deterministic: same bits) instead of being held since the top. */
  eigenvalues(M, dl, vxl, pl, bx, byl, bzl, lambdal);
  eigenvalues(M, dr, vxr, pr, bx, byr, bzr, lambdar);
#pragma unroll
  for (int n = 0; n < 7; n += 2) {
    real l1 = fmn(lambdal[n], lambda[n]);
    real l2 = fmx(lambdar[n], lambda[n]);
    if (l1 < zero && l2 > zero) lambda[n] = (lambda[n] * (l2 + l1) - two * l2 * l1) / (l2 - l1);
  }
  find_mhd_flux(M, ql, ul_, fl);
  find_mhd_flux(M, qr, ur_, fr);
  bool llf = false;
  real dim = dl, mxm = mxl, mym = myl, mzm = mzl, eim = el, bym = byl, bzm = bzl;
  real fluxd = fl[0] * zero_flux + fr[0] * zero_flux, fluxe = fl[1] * zero_flux + fr[1] * zero_flux;
  real fluxmx = fl[2] * zero_flux + fr[2] * zero_flux, fluxmy = fl[4] * zero_flux + fr[4] * zero_flux;
  real fluxby = fl[5] * zero_flux + fr[5] * zero_flux, fluxmz = fl[6] * zero_flux + fr[6] * zero_flux;
  real fluxbz = fl[7] * zero_flux + fr[7] * zero_flux;
  const real r15 = -s * beta_z / rdroot, r16 = s * beta_y / rdroot;          /* rem(2,6), rem(2,7) = rem(6,6), rem(6,7) */
  const real r14 = -(vy * beta_z - vz * beta_y);                             /* rem(2,5) */
  const real r05 = as_prime * beta_ystar, r06 = as_prime * beta_zstar;      /* rem(1,6:7) = rem(7,6:7) */
  const real r25 = -af_prime * beta_ystar, r26 = -af_prime * beta_zstar;    /* rem(3,6:7) = rem(5,6:7) */
#define ROW(n, R0, R1, R2, R3, R4, R5, R6)                                                                         \
  {                                                                                                                \
    const real c0 = (R0), c1 = (R1), c2 = (R2), c3 = (R3), c4 = (R4), c5 = (R5), c6 = (R6);                        \
    dim = dim + a[n] * c0; mxm = mxm + a[n] * c1; mym = mym + a[n] * c2; mzm = mzm + a[n] * c3;                    \
    eim = eim + a[n] * c4; bym = bym + a[n] * c5; bzm = bzm + a[n] * c6;                                           \
    real etm = eim - 0.5 * (mxm * mxm + mym * mym + mzm * mzm) / dim - 0.5 * (bx * bx + bym * bym + bzm * bzm);    \
    if (dim <= zero || etm <= zero) llf = true;                                                                    \
    real coef = rabs(lambda[n]) * a[n];                                                                            \
    fluxd = fluxd - coef * c0; fluxe = fluxe - coef * c4; fluxmx = fluxmx - coef * c1; fluxmy = fluxmy - coef * c2; \
    fluxby = fluxby - coef * c5; fluxmz = fluxmz - coef * c3; fluxbz = fluxbz - coef * c6;                         \
  }
  ROW(0, alpha_f, alpha_f * (vx - cfast), alpha_f * vy + Qslow * beta_ystar, alpha_f * vz + Qslow * beta_zstar,
      alpha_f * (hp - vx * cfast) + Qslow * vbeta + Aspbb, r05, r06)
  ROW(1, real(0.), real(0.), -beta_z, beta_y, r14, r15, r16)
  ROW(2, alpha_s, alpha_s * (vx - cslow), alpha_s * vy - Qfast * beta_ystar, alpha_s * vz - Qfast * beta_zstar,
      alpha_s * (hp - vx * cslow) - Qfast * vbeta - Afpbb, r25, r26)
  ROW(3, real(1.0), vx, vy, vz, 0.5 * vsq + (gamma - 2.) * Xfac / (gamma - 1.), real(0.), real(0.))
  ROW(4, alpha_s, alpha_s * (vx + cslow), alpha_s * vy + Qfast * beta_ystar, alpha_s * vz + Qfast * beta_zstar,
      alpha_s * (hp + vx * cslow) + Qfast * vbeta - Afpbb, r25, r26)
  ROW(5, real(0.), real(0.), beta_z, -beta_y, -r14, r15, r16)
  ROW(6, alpha_f, alpha_f * (vx + cfast), alpha_f * vy - Qslow * beta_ystar, alpha_f * vz - Qslow * beta_zstar,
      alpha_f * (hp + vx * cfast) - Qslow * vbeta + Aspbb, r05, r06)
#undef ROW
  if (llf) {   /* intermediate state unphysical: LLF flux :1018-1031 (rare: states and fluxes are formed again, same bits) */
    find_mhd_flux(M, ql, ul_, fl);
    find_mhd_flux(M, qr, ur_, fr);
    real vl = find_speed_info(M, ql), vr = find_speed_info(M, qr);
    real vm = fmx(vl, vr);
#pragma unroll
    for (int n = 0; n < 9; n++) {
      real fmean = half * (fr[n] + fl[n]) * zero_flux;
      real udiff = half * (ur_[n] - ul_[n]);
      fm[n] = fmean - vm * udiff;
    }
    return;
  }
  fm[0] = half * fluxd; fm[1] = half * fluxe; fm[2] = half * fluxmx; fm[3] = zero; fm[4] = half * fluxmy;
  fm[5] = half * fluxby; fm[6] = half * fluxmz; fm[7] = half * fluxbz; fm[8] = zero;
}

/* ======================================================================================================
 * 2-D Riemann problem at one cell edge: cmp_mag_flx mhd/umuscl.f90:1453.  The four corner states arrive already
 * permuted: s[0]=rho s[1]=P s[2]=v_p1 s[3]=v_p2 s[4]=v_or s[5]=B_p1 s[6]=B_p2 s[7]=B_or  (B_p1, B_p2 already replaced
 * by the pair means :1517-1528).
 * ====================================================================================================== */
__device__ __forceinline__ real fast_xy(const MPhys& M, const real* s, int y) {
  real qt[8];
  qt[0] = s[0]; qt[1] = s[1]; qt[6] = s[4]; qt[7] = s[7];
  if (!y) { qt[2] = s[2]; qt[3] = s[5]; qt[4] = s[3]; qt[5] = s[6]; }
  else { qt[2] = s[3]; qt[3] = s[6]; qt[4] = s[2]; qt[5] = s[5]; }
  return find_speed_fast(M, qt);
}
__device__ __forceinline__ real alfven_xy(const real* s, int y) {
  real qt[8];
  qt[0] = s[0]; qt[3] = y ? s[6] : s[5];
  return find_speed_alfven(qt);
}

template <int R2D>
__device__ __forceinline__ real emf_edge(const MPhys& M, const real* qLL, const real* qRL, const real* qLR, const real* qRR) {
  real ELL = qLL[2] * qLL[6] - qLL[3] * qLL[5];
  real ERL = qRL[2] * qRL[6] - qRL[3] * qRL[5];
  real ELR = qLR[2] * qLR[6] - qLR[3] * qLR[5];
  real ERR = qRR[2] * qRR[6] - qRR[3] * qRR[5];
  constexpr int r2d = R2D;
  if (r2d == MHD2D_HLLD) { /* :1567-1735 */
    real rLL = qLL[0], pLL = qLL[1], uLL = qLL[2], vLL = qLL[3], ALL = qLL[5], BLL = qLL[6], CLL = qLL[7];
    real rLR = qLR[0], pLR = qLR[1], uLR = qLR[2], vLR = qLR[3], ALR = qLR[5], BLR = qLR[6], CLR = qLR[7];
    real rRL = qRL[0], pRL = qRL[1], uRL = qRL[2], vRL = qRL[3], ARL = qRL[5], BRL = qRL[6], CRL = qRL[7];
    real rRR = qRR[0], pRR = qRR[1], uRR = qRR[2], vRR = qRR[3], ARR = qRR[5], BRR = qRR[6], CRR = qRR[7];
    real cfastLLx = fast_xy(M, qLL, 0), cfastLRx = fast_xy(M, qLR, 0), cfastRLx = fast_xy(M, qRL, 0), cfastRRx = fast_xy(M, qRR, 0);
    real cfastLLy = fast_xy(M, qLL, 1), cfastLRy = fast_xy(M, qLR, 1), cfastRLy = fast_xy(M, qRL, 1), cfastRRy = fast_xy(M, qRR, 1);
    real cmx = fmx4(cfastLLx, cfastLRx, cfastRLx, cfastRRx), cmy = fmx4(cfastLLy, cfastLRy, cfastRLy, cfastRRy);
    real SL = fmn4(uLL, uLR, uRL, uRR) - cmx;
    real SR = fmx4(uLL, uLR, uRL, uRR) + cmx;
    real SB = fmn4(vLL, vLR, vRL, vRR) - cmy;
    real ST = fmx4(vLL, vLR, vRL, vRR) + cmy;
    ELL = uLL * BLL - vLL * ALL;
    ELR = uLR * BLR - vLR * ALR;
    ERL = uRL * BRL - vRL * ARL;
    ERR = uRR * BRR - vRR * ARR;
    real PtotLL = pLL + half * (ALL * ALL + BLL * BLL + CLL * CLL);
    real PtotLR = pLR + half * (ALR * ALR + BLR * BLR + CLR * CLR);
    real PtotRL = pRL + half * (ARL * ARL + BRL * BRL + CRL * CRL);
    real PtotRR = pRR + half * (ARR * ARR + BRR * BRR + CRR * CRR);
    real rcLLx = rLL * (uLL - SL), rcRLx = rRL * (SR - uRL);
    real rcLRx = rLR * (uLR - SL), rcRRx = rRR * (SR - uRR);
    real rcLLy = rLL * (vLL - SB), rcLRy = rLR * (ST - vLR);
    real rcRLy = rRL * (vRL - SB), rcRRy = rRR * (ST - vRR);
    real ustar = (rcLLx * uLL + rcLRx * uLR + rcRLx * uRL + rcRRx * uRR + (PtotLL - PtotRL + PtotLR - PtotRR)) / (rcLLx + rcLRx + rcRLx + rcRRx);
    real vstar = (rcLLy * vLL + rcLRy * vLR + rcRLy * vRL + rcRRy * vRR + (PtotLL - PtotLR + PtotRL - PtotRR)) / (rcLLy + rcLRy + rcRLy + rcRRy);
    const RcpD dSL(SL - ustar), dSR(SR - ustar), dSB(SB - vstar), dST(ST - vstar);
    real rstarLLx = rLL * (SL - uLL) / dSL, BstarLL = BLL * (SL - uLL) / dSL;
    real rstarLLy = rLL * (SB - vLL) / dSB, AstarLL = ALL * (SB - vLL) / dSB;
    real rstarLL = rLL * (SL - uLL) / dSL * (SB - vLL) / dSB;
    real EstarLLx = ustar * BstarLL - vLL * ALL;
    real EstarLLy = uLL * BLL - vstar * AstarLL;
    real EstarLL = ustar * BstarLL - vstar * AstarLL;
    real rstarLRx = rLR * (SL - uLR) / dSL, BstarLR = BLR * (SL - uLR) / dSL;
    real rstarLRy = rLR * (ST - vLR) / dST, AstarLR = ALR * (ST - vLR) / dST;
    real rstarLR = rLR * (SL - uLR) / dSL * (ST - vLR) / dST;
    real EstarLRx = ustar * BstarLR - vLR * ALR;
    real EstarLRy = uLR * BLR - vstar * AstarLR;
    real EstarLR = ustar * BstarLR - vstar * AstarLR;
    real rstarRLx = rRL * (SR - uRL) / dSR, BstarRL = BRL * (SR - uRL) / dSR;
    real rstarRLy = rRL * (SB - vRL) / dSB, AstarRL = ARL * (SB - vRL) / dSB;
    real rstarRL = rRL * (SR - uRL) / dSR * (SB - vRL) / dSB;
    real EstarRLx = ustar * BstarRL - vRL * ARL;
    real EstarRLy = uRL * BRL - vstar * AstarRL;
    real EstarRL = ustar * BstarRL - vstar * AstarRL;
    real rstarRRx = rRR * (SR - uRR) / dSR, BstarRR = BRR * (SR - uRR) / dSR;
    real rstarRRy = rRR * (ST - vRR) / dST, AstarRR = ARR * (ST - vRR) / dST;
    real rstarRR = rRR * (SR - uRR) / dSR * (ST - vRR) / dST;
    real EstarRRx = ustar * BstarRR - vRR * ARR;
    real EstarRRy = uRR * BRR - vstar * AstarRR;
    real EstarRR = ustar * BstarRR - vstar * AstarRR;
    real sc = M.smallc;
    real calfvenL = fmx(fmx4(rabs(ALR) / rsqrt64(rstarLRx), rabs(AstarLR) / rsqrt64(rstarLR), rabs(ALL) / rsqrt64(rstarLLx), rabs(AstarLL) / rsqrt64(rstarLL)), sc);
    real calfvenR = fmx(fmx4(rabs(ARR) / rsqrt64(rstarRRx), rabs(AstarRR) / rsqrt64(rstarRR), rabs(ARL) / rsqrt64(rstarRLx), rabs(AstarRL) / rsqrt64(rstarRL)), sc);
    real calfvenB = fmx(fmx4(rabs(BLL) / rsqrt64(rstarLLy), rabs(BstarLL) / rsqrt64(rstarLL), rabs(BRL) / rsqrt64(rstarRLy), rabs(BstarRL) / rsqrt64(rstarRL)), sc);
    real calfvenT = fmx(fmx4(rabs(BLR) / rsqrt64(rstarLRy), rabs(BstarLR) / rsqrt64(rstarLR), rabs(BRR) / rsqrt64(rstarRRy), rabs(BstarRR) / rsqrt64(rstarRR)), sc);
    real SAL = fmn(ustar - calfvenL, zero), SAR = fmx(ustar + calfvenR, zero);
    real SAB = fmn(vstar - calfvenB, zero), SAT = fmx(vstar + calfvenT, zero);
    const RcpD dA(SAR - SAL), dB(SAT - SAB);
    real AstarT = (SAR * AstarRR - SAL * AstarLR) / dA, AstarB = (SAR * AstarRL - SAL * AstarLL) / dA;
    real BstarR = (SAT * BstarRR - SAB * BstarRL) / dB, BstarL = (SAT * BstarLR - SAB * BstarLL) / dB;
    real E;
    if (SB > 0.0) {
      if (SL > 0.0) E = ELL;
      else if (SR < 0.0) E = ERL;
      else E = (SAR * EstarLLx - SAL * EstarRLx + SAR * SAL * (BRL - BLL)) / dA;
    } else if (ST < 0.0) {
      if (SL > 0.0) E = ELR;
      else if (SR < 0.0) E = ERR;
      else E = (SAR * EstarLRx - SAL * EstarRRx + SAR * SAL * (BRR - BLR)) / dA;
    } else if (SL > 0.0) E = (SAT * EstarLLy - SAB * EstarLRy - SAT * SAB * (ALR - ALL)) / dB;
    else if (SR < 0.0) E = (SAT * EstarRLy - SAB * EstarRRy - SAT * SAB * (ARR - ARL)) / dB;
    else
      E = (SAL * SAB * EstarRR - SAL * SAT * EstarRL - SAR * SAB * EstarLR + SAR * SAT * EstarLL) / dA / dB -
          SAT * SAB / dB * (AstarT - AstarB) + SAR * SAL / dA * (BstarR - BstarL);
    return E;
  }
  if (r2d == MHD2D_HLL || r2d == MHD2D_HLLA) { /* :1737-1850 */
    real cLLx, cLRx, cRLx, cRRx, cLLy, cLRy, cRLy, cRRy;
    if (r2d == MHD2D_HLL) {
      cLLx = fast_xy(M, qLL, 0); cLRx = fast_xy(M, qLR, 0); cRLx = fast_xy(M, qRL, 0); cRRx = fast_xy(M, qRR, 0);
      cLLy = fast_xy(M, qLL, 1); cLRy = fast_xy(M, qLR, 1); cRLy = fast_xy(M, qRL, 1); cRRy = fast_xy(M, qRR, 1);
    } else {
      cLLx = alfven_xy(qLL, 0); cLRx = alfven_xy(qLR, 0); cRLx = alfven_xy(qRL, 0); cRRx = alfven_xy(qRR, 0);
      cLLy = alfven_xy(qLL, 1); cLRy = alfven_xy(qLR, 1); cRLy = alfven_xy(qRL, 1); cRRy = alfven_xy(qRR, 1);
    }
    real SL = fmn(fmn4(qLL[2], qLR[2], qRL[2], qRR[2]) - fmx4(cLLx, cLRx, cRLx, cRRx), zero);
    real SR = fmx(fmx4(qLL[2], qLR[2], qRL[2], qRR[2]) + fmx4(cLLx, cLRx, cRLx, cRRx), zero);
    real SB = fmn(fmn4(qLL[3], qLR[3], qRL[3], qRR[3]) - fmx4(cLLy, cLRy, cRLy, cRRy), zero);
    real ST = fmx(fmx4(qLL[3], qLR[3], qRL[3], qRR[3]) + fmx4(cLLy, cLRy, cRLy, cRRy), zero);
    return (SL * SB * ERR - SL * ST * ERL - SR * SB * ELR + SR * ST * ELL) / (SR - SL) / (ST - SB) -
           ST * SB / (ST - SB) * (qRR[5] - qLL[5]) + SR * SL / (SR - SL) * (qRR[6] - qLL[6]);
  }
  /* llf / roe / upwind: two 1-D problems on pair-averaged states :1852-1925 */
  real E = forth * (ELL + ERL + ELR + ERR);
  real ql[8], qr[8], fmean_x[9], fmean_y[9];
  const int mapx[8] = {0, 1, 2, 5, 3, 6, 4, 7}; /* qleft(1..8) <- s(1,2,3,6,4,7,5,8) */
  const int mapy[8] = {0, 1, 3, 6, 2, 5, 4, 7}; /* qleft(1..8) <- s(1,2,4,7,3,6,5,8) */
  for (int n = 0; n < 8; n++) {
    ql[n] = half * (qLL[mapx[n]] + qLR[mapx[n]]);
    qr[n] = half * (qRR[mapx[n]] + qRL[mapx[n]]);
  }
  if (r2d == MHD2D_ROE) athena_roe(M, ql, qr, fmean_x, 0.0);
  else if (r2d == MHD2D_LLF) lax_friedrich(M, ql, qr, fmean_x, 0.0);
  else upwind(M, ql, qr, fmean_x, 0.0);
  for (int n = 0; n < 8; n++) {
    ql[n] = half * (qLL[mapy[n]] + qRL[mapy[n]]);
    qr[n] = half * (qRR[mapy[n]] + qLR[mapy[n]]);
  }
  if (r2d == MHD2D_ROE) athena_roe(M, ql, qr, fmean_y, 0.0);
  else if (r2d == MHD2D_LLF) lax_friedrich(M, ql, qr, fmean_y, 0.0);
  else upwind(M, ql, qr, fmean_y, 0.0);
  return E + (fmean_x[5] - fmean_y[5]);
}


// 1-D solver dispatch of cmpflxm (mhd/umuscl.f90:1411-1437), allow_switch_solver=.false.
template <int R1D>
__device__ __forceinline__ void riemann1d(const MPhys& M, real* ql, real* qr, real* fg) {
  if (R1D == MHD_ROE) athena_roe(M, ql, qr, fg, 1.0);
  else if (R1D == MHD_LLF || R1D == MHD_UPWIND) lax_friedrich(M, ql, qr, fg, 1.0);   // CASE (4) also calls lax_friedrich
  else if (R1D == MHD_HLL) hll(M, ql, qr, fg);
  else if (R1D == MHD_HLLD) hlld(M, ql, qr, fg);
  else hydro_acoustic(M, ql, qr, fg);
}

// minmod / moncen limiter in the algebraic form of the MHD build (mhd/umuscl.f90:2367-2376)
__device__ __forceinline__ real slope_mm(real st, real ql, real qc, real qr) {
  const real dlft = st * (qc - ql);
  const real drgt = st * (qr - qc);
  const real dcen = half * (dlft + drgt) * real(1.0 / st.v);   // st is 1 or 2: x/st == x*(1/st) exactly
  const real dsgn = rsign(1.0, dcen);
  real dlim = fmn(rabs(dlft), rabs(drgt));
  if ((dlft * drgt) <= zero) dlim = zero;
  return dsgn * fmn(dlim, rabs(dcen));
}

// cmpdt for one cell (mhd/godunov_utils.f90:5-111, no gravity); uu[11] is destroyed like in the reference
__device__ __forceinline__ real mhd_cmpdt_cell(const MPhys& M, real* uu, real dx) {
  uu[0] = fmx(uu[0], real(M.smallr));
  const real rho = uu[0];
  const RcpD rrho(rho);
  for (int d = 1; d <= 3; d++) uu[d] = uu[d] / rrho;
  real B2 = zero;
  for (int d = 1; d <= 3; d++) {
    const real Bc = half * (uu[4 + d] + uu[7 + d]);
    B2 = B2 + Bc * Bc;
    uu[4] = uu[4] - half * uu[0] * (uu[d] * uu[d]) - half * (Bc * Bc);
  }
  uu[4] = fmx((M.gamma - one) * uu[4], M.smallp);
  const real a2 = M.gamma * uu[4] / rrho;
  real ctot = zero;
  for (int d = 1; d <= 3; d++) {
    const real cc = half * (B2 / rrho + a2);
    const real BN = half * (uu[4 + d] + uu[7 + d]);
    const real cf = rsqrt64(cc + rsqrt64(cc * cc - a2 * (BN * BN) / rrho));
    ctot = ctot + rabs(uu[d]) + cf;
  }
  real r = zero * dx / (ctot * ctot);
  r = fmx(r, 0.0001);
  return dx / ctot * (rsqrt64(one + two * M.courant_factor * r) - one) / r;
}

#undef zero
#undef one
#undef two
#undef half
#undef forth
}  // namespace rgpu
