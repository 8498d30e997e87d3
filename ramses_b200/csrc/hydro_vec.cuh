// hydro_vec.cuh -- lane-generic forms of the hydro Riemann solvers (sm_100a FP64 pipes).
//
// The FP64 pipe of a B200 SM needs ~4 independent FP64 instructions in flight per scheduler to run at its peak
// (profiles/microbench/fp64_latency_b200.txt: dependent DFMA chains reach 59 % of the peak with 12 warps per SM, 91 % with
// ILP 4 at 8 warps), and a Riemann solver is one long dependent chain.  The three face solves of a cell (x, y and z
// direction) are independent, so the dense sweep evaluates them as ONE solver on a 3-lane value type: every statement below
// then expands to three independent instructions inside one basic block and the instruction scheduler interleaves the
// chains.  To make that possible the solvers are written without data-dependent branches: the branches of the reference
// (hydro/godunov_utils.f90) become selects between values that are computed with exactly the reference's operations in the
// reference's order, so every lane is bit-identical to the scalar, branching form of hydro_device.cuh
// (tests/test_device_numerics_host.py::test_vec_solvers_equal_scalar compares them on random face states, V = double and
// V = V3).  Compile with -fmad=false like the rest of the strict build.
#pragma once
#include "hydro_device.cuh"

namespace rgpu {

struct V3 { double a, b, c; };
struct B3 { bool a, b, c; };

#define RGPU_V3_BINOP(OP)                                                                                             \
  __device__ __forceinline__ V3 operator OP(const V3& x, const V3& y) { return {x.a OP y.a, x.b OP y.b, x.c OP y.c}; } \
  __device__ __forceinline__ V3 operator OP(const V3& x, double y) { return {x.a OP y, x.b OP y, x.c OP y}; }          \
  __device__ __forceinline__ V3 operator OP(double x, const V3& y) { return {x OP y.a, x OP y.b, x OP y.c}; }
RGPU_V3_BINOP(+)
RGPU_V3_BINOP(-)
RGPU_V3_BINOP(*)
#undef RGPU_V3_BINOP
__device__ __forceinline__ V3 operator-(const V3& x) { return {-x.a, -x.b, -x.c}; }

// ---- masks ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool vnot(bool m) { return !m; }
__device__ __forceinline__ B3 vnot(const B3& m) { return {!m.a, !m.b, !m.c}; }
__device__ __forceinline__ bool vand(bool x, bool y) { return x && y; }
__device__ __forceinline__ B3 vand(const B3& x, const B3& y) { return {x.a && y.a, x.b && y.b, x.c && y.c}; }
__device__ __forceinline__ bool vor(bool x, bool y) { return x || y; }
__device__ __forceinline__ B3 vor(const B3& x, const B3& y) { return {x.a || y.a, x.b || y.b, x.c || y.c}; }
__device__ __forceinline__ bool vall(bool m) { return m; }
__device__ __forceinline__ bool vall(const B3& m) { return m.a && m.b && m.c; }

// comparisons (false on NaN like the Fortran / C operators)
__device__ __forceinline__ bool vgt(double x, double y) { return x > y; }
__device__ __forceinline__ B3 vgt(const V3& x, double y) { return {x.a > y, x.b > y, x.c > y}; }
__device__ __forceinline__ B3 vgt(const V3& x, const V3& y) { return {x.a > y.a, x.b > y.b, x.c > y.c}; }
__device__ __forceinline__ bool vge(double x, double y) { return x >= y; }
__device__ __forceinline__ B3 vge(const V3& x, double y) { return {x.a >= y, x.b >= y, x.c >= y}; }
__device__ __forceinline__ B3 vge(const V3& x, const V3& y) { return {x.a >= y.a, x.b >= y.b, x.c >= y.c}; }
__device__ __forceinline__ bool vlt(double x, double y) { return x < y; }
__device__ __forceinline__ B3 vlt(const V3& x, double y) { return {x.a < y, x.b < y, x.c < y}; }
__device__ __forceinline__ bool vle(double x, double y) { return x <= y; }
__device__ __forceinline__ B3 vle(const V3& x, double y) { return {x.a <= y, x.b <= y, x.c <= y}; }
__device__ __forceinline__ bool veq(double x, double y) { return x == y; }
__device__ __forceinline__ B3 veq(const V3& x, double y) { return {x.a == y, x.b == y, x.c == y}; }

__device__ __forceinline__ double vsel(bool m, double x, double y) { return m ? x : y; }
__device__ __forceinline__ V3 vsel(const B3& m, const V3& x, const V3& y) { return {m.a ? x.a : y.a, m.b ? x.b : y.b, m.c ? x.c : y.c}; }

// ---- lane-wise forms of the primitives of hydro_device.cuh -------------------------------------------------------------
#define RGPU_V3_MAP1(NAME, F)                                                                    \
  __device__ __forceinline__ double NAME(double x) { return F(x); }                              \
  __device__ __forceinline__ V3 NAME(const V3& x) { return {F(x.a), F(x.b), F(x.c)}; }
RGPU_V3_MAP1(vrcp, rcp_rn)
RGPU_V3_MAP1(vsqrt, sqrt_rn)
RGPU_V3_MAP1(vabs, fabs)
RGPU_V3_MAP1(vsign1, fsign1)
#undef RGPU_V3_MAP1
#define RGPU_V3_MAP2(NAME, F)                                                                                       \
  __device__ __forceinline__ double NAME(double x, double y) { return F(x, y); }                                   \
  __device__ __forceinline__ V3 NAME(const V3& x, const V3& y) { return {F(x.a, y.a), F(x.b, y.b), F(x.c, y.c)}; } \
  __device__ __forceinline__ V3 NAME(const V3& x, double y) { return {F(x.a, y), F(x.b, y), F(x.c, y)}; }           \
  __device__ __forceinline__ V3 NAME(double x, const V3& y) { return {F(x, y.a), F(x, y.b), F(x, y.c)}; }
RGPU_V3_MAP2(vfmax, fmx)    // floors on ordinary numbers (hydro_device.cuh: fmx, 3 instructions; fmax() costs 8)
RGPU_V3_MAP2(vfmx, fmx)     // Fortran MAX as gfortran evaluates it (hydro_device.cuh)
RGPU_V3_MAP2(vfmn, fmn)
RGPU_V3_MAP2(vfdiv, fdiv)
#undef RGPU_V3_MAP2
__device__ __forceinline__ double vdiv(double a, double b, double y) { return div_rn(a, b, y); }
__device__ __forceinline__ V3 vdiv(const V3& a, const V3& b, const V3& y) { return {div_rn(a.a, b.a, y.a), div_rn(a.b, b.b, y.b), div_rn(a.c, b.c, y.c)}; }

// rho*pow(x/po, 1/gamma) of the rarefaction branches (godunov_utils.f90:415,453): only where the mask is set -- a real
// branch per lane, the library pow() is far too long to evaluate speculatively
__device__ __forceinline__ double vpow_where(bool m, double ro, double num, double po, double ex, double other) {
  return m ? ro * pow(num / po, ex) : other;
}
__device__ __forceinline__ V3 vpow_where(const B3& m, const V3& ro, const V3& num, const V3& po, double ex, const V3& other) {
  V3 r = other;
  if (m.a) r.a = ro.a * pow(num.a / po.a, ex);
  if (m.b) r.b = ro.b * pow(num.b / po.b, ex);
  if (m.c) r.c = ro.c * pow(num.c / po.c, ex);
  return r;
}
// sqrt(fabs(gamma*pstar/rstar)) with IEEE `/` and sqrt() (godunov_utils.f90:419, :598): kept as the library operations
// because pstar may be exactly zero, outside the fast-path range of sqrt_rn
#ifdef RGPU_FAST
__device__ __forceinline__ double vcstar(double gamma, double pstar, double rstar) {   // floor far below smallc^2: sqrt_rn needs x > 0
  return sqrt_rn(fmx(fabs(gamma * pstar * rcp_rn(rstar)), 1e-300));
}
#else
__device__ __forceinline__ double vcstar(double gamma, double pstar, double rstar) { return sqrt(fabs(gamma * pstar / rstar)); }
#endif
__device__ __forceinline__ V3 vcstar(double gamma, const V3& pstar, const V3& rstar) {
  return {vcstar(gamma, pstar.a, rstar.a), vcstar(gamma, pstar.b, rstar.b), vcstar(gamma, pstar.c, rstar.c)};
}

template <class V> struct MaskOf { using type = bool; };
template <> struct MaskOf<V3> { using type = B3; };

// ---------------------------------------------------------------------------------------------------------------------------
// Solvers.  ql/qr = (rho, u_normal, P, u_t1, u_t2) per lane, fg = (mass, normal momentum, total energy, transverse momenta).
// NDIM = 3 forms (the dense 3-D sweep is the only user); statement order = hydro_device.cuh = the reference.
// ---------------------------------------------------------------------------------------------------------------------------
template <class V>
__device__ __forceinline__ void riemann_llf_v(const V* ql, const V* qr, V* fg, const Phys& P) {
  // hydro/godunov_utils.f90:660-820
  const V rl = vfmax(ql[0], P.smallr), ul = ql[1], pl = vfmax(ql[2], rl * P.smallp);
  V cl = P.gamma * pl;
  cl = vsqrt(vfdiv(cl, rl));
  const V rr = vfmax(qr[0], P.smallr), ur = qr[1], pr = vfmax(qr[2], rr * P.smallp);
  V cr = P.gamma * pr;
  cr = vsqrt(vfdiv(cr, rr));
  const V cmax = vfmx(vabs(ul) + cl, vabs(ur) + cr);
  V uL[5], uR[5];
  uL[0] = ql[0]; uR[0] = qr[0];
  uL[1] = ql[0] * ql[1]; uR[1] = qr[0] * qr[1];
  uL[2] = ql[2] * P.entho + 0.5 * ql[0] * (ql[1] * ql[1]);
  uR[2] = qr[2] * P.entho + 0.5 * qr[0] * (qr[1] * qr[1]);
  uL[2] = uL[2] + 0.5 * ql[0] * (ql[3] * ql[3]); uR[2] = uR[2] + 0.5 * qr[0] * (qr[3] * qr[3]);
  uL[2] = uL[2] + 0.5 * ql[0] * (ql[4] * ql[4]); uR[2] = uR[2] + 0.5 * qr[0] * (qr[4] * qr[4]);
#pragma unroll
  for (int n = 3; n < 5; n++) { uL[n] = ql[0] * ql[n]; uR[n] = qr[0] * qr[n]; }
  V fL, fR;
  fL = ql[1] * uL[0]; fR = qr[1] * uR[0];
  fg[0] = 0.5 * (fL + fR - cmax * (uR[0] - uL[0]));
  fL = ql[1] * uL[1] + ql[2]; fR = qr[1] * uR[1] + qr[2];
  fg[1] = 0.5 * (fL + fR - cmax * (uR[1] - uL[1]));
  fL = ql[1] * (uL[2] + ql[2]); fR = qr[1] * (uR[2] + qr[2]);
  fg[2] = 0.5 * (fL + fR - cmax * (uR[2] - uL[2]));
#pragma unroll
  for (int n = 3; n < 5; n++) {
    fL = ql[1] * uL[n]; fR = qr[1] * uR[n];
    fg[n] = 0.5 * (fL + fR - cmax * (uR[n] - uL[n]));
  }
}

template <class V>
__device__ __forceinline__ void riemann_hll_v(const V* ql, const V* qr, V* fg, const Phys& P) {
  // hydro/godunov_utils.f90:825-983
  const V rl = vfmax(ql[0], P.smallr), ul = ql[1], pl = vfmax(ql[2], rl * P.smallp);
  V cl = P.gamma * pl;
  cl = vsqrt(vfdiv(cl, rl));
  const V rr = vfmax(qr[0], P.smallr), ur = qr[1], pr = vfmax(qr[2], rr * P.smallp);
  V cr = P.gamma * pr;
  cr = vsqrt(vfdiv(cr, rr));
  const V SL = vfmn(vfmn(ul, ur) - vfmax(cl, cr), 0.0);
  const V SR = vfmx(vfmx(ul, ur) + vfmax(cl, cr), 0.0);
  V uL[5], uR[5];
  uL[0] = ql[0]; uR[0] = qr[0];
  uL[1] = ql[0] * ql[1]; uR[1] = qr[0] * qr[1];
  uL[2] = ql[2] * P.entho + 0.5 * ql[0] * (ql[1] * ql[1]);
  uR[2] = qr[2] * P.entho + 0.5 * qr[0] * (qr[1] * qr[1]);
  uL[2] = uL[2] + 0.5 * ql[0] * (ql[3] * ql[3]); uR[2] = uR[2] + 0.5 * qr[0] * (qr[3] * qr[3]);
  uL[2] = uL[2] + 0.5 * ql[0] * (ql[4] * ql[4]); uR[2] = uR[2] + 0.5 * qr[0] * (qr[4] * qr[4]);
#pragma unroll
  for (int n = 3; n < 5; n++) { uL[n] = ql[0] * ql[n]; uR[n] = qr[0] * qr[n]; }
  V fL, fR;
  const V den = SR - SL, yd = vrcp(den);
  fL = uL[1]; fR = uR[1];
  fg[0] = vdiv(SR * fL - SL * fR + SR * SL * (uR[0] - uL[0]), den, yd);
  fL = ql[2] + uL[1] * ql[1]; fR = qr[2] + uR[1] * qr[1];
  fg[1] = vdiv(SR * fL - SL * fR + SR * SL * (uR[1] - uL[1]), den, yd);
  fL = ql[1] * (uL[2] + ql[2]); fR = qr[1] * (uR[2] + qr[2]);
  fg[2] = vdiv(SR * fL - SL * fR + SR * SL * (uR[2] - uL[2]), den, yd);
#pragma unroll
  for (int n = 3; n < 5; n++) {
    fL = ql[1] * uL[n]; fR = qr[1] * uR[n];
    fg[n] = vdiv(SR * fL - SL * fR + SR * SL * (uR[n] - uL[n]), den, yd);
  }
}

template <class V>
__device__ __forceinline__ void riemann_hllc_v(const V* ql, const V* qr, V* fg, const Phys& P) {
  // hydro/godunov_utils.f90:988-1209 (Toro's HLLC).  The four-way sampling :1126-1170
  //   SL>0: left state | ustar>0: left star state | SR>0: right star state | else: right state
  // is evaluated as: side = left iff (SL>0 or ustar>0); the star state of that side is formed unconditionally with the
  // reference's expressions and selected iff not(SL>0) and (ustar>0 or SR>0).
  using M = typename MaskOf<V>::type;
  const V rl = vfmax(ql[0], P.smallr), Pl = vfmax(ql[2], rl * P.smallp), ul = ql[1];
  const V el = Pl * P.entho;
  V ecinl = 0.5 * rl * ul * ul;
  ecinl = ecinl + 0.5 * rl * (ql[3] * ql[3]);
  ecinl = ecinl + 0.5 * rl * (ql[4] * ql[4]);
  const V etotl = el + ecinl;
  const V rr = vfmax(qr[0], P.smallr), Pr = vfmax(qr[2], rr * P.smallp), ur = qr[1];
  const V er = Pr * P.entho;
  V ecinr = 0.5 * rr * ur * ur;
  ecinr = ecinr + 0.5 * rr * (qr[3] * qr[3]);
  ecinr = ecinr + 0.5 * rr * (qr[4] * qr[4]);
  const V etotr = er + ecinr;
  V cfastl = P.gamma * Pl;
  cfastl = vsqrt(vfmax(vfdiv(cfastl, rl), P.smallc2));
  V cfastr = P.gamma * Pr;
  cfastr = vsqrt(vfmax(vfdiv(cfastr, rr), P.smallc2));
  const V cmaxlr = vfmax(cfastl, cfastr);
  const V SL = vfmn(ul, ur) - cmaxlr;
  const V SR = vfmx(ul, ur) + cmaxlr;
  const V rcl = rl * (ul - SL), rcr = rr * (SR - ur);
  const V rcs = rcr + rcl, yrc = vrcp(rcs);
  const V ustar = vdiv(rcr * ur + rcl * ul + (Pl - Pr), rcs, yrc);
  const V Pstar = vdiv(rcr * Pl + rcl * Pr + rcl * rcr * (ul - ur), rcs, yrc);
  const M c1 = vgt(SL, 0.0), c2 = vgt(ustar, 0.0), c3 = vgt(SR, 0.0);
  const M side_l = vor(c1, c2);
  const M use_star = vand(vnot(c1), vor(c2, c3));
  const V r = vsel(side_l, rl, rr), u = vsel(side_l, ul, ur), Pq = vsel(side_l, Pl, Pr), et = vsel(side_l, etotl, etotr);
  const V S = vsel(side_l, SL, SR);
  const V den = S - ustar, yd = vrcp(den);
  const V rs = vdiv(r * (S - u), den, yd);
  const V es = vdiv((S - u) * et - Pq * u + Pstar * ustar, den, yd);
  const V ro = vsel(use_star, rs, r), uo = vsel(use_star, ustar, u), Po = vsel(use_star, Pstar, Pq), eto = vsel(use_star, es, et);
  fg[0] = ro * uo;
  fg[1] = ro * uo * uo + Po;
  fg[2] = (eto + Po) * uo;
#pragma unroll
  for (int n = 3; n < 5; n++) fg[n] = ro * uo * vsel(c2, ql[n], qr[n]);   // :1190-1203 (the upwind value is selected first: one product)
}

// shared tail of 'exact' and 'acoustic' (godunov_utils.f90:474-493, :634-652)
template <class V, class M>
__device__ __forceinline__ void flux_from_sample_v(const V& qg1, const V& qg2, const V& qg3, const M& left, const V* ql, const V* qr, V* fg,
                                                   const Phys& P) {
  fg[0] = qg1 * qg2;
  fg[1] = qg3 + qg1 * (qg2 * qg2);
  V etot = qg3 * P.entho + 0.5 * qg1 * (qg2 * qg2);
#pragma unroll
  for (int n = 3; n < 5; n++) {
    const V qt = vsel(left, ql[n], qr[n]);
    etot = etot + 0.5 * qg1 * (qt * qt);
  }
  fg[2] = qg2 * (etot + qg3);
#pragma unroll
  for (int n = 3; n < 5; n++) fg[n] = fg[0] * vsel(left, ql[n], qr[n]);
}

template <class V>
__device__ __forceinline__ void riemann_acoustic_v(const V* ql, const V* qr, V* fg, const Phys& P) {
  // hydro/godunov_utils.f90:500-655
  using M = typename MaskOf<V>::type;
  const V rl = vfmax(ql[0], P.smallr), ul = ql[1], pl = vfmax(ql[2], rl * P.smallp);
  const V rr = vfmax(qr[0], P.smallr), ur = qr[1], pr = vfmax(qr[2], rr * P.smallp);
  const V cl = vsqrt(vfdiv(P.gamma * pl, rl)), cr = vsqrt(vfdiv(P.gamma * pr, rr));
  const V wl = cl * rl, wr = cr * rr;
  const V wsum = wl + wr, yw = vrcp(wsum);
  const V pstar = vdiv((wr * pl + wl * pr) + wl * wr * (ul - ur), wsum, yw);
  const V ustar = vdiv((wr * ur + wl * ul) + (pl - pr), wsum, yw);
  const V sgnm = vsign1(ustar);
  const M left = veq(sgnm, 1.0);
  const V ro = vsel(left, rl, rr), uo = vsel(left, ul, ur), po = vsel(left, pl, pr), co = vsel(left, cl, cr);
  V rstar = ro + vfdiv(pstar - po, co * co);
  rstar = vfmx(rstar, P.smallr);
  V cstar = vcstar(P.gamma, pstar, rstar);
  cstar = vfmx(cstar, P.smallc);
  V spout = co - sgnm * uo;
  V spin = cstar - sgnm * ustar;
  V ushock = 0.5 * (spin + spout);
  ushock = vfmx(ushock, -sgnm * ustar);
  const M shock = vge(pstar, po);
  spout = vsel(shock, ushock, spout);
  spin = vsel(shock, ushock, spin);            // :611-614: spout = ushock; spin = spout
  // sampling :618-633: spout<0 -> outer state; spin>=0 -> star state; else inside the fan
  const M outer = vlt(spout, 0.0), star = vge(spin, 0.0);
  const V frac = vfdiv(spout, spout - spin);
  const V f1 = frac * rstar + (1.0 - frac) * ro;
  const V f2 = frac * ustar + (1.0 - frac) * uo;
  const V f3 = frac * pstar + (1.0 - frac) * po;
  const V g1 = vsel(outer, ro, vsel(star, rstar, f1));
  const V g2 = vsel(outer, uo, vsel(star, ustar, f2));
  const V g3 = vsel(outer, po, vsel(star, pstar, f3));
  flux_from_sample_v(g1, g2, g3, left, ql, qr, fg, P);
}

template <class V>
__device__ __forceinline__ void riemann_exact_v(const V* ql, const V* qr, V* fg, const Phys& P) {
  // riemann_approx, hydro/godunov_utils.f90:268-495: two-shock Newton-Raphson.  Every lane iterates until ITS OWN
  // convergence test (:357-361) fires and is frozen afterwards, exactly the per-interface compaction of the reference.
  using M = typename MaskOf<V>::type;
  const V rl = vfmax(ql[0], P.smallr), ul = ql[1], pl = vfmax(ql[2], rl * P.smallp);
  const V rr = vfmax(qr[0], P.smallr), ur = qr[1], pr = vfmax(qr[2], rr * P.smallp);
  const V cl = P.gamma * pl * rl, cr = P.gamma * pr * rr;
  V wl = vsqrt(cl), wr = vsqrt(cr);
  V pstar = vfdiv((wr * pl + wl * pr) + wl * wr * (ul - ur), wl + wr);
  pstar = vfmx(pstar, 0.0);
  V pold = pstar;
  const V ypl = vrcp(pl), ypr = vrcp(pr);
  M done = vlt(pl, -1.0);                     // all false (pl >= smallr*smallp > 0)
  for (int iter = 0; iter < P.niter_riemann; iter++) {
    const V wwl = vsqrt(cl * (1.0 + vdiv(P.gamma6 * (pold - pl), pl, ypl)));
    const V wwr = vsqrt(cr * (1.0 + vdiv(P.gamma6 * (pold - pr), pr, ypr)));
    const V ywl = vrcp(wwl), ywr = vrcp(wwr);
    const V qql = vfdiv(2.0 * (wwl * wwl * wwl), wwl * wwl + cl);
    const V qqr = vfdiv(2.0 * (wwr * wwr * wwr), wwr * wwr + cr);
    const V usl = ul - vdiv(pold - pl, wwl, ywl);
    const V usr = ur + vdiv(pold - pr, wwr, ywr);
    const V delp = vfmx(vfdiv(qqr * qql, qqr + qql) * (usl - usr), -pold);
    const V pnew = pold + delp;
    const V conv = vabs(vfdiv(delp, pnew + P.smallpp));
    pold = vsel(done, pold, pnew);
    done = vor(done, vnot(vgt(conv, 1e-06)));
    if (vall(done)) break;
  }
  pstar = pold;
  wl = vsqrt(cl * (1.0 + vdiv(P.gamma6 * (pstar - pl), pl, ypl)));
  wr = vsqrt(cr * (1.0 + vdiv(P.gamma6 * (pstar - pr), pr, ypr)));
  const V ustar = 0.5 * (ul + vfdiv(pl - pstar, wl) + ur - vfdiv(pr - pstar, wr));
  const V sgnm = vsign1(ustar);
  const M left = veq(sgnm, 1.0);
  const V ro = vsel(left, rl, rr), uo = vsel(left, ul, ur), po = vsel(left, pl, pr), wo = vsel(left, wl, wr);
  const V yro = vrcp(ro);
  const V co = vfmx(P.smallc, vsqrt(vabs(vdiv(P.gamma * po, ro, yro))));
  const M shock = vge(pstar, po);
  V rstar = vfdiv(ro, 1.0 + vfdiv(ro * (po - pstar), wo * wo));          // :411-413 (shock); rarefaction :415
  rstar = vpow_where(vnot(shock), ro, pstar, po, P.inv_gamma, rstar);
  rstar = vfmx(rstar, P.smallr);
  V cstar = vcstar(P.gamma, pstar, rstar);
  cstar = vfmx(cstar, P.smallc);
  V spout = co - sgnm * uo;
  V spin = cstar - sgnm * ustar;
  const V ushock = vdiv(wo, ro, yro) - sgnm * uo;
  spout = vsel(shock, ushock, spout);
  spin = vsel(shock, ushock, spin);
  // sampling :440-456: spout<=0 -> outer; spin>=0 -> star; else fan with the isentropic density
  const M outer = vle(spout, 0.0), star = vge(spin, 0.0);
  const M fan = vand(vnot(outer), vnot(star));
  const V frac = vfdiv(spout, spout - spin);
  const V f2 = frac * ustar + (1.0 - frac) * uo;
  const V f3 = frac * pstar + (1.0 - frac) * po;
  V g1 = vsel(outer, ro, rstar);
  g1 = vpow_where(fan, ro, f3, po, P.inv_gamma, g1);
  const V g2 = vsel(outer, uo, vsel(star, ustar, f2));
  const V g3 = vsel(outer, po, vsel(star, pstar, f3));
  flux_from_sample_v(g1, g2, g3, left, ql, qr, fg, P);
}

template <int RIEMANN, class V>
__device__ __forceinline__ void riemann_v(const V* ql, const V* qr, V* fg, const Phys& P) {
  if (RIEMANN == RIEMANN_LLF) riemann_llf_v<V>(ql, qr, fg, P);
  else if (RIEMANN == RIEMANN_HLL) riemann_hll_v<V>(ql, qr, fg, P);
  else if (RIEMANN == RIEMANN_HLLC) riemann_hllc_v<V>(ql, qr, fg, P);
  else if (RIEMANN == RIEMANN_ACOUSTIC) riemann_acoustic_v<V>(ql, qr, fg, P);
  else riemann_exact_v<V>(ql, qr, fg, P);
}

}  // namespace rgpu
