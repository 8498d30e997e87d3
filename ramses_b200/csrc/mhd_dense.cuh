// mhd_dense.cuh -- ideal-MHD Godunov sweep (constrained transport) of one dense-box level on sm_100a.
//
// Reference path: mhd/godunov_fine.f90 godfine1:538 -> mhd/umuscl.f90 mag_unsplit:31 (ctoprim:2029, uslope:2187,
// trace3d:750, cmpflxm:1308, cmp_mag_flx:1453) -> mhd/godunov_utils.f90 solvers; NDIM=3, nvar=8 stored as 11 variables
// (left-face B in 6..8, right-face B in 9..11).
//
// The reference evaluates every oct on its own 6^3 patch (each face flux and edge EMF is recomputed by up to 4 octs).  Here
// each quantity is computed once per cell / face / edge of the dense box, in six passes over cell-ordered work arrays
// W[comp][z][y][x] (x fastest, coalesced):
//   1 mhd_prim_kernel    u -> q(8) and the staggered field bf(3)                      ctoprim   :2029-2153
//   2 mhd_efield_kernel  edge-centred E = -v x B from 4-cell averages                 trace3d   :812-836
//   3 mhd_trace_kernel   TVD slopes, CT half-step of the face fields, source terms -> the predicted cell state in a
//                        compact form (44 doubles) from which every face state qm/qp and edge state qRT/qRB/qLT/qLB
//                        is a few additions                                            uslope :2187, trace3d :838-1267
//   4 mhd_flux_kernel    1-D Riemann problems at the three low faces of a cell         cmpflxm   :1308
//   5 mhd_emf_kernel     2-D Riemann problems at the three low edges of a cell         cmp_mag_flx :1453
//   6 mhd_update_kernel  conservative update + CT update of both face copies + fused courant_fine of the new state
//                                                           godfine1 :883-995, mhd/courant_fine.f90, cmpdt godunov_utils.f90:5
// On consistent data (right face of cell i == left face of cell i+1, which the update preserves bit for bit) this equals
// the per-oct evaluation: identical inputs give identical fluxes whichever oct computes them.
#pragma once
#include "mhd_device.cuh"
#include "sweep_dense.cuh"

namespace rgpu {

enum { MW_Q = 0, MW_BF = 8, MW_E = 11, MW_TR = 14, MW_F = 58, MW_EM = 73, MW_NCOMP = 76 };
// compact trace record
enum { TR_R = 0, TR_U, TR_V, TR_W, TR_P, TR_AL, TR_AR, TR_BL, TR_BR, TR_CL, TR_CR,
       TR_SX = 11,   // drx,dux,dvx,dwx,dpx,dBx,dCx
       TR_SY = 18,   // dry,duy,dvy,dwy,dpy,dAy,dCy
       TR_SZ = 25,   // drz,duz,dvz,dwz,dpz,dAz,dBz
       TR_FS = 32,   // dALy,dARy,dALz,dARz,dBLx,dBRx,dBLz,dBRz,dCLx,dCRx,dCLy,dCRy
       TR_N = 44 };

struct MhdArgs {
  const double* uin;   // uold  [11][8][nslot]
  double* uout;        // unew after set_uold (other ping-pong buffer)
  DenseGeom g;
  MPhys P;
  const double* dt_dev;
  double dx;
  double* W;           // [MW_NCOMP][nc]
  long long nc;        // ncx*ncy*ncz
  double* part;        // per-CTA partials [5][gridDim.x]: min dt, mass, etot, eint, emag of the new state
};

struct Cxyz { int x, y, z; };
__device__ __forceinline__ Cxyz cell_xyz(const DenseGeom& g, long long c) {
  Cxyz r;
  r.x = (int)(c % g.ncx);
  const long long t = c / g.ncx;
  r.y = (int)(t % g.ncy);
  r.z = (int)(t / g.ncy);
  return r;
}
__device__ __forceinline__ long long cidx(const DenseGeom& g, int x, int y, int z) { return ((long long)z * g.ncy + y) * g.ncx + x; }
// neighbour index with periodic wrap inside the box (dimensions without ghost shell)
__device__ __forceinline__ int wm(int c, int n, int wrap) { return (c > 0) ? c - 1 : (wrap ? n - 1 : 0); }
__device__ __forceinline__ int wp(int c, int n, int wrap) { return (c < n - 1) ? c + 1 : (wrap ? 0 : n - 1); }
// range of cells whose traced state is needed: owned range widened by one cell (all cells of a wrapped dimension)
__device__ __forceinline__ bool in_range(int c, int lo, int hi, int n, int wrap, int wlo, int whi) {
  if (wrap) return true;
  (void)n;
  return c >= lo - wlo && c < hi + whi;
}

#ifdef MHD_DEFINE_KERNELS
// ---------------------------------------------------------------------------------------------------- pass 1
__global__ void __launch_bounds__(256) mhd_prim_kernel(const MhdArgs a) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  const DenseGeom& g = a.g;
  const Cxyz p = cell_xyz(g, c);
  const long long off = cell_offset<3>(g, p.x, p.y, p.z);
  const long long vs = 8 * g.nslot;
  double u[11];
#pragma unroll
  for (int n = 0; n < 11; n++) u[n] = __ldg(a.uin + n * vs + off);
  double q[8];
  q[0] = fmx(u[0], a.P.smallr);
  const double rq0 = rcp_rn(q[0]);   // shared reciprocal, same bits as the four IEEE divisions (hydro_device.cuh div_rn)
  q[1] = div_rn(u[1], q[0], rq0); q[2] = div_rn(u[2], q[0], rq0); q[3] = div_rn(u[3], q[0], rq0);
  q[5] = (u[5] + u[8]) * 0.5;
  q[6] = (u[6] + u[9]) * 0.5;
  q[7] = (u[7] + u[10]) * 0.5;
  const double eken = 0.5 * (q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double emag = 0.5 * (q[5] * q[5] + q[6] * q[6] + q[7] * q[7]);
  const double etot = u[4] - emag - 0.0;
  const double eint = div_rn(etot, q[0], rq0) - eken;
  q[4] = fmx((a.P.gamma - 1.0) * q[0] * eint, a.P.smallp);
  double* W = a.W;
#pragma unroll
  for (int n = 0; n < 8; n++) W[(MW_Q + n) * a.nc + c] = q[n];
  W[(MW_BF + 0) * a.nc + c] = u[5];
  W[(MW_BF + 1) * a.nc + c] = u[6];
  W[(MW_BF + 2) * a.nc + c] = u[7];
}

// ---------------------------------------------------------------------------------------------------- pass 2
__global__ void __launch_bounds__(256) mhd_efield_kernel(const MhdArgs a) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  const DenseGeom& g = a.g;
  const Cxyz p = cell_xyz(g, c);
  const bool hx = g.wrapx || p.x > 0, hy = g.wrapy || p.y > 0, hz = g.wrapz || p.z > 0;
  const int xm = wm(p.x, g.ncx, g.wrapx), ym = wm(p.y, g.ncy, g.wrapy), zm = wm(p.z, g.ncz, g.wrapz);
  const double* Q = a.W + MW_Q * a.nc;
  const double* BF = a.W + MW_BF * a.nc;
  double* E = a.W + MW_E * a.nc;
  const long long nc = a.nc;
  const long long c_y = cidx(g, p.x, ym, p.z), c_z = cidx(g, p.x, p.y, zm), c_x = cidx(g, xm, p.y, p.z);
#define QV(n, cc) Q[(n) * nc + (cc)]
  if (hy && hz) {
    const long long c_yz = cidx(g, p.x, ym, zm);
    const double v = 0.25 * (QV(2, c_yz) + QV(2, c_y) + QV(2, c_z) + QV(2, c));
    const double w = 0.25 * (QV(3, c_yz) + QV(3, c_y) + QV(3, c_z) + QV(3, c));
    const double B = 0.5 * (BF[1 * nc + c_z] + BF[1 * nc + c]);
    const double C = 0.5 * (BF[2 * nc + c_y] + BF[2 * nc + c]);
    E[0 * nc + c] = v * C - w * B;
  }
  if (hx && hz) {
    const long long c_xz = cidx(g, xm, p.y, zm);
    const double u = 0.25 * (QV(1, c_xz) + QV(1, c_x) + QV(1, c_z) + QV(1, c));
    const double w = 0.25 * (QV(3, c_xz) + QV(3, c_x) + QV(3, c_z) + QV(3, c));
    const double A = 0.5 * (BF[0 * nc + c_z] + BF[0 * nc + c]);
    const double C = 0.5 * (BF[2 * nc + c_x] + BF[2 * nc + c]);
    E[1 * nc + c] = w * A - u * C;
  }
  if (hx && hy) {
    const long long c_xy = cidx(g, xm, ym, p.z);
    const double u = 0.25 * (QV(1, c_xy) + QV(1, c_x) + QV(1, c_y) + QV(1, c));
    const double v = 0.25 * (QV(2, c_xy) + QV(2, c_x) + QV(2, c_y) + QV(2, c));
    const double A = 0.5 * (BF[0 * nc + c_y] + BF[0 * nc + c]);
    const double B = 0.5 * (BF[1 * nc + c_x] + BF[1 * nc + c]);
    E[2 * nc + c] = u * B - v * A;
  }
#undef QV
}

#endif  // MHD_DEFINE_KERNELS (passes 1, 2)
// ---------------------------------------------------------------------------------------------------- pass 3
__device__ __forceinline__ double slope_mmd(double st, double ql, double qc, double qr) { return slope_mm(st, ql, qc, qr).v; }
// one limited slope of the cell-centred variables (all slope types of the NDIM=3 MHD build except type 3, which needs the
// whole 3^3 neighbourhood and is handled by the caller)
__device__ __forceinline__ double mhd_slope1(int st, double theta, double ql, double qc, double qr) {
  if (st == 1 || st == 2) return slope_mm((double)st, ql, qc, qr).v;
  const double dlft = qc - ql, drgt = qr - qc;
  if (st == 7) return ((dlft * drgt) <= 0.0) ? 0.0 : (2 * dlft * drgt / (dlft + drgt));
  // st == 8
  const double dcen = 0.5 * (dlft + drgt);
  const double dsgn = copysign(1.0, dcen);
  double dlim = fmn(theta * fabs(dlft), theta * fabs(drgt));
  if ((dlft * drgt) <= 0.0) dlim = 0.0;
  return dsgn * fmn(dlim, fabs(dcen));
}

template <bool SL>
__global__ void __launch_bounds__(256) mhd_trace_kernel(const MhdArgs a) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  const DenseGeom& g = a.g;
  const Cxyz p = cell_xyz(g, c);
  if (!in_range(p.x, g.ox0, g.ox1, g.ncx, g.wrapx, 1, 1) || !in_range(p.y, g.oy0, g.oy1, g.ncy, g.wrapy, 1, 1) ||
      !in_range(p.z, g.oz0, g.oz1, g.ncz, g.wrapz, 1, 1))
    return;
  const MPhys& P = a.P;
  const long long nc = a.nc;
  const double* Q = a.W + MW_Q * nc;
  const double* BF = a.W + MW_BF * nc;
  const double* E = a.W + MW_E * nc;
  double* TR = a.W + MW_TR * nc;
  const int xm = wm(p.x, g.ncx, g.wrapx), ym = wm(p.y, g.ncy, g.wrapy), zm = wm(p.z, g.ncz, g.wrapz);
  const int xp = wp(p.x, g.ncx, g.wrapx), yp = wp(p.y, g.ncy, g.wrapy), zp = wp(p.z, g.ncz, g.wrapz);
  const long long cxm = cidx(g, xm, p.y, p.z), cxp = cidx(g, xp, p.y, p.z);
  const long long cym = cidx(g, p.x, ym, p.z), cyp = cidx(g, p.x, yp, p.z);
  const long long czm = cidx(g, p.x, p.y, zm), czp = cidx(g, p.x, p.y, zp);
  const double dt = *a.dt_dev;
  const double dtdx = dt / a.dx, dtdy = dtdx, dtdz = dtdx;
  double q[8];
#pragma unroll
  for (int n = 0; n < 8; n++) q[n] = Q[n * nc + c];
  double r = q[0], u = q[1], v = q[2], w = q[3], pp = q[4], A = q[5], B = q[6], C = q[7];
  double AL = BF[0 * nc + c], AR = BF[0 * nc + cxp], BL = BF[1 * nc + c], BR = BF[1 * nc + cyp], CL = BF[2 * nc + c], CR = BF[2 * nc + czp];
  double sx[8], sy[8], sz[8];   // half slopes of (r,u,v,w,p,A,B,C) per direction
  double fs[12];
#pragma unroll
  for (int n = 0; n < 8; n++) { sx[n] = 0.0; sy[n] = 0.0; sz[n] = 0.0; }
#pragma unroll
  for (int n = 0; n < 12; n++) fs[n] = 0.0;
  if (SL) {
    const int st = P.slope_type, smt = P.slope_mag_type;
    if (st == 3) {   // positivity preserving :2410-2470
      for (int n = 0; n < 8; n++) {
        const double qc = q[n];
        double vmin = 0.0, vmax = 0.0;
        bool first = true;
        for (int dk = -1; dk <= 1; dk++)
          for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) {
              const int xx = di < 0 ? xm : di > 0 ? xp : p.x, yy = dj < 0 ? ym : dj > 0 ? yp : p.y, zz = dk < 0 ? zm : dk > 0 ? zp : p.z;
              const double df = Q[n * nc + cidx(g, xx, yy, zz)] - qc;
              if (first) { vmin = df; vmax = df; first = false; }
              else { vmin = fmn(vmin, df); vmax = fmx(vmax, df); }
            }
        const double dfx = 0.5 * (Q[n * nc + cxp] - Q[n * nc + cxm]);
        const double dfy = 0.5 * (Q[n * nc + cyp] - Q[n * nc + cym]);
        const double dfz = 0.5 * (Q[n * nc + czp] - Q[n * nc + czm]);
        const double dff = 0.5 * (fabs(dfx) + fabs(dfy) + fabs(dfz));
        const double slop = (dff > 0.0) ? fmn(1.0, fmn(fabs(vmin), fabs(vmax)) / dff) : 1.0;
        sx[n] = 0.5 * (slop * dfx); sy[n] = 0.5 * (slop * dfy); sz[n] = 0.5 * (slop * dfz);
      }
    } else if (st != 0) {
#pragma unroll
      for (int n = 0; n < 8; n++) {
        sx[n] = 0.5 * mhd_slope1(st, P.slope_theta, Q[n * nc + cxm], q[n], Q[n * nc + cxp]);
        sy[n] = 0.5 * mhd_slope1(st, P.slope_theta, Q[n * nc + cym], q[n], Q[n * nc + cyp]);
        sz[n] = 0.5 * mhd_slope1(st, P.slope_theta, Q[n * nc + czm], q[n], Q[n * nc + czp]);
      }
    }
    if (smt == 1 || smt == 2) {   // transverse slopes of the face fields :2565-2650
      const double s = (double)smt;
      const double* Bx = BF;
      const double* By = BF + nc;
      const double* Bz = BF + 2 * nc;
      fs[0] = 0.5 * slope_mmd(s, Bx[cym], Bx[c], Bx[cyp]);                                                     // dALy
      fs[1] = 0.5 * slope_mmd(s, Bx[cidx(g, xp, ym, p.z)], Bx[cxp], Bx[cidx(g, xp, yp, p.z)]);                 // dARy
      fs[2] = 0.5 * slope_mmd(s, Bx[czm], Bx[c], Bx[czp]);                                                     // dALz
      fs[3] = 0.5 * slope_mmd(s, Bx[cidx(g, xp, p.y, zm)], Bx[cxp], Bx[cidx(g, xp, p.y, zp)]);                 // dARz
      fs[4] = 0.5 * slope_mmd(s, By[cxm], By[c], By[cxp]);                                                     // dBLx
      fs[5] = 0.5 * slope_mmd(s, By[cidx(g, xm, yp, p.z)], By[cyp], By[cidx(g, xp, yp, p.z)]);                 // dBRx
      fs[6] = 0.5 * slope_mmd(s, By[czm], By[c], By[czp]);                                                     // dBLz
      fs[7] = 0.5 * slope_mmd(s, By[cidx(g, p.x, yp, zm)], By[cyp], By[cidx(g, p.x, yp, zp)]);                 // dBRz
      fs[8] = 0.5 * slope_mmd(s, Bz[cxm], Bz[c], Bz[cxp]);                                                     // dCLx
      fs[9] = 0.5 * slope_mmd(s, Bz[cidx(g, xm, p.y, zp)], Bz[czp], Bz[cidx(g, xp, p.y, zp)]);                 // dCRx
      fs[10] = 0.5 * slope_mmd(s, Bz[cym], Bz[c], Bz[cyp]);                                                    // dCLy
      fs[11] = 0.5 * slope_mmd(s, Bz[cidx(g, p.x, ym, zp)], Bz[czp], Bz[cidx(g, p.x, yp, zp)]);                // dCRy
    }
  }
  // CT half-step of the face fields :905-947
  {
    const double* Ex = E; const double* Ey = E + nc; const double* Ez = E + 2 * nc;
    const double ELL = Ex[c], ELR = Ex[czp], ERL = Ex[cyp], ERR = Ex[cidx(g, p.x, yp, zp)];
    const double FLL = Ey[c], FLR = Ey[czp], FRL = Ey[cxp], FRR = Ey[cidx(g, xp, p.y, zp)];
    const double GLL = Ez[c], GLR = Ez[cyp], GRL = Ez[cxp], GRR = Ez[cidx(g, xp, yp, p.z)];
    const double sAL0 = +(GLR - GLL) * dtdy * 0.5 - (FLR - FLL) * dtdz * 0.5;
    const double sAR0 = +(GRR - GRL) * dtdy * 0.5 - (FRR - FRL) * dtdz * 0.5;
    const double sBL0 = -(GRL - GLL) * dtdx * 0.5 + (ELR - ELL) * dtdz * 0.5;
    const double sBR0 = -(GRR - GLR) * dtdx * 0.5 + (ERR - ERL) * dtdz * 0.5;
    const double sCL0 = +(FRL - FLL) * dtdx * 0.5 - (ERL - ELL) * dtdy * 0.5;
    const double sCR0 = +(FRR - FLR) * dtdx * 0.5 - (ERR - ELR) * dtdy * 0.5;
    AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0; CL = CL + sCL0; CR = CR + sCR0;
  }
  if (SL) {   // source terms :949-975 (all zero for vanishing slopes: x + 0 keeps x, the predicted state is the cell state)
    const double drx = sx[0], dux = sx[1], dvx = sx[2], dwx = sx[3], dpx = sx[4], dBx = sx[6], dCx = sx[7];
    const double dry = sy[0], duy = sy[1], dvy = sy[2], dwy = sy[3], dpy = sy[4], dAy = sy[5], dCy = sy[7];
    const double drz = sz[0], duz = sz[1], dvz = sz[2], dwz = sz[3], dpz = sz[4], dAz = sz[5], dBz = sz[6];
    const double gamma = P.gamma;
    const double sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy + (-w * drz - dwz * r) * dtdz;
    const double rr = rcp_rn(r);   // r >= smallr; the nine quotients share the reciprocal (same bits as IEEE `/`)
#define DR(x) div_rn((x), r, rr)
    const double su0 = (-u * dux - DR(dpx + B * dBx + C * dCx)) * dtdx + (-v * duy + DR(B * dAy)) * dtdy + (-w * duz + DR(C * dAz)) * dtdz;
    const double sv0 = (-u * dvx + DR(A * dBx)) * dtdx + (-v * dvy - DR(dpy + A * dAy + C * dCy)) * dtdy + (-w * dvz + DR(C * dBz)) * dtdz;
    const double sw0 = (-u * dwx + DR(A * dCx)) * dtdx + (-v * dwy + DR(B * dCy)) * dtdy + (-w * dwz - DR(dpz + A * dAz + B * dBz)) * dtdz;
#undef DR
    const double sp0 = (-u * dpx - dux * gamma * pp) * dtdx + (-v * dpy - dvy * gamma * pp) * dtdy + (-w * dpz - dwz * gamma * pp) * dtdz;
    r = r + sr0; u = u + su0; v = v + sv0; w = w + sw0; pp = pp + sp0;
  }
  TR[TR_R * nc + c] = r; TR[TR_U * nc + c] = u; TR[TR_V * nc + c] = v; TR[TR_W * nc + c] = w; TR[TR_P * nc + c] = pp;
  TR[TR_AL * nc + c] = AL; TR[TR_AR * nc + c] = AR; TR[TR_BL * nc + c] = BL; TR[TR_BR * nc + c] = BR;
  TR[TR_CL * nc + c] = CL; TR[TR_CR * nc + c] = CR;
  if (SL) {
    TR[(TR_SX + 0) * nc + c] = sx[0]; TR[(TR_SX + 1) * nc + c] = sx[1]; TR[(TR_SX + 2) * nc + c] = sx[2]; TR[(TR_SX + 3) * nc + c] = sx[3];
    TR[(TR_SX + 4) * nc + c] = sx[4]; TR[(TR_SX + 5) * nc + c] = sx[6]; TR[(TR_SX + 6) * nc + c] = sx[7];
    TR[(TR_SY + 0) * nc + c] = sy[0]; TR[(TR_SY + 1) * nc + c] = sy[1]; TR[(TR_SY + 2) * nc + c] = sy[2]; TR[(TR_SY + 3) * nc + c] = sy[3];
    TR[(TR_SY + 4) * nc + c] = sy[4]; TR[(TR_SY + 5) * nc + c] = sy[5]; TR[(TR_SY + 6) * nc + c] = sy[7];
    TR[(TR_SZ + 0) * nc + c] = sz[0]; TR[(TR_SZ + 1) * nc + c] = sz[1]; TR[(TR_SZ + 2) * nc + c] = sz[2]; TR[(TR_SZ + 3) * nc + c] = sz[3];
    TR[(TR_SZ + 4) * nc + c] = sz[4]; TR[(TR_SZ + 5) * nc + c] = sz[5]; TR[(TR_SZ + 6) * nc + c] = sz[6];
#pragma unroll
    for (int n = 0; n < 12; n++) TR[(TR_FS + n) * nc + c] = fs[n];
  }
}

// the compact trace record of one cell
struct TC {
  double r, u, v, w, p, AL, AR, BL, BR, CL, CR, A, B, C;
  double sx[7], sy[7], sz[7], fs[12];
};
template <bool SL>
__device__ __forceinline__ void load_tc(const double* __restrict__ TR, long long nc, long long c, TC& t) {
  t.r = TR[TR_R * nc + c]; t.u = TR[TR_U * nc + c]; t.v = TR[TR_V * nc + c]; t.w = TR[TR_W * nc + c]; t.p = TR[TR_P * nc + c];
  t.AL = TR[TR_AL * nc + c]; t.AR = TR[TR_AR * nc + c]; t.BL = TR[TR_BL * nc + c]; t.BR = TR[TR_BR * nc + c];
  t.CL = TR[TR_CL * nc + c]; t.CR = TR[TR_CR * nc + c];
  t.A = 0.5 * (t.AL + t.AR); t.B = 0.5 * (t.BL + t.BR); t.C = 0.5 * (t.CL + t.CR);   // :983-985
#pragma unroll
  for (int n = 0; n < 7; n++) {
    t.sx[n] = SL ? TR[(TR_SX + n) * nc + c] : 0.0;
    t.sy[n] = SL ? TR[(TR_SY + n) * nc + c] : 0.0;
    t.sz[n] = SL ? TR[(TR_SZ + n) * nc + c] : 0.0;
  }
#pragma unroll
  for (int n = 0; n < 12; n++) t.fs[n] = SL ? TR[(TR_FS + n) * nc + c] : 0.0;
}
__device__ __forceinline__ void clamp_state(const MPhys& P, double* s, double r) {   // :996-997
  if (s[0] < P.smallr) s[0] = r;
  s[4] = fmx(P.smallp, s[4]);
}
// face states qm (SGN=+1) / qp (SGN=-1) of direction DIR in cell-variable order (r,u,v,w,p,A,B,C) :986-1075
template <int DIR, int SGN>
__device__ __forceinline__ void face_state(const MPhys& P, const TC& t, double* s) {
  const double* d = DIR == 0 ? t.sx : DIR == 1 ? t.sy : t.sz;
  if (SGN > 0) { s[0] = t.r + d[0]; s[1] = t.u + d[1]; s[2] = t.v + d[2]; s[3] = t.w + d[3]; s[4] = t.p + d[4]; }
  else { s[0] = t.r - d[0]; s[1] = t.u - d[1]; s[2] = t.v - d[2]; s[3] = t.w - d[3]; s[4] = t.p - d[4]; }
  if (DIR == 0) {   // d[5]=dBx d[6]=dCx
    s[5] = SGN > 0 ? t.AR : t.AL;
    s[6] = SGN > 0 ? t.B + d[5] : t.B - d[5];
    s[7] = SGN > 0 ? t.C + d[6] : t.C - d[6];
  } else if (DIR == 1) {   // d[5]=dAy d[6]=dCy
    s[5] = SGN > 0 ? t.A + d[5] : t.A - d[5];
    s[6] = SGN > 0 ? t.BR : t.BL;
    s[7] = SGN > 0 ? t.C + d[6] : t.C - d[6];
  } else {   // d[5]=dAz d[6]=dBz
    s[5] = SGN > 0 ? t.A + d[5] : t.A - d[5];
    s[6] = SGN > 0 ? t.B + d[6] : t.B - d[6];
    s[7] = SGN > 0 ? t.CR : t.CL;
  }
  clamp_state(P, s, t.r);
}
// edge states of the edge along DIR; S1/S2 = sign along the first / second transverse axis (R/L and T/B of the reference
// names qRT,qRB,qLT,qLB) :1076-1267.  (+a+b), (+a-b), (-a+b), (-a-b) are evaluated as (+-a) + (+-b): same bits.
template <int DIR, int S1, int S2>
__device__ __forceinline__ void edge_state(const MPhys& P, const TC& t, double* s) {
  const double* d1 = DIR == 2 ? t.sx : DIR == 1 ? t.sx : t.sy;
  const double* d2 = DIR == 2 ? t.sy : DIR == 1 ? t.sz : t.sz;
#define PM(S, x) ((S) > 0 ? (x) : -(x))
#pragma unroll
  for (int n = 0; n < 5; n++) {
    const double c0 = n == 0 ? t.r : n == 1 ? t.u : n == 2 ? t.v : n == 3 ? t.w : t.p;
    s[n] = c0 + (PM(S1, d1[n]) + PM(S2, d2[n]));
  }
  // fs: 0 dALy,1 dARy,2 dALz,3 dARz,4 dBLx,5 dBRx,6 dBLz,7 dBRz,8 dCLx,9 dCRx,10 dCLy,11 dCRy
  if (DIR == 2) {          // axes (x,y): sx[5]=dBx sx[6]=dCx ; sy[5]=dAy sy[6]=dCy
    s[5] = (S1 > 0 ? t.AR : t.AL) + PM(S2, S1 > 0 ? t.fs[1] : t.fs[0]);
    s[6] = (S2 > 0 ? t.BR : t.BL) + PM(S1, S2 > 0 ? t.fs[5] : t.fs[4]);
    s[7] = t.C + (PM(S1, t.sx[6]) + PM(S2, t.sy[6]));
  } else if (DIR == 1) {   // axes (x,z): sx[5]=dBx ; sz[6]=dBz
    s[5] = (S1 > 0 ? t.AR : t.AL) + PM(S2, S1 > 0 ? t.fs[3] : t.fs[2]);
    s[6] = t.B + (PM(S1, t.sx[5]) + PM(S2, t.sz[6]));
    s[7] = (S2 > 0 ? t.CR : t.CL) + PM(S1, S2 > 0 ? t.fs[9] : t.fs[8]);
  } else {                 // axes (y,z): sy[5]=dAy ; sz[5]=dAz
    s[5] = t.A + (PM(S1, t.sy[5]) + PM(S2, t.sz[5]));
    s[6] = (S1 > 0 ? t.BR : t.BL) + PM(S2, S1 > 0 ? t.fs[7] : t.fs[6]);
    s[7] = (S2 > 0 ? t.CR : t.CL) + PM(S1, S2 > 0 ? t.fs[11] : t.fs[10]);
  }
#undef PM
  clamp_state(P, s, t.r);
}

// ---------------------------------------------------------------------------------------------------- pass 4
__device__ __forceinline__ double sel3(int dir, double x, double y, double z) { return dir == 0 ? x : (dir == 1 ? y : z); }

// cmpflxm for the three low faces of a cell.  For Roe the direction loop is a real loop (`#pragma unroll 1`): the solver is
// inlined ONCE, so that the kernel (2700 instructions per solve) stays inside the instruction cache -- with three
// inlined copies the top stall reason was `no_instruction` (profiles/r1_ncu_full_mhd_flux_roe.txt).  The variable
// permutations ln,lt1,lt2,bn,bt1,bt2 of the three cmpflxm calls (mhd/umuscl.f90:51,84,117) become selects.
struct FaceCtx {
  bool ox, oy, oz, ex, ey, ez;
  int x, y, z, xm, ym, zm;
  long long c;
  double dt, rdx;
};
template <int R1D, bool SL>
__device__ __forceinline__ void one_face(const MhdArgs& a, const FaceCtx& k, const TC& tr, const int dir) {
  const DenseGeom& g = a.g;
  const long long nc = a.nc, c = k.c;
  const double* TR = a.W + MW_TR * nc;
  const bool need = dir == 0 ? (k.ex && k.oy && k.oz) : dir == 1 ? (k.ox && k.ey && k.oz) : (k.ox && k.oy && k.ez);
  if (!need) return;
  const long long cl = dir == 0 ? cidx(g, k.xm, k.y, k.z) : dir == 1 ? cidx(g, k.x, k.ym, k.z) : cidx(g, k.x, k.y, k.zm);
  TC tl;
  load_tc<SL>(TR, nc, cl, tl);
  double sm[8], sp[8];   // qm of the cell below, qp of this cell
  if (dir == 0) { face_state<0, +1>(a.P, tl, sm); face_state<0, -1>(a.P, tr, sp); }
  else if (dir == 1) { face_state<1, +1>(a.P, tl, sm); face_state<1, -1>(a.P, tr, sp); }
  else { face_state<2, +1>(a.P, tl, sm); face_state<2, -1>(a.P, tr, sp); }
  real ql[8], qr[8], fg[9];
  const double bnl = sel3(dir, sm[5], sm[6], sm[7]), bnr = sel3(dir, sp[5], sp[6], sp[7]);
  const double bn_mean = 0.5 * (bnl + bnr);
  ql[0] = sm[0]; ql[1] = sm[4]; ql[2] = sel3(dir, sm[1], sm[2], sm[3]); ql[3] = bn_mean;
  ql[4] = sel3(dir, sm[2], sm[1], sm[1]); ql[5] = sel3(dir, sm[6], sm[5], sm[5]);
  ql[6] = sel3(dir, sm[3], sm[3], sm[2]); ql[7] = sel3(dir, sm[7], sm[7], sm[6]);
  qr[0] = sp[0]; qr[1] = sp[4]; qr[2] = sel3(dir, sp[1], sp[2], sp[3]); qr[3] = bn_mean;
  qr[4] = sel3(dir, sp[2], sp[1], sp[1]); qr[5] = sel3(dir, sp[6], sp[5], sp[5]);
  qr[6] = sel3(dir, sp[3], sp[3], sp[2]); qr[7] = sel3(dir, sp[7], sp[7], sp[6]);
  riemann1d<R1D>(a.P, ql, qr, fg);
  // (rho, mx, my, mz, E); the induction fluxes are dropped (flux(:,6:8)=0, godunov_fine.f90:778-879)
  double f[5];
  f[0] = fg[0].v; f[4] = fg[1].v;
  f[1] = sel3(dir, fg[2].v, fg[4].v, fg[4].v);
  f[2] = sel3(dir, fg[4].v, fg[2].v, fg[6].v);
  f[3] = sel3(dir, fg[6].v, fg[6].v, fg[2].v);
  double* F = a.W + (MW_F + 5 * dir) * nc;
#pragma unroll
  for (int n = 0; n < 5; n++) F[n * nc + c] = div_rn(f[n] * k.dt, a.dx, k.rdx);   // flux = fx*dt/dx :83
}

template <int R1D, bool SL, int MINB>
__global__ void __launch_bounds__(128, MINB) mhd_flux_kernel(const MhdArgs a) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  const DenseGeom& g = a.g;
  const Cxyz p = cell_xyz(g, c);
  FaceCtx k;
  // low faces are needed for the owned cells and for the first cell above the owned range (= high face of the last one)
  k.ox = in_range(p.x, g.ox0, g.ox1, g.ncx, g.wrapx, 0, 0); k.oy = in_range(p.y, g.oy0, g.oy1, g.ncy, g.wrapy, 0, 0);
  k.oz = in_range(p.z, g.oz0, g.oz1, g.ncz, g.wrapz, 0, 0);
  k.ex = in_range(p.x, g.ox0, g.ox1, g.ncx, g.wrapx, 0, 1); k.ey = in_range(p.y, g.oy0, g.oy1, g.ncy, g.wrapy, 0, 1);
  k.ez = in_range(p.z, g.oz0, g.oz1, g.ncz, g.wrapz, 0, 1);
  k.x = p.x; k.y = p.y; k.z = p.z;
  k.xm = wm(p.x, g.ncx, g.wrapx); k.ym = wm(p.y, g.ncy, g.wrapy); k.zm = wm(p.z, g.ncz, g.wrapz);
  k.c = c; k.dt = *a.dt_dev; k.rdx = rcp_rn(a.dx);
  TC tr;
  load_tc<SL>(a.W + MW_TR * a.nc, a.nc, c, tr);
  if (R1D == MHD_ROE) {   // measured: one copy wins for Roe (10.8 -> 10.35 ms/step), three copies for the small solvers
#pragma unroll 1
    for (int dir = 0; dir < 3; dir++) one_face<R1D, SL>(a, k, tr, dir);
  } else {
    one_face<R1D, SL>(a, k, tr, 0); one_face<R1D, SL>(a, k, tr, 1); one_face<R1D, SL>(a, k, tr, 2);
  }
}

// ---------------------------------------------------------------------------------------------------- pass 5
// cmp_mag_flx at the three low edges of a cell, one inlined copy of the 2-D solver (see pass 4).  The four corner states
// arrive in cell-variable order (r,u,v,w,p,A,B,C) under the dummy-argument names of cmp_mag_flx; the variable permutation
// lp1,lp2,lor,bp1,bp2,bor of the three calls in mag_unsplit :147-240 -- emfz (2,3,4,6,7,8), emfy (4,2,3,8,6,7),
// emfx (3,4,2,7,8,6) -- is applied with selects.  dir = direction of the edge (0: emfx, 1: emfy, 2: emfz).
template <int R2D>
__device__ __forceinline__ double emf_corners(const MPhys& P, const double* RT, const double* RB, const double* LT, const double* LB, int dir) {
  real qLL[8], qRL[8], qLR[8], qRR[8];   // :1506-1541 (qLL<-qRT, qRL<-qLT, qLR<-qRB, qRR<-qLB)
#define VP1(s) sel3(dir, s[2], s[3], s[1])
#define VP2(s) sel3(dir, s[3], s[1], s[2])
#define VOR(s) sel3(dir, s[1], s[2], s[3])
#define BP1(s) sel3(dir, s[6], s[7], s[5])
#define BP2(s) sel3(dir, s[7], s[5], s[6])
#define BOR(s) sel3(dir, s[5], s[6], s[7])
  qLL[0] = RT[0]; qRL[0] = LT[0]; qLR[0] = RB[0]; qRR[0] = LB[0];
  qLL[1] = RT[4]; qRL[1] = LT[4]; qLR[1] = RB[4]; qRR[1] = LB[4];
  qLL[2] = VP1(RT); qRL[2] = VP1(LT); qLR[2] = VP1(RB); qRR[2] = VP1(LB);
  qLL[3] = VP2(RT); qRL[3] = VP2(LT); qLR[3] = VP2(RB); qRR[3] = VP2(LB);
  const double b1RT = BP1(RT), b1LT = BP1(LT), b1RB = BP1(RB), b1LB = BP1(LB);
  const double b2RT = BP2(RT), b2LT = BP2(LT), b2RB = BP2(RB), b2LB = BP2(LB);
  qLL[5] = 0.5 * (b1RT + b1LT); qRL[5] = 0.5 * (b1RT + b1LT);
  qLR[5] = 0.5 * (b1RB + b1LB); qRR[5] = 0.5 * (b1RB + b1LB);
  qLL[6] = 0.5 * (b2RT + b2RB); qRL[6] = 0.5 * (b2LT + b2LB);
  qLR[6] = 0.5 * (b2RT + b2RB); qRR[6] = 0.5 * (b2LT + b2LB);
  qLL[4] = VOR(RT); qRL[4] = VOR(LT); qLR[4] = VOR(RB); qRR[4] = VOR(LB);
  qLL[7] = BOR(RT); qRL[7] = BOR(LT); qLR[7] = BOR(RB); qRR[7] = BOR(LB);
#undef VP1
#undef VP2
#undef VOR
#undef BP1
#undef BP2
#undef BOR
  return emf_edge<R2D>(P, qLL, qRL, qLR, qRR).v;
}

template <int R2D, bool SL>
__device__ __forceinline__ void one_edge(const MhdArgs& a, const FaceCtx& k, const int dir) {
  const DenseGeom& g = a.g;
  const long long nc = a.nc, c = k.c;
  const double* TR = a.W + MW_TR * nc;
  const MPhys& P = a.P;
  const bool need = dir == 2 ? (k.ex && k.ey && k.oz) : dir == 1 ? (k.ex && k.oy && k.ez) : (k.ox && k.ey && k.ez);
  if (!need) return;
  double RT[8], RB[8], LT[8], LB[8];
  TC t;
  if (dir == 2) {          // emfz: (qRT(i-1,j-1), qRB(i-1,j), qLT(i,j-1), qLB(i,j)) component 3
    load_tc<SL>(TR, nc, cidx(g, k.xm, k.ym, k.z), t); edge_state<2, +1, +1>(P, t, RT);
    load_tc<SL>(TR, nc, cidx(g, k.xm, k.y, k.z), t); edge_state<2, +1, -1>(P, t, RB);
    load_tc<SL>(TR, nc, cidx(g, k.x, k.ym, k.z), t); edge_state<2, -1, +1>(P, t, LT);
    load_tc<SL>(TR, nc, c, t); edge_state<2, -1, -1>(P, t, LB);
  } else if (dir == 1) {   // emfy: (qRT(i-1,k-1), qLT(i,k-1), qRB(i-1,k), qLB(i,k)) component 2
    load_tc<SL>(TR, nc, cidx(g, k.xm, k.y, k.zm), t); edge_state<1, +1, +1>(P, t, RT);
    load_tc<SL>(TR, nc, cidx(g, k.x, k.y, k.zm), t); edge_state<1, -1, +1>(P, t, RB);   // dummy qRB <- actual qLT
    load_tc<SL>(TR, nc, cidx(g, k.xm, k.y, k.z), t); edge_state<1, +1, -1>(P, t, LT);   // dummy qLT <- actual qRB
    load_tc<SL>(TR, nc, c, t); edge_state<1, -1, -1>(P, t, LB);
  } else {                 // emfx: (qRT(j-1,k-1), qRB(j-1,k), qLT(j,k-1), qLB(j,k)) component 1
    load_tc<SL>(TR, nc, cidx(g, k.x, k.ym, k.zm), t); edge_state<0, +1, +1>(P, t, RT);
    load_tc<SL>(TR, nc, cidx(g, k.x, k.ym, k.z), t); edge_state<0, +1, -1>(P, t, RB);
    load_tc<SL>(TR, nc, cidx(g, k.x, k.y, k.zm), t); edge_state<0, -1, +1>(P, t, LT);
    load_tc<SL>(TR, nc, c, t); edge_state<0, -1, -1>(P, t, LB);
  }
  double* EM = a.W + MW_EM * nc;
  EM[dir * nc + c] = div_rn(emf_corners<R2D>(P, RT, RB, LT, LB, dir) * k.dt, a.dx, k.rdx);
}

template <int R2D, bool SL, int MINB>
__global__ void __launch_bounds__(128, MINB) mhd_emf_kernel(const MhdArgs a) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= a.nc) return;
  const DenseGeom& g = a.g;
  const Cxyz p = cell_xyz(g, c);
  FaceCtx k;
  k.ox = in_range(p.x, g.ox0, g.ox1, g.ncx, g.wrapx, 0, 0); k.oy = in_range(p.y, g.oy0, g.oy1, g.ncy, g.wrapy, 0, 0);
  k.oz = in_range(p.z, g.oz0, g.oz1, g.ncz, g.wrapz, 0, 0);
  k.ex = in_range(p.x, g.ox0, g.ox1, g.ncx, g.wrapx, 0, 1); k.ey = in_range(p.y, g.oy0, g.oy1, g.ncy, g.wrapy, 0, 1);
  k.ez = in_range(p.z, g.oz0, g.oz1, g.ncz, g.wrapz, 0, 1);
  k.x = p.x; k.y = p.y; k.z = p.z;
  k.xm = wm(p.x, g.ncx, g.wrapx); k.ym = wm(p.y, g.ncy, g.wrapy); k.zm = wm(p.z, g.ncz, g.wrapz);
  k.c = c; k.dt = *a.dt_dev; k.rdx = rcp_rn(a.dx);
  if (R2D == MHD2D_ROE) {
#pragma unroll 1
    for (int dir = 2; dir >= 0; dir--) one_edge<R2D, SL>(a, k, dir);
  } else {
    one_edge<R2D, SL>(a, k, 2); one_edge<R2D, SL>(a, k, 1); one_edge<R2D, SL>(a, k, 0);
  }
}

#ifdef MHD_DEFINE_KERNELS
// ---------------------------------------------------------------------------------------------------- pass 6
__global__ void __launch_bounds__(256) mhd_update_kernel(const MhdArgs a) {
  const DenseGeom& g = a.g;
  const long long nc = a.nc;
  const long long nx = g.ox1 - g.ox0, ny = g.oy1 - g.oy0, nz = g.oz1 - g.oz0;
  const long long nown = nx * ny * nz;
  const double* F = a.W + MW_F * nc;
  const double* EM = a.W + MW_EM * nc;
  const long long vs = 8 * g.nslot;
  double my_dt = 1e300, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nown; i += (long long)gridDim.x * blockDim.x) {
    const int x = g.ox0 + (int)(i % nx), y = g.oy0 + (int)((i / nx) % ny), z = g.oz0 + (int)(i / (nx * ny));
    const int xp = wp(x, g.ncx, g.wrapx), yp = wp(y, g.ncy, g.wrapy), zp = wp(z, g.ncz, g.wrapz);
    const long long c = cidx(g, x, y, z), cx = cidx(g, xp, y, z), cy = cidx(g, x, yp, z), cz = cidx(g, x, y, zp);
    const long long off = cell_offset<3>(g, x, y, z);
    double u[11];
#pragma unroll
    for (int n = 0; n < 11; n++) u[n] = __ldg(a.uin + n * vs + off);   // set_unew: unew = uold
    // conservative update :883-913, idim = 1,2,3
#pragma unroll
    for (int n = 0; n < 5; n++) u[n] = u[n] + (F[(0 + n) * nc + c] - F[(0 + n) * nc + cx]);
#pragma unroll
    for (int n = 0; n < 5; n++) u[n] = u[n] + (F[(5 + n) * nc + c] - F[(5 + n) * nc + cy]);
#pragma unroll
    for (int n = 0; n < 5; n++) u[n] = u[n] + (F[(10 + n) * nc + c] - F[(10 + n) * nc + cz]);
    // constrained transport :943-995; EM[d][cell] = EMF on the low edge along d of the cell
    const double* EX = EM; const double* EY = EM + nc; const double* EZ = EM + 2 * nc;
    const long long cxy = cidx(g, xp, yp, z), cxz = cidx(g, xp, y, zp), cyz = cidx(g, x, yp, zp);
    u[5] = u[5] + ((EY[c] - EY[cz]) - (EZ[c] - EZ[cy]));
    u[8] = u[8] + ((EY[cx] - EY[cxz]) - (EZ[cx] - EZ[cxy]));
    u[6] = u[6] + ((EZ[c] - EZ[cx]) - (EX[c] - EX[cz]));
    u[9] = u[9] + ((EZ[cy] - EZ[cxy]) - (EX[cy] - EX[cyz]));
    u[7] = u[7] + ((EX[c] - EX[cy]) - (EY[c] - EY[cx]));
    u[10] = u[10] + ((EX[cz] - EX[cyz]) - (EY[cz] - EY[cxz]));
#pragma unroll
    for (int n = 0; n < 11; n++) a.uout[n * vs + off] = u[n];   // set_uold
    // fused courant_fine of the new state (mhd/courant_fine.f90:96-123)
    m0 += u[0]; m1 += u[4];
    double em = 0.0, ei = u[4];
#pragma unroll
    for (int d = 1; d <= 3; d++) {
      const double b2 = 0.125 * SQ(u[4 + d] + u[7 + d]);
      em = em + b2;
      ei = ei - fdiv(0.5 * (u[d] * u[d]), u[0]) - b2;
    }
    m2 += ei; m3 += em;
    real uu[11];
#pragma unroll
    for (int n = 0; n < 11; n++) uu[n] = u[n];
    const double dtc = mhd_cmpdt_cell(a.P, uu, a.dx).v;
    my_dt = dtc < my_dt ? dtc : my_dt;
  }
  __shared__ double red[5][32];
  my_dt = warp_min(my_dt); m0 = warp_sum(m0); m1 = warp_sum(m1); m2 = warp_sum(m2); m3 = warp_sum(m3);
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][wi] = my_dt; red[1][wi] = m0; red[2][wi] = m1; red[3][wi] = m2; red[4][wi] = m3; }
  __syncthreads();
  if (wi == 0) {
    const int nw = blockDim.x >> 5;
    double v0 = l < nw ? red[0][l] : 1e300, v1 = l < nw ? red[1][l] : 0, v2 = l < nw ? red[2][l] : 0, v3 = l < nw ? red[3][l] : 0,
           v4 = l < nw ? red[4][l] : 0;
    v0 = warp_min(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
    if (l == 0) {
      const size_t nb = gridDim.x;
      a.part[0 * nb + blockIdx.x] = v0; a.part[1 * nb + blockIdx.x] = v1; a.part[2 * nb + blockIdx.x] = v2;
      a.part[3 * nb + blockIdx.x] = v3; a.part[4 * nb + blockIdx.x] = v4;
    }
  }
}

// stand-alone courant_fine scan of the owned cells (mhd/courant_fine.f90:1)
__global__ void __launch_bounds__(256) mhd_courant_kernel(const double* __restrict__ uin, DenseGeom g, MPhys P, double dx, double* __restrict__ part) {
  const long long nx = g.ox1 - g.ox0, ny = g.oy1 - g.oy0, nz = g.oz1 - g.oz0;
  const long long nown = nx * ny * nz;
  const long long vs = 8 * g.nslot;
  double my_dt = 1e300, m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nown; i += (long long)gridDim.x * blockDim.x) {
    const int x = g.ox0 + (int)(i % nx), y = g.oy0 + (int)((i / nx) % ny), z = g.oz0 + (int)(i / (nx * ny));
    const long long off = cell_offset<3>(g, x, y, z);
    double u[11];
#pragma unroll
    for (int n = 0; n < 11; n++) u[n] = __ldg(uin + n * vs + off);
    m0 += u[0]; m1 += u[4];
    double em = 0.0, ei = u[4];
#pragma unroll
    for (int d = 1; d <= 3; d++) {
      const double b2 = 0.125 * SQ(u[4 + d] + u[7 + d]);
      em = em + b2;
      ei = ei - fdiv(0.5 * (u[d] * u[d]), u[0]) - b2;
    }
    m2 += ei; m3 += em;
    real uu[11];
#pragma unroll
    for (int n = 0; n < 11; n++) uu[n] = u[n];
    const double dtc = mhd_cmpdt_cell(P, uu, dx).v;
    my_dt = dtc < my_dt ? dtc : my_dt;
  }
  __shared__ double red[5][32];
  my_dt = warp_min(my_dt); m0 = warp_sum(m0); m1 = warp_sum(m1); m2 = warp_sum(m2); m3 = warp_sum(m3);
  const int wi = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][wi] = my_dt; red[1][wi] = m0; red[2][wi] = m1; red[3][wi] = m2; red[4][wi] = m3; }
  __syncthreads();
  if (wi == 0) {
    const int nw = blockDim.x >> 5;
    double v0 = l < nw ? red[0][l] : 1e300, v1 = l < nw ? red[1][l] : 0, v2 = l < nw ? red[2][l] : 0, v3 = l < nw ? red[3][l] : 0,
           v4 = l < nw ? red[4][l] : 0;
    v0 = warp_min(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3); v4 = warp_sum(v4);
    if (l == 0) {
      const size_t nb = gridDim.x;
      part[0 * nb + blockIdx.x] = v0; part[1 * nb + blockIdx.x] = v1; part[2 * nb + blockIdx.x] = v2; part[3 * nb + blockIdx.x] = v3;
      part[4 * nb + blockIdx.x] = v4;
    }
  }
}

#endif  // MHD_DEFINE_KERNELS (pass 6, courant)
// make_boundary_hydro for one boundary region (mhd/hydro_boundary.f90:1; reflexive :141-222, zero gradient :223-296)
struct MhdBoundArgs {
  int n; const int* slots; long long nslot; long long nbr_off;
  int ind_ref[8], ind_normal[8], alt[8];
  double gs[3];
  int kind, gdim /*1..3*/, iperp1 /*0-based variable*/;
  double smallr;
};
#ifdef MHD_DEFINE_KERNELS
__global__ void mhd_boundary_kernel(double* __restrict__ u, const MhdBoundArgs b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n * 8) return;
  const int o = i / 8, ind = i % 8;
  const long long s = b.slots[o], sr = s + b.nbr_off;
  const int indr = b.ind_ref[ind] - 1;
  const int vl = 4 + b.gdim, vr = 7 + b.gdim;   // 0-based: left / right face of the normal field component
  double uu[11];
#pragma unroll
  for (int iv = 0; iv < 11; iv++) uu[iv] = u[((size_t)iv * 8 + indr) * b.nslot + sr];
  double o_[11];
  if (b.kind == 0) {
    double emag = 0.125 * (SQ(uu[5] + uu[8]) + SQ(uu[6] + uu[9]) + SQ(uu[7] + uu[10]));
    uu[4] = uu[4] - emag;
    const double B_normal = u[((size_t)b.iperp1 * 8 + (b.ind_normal[ind] - 1)) * b.nslot + sr];
#pragma unroll
    for (int iv = 0; iv < 11; iv++) {
      double sw = 1.0;
      if (iv >= 1 && iv <= 3) sw = b.gs[iv - 1];
      o_[iv] = uu[iv] * sw;
    }
    o_[vl] = 2 * B_normal - uu[vr];
    o_[vr] = 2 * B_normal - uu[vl];
    emag = 0.125 * (SQ(o_[5] + o_[8]) + SQ(o_[6] + o_[9]) + SQ(o_[7] + o_[10]));
    o_[4] = o_[4] + emag;
  } else {
    double emag = 0.125 * (SQ(uu[5] + uu[8]) + SQ(uu[6] + uu[9]) + SQ(uu[7] + uu[10]));
    double ekin = 0.0, d = fmx(uu[0], b.smallr);
    for (int idim = 1; idim <= 3; idim++) { const double v = uu[idim] / d; ekin = ekin + 0.5 * d * (v * v); }
    uu[4] = uu[4] - emag - ekin;
#pragma unroll
    for (int iv = 0; iv < 11; iv++) o_[iv] = uu[iv];
    const double alt = (double)b.alt[ind];
    o_[vl] = uu[vl] + (uu[vr] - uu[vl]) * alt;
    o_[vr] = uu[vr] + (uu[vr] - uu[vl]) * alt;
    emag = 0.125 * (SQ(o_[5] + o_[8]) + SQ(o_[6] + o_[9]) + SQ(o_[7] + o_[10]));
    ekin = 0.0; d = fmx(o_[0], b.smallr);
    for (int idim = 1; idim <= 3; idim++) { const double v = o_[idim] / d; ekin = ekin + 0.5 * d * (v * v); }
    o_[4] = o_[4] + emag + ekin;
  }
#pragma unroll
  for (int iv = 0; iv < 11; iv++) u[((size_t)iv * 8 + ind) * b.nslot + s] = o_[iv];
}

#endif  // MHD_DEFINE_KERNELS (boundary)

// launchers (mhd_inst_*.cu)
cudaError_t launch_mhd_sweep(const MhdArgs& a, int r1d, int r2d, bool has_slope, int nb_update, cudaStream_t st);
cudaError_t launch_mhd_courant(const double* u, const DenseGeom& g, const MPhys& P, double dx, double* part, int nb, cudaStream_t st);
cudaError_t launch_mhd_boundary(double* u, const MhdBoundArgs& b, cudaStream_t st);

}  // namespace rgpu
