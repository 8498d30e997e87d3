// mhd_amr.cuh -- ideal MHD on adaptively refined meshes in one and two dimensions (AMR mode, mirrored arrays).
//
// The reference's MHD test problems are a 1-D AMR shock tube (tests/mhd/imhd-tube) and the 2-D AMR Orszag-Tang vortex
// (tests/mhd/orszag-tang).  These kernels reproduce godfine1 of the MHD build (mhd/godunov_fine.f90:538-1459) for them:
//   gather of the 6^ndim patch with the divergence-free prolongation of missing neighbour octs (interpol_hydro
//   mhd/interpol_hydro.f90:612 + interpol_mag :990 = interpol_faces :1052, copy_from_refined_faces :1246,
//   cmp_central_faces :1354) -> mag_unsplit (mhd/umuscl.f90:31: ctoprim :2029, uslope :2187, trace1d :244 / trace2d :410,
//   cmpflxm :1308, cmp_mag_flx :1453) -> reset of fluxes / EMFs at refined faces and edges (:742-837) -> conservative update
//   of the Euler system and constrained-transport update of the face fields (:883-976) -> storage of the outer-face fluxes and
//   corner EMFs for the coarse refluxing (:997-1270), which separate kernels apply in the reference's accumulation order.
// upload_fine (mhd/interpol_hydro.f90:5-231, upl :233, upl_left :516, upl_right :564), courant_fine / cmpdt
// (mhd/godunov_utils.f90:5, ctot summed over 1..ndim) and make_boundary_hydro (mhd/hydro_boundary.f90, 1-D) complete the level step.
// State layout: uold/unew(1:ncell,1:nvar+3), variables 6:8 = left-face B, nvar+1:nvar+3 = right-face B (0-based 5..7, 8..10).
#pragma once
#include "amr_kernels.cuh"
#include "mhd_dense.cuh"

namespace rgpu {

constexpr int MNV = 8, MNVS = 11;

struct MhdAmrArgs {
  AmrTree t;
  const int* active;
  int nact, ilevel;
  const double* uold;    // [11][ncell]
  double* unew;
  double* rflux;         // [nact][2*ndim sides][2^(ndim-1) faces][8]: scaled, masked fluxes through the oct's outer faces
  double* remf;          // 2-D: [nact][4]: scaled, masked E_z at the corners (1,1), (1,3), (3,3), (3,1) of the oct
  MPhys P;
  double dt, dx;
  const double* dt_dev;
  int interpol_type, interpol_mag_type;
  int riemann, riemann2d;
};

// run-time dispatch of the 1-D and 2-D Riemann solvers (one out-of-line copy of each: these kernels are latency-, not
// throughput-critical -- the levels they serve hold 10^2..10^5 octs)
__device__ __noinline__ void mhd_riemann1d_rt(int r1d, const MPhys& M, real* ql, real* qr, real* fg) {
  switch (r1d) {
    case MHD_ROE: riemann1d<MHD_ROE>(M, ql, qr, fg); break;
    case MHD_LLF: case MHD_UPWIND: riemann1d<MHD_LLF>(M, ql, qr, fg); break;
    case MHD_HLL: riemann1d<MHD_HLL>(M, ql, qr, fg); break;
    case MHD_HLLD: riemann1d<MHD_HLLD>(M, ql, qr, fg); break;
    default: riemann1d<MHD_HYDRO>(M, ql, qr, fg); break;
  }
}
__device__ __noinline__ double mhd_emfz_rt(int r2d, const MPhys& M, const double* RT, const double* RB, const double* LT, const double* LB) {
  switch (r2d) {
    case MHD2D_LLF: return emf_corners<MHD2D_LLF>(M, RT, RB, LT, LB, 2);
    case MHD2D_ROE: return emf_corners<MHD2D_ROE>(M, RT, RB, LT, LB, 2);
    case MHD2D_UPWIND: return emf_corners<MHD2D_UPWIND>(M, RT, RB, LT, LB, 2);
    case MHD2D_HLL: return emf_corners<MHD2D_HLL>(M, RT, RB, LT, LB, 2);
    case MHD2D_HLLA: return emf_corners<MHD2D_HLLA>(M, RT, RB, LT, LB, 2);
    default: return emf_corners<MHD2D_HLLD>(M, RT, RB, LT, LB, 2);
  }
}
__device__ __forceinline__ double mslope(double st, double ql, double qc, double qr) { return slope_mm(real(st), real(ql), real(qc), real(qr)).v; }
// compute_1d_tvd (mhd/interpol_hydro.f90:1532)
__device__ __forceinline__ double mhd_tvd1(int mt, double b0, double b1, double b2) {
  if (mt == 3) { const double dlft = 0.5 * (b0 - b1), drgt = 0.5 * (b2 - b0); return dlft + drgt; }
  return mslope((double)mt, b1, b0, b2);
}
// ctoprim of one cell (mhd/umuscl.f90:2029): u[11] -> q[8] = (rho, u, v, w, P, A, B, C) with cell-centred B = mean of the faces
__device__ __forceinline__ void mhd_ctoprim_cell(const MPhys& P, const double* u, double* q) {
  q[0] = fmx(u[0], P.smallr);
  q[1] = u[1] / q[0]; q[2] = u[2] / q[0]; q[3] = u[3] / q[0];
  q[5] = (u[5] + u[MNV + 0]) * 0.5;
  q[6] = (u[6] + u[MNV + 1]) * 0.5;
  q[7] = (u[7] + u[MNV + 2]) * 0.5;
  const double eken = 0.5 * (q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double emag = 0.5 * (q[5] * q[5] + q[6] * q[6] + q[7] * q[7]);
  const double etot = u[4] - emag - 0.0;
  const double eint = etot / q[0] - eken;
  q[4] = fmx((P.gamma - 1.0) * q[0] * eint, P.smallp);
}

// ====================================================================================================== NDIM = 1
// interpol_hydro + interpol_mag for one father cell (mhd/interpol_hydro.f90:612,990): u2[2][11]
__device__ void mhd1_interpol_cell(const AmrTree& t, const double* __restrict__ uold, int ind_cell, int ilevel, int interpol_type,
                                   double smallr, double (*u2)[MNVS]) {
  const size_t NC = (size_t)t.ncell;
  auto UO = [&](int ic, int iv) -> double { return uold[(size_t)iv * NC + ic - 1]; };
  int fa[7];
  amr_getnborfather<1>(t, ind_cell, ilevel, fa);
  double u1[7][MNVS];
  for (int j = 0; j < 3; j++)
    for (int iv = 0; iv < MNVS; iv++) u1[j][iv] = UO(fa[j], iv);
  amr_interpol_hydro<1, MNVS>(u1, interpol_type, 0, smallr, u2);
  for (int ind = 0; ind < 2; ind++) { u2[ind][MNV + 1] = u2[ind][6]; u2[ind][MNV + 2] = u2[ind][7]; }   // :712-718
  double um1 = u1[0][5] + 0.5 * 0.0 * (0.0 - 0.5) + 0.5 * 0.0 * (0.0 - 0.5);                           // interpol_faces :1091
  double up1 = u1[0][MNV + 0] + 0.5 * 0.0 * (0.0 - 0.5) + 0.5 * 0.0 * (0.0 - 0.5);
  const int s1 = t.son[fa[1]], s2 = t.son[fa[2]];
  if (s1 > 0) um1 = UO(amr_cell(t, 1, s1), MNV + 0);   // right face of the right son of the left neighbour :1246
  if (s2 > 0) up1 = UO(amr_cell(t, 0, s2), 5);         // left face of the left son of the right neighbour
  const double u0 = 0.5 * (um1 + up1);                 // cmp_central_faces :1354
  u2[0][5] = um1; u2[0][MNV + 0] = u0;
  u2[1][5] = u0;  u2[1][MNV + 0] = up1;
}

// one thread per oct: the 6-cell patch lives in local memory
__global__ void mhd_amr1_godfine_kernel(const MhdAmrArgs a) {
  const int io = blockIdx.x * blockDim.x + threadIdx.x;
  if (io >= a.nact) return;
  const AmrTree& t = a.t;
  const MPhys& P = a.P;
  const size_t NC = (size_t)t.ncell;
  const int igrid = a.active[io];
  const double dt = a.dt_dev ? *a.dt_dev : a.dt, dx = a.dx;
  double uloc[6][MNVS];
  bool ok[6];
  {
    int nfc[27], ng[8];
    amr_get3cubefather<1>(t, igrid, a.ilevel, nfc, ng);
    for (int i1 = 0; i1 <= 2; i1++) {
      const int gn = nfc[i1] > 0 ? t.son[nfc[i1]] : 0;
      double u2[2][MNVS];
      if (gn <= 0) mhd1_interpol_cell(t, a.uold, nfc[i1], a.ilevel, a.interpol_type, P.smallr, u2);
      for (int i2 = 0; i2 <= 1; i2++) {
        const int x = 2 * i1 + i2;   // Fortran i3 = 1+2*(i1-1)+i2 = -1..4 -> 0..5
        if (gn > 0) {
          const int ic = amr_cell(t, i2, gn);
          for (int iv = 0; iv < MNVS; iv++) uloc[x][iv] = a.uold[(size_t)iv * NC + ic - 1];
          ok[x] = t.son[ic] > 0;
        } else {
          for (int iv = 0; iv < MNVS; iv++) uloc[x][iv] = u2[i2][iv];
          ok[x] = false;
        }
      }
    }
  }
  // ---- mag_unsplit, NDIM = 1 ----
  double q[6][MNV], qm[6][MNV], qp[6][MNV];
  for (int i = 0; i < 6; i++) mhd_ctoprim_cell(P, uloc[i], q[i]);
  const double dtdx = dt / dx;
  for (int i = 1; i <= 4; i++) {   // uslope :2222-2245 + trace1d :244 (Fortran cells 0..3)
    double dq[MNV];
    for (int n = 0; n < MNV; n++)
      dq[n] = (P.slope_type == 1 || P.slope_type == 2) ? mslope((double)P.slope_type, q[i - 1][n], q[i][n], q[i + 1][n]) : 0.0;
    double r = q[i][0], u = q[i][1], v = q[i][2], ww = q[i][3], pp = q[i][4], A = q[i][5], B = q[i][6], C = q[i][7];
    const double drx = 0.5 * dq[0], dux = 0.5 * dq[1], dvx = 0.5 * dq[2], dwx = 0.5 * dq[3], dpx = 0.5 * dq[4];
    const double dBx = 0.5 * dq[6], dCx = 0.5 * dq[7];
    const double sr0 = -u * drx - r * dux;
    const double su0 = -u * dux - (dpx + B * dBx + C * dCx) / r;
    const double sv0 = -u * dvx + (A * dBx) / r;
    const double sw0 = -u * dwx + (A * dCx) / r;
    const double sp0 = -u * dpx - P.gamma * pp * dux;
    const double sB0 = -u * dBx + A * dvx - B * dux;
    const double sC0 = -u * dCx + A * dwx - C * dux;
    r = r + sr0 * dtdx; u = u + su0 * dtdx; v = v + sv0 * dtdx; ww = ww + sw0 * dtdx; pp = pp + sp0 * dtdx;
    B = B + sB0 * dtdx; C = C + sC0 * dtdx;
    qp[i][0] = r - drx; qp[i][1] = u - dux; qp[i][2] = v - dvx; qp[i][3] = ww - dwx; qp[i][4] = pp - dpx;
    qp[i][5] = A; qp[i][6] = B - dBx; qp[i][7] = C - dCx;
    if (qp[i][0] < P.smallr) qp[i][0] = r;
    qm[i][0] = r + drx; qm[i][1] = u + dux; qm[i][2] = v + dvx; qm[i][3] = ww + dwx; qm[i][4] = pp + dpx;
    qm[i][5] = A; qm[i][6] = B + dBx; qm[i][7] = C + dCx;
    if (qm[i][0] < P.smallr) qm[i][0] = r;
  }
  double flux[3][MNV];
  for (int i3 = 1; i3 <= 3; i3++) {   // cmpflxm(..., 2,3,4,6,7,8)
    const double* m_ = qm[i3];
    const double* p_ = qp[i3 + 1];
    real ql[8], qr[8], fg[9];
    const double bn_mean = 0.5 * (m_[5] + p_[5]);
    ql[0] = m_[0]; ql[1] = m_[4]; ql[2] = m_[1]; ql[3] = bn_mean; ql[4] = m_[2]; ql[5] = m_[6]; ql[6] = m_[3]; ql[7] = m_[7];
    qr[0] = p_[0]; qr[1] = p_[4]; qr[2] = p_[1]; qr[3] = bn_mean; qr[4] = p_[2]; qr[5] = p_[6]; qr[6] = p_[3]; qr[7] = p_[7];
    mhd_riemann1d_rt(a.riemann, P, ql, qr, fg);
    double* f = flux[i3 - 1];
    f[0] = fg[0].v; f[4] = fg[1].v; f[1] = fg[2].v; f[5] = fg[3].v; f[2] = fg[4].v; f[6] = fg[5].v; f[3] = fg[6].v; f[7] = fg[7].v;
    for (int n = 0; n < MNV; n++) f[n] = f[n] * dt / dx;
    if (ok[i3] || ok[i3 + 1])         // :742-747
      for (int n = 0; n < MNV; n++) f[n] = 0.0;
    f[5] = 0.0;                       // flux(:,6,idim) = 0 :778
  }
  for (int i2 = 0; i2 <= 1; i2++) {   // :883-956 (emfy = emfz = 0 in one dimension)
    const int ic = amr_cell(t, i2, igrid);
    for (int iv = 0; iv < MNV; iv++) {
      double u = a.unew[(size_t)iv * NC + ic - 1];
      u = u + (flux[i2][iv] - flux[i2 + 1][iv]);
      if (iv == 5) u = u + ((0.0 - 0.0) - (0.0 - 0.0));
      a.unew[(size_t)iv * NC + ic - 1] = u;
    }
    for (int iv = 0; iv < 3; iv++) {
      double u = a.unew[(size_t)(MNV + iv) * NC + ic - 1];
      u = u + (flux[i2][5 + iv] - flux[i2 + 1][5 + iv]);
      if (iv == 0) u = u + ((0.0 - 0.0) - (0.0 - 0.0));
      a.unew[(size_t)(MNV + iv) * NC + ic - 1] = u;
    }
  }
  for (int n = 0; n < MNV; n++) {
    a.rflux[((size_t)io * 2 + 0) * MNV + n] = flux[0][n];
    a.rflux[((size_t)io * 2 + 1) * MNV + n] = flux[2][n];
  }
}

// ====================================================================================================== NDIM = 2
// interpol_hydro + interpol_mag for one father cell, NDIM = 2: u2[4][11]
__device__ void mhd2_interpol_cell(const AmrTree& t, const double* __restrict__ uold, int ind_cell, int ilevel, int interpol_type,
                                   int mt, double smallr, double (*u2)[MNVS]) {
  const size_t NC = (size_t)t.ncell;
  auto UO = [&](int ic, int iv) -> double { return uold[(size_t)iv * NC + ic - 1]; };
  int fa[7], ind1[5];
  amr_getnborfather<2>(t, ind_cell, ilevel, fa);
  double u1[7][MNVS];
  for (int j = 0; j < 5; j++) {
    for (int iv = 0; iv < MNVS; iv++) u1[j][iv] = UO(fa[j], iv);
    ind1[j] = t.son[fa[j]];
  }
  amr_interpol_hydro<2, MNVS>(u1, interpol_type, 0, smallr, u2);   // variables 1..5 and 8 (cell centred); the rest is overwritten
  for (int ind = 0; ind < 4; ind++) u2[ind][MNV + 2] = u2[ind][7];  // :712-718
  double u[3][2], v[2][3];
  // B1(j,c): c = 1..3 -> variables 6..8 (0-based 5..7), c = 4..6 -> nvar+1..nvar+3 (0-based 8..10)
#define B1(j_, c_) u1[(j_)][((c_) <= 3 ? 4 + (c_) : MNV + (c_)-4)]
  {   // interpol_faces :1052
    double s1;
    const double s2 = 0.0;
    s1 = 0.0; if (mt > 0) s1 = mhd_tvd1(mt, B1(0, 1), B1(3, 1), B1(4, 1));
    for (int j = 0; j <= 1; j++) u[0][j] = B1(0, 1) + 0.5 * s1 * ((double)j - 0.5) + 0.5 * s2 * ((double)0 - 0.5);
    s1 = 0.0; if (mt > 0) s1 = mhd_tvd1(mt, B1(0, 4), B1(3, 4), B1(4, 4));
    for (int j = 0; j <= 1; j++) u[2][j] = B1(0, 4) + 0.5 * s1 * ((double)j - 0.5) + 0.5 * s2 * ((double)0 - 0.5);
    s1 = 0.0; if (mt > 0) s1 = mhd_tvd1(mt, B1(0, 2), B1(1, 2), B1(2, 2));
    for (int i = 0; i <= 1; i++) v[i][0] = B1(0, 2) + 0.5 * s1 * ((double)i - 0.5) + 0.5 * s2 * ((double)0 - 0.5);
    s1 = 0.0; if (mt > 0) s1 = mhd_tvd1(mt, B1(0, 5), B1(1, 5), B1(2, 5));
    for (int i = 0; i <= 1; i++) v[i][2] = B1(0, 5) + 0.5 * s1 * ((double)i - 0.5) + 0.5 * s2 * ((double)0 - 0.5);
  }
#undef B1
  for (int j = 0; j <= 1; j++) {   // copy_from_refined_faces :1246
    if (ind1[1] > 0) u[0][j] = UO(amr_cell(t, 1 + j * 2, ind1[1]), MNV + 0);
    if (ind1[2] > 0) u[2][j] = UO(amr_cell(t, 0 + j * 2, ind1[2]), 5);
  }
  for (int i = 0; i <= 1; i++) {
    if (ind1[3] > 0) v[i][0] = UO(amr_cell(t, i + 1 * 2, ind1[3]), MNV + 1);
    if (ind1[4] > 0) v[i][2] = UO(amr_cell(t, i + 0 * 2, ind1[4]), 6);
  }
  double UXX = 0.0, VYY = 0.0;     // cmp_central_faces :1354, NDIM == 2
  for (int i = 0; i <= 1; i++)
    for (int j = 0; j <= 1; j++) {
      const int ii = 2 * i - 1, jj = 2 * j - 1;
      UXX = UXX + ((double)(ii * jj) * v[i][jj + 1]) * 0.25;
      VYY = VYY + ((double)(ii * jj) * u[ii + 1][j]) * 0.25;
    }
  for (int j = 0; j <= 1; j++) u[1][j] = 0.5 * (u[0][j] + u[2][j]) + UXX;
  for (int i = 0; i <= 1; i++) v[i][1] = 0.5 * (v[i][0] + v[i][2]) + VYY;
  for (int i = 0; i <= 1; i++)
    for (int j = 0; j <= 1; j++) {
      const int ind = i + 2 * j;
      u2[ind][5] = u[i][j];
      u2[ind][6] = v[i][j];
      u2[ind][MNV + 0] = u[i + 1][j];
      u2[ind][MNV + 1] = v[i][j + 1];
    }
}

constexpr int MHD2_TPO = 64;   // threads per oct

// one block of 64 threads per oct; the 6x6 patch and every intermediate of mag_unsplit in shared memory
__global__ void __launch_bounds__(MHD2_TPO) mhd_amr2_godfine_kernel(const MhdAmrArgs a) {
  struct Sm {
    double uloc[36][MNVS];
    unsigned char ok[36];
    int nfc[27], gnb[9], ng[8];
    double q[36][MNV];
    double bf[49][2];        // [j+1][i+1], Fortran -1..5
    double dq[36][MNV][2];
    double dbf[49][2];
    double Ez[36];           // corners 0..4
    double qm[36][MNV][2], qp[36][MNV][2];
    double qRT[36][MNV], qRB[36][MNV], qLT[36][MNV], qLB[36][MNV];
    double flux[2][3][3][MNV];
    double emfz[3][3];
  };
  __shared__ Sm s;
  const int tl = threadIdx.x, io = blockIdx.x;
  if (io >= a.nact) return;
  const AmrTree& t = a.t;
  const MPhys& P = a.P;
  const size_t NC = (size_t)t.ncell;
  const int igrid = a.active[io];
  const double dt = a.dt_dev ? *a.dt_dev : a.dt, dx = a.dx;
  const double smallr = P.smallr, smallp = P.smallp, gamma = P.gamma;
#define X_(i) ((i) + 1)
#define C36(j, i) (X_(j) * 6 + X_(i))
#define C49(j, i) (X_(j) * 7 + X_(i))
  if (tl == 0) {
    amr_get3cubefather<2>(t, igrid, a.ilevel, s.nfc, s.ng);
    for (int j = 0; j < 9; j++) s.gnb[j] = s.nfc[j] > 0 ? t.son[s.nfc[j]] : 0;
  }
  __syncthreads();
  // ---- gather (mhd/godunov_fine.f90:600-700) ----
  for (int e = tl; e < 9 * 4; e += MHD2_TPO) {
    const int jf = e / 4, is = e % 4;
    const int g = s.gnb[jf];
    if (g <= 0) continue;
    const int i1 = jf % 3, j1 = jf / 3;
    const int i3 = 1 + 2 * (i1 - 1) + (is & 1), j3 = 1 + 2 * (j1 - 1) + (is >> 1);
    const int ic = amr_cell(t, is, g);
    for (int iv = 0; iv < MNVS; iv++) s.uloc[C36(j3, i3)][iv] = a.uold[(size_t)iv * NC + ic - 1];
    s.ok[C36(j3, i3)] = t.son[ic] > 0;
  }
  for (int jf = tl; jf < 9; jf += MHD2_TPO) {
    if (s.gnb[jf] > 0) continue;
    double u2[4][MNVS];
    mhd2_interpol_cell(t, a.uold, s.nfc[jf], a.ilevel, a.interpol_type, a.interpol_mag_type, smallr, u2);
    const int i1 = jf % 3, j1 = jf / 3;
    for (int is = 0; is < 4; is++) {
      const int i3 = 1 + 2 * (i1 - 1) + (is & 1), j3 = 1 + 2 * (j1 - 1) + (is >> 1);
      for (int iv = 0; iv < MNVS; iv++) s.uloc[C36(j3, i3)][iv] = u2[is][iv];
      s.ok[C36(j3, i3)] = 0;
    }
  }
  __syncthreads();
  // ---- ctoprim :2029 ----
  for (int e = tl; e < 36; e += MHD2_TPO) mhd_ctoprim_cell(P, s.uloc[e], s.q[e]);
  for (int e = tl; e < 6 * 7; e += MHD2_TPO) {   // bf(:,:,1): j = -1..4, i = -1..5
    const int j = e / 7 - 1, i = e % 7 - 1;
    s.bf[C49(j, i)][0] = (i <= 4) ? s.uloc[C36(j, i)][5] : s.uloc[C36(j, i - 1)][MNV + 0];
  }
  for (int e = tl; e < 7 * 6; e += MHD2_TPO) {   // bf(:,:,2): j = -1..5, i = -1..4
    const int j = e / 6 - 1, i = e % 6 - 1;
    s.bf[C49(j, i)][1] = (j <= 4) ? s.uloc[C36(j, i)][6] : s.uloc[C36(j - 1, i)][MNV + 1];
  }
  __syncthreads();
  // ---- uslope, NDIM == 2 :2253-2372 ----
  for (int e = tl; e < 16 * MNV; e += MHD2_TPO) {
    const int n = e % MNV, c = e / MNV, i = c % 4, j = c / 4;
    double dx_ = 0.0, dy_ = 0.0;
    if (P.slope_type == 1 || P.slope_type == 2) {
      const double st = (double)P.slope_type;
      dx_ = mslope(st, s.q[C36(j, i - 1)][n], s.q[C36(j, i)][n], s.q[C36(j, i + 1)][n]);
      dy_ = mslope(st, s.q[C36(j - 1, i)][n], s.q[C36(j, i)][n], s.q[C36(j + 1, i)][n]);
    }
    s.dq[C36(j, i)][n][0] = dx_;
    s.dq[C36(j, i)][n][1] = dy_;
  }
  for (int e = tl; e < 49; e += MHD2_TPO) { s.dbf[e][0] = 0.0; s.dbf[e][1] = 0.0; }
  __syncthreads();
  if (P.slope_mag_type == 1 || P.slope_mag_type == 2) {
    const double st = (double)P.slope_mag_type;
    for (int e = tl; e < 4 * 5; e += MHD2_TPO) {   // j = 0..3, i = 0..4
      const int j = e / 5, i = e % 5;
      s.dbf[C49(j, i)][0] = mslope(st, s.bf[C49(j - 1, i)][0], s.bf[C49(j, i)][0], s.bf[C49(j + 1, i)][0]);
    }
    for (int e = tl; e < 5 * 4; e += MHD2_TPO) {   // j = 0..4, i = 0..3
      const int j = e / 4, i = e % 4;
      s.dbf[C49(j, i)][1] = mslope(st, s.bf[C49(j, i - 1)][1], s.bf[C49(j, i)][1], s.bf[C49(j, i + 1)][1]);
    }
  }
  // ---- trace2d :410: E_z at the corners i,j = 0..4 ----
  for (int e = tl; e < 25; e += MHD2_TPO) {
    const int j = e / 5, i = e % 5;
    const double u = 0.25 * (s.q[C36(j - 1, i - 1)][1] + s.q[C36(j, i - 1)][1] + s.q[C36(j - 1, i)][1] + s.q[C36(j, i)][1]);
    const double v = 0.25 * (s.q[C36(j - 1, i - 1)][2] + s.q[C36(j, i - 1)][2] + s.q[C36(j - 1, i)][2] + s.q[C36(j, i)][2]);
    const double A = 0.5 * (s.bf[C49(j - 1, i)][0] + s.bf[C49(j, i)][0]);
    const double B = 0.5 * (s.bf[C49(j, i - 1)][1] + s.bf[C49(j, i)][1]);
    s.Ez[C36(j, i)] = u * B - v * A;
  }
  __syncthreads();
  const double dtdx = dt / dx, dtdy = dt / dx;
  for (int e = tl; e < 16; e += MHD2_TPO) {
    const int j = e / 4, i = e % 4;
    const double* qq = s.q[C36(j, i)];
    double(*d)[2] = s.dq[C36(j, i)];
    double r = qq[0], u = qq[1], v = qq[2], ww = qq[3], pp = qq[4], A = qq[5], B = qq[6], C = qq[7];
    double AL = s.bf[C49(j, i)][0], AR = s.bf[C49(j, i + 1)][0], BL = s.bf[C49(j, i)][1], BR = s.bf[C49(j + 1, i)][1];
    const double drx = 0.5 * d[0][0], dux = 0.5 * d[1][0], dvx = 0.5 * d[2][0], dwx = 0.5 * d[3][0], dpx = 0.5 * d[4][0];
    const double dBx = 0.5 * d[6][0], dCx = 0.5 * d[7][0];
    const double dry = 0.5 * d[0][1], duy = 0.5 * d[1][1], dvy = 0.5 * d[2][1], dwy = 0.5 * d[3][1], dpy = 0.5 * d[4][1];
    const double dAy = 0.5 * d[5][1], dCy = 0.5 * d[7][1];
    const double dALy = 0.5 * s.dbf[C49(j, i)][0], dARy = 0.5 * s.dbf[C49(j, i + 1)][0];
    const double dBLx = 0.5 * s.dbf[C49(j, i)][1], dBRx = 0.5 * s.dbf[C49(j + 1, i)][1];
    const double ELL = s.Ez[C36(j, i)], ELR = s.Ez[C36(j + 1, i)], ERL = s.Ez[C36(j, i + 1)], ERR = s.Ez[C36(j + 1, i + 1)];
    const double sAL0 = +(ELR - ELL) * dtdy * 0.5;
    const double sAR0 = +(ERR - ERL) * dtdy * 0.5;
    const double sBL0 = -(ERL - ELL) * dtdx * 0.5;
    const double sBR0 = -(ERR - ELR) * dtdx * 0.5;
    AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0;
    const double sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy;
    const double su0 = (-u * dux - (dpx + B * dBx + C * dCx) / r) * dtdx + (-v * duy + B * dAy / r) * dtdy;
    const double sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - (dpy + A * dAy + C * dCy) / r) * dtdy;
    const double sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy;
    const double sp0 = (-u * dpx - dux * gamma * pp) * dtdx + (-v * dpy - dvy * gamma * pp) * dtdy;
    const double sC0 = (-u * dCx - C * dux + A * dwx) * dtdx + (-v * dCy - C * dvy + B * dwy) * dtdy;
    r = r + sr0; u = u + su0; v = v + sv0; ww = ww + sw0; pp = pp + sp0; C = C + sC0;
    A = 0.5 * (AL + AR); B = 0.5 * (BL + BR);
    auto set2 = [&](double* s_, double R, double U, double V, double W, double P_, double A_, double B_, double C_) {
      s_[0] = R; s_[1] = U; s_[2] = V; s_[3] = W; s_[4] = P_; s_[5] = A_; s_[6] = B_; s_[7] = C_;
      if (s_[0] < smallr) s_[0] = r;
      s_[4] = fmx(smallp, s_[4]);
    };
    double tt[MNV];
    const int c = C36(j, i);
    set2(tt, r - drx, u - dux, v - dvx, ww - dwx, pp - dpx, AL, B - dBx, C - dCx); for (int n = 0; n < MNV; n++) s.qp[c][n][0] = tt[n];
    set2(tt, r + drx, u + dux, v + dvx, ww + dwx, pp + dpx, AR, B + dBx, C + dCx); for (int n = 0; n < MNV; n++) s.qm[c][n][0] = tt[n];
    set2(tt, r - dry, u - duy, v - dvy, ww - dwy, pp - dpy, A - dAy, BL, C - dCy); for (int n = 0; n < MNV; n++) s.qp[c][n][1] = tt[n];
    set2(tt, r + dry, u + duy, v + dvy, ww + dwy, pp + dpy, A + dAy, BR, C + dCy); for (int n = 0; n < MNV; n++) s.qm[c][n][1] = tt[n];
    set2(s.qRT[c], r + (+drx + dry), u + (+dux + duy), v + (+dvx + dvy), ww + (+dwx + dwy), pp + (+dpx + dpy), AR + (+dARy), BR + (+dBRx), C + (+dCx + dCy));
    set2(s.qRB[c], r + (+drx - dry), u + (+dux - duy), v + (+dvx - dvy), ww + (+dwx - dwy), pp + (+dpx - dpy), AR + (-dARy), BL + (+dBLx), C + (+dCx - dCy));
    set2(s.qLT[c], r + (-drx + dry), u + (-dux + duy), v + (-dvx + dvy), ww + (-dwx + dwy), pp + (-dpx + dpy), AL + (+dALy), BR + (-dBRx), C + (-dCx + dCy));
    set2(s.qLB[c], r + (-drx - dry), u + (-dux - duy), v + (-dvx - dvy), ww + (-dwx - dwy), pp + (-dpx - dpy), AL + (-dALy), BL + (-dBLx), C + (-dCx - dCy));
  }
  __syncthreads();
  // ---- cmpflxm :1308 (12 faces) and cmp_mag_flx :1453 (9 corners), then the resets at refined faces / edges :760-837 ----
  for (int e = tl; e < 21; e += MHD2_TPO) {
    if (e < 12) {
      const int idim = e / 6, f = e % 6;
      const int i0 = idim == 0, j0 = idim == 1;
      const int ni = 2 + i0;
      const int i = 1 + f % ni, j = 1 + f / ni;
      const int ln = idim == 0 ? 1 : 2, lt1 = idim == 0 ? 2 : 1, lt2 = 3, bn = idim == 0 ? 5 : 6, bt1 = idim == 0 ? 6 : 5, bt2 = 7;
      const int cm = C36(j - j0, i - i0), cp = C36(j, i);
      real ql[8], qr[8], fg[9];
      const double bn_mean = 0.5 * (s.qm[cm][bn][idim] + s.qp[cp][bn][idim]);
      ql[0] = s.qm[cm][0][idim]; ql[1] = s.qm[cm][4][idim]; ql[2] = s.qm[cm][ln][idim]; ql[3] = bn_mean;
      ql[4] = s.qm[cm][lt1][idim]; ql[5] = s.qm[cm][bt1][idim]; ql[6] = s.qm[cm][lt2][idim]; ql[7] = s.qm[cm][bt2][idim];
      qr[0] = s.qp[cp][0][idim]; qr[1] = s.qp[cp][4][idim]; qr[2] = s.qp[cp][ln][idim]; qr[3] = bn_mean;
      qr[4] = s.qp[cp][lt1][idim]; qr[5] = s.qp[cp][bt1][idim]; qr[6] = s.qp[cp][lt2][idim]; qr[7] = s.qp[cp][bt2][idim];
      mhd_riemann1d_rt(a.riemann, P, ql, qr, fg);
      double* fl = s.flux[idim][j - 1][i - 1];
      fl[0] = fg[0].v; fl[4] = fg[1].v; fl[ln] = fg[2].v; fl[bn] = fg[3].v; fl[lt1] = fg[4].v; fl[bt1] = fg[5].v; fl[lt2] = fg[6].v; fl[bt2] = fg[7].v;
      for (int n = 0; n < MNV; n++) fl[n] = fl[n] * dt / dx;
      if (s.ok[cm] || s.ok[cp])
        for (int n = 0; n < MNV; n++) fl[n] = 0.0;
      fl[5] = 0.0;
      fl[6] = 0.0;
    } else {
      const int k = e - 12, i = 1 + k % 3, j = 1 + k / 3;
      double ez = mhd_emfz_rt(a.riemann2d, P, s.qRT[C36(j - 1, i - 1)], s.qRB[C36(j, i - 1)], s.qLT[C36(j - 1, i)], s.qLB[C36(j, i)]) * dt / dx;
      if (s.ok[C36(j, i)] || s.ok[C36(j - 1, i)] || s.ok[C36(j, i - 1)] || s.ok[C36(j - 1, i - 1)]) ez = 0.0;
      s.emfz[j - 1][i - 1] = ez;
    }
  }
  __syncthreads();
  // ---- update of the oct's own cells: Euler system x then y (:886-934), then the induction system (:939-976) ----
  for (int e = tl; e < 4 * MNVS; e += MHD2_TPO) {
    const int is = e % 4, iv = e / 4;               // iv 0..7: variables 1..8; 8..10: nvar+1..nvar+3 <- fluxes 6..8
    const int i3 = 1 + (is & 1), j3 = 1 + (is >> 1);
    const int ic = amr_cell(t, is, igrid);
    const int src = iv < MNV ? iv : 5 + (iv - MNV);
    double u = a.unew[(size_t)iv * NC + ic - 1];
    u = u + (s.flux[0][j3 - 1][i3 - 1][src] - s.flux[0][j3 - 1][i3][src]);
    u = u + (s.flux[1][j3 - 1][i3 - 1][src] - s.flux[1][j3][i3 - 1][src]);
    if (iv == 5) u = u + ((0.0 - 0.0) - (s.emfz[j3 - 1][i3 - 1] - s.emfz[j3][i3 - 1]));
    if (iv == MNV + 0) u = u + ((0.0 - 0.0) - (s.emfz[j3 - 1][i3] - s.emfz[j3][i3]));
    if (iv == 6) u = u + ((s.emfz[j3 - 1][i3 - 1] - s.emfz[j3 - 1][i3]) - (0.0 - 0.0));
    if (iv == MNV + 1) u = u + ((s.emfz[j3][i3 - 1] - s.emfz[j3][i3]) - (0.0 - 0.0));
    a.unew[(size_t)iv * NC + ic - 1] = u;
  }
  // ---- outer-face fluxes [side = 2*idim + (0 left | 1 right)][face 0..1][8] and the four corner EMFs ----
  for (int e = tl; e < 4 * 2 * MNV; e += MHD2_TPO) {
    const int n = e % MNV, fs = (e / MNV) % 2, side = e / (2 * MNV);
    const int idim = side / 2, right = side % 2;
    const double v = idim == 0 ? s.flux[0][fs][right ? 2 : 0][n] : s.flux[1][right ? 2 : 0][fs][n];
    a.rflux[(((size_t)io * 4 + side) * 2 + fs) * MNV + n] = v;
  }
  if (tl < 4) {
    const int ci[4] = {1, 1, 3, 3}, cj[4] = {1, 3, 3, 1};   // corners (i3,j3) of the edges X0Y0, X0Y1, X1Y1, X1Y0 (:1176-1270)
    a.remf[(size_t)io * 4 + tl] = s.emfz[cj[tl] - 1][ci[tl] - 1];
  }
#undef X_
#undef C36
#undef C49
}

// ====================================================================================================== NDIM = 3
// compute_2d_tvd, one direction (mhd/interpol_hydro.f90:1478)
__device__ __forceinline__ double mhd_tvd2(int mt, double b0, double bl, double br) {
  if (mt == 3) { const double dlft = 0.5 * (b0 - bl), drgt = 0.5 * (br - b0); return dlft + drgt; }
  return mslope((double)mt, bl, b0, br);
}
// interpol_hydro + interpol_mag for one father cell, NDIM = 3: u2[8][11]
__device__ void mhd3_interpol_cell(const AmrTree& t, const double* __restrict__ uold, int ind_cell, int ilevel, int interpol_type,
                                   int mt, double smallr, double (*u2)[MNVS]) {
  const size_t NC = (size_t)t.ncell;
  auto UO = [&](int ic, int iv) -> double { return uold[(size_t)iv * NC + ic - 1]; };
  int fa[7], ind1[7];
  amr_getnborfather<3>(t, ind_cell, ilevel, fa);
  double u1[7][MNVS];
  for (int j = 0; j < 7; j++) {
    for (int iv = 0; iv < MNVS; iv++) u1[j][iv] = UO(fa[j], iv);
    ind1[j] = t.son[fa[j]];
  }
  amr_interpol_hydro<3, MNVS>(u1, interpol_type, 0, smallr, u2);   // variables 1..5 (cell centred); the face fields are overwritten
  double u[3][2][2], v[2][3][2], w[2][2][3];                       // u[i+1][j][k], v[i][j+1][k], w[i][j][k+1]
#define B1(j_, c_) u1[(j_)][((c_) <= 3 ? 4 + (c_) : MNV + (c_)-4)]
  for (int side = 0; side < 2; side++) {   // interpol_faces :1052
    const int cx = side == 0 ? 1 : 4, cy = side == 0 ? 2 : 5, cz = side == 0 ? 3 : 6, f = side == 0 ? 0 : 2;
    double s1 = 0.0, s2 = 0.0;
    if (mt > 0) { s1 = mhd_tvd2(mt, B1(0, cx), B1(3, cx), B1(4, cx)); s2 = mhd_tvd2(mt, B1(0, cx), B1(5, cx), B1(6, cx)); }
    for (int j = 0; j <= 1; j++)
      for (int k = 0; k <= 1; k++) u[f][j][k] = B1(0, cx) + 0.5 * s1 * ((double)j - 0.5) + 0.5 * s2 * ((double)k - 0.5);
    s1 = s2 = 0.0;
    if (mt > 0) { s1 = mhd_tvd2(mt, B1(0, cy), B1(1, cy), B1(2, cy)); s2 = mhd_tvd2(mt, B1(0, cy), B1(5, cy), B1(6, cy)); }
    for (int i = 0; i <= 1; i++)
      for (int k = 0; k <= 1; k++) v[i][f][k] = B1(0, cy) + 0.5 * s1 * ((double)i - 0.5) + 0.5 * s2 * ((double)k - 0.5);
    s1 = s2 = 0.0;
    if (mt > 0) { s1 = mhd_tvd2(mt, B1(0, cz), B1(1, cz), B1(2, cz)); s2 = mhd_tvd2(mt, B1(0, cz), B1(3, cz), B1(4, cz)); }
    for (int i = 0; i <= 1; i++)
      for (int j = 0; j <= 1; j++) w[i][j][f] = B1(0, cz) + 0.5 * s1 * ((double)i - 0.5) + 0.5 * s2 * ((double)j - 0.5);
  }
#undef B1
  for (int a = 0; a <= 1; a++)             // copy_from_refined_faces :1246
    for (int b = 0; b <= 1; b++) {
      if (ind1[1] > 0) u[0][a][b] = UO(amr_cell(t, 1 + a * 2 + b * 4, ind1[1]), MNV + 0);
      if (ind1[2] > 0) u[2][a][b] = UO(amr_cell(t, 0 + a * 2 + b * 4, ind1[2]), 5);
      if (ind1[3] > 0) v[a][0][b] = UO(amr_cell(t, a + 1 * 2 + b * 4, ind1[3]), MNV + 1);
      if (ind1[4] > 0) v[a][2][b] = UO(amr_cell(t, a + 0 * 2 + b * 4, ind1[4]), 6);
      if (ind1[5] > 0) w[a][b][0] = UO(amr_cell(t, a + b * 2 + 1 * 4, ind1[5]), MNV + 2);
      if (ind1[6] > 0) w[a][b][2] = UO(amr_cell(t, a + b * 2 + 0 * 4, ind1[6]), 7);
    }
  double UXX = 0, VYY = 0, WZZ = 0, UXYZ = 0, VXYZ = 0, WXYZ = 0;   // cmp_central_faces :1354, NDIM == 3
  for (int i = 0; i <= 1; i++)
    for (int j = 0; j <= 1; j++)
      for (int k = 0; k <= 1; k++) {
        const int ii = 2 * i - 1, jj = 2 * j - 1, kk = 2 * k - 1;
        UXX = UXX + ((double)(ii * jj) * v[i][jj + 1][k] + (double)(ii * kk) * w[i][j][kk + 1]) * 0.125;
        VYY = VYY + ((double)(jj * kk) * w[i][j][kk + 1] + (double)(ii * jj) * u[ii + 1][j][k]) * 0.125;
        WZZ = WZZ + ((double)(ii * kk) * u[ii + 1][j][k] + (double)(jj * kk) * v[i][jj + 1][k]) * 0.125;
        UXYZ = UXYZ + ((double)(ii * jj * kk) * u[ii + 1][j][k]) * 0.125;
        VXYZ = VXYZ + ((double)(ii * jj * kk) * v[i][jj + 1][k]) * 0.125;
        WXYZ = WXYZ + ((double)(ii * jj * kk) * w[i][j][kk + 1]) * 0.125;
      }
  for (int j = 0; j <= 1; j++)
    for (int k = 0; k <= 1; k++)
      u[1][j][k] = 0.5 * (u[0][j][k] + u[2][j][k]) + UXX + ((double)k - 0.5) * VXYZ + ((double)j - 0.5) * WXYZ;
  for (int i = 0; i <= 1; i++)
    for (int k = 0; k <= 1; k++)
      v[i][1][k] = 0.5 * (v[i][0][k] + v[i][2][k]) + VYY + ((double)i - 0.5) * WXYZ + ((double)k - 0.5) * UXYZ;
  for (int i = 0; i <= 1; i++)
    for (int j = 0; j <= 1; j++)
      w[i][j][1] = 0.5 * (w[i][j][0] + w[i][j][2]) + WZZ + ((double)j - 0.5) * UXYZ + ((double)i - 0.5) * VXYZ;
  for (int i = 0; i <= 1; i++)
    for (int j = 0; j <= 1; j++)
      for (int k = 0; k <= 1; k++) {
        const int ind = i + 2 * j + 4 * k;
        u2[ind][5] = u[i][j][k]; u2[ind][MNV + 0] = u[i + 1][j][k];
        u2[ind][6] = v[i][j][k]; u2[ind][MNV + 1] = v[i][j + 1][k];
        u2[ind][7] = w[i][j][k]; u2[ind][MNV + 2] = w[i][j][k + 1];
      }
}

__device__ __noinline__ double mhd_emf_rt(int r2d, const MPhys& M, const double* RT, const double* RB, const double* LT, const double* LB, int dir) {
  switch (r2d) {
    case MHD2D_LLF: return emf_corners<MHD2D_LLF>(M, RT, RB, LT, LB, dir);
    case MHD2D_ROE: return emf_corners<MHD2D_ROE>(M, RT, RB, LT, LB, dir);
    case MHD2D_UPWIND: return emf_corners<MHD2D_UPWIND>(M, RT, RB, LT, LB, dir);
    case MHD2D_HLL: return emf_corners<MHD2D_HLL>(M, RT, RB, LT, LB, dir);
    case MHD2D_HLLA: return emf_corners<MHD2D_HLLA>(M, RT, RB, LT, LB, dir);
    default: return emf_corners<MHD2D_HLLD>(M, RT, RB, LT, LB, dir);
  }
}

constexpr int MHD3_TPO = 128;   // threads per oct
// shared-memory image of mag_unsplit on one 6^3 patch (mhd/umuscl.f90:31): ~160 KB, one oct per SM at a time (dynamic shared memory)
struct Mhd3Sm {
  double uloc[216][MNVS];
  double q[216][MNV];
  double bf[343][3];            // [k+1][j+1][i+1], Fortran -1..5
  double dq[64][MNV][3];        // cells 0..3
  double dbf[343][3][2];
  double Ex[216], Ey[216], Ez[216];
  double qm[64][MNV][3], qp[64][MNV][3];
  double qRT[64][MNV][3], qRB[64][MNV][3], qLT[64][MNV][3], qLB[64][MNV][3];
  double flux[3][27][MNV];      // [idim][(k3-1)*9 + (j3-1)*3 + (i3-1)]
  double emf[3][27];            // [dir: x, y, z][(k3-1)*9 + (j3-1)*3 + (i3-1)]
  int nfc[27], gnb[27], ng[8];
  unsigned char ok[216];
};
// one block of 128 threads per oct
__global__ void __launch_bounds__(MHD3_TPO) mhd_amr3_godfine_kernel(const MhdAmrArgs a) {
#ifdef RGPU_HOST_NUMERICS
  double* smem = rgpu_host_dyn_smem;
#else
  extern __shared__ double smem[];
#endif
  Mhd3Sm& s = *reinterpret_cast<Mhd3Sm*>(smem);
  const int tl = threadIdx.x, io = blockIdx.x;
  if (io >= a.nact) return;
  const AmrTree& t = a.t;
  const MPhys& P = a.P;
  const size_t NC = (size_t)t.ncell;
  const int igrid = a.active[io];
  const double dt = a.dt_dev ? *a.dt_dev : a.dt, dx = a.dx;
  const double smallr = P.smallr, smallp = P.smallp, gamma = P.gamma;
#define X_(i) ((i) + 1)
#define C6(k, j, i) ((X_(k) * 6 + X_(j)) * 6 + X_(i))
#define C7(k, j, i) ((X_(k) * 7 + X_(j)) * 7 + X_(i))
#define C4(k, j, i) (((k) * 4 + (j)) * 4 + (i))            /* trace region 0..3 */
#define F3(k3, j3, i3) ((((k3)-1) * 3 + ((j3)-1)) * 3 + ((i3)-1))
  if (tl == 0) {
    amr_get3cubefather<3>(t, igrid, a.ilevel, s.nfc, s.ng);
    for (int j = 0; j < 27; j++) s.gnb[j] = s.nfc[j] > 0 ? t.son[s.nfc[j]] : 0;
  }
  __syncthreads();
  // ---- gather ----
  for (int e = tl; e < 27 * 8; e += MHD3_TPO) {
    const int jf = e / 8, is = e % 8;
    const int g = s.gnb[jf];
    if (g <= 0) continue;
    const int i1 = jf % 3, j1 = (jf / 3) % 3, k1 = jf / 9;
    const int i3 = 1 + 2 * (i1 - 1) + (is & 1), j3 = 1 + 2 * (j1 - 1) + ((is >> 1) & 1), k3 = 1 + 2 * (k1 - 1) + (is >> 2);
    const int ic = amr_cell(t, is, g);
    for (int iv = 0; iv < MNVS; iv++) s.uloc[C6(k3, j3, i3)][iv] = a.uold[(size_t)iv * NC + ic - 1];
    s.ok[C6(k3, j3, i3)] = t.son[ic] > 0;
  }
  for (int jf = tl; jf < 27; jf += MHD3_TPO) {
    if (s.gnb[jf] > 0) continue;
    double u2[8][MNVS];
    mhd3_interpol_cell(t, a.uold, s.nfc[jf], a.ilevel, a.interpol_type, a.interpol_mag_type, smallr, u2);
    const int i1 = jf % 3, j1 = (jf / 3) % 3, k1 = jf / 9;
    for (int is = 0; is < 8; is++) {
      const int i3 = 1 + 2 * (i1 - 1) + (is & 1), j3 = 1 + 2 * (j1 - 1) + ((is >> 1) & 1), k3 = 1 + 2 * (k1 - 1) + (is >> 2);
      for (int iv = 0; iv < MNVS; iv++) s.uloc[C6(k3, j3, i3)][iv] = u2[is][iv];
      s.ok[C6(k3, j3, i3)] = 0;
    }
  }
  __syncthreads();
  // ---- ctoprim :2029 ----
  for (int e = tl; e < 216; e += MHD3_TPO) mhd_ctoprim_cell(P, s.uloc[e], s.q[e]);
  for (int e = tl; e < 343; e += MHD3_TPO) {
    const int i = e % 7 - 1, j = (e / 7) % 7 - 1, k = e / 49 - 1;
    if (j <= 4 && k <= 4) s.bf[e][0] = (i <= 4) ? s.uloc[C6(k, j, i)][5] : s.uloc[C6(k, j, i - 1)][MNV + 0];
    if (i <= 4 && k <= 4) s.bf[e][1] = (j <= 4) ? s.uloc[C6(k, j, i)][6] : s.uloc[C6(k, j - 1, i)][MNV + 1];
    if (i <= 4 && j <= 4) s.bf[e][2] = (k <= 4) ? s.uloc[C6(k, j, i)][7] : s.uloc[C6(k - 1, j, i)][MNV + 2];
  }
  for (int e = tl; e < 343 * 6; e += MHD3_TPO) (&s.dbf[0][0][0])[e] = 0.0;
  __syncthreads();
  // ---- uslope :2187 (slope types 0, 1, 2) ----
  for (int e = tl; e < 64 * MNV; e += MHD3_TPO) {
    const int n = e % MNV, c = e / MNV, i = c % 4, j = (c / 4) % 4, k = c / 16;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    if (P.slope_type == 1 || P.slope_type == 2) {
      const double st = (double)P.slope_type;
      d0 = mslope(st, s.q[C6(k, j, i - 1)][n], s.q[C6(k, j, i)][n], s.q[C6(k, j, i + 1)][n]);
      d1 = mslope(st, s.q[C6(k, j - 1, i)][n], s.q[C6(k, j, i)][n], s.q[C6(k, j + 1, i)][n]);
      d2 = mslope(st, s.q[C6(k - 1, j, i)][n], s.q[C6(k, j, i)][n], s.q[C6(k + 1, j, i)][n]);
    }
    s.dq[C4(k, j, i)][n][0] = d0; s.dq[C4(k, j, i)][n][1] = d1; s.dq[C4(k, j, i)][n][2] = d2;
  }
  if (P.slope_mag_type == 1 || P.slope_mag_type == 2) {
    const double st = (double)P.slope_mag_type;
    for (int e = tl; e < 4 * 4 * 5; e += MHD3_TPO) {
      {   // bf x: k 0..3, j 0..3, i 0..4
        const int i = e % 5, j = (e / 5) % 4, k = e / 20;
        s.dbf[C7(k, j, i)][0][0] = mslope(st, s.bf[C7(k, j - 1, i)][0], s.bf[C7(k, j, i)][0], s.bf[C7(k, j + 1, i)][0]);
        s.dbf[C7(k, j, i)][0][1] = mslope(st, s.bf[C7(k - 1, j, i)][0], s.bf[C7(k, j, i)][0], s.bf[C7(k + 1, j, i)][0]);
      }
      {   // bf y: k 0..3, j 0..4, i 0..3
        const int i = e % 4, j = (e / 4) % 5, k = e / 20;
        s.dbf[C7(k, j, i)][1][0] = mslope(st, s.bf[C7(k, j, i - 1)][1], s.bf[C7(k, j, i)][1], s.bf[C7(k, j, i + 1)][1]);
        s.dbf[C7(k, j, i)][1][1] = mslope(st, s.bf[C7(k - 1, j, i)][1], s.bf[C7(k, j, i)][1], s.bf[C7(k + 1, j, i)][1]);
      }
      {   // bf z: k 0..4, j 0..3, i 0..3
        const int i = e % 4, j = (e / 4) % 4, k = e / 16;
        s.dbf[C7(k, j, i)][2][0] = mslope(st, s.bf[C7(k, j, i - 1)][2], s.bf[C7(k, j, i)][2], s.bf[C7(k, j, i + 1)][2]);
        s.dbf[C7(k, j, i)][2][1] = mslope(st, s.bf[C7(k, j - 1, i)][2], s.bf[C7(k, j, i)][2], s.bf[C7(k, j + 1, i)][2]);
      }
    }
  }
  // ---- trace3d :750: edge-centred electric fields :812-836 on i,j,k = 0..4 ----
  for (int e = tl; e < 125; e += MHD3_TPO) {
    const int i = e % 5, j = (e / 5) % 5, k = e / 25;
#define Q(di, dj, dk, n) s.q[C6(k + (dk), j + (dj), i + (di))][n]
#define BF(di, dj, dk, n) s.bf[C7(k + (dk), j + (dj), i + (di))][n]
    double v = 0.25 * (Q(0, -1, -1, 2) + Q(0, -1, 0, 2) + Q(0, 0, -1, 2) + Q(0, 0, 0, 2));
    double ww = 0.25 * (Q(0, -1, -1, 3) + Q(0, -1, 0, 3) + Q(0, 0, -1, 3) + Q(0, 0, 0, 3));
    double B = 0.5 * (BF(0, 0, -1, 1) + BF(0, 0, 0, 1));
    double C = 0.5 * (BF(0, -1, 0, 2) + BF(0, 0, 0, 2));
    s.Ex[C6(k, j, i)] = v * C - ww * B;
    double u = 0.25 * (Q(-1, 0, -1, 1) + Q(-1, 0, 0, 1) + Q(0, 0, -1, 1) + Q(0, 0, 0, 1));
    ww = 0.25 * (Q(-1, 0, -1, 3) + Q(-1, 0, 0, 3) + Q(0, 0, -1, 3) + Q(0, 0, 0, 3));
    double A = 0.5 * (BF(0, 0, -1, 0) + BF(0, 0, 0, 0));
    C = 0.5 * (BF(-1, 0, 0, 2) + BF(0, 0, 0, 2));
    s.Ey[C6(k, j, i)] = ww * A - u * C;
    u = 0.25 * (Q(-1, -1, 0, 1) + Q(-1, 0, 0, 1) + Q(0, -1, 0, 1) + Q(0, 0, 0, 1));
    v = 0.25 * (Q(-1, -1, 0, 2) + Q(-1, 0, 0, 2) + Q(0, -1, 0, 2) + Q(0, 0, 0, 2));
    A = 0.5 * (BF(0, -1, 0, 0) + BF(0, 0, 0, 0));
    B = 0.5 * (BF(-1, 0, 0, 1) + BF(0, 0, 0, 1));
    s.Ez[C6(k, j, i)] = u * B - v * A;
  }
  __syncthreads();
  const double dtdx = dt / dx, dtdy = dt / dx, dtdz = dt / dx;
  for (int e = tl; e < 64; e += MHD3_TPO) {
    const int i = e % 4, j = (e / 4) % 4, k = e / 16;
    const double* q = s.q[C6(k, j, i)];
    double(*dq)[3] = s.dq[C4(k, j, i)];
    double r = q[0], u = q[1], v = q[2], ww = q[3], pp = q[4], A = q[5], B = q[6], C = q[7];
    double AL = BF(0, 0, 0, 0), AR = BF(1, 0, 0, 0), BL = BF(0, 0, 0, 1), BR = BF(0, 1, 0, 1), CL = BF(0, 0, 0, 2), CR = BF(0, 0, 1, 2);
    const double drx = 0.5 * dq[0][0], dux = 0.5 * dq[1][0], dvx = 0.5 * dq[2][0], dwx = 0.5 * dq[3][0], dpx = 0.5 * dq[4][0];
    const double dBx = 0.5 * dq[6][0], dCx = 0.5 * dq[7][0];
    const double dry = 0.5 * dq[0][1], duy = 0.5 * dq[1][1], dvy = 0.5 * dq[2][1], dwy = 0.5 * dq[3][1], dpy = 0.5 * dq[4][1];
    const double dAy = 0.5 * dq[5][1], dCy = 0.5 * dq[7][1];
    const double drz = 0.5 * dq[0][2], duz = 0.5 * dq[1][2], dvz = 0.5 * dq[2][2], dwz = 0.5 * dq[3][2], dpz = 0.5 * dq[4][2];
    const double dAz = 0.5 * dq[5][2], dBz = 0.5 * dq[6][2];
#define DBF(di, dj, dk, c, tt) s.dbf[C7(k + (dk), j + (dj), i + (di))][c][tt]
    const double dALy = 0.5 * DBF(0, 0, 0, 0, 0), dARy = 0.5 * DBF(1, 0, 0, 0, 0), dALz = 0.5 * DBF(0, 0, 0, 0, 1), dARz = 0.5 * DBF(1, 0, 0, 0, 1);
    const double dBLx = 0.5 * DBF(0, 0, 0, 1, 0), dBRx = 0.5 * DBF(0, 1, 0, 1, 0), dBLz = 0.5 * DBF(0, 0, 0, 1, 1), dBRz = 0.5 * DBF(0, 1, 0, 1, 1);
    const double dCLx = 0.5 * DBF(0, 0, 0, 2, 0), dCRx = 0.5 * DBF(0, 0, 1, 2, 0), dCLy = 0.5 * DBF(0, 0, 0, 2, 1), dCRy = 0.5 * DBF(0, 0, 1, 2, 1);
#define EX(dj, dk) s.Ex[C6(k + (dk), j + (dj), i)]
#define EY(di, dk) s.Ey[C6(k + (dk), j, i + (di))]
#define EZ(di, dj) s.Ez[C6(k, j + (dj), i + (di))]
    const double ELL = EX(0, 0), ELR = EX(0, 1), ERL = EX(1, 0), ERR = EX(1, 1);
    const double FLL = EY(0, 0), FLR = EY(0, 1), FRL = EY(1, 0), FRR = EY(1, 1);
    const double GLL = EZ(0, 0), GLR = EZ(0, 1), GRL = EZ(1, 0), GRR = EZ(1, 1);
    const double sAL0 = +(GLR - GLL) * dtdy * 0.5 - (FLR - FLL) * dtdz * 0.5;
    const double sAR0 = +(GRR - GRL) * dtdy * 0.5 - (FRR - FRL) * dtdz * 0.5;
    const double sBL0 = -(GRL - GLL) * dtdx * 0.5 + (ELR - ELL) * dtdz * 0.5;
    const double sBR0 = -(GRR - GLR) * dtdx * 0.5 + (ERR - ERL) * dtdz * 0.5;
    const double sCL0 = +(FRL - FLL) * dtdx * 0.5 - (ERL - ELL) * dtdy * 0.5;
    const double sCR0 = +(FRR - FLR) * dtdx * 0.5 - (ERR - ELR) * dtdy * 0.5;
    AL = AL + sAL0; AR = AR + sAR0; BL = BL + sBL0; BR = BR + sBR0; CL = CL + sCL0; CR = CR + sCR0;
    const double sr0 = (-u * drx - dux * r) * dtdx + (-v * dry - dvy * r) * dtdy + (-ww * drz - dwz * r) * dtdz;
    const double su0 = (-u * dux - (dpx + B * dBx + C * dCx) / r) * dtdx + (-v * duy + B * dAy / r) * dtdy + (-ww * duz + C * dAz / r) * dtdz;
    const double sv0 = (-u * dvx + A * dBx / r) * dtdx + (-v * dvy - (dpy + A * dAy + C * dCy) / r) * dtdy + (-ww * dvz + C * dBz / r) * dtdz;
    const double sw0 = (-u * dwx + A * dCx / r) * dtdx + (-v * dwy + B * dCy / r) * dtdy + (-ww * dwz - (dpz + A * dAz + B * dBz) / r) * dtdz;
    const double sp0 = (-u * dpx - dux * gamma * pp) * dtdx + (-v * dpy - dvy * gamma * pp) * dtdy + (-ww * dpz - dwz * gamma * pp) * dtdz;
    r = r + sr0; u = u + su0; v = v + sv0; ww = ww + sw0; pp = pp + sp0;
    A = 0.5 * (AL + AR); B = 0.5 * (BL + BR); C = 0.5 * (CL + CR);
    const int c4 = C4(k, j, i);
    auto SET = [&](double (*arr)[MNV][3], int d, double R, double U, double V, double W, double P_, double A_, double B_, double C_) {
      double(*s_)[3] = arr[c4];
      s_[0][d] = R; s_[1][d] = U; s_[2][d] = V; s_[3][d] = W; s_[4][d] = P_; s_[5][d] = A_; s_[6][d] = B_; s_[7][d] = C_;
      if (s_[0][d] < smallr) s_[0][d] = r;
      s_[4][d] = fmx(smallp, s_[4][d]);
    };
    SET(s.qp, 0, r - drx, u - dux, v - dvx, ww - dwx, pp - dpx, AL, B - dBx, C - dCx);
    SET(s.qm, 0, r + drx, u + dux, v + dvx, ww + dwx, pp + dpx, AR, B + dBx, C + dCx);
    SET(s.qp, 1, r - dry, u - duy, v - dvy, ww - dwy, pp - dpy, A - dAy, BL, C - dCy);
    SET(s.qm, 1, r + dry, u + duy, v + dvy, ww + dwy, pp + dpy, A + dAy, BR, C + dCy);
    SET(s.qp, 2, r - drz, u - duz, v - dvz, ww - dwz, pp - dpz, A - dAz, B - dBz, CL);
    SET(s.qm, 2, r + drz, u + duz, v + dvz, ww + dwz, pp + dpz, A + dAz, B + dBz, CR);
    SET(s.qRT, 0, r + (+dry + drz), u + (+duy + duz), v + (+dvy + dvz), ww + (+dwy + dwz), pp + (+dpy + dpz), A + (+dAy + dAz), BR + (+dBRz), CR + (+dCRy));
    SET(s.qRB, 0, r + (+dry - drz), u + (+duy - duz), v + (+dvy - dvz), ww + (+dwy - dwz), pp + (+dpy - dpz), A + (+dAy - dAz), BR + (-dBRz), CL + (+dCLy));
    SET(s.qLT, 0, r + (-dry + drz), u + (-duy + duz), v + (-dvy + dvz), ww + (-dwy + dwz), pp + (-dpy + dpz), A + (-dAy + dAz), BL + (+dBLz), CR + (-dCRy));
    SET(s.qLB, 0, r + (-dry - drz), u + (-duy - duz), v + (-dvy - dvz), ww + (-dwy - dwz), pp + (-dpy - dpz), A + (-dAy - dAz), BL + (-dBLz), CL + (-dCLy));
    SET(s.qRT, 1, r + (+drx + drz), u + (+dux + duz), v + (+dvx + dvz), ww + (+dwx + dwz), pp + (+dpx + dpz), AR + (+dARz), B + (+dBx + dBz), CR + (+dCRx));
    SET(s.qRB, 1, r + (+drx - drz), u + (+dux - duz), v + (+dvx - dvz), ww + (+dwx - dwz), pp + (+dpx - dpz), AR + (-dARz), B + (+dBx - dBz), CL + (+dCLx));
    SET(s.qLT, 1, r + (-drx + drz), u + (-dux + duz), v + (-dvx + dvz), ww + (-dwx + dwz), pp + (-dpx + dpz), AL + (+dALz), B + (-dBx + dBz), CR + (-dCRx));
    SET(s.qLB, 1, r + (-drx - drz), u + (-dux - duz), v + (-dvx - dvz), ww + (-dwx - dwz), pp + (-dpx - dpz), AL + (-dALz), B + (-dBx - dBz), CL + (-dCLx));
    SET(s.qRT, 2, r + (+drx + dry), u + (+dux + duy), v + (+dvx + dvy), ww + (+dwx + dwy), pp + (+dpx + dpy), AR + (+dARy), BR + (+dBRx), C + (+dCx + dCy));
    SET(s.qRB, 2, r + (+drx - dry), u + (+dux - duy), v + (+dvx - dvy), ww + (+dwx - dwy), pp + (+dpx - dpy), AR + (-dARy), BL + (+dBLx), C + (+dCx - dCy));
    SET(s.qLT, 2, r + (-drx + dry), u + (-dux + duy), v + (-dvx + dvy), ww + (-dwx + dwy), pp + (-dpx + dpy), AL + (+dALy), BR + (-dBRx), C + (-dCx + dCy));
    SET(s.qLB, 2, r + (-drx - dry), u + (-dux - duy), v + (-dvx - dvy), ww + (-dwx - dwy), pp + (-dpx - dpy), AL + (-dALy), BL + (-dBLx), C + (-dCx - dCy));
  }
#undef Q
#undef BF
#undef DBF
#undef EX
#undef EY
#undef EZ
  __syncthreads();
  // ---- cmpflxm :1308: 3 x 12 faces; cmp_mag_flx :1453: 3 x 18 edges; resets at refined faces / edges :760-879 ----
  for (int e = tl; e < 36 + 54; e += MHD3_TPO) {
    if (e < 36) {
      const int idim = e / 12, f = e % 12;
      const int i0 = idim == 0, j0 = idim == 1, k0 = idim == 2;
      const int ni = 2 + i0, nj = 2 + j0;
      const int i = 1 + f % ni, j = 1 + (f / ni) % nj, k = 1 + f / (ni * nj);
      const int perm[3][6] = {{2, 3, 4, 6, 7, 8}, {3, 2, 4, 7, 6, 8}, {4, 2, 3, 8, 6, 7}};
      const int ln = perm[idim][0] - 1, lt1 = perm[idim][1] - 1, lt2 = perm[idim][2] - 1;
      const int bn = perm[idim][3] - 1, bt1 = perm[idim][4] - 1, bt2 = perm[idim][5] - 1;
      double(*qm)[3] = s.qm[C4(k - k0, j - j0, i - i0)];
      double(*qp)[3] = s.qp[C4(k, j, i)];
      real ql[8], qr[8], fg[9];
      const double bn_mean = 0.5 * (qm[bn][idim] + qp[bn][idim]);
      ql[0] = qm[0][idim]; ql[1] = qm[4][idim]; ql[2] = qm[ln][idim]; ql[3] = bn_mean;
      ql[4] = qm[lt1][idim]; ql[5] = qm[bt1][idim]; ql[6] = qm[lt2][idim]; ql[7] = qm[bt2][idim];
      qr[0] = qp[0][idim]; qr[1] = qp[4][idim]; qr[2] = qp[ln][idim]; qr[3] = bn_mean;
      qr[4] = qp[lt1][idim]; qr[5] = qp[bt1][idim]; qr[6] = qp[lt2][idim]; qr[7] = qp[bt2][idim];
      mhd_riemann1d_rt(a.riemann, P, ql, qr, fg);
      double* fl = s.flux[idim][F3(k, j, i)];
      fl[0] = fg[0].v; fl[4] = fg[1].v; fl[ln] = fg[2].v; fl[bn] = fg[3].v; fl[lt1] = fg[4].v; fl[bt1] = fg[5].v; fl[lt2] = fg[6].v; fl[bt2] = fg[7].v;
      for (int n = 0; n < MNV; n++) fl[n] = fl[n] * dt / dx;
      if (s.ok[C6(k - k0, j - j0, i - i0)] || s.ok[C6(k, j, i)])
        for (int n = 0; n < MNV; n++) fl[n] = 0.0;
      fl[5] = 0.0; fl[6] = 0.0; fl[7] = 0.0;
    } else {
      const int g = e - 36, dir = 2 - g / 18, h = g % 18;   // emfz first like the reference (:147-240); the order is immaterial
      double RT[8], RB[8], LT[8], LB[8];
      int i, j, k;
      bool masked;
      if (dir == 2) {        // k 1..2, j 1..3, i 1..3
        i = 1 + h % 3; j = 1 + (h / 3) % 3; k = 1 + h / 9;
        for (int n = 0; n < 8; n++) {
          RT[n] = s.qRT[C4(k, j - 1, i - 1)][n][2]; RB[n] = s.qRB[C4(k, j, i - 1)][n][2];
          LT[n] = s.qLT[C4(k, j - 1, i)][n][2];     LB[n] = s.qLB[C4(k, j, i)][n][2];
        }
        masked = s.ok[C6(k, j, i)] || s.ok[C6(k, j - 1, i)] || s.ok[C6(k, j, i - 1)] || s.ok[C6(k, j - 1, i - 1)];
      } else if (dir == 1) { // k 1..3, j 1..2, i 1..3; second and third arguments swapped (:213-218)
        i = 1 + h % 3; j = 1 + (h / 3) % 2; k = 1 + h / 6;
        for (int n = 0; n < 8; n++) {
          RT[n] = s.qRT[C4(k - 1, j, i - 1)][n][1]; RB[n] = s.qLT[C4(k - 1, j, i)][n][1];
          LT[n] = s.qRB[C4(k, j, i - 1)][n][1];     LB[n] = s.qLB[C4(k, j, i)][n][1];
        }
        masked = s.ok[C6(k, j, i)] || s.ok[C6(k - 1, j, i)] || s.ok[C6(k, j, i - 1)] || s.ok[C6(k - 1, j, i - 1)];
      } else {               // k 1..3, j 1..3, i 1..2
        i = 1 + h % 2; j = 1 + (h / 2) % 3; k = 1 + h / 6;
        for (int n = 0; n < 8; n++) {
          RT[n] = s.qRT[C4(k - 1, j - 1, i)][n][0]; RB[n] = s.qRB[C4(k, j - 1, i)][n][0];
          LT[n] = s.qLT[C4(k - 1, j, i)][n][0];     LB[n] = s.qLB[C4(k, j, i)][n][0];
        }
        masked = s.ok[C6(k, j, i)] || s.ok[C6(k - 1, j, i)] || s.ok[C6(k, j - 1, i)] || s.ok[C6(k - 1, j - 1, i)];
      }
      double ez = mhd_emf_rt(a.riemann2d, P, RT, RB, LT, LB, dir) * dt / dx;
      if (masked) ez = 0.0;
      s.emf[dir][F3(k, j, i)] = ez;
    }
  }
  __syncthreads();
  // ---- update of the oct's own cells: Euler system x, y, z (:886-934), then constrained transport (:939-995) ----
#define RX(i3, j3, k3) s.emf[0][F3(k3, j3, i3)]
#define RY(i3, j3, k3) s.emf[1][F3(k3, j3, i3)]
#define RZ(i3, j3, k3) s.emf[2][F3(k3, j3, i3)]
  for (int e = tl; e < 8 * MNVS; e += MHD3_TPO) {
    const int is = e % 8, iv = e / 8;
    const int i3 = 1 + (is & 1), j3 = 1 + ((is >> 1) & 1), k3 = 1 + (is >> 2);
    const int ic = amr_cell(t, is, igrid);
    const int src = iv < MNV ? iv : 5 + (iv - MNV);
    double u = a.unew[(size_t)iv * NC + ic - 1];
    u = u + (s.flux[0][F3(k3, j3, i3)][src] - s.flux[0][F3(k3, j3, i3 + 1)][src]);
    u = u + (s.flux[1][F3(k3, j3, i3)][src] - s.flux[1][F3(k3, j3 + 1, i3)][src]);
    u = u + (s.flux[2][F3(k3, j3, i3)][src] - s.flux[2][F3(k3 + 1, j3, i3)][src]);
    if (iv == 5) u = u + ((RY(i3, j3, k3) - RY(i3, j3, k3 + 1)) - (RZ(i3, j3, k3) - RZ(i3, j3 + 1, k3)));
    if (iv == MNV + 0) u = u + ((RY(i3 + 1, j3, k3) - RY(i3 + 1, j3, k3 + 1)) - (RZ(i3 + 1, j3, k3) - RZ(i3 + 1, j3 + 1, k3)));
    if (iv == 6) u = u + ((RZ(i3, j3, k3) - RZ(i3 + 1, j3, k3)) - (RX(i3, j3, k3) - RX(i3, j3, k3 + 1)));
    if (iv == MNV + 1) u = u + ((RZ(i3, j3 + 1, k3) - RZ(i3 + 1, j3 + 1, k3)) - (RX(i3, j3 + 1, k3) - RX(i3, j3 + 1, k3 + 1)));
    if (iv == 7) u = u + ((RX(i3, j3, k3) - RX(i3, j3 + 1, k3)) - (RY(i3, j3, k3) - RY(i3 + 1, j3, k3)));
    if (iv == MNV + 2) u = u + ((RX(i3, j3, k3 + 1) - RX(i3, j3 + 1, k3 + 1)) - (RY(i3, j3, k3 + 1) - RY(i3 + 1, j3, k3 + 1)));
    a.unew[(size_t)iv * NC + ic - 1] = u;
  }
#undef RX
#undef RY
#undef RZ
  // ---- outer-face fluxes [side][face = j/k fastest-first transverse index][8] and all 54 edge EMFs for the coarse refluxing ----
  for (int e = tl; e < 6 * 4 * MNV; e += MHD3_TPO) {
    const int n = e % MNV, fs = (e / MNV) % 4, side = e / (4 * MNV);
    const int idim = side / 2, right = side % 2;
    const int a0 = fs & 1, a1 = fs >> 1;   // transverse face coordinates, lower dimension first (the loop order k3, j3, i3 of :1030-1168)
    int i3, j3, k3;
    if (idim == 0) { i3 = right ? 3 : 1; j3 = 1 + a0; k3 = 1 + a1; }
    else if (idim == 1) { j3 = right ? 3 : 1; i3 = 1 + a0; k3 = 1 + a1; }
    else { k3 = right ? 3 : 1; i3 = 1 + a0; j3 = 1 + a1; }
    a.rflux[(((size_t)io * 6 + side) * 4 + fs) * MNV + n] = s.flux[idim][F3(k3, j3, i3)][n];
  }
  for (int e = tl; e < 81; e += MHD3_TPO) a.remf[(size_t)io * 81 + e] = s.emf[e / 27][e % 27];
#undef X_
#undef C6
#undef C7
#undef C4
#undef F3
}

// coarse refluxing of the edge EMFs, NDIM = 3 (:1176-1455): per target (cell, variable) the contributions
// (emf(c0) + emf(c1)) * 0.25 * weight [* 0.5], code = oct << 7 | edge (0..11) << 3 | (weight 0.5 ? 4 : 0) | (half ? 2 : 0) | (minus ? 1 : 0)
struct MhdEdge3 { signed char dir, c0, c1; };   // c = (k3-1)*9 + (j3-1)*3 + (i3-1) of the two emf entries of the edge
struct Emf3RefluxArgs {
  int nent;
  const int* cell; const int* var; const int* start; const int* code;
  const double* remf;  // [nact][3][27]
  double* unew;
  long long ncell;
  signed char edir[12], ec0[12], ec1[12];
};
__global__ void mhd_amr_emf3_reflux_kernel(const Emf3RefluxArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.nent) return;
  const size_t idx = (size_t)a.var[e] * a.ncell + a.cell[e] - 1;
  double u = a.unew[idx];
  for (int k = a.start[e]; k < a.start[e + 1]; k++) {
    const int c = a.code[k];
    const int oct = c >> 7, edge = (c >> 3) & 15;
    const double weight = (c & 4) ? 0.5 : 1.0;
    const double* em = a.remf + (size_t)oct * 81 + (size_t)a.edir[edge] * 27;
    const double dflux = (em[a.ec0[edge]] + em[a.ec1[edge]]) * 0.25 * weight;
    const double v = (c & 2) ? dflux * 0.5 : dflux;
    if (c & 1) u = u - v; else u = u + v;
  }
  a.unew[idx] = u;
}

// coarse refluxing of the Euler fluxes (mhd/godunov_fine.f90:1030-1168): same schedule as the hydro pass (RefluxArgs: target
// cells with their contributions (oct, side, face) in the reference's visiting order); variables 1..nvar take the fluxes
// 1..nvar, nvar+1..nvar+3 the fluxes 6..8
__global__ void mhd_amr_reflux_kernel(const RefluxArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nent * MNVS) return;
  const int e = i / MNVS, n = i % MNVS;
  const int src = n < MNV ? n : 5 + (n - MNV);
  const size_t idx = (size_t)n * a.ncell + a.cell[e] - 1;
  double u = a.unew[idx];
  for (int k = a.start[e]; k < a.start[e + 1]; k++) {
    const int sr = a.src[k];
    const int oct = sr >> 6, side = (sr >> 3) & 7, face = sr & 7;
    const double f = a.rflux[(((size_t)oct * a.nsides + side) * a.nsf + face) * MNV + src] * a.oneontwotondim;
    if (side & 1) u = u + f; else u = u - f;
  }
  a.unew[idx] = u;
}
// coarse refluxing of the corner EMFs, NDIM = 2 (:1176-1270): entries grouped per target (cell, variable) with their
// contributions in the reference's visiting order; code = oct*32 + corner*8 + (weight 0.5 ? 4 : 0) + (half ? 2 : 0) + (minus ? 1 : 0)
struct EmfRefluxArgs {
  int nent;
  const int* cell;     // [nent] target cell
  const int* var;      // [nent] target variable (0-based: 5, 6, 8, 9)
  const int* start;    // [nent+1]
  const int* code;     // contributions
  const double* remf;  // [nact][4]
  double* unew;
  long long ncell;
};
__global__ void mhd_amr_emf_reflux_kernel(const EmfRefluxArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.nent) return;
  const size_t idx = (size_t)a.var[e] * a.ncell + a.cell[e] - 1;
  double u = a.unew[idx];
  for (int k = a.start[e]; k < a.start[e + 1]; k++) {
    const int c = a.code[k];
    const int oct = c >> 5, corner = (c >> 3) & 3;
    const double weight = (c & 4) ? 0.5 : 1.0;
    const double ez = a.remf[(size_t)oct * 4 + corner];
    const double dflux = (ez + ez) * 0.25 * weight;   // emfz(:,:,:,2) is a copy of emfz(:,:,:,1) (umuscl.f90:190-199)
    const double v = (c & 2) ? dflux * 0.5 : dflux;
    if (c & 1) u = u - v; else u = u + v;
  }
  a.unew[idx] = u;
}
// the 3^ndim father cells of every active oct (get3cubefather), for the host-built EMF reflux schedule
template <int NDIM>
__global__ void amr_nfc_kernel(const AmrTree t, const int* __restrict__ active, int nact, int ilevel, int* __restrict__ out) {
  const int io = blockIdx.x * blockDim.x + threadIdx.x;
  if (io >= nact) return;
  int nfc[27], ng[8];
  amr_get3cubefather<NDIM>(t, active[io], ilevel, nfc, ng);
  constexpr int N3 = NDIM == 1 ? 3 : (NDIM == 2 ? 9 : 27);
  for (int j = 0; j < N3; j++) out[(size_t)io * N3 + j] = nfc[j];
}

// ====================================================================================================== list passes
// upload_fine (mhd/interpol_hydro.f90:5-231): pass 0 = upl (:233) on the split cells, pass 1 = upl_left / upl_right on the leaf
// cells next to a refined cell (:70-231); two launches (pass 1 reads the fine level only, pass 0 writes split cells only)
__global__ void mhd_amr_upload_kernel(double* __restrict__ u, const AmrTree t, const int* __restrict__ igrid, int n, int ndim, double smallr,
                                      int pass) {
  const int hhh[6][4] = {{1, 3, 5, 7}, {2, 4, 6, 8}, {1, 2, 5, 6}, {3, 4, 7, 8}, {1, 2, 3, 4}, {5, 6, 7, 8}};
  const int T = 1 << ndim, Th = T / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * T) return;
  const int o = i % n, ind = i / n;
  const int ig = igrid[o];
  const size_t NC = (size_t)t.ncell;
  const int ic = amr_cell(t, ind, ig);
  auto U = [&](int c, int iv) -> double& { return u[(size_t)iv * NC + c - 1]; };
  if (pass == 0) {
    const int gs = t.son[ic];
    if (gs <= 0) return;
    double getx = 0.0;
    for (int is = 0; is < T; is++) getx = getx + fmx(U(amr_cell(t, is, gs), 0), smallr);
    U(ic, 0) = getx / (double)T;
    for (int iv = 1; iv < MNV; iv++)
      if (iv <= 4 || iv > 4 + ndim) {
        getx = 0.0;
        for (int is = 0; is < T; is++) getx = getx + U(amr_cell(t, is, gs), iv);
        U(ic, iv) = getx / (double)T;
      }
    if (ndim == 1) { U(ic, MNV + 1) = U(ic, 6); U(ic, MNV + 2) = U(ic, 7); }
    if (ndim == 2) U(ic, MNV + 2) = U(ic, 7);
    for (int idim = 1; idim <= ndim; idim++) {
      getx = 0.0;
      for (int k = 0; k < Th; k++) getx = getx + U(amr_cell(t, hhh[2 * idim - 2][k] - 1, gs), 4 + idim);
      U(ic, 4 + idim) = getx / (double)Th;
      getx = 0.0;
      for (int k = 0; k < Th; k++) getx = getx + U(amr_cell(t, hhh[2 * idim - 1][k] - 1, gs), MNV - 1 + idim);
      U(ic, MNV - 1 + idim) = getx / (double)Th;
    }
    return;
  }
  if (t.son[ic] != 0) return;
  for (int d = 0; d < ndim; d++) {
    const int b = (ind >> d) & 1, other = ind ^ (1 << d);
    // left neighbour cell (same level): inside the oct for b = 1, in the left neighbour oct for b = 0
    int gl = b == 1 ? ig : t.son[amr_nbor(t, ig, 2 * d + 1)];
    if (gl > 0) {
      const int sc = t.son[amr_cell(t, other, gl)];
      if (sc > 0) {   // upl_left: left B of the leaf = mean of the right B of the touching sons
        double getx = 0.0;
        for (int k = 0; k < Th; k++) getx = getx + U(amr_cell(t, hhh[2 * d + 1][k] - 1, sc), MNV + d);
        U(ic, 5 + d) = getx / (double)Th;
      }
    }
    int gr = b == 0 ? ig : t.son[amr_nbor(t, ig, 2 * d + 2)];
    if (gr > 0) {
      const int sc = t.son[amr_cell(t, other, gr)];
      if (sc > 0) {   // upl_right
        double getx = 0.0;
        for (int k = 0; k < Th; k++) getx = getx + U(amr_cell(t, hhh[2 * d][k] - 1, sc), 5 + d);
        U(ic, MNV + d) = getx / (double)Th;
      }
    }
  }
}

// cmpdt for one cell with ctot summed over the directions 1..ndim (mhd/godunov_utils.f90:5-111 as built for NDIM < 3)
__device__ __forceinline__ double mhd_cmpdt_cell_nd(const MPhys& M, double* uu, double dx, int ndim) {
  uu[0] = fmx(uu[0], M.smallr);
  const double rho = uu[0];
  for (int d = 1; d <= 3; d++) uu[d] = uu[d] / rho;
  double B2 = 0.0;
  for (int d = 1; d <= 3; d++) {
    const double Bc = 0.5 * (uu[4 + d] + uu[MNV + d - 1]);
    B2 = B2 + Bc * Bc;
    uu[4] = uu[4] - 0.5 * uu[0] * (uu[d] * uu[d]) - 0.5 * (Bc * Bc);
  }
  uu[4] = fmx((M.gamma - 1.0) * uu[4], M.smallp);
  const double a2 = M.gamma * uu[4] / uu[0];
  double ctot = 0.0;
  for (int d = 1; d <= ndim; d++) {
    const double cc = 0.5 * (B2 / rho + a2);
    const double BN = 0.5 * (uu[4 + d] + uu[MNV + d - 1]);
    const double cf = sqrt(cc + sqrt(cc * cc - a2 * (BN * BN) / rho));
    ctot = ctot + fabs(uu[d]) + cf;
  }
  double r = 0.0 * dx / (ctot * ctot);
  r = fmx(r, 0.0001);
  return dx / ctot * (sqrt(1.0 + 2.0 * M.courant_factor * r) - 1.0) / r;
}
// courant_fine over the leaf cells of the listed octs: part[0][block] = min dt (the diagnostics sums are not accumulated here)
__global__ void mhd_amr_courant_kernel(const double* __restrict__ u, const AmrTree t, const int* __restrict__ igrid, int n, MPhys P, double dx,
                                       int ndim, double* __restrict__ part) {
  __shared__ double red[32];
  const int T = 1 << ndim;
  const size_t NC = (size_t)t.ncell;
  double my_dt = 1e300;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n * T; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % n), ind = (int)(i / n);
    const int ic = amr_cell(t, ind, igrid[o]);
    if (t.son[ic] != 0) continue;
    double uu[MNVS];
    for (int k = 0; k < MNVS; k++) uu[k] = u[(size_t)k * NC + ic - 1];
    const double dtc = mhd_cmpdt_cell_nd(P, uu, dx, ndim);
    my_dt = dtc < my_dt ? dtc : my_dt;
  }
  my_dt = warp_min(my_dt);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = my_dt;
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    double v0 = l < nw ? red[l] : 1e300;
    v0 = warp_min(v0);
    if (l == 0) {
      const size_t nb = gridDim.x;
      part[0 * nb + blockIdx.x] = v0; part[1 * nb + blockIdx.x] = 0.0; part[2 * nb + blockIdx.x] = 0.0; part[3 * nb + blockIdx.x] = 0.0;
    }
  }
}

// make_boundary_hydro of the MHD build (mhd/hydro_boundary.f90:53-139) on the mirrored arrays, NDIM = 1: reflexive and outflow
struct MhdAmrBoundArgs { int n; const int* igrid; int dir; int kind; double smallr; };
__global__ void mhd_amr1_boundary_kernel(double* __restrict__ u, const AmrTree t, const MhdAmrBoundArgs b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n * 2) return;
  const int o = i / 2, ind = i % 2;
  const size_t NC = (size_t)t.ncell;
  auto U = [&](int c, int iv) -> double& { return u[(size_t)(iv - 1) * NC + c - 1]; };   // 1-based variable index like the reference
  const int dir = b.dir;
  const int inbor = dir == 1 ? 2 : 1;
  const int ref_x[2] = {2, 1}, free1[2] = {1, 1}, free2[2] = {2, 2}, alt1[2] = {-2, -1}, alt2[2] = {1, 2};
  const int* ind_ref = (b.kind == 0) ? ref_x : (dir == 1 ? free1 : free2);
  const int* ind_normal = dir == 1 ? free1 : free2;
  const int* alt = dir == 1 ? alt1 : alt2;
  const int iperp1 = dir == 1 ? 6 : MNV + 1;
  const int gdim = 1;
  const int ig = b.igrid[o];
  const int igr = t.son[amr_nbor(t, ig, inbor)];
  const int ic = amr_cell(t, ind, ig), icr = amr_cell(t, ind_ref[ind] - 1, igr);
  double uu[MNVS + 1];
  for (int iv = 1; iv <= MNVS; iv++) uu[iv] = U(icr, iv);
  auto sq = [](double x) { return x * x; };
  if (b.kind == 0) {
    const int icn = amr_cell(t, ind_normal[ind] - 1, igr);
    double emag = 0.125 * (sq(uu[6] + uu[MNV + 1]) + sq(uu[7] + uu[MNV + 2]) + sq(uu[8] + uu[MNV + 3]));
    uu[5] = uu[5] - emag;
    const double B_normal = U(icn, iperp1);
    for (int iv = 1; iv <= MNVS; iv++) {
      double sw = 1;
      if (iv == 2) sw = -1;
      if (iv != 5 + gdim && iv != MNV + gdim) U(ic, iv) = uu[iv] * sw;
      if (iv == 5 + gdim) U(ic, 5 + gdim) = 2 * B_normal - uu[MNV + gdim];
      if (iv == MNV + gdim) U(ic, MNV + gdim) = 2 * B_normal - uu[5 + gdim];
    }
    emag = 0.125 * (sq(U(ic, 6) + U(ic, MNV + 1)) + sq(U(ic, 7) + U(ic, MNV + 2)) + sq(U(ic, 8) + U(ic, MNV + 3)));
    U(ic, 5) = U(ic, 5) + emag;
  } else {
    double emag = 0.125 * (sq(uu[6] + uu[MNV + 1]) + sq(uu[7] + uu[MNV + 2]) + sq(uu[8] + uu[MNV + 3]));
    double ekin = 0.0, d = fmx(uu[1], b.smallr);
    { const double v = uu[2] / d; ekin = ekin + 0.5 * d * (v * v); }
    uu[5] = uu[5] - emag - ekin;
    for (int iv = 1; iv <= MNVS; iv++) {
      if (iv != 5 + gdim && iv != MNV + gdim) U(ic, iv) = uu[iv];
      if (iv == 5 + gdim) U(ic, 5 + gdim) = uu[5 + gdim] + (uu[MNV + gdim] - uu[5 + gdim]) * (double)alt[ind];
      if (iv == MNV + gdim) U(ic, MNV + gdim) = uu[MNV + gdim] + (uu[MNV + gdim] - uu[5 + gdim]) * (double)alt[ind];
    }
    emag = 0.125 * (sq(U(ic, 6) + U(ic, MNV + 1)) + sq(U(ic, 7) + U(ic, MNV + 2)) + sq(U(ic, 8) + U(ic, MNV + 3)));
    ekin = 0.0; d = fmx(U(ic, 1), b.smallr);
    { const double v = U(ic, 2) / d; ekin = ekin + 0.5 * d * (v * v); }
    U(ic, 5) = U(ic, 5) + emag + ekin;
  }
}

}  // namespace rgpu
