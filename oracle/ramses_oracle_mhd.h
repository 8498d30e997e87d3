/*
 * ramses_oracle_mhd.h -- CPU ORACLE, ideal-MHD variant (test infrastructure, NOT product code).
 * See ramses_oracle_mhd.c for the reference citations and the parity-pinning status.
 */
#ifndef RAMSES_ORACLE_MHD_H
#define RAMSES_ORACLE_MHD_H
#include "ramses_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* iriemann (hydro/read_hydro_params.f90:190-205) and iriemann2d (:207-220) */
enum { ORC_MHD_LLF = 0, ORC_MHD_ROE = 1, ORC_MHD_HLL = 2, ORC_MHD_HLLD = 3, ORC_MHD_UPWIND = 4, ORC_MHD_HYDRO = 5 };
enum { ORC_MHD2D_LLF = 0, ORC_MHD2D_ROE = 1, ORC_MHD2D_UPWIND = 2, ORC_MHD2D_HLL = 3, ORC_MHD2D_HLLA = 4, ORC_MHD2D_HLLD = 5 };

typedef struct {
  int slope_type;      /* mhd/hydro_parameters.f90:92 */
  int slope_mag_type;  /* :93 (-1 in the namelist means slope_type, resolved by the caller) */
  int riemann;         /* ORC_MHD_*   */
  int riemann2d;       /* ORC_MHD2D_* */
  double gamma, smallr, smallc, slope_theta, courant_factor, boxlen;
} orc_mhd_params;

typedef struct orc_mhd_work orc_mhd_work;
orc_mhd_work* orc_mhd_work_new(void);
void orc_mhd_work_free(orc_mhd_work*);
double* orc_mhd_work_uloc(orc_mhd_work*);          /* [6][6][6][11]  (k,j,i,ivar)          */
double* orc_mhd_work_flux(orc_mhd_work*);          /* [3][3][3][3][8] (idim,k3,j3,i3,ivar) */
double* orc_mhd_work_emf(orc_mhd_work*, int dir);  /* [3][3][3] emfx/emfy/emfz             */

/* q = (rho, P, v_n, B_n, v_t1, B_t1, v_t2, B_t2); fg[9] */
void orc_mhd_riemann(const orc_mhd_params*, const double* qleft, const double* qright, double* fgdnv);
/* corner states (rho,u,v,w,P,A,B,C) handed to cmp_mag_flx as (qRT,qRB,qLT,qLB); dir 0/1/2 = emfx/emfy/emfz call */
double orc_mhd_emf(const orc_mhd_params*, const double* RT, const double* RB, const double* LT, const double* LB, int dir);
void orc_mhd_unsplit(const orc_mhd_params*, orc_mhd_work*, double dx, double dt);
double orc_mhd_cmpdt_cell(const orc_mhd_params*, double* uu /*[11], destroyed*/, double dx);

void orc_mhd_set_unew(const orc_mesh*, int ilevel, const double* uold, double* unew);
void orc_mhd_set_uold(const orc_mesh*, int ilevel, double* uold, const double* unew);
void orc_mhd_godunov_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, double dt, const double* uold, double* unew, int nthreads);
double orc_mhd_courant_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, double dt_in, const double* uold, double sums[4]);
void orc_mhd_make_boundary_hydro(const orc_mhd_params*, const orc_mesh*, int ilevel, double* uold);
void orc_mhd_run_uniform(const orc_mhd_params*, const orc_mesh*, int ilevel, int nstep, double* uold, double* unew,
                         double* dt_hist, double* t_io, int nthreads);
void orc_mhd_set_threads(int n);

/* NDIM = 1 with AMR (tests/mhd/imhd-tube) */
void orc_mhd1_interpol_cell(const orc_mesh*, int ind_cell, int ilevel, const double* uold, double* u2 /*[2][11]*/);
void orc_mhd1_godunov_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, int levelmin, int nvector, double dt, const double* uold, double* unew);
void orc_mhdn_set_unew(const orc_mesh*, int ilevel, const double* uold, double* unew);
void orc_mhdn_set_uold(const orc_mesh*, int ilevel, double* uold, const double* unew);
double orc_mhd1_courant_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, double dt_in, const double* uold);
void orc_mhd1_make_boundary_hydro(const orc_mhd_params*, const orc_mesh*, int ilevel, double* uold);
void orc_mhd1_upload_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, double* uold);


/* NDIM = 2 with AMR (tests/mhd/orszag-tang) and dimension-generic passes */
void orc_mhd_set_interpol(int interpol_type, int interpol_mag_type);
void orc_mhd2_unsplit(const orc_mhd_params*, const double* uloc /*[6][6][11] (j,i,ivar)*/, double dx, double dt,
                      double* flux /*[2][3][3][8] (idim,j3,i3,ivar)*/, double* emfz /*[3][3] (j3,i3)*/);
void orc_mhd2_interpol_cell(const orc_mesh*, int ind_cell, int ilevel, const double* uold, double* u2 /*[4][11]*/);
void orc_mhd2_godunov_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, int levelmin, int nvector, double dt, const double* uold, double* unew);
double orc_mhdn_courant_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, double dt_in, const double* uold);
void orc_mhdn_upload_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, double* uold);
/* tests/mhd/orszag-tang/condinit.f90 on the active octs of a level */
void orc_mhd2_condinit_orszag_tang(const orc_mhd_params*, const orc_mesh*, int ilevel, double* uold);
/* NDIM = 3 with AMR (no golden file: held by equality with the NDIM=2 routines on z-invariant runs) */
void orc_mhd3_interpol_cell(const orc_mesh*, int ind_cell, int ilevel, const double* uold, double* u2 /*[8][11]*/);
void orc_mhd3_godunov_fine(const orc_mhd_params*, const orc_mesh*, int ilevel, int levelmin, int nvector, double dt, const double* uold, double* unew);
void orc_mhd_set_courant_ndim(int mask);   /* test hook: bit mask of the directions in cmpdt, 0 = all NDIM */
void orc_mhd3_condinit_orszag_tang(const orc_mhd_params*, const orc_mesh*, int ilevel, double* uold);
/* hydro_flag (hydro/hydro_flag.f90, SOLVERmhd) + hydro_refine mhd/godunov_utils.f90:113; err/flo = (d, p, b2, A, B, C, u) */
void orc_amr_mhd_hydro_flag(const orc_mhd_params*, const orc_mesh*, int ilevel, const double* uold, int* flag1,
                            const double err[7], const double flo[7]);

#ifdef __cplusplus
}
#endif
#endif
