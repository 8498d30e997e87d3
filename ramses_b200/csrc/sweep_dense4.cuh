// sweep_dense4.cuh -- one-barrier form of the fused 3-D dense-box Godunov sweep (sm_100a).
//
// Same contract, tiling, data layout and arithmetic as sweep3_kernel (sweep_dense3.cuh).  What changes is the schedule of the
// plane loop.  The ncu source page of sweep3 (profiles/r2_tune_sweep.md) shows two phases of equal length per plane: the
// predictor (slopes + trace; issue / shared-memory bound, FP64 pipe 35 % busy) and the solve (three Riemann problems + update;
// FP64 pipe 70 % busy) -- and, with two CTA barriers per plane, ALL warps are in the same phase at the same time.  Here the
// predictor of plane k+1 and the solves of plane k sit between the SAME pair of barriers:
//
//     barrier | late y-update of plane k-1 | solve(k)  and  predictor(k+1), in either order | barrier | ...
//
// They are independent (the predictor reads the ring of primitive planes, the solve reads the face states the previous
// iteration produced), so half of the warps run solve -> predictor and the other half predictor -> solve (ORD = 1): at any
// time some warps feed the FP64 pipe while the others use the issue slots / the shared-memory pipe.  Price: the face states of
// plane k (qp_x, qp_y, qp_z, the shuffled qm_x and the latest qm_z: 25 doubles) stay in registers across the barrier, the
// y exchanges (qm_y, F_y) are double buffered and the ring holds four planes: 218 KB of shared memory at BY = 12.
// The halo-row warps stage and convert the planes ahead (converter warps, as in sweep3).
//
// Bit-identical to sweep3 / the round-1 kernel / the oracle (tests/test_device_numerics_host.py runs it on the CPU).
#pragma once
#include "sweep_dense3.cuh"

namespace rgpu {

template <int BY>
struct Sweep4Smem {
  static constexpr int NV = 5, QX = 34, QY = BY + 2, NQ = 6, PL = QX * QY, NT = 32 * BY;
  static constexpr size_t ring = (size_t)4 * NQ * PL;     // planes k, k+1, k+2 read by the predictor + k+3 being converted
  static constexpr size_t stage = (size_t)NV * PL;
  static constexpr size_t exq = (size_t)2 * NV * NT;      // qm_y, double buffered by plane parity
  static constexpr size_t exf = (size_t)2 * NV * NT;      // F_y, double buffered by plane parity
  static constexpr size_t carry = (size_t)3 * NV * NT;    // per-thread: qm_z of plane k-1, Fz, partial update of the pending plane
  static constexpr size_t doubles = ring + stage + exq + exf + carry;
};

// ORD: 0 every warp solve -> predictor; 1 warps alternate (by groups of four: the warps of one scheduler get both orders);
//      2 every warp predictor -> solve
template <int RIEMANN, int SLOPE, int BY, int VEC, int ORD>
__global__ void __launch_bounds__(32 * BY, 1) sweep4_kernel(const SweepArgs a) {
  using S = Sweep4Smem<BY>;
  constexpr int NDIM = 3, BX = 32, NV = S::NV, QX = S::QX, NQ = S::NQ, NT = S::NT, PL = S::PL;
  constexpr int TXO = BX - 2, TYO = BY - 2;
#ifdef RGPU_HOST_NUMERICS
  double* smem = rgpu_host_dyn_smem;
#else
  extern __shared__ double smem[];
#endif
  double* qring = smem;
  double* stage = qring + S::ring;
  double* exq = stage + S::stage;
  double* exf = exq + S::exq;
  double* carry = exf + S::exf;                // [0..NV): qm_z(k-1)   [NV..2NV): Fz   [2NV..3NV): partial update
  __shared__ double red[4][NT / 32];

  const DenseGeom& g = a.g;
  const Phys& P = a.P;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * BX + tx;
  const double dt = a.dt_dev ? *a.dt_dev : a.dt_val;
  const double dtdx = dt / a.dx;
  const unsigned vstride = 8u * (unsigned)g.nslot;
  const bool trace_first = (ORD == 2) || (ORD == 1 && ((ty >> 2) & 1));   // warp uniform

  double my_dt = 1e300, my_mass = 0.0, my_etot = 0.0, my_eint = 0.0;

  long long w0 = a.nwork * blockIdx.x / gridDim.x;
  const long long w1 = a.nwork * (blockIdx.x + 1) / gridDim.x;
  while (w0 < w1) {
    int tix, tiy, z0, zn;
    sweep_work_decode(a, w0, w1, tix, tiy, z0, zn);
    w0 += zn;
    const int x0 = g.ox0 + tix * TXO, y0 = g.oy0 + tiy * TYO;
    const int z1 = z0 + zn;
    const int cx = x0 - 1 + tx, cy = y0 - 1 + ty;

    const bool col_own = (tx >= 1) && (tx <= BX - 2) && (cx < g.ox1);
    const bool row_own = (ty >= 1) && (ty <= BY - 2) && (cy < g.oy1);
    const bool own = col_own && row_own;

    constexpr int NOWN = (PL + NT - 1) / NT;
    const unsigned off_own = (unsigned)cell_offset<NDIM>(g, wrap_or_clamp(cx, g.ncx, g.wrapx), wrap_or_clamp(cy, g.ncy, g.wrapy), 0);
    auto zoff = [&](int z) -> unsigned {
      const int zc = wrap_or_clamp(z, g.ncz, g.wrapz);
      return (unsigned)((zc & 1) << 2) * (unsigned)g.nslot + (unsigned)(g.nox * g.noy) * (unsigned)(zc >> 1);
    };
    auto to_ring = [&](const double* u, int slot, int i) {   // ctoprim (hydro/umuscl.f90:861) of one cell into ring slot `slot`
      double q[NV];
      const double r = fmx(u[0], P.smallr);
      const double oneoverrho = rcp_rn(r);
      q[0] = r;
      double eken;
      q[1] = u[1] * oneoverrho;
      eken = 0.5 * q[1] * q[1];
      q[2] = u[2] * oneoverrho; eken = eken + 0.5 * q[2] * q[2];
      q[3] = u[3] * oneoverrho; eken = eken + 0.5 * q[3] * q[3];
      const double eint = fmx(u[4] * oneoverrho - eken - 0.0, P.smalle);
      q[4] = (P.gamma - 1.0) * r * eint;
      q[1] = q[1] + 0.0;                       // gravity predictor with gloc = 0 (:932-938): -0 -> +0
      q[2] = q[2] + 0.0;
      q[3] = q[3] + 0.0;
      double* qs = qring + (size_t)slot * NQ * PL + i;
#pragma unroll
      for (int n = 0; n < NV; n++) qs[n * PL] = q[n];
      qs[NV * PL] = oneoverrho;
    };
    auto load_plane_direct = [&](int z, int slot) {          // segment prologue: every thread loads its share of a plane
      const unsigned zo = zoff(z);
#pragma unroll
      for (int j = 0; j < NOWN; j++) {
        const int i = tid + j * NT;
        if (i >= PL) break;
        const int xc = wrap_or_clamp(x0 - 2 + i % QX, g.ncx, g.wrapx);
        const int yc = wrap_or_clamp(y0 - 2 + i / QX, g.ncy, g.wrapy);
        const unsigned off = (unsigned)cell_offset<NDIM>(g, xc, yc, 0) + zo;
        double u[NV];
#pragma unroll
        for (int n = 0; n < NV; n++) u[n] = __ldg(a.uin + ((unsigned)n * vstride + off));
        to_ring(u, slot, i);
      }
    };
    // converter threads: rows 0 and BY-1 own the cells i = ctid + 64 j of the staged plane
    constexpr int NCONV = (PL + 63) / 64;
    const bool conv_thread = (ty == 0 || ty == BY - 1);
    const int ctid = tx + (ty == 0 ? 0 : 32);
    unsigned coff[NCONV];
    if (conv_thread) {
#pragma unroll
      for (int j = 0; j < NCONV; j++) {
        const int i = min(ctid + j * 64, PL - 1);
        const int xc = wrap_or_clamp(x0 - 2 + i % QX, g.ncx, g.wrapx);
        const int yc = wrap_or_clamp(y0 - 2 + i / QX, g.ncy, g.wrapy);
        coff[j] = (unsigned)cell_offset<NDIM>(g, xc, yc, 0);
      }
    }
    auto conv_stage = [&](unsigned zo) {
#pragma unroll
      for (int j = 0; j < NCONV; j++) {
        const int i = ctid + j * 64;
        if (i >= PL) break;
        const unsigned off = coff[j] + zo;
#pragma unroll
        for (int n = 0; n < NV; n++) cp_async8(stage + n * PL + i, a.uin + ((unsigned)n * vstride + off));
      }
    };
    auto conv_to_ring = [&](int slot) {
      cp_async_wait_all();
#pragma unroll 1
      for (int i = ctid; i < PL; i += 64) {
        double u[NV];
#pragma unroll
        for (int n = 0; n < NV; n++) u[n] = stage[n * PL + i];
        to_ring(u, slot, i);
      }
    };
    // ---- uslope + trace3d (hydro/umuscl.f90:970, :483) of my cell of the plane in ring slot sc (neighbours in sm1 / sp1): the six
    //      face states; qm_y goes to the exchange buffer `eq`
    const int qx = tx + 1, qy = ty + 1;
    auto predictor = [&](int sm1, int sc, int sp1, double* eq, double* qmx, double* qpx, double* qpy, double* qpz, double* qmz) {
      const double* qc = qring + sc * (NQ * PL) + qy * QX + qx;
      const double* qb_ = qring + sm1 * (NQ * PL) + qy * QX + qx;
      const double* qf_ = qring + sp1 * (NQ * PL) + qy * QX + qx;
      double q[NV], dq[NDIM][NV], t0[NV];
#pragma unroll
      for (int n = 0; n < NV; n++) q[n] = qc[n * PL];
      const double rinv = qc[NV * PL];
      if (SLOPE < 0 && P.slope_type == 3) {      // positivity preserving unsplit slope :1328-1391
#pragma unroll
        for (int n = 0; n < NV; n++) {
          const double* qn = qc + n * PL;
          double vmin = 0, vmax = 0;
          bool first = true;
          for (int cc = -1; cc <= 1; cc++) {
            const double* qz = (cc < 0 ? qb_ : (cc > 0 ? qf_ : qc)) + n * PL;
            for (int aa = -1; aa <= 1; aa++)
              for (int bb = -1; bb <= 1; bb++) {
                const double d = qz[bb * QX + aa] - q[n];
                if (first) { vmin = d; vmax = d; first = false; }
                else { vmin = fmn(vmin, d); vmax = fmx(vmax, d); }
              }
          }
          const double dfx = 0.5 * (qn[1] - qn[-1]);
          const double dfy = 0.5 * (qn[QX] - qn[-QX]);
          const double dfz = 0.5 * (qf_[n * PL] - qb_[n * PL]);
          const double dff = 0.5 * (fabs(dfx) + fabs(dfy) + fabs(dfz));
          double slop;
          if (dff > 0.0) slop = fmn(1.0, fdiv(fmn(fabs(vmin), fabs(vmax)), dff));
          else slop = 1.0;
          dq[0][n] = slop * dfx;
          dq[1][n] = slop * dfy;
          dq[2][n] = slop * dfz;
        }
      } else {
#pragma unroll
        for (int n = 0; n < NV; n++) {
          const double* qn = qc + n * PL;
          dq[0][n] = slope_lcr<NDIM, SLOPE>(qn[-1], q[n], qn[1], P);
          dq[1][n] = slope_lcr<NDIM, SLOPE>(qn[-QX], q[n], qn[QX], P);
          dq[2][n] = slope_lcr<NDIM, SLOPE>(qb_[n * PL], q[n], qf_[n * PL], P);
        }
      }
      double s0[NV];
      trace_sources<NDIM>(q, dq, rinv, s0, P);
#ifdef RGPU_FAST
      const double hdtdx = dtdx * 0.5;
#pragma unroll
      for (int n = 0; n < NV; n++) t0[n] = s0[n] * hdtdx;
#else
#pragma unroll
      for (int n = 0; n < NV; n++) t0[n] = s0[n] * dtdx * 0.5;
#endif
#pragma unroll
      for (int n = 0; n < NV; n++) {             // face states :592-673
        const double hx = 0.5 * dq[0][n], hy = 0.5 * dq[1][n], hz = 0.5 * dq[2][n];
        qmx[n] = q[n] + hx + t0[n];
        qpx[n] = q[n] - hx + t0[n];
        const double qmy = q[n] + hy + t0[n];
        qpy[n] = q[n] - hy + t0[n];
        qmz[n] = q[n] + hz + t0[n];
        qpz[n] = q[n] - hz + t0[n];
        if (n == 0) {
          if (qmx[0] < P.smallr) qmx[0] = q[0];
          if (qpx[0] < P.smallr) qpx[0] = q[0];
          if (qpy[0] < P.smallr) qpy[0] = q[0];
          if (qmz[0] < P.smallr) qmz[0] = q[0];
          if (qpz[0] < P.smallr) qpz[0] = q[0];
          eq[tid] = (qmy < P.smallr) ? q[0] : qmy;
        } else {
          eq[n * NT + tid] = qmy;
        }
      }
    };

    __syncthreads();                           // previous segment done with shared memory
    load_plane_direct(z0 - 2, 0);
    load_plane_direct(z0 - 1, 1);
    load_plane_direct(z0, 2);
    load_plane_direct(z0 + 1, 3);
    const int kbeg = z0 - 1, kend = z1;
    if (conv_thread && kbeg + 3 <= kend + 1) conv_stage(zoff(kbeg + 3));
    __syncthreads();                           // ring planes kbeg-1 .. kbeg+2 complete
    // ring slots of planes k-1, k, k+1, k+2 (rotating); plane offsets of k-1, k and the converters' k+4 (incremental)
    int sA = 0, sB = 1, sC = 2, sD = 3;
    int zc4 = wrap_or_clamp(kbeg + 4, g.ncz, g.wrapz);
    unsigned zo_m1 = zoff(kbeg - 1), zo_0 = zoff(kbeg), zo_1 = zoff(kbeg + 1), zo_2 = zoff(kbeg + 2), zo_3 = zoff(kbeg + 3), zo_4 = zoff(kbeg + 4);
    // prologue: predictor of plane kbeg -> the register set the first iteration consumes
    double qlx[NV], qpx[NV], qpy[NV], qpz[NV], qmzl[NV];       // face states of plane k; qmzl = qm_z of plane k
    {
      double qmx[NV];
      predictor(sA, sB, sC, exq + (kbeg & 1) * (NV * NT), qmx, qpx, qpy, qpz, qmzl);
#pragma unroll
      for (int n = 0; n < NV; n++) qlx[n] = __shfl_up_sync(0xffffffffu, qmx[n], 1);
    }
    bool pend = false;
    for (int k = kbeg; k <= kend; k++) {
      __syncthreads();   // qm_y(k) and F_y(k-1) of every row visible; ring plane k+2 complete; plane k-1's slot is free
      const int par = k & 1;
      double* exq_k = exq + par * (NV * NT);              // qm_y of plane k (read), F_y of plane k (written)
      double* exf_k = exf + par * (NV * NT);
      double* exq_n = exq + (par ^ 1) * (NV * NT);        // qm_y of plane k+1 (written by the predictor)
      const double* exf_p = exf + (par ^ 1) * (NV * NT);  // F_y of plane k-1
      if (own && pend) {                       // y part of the update of plane k-1 (godfine1 :751-792: x, then y, then z)
#pragma unroll
        for (int n = 0; n < NV; n++) {
          double u = carry[(2 * NV + n) * NT + tid];
          u = u + (exf_p[n * NT + tid] - exf_p[n * NT + tid + BX]);
          carry[(2 * NV + n) * NT + tid] = u;
        }
      }
      const bool plane_flux = (k >= z0 && k < z1);
      const bool do_pred = (k < kend);         // the predictor of plane kend+1 is never needed
      double ucur[NV];
#pragma unroll
      for (int n = 0; n < NV; n++) ucur[n] = 0.0;
      if (own && plane_flux) {   // set_unew: unew = uold
        const unsigned off = off_own + zo_0;
#pragma unroll
        for (int n = 0; n < NV; n++) ucur[n] = __ldg(a.uin + ((unsigned)n * vstride + off));
      }
      if (conv_thread) {                       // convert plane k+3 (slot of plane k-1), stage plane k+4
        if (k + 3 <= kend + 1) conv_to_ring(sA);
        if (k + 4 <= kend + 1) conv_stage(zo_4);
      }
      // face states of plane k+1 (set by the predictor below)
      double nqlx[NV], nqpx[NV], nqpy[NV], nqpz[NV], nqmz[NV];
      auto run_predictor = [&]() {
        double qmx[NV];
        predictor(sB, sC, sD, exq_n, qmx, nqpx, nqpy, nqpz, nqmz);
#pragma unroll
        for (int n = 0; n < NV; n++) nqlx[n] = __shfl_up_sync(0xffffffffu, qmx[n], 1);
      };
      if (trace_first && do_pred) run_predictor();

      // ---- solve(k): one inlined solver per direction; rows decide which faces they own (warp uniform) ----
      double fx[NV], fy[NV], fz[NV];
#pragma unroll
      for (int n = 0; n < NV; n++) { fx[n] = 0.0; fy[n] = 0.0; fz[n] = 0.0; }
      const double* e = exq_k + tid - BX;      // qm_y(k) of row ty-1
      const double* cq = carry + tid;          // qm_z of plane k-1
      const bool do_x = row_own && plane_flux;
      const bool do_y = (ty >= 1) && (cy <= g.oy1) && plane_flux;
      const bool do_z = row_own && (k >= z0);
      if (VEC == 1 && do_x && do_y && do_z) {
        V3 QL[NV], QR[NV], FG[NV];
        QL[0] = {qlx[0], e[0 * NT], cq[0 * NT]}; QR[0] = {qpx[0], qpy[0], qpz[0]};
        QL[1] = {qlx[1], e[2 * NT], cq[3 * NT]}; QR[1] = {qpx[1], qpy[2], qpz[3]};
        QL[2] = {qlx[4], e[4 * NT], cq[4 * NT]}; QR[2] = {qpx[4], qpy[4], qpz[4]};
        QL[3] = {qlx[2], e[1 * NT], cq[1 * NT]}; QR[3] = {qpx[2], qpy[1], qpz[1]};
        QL[4] = {qlx[3], e[3 * NT], cq[2 * NT]}; QR[4] = {qpx[3], qpy[3], qpz[2]};
        riemann_v<RIEMANN, V3>(QL, QR, FG, P);
        fx[0] = FG[0].a; fx[1] = FG[1].a; fx[4] = FG[2].a; fx[2] = FG[3].a; fx[3] = FG[4].a;
        fy[0] = FG[0].b; fy[2] = FG[1].b; fy[4] = FG[2].b; fy[1] = FG[3].b; fy[3] = FG[4].b;
        fz[0] = FG[0].c; fz[3] = FG[1].c; fz[4] = FG[2].c; fz[1] = FG[3].c; fz[2] = FG[4].c;
        scale_fluxes<NV>(fx, dt, a.dx, a.inv_dx, a.dx_pow2);
        scale_fluxes<NV>(fy, dt, a.dx, a.inv_dx, a.dx_pow2);
        scale_fluxes<NV>(fz, dt, a.dx, a.inv_dx, a.dx_pow2);
      } else {
        if (do_x) {                            // cmpflxm(...,2,3,4) hydro/umuscl.f90:97
          double ql[NV], qr[NV], fg[NV];
          ql[0] = qlx[0]; ql[1] = qlx[1]; ql[2] = qlx[4]; ql[3] = qlx[2]; ql[4] = qlx[3];
          qr[0] = qpx[0]; qr[1] = qpx[1]; qr[2] = qpx[4]; qr[3] = qpx[2]; qr[4] = qpx[3];
          if (VEC == 0) riemann<NDIM, RIEMANN>(ql, qr, fg, P); else riemann_v<RIEMANN, double>(ql, qr, fg, P);
          fx[0] = fg[0]; fx[1] = fg[1]; fx[4] = fg[2]; fx[2] = fg[3]; fx[3] = fg[4];
          scale_fluxes<NV>(fx, dt, a.dx, a.inv_dx, a.dx_pow2);
        }
        if (do_y) {                            // cmpflxm(...,3,2,4) :120
          double ql[NV], qr[NV], fg[NV];
          ql[0] = e[0 * NT]; ql[1] = e[2 * NT]; ql[2] = e[4 * NT]; ql[3] = e[1 * NT]; ql[4] = e[3 * NT];
          qr[0] = qpy[0]; qr[1] = qpy[2]; qr[2] = qpy[4]; qr[3] = qpy[1]; qr[4] = qpy[3];
          if (VEC == 0) riemann<NDIM, RIEMANN>(ql, qr, fg, P); else riemann_v<RIEMANN, double>(ql, qr, fg, P);
          fy[0] = fg[0]; fy[2] = fg[1]; fy[4] = fg[2]; fy[1] = fg[3]; fy[3] = fg[4];
          scale_fluxes<NV>(fy, dt, a.dx, a.inv_dx, a.dx_pow2);
        }
        if (do_z) {                            // cmpflxm(...,4,2,3) :144; left state = qm_z of the previous plane
          double ql[NV], qr[NV], fg[NV];
          ql[0] = cq[0 * NT]; ql[1] = cq[3 * NT]; ql[2] = cq[4 * NT]; ql[3] = cq[1 * NT]; ql[4] = cq[2 * NT];
          qr[0] = qpz[0]; qr[1] = qpz[3]; qr[2] = qpz[4]; qr[3] = qpz[1]; qr[4] = qpz[2];
          if (VEC == 0) riemann<NDIM, RIEMANN>(ql, qr, fg, P); else riemann_v<RIEMANN, double>(ql, qr, fg, P);
          fz[0] = fg[0]; fz[3] = fg[1]; fz[4] = fg[2]; fz[1] = fg[3]; fz[2] = fg[4];
          scale_fluxes<NV>(fz, dt, a.dx, a.inv_dx, a.dx_pow2);
        }
      }
      if (do_y) {
#pragma unroll
        for (int n = 0; n < NV; n++) exf_k[n * NT + tid] = fy[n];
      }
      // qm_z of plane k becomes the left state of the z faces of plane k+1 (the slot was read by the z solve above)
#pragma unroll
      for (int n = 0; n < NV; n++) carry[n * NT + tid] = qmzl[n];
      double fxr[NV];                          // flux through my +x face comes from lane+1
#pragma unroll
      for (int n = 0; n < NV; n++) fxr[n] = __shfl_down_sync(0xffffffffu, fx[n], 1);
      if (own) {
        if (k > z0) {                          // plane k-1 is complete with the z fluxes: set_uold + courant_fine
          double unew_[NV];
#pragma unroll
          for (int n = 0; n < NV; n++) unew_[n] = carry[(2 * NV + n) * NT + tid] + (carry[(NV + n) * NT + tid] - fz[n]);
          const unsigned off = off_own + zo_m1;
#pragma unroll
          for (int n = 0; n < NV; n++) a.uout[(unsigned)n * vstride + off] = unew_[n];
          double ei;
          const double dtc = cmpdt_cell<NDIM>(unew_, a.dx, P, ei);
          my_dt = dtc < my_dt ? dtc : my_dt;
          my_mass += unew_[0];
          my_etot += unew_[NDIM + 1];
          my_eint += ei;
        }
        if (k >= z0) {
#pragma unroll
          for (int n = 0; n < NV; n++) carry[(NV + n) * NT + tid] = fz[n];
        }
        if (plane_flux) {
#pragma unroll
          for (int n = 0; n < NV; n++) carry[(2 * NV + n) * NT + tid] = ucur[n] + (fx[n] - fxr[n]);
        }
      }
      pend = plane_flux;

      if (!trace_first && do_pred) run_predictor();
      if (do_pred) {
#pragma unroll
        for (int n = 0; n < NV; n++) { qlx[n] = nqlx[n]; qpx[n] = nqpx[n]; qpy[n] = nqpy[n]; qpz[n] = nqpz[n]; qmzl[n] = nqmz[n]; }
      }
      { const int t = sA; sA = sB; sB = sC; sC = sD; sD = t; }
      zo_m1 = zo_0; zo_0 = zo_1; zo_1 = zo_2; zo_2 = zo_3; zo_3 = zo_4;
      zc4 = g.wrapz ? (zc4 + 1 == g.ncz ? 0 : zc4 + 1) : min(zc4 + 1, g.ncz - 1);
      zo_4 = (unsigned)((zc4 & 1) << 2) * (unsigned)g.nslot + (unsigned)(g.nox * g.noy) * (unsigned)(zc4 >> 1);
    }
  }

  // ---- warp-shuffle + shared reduction of the Courant scan partials ----
  my_dt = warp_min(my_dt);
  my_mass = warp_sum(my_mass); my_etot = warp_sum(my_etot); my_eint = warp_sum(my_eint);
  const int w = tid >> 5, l = tid & 31;
  __syncthreads();
  if (l == 0) { red[0][w] = my_dt; red[1][w] = my_mass; red[2][w] = my_etot; red[3][w] = my_eint; }
  __syncthreads();
  if (w == 0) {
    double v0 = 1e300, v1 = 0, v2 = 0, v3 = 0;
    for (int i = l; i < NT / 32; i += 32) { v0 = red[0][i] < v0 ? red[0][i] : v0; v1 += red[1][i]; v2 += red[2][i]; v3 += red[3][i]; }
    v0 = warp_min(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
    if (l == 0 && a.part) {
      const size_t nb = a.part_stride ? (size_t)a.part_stride : (size_t)gridDim.x, c0 = (size_t)a.part_off + blockIdx.x;
      a.part[0 * nb + c0] = v0; a.part[1 * nb + c0] = v1;
      a.part[2 * nb + c0] = v2; a.part[3 * nb + c0] = v3;
    }
  }
}

#ifndef RGPU_HOST_NUMERICS
template <int RIEMANN, int SLOPE, int BY, int VEC, int ORD>
cudaError_t launch_sweep4_v(const SweepArgs& a, int nblocks, cudaStream_t st) {
  constexpr size_t smem = sizeof(double) * Sweep4Smem<BY>::doubles;
  auto kern = sweep4_kernel<RIEMANN, SLOPE, BY, VEC, ORD>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<nblocks, dim3(32, BY, 1), smem, st>>>(a);
  return cudaGetLastError();
}
// variant 4BBOV: BB tile rows, O order (0 solve first, 1 alternating, 2 predictor first), V solver form (0, 1, 2)
template <int RIEMANN>
cudaError_t launch_sweep4(const SweepArgs& a, int nblocks, cudaStream_t st, int variant) {
  if (a.P.slope_type != 1) return cudaErrorInvalidValue;      // tuning variants: slope_type 1 only
  switch (variant) {
#ifdef SWEEP3_TUNING_VARIANTS
    case 41202: return launch_sweep4_v<RIEMANN, 1, 12, 2, 0>(a, nblocks, st);
    case 41212: return launch_sweep4_v<RIEMANN, 1, 12, 2, 1>(a, nblocks, st);
    case 41222: return launch_sweep4_v<RIEMANN, 1, 12, 2, 2>(a, nblocks, st);
    case 41211: return launch_sweep4_v<RIEMANN, 1, 12, 1, 1>(a, nblocks, st);
    case 41012: return launch_sweep4_v<RIEMANN, 1, 10, 2, 1>(a, nblocks, st);
#endif
    default: break;
  }
  return cudaErrorInvalidValue;
}
#endif

}  // namespace rgpu
