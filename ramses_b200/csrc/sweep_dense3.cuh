// sweep_dense3.cuh -- round-2 form of the fused 3-D dense-box Godunov sweep (sm_100a).
//
// Same contract, data layout, tiling and arithmetic as sweep_dense_kernel<3,...> (sweep_dense.cuh; reference
// hydro/godunov_fine.f90:486-911 + hydro/umuscl.f90:22-171), restructured around what the round-1 profile showed
// (profiles/r1_ncu_full_final_*.txt: FP64 pipe 47 % busy, stall reasons `wait` 2.4 / `barrier` 1.1 warps per issue):
//
//  * the x, y and z face solves of a cell run as ONE branch-free solver on a 3-lane value type (hydro_vec.cuh): three
//    independent dependency chains per thread instead of three serial ones (the FP64 pipe needs ILP, not more warps:
//    profiles/microbench/fp64_latency_b200.txt);
//  * two CTA barriers per plane instead of three: the y part of the update of plane k is applied after the first barrier
//    of plane k+1, when every row has published its y fluxes (same operation order: x, then y, then z);
//  * all six face states of a cell are formed right after the predictor, so slopes and source terms are dead before the
//    solver starts (lower register pressure across the solve).
//
// Bit-identical to the round-1 kernel and to the oracle: tests/test_device_numerics_host.py runs this kernel on the CPU
// (emulated launch) against the oracle; tests/test_gpu_parity.py runs it on the GPU.
#pragma once
#include "sweep_dense.cuh"
#include "hydro_vec.cuh"

namespace rgpu {

template <int BY>
struct Sweep3Smem {
  static constexpr int NV = 5, QX = 34, QY = BY + 2, NQ = 6, PL = QX * QY, NT = 32 * BY;
  static constexpr size_t ring = (size_t)3 * NQ * PL;     // primitive variables + 1/rho, planes k-1, k, k+1
  static constexpr size_t stage = (size_t)NV * PL;        // raw conserved state of plane k+2 (cp.async)
  static constexpr size_t exq = (size_t)NV * NT;          // qm_y of plane k      (read by row+1)
  static constexpr size_t exf = (size_t)NV * NT;          // Fy of plane k        (read by row-1 after the next barrier)
  static constexpr size_t carry = (size_t)3 * NV * NT;    // per-thread: qm_z, Fz, partial update of the pending plane
  static constexpr size_t doubles = ring + stage + exq + exf + carry;
};

// VEC: 0 = three scalar solves with the branching solvers of hydro_device.cuh (round-1 arithmetic path, new loop);
//      1 = one 3-lane solve (hydro_vec.cuh); 2 = three scalar solves with the branch-free forms of hydro_vec.cuh.
// CONVW: the two halo-row warps (ty = 0 and ty = BY-1, idle while the owned rows solve their faces) stage plane k+3 and convert
//        plane k+2 to primitive variables during the solve phase of plane k; the other warps never touch the staging buffer.
template <int RIEMANN, int SLOPE, int BY, int MINB, int VEC, bool CONVW = false>
__global__ void __launch_bounds__(32 * BY, MINB) sweep3_kernel(const SweepArgs a) {
  using S = Sweep3Smem<BY>;
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
  double* carry = exf + S::exf;                // [0..NV): qm_z   [NV..2NV): Fz   [2NV..3NV): partial update
  __shared__ double red[4][NT / 32];

  const DenseGeom& g = a.g;
  const Phys& P = a.P;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * BX + tx;
  const double dt = a.dt_dev ? *a.dt_dev : a.dt_val;
  const double dtdx = dt / a.dx;               // trace3d: dtdx = dt/dx (hydro/umuscl.f90:516)
  // 32-bit ELEMENT indices (host side guarantees nvar*8*nslot < 2^32): one IMAD.WIDE per access instead of 64-bit adds
  const unsigned vstride = 8u * (unsigned)g.nslot;

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
    const bool row_own = (ty >= 1) && (ty <= BY - 2) && (cy < g.oy1);    // warp uniform (a warp is one row)
    const bool own = col_own && row_own;

    constexpr int NOWN = (PL + NT - 1) / NT;
    unsigned offxy[NOWN];
#pragma unroll
    for (int j = 0; j < NOWN; j++) {
      const int i = tid + j * NT;
      const int qx_ = i % QX, qy_ = i / QX;
      const int xc = wrap_or_clamp(x0 - 2 + qx_, g.ncx, g.wrapx);
      const int yc = wrap_or_clamp(y0 - 2 + qy_, g.ncy, g.wrapy);
      offxy[j] = (unsigned)cell_offset<NDIM>(g, xc, yc, 0);
    }
    const unsigned off_own = (unsigned)cell_offset<NDIM>(g, wrap_or_clamp(cx, g.ncx, g.wrapx), wrap_or_clamp(cy, g.ncy, g.wrapy), 0);
    auto zoff = [&](int z) -> unsigned {
      const int zc = wrap_or_clamp(z, g.ncz, g.wrapz);
      return (unsigned)((zc & 1) << 2) * (unsigned)g.nslot + (unsigned)(g.nox * g.noy) * (unsigned)(zc >> 1);
    };
    // ctoprim (hydro/umuscl.f90:861) of one cell into ring slot `slot`
    auto to_ring = [&](const double* u, int slot, int i) {
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
    auto load_plane_direct = [&](int z, int slot) {
      const unsigned zo = zoff(z);
#pragma unroll
      for (int j = 0; j < NOWN; j++) {
        const int i = tid + j * NT;
        if (i >= PL) break;
        const unsigned off = offxy[j] + zo;
        double u[NV];
#pragma unroll
        for (int n = 0; n < NV; n++) u[n] = __ldg(a.uin + ((unsigned)n * vstride + off));
        to_ring(u, slot, i);
      }
    };
    auto stage_plane_async_off = [&](unsigned zo) {
#pragma unroll
      for (int j = 0; j < NOWN; j++) {
        const int i = tid + j * NT;
        if (i >= PL) break;
        const unsigned off = offxy[j] + zo;
#pragma unroll
        for (int n = 0; n < NV; n++) cp_async8(stage + n * PL + i, a.uin + ((unsigned)n * vstride + off));
      }
    };
    auto stage_plane_async = [&](int z) { stage_plane_async_off(zoff(z)); };
    auto stage_to_ring = [&](int slot) {
      cp_async_wait_all();
      for (int i = tid; i < PL; i += NT) {
        double u[NV];
#pragma unroll
        for (int n = 0; n < NV; n++) u[n] = stage[n * PL + i];
        to_ring(u, slot, i);
      }
    };

    // converter threads (CONVW): 64 threads (rows 0 and BY-1) own the cells i = ctid + 64 j of the staged plane
    constexpr int NCONV = (PL + 63) / 64;
    const bool conv_thread = CONVW && (ty == 0 || ty == BY - 1);
    const int ctid = tx + (ty == 0 ? 0 : 32);
    unsigned coff[CONVW ? NCONV : 1];
    if (conv_thread) {
#pragma unroll
      for (int j = 0; j < (CONVW ? NCONV : 1); j++) {
        const int i = min(ctid + j * 64, PL - 1);
        const int xc = wrap_or_clamp(x0 - 2 + i % QX, g.ncx, g.wrapx);
        const int yc = wrap_or_clamp(y0 - 2 + i / QX, g.ncy, g.wrapy);
        coff[j] = (unsigned)cell_offset<NDIM>(g, xc, yc, 0);
      }
    }
    auto conv_stage = [&](unsigned zo) {       // cp.async of the converter's cells of one plane
#pragma unroll
      for (int j = 0; j < (CONVW ? NCONV : 1); j++) {
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

    __syncthreads();                           // previous segment done with shared memory
    load_plane_direct(z0 - 2, 0);
    load_plane_direct(z0 - 1, 1);
    if (CONVW) {
      load_plane_direct(z0, 2);
      if (conv_thread) conv_stage(zoff(z0 + 1));
    } else {
      stage_plane_async(z0);
    }
    const int kbeg = z0 - 1, kend = z1;
    bool pend = false;                         // the x part of the update of the previous plane waits for its y and z parts
    // ring slots of planes k-1, k, k+1 rotate; the plane offsets of k-1, k, k+2 advance incrementally (no division, no
    // modulo inside the plane loop)
    int sm1 = 0, sc = 1, sp1 = 2;
    int zc3 = wrap_or_clamp(kbeg + 3, g.ncz, g.wrapz);
    unsigned zo_m1 = zoff(kbeg - 1), zo_0 = zoff(kbeg), zo_1 = zoff(kbeg + 1), zo_2 = zoff(kbeg + 2), zo_3 = zoff(kbeg + 3);
    for (int k = kbeg; k <= kend; k++) {
      if (!CONVW) stage_to_ring(sp1);
      __syncthreads();                         // ring plane k+1 complete; every row has published Fy(k-1)
      if (!CONVW && k < kend) stage_plane_async_off(zo_2);
      if (own && pend) {                       // y part of the update of plane k-1 (godfine1 :751-792: x, then y, then z)
#pragma unroll
        for (int n = 0; n < NV; n++) {
          double u = carry[(2 * NV + n) * NT + tid];
          u = u + (exf[n * NT + tid] - exf[n * NT + tid + BX]);
          carry[(2 * NV + n) * NT + tid] = u;
        }
      }
      const bool plane_flux = (k >= z0 && k < z1);

      // ---- uslope + trace3d of my cell (hydro/umuscl.f90:970, :483): all six face states ----
      const int qx = tx + 1, qy = ty + 1;
      const double* qc = qring + sc * (NQ * PL) + qy * QX + qx;
      double qmx[NV], qpx[NV], qpy[NV], qpz[NV], qmz[NV];
      {
        // every thread traces the ring cell under it: the ring holds a valid primitive state at all 34 x (BY+2) positions
        // (loaded with wrap / clamp), so threads beyond the owned range need no special case (their results are never used)
        double q[NV], dq[NDIM][NV], t0[NV];
        {
#pragma unroll
          for (int n = 0; n < NV; n++) q[n] = qc[n * PL];
          const double rinv = qc[NV * PL];
          const double* qb_ = qring + sm1 * (NQ * PL) + qy * QX + qx;
          const double* qf_ = qring + sp1 * (NQ * PL) + qy * QX + qx;
          if (SLOPE < 0 && P.slope_type == 3) {
            // positivity preserving unsplit slope :1328-1391
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
        }
        // face states :592-673
#pragma unroll
        for (int n = 0; n < NV; n++) {
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
            exq[tid] = (qmy < P.smallr) ? q[0] : qmy;
          } else {
            exq[n * NT + tid] = qmy;
          }
        }
      }
      // left state of my -x face comes from lane-1 (a warp is one x-row)
      double qlx[NV];
#pragma unroll
      for (int n = 0; n < NV; n++) qlx[n] = __shfl_up_sync(0xffffffffu, qmx[n], 1);
      double ucur[NV];
#pragma unroll
      for (int n = 0; n < NV; n++) ucur[n] = 0.0;
      if (own && plane_flux) {   // set_unew: unew = uold; issued early so the latency hides under the Riemann solves
        const unsigned off = off_own + zo_0;
#pragma unroll
        for (int n = 0; n < NV; n++) ucur[n] = __ldg(a.uin + ((unsigned)n * vstride + off));
      }
      __syncthreads();                         // qm_y of every row is visible

      if (CONVW && conv_thread) {              // the solve phase of the owned rows: convert plane k+2, stage plane k+3
        if (k + 2 <= kend + 1) conv_to_ring(sm1);          // slot of plane k-1: nobody reads it after the barrier above
        if (k + 3 <= kend + 1) conv_stage(zo_3);
      }
      double fx[NV], fy[NV], fz[NV];
#pragma unroll
      for (int n = 0; n < NV; n++) { fx[n] = 0.0; fy[n] = 0.0; fz[n] = 0.0; }
      const double* e = exq + tid - BX;        // qm_y of row ty-1 (rows ty >= 1 only)
      const double* cq = carry + tid;          // qm_z of plane k-1
      // which faces this ROW solves (warp uniform: a warp is one row): -x and -z faces of the owned rows, the -y face of rows
      // 1 .. first row above the owned ones.  One inlined solver per direction and no other copy (I-cache)
      const bool do_x = row_own && plane_flux;
      const bool do_y = (ty >= 1) && (cy <= g.oy1) && plane_flux;
      const bool do_z = row_own && (k >= z0);
      if (VEC == 1 && do_x && do_y && do_z) {
        // cmpflxm permutations (hydro/umuscl.f90:97,120,144): lane a = x (2,3,4), b = y (3,2,4), c = z (4,2,3)
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
        if (do_z) {                            // cmpflxm(...,4,2,3) :144; left state carried from the previous plane
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
        for (int n = 0; n < NV; n++) exf[n * NT + tid] = fy[n];
      }
      // qm_z of this plane becomes the left state of the next plane's z faces (after the z solve read the old one)
#pragma unroll
      for (int n = 0; n < NV; n++) carry[n * NT + tid] = qmz[n];
      // flux through my +x face comes from lane+1
      double fxr[NV];
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
      { const int t = sm1; sm1 = sc; sc = sp1; sp1 = t; }
      zo_m1 = zo_0; zo_0 = zo_1; zo_1 = zo_2; zo_2 = zo_3;
      zc3 = g.wrapz ? (zc3 + 1 == g.ncz ? 0 : zc3 + 1) : min(zc3 + 1, g.ncz - 1);
      zo_3 = (unsigned)((zc3 & 1) << 2) * (unsigned)g.nslot + (unsigned)(g.nox * g.noy) * (unsigned)(zc3 >> 1);
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
template <int RIEMANN, int SLOPE, int BY, int MINB, int VEC, bool CONVW = false>
cudaError_t launch_sweep3_v(const SweepArgs& a, int nblocks, cudaStream_t st) {
  constexpr size_t smem = sizeof(double) * Sweep3Smem<BY>::doubles;
  auto kern = sweep3_kernel<RIEMANN, SLOPE, BY, MINB, VEC, CONVW>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // MINB > 1: all of the SM's unified L1/shared storage as shared memory so that MINB CTAs are resident together (costs L1:
    // measured -9 % on the one-CTA variants, profiles/r2_tune_sweep.md)
    if (MINB > 1) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return e;
    }
    configured = true;
  }
  kern<<<nblocks, dim3(32, BY, 1), smem, st>>>(a);
  return cudaGetLastError();
}
// variant = 100*BY + 10*MINB + VEC (MINB digit 9: one CTA per SM with converter warps).  The product default is SWEEP3_DEFAULT_VARIANT; the others exist for the tuning runs
// recorded under profiles/ (RGPU_SWEEP=<variant> at bind time) and are compiled for slope_type 1 only.
#ifndef SWEEP3_DEFAULT_VARIANT
#define SWEEP3_DEFAULT_VARIANT 1292
#endif
constexpr int sweep3_by_of(int variant) { return (variant % 10000) / 100; }   // 4BBOV: sweep4_kernel (sweep_dense4.cuh), O = order, V = solver form
template <int RIEMANN, int SLOPE>
cudaError_t launch_sweep3_s(const SweepArgs& a, int nblocks, cudaStream_t st, int variant) {
  switch (variant) {
    case 1292: return launch_sweep3_v<RIEMANN, SLOPE, 12, 1, 2, true>(a, nblocks, st);   // 9 = converter warps, one CTA per SM
#ifdef SWEEP3_TUNING_VARIANTS
    case 1212: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 12, 1, 2>(a, nblocks, st); break;
    case 1211: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 12, 1, 1>(a, nblocks, st); break;
    case 1412: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 14, 1, 2>(a, nblocks, st); break;
    case 1492: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 14, 1, 2, true>(a, nblocks, st); break;
    case 1692: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 16, 1, 2, true>(a, nblocks, st); break;
    case 1291: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 12, 1, 1, true>(a, nblocks, st); break;
    case 1210: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 12, 1, 0>(a, nblocks, st); break;
    case 811: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 8, 1, 1>(a, nblocks, st); break;
    case 821: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 8, 2, 1>(a, nblocks, st); break;
    case 820: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 8, 2, 0>(a, nblocks, st); break;
    case 822: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 8, 2, 2>(a, nblocks, st); break;
    case 1612: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 16, 1, 2>(a, nblocks, st); break;
    case 812: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 8, 1, 2>(a, nblocks, st); break;
    case 1611: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 16, 1, 1>(a, nblocks, st); break;
    case 1610: if (SLOPE == 1) return launch_sweep3_v<RIEMANN, SLOPE == 1 ? 1 : SLOPE, 16, 1, 0>(a, nblocks, st); break;
#endif
    default: break;
  }
  return cudaErrorInvalidValue;
}
template <int RIEMANN>
cudaError_t launch_sweep3(const SweepArgs& a, int nblocks, cudaStream_t st, int variant) {
  if (a.P.slope_type == 1) return launch_sweep3_s<RIEMANN, 1>(a, nblocks, st, variant);
  if (a.P.slope_type == 2) return launch_sweep3_s<RIEMANN, 2>(a, nblocks, st, variant);
  return launch_sweep3_s<RIEMANN, -1>(a, nblocks, st, variant);
}
#endif

}  // namespace rgpu
