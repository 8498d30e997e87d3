// amr_schedules.h -- host-only builders of the deterministic coarse-reflux schedules of AMR mode (no CUDA in here; used by
// rgpu_api.cu at bind time and by the CPU test harness).  godfine1 adds the fluxes / EMFs of an oct's outer faces and corners
// into cells of the coarser level, several octs into the same cell; the reference's result depends on the order of these
// floating-point additions, so the schedules list, per target, the contributions in the order the reference visits them:
// batch of nvector octs, then the loop nest of the routine.  One device thread per target then reproduces the sum bit for bit.
#pragma once
#include <algorithm>
#include <vector>

namespace rgpu {

// Euler fluxes (hydro/godunov_fine.f90:798-908; mhd/godunov_fine.f90:1030-1168): batch, direction, left then right, face, oct of
// the batch whose neighbour father cell is a leaf (son(nbor) == 0).  Out: target cells, start offsets, packed sources
// (oct position in the active list << 6 | side << 3 | face); src_unsorted = the same sources in visiting order.
// nbor = nbor(1:ngridmax,1:2*ndim), son = son(1:ncell), both 0-based views of the Fortran arrays.
inline void build_reflux_schedule(int ndim, int nvector, int nact, const int* igrid_active, const int* nbor, const int* son, int ngridmax,
                                  std::vector<int>& cells, std::vector<int>& start, std::vector<int>& srcs, std::vector<int>& src_unsorted) {
  const int nv = std::max(1, nvector), NSF = 1 << (ndim - 1);
  std::vector<int> tgt;
  std::vector<int>& src = src_unsorted;
  src.clear(); cells.clear(); start.clear(); srcs.clear();
  for (int i0 = 0; i0 < nact; i0 += nv) {
    const int ng = std::min(nv, nact - i0);
    for (int d = 0; d < ndim; d++)
      for (int s = 0; s < 2; s++)
        for (int f = 0; f < NSF; f++)
          for (int i = 0; i < ng; i++) {
            const int ig = igrid_active[i0 + i];
            const int nb = nbor[(size_t)(2 * d + s) * ngridmax + ig - 1];
            if (nb > 0 && son[nb - 1] == 0) { tgt.push_back(nb); src.push_back(((i0 + i) << 6) | ((2 * d + s) << 3) | f); }
          }
  }
  std::vector<int> order(tgt.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tgt[x] < tgt[y]; });   // stable: visiting order kept per cell
  for (size_t k = 0; k < order.size(); k++) {
    if (k == 0 || tgt[order[k]] != tgt[order[k - 1]]) { cells.push_back(tgt[order[k]]); start.push_back((int)k); }
    srcs.push_back(src[order[k]]);
  }
  start.push_back((int)order.size());
}

// Corner EMFs of the MHD build, NDIM = 2 (mhd/godunov_fine.f90:1176-1270): batch, edge X0Y0 / X0Y1 / X1Y1 / X1Y0, oct of the
// batch, the statements of the edge in source order.  nfc = the 3x3 father cells of every oct (get3cubefather),
// [nact][9], index i1 + 3*j1.  Out: per target (cell, variable 0-based: 5 = B_x left, 6 = B_y left, 8 = B_x right, 9 = B_y right)
// the contribution codes  oct << 5 | edge << 3 | (weight one half ? 4 : 0) | (dflux*half ? 2 : 0) | (subtract ? 1 : 0).
inline void build_emf_schedule_2d(int nvector, int nact, const int* nfc, const int* son, std::vector<int>& cells, std::vector<int>& vars,
                                  std::vector<int>& start, std::vector<int>& codes) {
  auto sonh = [&](int c) { return c > 0 ? son[c - 1] : 0; };
  static const int fo[4][3][2] = {{{1, 0}, {0, 0}, {0, 1}}, {{0, 1}, {0, 2}, {1, 2}}, {{1, 2}, {2, 2}, {2, 1}}, {{2, 1}, {2, 0}, {1, 0}}};
  // per edge: (which of ind_father 1,2,3, variable, subtract, half) in statement order; the last two only when all three are leaves
  struct St { int b, var, minus, half; };
  static const St stm[4][6] = {
      {{0, 5, 0, 0}, {1, 8, 0, 0}, {1, 9, 1, 0}, {2, 6, 1, 0}, {2, 8, 1, 1}, {0, 9, 0, 1}},
      {{0, 9, 1, 0}, {1, 6, 1, 0}, {1, 8, 1, 0}, {2, 5, 1, 0}, {2, 6, 0, 1}, {0, 8, 0, 1}},
      {{0, 8, 1, 0}, {1, 5, 1, 0}, {1, 6, 0, 0}, {2, 9, 0, 0}, {2, 5, 0, 1}, {0, 6, 1, 1}},
      {{0, 6, 0, 0}, {1, 9, 0, 0}, {1, 5, 0, 0}, {2, 8, 0, 0}, {2, 9, 1, 1}, {0, 5, 1, 1}}};
  std::vector<long long> key;
  std::vector<int> code;
  const int nv = std::max(1, nvector);
  cells.clear(); vars.clear(); start.clear(); codes.clear();
  for (int i0 = 0; i0 < nact; i0 += nv) {
    const int ng = std::min(nv, nact - i0);
    for (int e = 0; e < 4; e++)
      for (int i = 0; i < ng; i++) {
        const int* f = &nfc[(size_t)(i0 + i) * 9];
        const int b[3] = {f[fo[e][0][0] + 3 * fo[e][0][1]], f[fo[e][1][0] + 3 * fo[e][1][1]], f[fo[e][2][0] + 3 * fo[e][2][1]]};
        const int s1 = sonh(b[0]), s2 = sonh(b[1]), s3 = sonh(b[2]);
        if (s1 > 0 && s3 > 0) continue;
        const int whalf = (s1 > 0 || s2 > 0 || s3 > 0) ? 1 : 0;
        const bool all_leaf = s1 == 0 && s2 == 0 && s3 == 0;
        for (int k = 0; k < (all_leaf ? 6 : 4); k++) {
          const St& q = stm[e][k];
          if (b[q.b] <= 0) continue;
          key.push_back((long long)b[q.b] * 16 + q.var);
          code.push_back(((i0 + i) << 5) | (e << 3) | (whalf << 2) | (q.half << 1) | q.minus);
        }
      }
  }
  std::vector<int> order(key.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
  for (size_t k = 0; k < order.size(); k++) {
    if (k == 0 || key[order[k]] != key[order[k - 1]]) {
      cells.push_back((int)(key[order[k]] / 16)); vars.push_back((int)(key[order[k]] % 16)); start.push_back((int)k);
    }
    codes.push_back(code[order[k]]);
  }
  start.push_back((int)order.size());
}

// Edge EMFs of the MHD build, NDIM = 3 (mhd/godunov_fine.f90:1176-1455): the twelve edges of an oct, in the order of the source
// (EMFz: X0Y0, X0Y1, X1Y1, X1Y0; EMFx: Y0Z0, Y0Z1, Y1Z1, Y1Z0; EMFy: X0Z0, X0Z1, X1Z1, X1Z0).  Per edge: the offsets (i1,j1,k1) of
// ind_father1..3 in the 3x3x3 father-cell cube, the EMF component, the (i3,j3,k3) of its two emf entries (the edge spans two fine
// cells), four updates {ind_father 1..3, variable 1-based (6..8 left faces, 9..11 right faces), sign} and the two extra updates
// applied with dflux/2 when all three father cells are leaves.
struct MhdEdge3Host { int f[3][3]; int dir; int c[2][3]; int upd[4][3]; int leaf[2][3]; };
inline const MhdEdge3Host* mhd_edges3() {
  static const MhdEdge3Host E[12] = {
      {{{1, 0, 1}, {0, 0, 1}, {0, 1, 1}}, 2, {{1, 1, 1}, {1, 1, 2}}, {{1, 6, +1}, {2, 9, +1}, {2, 10, -1}, {3, 7, -1}}, {{3, 9, -1}, {1, 10, +1}}},
      {{{0, 1, 1}, {0, 2, 1}, {1, 2, 1}}, 2, {{1, 3, 1}, {1, 3, 2}}, {{1, 10, -1}, {2, 7, -1}, {2, 9, -1}, {3, 6, -1}}, {{3, 7, +1}, {1, 9, +1}}},
      {{{1, 2, 1}, {2, 2, 1}, {2, 1, 1}}, 2, {{3, 3, 1}, {3, 3, 2}}, {{1, 9, -1}, {2, 6, -1}, {2, 7, +1}, {3, 10, +1}}, {{3, 6, +1}, {1, 7, -1}}},
      {{{2, 1, 1}, {2, 0, 1}, {1, 0, 1}}, 2, {{3, 1, 1}, {3, 1, 2}}, {{1, 7, +1}, {2, 10, +1}, {2, 6, +1}, {3, 9, +1}}, {{3, 10, -1}, {1, 6, -1}}},
      {{{1, 1, 0}, {1, 0, 0}, {1, 0, 1}}, 0, {{1, 1, 1}, {2, 1, 1}}, {{1, 7, +1}, {2, 10, +1}, {2, 11, -1}, {3, 8, -1}}, {{1, 11, +1}, {3, 10, -1}}},
      {{{1, 0, 1}, {1, 0, 2}, {1, 1, 2}}, 0, {{1, 1, 3}, {2, 1, 3}}, {{1, 11, -1}, {2, 8, -1}, {2, 10, -1}, {3, 7, -1}}, {{1, 10, +1}, {3, 8, +1}}},
      {{{1, 1, 2}, {1, 2, 2}, {1, 2, 1}}, 0, {{1, 3, 3}, {2, 3, 3}}, {{1, 10, -1}, {2, 7, -1}, {2, 8, +1}, {3, 11, +1}}, {{3, 7, +1}, {1, 8, -1}}},
      {{{1, 2, 1}, {1, 2, 0}, {1, 1, 0}}, 0, {{1, 3, 1}, {2, 3, 1}}, {{1, 8, +1}, {2, 11, +1}, {2, 7, +1}, {3, 10, +1}}, {{3, 11, -1}, {1, 7, -1}}},
      {{{1, 1, 0}, {0, 1, 0}, {0, 1, 1}}, 1, {{1, 1, 1}, {1, 2, 1}}, {{1, 6, -1}, {2, 9, -1}, {2, 11, +1}, {3, 8, +1}}, {{3, 9, +1}, {1, 11, -1}}},
      {{{0, 1, 1}, {0, 1, 2}, {1, 1, 2}}, 1, {{1, 1, 3}, {1, 2, 3}}, {{1, 11, +1}, {2, 8, +1}, {2, 9, +1}, {3, 6, +1}}, {{3, 8, -1}, {1, 9, -1}}},
      {{{1, 1, 2}, {2, 1, 2}, {2, 1, 1}}, 1, {{3, 1, 3}, {3, 2, 3}}, {{1, 9, +1}, {2, 6, +1}, {2, 8, -1}, {3, 11, -1}}, {{3, 6, -1}, {1, 8, +1}}},
      {{{2, 1, 1}, {2, 1, 0}, {1, 1, 0}}, 1, {{3, 1, 1}, {3, 2, 1}}, {{1, 8, -1}, {2, 11, -1}, {2, 6, -1}, {3, 9, -1}}, {{3, 11, +1}, {1, 6, +1}}}};
  return E;
}
// nfc [nact][27] (index i1 + 3*j1 + 9*k1).  Out: per target (cell, variable 0-based 5..10) the contribution codes
// oct << 7 | edge << 3 | (weight one half ? 4 : 0) | (dflux*half ? 2 : 0) | (subtract ? 1 : 0), in the reference's visiting order
// (batch, edge, oct of the batch, statement).
inline void build_emf_schedule_3d(int nvector, int nact, const int* nfc, const int* son, std::vector<int>& cells, std::vector<int>& vars,
                                  std::vector<int>& start, std::vector<int>& codes) {
  auto sonh = [&](int c) { return c > 0 ? son[c - 1] : 0; };
  const MhdEdge3Host* E = mhd_edges3();
  std::vector<long long> key;
  std::vector<int> code;
  const int nv = std::max(1, nvector);
  cells.clear(); vars.clear(); start.clear(); codes.clear();
  for (int i0 = 0; i0 < nact; i0 += nv) {
    const int ng = std::min(nv, nact - i0);
    for (int e = 0; e < 12; e++)
      for (int i = 0; i < ng; i++) {
        const int* f = &nfc[(size_t)(i0 + i) * 27];
        int b[4];
        for (int q = 0; q < 3; q++) b[q + 1] = f[E[e].f[q][0] + 3 * E[e].f[q][1] + 9 * E[e].f[q][2]];
        const int s1 = sonh(b[1]), s2 = sonh(b[2]), s3 = sonh(b[3]);
        if (s1 > 0 && s3 > 0) continue;
        const int whalf = (s1 > 0 || s2 > 0 || s3 > 0) ? 1 : 0;
        const bool all_leaf = s1 == 0 && s2 == 0 && s3 == 0;
        for (int k = 0; k < (all_leaf ? 6 : 4); k++) {
          const int* u = k < 4 ? E[e].upd[k] : E[e].leaf[k - 4];
          if (b[u[0]] <= 0) continue;
          key.push_back((long long)b[u[0]] * 16 + (u[1] - 1));
          code.push_back(((i0 + i) << 7) | (e << 3) | (whalf << 2) | ((k >= 4 ? 1 : 0) << 1) | (u[2] > 0 ? 0 : 1));
        }
      }
  }
  std::vector<int> order(key.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key[x] < key[y]; });
  for (size_t k = 0; k < order.size(); k++) {
    if (k == 0 || key[order[k]] != key[order[k - 1]]) {
      cells.push_back((int)(key[order[k]] / 16)); vars.push_back((int)(key[order[k]] % 16)); start.push_back((int)k);
    }
    codes.push_back(code[order[k]]);
  }
  start.push_back((int)order.size());
}

}  // namespace rgpu
