// disperse.hpp — the soft-disc relaxation of cell centres shared by
// Tissue2D::Disperse (reference src/Tissue2D.cpp:42-99) and Tissue3D::Disperse2D
// (reference src/Tissue3D.cpp:40-105).  Host-only setup code, outside the accelerated
// path; restated so that identical initial conditions are produced:
// centres drawn with the unseeded drand48() sequence, X then Y per cell; periodic
// minimum image; step 0.01; stop when |dU| <= 1e-6 or after 1e5 iterations.
#ifndef DPM_B200_DISPERSE_HPP
#define DPM_B200_DISPERSE_HPP
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace DPM {
namespace detail {

// radius[i] is the interaction radius of cell i (r0 in 2D, 2*r0 in 3D).
// Returns true when the iteration cap was hit.
//
// The reference visits all N^2 ordered pairs every iteration (src/Tissue2D.cpp:60-86, src/Tissue3D.cpp:62-90), which
// makes its own initialiser unusable beyond ~1e3 cells.  Only OVERLAPPING pairs change anything
// (forces and the energy U are accumulated inside `if (dist <= ri + rj)`), so visiting, for i ascending, a superset of
// i's overlapping partners in ascending j reproduces every floating-point operation of the reference in the same
// order: the result is bit-identical, and so is the iteration at which |dU| <= 1e-6 stops the loop.  The supersets
// are Verlet lists over a periodic bin grid: built with reach + skin, reused until a centre has moved skin/2.
// Boxes with fewer than 3 bins per side keep the all-pairs lists.
inline bool relax_centres(const std::vector<float> &radius, float L, std::vector<float> &X, std::vector<float> &Y) {
  const int n = (int)radius.size();
  X.resize(n);
  Y.resize(n);
  std::vector<float> Fx(n), Fy(n);
  for (int i = 0; i < n; i++) {
    X[i] = drand48() * L;
    Y[i] = drand48() * L;
  }
  float rmax = 0.0f;
  for (int i = 0; i < n; i++) rmax = std::fmax(rmax, radius[i]);
  const double reach = 2.0 * (double)rmax;           // largest ri + rj
  const double skin = 0.25 * reach + 1e-6;           // list margin; lists are valid while every centre moved < skin/2
  const double cut = reach + skin;
  std::vector<int> list_start, list;                 // CSR: partners of i, ascending
  std::vector<float> Xb, Yb;                         // centres at build time
  std::vector<int> bin_start, bin_items, tmp;
  bool have_lists = false;
  auto build_lists = [&]() {
    list_start.assign(n + 1, 0);
    list.clear();
    const bool finite_box = std::isfinite(L) && L > 0.0f;
    const int nb = finite_box ? (int)std::floor((double)L / cut) : 0;
    if (nb < 3 || n < 64) {  // tiny box or tiny tissue: every other cell, ascending (the reference's own loop)
      for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) if (j != i) list.push_back(j);
        list_start[i + 1] = (int)list.size();
      }
    } else {
      auto bin_of = [&](float x) {
        double w = (double)x - (double)L * std::floor((double)x / (double)L);  // wrapped into [0, L)
        int b = (int)(w / (double)L * nb);
        return b < 0 ? 0 : (b >= nb ? nb - 1 : b);
      };
      bin_start.assign(nb * nb + 1, 0);
      for (int i = 0; i < n; i++) bin_start[bin_of(Y[i]) * nb + bin_of(X[i]) + 1]++;
      for (int b = 0; b < nb * nb; b++) bin_start[b + 1] += bin_start[b];
      bin_items.assign(n, 0);
      tmp.assign(bin_start.begin(), bin_start.end() - 1);
      for (int i = 0; i < n; i++) bin_items[tmp[bin_of(Y[i]) * nb + bin_of(X[i])]++] = i;
      std::vector<int> near;
      for (int i = 0; i < n; i++) {
        near.clear();
        const int bx = bin_of(X[i]), by = bin_of(Y[i]);
        for (int oy = -1; oy <= 1; oy++)
          for (int ox = -1; ox <= 1; ox++) {
            const int b = ((by + oy + nb) % nb) * nb + (bx + ox + nb) % nb;
            for (int k = bin_start[b]; k < bin_start[b + 1]; k++) {
              const int j = bin_items[k];
              if (j == i) continue;
              double dx = (double)X[j] - (double)X[i], dy = (double)Y[j] - (double)Y[i];
              dx -= (double)L * std::round(dx / (double)L);
              dy -= (double)L * std::round(dy / (double)L);
              if (dx * dx + dy * dy <= cut * cut) near.push_back(j);
            }
          }
        std::sort(near.begin(), near.end());
        list.insert(list.end(), near.begin(), near.end());
        list_start[i + 1] = (int)list.size();
      }
    }
    Xb = X;
    Yb = Y;
    have_lists = true;
  };
  float oldU = 100, dU = 100;
  int count = 0;
  while (dU > 1e-6) {
    if (have_lists) {  // still valid?  (a non-finite centre also forces a rebuild: the comparison is false for NaN)
      const double lim = 0.45 * skin;
      for (int i = 0; i < n; i++) {
        const double mx = (double)X[i] - (double)Xb[i], my = (double)Y[i] - (double)Yb[i];
        if (!(mx * mx + my * my < lim * lim)) { have_lists = false; break; }
      }
    }
    if (!have_lists) build_lists();
    float U = 0;
    for (int i = 0; i < n; i++) Fx[i] = Fy[i] = 0.0;
    for (int i = 0; i < n; i++) {
      const float xi = X[i], yi = Y[i], ri = radius[i];
      for (int k = list_start[i]; k < list_start[i + 1]; k++) {
        const int j = list[k];
        const float rj = radius[j];
        float dx = X[j] - xi;
        dx -= L * round(dx / L);
        float dy = Y[j] - yi;
        dy -= L * round(dy / L);
        float dist = sqrt(dx * dx + dy * dy);
        if (dist <= (ri + rj)) {
          const float ux = dx / dist, uy = dy / dist;
          const float ftmp = (1.0 - dist / (ri + rj)) / (ri + rj);
          const float fx = ftmp * ux, fy = ftmp * uy;
          Fx[i] -= fx;
          Fy[i] -= fy;
          Fy[j] += fy;
          Fx[j] += fx;
          U += 0.5 * (1 - (dist / (ri + rj)) * (1 - dist / (ri + rj)));
        }
      }
    }
    for (int i = 0; i < n; i++) {
      X[i] += 0.01 * Fx[i];
      Y[i] += 0.01 * Fy[i];
    }
    dU = U - oldU;
    if (dU < 0.0) dU *= -1;
    oldU = U;
    count++;
    if (count > 1e5) return true;
  }
  return false;
}

// The reference's loop verbatim in structure (all ordered pairs): kept as the checker of the binned form above
// (tests/test_host_cpu.py compares the two bit for bit on tissues too large for the real reference to finish quickly).
inline bool relax_centres_allpairs(const std::vector<float> &radius, float L, std::vector<float> &X, std::vector<float> &Y) {
  const int n = (int)radius.size();
  X.resize(n);
  Y.resize(n);
  std::vector<float> Fx(n), Fy(n);
  for (int i = 0; i < n; i++) {
    X[i] = drand48() * L;
    Y[i] = drand48() * L;
  }
  float oldU = 100, dU = 100;
  int count = 0;
  while (dU > 1e-6) {
    float U = 0;
    for (int i = 0; i < n; i++) Fx[i] = Fy[i] = 0.0;
    for (int i = 0; i < n; i++) {
      const float xi = X[i], yi = Y[i], ri = radius[i];
      for (int j = 0; j < n; j++) {
        if (j == i) continue;
        const float rj = radius[j];
        float dx = X[j] - xi;
        dx -= L * round(dx / L);
        float dy = Y[j] - yi;
        dy -= L * round(dy / L);
        float dist = sqrt(dx * dx + dy * dy);
        if (dist <= (ri + rj)) {
          const float ux = dx / dist, uy = dy / dist;
          const float ftmp = (1.0 - dist / (ri + rj)) / (ri + rj);
          const float fx = ftmp * ux, fy = ftmp * uy;
          Fx[i] -= fx;
          Fy[i] -= fy;
          Fy[j] += fy;
          Fx[j] += fx;
          U += 0.5 * (1 - (dist / (ri + rj)) * (1 - dist / (ri + rj)));
        }
      }
    }
    for (int i = 0; i < n; i++) {
      X[i] += 0.01 * Fx[i];
      Y[i] += 0.01 * Fy[i];
    }
    dU = U - oldU;
    if (dU < 0.0) dU *= -1;
    oldU = U;
    count++;
    if (count > 1e5) return true;
  }
  return false;
}

}  // namespace detail
}  // namespace DPM
#endif
