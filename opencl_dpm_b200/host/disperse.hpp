// disperse.hpp — the O(N^2) soft-disc relaxation of cell centres shared by
// Tissue2D::Disperse (reference src/Tissue2D.cpp:42-99) and Tissue3D::Disperse2D
// (reference src/Tissue3D.cpp:40-105).  Host-only setup code, outside the accelerated
// path; restated so that identical initial conditions are produced:
// centres drawn with the unseeded drand48() sequence, X then Y per cell; periodic
// minimum image; step 0.01; stop when |dU| <= 1e-6 or after 1e5 iterations.
#ifndef DPM_B200_DISPERSE_HPP
#define DPM_B200_DISPERSE_HPP
#include <cmath>
#include <cstdlib>
#include <vector>

namespace DPM {
namespace detail {

// radius[i] is the interaction radius of cell i (r0 in 2D, 2*r0 in 3D).
// Returns true when the iteration cap was hit.
inline bool relax_centres(const std::vector<float> &radius, float L, std::vector<float> &X, std::vector<float> &Y) {
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
