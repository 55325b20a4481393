// Tissue2D.cpp — host side of the 2D hot path over the C ABI.
// Mirrors DPM::Tissue2D of the reference (src/Tissue2D.cpp): constructor (:15-33),
// Disperse (:42-114) and CLEulerUpdate (:116-241), with the OpenCL build / 9 buffers /
// 6 kernels / step loop (:142-233) replaced by one dpm2d_euler_update call.
#include <cmath>
#include <iostream>
#include <stdexcept>

#include "Tissue.hpp"
#include "disperse.hpp"
#include "dpm_b200.h"

namespace DPM {

struct DeviceHandle2D {
  dpm2d_t *h = nullptr;
  int ncells = 0, max_nv = 0;
  ~DeviceHandle2D() {
    if (h) dpm2d_destroy(h);
  }
};

Tissue2D::Tissue2D(std::vector<Cell2D> inputCells, float packingFraction) {
  cells = inputCells;
  NCELLS = cells.size();
  PBC = true;
  Kre = 1.0f;
  Kat = 0.0f;
  phi0 = packingFraction;
  float totalArea = 0.0f;
  maxNV = cells[0].NV;
  for (auto &c : cells) {
    totalArea += c.GetArea();
    if (c.NV > (unsigned int)maxNV) maxNV = c.NV;
  }
  L = sqrt(totalArea) / phi0;
}

void Tissue2D::Disperse() {
  std::vector<float> radius(NCELLS), X, Y;
  for (int i = 0; i < NCELLS; i++) radius[i] = cells[i].r0;
  detail::relax_centres(radius, L, X, Y);
  // rebuild each regular polygon around its relaxed centre (reference :105-113)
  for (int i = 0; i < NCELLS; i++)
    for (int j = 0; j < (int)cells[i].NV; j++) {
      cells[i].Verticies[j][0] = cells[i].r0 * (cos(2.0 * M_PI * (j + 1) / cells[i].NV)) + X[i];
      cells[i].Verticies[j][1] = cells[i].r0 * (sin(2.0 * M_PI * (j + 1) / cells[i].NV)) + Y[i];
    }
}

void Tissue2D::CLEulerUpdate(int nsteps, float dt) {
  // pack, padded to maxNV per cell with zeros (reference :117-140)
  maxNV = cells.empty() ? 0 : (int)cells[0].NV;
  for (auto &c : cells) maxNV = std::max(maxNV, (int)c.NV);
  const size_t stride = (size_t)maxNV * 2;
  std::vector<float> verts(stride * NCELLS, 0.0f), forces(stride * NCELLS, 0.0f);
  std::vector<float> Ka(NCELLS), Kl(NCELLS), Kb(NCELLS), l0(NCELLS), a0(NCELLS), r0(NCELLS);
  std::vector<int32_t> NV(NCELLS);
  for (int ci = 0; ci < NCELLS; ci++) {
    const Cell2D &c = cells[ci];
    Ka[ci] = c.Ka; Kl[ci] = c.Kl; Kb[ci] = c.Kb; l0[ci] = c.l0; a0[ci] = c.a0; r0[ci] = c.r0; NV[ci] = (int32_t)c.NV;
    for (unsigned int vi = 0; vi < c.NV; vi++) {
      verts[ci * stride + 2 * vi] = c.Verticies[vi][0];
      verts[ci * stride + 2 * vi + 1] = c.Verticies[vi][1];
    }
  }
  if (!dev) dev = std::make_shared<DeviceHandle2D>();
  auto err = [] { char b[1024]; dpm_last_error(b, sizeof b); return std::string(b); };
  if (!dev->h || dev->ncells != NCELLS || dev->max_nv != maxNV) {
    if (dev->h) { dpm2d_destroy(dev->h); dev->h = nullptr; }
    if (dpm2d_create(&dev->h, 0, NCELLS, maxNV) != DPM_OK) throw std::runtime_error(err());
    dev->ncells = NCELLS;
    dev->max_nv = maxNV;
  }
  const int rc = dpm2d_euler_update(dev->h, verts.data(), forces.data(), NV.data(), Ka.data(), Kl.data(), Kb.data(),
                                    a0.data(), l0.data(), r0.data(), nsteps, dt, Kre, Kat, (int)PBC, L, nullptr);
  if (rc == DPM_ERR_INVALID_ARGUMENT) throw std::invalid_argument(err());
  if (rc != DPM_OK) throw std::runtime_error(err());
  // unpack real vertices only (reference :235-240)
  for (int ci = 0; ci < NCELLS; ci++)
    for (unsigned int vi = 0; vi < cells[ci].NV; vi++) {
      cells[ci].Verticies[vi] = {verts[ci * stride + 2 * vi], verts[ci * stride + 2 * vi + 1]};
      cells[ci].Forces[vi] = {forces[ci * stride + 2 * vi], forces[ci * stride + 2 * vi + 1]};
    }
}

}  // namespace DPM
