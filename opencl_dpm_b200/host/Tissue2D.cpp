// Tissue2D.cpp — host side of the 2D hot path over the C ABI.
// Mirrors DPM::Tissue2D of the reference (src/Tissue2D.cpp): constructor (:15-33),
// Disperse (:42-114) and CLEulerUpdate (:116-241), with the OpenCL build / 9 buffers /
// 6 kernels / step loop (:142-233) replaced by one dpm2d_euler_update call.
#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <string>

#include "Tissue.hpp"
#include "disperse.hpp"
#include "dpm_b200.h"

namespace DPM {

// flat arrays in the C ABI's layout: vertices padded to maxNV per cell with zeros (reference :117-140)
struct Packed2D {
  int maxNV = 0;
  std::vector<float> verts, forces, Ka, Kl, Kb, l0, a0, r0;
  std::vector<int32_t> NV;
};

struct DeviceHandle2D {
  dpm2d_t *h = nullptr;
  int ncells = 0, max_nv = 0;
  bool resident = false;  // the device holds the tissue's current state (StepResident)
  Packed2D staging;
  ~DeviceHandle2D() {
    if (h) dpm2d_destroy(h);
  }
};

static std::string err2d() {
  char b[1024];
  dpm_last_error(b, sizeof b);
  return std::string(b);
}

static void pack2d(const std::vector<Cell2D> &cells, int NCELLS, Packed2D &P) {
  int maxNV = cells.empty() ? 0 : (int)cells[0].NV;
  for (auto &c : cells) maxNV = std::max(maxNV, (int)c.NV);
  P.maxNV = maxNV;
  const size_t stride = (size_t)maxNV * 2;
  P.verts.assign(stride * NCELLS, 0.0f);
  P.forces.assign(stride * NCELLS, 0.0f);
  for (auto *x : {&P.Ka, &P.Kl, &P.Kb, &P.l0, &P.a0, &P.r0}) x->assign(NCELLS, 0.0f);
  P.NV.assign(NCELLS, 0);
  for (int ci = 0; ci < NCELLS; ci++) {
    const Cell2D &c = cells[ci];
    P.Ka[ci] = c.Ka; P.Kl[ci] = c.Kl; P.Kb[ci] = c.Kb; P.l0[ci] = c.l0; P.a0[ci] = c.a0; P.r0[ci] = c.r0; P.NV[ci] = (int32_t)c.NV;
    for (unsigned int vi = 0; vi < c.NV; vi++) {
      P.verts[ci * stride + 2 * vi] = c.Verticies[vi][0];
      P.verts[ci * stride + 2 * vi + 1] = c.Verticies[vi][1];
    }
  }
}

// real vertices only (reference :235-240)
static void unpack2d(std::vector<Cell2D> &cells, int NCELLS, const Packed2D &P) {
  const size_t stride = (size_t)P.maxNV * 2;
  for (int ci = 0; ci < NCELLS; ci++)
    for (unsigned int vi = 0; vi < cells[ci].NV; vi++) {
      cells[ci].Verticies[vi] = {P.verts[ci * stride + 2 * vi], P.verts[ci * stride + 2 * vi + 1]};
      cells[ci].Forces[vi] = {P.forces[ci * stride + 2 * vi], P.forces[ci * stride + 2 * vi + 1]};
    }
}

static void ensure_handle2d(DeviceHandle2D &dev, int NCELLS, int maxNV) {
  if (!dev.h || dev.ncells != NCELLS || dev.max_nv != maxNV) {
    if (dev.h) { dpm2d_destroy(dev.h); dev.h = nullptr; }
    if (dpm2d_create(&dev.h, 0, NCELLS, maxNV) != DPM_OK) throw std::runtime_error(err2d());
    dev.ncells = NCELLS;
    dev.max_nv = maxNV;
  }
}

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
  Packed2D P;
  pack2d(cells, NCELLS, P);
  maxNV = P.maxNV;
  if (!dev) dev = std::make_shared<DeviceHandle2D>();
  dev->resident = false;  // this call re-uploads; afterwards callers may edit cells, so nothing is assumed resident
  ensure_handle2d(*dev, NCELLS, maxNV);
  const int rc = dpm2d_euler_update(dev->h, P.verts.data(), P.forces.data(), P.NV.data(), P.Ka.data(), P.Kl.data(), P.Kb.data(),
                                    P.a0.data(), P.l0.data(), P.r0.data(), nsteps, dt, Kre, Kat, (int)PBC, L, nullptr);
  if (rc == DPM_ERR_INVALID_ARGUMENT) throw std::invalid_argument(err2d());
  if (rc != DPM_OK) throw std::runtime_error(err2d());
  unpack2d(cells, NCELLS, P);
}

// ---- device-resident stepping (extension, see Tissue.hpp) -------------------------------------------------------
void Tissue2D::InvalidateDevice() {
  if (dev) dev->resident = false;
}

void Tissue2D::StepResident(int nsteps, float dt) {
  if (!dev) dev = std::make_shared<DeviceHandle2D>();
  Packed2D &P = dev->staging;
  if (!dev->resident) {
    pack2d(cells, NCELLS, P);
    maxNV = P.maxNV;
    ensure_handle2d(*dev, NCELLS, maxNV);
    if (dpm2d_upload(dev->h, P.verts.data(), P.NV.data(), P.Ka.data(), P.Kl.data(), P.Kb.data(), P.a0.data(), P.l0.data(), P.r0.data()) != DPM_OK)
      throw std::runtime_error(err2d());
    dev->resident = true;
  }
  const int rc = dpm2d_step(dev->h, nsteps, dt, Kre, Kat, (int)PBC, L);  // asynchronous
  if (rc == DPM_ERR_INVALID_ARGUMENT) throw std::invalid_argument(err2d());
  if (rc != DPM_OK) throw std::runtime_error(err2d());
}

void Tissue2D::SyncCells() {
  if (!dev || !dev->h || !dev->resident) return;
  Packed2D &P = dev->staging;
  if (dpm2d_download(dev->h, P.verts.data(), P.forces.data()) != DPM_OK) {
    dev->resident = false;
    throw std::runtime_error(err2d());
  }
  unpack2d(cells, NCELLS, P);
}

}  // namespace DPM
