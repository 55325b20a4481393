// Tissue3D.cpp — host side of the 3D hot path over the C ABI.
//
// Mirrors DPM::Tissue3D of the reference (src/Tissue3D.cpp): same constructor arithmetic
// (:16-29), same Disperse2D behaviour (:40-116), and a CLEulerUpdate with the same
// validation, exceptions, stdout/stderr messages and post-conditions (:118-522) — but the
// OpenCL program build, the ten cl::Buffers, six cl::Kernels and the per-step enqueue loop
// (:199-470) are ONE call into dpm3d_euler_update (include/dpm_b200.h).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <thread>

#include "Tissue.hpp"
#include "disperse.hpp"
#include "dpm_b200.h"

namespace DPM {

// flat arrays in the C ABI's layout (float4-strided vertices, per-cell scalars), as the reference packs them (:137-197)
struct Packed3D {
  int NV = 0, NF = 0, NC = 0;
  std::vector<uint32_t> faces;
  // the two large arrays are shared-owned: zero-copy numpy views (Tissue3D::PackedPositions / PackedForces) keep the array
  // they look at alive even if a later call re-sizes the staging
  std::shared_ptr<std::vector<float>> vbuf = std::make_shared<std::vector<float>>(), fbuf = std::make_shared<std::vector<float>>();
  std::vector<float> &verts() const { return *vbuf; }
  std::vector<float> &forces() const { return *fbuf; }
  // a NEW pair of arrays (zeroed) when the tissue's size changed; views of the old pair stay valid
  void resize(size_t n) {
    if (vbuf->size() == n) return;
    vbuf = std::make_shared<std::vector<float>>(n, 0.0f);
    fbuf = std::make_shared<std::vector<float>>(n, 0.0f);
  }
  std::vector<float> Kv, Ka, Ks, v0, a0, l0;
};

struct DeviceHandle3D {
  dpm3d_t *h = nullptr;
  int ncells = 0;
  std::vector<uint32_t> faces;
  bool resident = false;  // the device holds the tissue's current state (StepResident)
  Packed3D staging;       // host staging, kept between calls (no 84 MB allocation + first touch per call at BASELINE sizes)
  const float *pinned_v = nullptr, *pinned_f = nullptr;  // page-locked state of the two large staging arrays
  // (re-)page-lock the large staging arrays after pack3d may have reallocated them; small tissues are not worth the system call
  void pin() {
    if (staging.verts().size() < (1u << 18)) return;
    if (staging.verts().data() != pinned_v) {
      if (pinned_v) dpm_unpin_host_buffer(const_cast<float *>(pinned_v));
      pinned_v = dpm_pin_host_buffer(staging.verts().data(), staging.verts().size() * sizeof(float)) == DPM_OK ? staging.verts().data() : nullptr;
    }
    if (staging.forces().data() != pinned_f) {
      if (pinned_f) dpm_unpin_host_buffer(const_cast<float *>(pinned_f));
      pinned_f = dpm_pin_host_buffer(staging.forces().data(), staging.forces().size() * sizeof(float)) == DPM_OK ? staging.forces().data() : nullptr;
    }
  }
  ~DeviceHandle3D() {
    if (pinned_v) dpm_unpin_host_buffer(const_cast<float *>(pinned_v));
    if (pinned_f) dpm_unpin_host_buffer(const_cast<float *>(pinned_f));
    if (h) dpm3d_destroy(h);
  }
};

static std::string last_error() {
  char buf[1024];
  dpm_last_error(buf, sizeof buf);
  return std::string(buf);
}

Tissue3D::Tissue3D(std::vector<DPM::Cell3D> cells, float phi0) {
  Cells = cells;
  NCELLS = Cells.size();
  if (NCELLS > 100) {
    // kept for output fidelity (reference :19-23); the all-pairs limit behind it is gone
    std::cerr << "Warning: Large number of cells, too many cells may crash program!" << std::endl;
  }
  float volume = 0.0f;
  for (int ci = 0; ci < NCELLS; ci++) volume += Cells[ci].v0;
  L = cbrt(volume) / phi0;
  PBC = true;
  Kre = 0.0f;  // uninitialised in the reference; callers always set it
  Kat = 0.0f;
  attractionMethod.assign("General");
}

void Tissue3D::Disperse2D() {
  std::vector<float> radius(NCELLS), X, Y;
  for (int i = 0; i < NCELLS; i++) radius[i] = Cells[i].r0 * 2;
  if (detail::relax_centres(radius, L, X, Y))
    std::cerr << "Warning: Max timesteps for dispersion reached" << std::endl;
  // reference :106-115 — the centre is SUBTRACTED (sic), z is left untouched
  for (int i = 0; i < NCELLS; i++) {
    std::array<float, 3> com = Cells[i].GetCOM();
    for (unsigned int j = 0; j < Cells[i].nverts(); j++) {
      Cells[i].Verts[j][0] -= com[0];
      Cells[i].Verts[j][1] -= com[1];
      Cells[i].Verts[j][0] -= X[i];
      Cells[i].Verts[j][1] -= Y[i];
    }
  }
}

// Packing and unpacking walk every vertex of every cell on the host (AoS <-> the C ABI's flat arrays); for BASELINE-size
// tissues (millions of vertices) that costs more than the device call itself.  Cells are independent, so ranges of cells
// go to a few threads.  The threads only run the silent happy path: any condition that would make the reference print or
// throw (:149-190, :477-521) sets `bad`, and the caller then repeats the pass serially with the original, ordered
// messages and exceptions.
template <typename Fn>
static void parallel_cells(int ncells, size_t work_per_cell, Fn &&fn) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int nt = (int)std::min<size_t>(std::min<unsigned>(hw ? hw : 1u, 16u), (size_t)ncells * work_per_cell / 100000);
  if (nt < 2) { fn(0, ncells); return; }
  std::vector<std::thread> th;
  const int chunk = (ncells + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    const int c0 = t * chunk, c1 = std::min(ncells, c0 + chunk);
    if (c0 < c1) th.emplace_back([&fn, c0, c1] { fn(c0, c1); });
  }
  for (auto &x : th) x.join();
}

// validation of the step arguments: same conditions, messages and exception types as reference :123-135
static void validate_step(int nsteps, float dt, int NCELLS) {
  if (nsteps <= 0) {
    std::cerr << "[ERROR] Invalid nsteps: " << nsteps << std::endl;
    throw std::invalid_argument("nsteps must be positive");
  }
  if (dt <= 0.0f || dt > 0.1f) {
    std::cerr << "[ERROR] Invalid dt: " << dt << " (should be in range (0, 0.1])" << std::endl;
    throw std::invalid_argument("dt must be positive and reasonable");
  }
  if (NCELLS <= 0) {
    std::cerr << "[ERROR] Invalid NCELLS: " << NCELLS << std::endl;
    throw std::invalid_argument("NCELLS must be positive");
  }

}

// topology from Cells[0] only (:144-155), per-cell scalars and float4 vertices (:157-197), with the reference's checks
static void pack3d(const std::vector<Cell3D> &Cells, int NCELLS, Packed3D &P) {
  // the reference hard-codes Cell3D::NF / Cell3D::NV (:119-120); cells built with a subdivision level carry their own
  const int NF = P.NF = !Cells.empty() ? (int)Cells[0].nfaces() : (int)Cell3D::NF;
  const int NV = P.NV = !Cells.empty() ? (int)Cells[0].nverts() : (int)Cell3D::NV;
  std::vector<uint32_t> &faces = P.faces;
  std::vector<float> &Kv = P.Kv, &Ka = P.Ka, &Ks = P.Ks, &v0 = P.v0, &a0 = P.a0, &l0 = P.l0;
  faces.assign(3 * (size_t)NF, 0);
  for (int fi = 0; fi < NF; fi++) {
    const auto &f = Cells[0].Faces[fi];
    if (f[0] >= (unsigned)NV || f[1] >= (unsigned)NV || f[2] >= (unsigned)NV) {
      std::cerr << "[ERROR] Invalid face index at face " << fi << ": (" << f[0] << "," << f[1] << "," << f[2] << ")"
                << std::endl;
      throw std::runtime_error("Invalid face indices");
    }
    faces[3 * fi] = f[0]; faces[3 * fi + 1] = f[1]; faces[3 * fi + 2] = f[2];
  }
  // the staging arrays persist between calls: only (re)sized when the tissue changed — no re-zeroing of 2 x 42 MB per call at
  // BASELINE sizes; every xyz is written below, the pad lane stays 0 from the first sizing, forces are overwritten by the call
  P.resize((size_t)NCELLS * NV * 4);
  P.NC = NCELLS;
  std::vector<float> &verts = P.verts();
  for (auto *x : {&Kv, &Ka, &Ks, &v0, &a0, &l0}) x->assign(NCELLS, 0.0f);
  {  // threaded happy path; anything the checks below would report makes the serial pass run instead
    std::atomic<bool> bad{false};
    parallel_cells(NCELLS, (size_t)NV, [&](int c0, int c1) {
      for (int ci = c0; ci < c1 && !bad.load(std::memory_order_relaxed); ci++) {
        const Cell3D &c = Cells[ci];
        bool ok = (int)c.nverts() == NV && (int)c.nfaces() == NF && (int)c.Forces.size() == NV && c.Kv > 0 && c.Ka > 0 && c.Ks > 0 &&
                  c.v0 > 0 && c.a0 > 0;
        if (ok) {
          Kv[ci] = c.Kv; Ka[ci] = c.Ka; Ks[ci] = c.Ks; v0[ci] = c.v0; a0[ci] = c.a0;
          l0[ci] = sqrt(4.0f * c.a0) / sqrt(3.0f);
          float *dst = &verts[(size_t)ci * NV * 4];
          for (int vi = 0; vi < NV; vi++, dst += 4) {
            const auto &q = c.Verts[vi];
            ok = ok && std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]);
            dst[0] = q[0]; dst[1] = q[1]; dst[2] = q[2];
          }
        }
        if (!ok) bad.store(true, std::memory_order_relaxed);
      }
    });
    if (!bad.load()) return;
  }
  for (int ci = 0; ci < NCELLS; ci++) {
    const Cell3D &c = Cells[ci];
    if ((int)c.nverts() != NV || (int)c.nfaces() != NF || (int)c.Forces.size() != NV) {
      // one topology for the whole tissue, as in the reference (Cells[0].Faces is the only face list it uploads)
      std::cerr << "[ERROR] Cell " << ci << " has " << c.nverts() << " vertices, cell 0 has " << NV << std::endl;
      throw std::runtime_error("All cells of a tissue must share one mesh");
    }
    if (c.Kv <= 0 || c.Ka <= 0 || c.Ks <= 0) {
      std::cerr << "[ERROR] Invalid spring constants for cell " << ci << ": Kv=" << c.Kv << ", Ka=" << c.Ka
                << ", Ks=" << c.Ks << std::endl;
      throw std::runtime_error("Invalid spring constants");
    }
    if (c.v0 <= 0 || c.a0 <= 0) {
      std::cerr << "[ERROR] Invalid reference values for cell " << ci << ": v0=" << c.v0 << ", a0=" << c.a0 << std::endl;
      throw std::runtime_error("Invalid reference values");
    }
    Kv[ci] = c.Kv; Ka[ci] = c.Ka; Ks[ci] = c.Ks; v0[ci] = c.v0; a0[ci] = c.a0;
    l0[ci] = sqrt(4.0f * c.a0) / sqrt(3.0f);  // rest edge length of an equilateral triangle of area a0 (:177)
    for (int vi = 0; vi < NV; vi++) {
      float *dst = &verts[((size_t)ci * NV + vi) * 4];
      for (int d = 0; d < 3; d++) {
        if (!std::isfinite(c.Verts[vi][d])) {
          std::cerr << "[ERROR] Non-finite vertex coordinate at cell " << ci << ", vertex " << vi << ", coord " << d
                    << ": " << c.Verts[vi][d] << std::endl;
          throw std::runtime_error("Non-finite vertex coordinates");
        }
        dst[d] = c.Verts[vi][d];
      }
    }
  }

}

// flat -> Cells with the reference's post-conditions and checks (:477-521)
static void unpack3d(std::vector<Cell3D> &Cells, int NCELLS, const Packed3D &P) {
  const int NV = P.NV;
  const std::vector<float> &verts = P.verts(), &forces = P.forces();
  {  // threaded happy path; anything the checks below would report makes the serial pass run instead
    std::atomic<bool> bad{false};
    parallel_cells(NCELLS, (size_t)NV * 3, [&](int c0, int c1) {
      for (int ci = c0; ci < c1 && !bad.load(std::memory_order_relaxed); ci++) {
        Cell3D &c = Cells[ci];
        bool ok = true;
        const float *pv = &verts[(size_t)ci * NV * 4], *pf = &forces[(size_t)ci * NV * 4];
        for (int vi = 0; vi < NV; vi++, pv += 4, pf += 4) {
          ok = ok && std::isfinite(pv[0]) && std::isfinite(pv[1]) && std::isfinite(pv[2]) && std::isfinite(pf[0]) && std::isfinite(pf[1]) &&
               std::isfinite(pf[2]);
          c.Verts[vi] = {pv[0], pv[1], pv[2]};
          c.Forces[vi] = {pf[0], pf[1], pf[2]};
        }
        c.Volume = c.GetVolume();
        c.SurfaceArea = c.GetSurfaceArea();
        ok = ok && std::isfinite(c.Volume) && c.Volume > 0 && std::isfinite(c.SurfaceArea) && c.SurfaceArea > 0;
        if (!ok) bad.store(true, std::memory_order_relaxed);
      }
    });
    if (!bad.load()) return;
  }
  for (int ci = 0; ci < NCELLS; ci++) {
    for (int vi = 0; vi < NV; vi++) {
      const float *pv = &verts[((size_t)ci * NV + vi) * 4], *pf = &forces[((size_t)ci * NV + vi) * 4];
      for (int d = 0; d < 3; d++) {
        if (!std::isfinite(pv[d])) {
          std::cerr << "[ERROR] Non-finite result vertex at cell " << ci << ", vertex " << vi << ", coord " << d << ": "
                    << pv[d] << std::endl;
          throw std::runtime_error("Non-finite simulation results");
        }
        if (!std::isfinite(pf[d]))
          std::cerr << "[WARNING] Non-finite force at cell " << ci << ", vertex " << vi << ", coord " << d << ": " << pf[d]
                    << std::endl;
      }
      Cells[ci].Verts[vi] = {pv[0], pv[1], pv[2]};
      Cells[ci].Forces[vi] = {pf[0], pf[1], pf[2]};
    }
    Cells[ci].Volume = Cells[ci].GetVolume();
    Cells[ci].SurfaceArea = Cells[ci].GetSurfaceArea();
    if (!std::isfinite(Cells[ci].Volume) || Cells[ci].Volume <= 0)
      std::cerr << "[WARNING] Invalid volume for cell " << ci << ": " << Cells[ci].Volume << std::endl;
    if (!std::isfinite(Cells[ci].SurfaceArea) || Cells[ci].SurfaceArea <= 0)
      std::cerr << "[WARNING] Invalid surface area for cell " << ci << ": " << Cells[ci].SurfaceArea << std::endl;
  }
}

void Tissue3D::CLEulerUpdate(int nsteps, float dt) {
  validate_step(nsteps, dt, NCELLS);
  if (!dev) dev = std::make_shared<DeviceHandle3D>();
  Packed3D &P = dev->staging;  // persistent between calls; the arrays are rewritten from Cells by every call
  pack3d(Cells, NCELLS, P);
  dev->pin();
  const int NF = P.NF, NV = P.NV;
  std::vector<uint32_t> &faces = P.faces;
  std::vector<float> &verts = P.verts(), &forces = P.forces(), &Kv = P.Kv, &Ka = P.Ka, &Ks = P.Ks, &v0 = P.v0, &a0 = P.a0, &l0 = P.l0;

  // ---- device: one handle per tissue, re-created only if the size or topology changed
  try {
    if (!dev) dev = std::make_shared<DeviceHandle3D>();
    dev->resident = false;  // this call re-uploads; afterwards callers may edit Cells, so nothing is assumed resident
    if (!dev->h || dev->ncells != NCELLS || dev->faces != faces) {
      if (dev->h) { dpm3d_destroy(dev->h); dev->h = nullptr; }
      if (dpm3d_create(&dev->h, 0, NCELLS, NV, NF, faces.data()) != DPM_OK) throw std::runtime_error(last_error());
      dev->ncells = NCELLS;
      dev->faces = faces;
    }
    // AllVertAttraction (shaders/Cell3D_Kernel.cl:313-364) is compiled but never enqueued by the reference, whatever Kat
    // is, and `attractionMethod` ("General", src/Tissue3D.cpp:30) is never read.  Default behaviour is
    // therefore unchanged; setting attractionMethod = "AllVertAttraction" opts in to that kernel with strength Kat.
    const unsigned mask = DPM3D_ALL | (attractionMethod == "AllVertAttraction" ? DPM3D_ATTRACT : 0u);
    if (dpm3d_set_force_mask(dev->h, mask) != DPM_OK) throw std::runtime_error(last_error());
    float loop_ms = 0.0f;
    const int rc = dpm3d_euler_update(dev->h, verts.data(), forces.data(), Kv.data(), Ka.data(), Ks.data(), v0.data(),
                                      a0.data(), l0.data(), nsteps, dt, Kre, Kat, PBC, L, &loop_ms);
    if (rc == DPM_ERR_INVALID_ARGUMENT) throw std::invalid_argument(last_error());
    if (rc != DPM_OK) throw std::runtime_error(last_error());
    // same two lines the reference prints after its step loop (:454-461); the time is the
    // CUDA-event time of the step loop, rounded to whole milliseconds as the reference does
    const long long ms = (long long)loop_ms;
    std::cout << nsteps << " timesteps completed in " << ms << " ms" << std::endl;
    std::cout << "Average time per step: " << (ms / static_cast<double>(nsteps)) << " ms" << std::endl;
  } catch (const std::exception &e) {
    std::cerr << "[ERROR] Exception caught: " << e.what() << std::endl;
    throw;
  }

  unpack3d(Cells, NCELLS, P);
}

// ---- device-resident stepping (extension, see Tissue.hpp) -------------------------------------------------------
void Tissue3D::InvalidateDevice() {
  if (dev) dev->resident = false;
}

void Tissue3D::StepResident(int nsteps, float dt) {
  validate_step(nsteps, dt, NCELLS);
  try {
    if (!dev) dev = std::make_shared<DeviceHandle3D>();
    Packed3D &P = dev->staging;
    if (!dev->resident) {
      pack3d(Cells, NCELLS, P);
      if (!dev->h || dev->ncells != NCELLS || dev->faces != P.faces) {
        if (dev->h) { dpm3d_destroy(dev->h); dev->h = nullptr; }
        if (dpm3d_create(&dev->h, 0, NCELLS, P.NV, P.NF, P.faces.data()) != DPM_OK) throw std::runtime_error(last_error());
        dev->ncells = NCELLS;
        dev->faces = P.faces;
      }
      if (dpm3d_upload(dev->h, P.verts().data(), P.Kv.data(), P.Ka.data(), P.Ks.data(), P.v0.data(), P.a0.data(), P.l0.data()) != DPM_OK)
        throw std::runtime_error(last_error());
      dev->resident = true;
    }
    const unsigned mask = DPM3D_ALL | (attractionMethod == "AllVertAttraction" ? DPM3D_ATTRACT : 0u);
    if (dpm3d_set_force_mask(dev->h, mask) != DPM_OK) throw std::runtime_error(last_error());
    const int rc = dpm3d_step(dev->h, nsteps, dt, Kre, Kat, PBC, L);  // asynchronous: errors of the run surface in SyncCells
    if (rc == DPM_ERR_INVALID_ARGUMENT) throw std::invalid_argument(last_error());
    if (rc != DPM_OK) throw std::runtime_error(last_error());
  } catch (const std::exception &e) {
    std::cerr << "[ERROR] Exception caught: " << e.what() << std::endl;
    throw;
  }
}

void Tissue3D::SyncCells() {
  if (!dev || !dev->h || !dev->resident) return;  // nothing newer on the device than Cells
  Packed3D &P = dev->staging;
  if (dpm3d_download(dev->h, P.verts().data(), P.forces().data()) != DPM_OK) {
    dev->resident = false;
    std::cerr << "[ERROR] Exception caught: " << last_error() << std::endl;
    throw std::runtime_error(last_error());
  }
  unpack3d(Cells, NCELLS, P);
}

// ---- zero-copy views of the packed host arrays (extension) ------------------------------------------
// The arrays of the last CLEulerUpdate / SyncCells: [NCELLS][NV][4] floats (x, y, z, pad), positions and last-step forces.
// The returned owner keeps the array alive; the next call on the same tissue rewrites it in place (a live view).
std::shared_ptr<std::vector<float>> Tissue3D::PackedPositions(int *ncells, int *nv) const {
  if (!dev || dev->staging.vbuf->empty()) return nullptr;
  if (ncells) *ncells = dev->staging.NC;
  if (nv) *nv = dev->staging.NV;
  return dev->staging.vbuf;
}
std::shared_ptr<std::vector<float>> Tissue3D::PackedForces(int *ncells, int *nv) const {
  if (!dev || dev->staging.fbuf->empty()) return nullptr;
  if (ncells) *ncells = dev->staging.NC;
  if (nv) *nv = dev->staging.NV;
  return dev->staging.fbuf;
}

}  // namespace DPM
