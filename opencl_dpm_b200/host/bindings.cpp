// bindings.cpp — pybind11 module `clDPM`: the reference's Python surface
// (src/DPMWrapper.cpp:11-17, src/CellWrapper.cpp:7-29, src/TissueWrapper.cpp:7-29),
// bound to the B200 host classes.  Attribute names, read/write-ness and method names
// are identical; T.Cells stays copy-on-access through the STL casters.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "Tissue.hpp"
#include "disperse.hpp"

namespace py = pybind11;
using namespace DPM;

// numpy view [NCELLS][NV][4] of a shared-owned packed array: no copy, the capsule keeps the array alive
static py::object packed_view(const std::shared_ptr<std::vector<float>> &buf, int nc, int nv) {
  if (!buf) return py::none();
  auto *keep = new std::shared_ptr<std::vector<float>>(buf);
  py::capsule owner(keep, [](void *p) { delete static_cast<std::shared_ptr<std::vector<float>> *>(p); });
  return py::array_t<float>({(py::ssize_t)nc, (py::ssize_t)nv, (py::ssize_t)4},
                            {(py::ssize_t)(sizeof(float) * 4 * nv), (py::ssize_t)(sizeof(float) * 4), (py::ssize_t)sizeof(float)}, buf->data(), owner);
}

PYBIND11_MODULE(clDPM, m) {
  m.doc() = "Deformable Particle Model — B200-native (CUDA sm_100a) drop-in for OpenCL_DPM's clDPM";

  // test hook (not part of the reference surface): the centre relaxation behind Disperse()/Disperse2D(), in its binned
  // form (what the classes use) or in the reference's all-pairs form; returns (X, Y, hit_iteration_cap)
  m.def("_relax_centres", [](const std::vector<float> &radius, float L, bool allpairs) {
    std::vector<float> X, Y;
    const bool cap = allpairs ? DPM::detail::relax_centres_allpairs(radius, L, X, Y) : DPM::detail::relax_centres(radius, L, X, Y);
    return py::make_tuple(X, Y, cap);
  });

  py::class_<Cell2D>(m, "Cell2D")
      .def(py::init<float, float, float, unsigned int, float>())
      .def_readwrite("Ka", &Cell2D::Ka)
      .def_readwrite("Kl", &Cell2D::Kl)
      .def_readwrite("Kb", &Cell2D::Kb)
      .def_readwrite("Verts", &Cell2D::Verticies)
      .def_readwrite("Forces", &Cell2D::Forces);

  py::class_<Tissue2D>(m, "Tissue2D")
      .def(py::init<std::vector<Cell2D>, float>())
      .def_readwrite("Cells", &Tissue2D::cells)
      .def_readonly("NCELLS", &Tissue2D::NCELLS)
      .def_readonly("L", &Tissue2D::L)
      .def_readonly("PBC", &Tissue2D::PBC)
      .def_readwrite("Kre", &Tissue2D::Kre)
      .def_readwrite("Kat", &Tissue2D::Kat)
      .def("CLEulerUpdate", &Tissue2D::CLEulerUpdate)
      .def("AppendFrame", &Tissue2D::AppendFrame)
      .def("StepResident", &Tissue2D::StepResident)
      .def("SyncCells", &Tissue2D::SyncCells)
      .def("InvalidateDevice", &Tissue2D::InvalidateDevice)
      .def("Disperse", &Tissue2D::Disperse);

  py::class_<Cell3D>(m, "Cell3D")
      .def(py::init<std::array<float, 3>, float, float>())
      // extension: Cell3D(start, calA, r0, subdivisions) — 3 gives the 642-vertex mesh of BASELINE configs D/E
      .def(py::init<std::array<float, 3>, float, float, int>())
      .def_readonly("subdivisions", &Cell3D::subdivisions)
      .def_property_readonly("NV", &Cell3D::nverts)
      .def_property_readonly("NF", &Cell3D::nfaces)
      .def_readwrite("Kv", &Cell3D::Kv)
      .def_readwrite("Ka", &Cell3D::Ka)
      .def_readwrite("Ks", &Cell3D::Ks)
      .def_readwrite("Verts", &Cell3D::Verts)
      .def("GetVolume", &Cell3D::GetVolume)
      .def("GetPositions", &Cell3D::GetPositionsV)
      .def("GetVesselPositions", &Cell3D::GetVesselPositionsV)
      .def("GetFaces", &Cell3D::GetFacesV)
      .def("GetForces", &Cell3D::GetForcesV);

  py::class_<Tissue3D>(m, "Tissue3D")
      .def(py::init<std::vector<Cell3D>, float>())
      .def_readwrite("Kre", &Tissue3D::Kre)
      .def_readwrite("Kat", &Tissue3D::Kat)
      // extension: "General" (default, reference behaviour: Kat is inert in 3D) or "AllVertAttraction"
      .def_readwrite("attractionMethod", &Tissue3D::attractionMethod)
      .def_readwrite("Cells", &Tissue3D::Cells)
      .def_readonly("NCELLS", &Tissue3D::NCELLS)
      .def_readonly("L", &Tissue3D::L)
      .def_readonly("PBC", &Tissue3D::PBC)
      .def("CLEulerUpdate", &Tissue3D::CLEulerUpdate)
      .def("AppendFrame", &Tissue3D::AppendFrame)
      // extensions: device-resident stepping (no per-call upload/download), see Tissue.hpp
      .def("StepResident", &Tissue3D::StepResident)
      .def("SyncCells", &Tissue3D::SyncCells)
      .def("InvalidateDevice", &Tissue3D::InvalidateDevice)
      // extension: zero-copy numpy views [NCELLS][NV][4] (x, y, z, pad) of the positions / last-step forces as the last
      // CLEulerUpdate or SyncCells left them (None before the first call); the next call rewrites them in place
      .def("PositionsView", [](const Tissue3D &T) { int nc = 0, nv = 0; auto b = T.PackedPositions(&nc, &nv); return packed_view(b, nc, nv); })
      .def("ForcesView", [](const Tissue3D &T) { int nc = 0, nv = 0; auto b = T.PackedForces(&nc, &nv); return packed_view(b, nc, nv); })
      .def("Disperse2D", &Tissue3D::Disperse2D);
}
