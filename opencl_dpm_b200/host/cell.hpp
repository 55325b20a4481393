// cell.hpp — DPM::Cell2D / DPM::Cell3D with the public surface of the reference's
// include/cell.hpp:11-58 (same member names, types and method signatures) so that
// existing C++ and Python callers compile and run unchanged.  No OpenCL headers.
#ifndef DPM_B200_CELL_HPP
#define DPM_B200_CELL_HPP
#include <array>
#include <vector>

namespace DPM {

struct Cell2D {
  unsigned int NV;
  float calA0;
  float a0;
  float l0;
  float r0;
  float Ka;
  float Kb;
  float Kl;
  float Ks;
  std::vector<std::array<float, 2>> Verticies;  // (sic) spelling kept for source compatibility
  std::vector<std::array<float, 2>> Forces;
  Cell2D(float x0, float y0, float calA, unsigned int NV, float r0);
  float GetArea();
  float GetPerim();
};

class Cell3D {
public:
  static const unsigned int NV = 162;  // icosphere, 2 subdivisions (reference include/cell.hpp:30-31): the DEFAULT mesh
  static const unsigned int NF = 320;
  // Extension: the reference hard-codes two subdivisions; BASELINE's 642-vertex configurations need
  // three.  A cell built with the 4-argument constructor carries its own mesh size; everything that loops over a cell
  // uses nverts()/nfaces() (== NV/NF for the reference mesh).  0 <= subdivisions <= 3 (12, 42, 162, 642 vertices).
  int subdivisions = 2;
  unsigned int nverts() const { return (unsigned int)Verts.size(); }
  unsigned int nfaces() const { return (unsigned int)Faces.size(); }
  float calA0;
  float r0;
  float v0;
  float sa0;
  float a0;
  float Kv;
  float Ka;
  float Ks;
  float Volume;
  float SurfaceArea;
  std::vector<std::array<float, 3>> Verts;
  std::vector<std::array<float, 3>> Forces;
  std::vector<std::array<unsigned int, 3>> Faces;

  float GetVolume();
  float GetSurfaceArea();
  std::array<std::array<float, NV>, 3> GetPositions();
  std::array<std::array<float, 162>, 3> GetVesselPositions(float L);
  std::array<std::array<float, NV>, 3> GetForces();
  std::array<std::array<int, 3>, NF> GetFaces();
  std::array<float, 3> GetCOM();
  Cell3D(std::array<float, 3> starting_point, float CalA0, float r0);
  Cell3D(std::array<float, 3> starting_point, float CalA0, float r0, int subdivisions);
  // size-agnostic forms of the four getters above (the fixed-size ones throw std::logic_error for a non-reference mesh);
  // the Python module binds GetPositions/GetVesselPositions/GetForces/GetFaces to these: same list-of-lists result
  std::vector<std::vector<float>> GetPositionsV();
  std::vector<std::vector<float>> GetVesselPositionsV(float L);
  std::vector<std::vector<float>> GetForcesV();
  std::vector<std::array<int, 3>> GetFacesV();
};

}  // namespace DPM
#endif
