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
  static const unsigned int NV = 162;  // icosphere, 2 subdivisions (reference include/cell.hpp:30-31)
  static const unsigned int NF = 320;
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
};

}  // namespace DPM
#endif
