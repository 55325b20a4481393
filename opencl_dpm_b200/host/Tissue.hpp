// Tissue.hpp — DPM::Tissue3D / DPM::Tissue2D with the public surface of the reference's
// include/Tissue.hpp:13-55.  The four cl:: members of each class (platform, device,
// program, context) are replaced by one opaque device handle over the C ABI in
// include/dpm_b200.h; everything a caller can name is unchanged.
#ifndef DPM_B200_TISSUE_HPP
#define DPM_B200_TISSUE_HPP
#include <memory>
#include <string>
#include <vector>

#include "cell.hpp"

namespace DPM {

struct DeviceHandle3D;  // owns a dpm3d_t*
struct DeviceHandle2D;  // owns a dpm2d_t*

class Tissue3D {
public:
  int NCELLS;
  int PBC;  // (boolean) periodic boundary conditions
  float L;
  float Kre;
  float Kat;
  std::string attractionMethod;
  std::vector<DPM::Cell3D> Cells;

  Tissue3D(std::vector<DPM::Cell3D> cells, float phi0);

  void CLEulerUpdate(int nsteps, float dt);
  void Disperse2D();
  // extension: append the current vertex positions to a flat binary trajectory (host/trajectory.cpp)
  void AppendFrame(const std::string &path);
  // extension: device-resident stepping.  CLEulerUpdate must re-upload Cells and read everything back
  // on every call because callers may edit Cells between calls (SURVEY §8b "Ownership"); a loop of many short calls pays
  // pack + H2D + D2H + unpack each time.  StepResident uploads Cells only when the device copy is stale (first use, after
  // CLEulerUpdate, after InvalidateDevice()) and then only enqueues timesteps; SyncCells() brings positions, last-step
  // forces, Volume and SurfaceArea back into Cells (same checks and warnings as CLEulerUpdate).  Kre/Kat/PBC/L are read on
  // every call; per-cell stiffnesses at upload time.  Same validation and exception types as CLEulerUpdate.
  void StepResident(int nsteps, float dt);
  void SyncCells();
  void InvalidateDevice();
  // extension: zero-copy access to the packed host arrays of the last CLEulerUpdate / SyncCells,
  // [NCELLS][NV][4] floats; the Python module exposes them as numpy views (PositionsView / ForcesView) instead of the
  // list-of-lists copies of Cell3D::GetPositions / GetForces (src/CellWrapper.cpp:7-19)
  std::shared_ptr<std::vector<float>> PackedPositions(int *ncells, int *nv) const;
  std::shared_ptr<std::vector<float>> PackedForces(int *ncells, int *nv) const;

private:
  std::shared_ptr<DeviceHandle3D> dev;  // created on first use; copies of a Tissue share it
};

class Tissue2D {
public:
  std::vector<Cell2D> cells;
  int NCELLS;
  bool PBC;
  float Kre;
  float Kat;
  float phi0;
  float L;
  Tissue2D(std::vector<DPM::Cell2D> cells, float phi0);
  void Disperse();
  void CLEulerUpdate(int nsteps, float dt);
  void AppendFrame(const std::string &path);  // extension, see Tissue3D::AppendFrame
  void StepResident(int nsteps, float dt);     // extensions, see Tissue3D
  void SyncCells();
  void InvalidateDevice();

private:
  int maxNV;
  std::shared_ptr<DeviceHandle2D> dev;
};

}  // namespace DPM
#endif
