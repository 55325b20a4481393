// trajectory.cpp — flat binary trajectory writer for Tissue2D / Tissue3D.
//
// The reference has no on-disk format: its demos render every frame to a PNG through matplotlib (plot.py,
// test3D.py:21-27, test2D.py:23-31).  AppendFrame(path) appends the current vertex positions to a self-describing
// little-endian file instead (reader: opencl_dpm_b200/traj.py):
//
//   header   char[8] "DPMTRAJ1" | int32 dim | int32 ncells | int32 nv (3D: vertices per cell; 2D: largest NV) |
//            int32 nf | float32 L | int32 PBC | int32[2] reserved                                   (40 bytes)
//   3D       int32 faces[nf][3]                      (the shared topology, Cells[0].Faces)
//   2D       int32 NV[ncells]
//   frames   float32 positions[ncells][nv][dim], back to back (2D rows are padded to nv with zeros)
//
// A file is created on the first call; later calls check that the header still describes the tissue.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "Tissue.hpp"

namespace DPM {
namespace {

struct Header {
  char magic[8];
  int32_t dim, ncells, nv, nf;
  float L;
  int32_t pbc, reserved[2];
};
static_assert(sizeof(Header) == 40, "trajectory header layout");

void append(const std::string &path, const Header &h, const std::vector<int32_t> &table, const std::vector<float> &frame) {
  FILE *f = std::fopen(path.c_str(), "rb");
  const bool fresh = (f == nullptr);
  if (f) {
    Header old;
    const bool ok = std::fread(&old, sizeof old, 1, f) == 1 && std::memcmp(old.magic, h.magic, 8) == 0 && old.dim == h.dim &&
                    old.ncells == h.ncells && old.nv == h.nv && old.nf == h.nf;
    std::fclose(f);
    if (!ok) throw std::runtime_error("AppendFrame: " + path + " exists and does not describe this tissue");
  }
  f = std::fopen(path.c_str(), fresh ? "wb" : "ab");
  if (!f) throw std::runtime_error("AppendFrame: cannot open " + path);
  bool ok = true;
  if (fresh) {
    ok = std::fwrite(&h, sizeof h, 1, f) == 1;
    if (ok && !table.empty()) ok = std::fwrite(table.data(), sizeof(int32_t), table.size(), f) == table.size();
  }
  if (ok) ok = std::fwrite(frame.data(), sizeof(float), frame.size(), f) == frame.size();
  std::fclose(f);
  if (!ok) throw std::runtime_error("AppendFrame: short write to " + path);
}

}  // namespace

void Tissue3D::AppendFrame(const std::string &path) {
  if (Cells.empty()) throw std::runtime_error("AppendFrame: empty tissue");
  Header h{};
  std::memcpy(h.magic, "DPMTRAJ1", 8);
  h.dim = 3; h.ncells = (int32_t)Cells.size(); h.nv = (int32_t)Cells[0].nverts(); h.nf = (int32_t)Cells[0].nfaces();
  h.L = L; h.pbc = PBC;
  std::vector<int32_t> faces;
  for (const auto &t : Cells[0].Faces) for (int k = 0; k < 3; k++) faces.push_back((int32_t)t[k]);
  std::vector<float> frame;
  frame.reserve((size_t)h.ncells * h.nv * 3);
  for (const auto &c : Cells) {
    if ((int32_t)c.nverts() != h.nv) throw std::runtime_error("AppendFrame: all cells of a tissue must share one mesh");
    for (const auto &v : c.Verts) { frame.push_back(v[0]); frame.push_back(v[1]); frame.push_back(v[2]); }
  }
  append(path, h, faces, frame);
}

void Tissue2D::AppendFrame(const std::string &path) {
  if (cells.empty()) throw std::runtime_error("AppendFrame: empty tissue");
  Header h{};
  std::memcpy(h.magic, "DPMTRAJ1", 8);
  int32_t mx = 0;
  std::vector<int32_t> nvs;
  for (const auto &c : cells) { nvs.push_back((int32_t)c.NV); mx = c.NV > (unsigned)mx ? (int32_t)c.NV : mx; }
  h.dim = 2; h.ncells = (int32_t)cells.size(); h.nv = mx; h.nf = 0; h.L = L; h.pbc = PBC ? 1 : 0;
  std::vector<float> frame((size_t)h.ncells * mx * 2, 0.0f);
  for (size_t ci = 0; ci < cells.size(); ci++)
    for (unsigned int i = 0; i < cells[ci].NV; i++) {
      frame[(ci * mx + i) * 2] = cells[ci].Verticies[i][0];
      frame[(ci * mx + i) * 2 + 1] = cells[ci].Verticies[i][1];
    }
  append(path, h, nvs, frame);
}

}  // namespace DPM
