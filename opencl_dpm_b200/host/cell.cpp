// cell.cpp — construction and measurement of single cells.
// Behavioural restatement of the reference's src/cell.cpp (same initial conditions:
// vertex numbering, the not-quite-unit icosphere of SURVEY F13, v0/sa0/a0, calA0/a0/l0);
// the icosphere itself comes from the C ABI helper dpm_icosphere so that flat-array users
// and the C++ classes share one implementation.
#include "cell.hpp"

#include <cmath>
#include <stdexcept>
#include <string>

#include "dpm_b200.h"

namespace DPM {

// ---- 2D ---------------------------------------------------------------------------
// reference src/cell.cpp:12-33: regular NV-gon of circumradius r0 starting at angle
// 2*pi*1/NV; calA0 is the requested shape parameter rescaled by the polygon/circle
// ratio; a0 is the polygon's own area, so a fresh cell has zero area strain.
Cell2D::Cell2D(float x0, float y0, float calA, unsigned int nverts, float radius) {
  NV = nverts;
  r0 = radius;
  Ka = Kb = Kl = Ks = 0.0f;  // the reference leaves these uninitialised; callers set them
  calA0 = calA * (NV * tan(M_PI / NV) / M_PI);
  Verticies.assign(NV, {0.0f, 0.0f});
  Forces.assign(NV, {0.0f, 0.0f});
  for (unsigned int i = 0; i < NV; i++) {
    const double ang = 2.0 * M_PI * (i + 1.0) / (float)NV;
    Verticies[i][0] = r0 * cos(ang) + x0;
    Verticies[i][1] = r0 * sin(ang) + y0;
  }
  a0 = GetArea();
  l0 = 2.0 * sqrt(M_PI * calA0 * a0) / (float)NV;
}

// shoelace area, accumulated in float with the 0.5 factor applied in double (src/cell.cpp:35-46)
float Cell2D::GetArea() {
  float area = 0.0f;
  for (unsigned int i = 0, j = NV - 1; i < NV; j = i++)
    area += 0.5 * ((Verticies[j][0] + Verticies[i][0]) * (Verticies[j][1] - Verticies[i][1]));
  return area < 0.0f ? -area : area;
}

// Closed-polygon perimeter.  The reference's closing edge reads Verticies[NV] (one past the
// end, src/cell.cpp:56-57 — undefined behaviour); the intended edge NV-1 -> 0 is used here.
float Cell2D::GetPerim() {
  float dist = 0.0;
  for (unsigned int i = 0; i < NV; i++) {
    const unsigned int n = (i + 1 == NV) ? 0 : i + 1;
    const float dx = Verticies[n][0] - Verticies[i][0], dy = Verticies[n][1] - Verticies[i][1];
    dist += sqrt(dx * dx + dy * dy);
  }
  return dist;
}

// ---- 3D ---------------------------------------------------------------------------
// reference src/cell.cpp:62-158
Cell3D::Cell3D(std::array<float, 3> start, float calA, float radius) : Cell3D(start, calA, radius, 2) {}

Cell3D::Cell3D(std::array<float, 3> start, float calA, float radius, int subdiv) {
  if (subdiv < 0 || subdiv > 3) throw std::invalid_argument("Cell3D: subdivisions must be 0..3 (12, 42, 162 or 642 vertices)");
  subdivisions = subdiv;
  calA0 = calA;
  r0 = radius;
  Kv = 0.0f;
  Ka = 0.0f;
  Ks = 0.0f;  // uninitialised in the reference
  const unsigned int nf = 20u << (2 * subdiv), nv = nf / 2 + 2;  // 4^s * 20 faces; Euler: V = F/2 + 2
  std::vector<float> unit(3 * nv);
  std::vector<uint32_t> tri(3 * nf);
  int gv = 0, gf = 0;
  if (dpm_icosphere(subdiv, unit.data(), tri.data(), &gv, &gf) != DPM_OK || gv != (int)nv || gf != (int)nf)
    throw std::runtime_error("icosphere construction failed");
  Verts.resize(nv);
  Forces.assign(nv, {0.0f, 0.0f, 0.0f});
  Faces.resize(nf);
  for (unsigned int f = 0; f < nf; f++) Faces[f] = {tri[3 * f], tri[3 * f + 1], tri[3 * f + 2]};
  for (unsigned int v = 0; v < nv; v++)
    for (int d = 0; d < 3; d++) {
      float x = unit[3 * v + d];
      x *= r0;
      x += start[d];
      Verts[v][d] = x;
    }
  v0 = (4.0f / 3.0f) * M_PI * pow(r0, 3);
  sa0 = pow((6 * sqrt(M_PI) * v0 * calA), (2.0f / 3.0f));
  a0 = (sa0 / (float)nf);
  Volume = GetVolume();
  SurfaceArea = GetSurfaceArea();
}

// |sum_f (v1 x v2) . v0| / 6 accumulated in float (src/cell.cpp:196-211)
float Cell3D::GetVolume() {
  float vol = 0.0;
  for (const auto &t : Faces) {
    const auto &p0 = Verts[t[0]], &p1 = Verts[t[1]], &p2 = Verts[t[2]];
    const float cx = p1[1] * p2[2] - p1[2] * p2[1];
    const float cy = p1[2] * p2[0] - p1[0] * p2[2];
    const float cz = p1[0] * p2[1] - p1[1] * p2[0];
    const float part = cx * p0[0] + cy * p0[1] + cz * p0[2];
    vol += part;
  }
  return std::abs(vol) / 6.0;
}

// sum_f |(v1-v0) x (v2-v0)| — NOT halved: the reference returns twice the area
// (src/cell.cpp:213-230) and callers only test it for finiteness/sign.
float Cell3D::GetSurfaceArea() {
  float area = 0.0;
  for (const auto &t : Faces) {
    const auto &p0 = Verts[t[0]], &p1 = Verts[t[1]], &p2 = Verts[t[2]];
    const float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
    const float bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
    const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    const float part = std::sqrt(cx * cx + cy * cy + cz * cz);
    area += part;
  }
  return area;
}

static void need_reference_mesh(const Cell3D &c, const char *what) {
  if (c.nverts() != Cell3D::NV || c.nfaces() != Cell3D::NF)
    throw std::logic_error(std::string(what) + ": fixed-size result needs the 162-vertex mesh; use the ...V() form");
}

std::array<std::array<float, Cell3D::NV>, 3> Cell3D::GetPositions() {
  need_reference_mesh(*this, "GetPositions");
  std::array<std::array<float, NV>, 3> out;
  for (unsigned int v = 0; v < NV; v++)
    for (int d = 0; d < 3; d++) out[d][v] = Verts[v][d];
  return out;
}

// wraps the monolayer around a cylinder of circumference L (src/cell.cpp:242-253)
std::array<std::array<float, 162>, 3> Cell3D::GetVesselPositions(float L) {
  need_reference_mesh(*this, "GetVesselPositions");
  std::array<std::array<float, NV>, 3> out;
  const std::vector<std::vector<float>> v = GetVesselPositionsV(L);
  for (int d = 0; d < 3; d++)
    for (unsigned int i = 0; i < NV; i++) out[d][i] = v[d][i];
  return out;
}

std::array<std::array<float, Cell3D::NV>, 3> Cell3D::GetForces() {
  need_reference_mesh(*this, "GetForces");
  std::array<std::array<float, NV>, 3> out;
  for (unsigned int v = 0; v < NV; v++)
    for (int d = 0; d < 3; d++) out[d][v] = Forces[v][d];
  return out;
}

std::array<std::array<int, 3>, Cell3D::NF> Cell3D::GetFaces() {
  need_reference_mesh(*this, "GetFaces");
  std::array<std::array<int, 3>, NF> out;
  for (unsigned int f = 0; f < NF; f++)
    for (int k = 0; k < 3; k++) out[f][k] = (int)Faces[f][k];
  return out;
}

std::vector<std::vector<float>> Cell3D::GetPositionsV() {
  std::vector<std::vector<float>> out(3, std::vector<float>(nverts()));
  for (unsigned int v = 0; v < nverts(); v++)
    for (int d = 0; d < 3; d++) out[d][v] = Verts[v][d];
  return out;
}

std::vector<std::vector<float>> Cell3D::GetVesselPositionsV(float L) {
  std::vector<std::vector<float>> out(3, std::vector<float>(nverts()));
  const float scale = (2.0 * M_PI) / L;
  const float radius = L / (2 * M_PI);
  for (unsigned int v = 0; v < nverts(); v++) {
    const float theta = Verts[v][0] * scale;
    out[0][v] = (radius - Verts[v][2]) * cos(theta);
    out[2][v] = (radius - Verts[v][2]) * sin(theta);
    out[1][v] = Verts[v][1];
  }
  return out;
}

std::vector<std::vector<float>> Cell3D::GetForcesV() {
  std::vector<std::vector<float>> out(3, std::vector<float>(nverts()));
  for (unsigned int v = 0; v < nverts(); v++)
    for (int d = 0; d < 3; d++) out[d][v] = Forces[v][d];
  return out;
}

std::vector<std::array<int, 3>> Cell3D::GetFacesV() {
  std::vector<std::array<int, 3>> out(nfaces());
  for (unsigned int f = 0; f < nfaces(); f++)
    for (int k = 0; k < 3; k++) out[f][k] = (int)Faces[f][k];
  return out;
}

std::array<float, 3> Cell3D::GetCOM() {
  std::array<float, 3> com = {0.0, 0.0, 0.0};
  const unsigned int n = nverts();
  for (unsigned int v = 0; v < n; v++)
    for (int d = 0; d < 3; d++) com[d] += Verts[v][d];
  for (int d = 0; d < 3; d++) com[d] /= (float)n;
  return com;
}

}  // namespace DPM
