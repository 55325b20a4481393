// dpm_capi.cu — general entry points of the C ABI: errors, device query, geometry helpers.
#include <array>
#include <cmath>
#include <cstring>
#include <vector>

#include "dpm_common.cuh"

namespace dpm {
static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}
}  // namespace dpm

extern "C" {

const char *dpm_version(void) { return "opencl_dpm_b200 0.1 (sm_100a)"; }

int dpm_last_error(char *buf, size_t n) {
  if (!buf || n == 0) return DPM_ERR_INVALID_ARGUMENT;
  strncpy(buf, dpm::g_last_error.c_str(), n - 1);
  buf[n - 1] = 0;
  return DPM_OK;
}

int dpm_pin_host_buffer(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return dpm::fail(DPM_ERR_INVALID_ARGUMENT, "NULL buffer");
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return DPM_OK; }
  if (e != cudaSuccess) { cudaGetLastError(); return dpm::fail(DPM_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
  return DPM_OK;
}

int dpm_unpin_host_buffer(void *ptr) {
  if (!ptr) return DPM_OK;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) cudaGetLastError();  // not registered (any more): nothing to release
  return DPM_OK;
}

int dpm_device_count(int *count) {
  if (!count) return dpm::fail(DPM_ERR_INVALID_ARGUMENT, "count is NULL");
  *count = 0;
  DPM_CUDA_TRY(cudaGetDeviceCount(count));
  return DPM_OK;
}

// Icosphere with the reference's construction semantics (src/cell.cpp:62-140,:160-194):
// 12 normalised icosahedron vertices, 20 faces, `subdiv` rounds of 1->4 splitting where
// each new midpoint is "normalised" component by component with the norm re-evaluated
// after every division (so midpoints are NOT exactly on the unit sphere).
// Midpoints are shared through a (Cantor-pair key -> index) cache searched linearly in
// insertion order, which fixes the vertex numbering.
int dpm_icosphere(int subdiv, float *V, uint32_t *F, int *nv_out, int *nf_out) {
  if (subdiv < 0 || subdiv > 5 || !V || !F) return dpm::fail(DPM_ERR_INVALID_ARGUMENT, "dpm_icosphere: bad arguments");
  const float t = (float)((1 + std::sqrt(5)) / 2);
  const float base[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                             {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
  std::vector<std::array<float, 3>> verts(12);
  for (int i = 0; i < 12; i++) {
    float norm = std::sqrt(base[i][0] * base[i][0] + base[i][1] * base[i][1] + base[i][2] * base[i][2]);
    for (int d = 0; d < 3; d++) verts[i][d] = base[i][d] / norm;
  }
  std::vector<std::array<uint32_t, 3>> faces = {
      {0, 11, 5}, {0, 5, 1},  {0, 1, 7},   {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
      {3, 9, 4},  {3, 4, 2},  {3, 2, 6},   {3, 6, 8},  {3, 8, 9},   {4, 9, 5}, {2, 4, 11}, {6, 2, 10},  {8, 6, 7},  {9, 8, 1}};
  std::vector<std::pair<long long, uint32_t>> cache;
  auto midpoint = [&](uint32_t p1, uint32_t p2) -> uint32_t {
    long long key = (long long)(p1 + p2) * (p1 + p2 + 1) / 2 + (p1 < p2 ? p1 : p2);
    key = (int)key;  // the reference keeps the key in an int
    for (auto &kv : cache)
      if (kv.first == key) return kv.second;
    std::array<float, 3> m;
    for (int d = 0; d < 3; d++) {
      m[d] = verts[p2][d] + verts[p1][d];
      m[d] = (float)((double)m[d] * 0.5);
    }
    for (int d = 0; d < 3; d++) {
      float norm = std::sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
      m[d] /= norm;
    }
    verts.push_back(m);
    uint32_t idx = (uint32_t)verts.size() - 1;
    cache.emplace_back(key, idx);
    return idx;
  };
  for (int s = 0; s < subdiv; s++) {
    std::vector<std::array<uint32_t, 3>> nf;
    nf.reserve(faces.size() * 4);
    for (auto &f : faces) {
      uint32_t a = midpoint(f[0], f[1]), b = midpoint(f[1], f[2]), c = midpoint(f[2], f[0]);
      nf.push_back({f[0], a, c});
      nf.push_back({f[1], b, a});
      nf.push_back({f[2], c, b});
      nf.push_back({a, b, c});
    }
    faces.swap(nf);
  }
  for (size_t i = 0; i < verts.size(); i++)
    for (int d = 0; d < 3; d++) V[3 * i + d] = verts[i][d];
  for (size_t i = 0; i < faces.size(); i++)
    for (int d = 0; d < 3; d++) F[3 * i + d] = faces[i][d];
  if (nv_out) *nv_out = (int)verts.size();
  if (nf_out) *nf_out = (int)faces.size();
  return DPM_OK;
}

int dpm_cell3d_params(float calA, float r0, int nf, float *out4) {
  if (!out4 || nf <= 0) return dpm::fail(DPM_ERR_INVALID_ARGUMENT, "dpm_cell3d_params: bad arguments");
  float v0 = (float)((double)(4.0f / 3.0f) * M_PI * std::pow((double)r0, 3));           // src/cell.cpp:153
  float sa0 = (float)std::pow(6 * std::sqrt(M_PI) * (double)v0 * (double)calA, (double)(2.0f / 3.0f));  // :154
  float a0 = sa0 / (float)nf;                                                            // :155
  float l0 = (float)(std::sqrt((double)(4.0f * a0)) / std::sqrt((double)3.0f));        // src/Tissue3D.cpp:177
  out4[0] = v0; out4[1] = sa0; out4[2] = a0; out4[3] = l0;
  return DPM_OK;
}

}  // extern "C"
