// TEST INFRASTRUCTURE — C entry points over the REAL reference (sudo-shaka/OpenCL_DPM), compiled from its own
// sources where they lie (/root/reference/src/{cell,Tissue2D,Tissue3D}.cpp, see oracle/Makefile) into
// oracle/_ref/libref_dpm.so.  The only pieces that are not the reference's are
//   * the OpenCL header shim (oracle/clshim), because the image has no CL headers, and
//   * readKernelSource(), which returns the reference's .cl texts embedded at build time with .incbin
//     (the reference reads ./shaders/*.cl relative to the CWD; on the GPU box /root/reference does not exist).
// With NVIDIA's OpenCL ICD (present on the B200 box) this runs the reference's actual CLEulerUpdate — host
// loop and OpenCL kernels — and is used to pin the CPU oracle and to time the reference arm of bench.py.
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "Tissue.hpp"
#include "cell.hpp"

__asm__(".section .rodata\n"
        ".global ref_cl3d_src\nref_cl3d_src:\n.incbin \"" REF_SHADER_DIR "/Cell3D_Kernel.cl\"\n.byte 0\n"
        ".global ref_cl2d_src\nref_cl2d_src:\n.incbin \"" REF_SHADER_DIR "/Cell2D_kernel.cl\"\n.byte 0\n"
        ".text\n");
extern "C" const char ref_cl3d_src[];
extern "C" const char ref_cl2d_src[];

// replaces src/readKernel.cpp: same signature, embedded text instead of a CWD-relative file
std::string readKernelSource(const std::string &filename) {
  if (filename.find("Cell3D_Kernel.cl") != std::string::npos) return std::string(ref_cl3d_src);
  if (filename.find("Cell2D_kernel.cl") != std::string::npos) return std::string(ref_cl2d_src);
  throw std::runtime_error("Failed to  find kernel file: " + filename);
}

static std::string g_err;

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// 1 if an OpenCL device can be had (the reference's compute path can run), else 0
int ref_available() {
  cl::Device d = cl::Device::getDefault();
  return d.id != nullptr;
}

int ref_device_name(char *buf, int n) {
  cl::Device d = cl::Device::getDefault();
  if (!d.id || !cl::detail::api().GetDeviceInfo) return 0;
  return cl::detail::api().GetDeviceInfo(d.id, CL_DEVICE_NAME, (size_t)n, buf, nullptr) == CL_SUCCESS;
}

// ---- geometry straight from the reference's constructors ------------------------------------------
// scal = {calA0, r0, v0, sa0, a0, Volume, SurfaceArea}
void ref_cell3d(const float *start, float calA, float r0, float *verts3, unsigned *faces3, float *scal) {
  DPM::Cell3D c({start[0], start[1], start[2]}, calA, r0);
  for (unsigned i = 0; i < c.NV; i++) for (int d = 0; d < 3; d++) verts3[3 * i + d] = c.Verts[i][d];
  for (unsigned i = 0; i < c.NF; i++) for (int d = 0; d < 3; d++) faces3[3 * i + d] = c.Faces[i][d];
  scal[0] = c.calA0; scal[1] = c.r0; scal[2] = c.v0; scal[3] = c.sa0; scal[4] = c.a0; scal[5] = c.Volume; scal[6] = c.SurfaceArea;
}
// scal = {calA0, a0, l0, r0, GetArea()}
void ref_cell2d(float x0, float y0, float calA, unsigned nv, float r0, float *verts2, float *scal) {
  DPM::Cell2D c(x0, y0, calA, nv, r0);
  for (unsigned i = 0; i < nv; i++) { verts2[2 * i] = c.Verticies[i][0]; verts2[2 * i + 1] = c.Verticies[i][1]; }
  scal[0] = c.calA0; scal[1] = c.a0; scal[2] = c.l0; scal[3] = c.r0; scal[4] = c.GetArea();
}

// ---- the reference's own initialisers (CPU only) ---------------------------------------------------
// n identical Cell3D(start, calA, r0) -> Tissue3D(cells, phi0) -> Disperse2D(); returns L
float ref_disperse3d(int n, const float *start, float calA, float r0, float phi0, float *verts3) {
  DPM::Cell3D c({start[0], start[1], start[2]}, calA, r0);
  std::vector<DPM::Cell3D> cells(n, c);
  DPM::Tissue3D T(cells, phi0);
  T.Disperse2D();
  for (int ci = 0; ci < n; ci++)
    for (unsigned i = 0; i < 162; i++) for (int d = 0; d < 3; d++) verts3[(ci * 162 + i) * 3 + d] = T.Cells[ci].Verts[i][d];
  return T.L;
}
float ref_disperse2d(int n, float calA, unsigned nv, float r0, float phi0, float *verts2) {
  DPM::Cell2D c(0, 0, calA, nv, r0);
  std::vector<DPM::Cell2D> cells(n, c);
  DPM::Tissue2D T(cells, phi0);
  T.Disperse();
  for (int ci = 0; ci < n; ci++)
    for (unsigned i = 0; i < nv; i++) { verts2[(ci * nv + i) * 2] = T.cells[ci].Verticies[i][0]; verts2[(ci * nv + i) * 2 + 1] = T.cells[ci].Verticies[i][1]; }
  return T.L;
}

// ---- Tissue3D::CLEulerUpdate of the reference --------------------------------------------------------
// verts3: nc*162*3 in/out; forces3: nc*162*3 out (may be NULL). Returns 0, or -1 with ref_last_error().
// seconds (may be NULL): wall time of the CLEulerUpdate call (JIT build + upload + loop + download).
int ref3d_euler(int nc, float *verts3, float *forces3, const float *Kv, const float *Ka, const float *Ks, const float *v0,
                const float *a0, float Kre, int PBC, float L, int nsteps, float dt, double *seconds) {
  try {
    DPM::Cell3D proto({0.f, 0.f, 0.f}, 1.0f, 1.0f);
    std::vector<DPM::Cell3D> cells(nc, proto);
    for (int ci = 0; ci < nc; ci++) {
      DPM::Cell3D &c = cells[ci];
      c.Kv = Kv[ci]; c.Ka = Ka[ci]; c.Ks = Ks[ci]; c.v0 = v0[ci]; c.a0 = a0[ci];
      for (unsigned i = 0; i < 162; i++) for (int d = 0; d < 3; d++) c.Verts[i][d] = verts3[(ci * 162 + i) * 3 + d];
    }
    DPM::Tissue3D T(cells, 1.0f);
    T.Kre = Kre; T.Kat = 0.0f; T.PBC = PBC; T.L = L;
    auto t0 = std::chrono::steady_clock::now();
    T.CLEulerUpdate(nsteps, dt);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int ci = 0; ci < nc; ci++)
      for (unsigned i = 0; i < 162; i++) for (int d = 0; d < 3; d++) {
        verts3[(ci * 162 + i) * 3 + d] = T.Cells[ci].Verts[i][d];
        if (forces3) forces3[(ci * 162 + i) * 3 + d] = T.Cells[ci].Forces[i][d];
      }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// ---- AllVertAttraction of the reference's Cell3D_Kernel.cl, run on its own -----------------------------------------
// The reference host never enqueues this kernel (it is compiled with the program but no cl::Kernel is made for it),
// so there is no host code of the reference to call: this entry builds the reference's UNMODIFIED kernel text with the
// reference's build options (src/Tissue3D.cpp:199) and launches that one kernel on zeroed forces over the
// (NV, NCELLS) range its get_global_id(0/1) indexing expects (the reference's vertCellSize, src/Tissue3D.cpp:360).
// verts3: nc*162*3 in; forces3: nc*162*3 out.  Returns 0, or -1 with ref_last_error().
int ref3d_attract(int nc, const float *verts3, float *forces3, const float *l0, float L, int PBC, float Kat) {
  try {
    cl::Device device = cl::Device::getDefault();
    if (!device.id) { g_err = "no OpenCL device"; return -1; }
    cl::Context context({device});
    cl::Program program(context, std::string(ref_cl3d_src));
    cl_int err = program.build({device}, "-cl-opt-disable -Werror");
    if (err != CL_SUCCESS) { g_err = "build failed: " + program.getBuildInfo<CL_PROGRAM_BUILD_LOG>(device); return -1; }
    const int NV = 162;
    std::vector<cl_float3> V((size_t)nc * NV), F((size_t)nc * NV);
    for (size_t i = 0; i < V.size(); i++) {
      V[i] = cl_float3{};
      F[i] = cl_float3{};
      V[i].s[0] = verts3[3 * i]; V[i].s[1] = verts3[3 * i + 1]; V[i].s[2] = verts3[3 * i + 2];
    }
    std::vector<float> l0v(l0, l0 + nc);
    cl::Buffer gV(context, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, sizeof(cl_float3) * V.size(), V.data(), &err);
    if (err != CL_SUCCESS) { g_err = "verts buffer"; return -1; }
    cl::Buffer gF(context, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, sizeof(cl_float3) * F.size(), F.data(), &err);
    if (err != CL_SUCCESS) { g_err = "forces buffer"; return -1; }
    cl::Buffer gl0(context, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, sizeof(float) * nc, l0v.data(), &err);
    if (err != CL_SUCCESS) { g_err = "l0 buffer"; return -1; }
    cl::Kernel k(program, "AllVertAttraction", &err);
    if (err != CL_SUCCESS) { g_err = "kernel AllVertAttraction not found"; return -1; }
    k.setArg(0, gV); k.setArg(1, gF); k.setArg(2, gl0); k.setArg(3, L); k.setArg(4, nc); k.setArg(5, PBC); k.setArg(6, Kat);
    cl::CommandQueue queue(context, device, 0, &err);
    if (err != CL_SUCCESS) { g_err = "queue"; return -1; }
    err = queue.enqueueNDRangeKernel(k, cl::NullRange, cl::NDRange(NV, nc));
    if (err != CL_SUCCESS) { g_err = "enqueue failed: " + std::to_string(err); return -1; }
    err = queue.enqueueReadBuffer(gF, CL_TRUE, 0, sizeof(cl_float3) * F.size(), F.data());
    if (err != CL_SUCCESS) { g_err = "read failed: " + std::to_string(err); return -1; }
    for (size_t i = 0; i < F.size(); i++) for (int d = 0; d < 3; d++) forces3[3 * i + d] = F[i].s[d];
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// ---- Tissue2D::CLEulerUpdate of the reference --------------------------------------------------------
// verts2/forces2: nc*S*2 padded (S >= max nv). NOTE the reference calls exit(0) if the OpenCL build fails
// (src/Tissue2D.cpp:143-148): call ref_available() first.
int ref2d_euler(int nc, int S, const int *nv, float *verts2, float *forces2, const float *Ka, const float *Kl, const float *Kb,
                const float *a0, const float *l0, const float *r0, float Kre, float Kat, int PBC, float L, int nsteps, float dt,
                double *seconds) {
  try {
    if (!ref_available()) { g_err = "no OpenCL device"; return -1; }
    std::vector<DPM::Cell2D> cells;
    for (int ci = 0; ci < nc; ci++) {
      DPM::Cell2D c(0.f, 0.f, 1.0f, (unsigned)nv[ci], r0[ci]);
      c.Ka = Ka[ci]; c.Kl = Kl[ci]; c.Kb = Kb[ci]; c.a0 = a0[ci]; c.l0 = l0[ci]; c.r0 = r0[ci];
      for (int i = 0; i < nv[ci]; i++) { c.Verticies[i][0] = verts2[(ci * S + i) * 2]; c.Verticies[i][1] = verts2[(ci * S + i) * 2 + 1]; }
      cells.push_back(c);
    }
    DPM::Tissue2D T(cells, 1.0f);
    T.Kre = Kre; T.Kat = Kat; T.PBC = PBC != 0; T.L = L;
    auto t0 = std::chrono::steady_clock::now();
    T.CLEulerUpdate(nsteps, dt);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int ci = 0; ci < nc; ci++)
      for (int i = 0; i < nv[ci]; i++) for (int d = 0; d < 2; d++) {
        verts2[(ci * S + i) * 2 + d] = T.cells[ci].Verticies[i][d];
        if (forces2) forces2[(ci * S + i) * 2 + d] = T.cells[ci].Forces[i][d];
      }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
