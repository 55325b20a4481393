// TEST INFRASTRUCTURE (oracle/): asks the box's OpenCL runtime (the one the reference runs on) what its built-ins return
// at the edge cases the DPM kernels can reach: normalize() of a zero vector (AttractionForceUpdate when two vertices of
// different cells coincide, shaders/Cell2D_kernel.cl:263-264) and a few others.  Build: g++ -std=c++17 -Iclshim
// cl_semantics_probe.cpp -ldl -o _ref/cl_semantics_probe ; run with OCL_ICD_FILENAMES=libnvidia-opencl.so.1 if needed.
#include <cstdio>
#include <string>
#include <vector>
#include <CL/opencl.hpp>
static const char *SRC = R"CLC(
__kernel void probe(__global float *out) {
  float2 z2 = (float2)(0.0f, 0.0f);
  float3 z3 = (float3)(0.0f, 0.0f, 0.0f);
  float2 n2 = normalize(z2);
  float3 n3 = normalize(z3);
  float2 t2 = (float2)(1e-30f, 0.0f);
  float2 m2 = normalize(t2);
  float d = 0.0f;
  float2 f = ((0.5f / 64) * d / 0.1f) * normalize(z2);
  out[0] = n2.x; out[1] = n2.y; out[2] = n3.x; out[3] = n3.y; out[4] = n3.z;
  out[5] = m2.x; out[6] = m2.y; out[7] = f.x; out[8] = f.y;
  out[9] = length(z2); out[10] = atan2(0.0f, 0.0f); out[11] = round(2.5f); out[12] = round(-2.5f);
  float2 tiny = (float2)(1e-22f, 1e-22f);
  float2 nt = normalize(tiny);
  out[13] = nt.x; out[14] = nt.y;
}
)CLC";
int main() {
  cl::Device dev = cl::Device::getDefault();
  cl::Context ctx({dev});
  cl::Program prog(ctx, std::string(SRC));
  cl_int e = prog.build({dev}, "-cl-opt-disable");
  if (e != CL_SUCCESS) { printf("build failed %d: %s\n", e, prog.getBuildInfo<CL_PROGRAM_BUILD_LOG>(dev).c_str()); return 1; }
  std::vector<float> out(16, -777.0f);
  cl::Buffer buf(ctx, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, sizeof(float) * out.size(), out.data());
  cl::Kernel k(prog, "probe");
  k.setArg(0, buf);
  cl::CommandQueue q(ctx, dev);
  e = q.enqueueNDRangeKernel(k, cl::NullRange, cl::NDRange(1), cl::NullRange);
  q.finish();
  q.enqueueReadBuffer(buf, CL_TRUE, 0, sizeof(float) * out.size(), out.data());
  printf("enqueue err %d\n", e);
  printf("normalize((float2)0)      = (%g, %g)\n", out[0], out[1]);
  printf("normalize((float3)0)      = (%g, %g, %g)\n", out[2], out[3], out[4]);
  printf("normalize((1e-30,0))      = (%g, %g)\n", out[5], out[6]);
  printf("(Kat/n*0/l0)*normalize(0) = (%g, %g)\n", out[7], out[8]);
  printf("length(0)=%g atan2(0,0)=%g round(2.5)=%g round(-2.5)=%g\n", out[9], out[10], out[11], out[12]);
  printf("normalize((1e-22,1e-22))  = (%g, %g)\n", out[13], out[14]);
  return 0;
}
