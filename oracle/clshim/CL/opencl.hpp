// TEST INFRASTRUCTURE — a minimal hand-written stand-in for the Khronos OpenCL C++ bindings, covering exactly
// the subset the reference's host code uses (src/Tissue2D.cpp, src/Tissue3D.cpp): Platform/Device::getDefault,
// Context, Program(build, getBuildInfo), Buffer, Kernel::setArg, NDRange/NullRange, CommandQueue
// (enqueueNDRangeKernel, enqueueReadBuffer, finish).  It lets the reference's UNMODIFIED sources compile in an
// image without CL headers.  All OpenCL entry points are resolved at run time from the ICD loader, so the
// library loads (and the reference's CPU-only methods such as Disperse work) where no OpenCL platform exists;
// compute calls then fail with an OpenCL error code exactly as they would with no device.
#ifndef ORACLE_CLSHIM_OPENCL_HPP
#define ORACLE_CLSHIM_OPENCL_HPP
#include <dlfcn.h>

#include <chrono>      // the Khronos header pulls these in transitively; the reference relies on that
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "cl.h"

namespace cl {
namespace detail {
struct Api {
  void *lib = nullptr;
  cl_int (*GetPlatformIDs)(cl_uint, cl_platform_id *, cl_uint *) = nullptr;
  cl_int (*GetDeviceIDs)(cl_platform_id, cl_device_type, cl_uint, cl_device_id *, cl_uint *) = nullptr;
  cl_context (*CreateContext)(const intptr_t *, cl_uint, const cl_device_id *, void *, void *, cl_int *) = nullptr;
  cl_program (*CreateProgramWithSource)(cl_context, cl_uint, const char **, const size_t *, cl_int *) = nullptr;
  cl_int (*BuildProgram)(cl_program, cl_uint, const cl_device_id *, const char *, void *, void *) = nullptr;
  cl_int (*GetProgramBuildInfo)(cl_program, cl_device_id, cl_program_build_info, size_t, void *, size_t *) = nullptr;
  cl_mem (*CreateBuffer)(cl_context, cl_mem_flags, size_t, void *, cl_int *) = nullptr;
  cl_kernel (*CreateKernel)(cl_program, const char *, cl_int *) = nullptr;
  cl_int (*SetKernelArg)(cl_kernel, cl_uint, size_t, const void *) = nullptr;
  cl_command_queue (*CreateCommandQueue)(cl_context, cl_device_id, cl_command_queue_properties, cl_int *) = nullptr;
  cl_int (*EnqueueNDRangeKernel)(cl_command_queue, cl_kernel, cl_uint, const size_t *, const size_t *, const size_t *, cl_uint,
                                 const cl_event *, cl_event *) = nullptr;
  cl_int (*EnqueueReadBuffer)(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void *, cl_uint, const cl_event *, cl_event *) = nullptr;
  cl_int (*Finish)(cl_command_queue) = nullptr;
  cl_int (*ReleaseMemObject)(cl_mem) = nullptr;
  cl_int (*ReleaseKernel)(cl_kernel) = nullptr;
  cl_int (*ReleaseCommandQueue)(cl_command_queue) = nullptr;
  cl_int (*GetDeviceInfo)(cl_device_id, cl_device_info, size_t, void *, size_t *) = nullptr;
  bool ok = false;
  Api() {
    // the NVIDIA driver ships an ICD but this image has no /etc/OpenCL/vendors entry for it
    setenv("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1", 0);
    const char *names[] = {"libOpenCL.so.1", "libOpenCL.so", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1", nullptr};
    for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
#define ORACLE_CL_SYM(member, name) member = reinterpret_cast<decltype(member)>(dlsym(lib, name))
    ORACLE_CL_SYM(GetPlatformIDs, "clGetPlatformIDs");
    ORACLE_CL_SYM(GetDeviceIDs, "clGetDeviceIDs");
    ORACLE_CL_SYM(CreateContext, "clCreateContext");
    ORACLE_CL_SYM(CreateProgramWithSource, "clCreateProgramWithSource");
    ORACLE_CL_SYM(BuildProgram, "clBuildProgram");
    ORACLE_CL_SYM(GetProgramBuildInfo, "clGetProgramBuildInfo");
    ORACLE_CL_SYM(CreateBuffer, "clCreateBuffer");
    ORACLE_CL_SYM(CreateKernel, "clCreateKernel");
    ORACLE_CL_SYM(SetKernelArg, "clSetKernelArg");
    ORACLE_CL_SYM(CreateCommandQueue, "clCreateCommandQueue");
    ORACLE_CL_SYM(EnqueueNDRangeKernel, "clEnqueueNDRangeKernel");
    ORACLE_CL_SYM(EnqueueReadBuffer, "clEnqueueReadBuffer");
    ORACLE_CL_SYM(Finish, "clFinish");
    ORACLE_CL_SYM(ReleaseMemObject, "clReleaseMemObject");
    ORACLE_CL_SYM(ReleaseKernel, "clReleaseKernel");
    ORACLE_CL_SYM(ReleaseCommandQueue, "clReleaseCommandQueue");
    ORACLE_CL_SYM(GetDeviceInfo, "clGetDeviceInfo");
#undef ORACLE_CL_SYM
    ok = GetPlatformIDs && GetDeviceIDs && CreateContext && CreateProgramWithSource && BuildProgram && CreateBuffer &&
         CreateKernel && SetKernelArg && CreateCommandQueue && EnqueueNDRangeKernel && EnqueueReadBuffer && Finish;
  }
};
inline Api &api() {
  static Api a;
  return a;
}
constexpr cl_int kNoOpenCL = -1001;  // CL_PLATFORM_NOT_FOUND_KHR
}  // namespace detail

class Platform {
public:
  cl_platform_id id = nullptr;
  static Platform getDefault() {
    Platform p;
    auto &a = detail::api();
    if (a.ok) {
      cl_uint n = 0;
      if (a.GetPlatformIDs(1, &p.id, &n) != CL_SUCCESS || n == 0) p.id = nullptr;
    }
    return p;
  }
};

class Device {
public:
  cl_device_id id = nullptr;
  static Device getDefault() {
    Device d;
    auto &a = detail::api();
    Platform p = Platform::getDefault();
    if (a.ok && p.id) {
      cl_uint n = 0;
      if (a.GetDeviceIDs(p.id, CL_DEVICE_TYPE_DEFAULT, 1, &d.id, &n) != CL_SUCCESS || n == 0) {
        d.id = nullptr;
        if (a.GetDeviceIDs(p.id, CL_DEVICE_TYPE_ALL, 1, &d.id, &n) != CL_SUCCESS || n == 0) d.id = nullptr;
      }
    }
    return d;
  }
};

class Context {
public:
  cl_context ctx = nullptr;
  Context() = default;
  explicit Context(const std::vector<Device> &devs) {
    auto &a = detail::api();
    if (!a.ok || devs.empty() || !devs[0].id) return;
    cl_int err = 0;
    cl_device_id d = devs[0].id;
    ctx = a.CreateContext(nullptr, 1, &d, nullptr, nullptr, &err);
    if (err != CL_SUCCESS) ctx = nullptr;
  }
};

class Program {
public:
  cl_program prog = nullptr;
  Program() = default;
  Program(const Context &c, const std::string &src) {
    auto &a = detail::api();
    if (!a.ok || !c.ctx) return;
    const char *s = src.c_str();
    size_t n = src.size();
    cl_int err = 0;
    prog = a.CreateProgramWithSource(c.ctx, 1, &s, &n, &err);
    if (err != CL_SUCCESS) prog = nullptr;
  }
  cl_int build(const std::vector<Device> &devs, const char *options = nullptr) {
    auto &a = detail::api();
    if (!a.ok || !prog || devs.empty() || !devs[0].id) return detail::kNoOpenCL;
    cl_device_id d = devs[0].id;
    return a.BuildProgram(prog, 1, &d, options, nullptr, nullptr);
  }
  template <cl_uint name>
  std::string getBuildInfo(const Device &dev) const {
    auto &a = detail::api();
    if (!a.ok || !prog || !a.GetProgramBuildInfo) return "no OpenCL platform";
    size_t n = 0;
    a.GetProgramBuildInfo(prog, dev.id, name, 0, nullptr, &n);
    std::string s(n, '\0');
    if (n) a.GetProgramBuildInfo(prog, dev.id, name, n, &s[0], nullptr);
    return s;
  }
};

class Buffer {
public:
  cl_mem mem = nullptr;
  Buffer() = default;
  Buffer(const Context &c, cl_mem_flags flags, size_t size, void *host_ptr = nullptr, cl_int *err = nullptr) {
    auto &a = detail::api();
    cl_int e = detail::kNoOpenCL;
    if (a.ok && c.ctx) mem = a.CreateBuffer(c.ctx, flags, size, host_ptr, &e);
    if (err) *err = e;
  }
  Buffer(const Buffer &) = delete;
  Buffer &operator=(const Buffer &) = delete;
  ~Buffer() {
    if (mem && detail::api().ReleaseMemObject) detail::api().ReleaseMemObject(mem);
  }
};

class Kernel {
public:
  cl_kernel k = nullptr;
  Kernel() = default;
  Kernel(const Program &p, const char *name, cl_int *err = nullptr) {
    auto &a = detail::api();
    cl_int e = detail::kNoOpenCL;
    if (a.ok && p.prog) k = a.CreateKernel(p.prog, name, &e);
    if (err) *err = e;
  }
  Kernel(const Kernel &) = delete;
  Kernel &operator=(const Kernel &) = delete;
  ~Kernel() {
    if (k && detail::api().ReleaseKernel) detail::api().ReleaseKernel(k);
  }
  cl_int setArg(cl_uint i, const Buffer &b) { return k ? detail::api().SetKernelArg(k, i, sizeof(cl_mem), &b.mem) : detail::kNoOpenCL; }
  template <typename T>
  cl_int setArg(cl_uint i, const T &v) {
    return k ? detail::api().SetKernelArg(k, i, sizeof(T), &v) : detail::kNoOpenCL;
  }
};

class NDRange {
public:
  size_t sizes[3] = {0, 0, 0};
  cl_uint dims = 0;
  NDRange() = default;
  NDRange(size_t a) : sizes{a, 1, 1}, dims(1) {}
  NDRange(size_t a, size_t b) : sizes{a, b, 1}, dims(2) {}
  NDRange(size_t a, size_t b, size_t c) : sizes{a, b, c}, dims(3) {}
};
static const NDRange NullRange;

class CommandQueue {
public:
  cl_command_queue q = nullptr;
  CommandQueue(const Context &c, const Device &d, cl_command_queue_properties props = 0, cl_int *err = nullptr) {
    auto &a = detail::api();
    cl_int e = detail::kNoOpenCL;
    if (a.ok && c.ctx && d.id) q = a.CreateCommandQueue(c.ctx, d.id, props, &e);
    if (err) *err = e;
  }
  CommandQueue(const CommandQueue &) = delete;
  ~CommandQueue() {
    if (q && detail::api().ReleaseCommandQueue) detail::api().ReleaseCommandQueue(q);
  }
  cl_int enqueueNDRangeKernel(const Kernel &k, const NDRange &offset, const NDRange &global, const NDRange &local = NullRange) {
    if (!q || !k.k) return detail::kNoOpenCL;
    return detail::api().EnqueueNDRangeKernel(q, k.k, global.dims, offset.dims ? offset.sizes : nullptr, global.sizes,
                                              local.dims ? local.sizes : nullptr, 0, nullptr, nullptr);
  }
  cl_int enqueueReadBuffer(const Buffer &b, cl_bool blocking, size_t offset, size_t size, void *ptr) {
    if (!q || !b.mem) return detail::kNoOpenCL;
    return detail::api().EnqueueReadBuffer(q, b.mem, blocking, offset, size, ptr, 0, nullptr, nullptr);
  }
  cl_int finish() { return q ? detail::api().Finish(q) : detail::kNoOpenCL; }
};

}  // namespace cl
#endif
