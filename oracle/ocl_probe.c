/* TEST INFRASTRUCTURE (oracle/): probes whether an OpenCL platform exists on
 * this machine, through the ICD loader, without needing CL headers.
 * Prototypes are hand-declared from the OpenCL 1.2 C API. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef int cl_int; typedef unsigned cl_uint; typedef void *cl_platform_id; typedef void *cl_device_id;
typedef unsigned long cl_bitfield;
typedef cl_int (*pfn_getplat)(cl_uint, cl_platform_id *, cl_uint *);
typedef cl_int (*pfn_platinfo)(cl_platform_id, cl_uint, size_t, void *, size_t *);
typedef cl_int (*pfn_getdev)(cl_platform_id, cl_bitfield, cl_uint, cl_device_id *, cl_uint *);
typedef cl_int (*pfn_devinfo)(cl_device_id, cl_uint, size_t, void *, size_t *);
int main(void) {
  const char *names[] = {"libOpenCL.so.1", "libOpenCL.so", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1", 0};
  void *h = 0;
  for (int i = 0; names[i] && !h; i++) { h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL); if (h) printf("loader: %s\n", names[i]); }
  if (!h) { printf("no OpenCL loader: %s\n", dlerror()); return 0; }
  pfn_getplat gp = (pfn_getplat)dlsym(h, "clGetPlatformIDs");
  pfn_platinfo pi = (pfn_platinfo)dlsym(h, "clGetPlatformInfo");
  pfn_getdev gd = (pfn_getdev)dlsym(h, "clGetDeviceIDs");
  pfn_devinfo di = (pfn_devinfo)dlsym(h, "clGetDeviceInfo");
  cl_platform_id plats[8]; cl_uint np = 0;
  cl_int e = gp(8, plats, &np);
  printf("clGetPlatformIDs err=%d nplat=%u\n", e, np);
  for (cl_uint p = 0; p < np && p < 8; p++) {
    char buf[512]; buf[0] = 0;
    pi(plats[p], 0x0902 /*CL_PLATFORM_NAME*/, sizeof buf, buf, 0); printf("platform %u: %s", p, buf);
    pi(plats[p], 0x0901 /*CL_PLATFORM_VERSION*/, sizeof buf, buf, 0); printf(" | %s\n", buf);
    cl_device_id devs[16]; cl_uint nd = 0;
    e = gd(plats[p], 0xFFFFFFFF /*ALL*/, 16, devs, &nd);
    printf("  devices err=%d n=%u\n", e, nd);
    for (cl_uint d = 0; d < nd && d < 16; d++) { di(devs[d], 0x102B /*CL_DEVICE_NAME*/, sizeof buf, buf, 0); printf("  dev %u: %s\n", d, buf); }
  }
  return 0;
}
