/* TEST INFRASTRUCTURE — minimal hand-written OpenCL C declarations (the image ships no CL headers).
 * Only what the reference's host code touches is declared; values are those of the Khronos
 * OpenCL specification.  Entry points are NOT linked: opencl.hpp resolves them at run time from
 * the ICD loader with dlopen/dlsym, so a binary built here loads on machines without OpenCL. */
#ifndef ORACLE_CLSHIM_CL_H
#define ORACLE_CLSHIM_CL_H
#include <stddef.h>
#include <stdint.h>

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef float cl_float;
typedef cl_uint cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_command_queue_properties;
typedef cl_uint cl_program_build_info;
typedef cl_uint cl_platform_info;
typedef cl_uint cl_device_info;

typedef struct _cl_platform_id *cl_platform_id;
typedef struct _cl_device_id *cl_device_id;
typedef struct _cl_context *cl_context;
typedef struct _cl_command_queue *cl_command_queue;
typedef struct _cl_mem *cl_mem;
typedef struct _cl_program *cl_program;
typedef struct _cl_kernel *cl_kernel;
typedef struct _cl_event *cl_event;

/* 3-component vectors are 4-component vectors (16 bytes, 16-byte aligned) */
typedef union {
  cl_float s[4];
  struct { cl_float x, y, z, w; };
} __attribute__((aligned(16))) cl_float4;
typedef cl_float4 cl_float3;
typedef union {
  cl_uint s[4];
  struct { cl_uint x, y, z, w; };
} __attribute__((aligned(16))) cl_uint4;
typedef cl_uint4 cl_uint3;

#define CL_SUCCESS 0
#define CL_DEVICE_NOT_FOUND -1
#define CL_FALSE 0
#define CL_TRUE 1
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_DEVICE_TYPE_DEFAULT (1 << 0)
#define CL_DEVICE_TYPE_CPU (1 << 1)
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PLATFORM_NAME 0x0902
#define CL_DEVICE_NAME 0x102B
#endif
