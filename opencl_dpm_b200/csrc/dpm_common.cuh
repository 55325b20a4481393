// dpm_common.cuh — shared device/host helpers for the B200 DPM hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/dpm_b200.h"

namespace dpm {

// ---- error plumbing (no exceptions across the C ABI) ------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define DPM_CUDA_TRY(expr)                                                                       \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::dpm::fail(DPM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

// ---- neighbour-search state kept on the device ------------------------------
struct NbrState {
  dpm_grid_t grid;
  int rebuild;   // lists are stale: rebuild before the next step
  int overflow;  // a candidate list exceeded K
  int nbuilds;
  int next;  // 2D: cells near the global extremes (the only possible |d| > L wrap partners), length of NbrBuffers::ext_list
  unsigned long long contact_evals;
  unsigned long long literal_evals;  // 3D: units that needed the literal all-faces sum (fallback of the fast path)
  unsigned long long fallback_why[4];
  int unit_total;     // 3D: contact units queued this step (reset by the rebuild kernel, which runs first every step)
  int unit_overflow;  // 3D: the global unit list was too small  // [0] neighbour not star-shaped, [1] vertex within the pad of the COM, [2] walk limit, [3] ring limit
  // global raw (unwrapped) extent, for the 2D |d|>L partner search
  float glo[3];
  float ghi[3];
  float range;  // interaction range the current lists were built for (3D: 1.16 * max edge bound * headroom)
  float pad1;
};

// Cell-list buffers (all device pointers)
struct NbrBuffers {
  NbrState *st;
  const float4 *blo;  // per cell: (lo.xyz, *)  exact AABB of current positions
  const float4 *bhi;  // per cell: (hi.xyz, *)
  int blo_stride;     // float4 stride between cells in blo/bhi arrays
  float4 *bbox_lo;    // AABB at build time, grown by skin/2
  float4 *bbox_hi;
  int *bin_id;
  int *order;
  int *bin_count;  // cap+1, doubles as scatter fill counter
  int *bin_start;  // cap+1
  int *cand_count;
  int *cand;
  float *partial;  // [grid][16] block partials of the reduction
  int *chunk_sum;  // [grid]
  int nc;          // cells binned (owned + ghosts)
  const int *nc_dev;  // if non-null the binned cell count is read from the device (sharded runs: owned + received ghosts)
  const int *gid;     // if non-null, global cell ids: candidate lists are ordered by gid so that the force summation
                      // order (and hence every bit of the result) is independent of the decomposition
  int nc_list;     // cells that get candidate lists (owned)
  int nd;
  int cap;
  int K;
  int pbc;
  float L;
  float skin_rel;
  float range;        // explicit interaction range (range_from_bounds == 0)
  int range_from_bounds;  // 1: range = range_scale * max_c bhi[c].w (per-cell contact pad), recorded in st->range
  float range_scale;
  float att_pad_scale;  // 3D vertex-vertex attraction on: a cell's pad counts as max(pad, att_pad_scale * l0), l0 = blo[3].w
  int far2d;
  int *ext_list;  // far2d: [nc] scratch, the cells near the global extremes (filled by the rebuild kernel)
};

// Launches the rebuild kernel as one thread-block cluster, with programmatic dependent launch (no-op on the device when
// st->rebuild == 0); coop_grid = rebuild_max_grid(): the CTAs of that cluster (sizes of NbrBuffers::partial / chunk_sum).
cudaError_t launch_rebuild(const NbrBuffers &nb, cudaStream_t stream, int coop_grid);
int rebuild_max_grid(int device);

// ---- small device helpers -----------------------------------------------------
#ifdef __CUDACC__
// Programmatic dependent launch (kernels chained with cudaLaunchAttributeProgrammaticStreamSerialization):
// griddep_wait() blocks until the preceding kernel of the stream has completed and its writes are visible (a no-op when
// the kernel was launched without the attribute); griddep_launch() lets the following kernel's CTAs be scheduled once
// every CTA of this grid has called it or exited.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// x - L*roundf(x/L) with every operation individually rounded (bit-exact with the CPU spec)
__device__ __forceinline__ float minimg_rn(float d, float L) {
  return __fsub_rn(d, __fmul_rn(L, roundf(__fdiv_rn(d, L))));
}
__device__ __forceinline__ float wrap_rn(float c, int pbc, float L) {
  return pbc ? __fsub_rn(c, __fmul_rn(L, floorf(__fdiv_rn(c, L)))) : c;
}
#endif

}  // namespace dpm
