/*
 * dpm_b200.h — C ABI of the B200-native replacement for the per-timestep hot
 * path of sudo-shaka/OpenCL_DPM (force evaluation + overdamped Euler step).
 *
 * This is the seam that replaces, in the reference,
 *   - readKernelSource() + cl::Program::build()        src/readKernel.cpp:4-10, src/Tissue3D.cpp:199, src/Tissue2D.cpp:142
 *   - the cl::Buffer uploads                           src/Tissue3D.cpp:208-281, src/Tissue2D.cpp:149-168
 *   - the per-step enqueueNDRangeKernel sequence       src/Tissue3D.cpp:372-445, src/Tissue2D.cpp:215-229
 *   - the blocking enqueueReadBuffer calls             src/Tissue3D.cpp:425-434,:464-470, src/Tissue2D.cpp:223-233
 * i.e. everything inside Tissue{2D,3D}::CLEulerUpdate between "pack" and
 * "unpack".  The host classes (opencl_dpm_b200/host) call only this header.
 *
 * Conventions
 *   - plain C, borrowed host pointers valid for the duration of the call,
 *     device memory owned by the handle; no C++ exceptions cross the ABI;
 *   - every function returns 0 on success or a DPM_ERR_* code; the message of
 *     the last failure on the calling thread is read with dpm_last_error();
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point returns DPM_ERR_CUDA;
 *   - one CUDA stream per handle (its own unless dpm*_set_stream is called);
 *     dpm*_step is asynchronous on that stream, dpm*_download / dpm*_euler_update
 *     synchronise before returning;
 *   - 3D positions/forces use the reference's device layout: float4-strided
 *     (cl_float3 == 16 B, src/Tissue3D.cpp:139-141), 2D uses float2 padded to
 *     maxNV per cell (src/Tissue2D.cpp:117-120,139).
 */
#ifndef DPM_B200_H
#define DPM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPM_OK 0
#define DPM_ERR_INVALID_ARGUMENT 1 /* maps to std::invalid_argument (src/Tissue3D.cpp:123-135) */
#define DPM_ERR_RUNTIME 2          /* maps to std::runtime_error   (src/Tissue3D.cpp:149-206) */
#define DPM_ERR_CUDA 3             /* device API failure -> std::runtime_error with cudaGetErrorString */
#define DPM_ERR_TOPOLOGY 4         /* face list is not a closed, consistently oriented 2-manifold */
#define DPM_ERR_NCCL 5

/* Which force terms run (all on by default; used by the parity tests to isolate kernels). */
#define DPM3D_VOLUME 1u /* VolumeForceUpdate       shaders/Cell3D_Kernel.cl:66-112  */
#define DPM3D_AREA 2u   /* SurfaceAreaForceUpdate  shaders/Cell3D_Kernel.cl:114-177 */
#define DPM3D_STICK 4u  /* StickToSurface          shaders/Cell3D_Kernel.cl:180-247 */
#define DPM3D_REPEL 8u  /* RepellingForces         shaders/Cell3D_Kernel.cl:251-310 */
#define DPM3D_ALL 15u   /* the kernels the reference host enqueues (src/Tissue3D.cpp:372-423): the default mask */
/* AllVertAttraction (shaders/Cell3D_Kernel.cl:313-364): vertex-vertex attraction between cells, strength Kat.  The
 * reference compiles it but never enqueues it, so it is NOT part of the default mask: a caller opts in
 * with dpm3d_set_force_mask(h, DPM3D_ALL | DPM3D_ATTRACT) and a non-zero Kat; it then runs between RepellingForces
 * and EulerPosition.  Evaluated in gather form (each vertex sums what it adds to itself and what the other vertex's
 * work-item would scatter onto it), so ghost cells of a sharded run stay read-only. */
#define DPM3D_ATTRACT 16u

#define DPM2D_AREA 1u      /* AreaForceUpdates       shaders/Cell2D_kernel.cl:13-42   */
#define DPM2D_PERIMETER 2u /* PerimeterForceUpdates  shaders/Cell2D_kernel.cl:92-119  */
#define DPM2D_BENDING 4u   /* BendingForceUpdates    shaders/Cell2D_kernel.cl:44-90   */
#define DPM2D_ATTRACT 8u   /* AttractionForceUpdate  shaders/Cell2D_kernel.cl:222-268 */
#define DPM2D_REPEL 16u    /* RepulsionForceUpdate   shaders/Cell2D_kernel.cl:121-220 */
#define DPM2D_ALL 31u

typedef struct dpm3d_ctx dpm3d_t;
typedef struct dpm2d_ctx dpm2d_t;

/* Grid of the cell list (integer artefact; must equal oracle_cell_list's bit for bit). */
typedef struct {
  int32_t nb[3];
  int32_t periodic[3];
  int32_t allpass[3];
  float origin[3];
  float inv_binw[3];
  float max_ext;
  float margin;
  int32_t nbins;
  int32_t pad;
} dpm_grid_t;

/* Counters accumulated since create (or the last dpm*_reset_stats). */
typedef struct {
  uint64_t steps;         /* timesteps executed */
  uint64_t launches;      /* kernels launched by this library */
  uint64_t rebuilds;      /* neighbour-list rebuilds */
  uint64_t contact_evals; /* 3D: (vertex, cell) winding-number evaluations; 2D: (vertex, cell) polygon tests */
  uint64_t halo_bytes;    /* bytes sent to other ranks */
  uint64_t reserved[3];
} dpm_stats_t;

/* ---- general ------------------------------------------------------------ */
const char *dpm_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf. */
int dpm_last_error(char *buf, size_t n);
int dpm_device_count(int *count);
/* Page-lock (and later release) a caller-owned host buffer so that the uploads / downloads of dpm*_upload, dpm*_download
 * and dpm*_euler_update run at full PCIe speed without a bounce buffer.  Optional: every entry point also accepts pageable
 * memory (the reference's host arrays are plain std::vector storage, src/Tissue3D.cpp:139-141).  Re-pinning a buffer that is
 * already pinned is not an error. */
int dpm_pin_host_buffer(void *ptr, size_t bytes);
int dpm_unpin_host_buffer(void *ptr);

/* Geometry helpers restating the reference constructors (src/cell.cpp:62-158, :12-33) so
 * that flat tissues can be built without the C++ classes. subdiv = 2 is the reference mesh
 * (162 vertices / 320 faces); subdiv = 3 gives 642 / 1280.  verts3: nv*3, faces: nf*3. */
int dpm_icosphere(int subdiv, float *verts3, uint32_t *faces, int *nv, int *nf);
/* out4 = {v0, sa0, a0, l0}  (src/cell.cpp:153-155, src/Tissue3D.cpp:177) */
int dpm_cell3d_params(float calA, float r0, int nf, float *out4);

/* ---- 3D:  replaces shaders/Cell3D_Kernel.cl + src/Tissue3D.cpp:199-470 ---- */

/* faces: nf*3 vertex indices of the shared topology (the reference uploads Cells[0].Faces
 * only, src/Tissue3D.cpp:144-146).  The mesh must be a closed oriented 2-manifold. */
int dpm3d_create(dpm3d_t **h, int device, int ncells, int nv, int nf, const uint32_t *faces);
int dpm3d_destroy(dpm3d_t *h);
/* Use the caller's CUDA stream (cudaStream_t passed as void*); NULL restores the handle's own. */
int dpm3d_set_stream(dpm3d_t *h, void *cuda_stream);
/* skin_rel: Verlet skin as a fraction of the largest cell extent (default 0.1);
 * max_candidates: per-cell candidate-list capacity K (default 32). Takes effect at the next upload. */
int dpm3d_set_neighbor_params(dpm3d_t *h, float skin_rel, int max_candidates);
/* The values in force: dpm3d_euler_update doubles max_candidates when a candidate list overflows, so a caller that
 * sizes the `cand` buffer of dpm3d_get_neighbor_artifacts must ask for the current K first. */
int dpm3d_get_neighbor_params(dpm3d_t *h, float *skin_rel, int *max_candidates);
int dpm3d_set_force_mask(dpm3d_t *h, unsigned mask);
/* Reference-race compatibility (off by default).  VolumeForceUpdate writes cellVolumes[ci] from work-item fi==0
 * and reads it back behind a work-group-scoped barrier (shaders/Cell3D_Kernel.cl:74-83) while the launch leaves
 * the local size to the runtime (src/Tissue3D.cpp:383): on NVIDIA's OpenCL the local size is 160, so faces
 * 160..319 read the volume of the PREVIOUS step (0 on the first step of every CLEulerUpdate call).  The default
 * (stale_volume_from_face < 0) implements the intended semantics — every face sees the current volume;
 * dpm3d_set_compat(h, 160) reproduces what the reference actually computes on NVIDIA hardware, bit-for-bit in
 * structure, so the CUDA path can be compared with the reference's own outputs (tests/golden). */
int dpm3d_set_compat(dpm3d_t *h, int stale_volume_from_face);

/* verts4: ncells*nv*4 floats (x,y,z,pad).  Per-cell arrays of length ncells:
 * Kv,Ka,Ks (src/Tissue3D.cpp:172-174), v0,a0 (:175-176), l0 = sqrt(4 a0)/sqrt(3) (:177). */
int dpm3d_upload(dpm3d_t *h, const float *verts4, const float *Kv, const float *Ka, const float *Ks,
                 const float *v0, const float *a0, const float *l0);
/* Same, but verts4_dev is a DEVICE pointer (inputs already resident in HBM). */
int dpm3d_upload_device(dpm3d_t *h, const float *verts4_dev, const float *Kv, const float *Ka, const float *Ks,
                        const float *v0, const float *a0, const float *l0);
/* nsteps of {ClearForces, Volume, SurfaceArea, StickToSurface, Repelling, EulerPosition}
 * (src/Tissue3D.cpp:372-423) fused; asynchronous.  Kat is only used when the force mask contains DPM3D_ATTRACT
 * (the reference never launches AllVertAttraction, SURVEY F12: by default Kat has no effect, as in the reference).
 * Every 1000 timesteps the call synchronises and checks the device-side error flags, as the reference drains its
 * queue "to catch errors early" (src/Tissue3D.cpp:437-444): a capacity overflow is reported then, not nsteps later. */
int dpm3d_step(dpm3d_t *h, int nsteps, float dt, float Kre, float Kat, int pbc, float L);
int dpm3d_sync(dpm3d_t *h);
/* verts4 and/or forces4 may be NULL. forces4 = forces of the last executed step. */
int dpm3d_download(dpm3d_t *h, float *verts4, float *forces4);
/* Device pointers of the current state (valid until the next step/upload). */
int dpm3d_device_state(dpm3d_t *h, float **verts4_dev, float **forces4_dev);

/* The whole reference seam in one call: upload, nsteps, download (host buffers).
 * verts4 is updated in place, forces4 (may be NULL) receives the last step's forces,
 * loop_ms (may be NULL) the CUDA-event time of the step loop only — what the reference's
 * own timer brackets (src/Tissue3D.cpp:369,454). */
int dpm3d_euler_update(dpm3d_t *h, float *verts4, float *forces4, const float *Kv, const float *Ka,
                       const float *Ks, const float *v0, const float *a0, const float *l0, int nsteps, float dt,
                       float Kre, float Kat, int pbc, float L, float *loop_ms);

/* Neighbour-search artefacts of the most recent rebuild (any pointer may be NULL).
 * bin_start has grid.nbins+1 entries, cand is ncells*K with K = max_candidates. Syncs. */
int dpm3d_get_neighbor_artifacts(dpm3d_t *h, dpm_grid_t *grid, int32_t *bin_id, int32_t *order,
                                 int32_t *bin_start, int32_t *cand_count, int32_t *cand);
/* Forces a rebuild of the lists from the current positions (asynchronous). */
int dpm3d_rebuild_neighbors(dpm3d_t *h, int pbc, float L);
/* Per-cell scalars of the current state: bounds12 = ncells*12 floats
 * {lo.xyz, r2max, hi.xyz, contact pad, com.xyz, volume}; COM and volume are evaluated in the reference's serial
 * summation order with individually rounded operations (bit-comparable with the CPU oracle). Syncs. */
int dpm3d_get_cell_bounds(dpm3d_t *h, float *bounds12);
int dpm3d_get_stats(dpm3d_t *h, dpm_stats_t *out);
int dpm3d_reset_stats(dpm3d_t *h);

/* ---- 3D multi-GPU: x-slab decomposition with per-step halo exchange (NCCL) ---- */
/* 128-byte NCCL unique id: rank 0 calls dpm_nccl_unique_id and broadcasts the bytes
 * (torch.distributed / MPI / a file); every rank then calls dpm3d_shard_init. */
int dpm_nccl_unique_id(uint8_t id[128]);
/* Turns a fresh handle into one shard: its ncells are the cells OWNED by this rank (any static assignment is
 * correct; x-slabs keep the halo small); max_ghost = ghost cells accepted from EACH of the two neighbouring
 * slabs.  Call before the first upload, on every rank (it creates the NCCL communicator).  Afterwards
 * dpm3d_step exchanges the boundary cells' vertices with ncclSend/ncclRecv before every timestep. */
int dpm3d_shard_init(dpm3d_t *h, int rank, int nranks, const uint8_t id[128], int max_ghost);
/* Global ids of the owned cells; candidate lists are ordered by global id so that a sharded run sums forces in
 * the single-GPU order (bit-identical results). */
int dpm3d_set_global_ids(dpm3d_t *h, const int32_t *gid);

/* ---- 2D:  replaces shaders/Cell2D_kernel.cl + src/Tissue2D.cpp:142-233 ---- */
int dpm2d_create(dpm2d_t **h, int device, int ncells, int max_nv);
int dpm2d_destroy(dpm2d_t *h);
int dpm2d_set_stream(dpm2d_t *h, void *cuda_stream);
int dpm2d_set_neighbor_params(dpm2d_t *h, float skin_rel, int max_candidates);
int dpm2d_get_neighbor_params(dpm2d_t *h, float *skin_rel, int *max_candidates);
int dpm2d_set_force_mask(dpm2d_t *h, unsigned mask);
/* verts2: ncells*max_nv*2 floats padded as the reference pads (src/Tissue2D.cpp:139);
 * nv: vertices per cell; Ka,Kl,Kb,a0,l0,r0 per cell (src/Tissue2D.cpp:129-136). */
int dpm2d_upload(dpm2d_t *h, const float *verts2, const int32_t *nv, const float *Ka, const float *Kl,
                 const float *Kb, const float *a0, const float *l0, const float *r0);
int dpm2d_step(dpm2d_t *h, int nsteps, float dt, float Kre, float Kat, int pbc, float L);
int dpm2d_sync(dpm2d_t *h);
int dpm2d_download(dpm2d_t *h, float *verts2, float *forces2);
int dpm2d_euler_update(dpm2d_t *h, float *verts2, float *forces2, const int32_t *nv, const float *Ka,
                       const float *Kl, const float *Kb, const float *a0, const float *l0, const float *r0,
                       int nsteps, float dt, float Kre, float Kat, int pbc, float L, float *loop_ms);
int dpm2d_get_neighbor_artifacts(dpm2d_t *h, dpm_grid_t *grid, int32_t *bin_id, int32_t *order,
                                 int32_t *bin_start, int32_t *cand_count, int32_t *cand);
int dpm2d_rebuild_neighbors(dpm2d_t *h, float Kat, int pbc, float L);
int dpm2d_get_stats(dpm2d_t *h, dpm_stats_t *out);
int dpm2d_reset_stats(dpm2d_t *h);

#ifdef __cplusplus
}
#endif
#endif /* DPM_B200_H */
