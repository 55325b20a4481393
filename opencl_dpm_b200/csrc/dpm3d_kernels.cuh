// dpm3d_kernels.cuh — the 3D force + integrate timestep for sm_100a.
//
// Replaces the six per-step OpenCL kernels of shaders/Cell3D_Kernel.cl
// (ClearForces :366, VolumeForceUpdate :66, SurfaceAreaForceUpdate :114,
//  StickToSurface :180, RepellingForces :251, EulerPosition :371 — and, opt-in,
//  AllVertAttraction :313) and their enqueue sequence (src/Tissue3D.cpp:372-423) by three
// kernels per timestep, chained with programmatic dependent launch behind the neighbour
// rebuild kernel of neighbor.cu:
//
//   dpm3d_units_kernel    CTA = one cell: every vertex against the padded bounding box and
//                         sphere of each cell-list candidate -> "contact units" (vertex,
//                         neighbour) in one global queue (exact cull);
//   dpm3d_contact_kernel  8 lanes per unit, load-balanced over the whole GPU: the reference's
//                         winding number of the unit by the exact decomposition
//                         w_ref = W - sum(skipped faces)  (walk to the pierced face + ring search);
//   dpm3d_step_kernel     CTA = one cell.  The cell's vertices (float4) arrive in shared memory by
//                         a bulk async copy; shape forces are GATHERED per vertex over a constant
//                         ring adjacency (no float atomics, unlike atomic_add_f :10-32); the units'
//                         weights are folded in; the Euler update writes the other position buffer
//                         (all forces of a step use start-of-step positions); the epilogue
//                         produces next step's per-cell bounds (AABB, COM, volume, r^2max, contact pad,
//                         star-shape flag, per-face flags) and raises the rebuild flag when a cell
//                         leaves its build-time box.
//
// Numerical contract: per-cell COM and signed-volume sums are
// evaluated in the reference's serial order with individually rounded operations,
// because the volume sum is ill-conditioned (terms ~ |x|^2 * edge) and any other
// order changes the strain at the 1e-4 level; everything else uses fused fp32.
#pragma once
#include "dpm_common.cuh"

namespace dpm {

struct Step3DParams {
  const float4 *__restrict__ pos_in;
  float4 *__restrict__ pos_out;
  float4 *__restrict__ force_out;  // nullptr unless this is the last step of a call
  const float4 *__restrict__ bnd_in;  // 3 float4 per cell: (lo,r2max) (hi,0) (com,vol)
  float4 *__restrict__ bnd_out;
  const float4 *__restrict__ cellA;  // (Kv, Ka, Ks, v0)
  const float4 *__restrict__ cellB;  // (a0, l0, 0, 0)
  const ushort4 *__restrict__ faces;
  const uint16_t *__restrict__ ring_nbr;   // [nv][ring_stride] BYTE OFFSETS (16 * vertex) of the ring neighbours; ring_stride == 8 (one 16-byte load) or 16
  const uint16_t *__restrict__ ring_face;
  const uint8_t *__restrict__ valence;
  int ring_stride;
  const unsigned char *__restrict__ flag_in;  // [cell][vflag_stride(nv)] per-vertex flags of the CURRENT positions (see vflag_stride),
  unsigned char *__restrict__ flag_out;       //   written by the previous epilogue / the bounds kernel
  uint2 *__restrict__ vlist;                  // [cell][<= nv] the cell's vertices that have units: (vertex, (offset within the
  int *__restrict__ vlist_cnt;                //   cell's unit range) << 8 | count), and how many there are
  const ushort4 *__restrict__ face_adj;    // face across edge (a,b), (b,c), (c,a)
  const uint16_t *__restrict__ ring_tab;   // per face: faces in BFS (edge-adjacency) order, RING_TAB entries
  const uint8_t *__restrict__ ring_end;    // per face: cumulative end of rings 0..RING_MAX
  const uint16_t *__restrict__ dir_table;  // octahedral direction map (DIR_N x DIR_N) -> face, walk start guess
  const int *__restrict__ cand_count;
  const int *__restrict__ cand;
  int K;
  const float4 *__restrict__ bbox_lo;
  const float4 *__restrict__ bbox_hi;
  NbrState *st;
  // contact units: (global vertex id, neighbour cell) pairs that survive the culls, built by dpm3d_units_kernel in a
  // per-cell contiguous range [unit_base[ci], +unit_cnt[ci]) of one global list, evaluated by dpm3d_contact_kernel
  int2 *__restrict__ unit_rec;
  float *__restrict__ unit_w;
  float4 *__restrict__ unit_att;  // DPM3D_ATTRACT: the unit's AllVertAttraction force on its vertex (else unused)
  int *__restrict__ unit_base;
  int *__restrict__ unit_cnt;
  int unit_cap;
  // neighbours that are NOT star-shaped about their kernel point: bounding boxes of their face patches (PATCH_F consecutive
  // faces), refreshed every timestep by the units kernel, searched by the contact kernel (winding_patches)
  const uint16_t *__restrict__ vorder;     // [nv] vertex taken by slot r of the ring pass (chosen on the host against bank conflicts)
  const ushort4 *__restrict__ faces_proc;  // [nf] (a, b, c, face id) in the face pass's processing order (a permutation inside blocks of 32)
  float *__restrict__ terms;       // [owned cell][terms_stride(nf)] signed-volume terms of the NEW positions (face pass -> chain warp)
  int *__restrict__ grp_done;      // [ceil(nc / CHAIN_GROUP)] CTAs of the group that have published their cell (self-resetting)
  float4 *__restrict__ part;       // [owned cell][STEP_WARPS][3] per-warp partials of the new positions: (lo.xyz, r2) (hi.xyz, 0) (e2, star, 0, 0)
  float4 *__restrict__ patch_box;  // [cell slot][npatch][2]: (lo.xyz, 0) (hi.xyz, 0)
  int npatch;
  const int *__restrict__ n_total_dev;  // sharded runs: owned + ghost cells present (device-side count); else nullptr
  // sharded runs, peer-memory halo: a cell on a send list is stored by the step kernel's epilogue straight into the neighbouring
  // rank's inbox over NVLink (positions by a second bulk store out of shared memory, bounds and global id by the chain warp),
  // so the exchange in front of the next timestep only has to release the arrival flags (dpm_halo.cu)
  const int2 *__restrict__ push_slot;  // [owned cell] slot in the message to peer 0 / peer 1, or -1; nullptr: no fused push
  const int *__restrict__ push_gid;    // [owned cell] global ids
  float4 *push_pos[2], *push_bnd[2];   // position / bounds areas of the peers' inbox buffers of the NEXT exchange
  int *push_gidp[2];
  int nc;  // cells stepped by this launch (owned)
  int nv, nf;
  float dt, Kc;
  float Kat;  // != 0 only when DPM3D_ATTRACT is in the mask
  int pbc;
  float L;
  unsigned mask;
  int stale_from;  // >= 0: reference-race compatibility mode (dpm3d_set_compat), else -1
};

#define DPM_PRAGMA_(x) _Pragma(#x)
#define DPM_UNROLL(n) DPM_PRAGMA_(unroll n)
#ifndef DPM_FACE_UNROLL
#define DPM_FACE_UNROLL 2  // unroll factors of the step kernel's face and vertex loops (sweep in profiles/)
#endif
#ifndef DPM_RING_UNROLL
#define DPM_RING_UNROLL 1
#endif
#ifndef DPM_STEP_MINB
#define DPM_STEP_MINB 9  // CTAs per SM promised to ptxas for the step kernel: 56 registers, no spills in the default instantiation
#endif
constexpr int BND = 5;              // float4 per cell in the bounds arrays:
                                    //   (lo.xyz, r2max about the kernel point) (hi.xyz, contact pad) (com.xyz, volume)
                                    //   (0, star flag, previous volume, l0) (kernel point.xyz, 0)
// The KERNEL POINT of a cell is an interior point fixed BEFORE the cell's new positions exist (the serial-order COM of the
// positions one timestep earlier; after an upload the COM itself): the centre of the bounding sphere, the apex of the
// star-shape test and the point the contact kernel walks from.  Nothing of the step kernel's epilogue therefore waits for
// the serial COM / volume chains of the NEW positions; those only feed the next timestep (shift, force direction, strain).
#ifndef DPM_CHAIN_GROUP
#define DPM_CHAIN_GROUP 8
#endif
#ifndef DPM_CHAIN_STAGES
#define DPM_CHAIN_STAGES 3
#endif
#ifndef DPM_CHAIN_BULK
#define DPM_CHAIN_BULK 0  // 0: per-lane cp.async (LDGSTS) staging of the chain operands; 1: bulk async copies (TMA unit) — measured 6 % slower
                          // per timestep on config D: its latency per chunk is longer and the chains of the last groups are the kernel's tail
#endif
constexpr int CHAIN_GROUP = DPM_CHAIN_GROUP;  // cells whose serial chains one warp evaluates together (4 chains per cell)
constexpr int CHAIN_CH = 32;        // chain iterations staged per chunk
constexpr int CHAIN_STAGES = DPM_CHAIN_STAGES;  // chunks in flight
constexpr int CHAIN_POS_STRIDE = CHAIN_CH * 16 + 16;  // bytes per cell and stage, padded: the 8 cells' lanes hit 8 different bank quads
constexpr int CHAIN_TERM_STRIDE = CHAIN_CH * 8 + 16;
constexpr int CHAIN_STAGE_BYTES = CHAIN_GROUP * (CHAIN_POS_STRIDE + CHAIN_TERM_STRIDE);
constexpr int CHAIN_SMEM = CHAIN_STAGES * CHAIN_STAGE_BYTES;
__host__ __device__ inline int terms_stride(int nf) { return ((nf + 2 * CHAIN_CH - 1) / (2 * CHAIN_CH)) * (2 * CHAIN_CH); }  // floats per cell
// The reference skips faces with denom < 1e-8 (shaders/Cell3D_Kernel.cl:293-295), i.e. every face that subtends
// at least pi steradians from the vertex, so its "winding number" is
//        w_ref(p) = W(p) - (1/4pi) * sum over skipped faces of Omega_f(p),      W = true winding number (0 or 1).
// A planar triangle subtends >= pi only if the vertex projects INSIDE it (otherwise it lies in an open half-plane
// seen from the foot point, < pi) and its height h above the plane satisfies h <= R/sqrt(3) <= e/3, R <= e/sqrt(3)
// being the radius of the triangle's enclosing circle and e its longest edge.  Hence a face can only be skipped if
// some point of it is within e/3 of the vertex: a vertex farther than CONTACT_PAD * emax from a cell's bounding
// box / sphere gets w_ref = W = 0 from it and is culled exactly.
constexpr float CONTACT_PAD = 0.34f;
constexpr float RANGE_HEADROOM = 1.25f;  // lists are built for pads up to 1.25x the largest current one
// AllVertAttraction reaches 2 * l0 of either cell (shaders/Cell3D_Kernel.cl:350): a vertex can only take part if it is
// within ATT_REACH * max(l0_i, l0_j) of the other cell's box and bounding sphere (the 1e-3 covers fp32 rounding of the test)
constexpr float ATT_REACH = 2.002f;
constexpr int RING_MAX = 6;              // edge-adjacency rings examined around the radially hit face
constexpr int RING_TAB = 64;             // 1 + 3 + 6 + 9 + 12 + 15 + 18 = 64 faces: one 64-bit hit mask
constexpr int DIR_N = 16;                // octahedral map resolution
constexpr int MAX_WALK = 64;
constexpr int UNIT_LANES = 8;            // lanes cooperating on one (vertex, neighbour) unit: ring faces / literal faces in parallel

__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = UNIT_LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// sP[nv] | sF[max(nv, ceil(nf/2))] (forces; the bounds kernel reuses it for the signed-volume terms) | per-vertex flags in | out
__host__ __device__ inline int step3d_mid_slots(int nv, int nf) { return nv > (nf + 1) / 2 ? nv : (nf + 1) / 2; }
// per-VERTEX flags of a cell (one byte per vertex, rows padded to 16 bytes for the bulk copies): bits 0..6 = number of
// adjacent faces that face the substrate (StickToSurface applies once per such face, :226-246), bit 7 = some adjacent face
// has a degenerate edge (SurfaceAreaForceUpdate skips it, :151).  Accumulated by the face pass of the previous epilogue.
__host__ __device__ inline int vflag_stride(int nv) { return (nv + 15) & ~15; }
__device__ __forceinline__ void vflag_add(unsigned char *sV, int v, unsigned bits) {
  atomicAdd(reinterpret_cast<unsigned *>(sV) + (v >> 2), bits << (8 * (v & 3)));
}
__device__ __forceinline__ void vflag_or(unsigned char *sV, int v, unsigned bits) {
  atomicOr(reinterpret_cast<unsigned *>(sV) + (v >> 2), bits << (8 * (v & 3)));
}
inline size_t step3d_smem_bytes(int nv, int nf) {
  const size_t cell = sizeof(float4) * ((size_t)nv + step3d_mid_slots(nv, nf)) + 2 * (size_t)vflag_stride(nv) + 64;
  return cell > (size_t)CHAIN_SMEM ? cell : (size_t)CHAIN_SMEM;  // the chain warp of a group's last CTA stages 8 cells in the same memory
}

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 sub3(float4 a, float4 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 sub3(float4 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// x / 6.0f, correctly rounded (== __fdiv_rn(x, 6.0f) for normal results): q0 = x * fl(1/6) is within 1 ulp, the
// remainder r = x - 6 q0 is exact in an FMA, and q0 + r * fl(1/6) rounds to the correctly rounded quotient
// (Markstein).  Results that are not comfortably normal take the IEEE division.
__device__ __forceinline__ float div6_rn(float x) {
  const float c = 0.16666667163372039794921875f;  // fl(1/6)
  const float q0 = __fmul_rn(x, c);
  const float r = __fmaf_rn(-6.0f, q0, x);
  const float q1 = __fmaf_rn(r, c, q0);
  return (fabsf(x) > 1e-30f && fabsf(x) < 1e30f) ? q1 : __fdiv_rn(x, 6.0f);
}

// 1/sqrt(x) as ONE MUFU.RSQ (rsqrtf() without -use_fast_math wraps it in denormal scaling: ~5 instructions);
// identical to rsqrtf for normal x, which squared lengths of live edges are.
__device__ __forceinline__ float rsqrt_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Programmatic dependent launch (the timestep's kernels are chained with cudaLaunchAttributeProgrammaticStreamSerialization):
// griddep_wait() blocks until the preceding kernel of the stream has completed and its writes are visible (a no-op when
// the kernel was launched without the attribute); griddep_launch() lets the following kernel's CTAs be scheduled once
// every CTA of this grid has called it or exited.
// (griddep_wait / griddep_launch live in dpm_common.cuh: the rebuild kernel in neighbor.cu is part of the chain)

// 8-byte shared-memory load at a 32-bit shared address + immediate offset (one LDS.64, no generic addressing)
template <int OFF>
__device__ __forceinline__ float2 lds_v2(unsigned addr) {
  float2 t;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(t.x), "=f"(t.y) : "r"(addr), "n"(OFF));
  return t;
}

// ---- 1-D bulk copies through the async proxy (TMA unit): one instruction moves a whole cell -------------------
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// octahedral direction -> texel of the DIR_N x DIR_N walk-start table (mirrored on the host in dpm3d.cu)
__host__ __device__ inline int octa_texel(float x, float y, float z) {
  const float s = fabsf(x) + fabsf(y) + fabsf(z);
  float ox = x / s, oy = y / s;
  if (z < 0.0f) {
    const float tx = (1.0f - fabsf(oy)) * (ox >= 0.0f ? 1.0f : -1.0f), ty = (1.0f - fabsf(ox)) * (oy >= 0.0f ? 1.0f : -1.0f);
    ox = tx; oy = ty;
  }
  int ix = (int)((ox * 0.5f + 0.5f) * DIR_N), iy = (int)((oy * 0.5f + 0.5f) * DIR_N);
  ix = ix < 0 ? 0 : (ix > DIR_N - 1 ? DIR_N - 1 : ix);
  iy = iy < 0 ? 0 : (iy > DIR_N - 1 ? DIR_N - 1 : iy);
  return iy * DIR_N + ix;
}

// den / num of the reference's solid-angle formula for the face with corner vectors a, b, c = V + shift - p
// (shaders/Cell3D_Kernel.cl:285-298); unit vectors by MUFU.RSQ + one Newton step.
__device__ __forceinline__ void solid_angle_terms(float3 a, float3 b, float3 c, float &den, float &num) {
  float da = dot3(a, a), db = dot3(b, b), dc = dot3(c, c);
  float ra = rsqrtf(da), rb = rsqrtf(db), rc = rsqrtf(dc);
  // OpenCL normalize(0) = 0 (the vertex coincides with a corner; measured on the reference's runtime, see oracle/)
  ra = da > 0.0f ? ra * (1.5f - 0.5f * da * ra * ra) : 0.0f;
  rb = db > 0.0f ? rb * (1.5f - 0.5f * db * rb * rb) : 0.0f;
  rc = dc > 0.0f ? rc * (1.5f - 0.5f * dc * rc * rc) : 0.0f;
  a = f3(a.x * ra, a.y * ra, a.z * ra); b = f3(b.x * rb, b.y * rb, b.z * rb); c = f3(c.x * rc, c.y * rc, c.z * rc);
  den = 1.0f + dot3(a, b) + dot3(b, c) + dot3(c, a);
  num = dot3(a, cross3(b, c));
}

// The reference's sum, literally, for one (vertex, neighbour) unit: every face, one lane.  Fallback of
// winding_fast (neighbour not star-shaped about its COM, vertex within the pad of the COM, walk/ring limits).
static __device__ __noinline__ float winding_literal(const float4 *__restrict__ Vj, const ushort4 *__restrict__ faces, int nf, float4 sh,
                                              float4 p, int g, unsigned gmask) {
  float om = 0.0f;
  for (int f = g; f < nf; f += UNIT_LANES) {
    const ushort4 fc = __ldg(faces + f);
    const float4 q0 = __ldg(Vj + fc.x), q1 = __ldg(Vj + fc.y), q2 = __ldg(Vj + fc.z);
    const float3 a = f3((q0.x + sh.x) - p.x, (q0.y + sh.y) - p.y, (q0.z + sh.z) - p.z);
    const float3 b = f3((q1.x + sh.x) - p.x, (q1.y + sh.y) - p.y, (q1.z + sh.z) - p.z);
    const float3 c = f3((q2.x + sh.x) - p.x, (q2.y + sh.y) - p.y, (q2.z + sh.z) - p.z);
    float den, num;
    solid_angle_terms(a, b, c, den, num);
    if (!(den < 1e-8f)) om += 2.0f * atan2f(num, den);  // :293-299
  }
  return group_sum(om, gmask) / (4.0f * 3.14159274101257f);
}

// w_ref = W - (1/4pi) * sum_{faces with den < 1e-8} Omega_f, for a neighbour that is star-shaped about its COM C
// (checked every step by the owner's epilogue):
//   1. the face f* hit by the ray C -> p is found by walking the spherical triangulation seen from C, starting
//      from a direction-table guess;  p is inside  <=>  p is on the inner side of f*'s plane   =>  W;
//   2. a skipped face has a point within rho = pad of p (see CONTACT_PAD), hence intersects the cone of half-angle
//      asin(rho / |p - C|) about the ray; faces meeting a cone form an edge-connected patch containing f*, so the
//      rings of f* are examined outwards until a whole ring misses the (conservatively tested) cone;
//   3. only those few faces get the reference's den / num / atan2.
// Omega of a SKIPPED face (den < 1e-8) in double precision from the reference's fp32 corner vectors: for a vertex
// nearly in the plane of a face that subtends ~pi both num and den are ~1e-3, and fp32 would lose 4 digits of
// exactly the term that W - sum(...) needs (the literal sum never evaluates these faces, so it does not suffer).
// Only the CANCELLATIONS need the extra digits: num and den are formed in fp64 (DFMA — full rate on B200 — from the
// reference's fp32 corner vectors); the three lengths come from one MUFU.RSQ + one Newton step in fp64 (relative error
// ~1e-14) instead of the software fp64 sqrt, and the angle from the fp32 atan2f of the rounded pair (its ~1e-7 relative
// error is far below the 1e-5 force tolerance; the software fp64 atan2 was most of this function's ~600 instructions).
__device__ __forceinline__ double rsqrt_f64(double x) {
  double r = (double)rsqrt_fast((float)x);
  return r * (1.5 - 0.5 * x * r * r);
}
__device__ __forceinline__ void solid_angle_terms_f64(float3 a, float3 b, float3 c, double &den, double &num) {
  const double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z, cx = c.x, cy = c.y, cz = c.z;
  const double da = ax * ax + ay * ay + az * az, db = bx * bx + by * by + bz * bz, dc = cx * cx + cy * cy + cz * cz;
  const double la = da * rsqrt_f64(da), lb = db * rsqrt_f64(db), lc = dc * rsqrt_f64(dc);
  num = ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
  den = la * lb * lc + (ax * bx + ay * by + az * bz) * lc + (bx * cx + by * cy + bz * cz) * la + (cx * ax + cy * ay + cz * az) * lb;
}
static __device__ __noinline__ float omega_skipped_f64(float3 a, float3 b, float3 c) {
  double den, num;
  solid_angle_terms_f64(a, b, c, den, num);
  // scale the pair into fp32's comfortable range (atan2 is homogeneous): both are ~|a||b||c| * 1e-3 or smaller
  const double m = fmax(fabs(num), fabs(den));
  if (!(m > 0.0)) return 0.0f;  // atan2(0, 0) = 0 on the reference's runtime (oracle/cl_semantics_probe.cpp)
  const double s = 1.0 / m;
  return 2.0f * atan2f((float)(num * s), (float)(den * s));
}

// Returns 0 on success, else the reason the caller must fall back to the literal sum (1: vertex within the pad of
// the neighbour's COM or exactly on one of its vertices, 2: walk limit, 3: ring limit).
// Warp-uniform version: the warp's 4 groups (UNIT_LANES = 8 lanes each) hold 4 different units and advance in
// lockstep (every loop runs while ANY group needs it, idle groups are predicated off), so the heavy per-face math is
// issued once for all four.  `active` = this lane's group has a unit.  Returns 0 on success, else the reason the
// caller must fall back to the literal sum (1: vertex within the pad of the neighbour's COM, 2: walk limit,
// 3: ring limit); group-uniform.
__device__ __forceinline__ int winding_fast(const Step3DParams &P, bool active, const float4 *__restrict__ Vj, float4 sh, float4 p,
                                            float3 Cs, float rho, float &w_out, int g, int gshift) {
  const unsigned FULL = 0xffffffffu;
  const float3 u = f3(p.x - Cs.x, p.y - Cs.y, p.z - Cs.z);
  const float r2 = dot3(u, u);
  int why = 0;
  if (active && !(r2 > 1.0201f * rho * rho)) why = 1;
  // ---- 1. walk to the face pierced by the ray C -> p ----------------------------------------------------
  int f = (active && why == 0) ? (int)__ldg(P.dir_table + octa_texel(u.x, u.y, u.z)) : 0;
  float3 A = f3(0.f, 0.f, 0.f), B = A, C = A;
  bool found = !(active && why == 0);
  for (int it = 0; it < MAX_WALK; it++) {
    if (!__any_sync(FULL, !found)) break;
    if (!found) {
      const ushort4 fc = __ldg(P.faces + f);
      const float4 q0 = __ldg(Vj + fc.x), q1 = __ldg(Vj + fc.y), q2 = __ldg(Vj + fc.z);
      A = f3((q0.x + sh.x) - Cs.x, (q0.y + sh.y) - Cs.y, (q0.z + sh.z) - Cs.z);
      B = f3((q1.x + sh.x) - Cs.x, (q1.y + sh.y) - Cs.y, (q1.z + sh.z) - Cs.z);
      C = f3((q2.x + sh.x) - Cs.x, (q2.y + sh.y) - Cs.y, (q2.z + sh.z) - Cs.z);
      const float d0 = dot3(u, cross3(A, B)), d1 = dot3(u, cross3(B, C)), d2 = dot3(u, cross3(C, A));
      const float dm = fminf(d0, fminf(d1, d2));
      if (dm >= 0.0f) found = true;
      else {
        const ushort4 ad = __ldg(P.face_adj + f);
        f = (dm == d0) ? ad.x : (dm == d1 ? ad.y : ad.z);
      }
    }
  }
  if (active && why == 0 && !found) why = 2;
  bool open = active && why == 0;  // still examining rings
  // inside <=> p on the inner side of the hit face's plane (the COM is on the inner side of every face)
  const float3 n = cross3(f3(B.x - A.x, B.y - A.y, B.z - A.z), f3(C.x - A.x, C.y - A.y, C.z - A.z));
  const float W = (dot3(n, f3(u.x - A.x, u.y - A.y, u.z - A.z)) < 0.0f) ? 1.0f : 0.0f;
  const float sinp2 = open ? rho * rho / r2 : 0.0f;
  const float s2 = sinp2 * r2 * 1.002f, c2 = (1.0f - sinp2) * r2 * 0.998f;
  // ---- 2./3. rings of the hit face in chunks of UNIT_LANES table entries ----------------------------------
  const uint16_t *tab = P.ring_tab + (size_t)f * RING_TAB;
  unsigned long long rends = 0;  // ring ends 0..RING_MAX packed, 8 bits each
  if (open) {
    const uint8_t *re = P.ring_end + (size_t)f * (RING_MAX + 1);
#pragma unroll
    for (int r = 0; r <= RING_MAX; r++) rends |= (unsigned long long)__ldg(re + r) << (8 * r);
  }
  float corr = 0.0f;
  unsigned long long hits = 0;
  int ring = 1;  // next ring whose completeness is checked (ring 0 = the hit face itself always touches the cone)
  for (int base = 0; base < RING_TAB; base += UNIT_LANES) {
    if (!__any_sync(FULL, open)) break;
    const int j = base + g;
    bool hit = false, coincident = false;
    const int jlast = (int)((rends >> (8 * RING_MAX)) & 0xff);
    if (open && j < jlast) {
      const int gf = __ldg(tab + j);
      const ushort4 gc = __ldg(P.faces + gf);
      const float4 q0 = __ldg(Vj + gc.x), q1 = __ldg(Vj + gc.y), q2 = __ldg(Vj + gc.z);
      const float3 a = f3((q0.x + sh.x) - p.x, (q0.y + sh.y) - p.y, (q0.z + sh.z) - p.z);  // reference: V + shift - p
      const float3 b = f3((q1.x + sh.x) - p.x, (q1.y + sh.y) - p.y, (q1.z + sh.z) - p.z);
      const float3 c = f3((q2.x + sh.x) - p.x, (q2.y + sh.y) - p.y, (q2.z + sh.z) - p.z);
      // Is the face, seen from C, within the cone's half-angle phi of the ray?  Exact test on the sphere of
      // directions: the ray pierces the spherical triangle, or passes within phi of a corner, or within phi of the
      // interior of an edge arc (Lagrange identity for (ga x u).(ga x gb) keeps it to dot products).
      const float3 ga = f3(a.x + u.x, a.y + u.y, a.z + u.z), gb = f3(b.x + u.x, b.y + u.y, b.z + u.z), gc2 = f3(c.x + u.x, c.y + u.y, c.z + u.z);
      const float ua = dot3(u, ga), ub = dot3(u, gb), uc = dot3(u, gc2);
      const float aa = dot3(ga, ga), bb = dot3(gb, gb), cc = dot3(gc2, gc2);
      const float ab = dot3(ga, gb), bc = dot3(gb, gc2), ca = dot3(gc2, ga);
      const float3 n0 = cross3(ga, gb), n1 = cross3(gb, gc2), n2 = cross3(gc2, ga);
      const float e0 = dot3(u, n0), e1 = dot3(u, n1), e2 = dot3(u, n2);
      hit = (e0 >= 0.0f && e1 >= 0.0f && e2 >= 0.0f);
      hit = hit || (ua > 0.0f && ua * ua >= c2 * aa) || (ub > 0.0f && ub * ub >= c2 * bb) || (uc > 0.0f && uc * uc >= c2 * cc);
      hit = hit || (e0 < 0.0f && e0 * e0 <= s2 * dot3(n0, n0) && aa * ub - ab * ua >= 0.0f && ua * bb - ub * ab >= 0.0f);
      hit = hit || (e1 < 0.0f && e1 * e1 <= s2 * dot3(n1, n1) && bb * uc - bc * ub >= 0.0f && ub * cc - uc * bc >= 0.0f);
      hit = hit || (e2 < 0.0f && e2 * e2 <= s2 * dot3(n2, n2) && cc * ua - ca * uc >= 0.0f && uc * aa - ua * ca >= 0.0f);
      if (hit) {
        float den, num;
        solid_angle_terms(a, b, c, den, num);
        if (den < 1e-8f) corr += omega_skipped_f64(a, b, c);
        // p ON a corner: the reference's normalize(0) = 0 zeroes that face's term, which W - sum(skipped) cannot express
        coincident = dot3(a, a) == 0.0f || dot3(b, b) == 0.0f || dot3(c, c) == 0.0f;
      }
    }
    const unsigned bal = __ballot_sync(FULL, hit);
    hits |= (unsigned long long)((bal >> gshift) & 0xffu) << base;
    if ((__ballot_sync(FULL, coincident) >> gshift) & 0xffu) { why = 1; open = false; }
    // every ring that is now completely examined must have touched the cone, else the patch is closed
    while (open && ring <= RING_MAX) {
      const int rb = (int)((rends >> (8 * (ring - 1))) & 0xff), re2 = (int)((rends >> (8 * ring)) & 0xff);
      if (re2 > base + UNIT_LANES) break;  // ring not completely examined yet
      const unsigned long long rmask = (re2 >= 64 ? ~0ull : ((1ull << re2) - 1ull)) & ~((1ull << rb) - 1ull);
      if ((hits & rmask) == 0ull) open = false;  // closed: nothing beyond this ring can touch the cone
      else ring++;
    }
    if (open && ring > RING_MAX) { why = 3; open = false; }  // the outermost tabulated ring still touches the cone
  }
  corr += __shfl_xor_sync(FULL, corr, 4);
  corr += __shfl_xor_sync(FULL, corr, 2);
  corr += __shfl_xor_sync(FULL, corr, 1);
  w_out = W - corr / (4.0f * 3.14159274101257f);
  return why;
}

// w_ref = W - (1/4pi) * sum_{faces with den < 1e-8} Omega_f for ANY closed, consistently oriented mesh (a crumpled cell
// that is no longer star-shaped, even a self-intersecting one: the sum of all solid angles is 4 pi W for every closed
// oriented surface).  The neighbour's faces are grouped in patches of PATCH_F consecutive faces whose bounding boxes the
// units kernel refreshes every timestep (only for cells that fail the star-shape test):
//   W        signed crossings of the ray p + t * ez, t > 0: only patches whose box the ray can hit are opened; a face is
//            crossed iff the origin lies inside its projection on z = 0 — edge functions with individually rounded
//            products, so the two faces of an edge see exactly opposite values (watertight), ties broken by vertex index —
//            and the triple product a.(b x c) has the sign of the projected area (crossing point above p);
//   skipped  a skipped face has a point within rho = pad of p (CONTACT_PAD): only patches whose box comes within rho of p
//            get the reference's den / num; for those faces the triple product is taken in fp64 so that the crossing and
//            the sign of the skipped solid angle can never disagree about the side of the plane p is on.
// One group of UNIT_LANES lanes per unit, group-uniform control flow.  Cost: npatch box tests + the faces of a few
// patches, instead of all nf faces with three normalisations and an atan2 each (winding_literal).
// Returns true if p coincides with a vertex of the neighbour (the caller takes the literal sum: normalize(0) = 0).
constexpr int PATCH_F = 16;  // faces per patch: 2 per lane of a unit group
// Warp-uniform like winding_fast: the warp's 4 groups hold 4 units and advance in lockstep (full-mask votes; a vote with a
// partial mask in divergent code costs a convergence barrier each time: 27 % of this function's stall samples before).
// `active` = this lane's group has a unit whose neighbour is not star-shaped.  The ray runs along z, away from the
// neighbour's kernel point: a vertex beside or outside the neighbour (the usual unit) then meets few or no patches, and
// outside the neighbour's exact bounding box W = 0 without any test.
static __device__ __noinline__ bool winding_patches(const ushort4 *__restrict__ faces, int nf, int npatch, bool active,
                                                    const float4 *__restrict__ Vj, const float4 *__restrict__ box, float4 sh, float4 p,
                                                    float kz, float4 blo, float4 bhi, float rho, float &w_out, int g, int gshift) {
  const unsigned FULL = 0xffffffffu;
  // kz: z of the neighbour's kernel point (shifted), blo / bhi: its exact bounding box (unshifted)
  const float3 ps = f3(p.x - sh.x, p.y - sh.y, p.z - sh.z);  // the boxes are those of the unshifted neighbour
  const float eps = 1e-5f * (fmaxf(fabsf(ps.x), fmaxf(fabsf(ps.y), fabsf(ps.z))) + fmaxf(fabsf(sh.x), fmaxf(fabsf(sh.y), fabsf(sh.z))) + 1.0f);
  const float reach = rho * 1.01f + eps, reach2 = reach * reach;
  const bool up = p.z >= kz;
  const bool inbox = active && ps.x >= blo.x - eps && ps.x <= bhi.x + eps && ps.y >= blo.y - eps && ps.y <= bhi.y + eps &&
                     ps.z >= blo.z - eps && ps.z <= bhi.z + eps;
  // ---- 1. every lane classifies its patches (independent loads, no votes): bit it <=> patch it * 8 + g ------------------
  unsigned raybits = 0u, nearbits = 0u;
  const int nit = (npatch + UNIT_LANES - 1) / UNIT_LANES;
  if (active) {
#pragma unroll 4
    for (int it = 0; it < nit; it++) {
      const int k = it * UNIT_LANES + g;
      if (k < npatch) {
        const float4 lo = __ldg(box + 2 * k), hi = __ldg(box + 2 * k + 1);
        const bool ray = inbox && ps.x >= lo.x - eps && ps.x <= hi.x + eps && ps.y >= lo.y - eps && ps.y <= hi.y + eps &&
                         (up ? ps.z <= hi.z + eps : ps.z >= lo.z - eps);
        const float dx = fmaxf(fmaxf(lo.x - ps.x, ps.x - hi.x), 0.0f), dy = fmaxf(fmaxf(lo.y - ps.y, ps.y - hi.y), 0.0f),
                    dz = fmaxf(fmaxf(lo.z - ps.z, ps.z - hi.z), 0.0f);
        const bool near = dx * dx + dy * dy + dz * dz <= reach2;
        raybits |= (ray ? 1u : 0u) << it;
        nearbits |= (near ? 1u : 0u) << it;
      }
    }
  }
  // ---- 2. the faces of the flagged patches, 2 per lane; every group pops its own patches, the warp stays in lockstep ----
  int W = 0;
  float corr = 0.0f;
  bool coincident = false;
  for (int it = 0; it < nit; it++) {
    const unsigned mn = (__ballot_sync(FULL, (nearbits >> it) & 1u) >> gshift) & 0xffu;
    unsigned m = ((__ballot_sync(FULL, (raybits >> it) & 1u) >> gshift) & 0xffu) | mn;
    while (__any_sync(FULL, m != 0u)) {
      const bool has = m != 0u;
      const int j = has ? __ffs(m) - 1 : 0;
      m &= m - 1;
      const bool nearp = has && ((mn >> j) & 1u);
#pragma unroll
      for (int hf = 0; hf < PATCH_F / UNIT_LANES; hf++) {
        const int f = (it * UNIT_LANES + j) * PATCH_F + hf * UNIT_LANES + g;
        if (!has || f >= nf) continue;
        const ushort4 fc = __ldg(faces + f);
        const float4 q0 = __ldg(Vj + fc.x), q1 = __ldg(Vj + fc.y), q2 = __ldg(Vj + fc.z);
        const float3 a = f3((q0.x + sh.x) - p.x, (q0.y + sh.y) - p.y, (q0.z + sh.z) - p.z);  // reference: V + shift - p
        const float3 b = f3((q1.x + sh.x) - p.x, (q1.y + sh.y) - p.y, (q1.z + sh.z) - p.z);
        const float3 c = f3((q2.x + sh.x) - p.x, (q2.y + sh.y) - p.y, (q2.z + sh.z) - p.z);
        // edge functions of the projection on z = 0, exactly antisymmetric under swapping the end points
        const float eab = __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
        const float ebc = __fsub_rn(__fmul_rn(b.x, c.y), __fmul_rn(b.y, c.x));
        const float eca = __fsub_rn(__fmul_rn(c.x, a.y), __fmul_rn(c.y, a.x));
        const bool pab = eab > 0.0f || (eab == 0.0f && fc.x < fc.y), pbc = ebc > 0.0f || (ebc == 0.0f && fc.y < fc.z),
                   pca = eca > 0.0f || (eca == 0.0f && fc.z < fc.x);
        const bool ccw = pab && pbc && pca, cw = !pab && !pbc && !pca;
        float tf = a.x * (b.y * c.z - b.z * c.y) + a.y * (b.z * c.x - b.x * c.z) + a.z * (b.x * c.y - b.y * c.x);
        if (nearp) {
          float den, num;
          solid_angle_terms(a, b, c, den, num);
          if (den < 1e-8f) {  // skipped by the reference (:293-295)
            double den64, num64;
            solid_angle_terms_f64(a, b, c, den64, num64);
            tf = num64 > 0.0 ? 1.0f : (num64 < 0.0 ? -1.0f : 0.0f);  // the crossing uses the sign the skipped solid angle has
            const double mm = fmax(fabs(num64), fabs(den64));
            if (mm > 0.0) corr += 2.0f * atan2f((float)(num64 / mm), (float)(den64 / mm));
          }
          coincident = coincident || dot3(a, a) == 0.0f || dot3(b, b) == 0.0f || dot3(c, c) == 0.0f;
        }
        // the triple product has the sign of the projected area <=> the crossing point lies above p
        if (up) {
          if (ccw && tf > 0.0f) W += 1;   // the ray leaves through a face whose normal points up
          if (cw && tf < 0.0f) W -= 1;    // the ray enters
        } else {
          if (cw && tf > 0.0f) W += 1;
          if (ccw && tf < 0.0f) W -= 1;
        }
      }
    }
  }
#pragma unroll
  for (int o = UNIT_LANES / 2; o > 0; o >>= 1) {
    W += __shfl_xor_sync(FULL, W, o);
    corr += __shfl_xor_sync(FULL, corr, o);
  }
  w_out = (float)W - corr / (4.0f * 3.14159274101257f);
  return ((__ballot_sync(FULL, coincident) >> gshift) & 0xffu) != 0u;
}

// Bounding boxes of the face patches of one cell (not star-shaped): 8 lanes per patch, 2 faces each; exact fp32 min / max.
__device__ __forceinline__ void patch_boxes(const Step3DParams &P, int ci, int tid, int nthreads) {
  const float4 *V = P.pos_in + (size_t)ci * P.nv;
  float4 *box = P.patch_box + (size_t)ci * P.npatch * 2;
  const int g = tid & (UNIT_LANES - 1);
  const int per = nthreads / UNIT_LANES;
  for (int k0 = 0; k0 < P.npatch; k0 += per) {  // uniform trip count over the CTA (full-warp shuffles below)
    const int k = k0 + tid / UNIT_LANES;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (k < P.npatch) {
#pragma unroll
      for (int hf = 0; hf < PATCH_F / UNIT_LANES; hf++) {
        const int f = k * PATCH_F + hf * UNIT_LANES + g;
        if (f >= P.nf) continue;
        const ushort4 fc = __ldg(P.faces + f);
        const float4 q0 = V[fc.x], q1 = V[fc.y], q2 = V[fc.z];
        lo[0] = fminf(lo[0], fminf(q0.x, fminf(q1.x, q2.x))); hi[0] = fmaxf(hi[0], fmaxf(q0.x, fmaxf(q1.x, q2.x)));
        lo[1] = fminf(lo[1], fminf(q0.y, fminf(q1.y, q2.y))); hi[1] = fmaxf(hi[1], fmaxf(q0.y, fmaxf(q1.y, q2.y)));
        lo[2] = fminf(lo[2], fminf(q0.z, fminf(q1.z, q2.z))); hi[2] = fmaxf(hi[2], fmaxf(q0.z, fmaxf(q1.z, q2.z)));
      }
    }
#pragma unroll
    for (int o = UNIT_LANES / 2; o > 0; o >>= 1)
#pragma unroll
      for (int d = 0; d < 3; d++) {
        lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
        hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
      }
    if (k < P.npatch && g == 0) {
      box[2 * k] = make_float4(lo[0], lo[1], lo[2], 0.f);
      box[2 * k + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------
// K1a: contact units.  One CTA per cell: every vertex is tested against the padded bounding box and sphere of each
// candidate neighbour (cell list) whose box overlaps the cell's own; survivors are written, ordered by (thread,
// vertex, ascending neighbour), into a contiguous range of the global unit list reserved with one atomicAdd.
// ---------------------------------------------------------------------------------
#ifndef DPM_UNITS_THREADS
#define DPM_UNITS_THREADS 128
#endif
constexpr int UNITS_THREADS = DPM_UNITS_THREADS;
constexpr int UNITS_KMAX = 128;

// ATT: DPM3D_ATTRACT is selected (the default instantiation carries none of the attraction code)
template <bool ATT>
__global__ void __launch_bounds__(UNITS_THREADS) dpm3d_units_kernel(Step3DParams P) {
  __shared__ float4 sLo[UNITS_KMAX], sHi[UNITS_KMAX], sSph[UNITS_KMAX];
  __shared__ int sCand[UNITS_KMAX];
  __shared__ int sWarp[UNITS_THREADS / 32], sWarpV[UNITS_THREADS / 32];
  __shared__ int sBase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ci = blockIdx.x, nv = P.nv, K = P.K;
  griddep_launch();
  // state of the previous timestep's step kernel (complete before the kernel ahead of this one started): read it now
  if (ci >= (P.n_total_dev ? *P.n_total_dev : P.nc)) return;  // sharded: the grid covers the ghost capacity
  // a cell that is not star-shaped about its kernel point (owned or ghost): refresh the boxes of its face patches, which the
  // contact kernel searches for every unit that has this cell as the neighbour (winding_patches)
  const float4 bi0 = P.bnd_in[BND * (size_t)ci], bi1 = P.bnd_in[BND * (size_t)ci + 1], bi2 = P.bnd_in[BND * (size_t)ci + 2],
               bi3 = P.bnd_in[BND * (size_t)ci + 3];
  if (bi3.y == 0.0f) patch_boxes(P, ci, tid, UNITS_THREADS);
  if (ci >= P.nc) return;  // ghost cell: no units of its own
  constexpr bool att = ATT;
  const float l0i = att ? bi3.w : 0.0f;
  griddep_wait();  // the candidate lists (and the unit counter) belong to the rebuild kernel ahead
  const int ncand = min(P.cand_count[ci], K);
  int nact = 0;
  for (int k = tid; k < ncand; k += UNITS_THREADS) {
    const int cj = P.cand[(size_t)ci * K + k];
    const float4 bj0 = P.bnd_in[BND * (size_t)cj], bj1 = P.bnd_in[BND * (size_t)cj + 1], bj2 = P.bnd_in[BND * (size_t)cj + 2];
    float3 sh = f3(0.f, 0.f, 0.f);
    if (P.pbc) {  // shift = L * round((COMi - COMJ) / L)   (:277-281)
      sh.x = P.L * roundf((bi2.x - bj2.x) / P.L);
      sh.y = P.L * roundf((bi2.y - bj2.y) / P.L);
      sh.z = P.L * roundf((bi2.z - bj2.z) / P.L);
    }
    // padded, shifted bounding box / sphere of cj: outside them the reference's formula gives exactly zero
    float pad = bj1.w;  // CONTACT_PAD * (longest edge of cj)
    bool nocull = false;
    if (att) {
      // AllVertAttraction pairs lie within 2 * max(l0_i, l0_j) of each other (:350).  It takes the minimum image of every
      // vertex pair (:338-343); that is the image of the COM shift unless the box is so small that a second image of cj can
      // come within reach of this cell as well: then every vertex is kept for cj (the evaluation itself is literal).
      pad = fmaxf(pad, ATT_REACH * fmaxf(l0i, P.bnd_in[BND * (size_t)cj + 3].w));
      if (P.pbc) {
        const float dmx = fmaxf(fabsf((bj1.x + sh.x) - bi0.x), fabsf(bi1.x - (bj0.x + sh.x)));
        const float dmy = fmaxf(fabsf((bj1.y + sh.y) - bi0.y), fabsf(bi1.y - (bj0.y + sh.y)));
        const float dmz = fmaxf(fabsf((bj1.z + sh.z) - bi0.z), fabsf(bi1.z - (bj0.z + sh.z)));
        nocull = !(fmaxf(dmx, fmaxf(dmy, dmz)) + pad < P.L);
      }
    }
    float4 lo = make_float4((bj0.x + sh.x) - pad, (bj0.y + sh.y) - pad, (bj0.z + sh.z) - pad, 0.f);
    float4 hi = make_float4((bj1.x + sh.x) + pad, (bj1.y + sh.y) + pad, (bj1.z + sh.z) + pad, 0.f);
    const float rs = sqrtf(bj0.w) + pad;
    const float4 bj4 = P.bnd_in[BND * (size_t)cj + 4];  // kernel point: centre of cj's bounding sphere
    float4 sph = make_float4(bj4.x + sh.x, bj4.y + sh.y, bj4.z + sh.z, rs * rs * 1.0001f + 1e-30f);
    if (nocull) {
      lo = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
      hi = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
      sph.w = INFINITY;
    }
    const bool ov = !(lo.x > bi1.x || hi.x < bi0.x || lo.y > bi1.y || hi.y < bi0.y || lo.z > bi1.z || hi.z < bi0.z);
    sCand[k] = ov ? cj : -1;
    sLo[k] = lo;
    sHi[k] = hi;
    sSph[k] = sph;
    nact += ov ? 1 : 0;
  }
  if (__syncthreads_or(nact) == 0) {  // no neighbour's box reaches this cell: its vertices are not even read
    if (tid == 0) { P.unit_base[ci] = 0; P.unit_cnt[ci] = 0; P.vlist_cnt[ci] = 0; }
    return;
  }
  // The active candidates, compacted (ascending k = ascending neighbour id).
  __shared__ int sAct[UNITS_KMAX];
  __shared__ int sNact;
  if (warp == 0) {
    int nb = 0;
    for (int k0 = 0; k0 < ncand; k0 += 32) {
      const int k = k0 + lane;
      const bool a = k < ncand && sCand[k] >= 0;
      const unsigned b = __ballot_sync(0xffffffffu, a);
      if (a) sAct[nb + __popc(b & ((1u << lane) - 1u))] = k;
      nb += __popc(b);
    }
    if (lane == 0) sNact = nb;
  }
  __syncthreads();
  const int na = sNact;
  const float4 *gP = P.pos_in + (size_t)ci * nv;
  auto inside = [&](const float4 &p, int k) {
    const float4 lo = sLo[k], hi = sHi[k], sp = sSph[k];
    const float dx = p.x - sp.x, dy = p.y - sp.y, dz = p.z - sp.z;
    return !(p.x < lo.x || p.x > hi.x || p.y < lo.y || p.y > hi.y || p.z < lo.z || p.z > hi.z) && (dx * dx + dy * dy + dz * dz) <= sp.w;
  };
  // Two passes over the thread's vertices (count, then emit) instead of per-vertex state in registers: the kernel is bound by
  // the latency of its dependent global loads times the number of CTA waves, so small CTAs with few registers — many of them
  // resident — matter more than the repeated tests (the second pass reads the vertices out of L1 / L2).
  int cnt = 0, nvh = 0;
  for (int v = tid; v < nv; v += UNITS_THREADS) {
    const float4 p = gP[v];
    int c = 0;
    for (int a = 0; a < na; a++) c += inside(p, sAct[a]) ? 1 : 0;
    cnt += c;
    nvh += c > 0 ? 1 : 0;
  }
  // exclusive scans of the per-thread counts: units, and vertices that have any
  int incl = cnt, vincl = nvh;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o), m = __shfl_up_sync(0xffffffffu, vincl, o);
    if (lane >= o) { incl += n; vincl += m; }
  }
  if (lane == 31) { sWarp[warp] = incl; sWarpV[warp] = vincl; }
  __syncthreads();
  int woff = 0, total = 0, vwoff = 0, vtotal = 0;
#pragma unroll
  for (int w = 0; w < UNITS_THREADS / 32; w++) {
    const int c = sWarp[w], d = sWarpV[w];
    if (w < warp) { woff += c; vwoff += d; }
    total += c; vtotal += d;
  }
  if (tid == 0) {
    int base = 0;
    if (total > 0) {
      base = atomicAdd(&P.st->unit_total, total);
      if (base + total > P.unit_cap) { P.st->unit_overflow = 1; base = -1; }
      else atomicAdd(&P.st->contact_evals, (unsigned long long)total);
    }
    sBase = base;
    P.unit_base[ci] = base < 0 ? 0 : base;
    P.unit_cnt[ci] = base < 0 ? 0 : total;
    P.vlist_cnt[ci] = base < 0 ? 0 : vtotal;
  }
  __syncthreads();
  if (total == 0 || sBase < 0 || cnt == 0) return;
  // emit: the thread's units in (vertex, ascending neighbour) order, and for every vertex that has any the entry
  // (vertex, offset within the cell's unit range << 8 | count) of the list the step kernel visits
  int off = woff + incl - cnt;
  uint2 *vl = P.vlist + (size_t)ci * nv + (vwoff + vincl - nvh);
  int2 *out = P.unit_rec + sBase + off;
  for (int v = tid; v < nv; v += UNITS_THREADS) {
    const float4 p = gP[v];
    int c = 0;
    for (int a = 0; a < na; a++) {
      const int k = sAct[a];
      if (inside(p, k)) { out[c] = make_int2(ci * nv + v, sCand[k]); c++; }
    }
    if (c > 0) {
      *vl++ = make_uint2((unsigned)v, ((unsigned)off << 8) | (unsigned)min(c, 255));
      out += c;
      off += c;
    }
  }
}

// ---------------------------------------------------------------------------------
// K1b: evaluates every contact unit of the global list, UNIT_LANES lanes per unit, grid-stride (load balanced over
// the whole GPU instead of per cell).  Writes the reference's "winding number" of the unit.
// ---------------------------------------------------------------------------------
constexpr int CONTACT_THREADS = 256;

#ifndef DPM_CONTACT_MINB
#define DPM_CONTACT_MINB 4  // CTAs per SM promised to ptxas for the contact kernel: 64 registers (small spills), 32 instead of 24 warps per SM:
                            // -13 % in the contact-dominated phases against 80 registers, 48 registers (5 CTAs) is slower again
#endif
template <bool ATT>
__global__ void __launch_bounds__(CONTACT_THREADS, DPM_CONTACT_MINB) dpm3d_contact_kernel(Step3DParams P) {
  const int lane = threadIdx.x & 31;
  const int g = lane & (UNIT_LANES - 1);
  const int gshift = lane & ~(UNIT_LANES - 1);
  const unsigned gmask = ((1u << UNIT_LANES) - 1u) << gshift;
  const int ngroups = gridDim.x * (CONTACT_THREADS / UNIT_LANES);
  griddep_launch();  // the step kernel may start its shape-force pass next to this kernel
  griddep_wait();    // the unit list of the units kernel
  const int total = min(P.st->unit_total, P.unit_cap);
  const int nv = P.nv;
  // warp-uniform trip count: the 4 groups of a warp take 4 consecutive units and stay in lockstep
  const int u0 = blockIdx.x * (CONTACT_THREADS / UNIT_LANES) + (threadIdx.x / 32) * (32 / UNIT_LANES);
  for (int ub = u0; ub < total; ub += ngroups) {
    const int u = ub + (lane / UNIT_LANES);
    const bool active = u < total;
    int2 rec = make_int2(0, 0);
    if (active) rec = P.unit_rec[u];
    const int ci = rec.x / nv, cj = rec.y;
    const float4 p = P.pos_in[rec.x];
    const float4 bi2 = P.bnd_in[BND * (size_t)ci + 2];
    const float4 bj1 = P.bnd_in[BND * (size_t)cj + 1], bj2 = P.bnd_in[BND * (size_t)cj + 2], bj3 = P.bnd_in[BND * (size_t)cj + 3];
    const float4 bj4 = P.bnd_in[BND * (size_t)cj + 4];  // kernel point of cj: sphere centre, apex of its star-shape test
    constexpr bool att = ATT;
    float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (P.pbc) {  // shift = L * round((COMi - COMJ) / L)   (:277-281)
      sh.x = P.L * roundf((bi2.x - bj2.x) / P.L);
      sh.y = P.L * roundf((bi2.y - bj2.y) / P.L);
      sh.z = P.L * roundf((bi2.z - bj2.z) / P.L);
    }
    const float4 *Vj = P.pos_in + (size_t)cj * nv;
    const bool star = bj3.y != 0.0f;  // neighbour star-shaped about its kernel point (checked by its owner's epilogue)
    // With the attraction on, the units kernel admits vertices within the (larger) attraction pad: the contact term is
    // evaluated only for those that pass the units kernel's test with the CONTACT pad (same expressions), the others get
    // w = 0 exactly as if they had been culled.
    bool contact = active;
    if (att) {
      contact = active && (P.mask & DPM3D_REPEL) && P.Kc != 0.0f;
      const float4 bj0 = P.bnd_in[BND * (size_t)cj];
      const float pad = bj1.w;
      const float lx = (bj0.x + sh.x) - pad, ly = (bj0.y + sh.y) - pad, lz = (bj0.z + sh.z) - pad;
      const float hx = (bj1.x + sh.x) + pad, hy = (bj1.y + sh.y) + pad, hz = (bj1.z + sh.z) + pad;
      const float rs = sqrtf(bj0.w) + pad;
      const float dx = p.x - (bj4.x + sh.x), dy = p.y - (bj4.y + sh.y), dz = p.z - (bj4.z + sh.z);
      contact = contact && !(p.x < lx || p.x > hx || p.y < ly || p.y > hy || p.z < lz || p.z > hz) &&
                (dx * dx + dy * dy + dz * dz) <= rs * rs * 1.0001f + 1e-30f;
    }
    float w = 0.0f;
    int why = winding_fast(P, contact && star, Vj, sh, p, f3(bj4.x + sh.x, bj4.y + sh.y, bj4.z + sh.z), bj1.w, w, g, gshift);
    const bool nonstar = contact && !star;
    if (__any_sync(0xffffffffu, nonstar)) {  // warp-uniform: the general evaluation over the neighbours' patch boxes
      const float4 bj0 = P.bnd_in[BND * (size_t)cj];
      float wp = 0.0f;
      const bool coincident = winding_patches(P.faces, P.nf, P.npatch, nonstar, Vj, P.patch_box + (size_t)cj * P.npatch * 2, sh, p,
                                              bj4.z + sh.z, bj0, bj1, bj1.w, wp, g, gshift);
      if (nonstar) {
        w = wp;
        why = coincident ? 1 : 0;
        if (g == 0) atomicAdd(&P.st->fallback_why[0], 1ull);  // statistics: units of non-star-shaped neighbours
      }
    }
    if (contact && why != 0) {  // group-uniform branch
      w = winding_literal(Vj, P.faces, P.nf, sh, p, g, gmask);
      if (g == 0) { atomicAdd(&P.st->literal_evals, 1ull); atomicAdd(&P.st->fallback_why[why], 1ull); }
    }
    if (active && g == 0) P.unit_w[u] = contact ? w : 0.0f;
    if (att) {
      // AllVertAttraction (shaders/Cell3D_Kernel.cl:313-364) in gather form: what this vertex's work-item adds to itself
      // (rest length l0[ci]) plus what the work-item of every vertex vj of cj scatters onto it (rest length l0[cj]; its
      // delta is exactly -delta and its distance exactly dist).  The group's 8 lanes take different vj.
      const float l0i = fmaxf(P.bnd_in[BND * (size_t)ci + 3].w, 1e-12f), l0j = fmaxf(bj3.w, 1e-12f);
      float ax = 0.0f, ay = 0.0f, az = 0.0f;
      if (active) {
        for (int vj = g; vj < nv; vj += UNIT_LANES) {
          const float4 q = __ldg(Vj + vj);
          float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
          if (P.pbc) {  // :338-343
            dx -= P.L * roundf(dx / P.L);
            dy -= P.L * roundf(dy / P.L);
            dz -= P.L * roundf(dz / P.L);
          }
          const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
          float s = 0.0f;
          if (dist > 1e-12f) {  // :350
            if (dist < l0i * 2.0f) s += P.Kat * 0.5f * (dist / l0i - 1.0f);
            if (dist < l0j * 2.0f) s += P.Kat * 0.5f * (dist / l0j - 1.0f);
          }
          if (s != 0.0f) {
            const float r = s / dist;
            ax -= r * dx; ay -= r * dy; az -= r * dz;
          }
        }
      }
#pragma unroll
      for (int o = UNIT_LANES / 2; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
        az += __shfl_xor_sync(0xffffffffu, az, o);
      }
      if (active && g == 0) P.unit_att[u] = make_float4(ax, ay, az, 0.0f);
    }
  }
}

// signed volume of the tetrahedron (C, P0, P1, P2) = dot(P0 - C, n)/6, n = (P1-P0) x (P2-P0), is positive with a margin
// (the sine of the angle between P0 - C and the face plane exceeds 1e-3) for every face  <=>  the mesh is star-shaped
// about C (closed, consistently oriented): the precondition of winding_fast
__device__ __forceinline__ bool face_sees_centre(float4 P0, float3 n, float nn, float3 C) {
  const float3 a = sub3(P0, C);
  const float sv = dot3(a, n);
  return sv > 0.0f && sv * sv > 1e-6f * dot3(a, a) * nn;
}

// ---------------------------------------------------------------------------------
// Per-cell scalars of the positions held in shared memory (sP): exact AABB, serial-order COM, serial-order signed
// volume (the four chains run in four lanes of ONE warp, rotated over the CTA's warps by the cell index so that every
// SM sub-partition gets its share of them), r^2 max/min about the COM, longest edge (-> contact pad) and the
// star-shape flag.  Used by the bounds kernel (after an upload) and by the step kernel's epilogue (for the NEW
// positions).  Block of STEP_THREADS threads; sWide: max(nv, ceil(nf/2)) float4 of scratch.  Every thread passes the partial vertex
// sum and partial AABB of the vertices it staged / integrated (tid, tid + STEP_THREADS, ...).
//   bnd[0] = (lo.xyz, r2max)   bnd[1] = (hi.xyz, pad)   bnd[2] = (com.xyz, volume)   bnd[3] = (r2min, star, vol_prev, l0)
// ---------------------------------------------------------------------------------
#ifndef DPM_STEP_THREADS
#define DPM_STEP_THREADS 128
#endif
constexpr int STEP_THREADS = DPM_STEP_THREADS;
constexpr int STEP_WARPS = STEP_THREADS / 32;

struct CellTopo {
  const ushort4 *__restrict__ faces;
  const uint16_t *__restrict__ ring_nbr;
  const uint8_t *__restrict__ valence;
  int ring_stride, nv, nf;
};

struct VertPartial {  // per-thread partials over the thread's own vertices
  float sx, sy, sz;
  float lo[3], hi[3];
  __device__ __forceinline__ void init() {
    sx = sy = sz = 0.0f;
    lo[0] = lo[1] = lo[2] = INFINITY;
    hi[0] = hi[1] = hi[2] = -INFINITY;
  }
  __device__ __forceinline__ void add(float4 p) {
    sx += p.x; sy += p.y; sz += p.z;
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
  }
};

__device__ __forceinline__ void cell_scalars(const float4 *sP, float4 *sWide, unsigned char *sV, const CellTopo &T, VertPartial vp, float vol_prev,
                                             const float *l0_ptr, float4 *bnd_cell, unsigned char *flag_cell, const float4 *bbox_lo,
                                             const float4 *bbox_hi, NbrState *st) {
  __shared__ float sRed[STEP_WARPS][12];
  __shared__ float sSc[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = T.nv, nf = T.nf;
  // approximate centroid (tree sum): apex for the star-shape test only; the stored COM is the serial-order one
  {
    const float sx = warp_sum(vp.sx), sy = warp_sum(vp.sy), sz = warp_sum(vp.sz);
    float lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) { lo[d] = warp_min(vp.lo[d]); hi[d] = warp_max(vp.hi[d]); }
    if (lane == 0) {
      sRed[warp][0] = sx; sRed[warp][1] = sy; sRed[warp][2] = sz;
#pragma unroll
      for (int d = 0; d < 3; d++) { sRed[warp][3 + d] = lo[d]; sRed[warp][6 + d] = hi[d]; }
    }
  }
  __syncthreads();
  float3 capx = f3(0.f, 0.f, 0.f);
#pragma unroll
  for (int w = 0; w < STEP_WARPS; w++) { capx.x += sRed[w][0]; capx.y += sRed[w][1]; capx.z += sRed[w][2]; }
  {
    const float inv = 1.0f / (float)nv;
    capx = f3(capx.x * inv, capx.y * inv, capx.z * inv);
  }
  // ONE pass over the faces: signed-volume term dot(cross(P0,P1),P2)/6.0f in the reference's operation order, unfused
  // (shaders/Cell3D_Kernel.cl:58-61); star-shape test; next step's facing-the-substrate (StickToSurface :209-214) and
  // degenerate-edge (:151) flags
  float *sTerm = reinterpret_cast<float *>(sWide);  // term f at float 4*(f/2) + (f&1): pair k = (2k, 2k+1) in sWide[k].xy
  int star = 1;
  float e2 = 0.0f;
  DPM_UNROLL(DPM_FACE_UNROLL)
  for (int f = tid; f < nf; f += STEP_THREADS) {
    const ushort4 fc = __ldg(T.faces + f);
    const float4 P0 = sP[fc.x], P1 = sP[fc.y], P2 = sP[fc.z];
    const float cx = __fsub_rn(__fmul_rn(P0.y, P1.z), __fmul_rn(P0.z, P1.y));
    const float cy = __fsub_rn(__fmul_rn(P0.z, P1.x), __fmul_rn(P0.x, P1.z));
    const float cz = __fsub_rn(__fmul_rn(P0.x, P1.y), __fmul_rn(P0.y, P1.x));
    sTerm[4 * (f >> 1) + (f & 1)] = div6_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, P2.x), __fmul_rn(cy, P2.y)), __fmul_rn(cz, P2.z)));
    const float3 A = sub3(P1, P0), B = sub3(P2, P0), C = sub3(P2, P1);
    const float3 n = cross3(A, B);
    const float nn = dot3(n, n);
    star &= face_sees_centre(P0, n, nn, capx) ? 1 : 0;
    const bool down = n.z * rsqrt_fast(nn) < -0.1f;
    const float la = dot3(A, A), lb = dot3(B, B), lc = dot3(C, C);
    e2 = fmaxf(e2, fmaxf(la, fmaxf(lb, lc)));  // every edge is an edge of some face: longest edge for free
    const bool deg = fminf(la, fminf(lb, lc)) < 1e-24f;
    if (down) {  // only corners the substrate force can act on (below the plane or within 2 l0 of it)
      const float near_z = *l0_ptr * 2.0f;
      if (P0.z < near_z) vflag_add(sV, fc.x, 1u);
      if (P1.z < near_z) vflag_add(sV, fc.y, 1u);
      if (P2.z < near_z) vflag_add(sV, fc.z, 1u);
    }
    if (deg) { vflag_or(sV, fc.x, 0x80u); vflag_or(sV, fc.y, 0x80u); vflag_or(sV, fc.z, 0x80u); }
  }
  e2 = warp_max(e2);
  if (lane == 0) sRed[warp][9] = e2;
  __syncthreads();
  for (int v = tid; v < vflag_stride(nv); v += STEP_THREADS) flag_cell[v] = sV[v];
  // The serial chains (:35-44 COM, :46-64 volume): lane 0/1/2 = COM x/y/z, lane 3 = signed volume.  Every iteration
  // is one 8-byte LDS with the same 16-byte stride in all four lanes (lanes 0,1: sP[k].xy; lane 2: sP[k].zw; lane 3:
  // sWide[k].xy = terms 2k, 2k+1) and two predicated FADDs.
  if (warp == (int)(blockIdx.x & (STEP_WARPS - 1)) && lane < 4) {
    unsigned addr = (lane == 3) ? smem_addr(sWide) : smem_addr(sP) + (lane == 2 ? 8u : 0u);
    const bool useA = lane != 1, useB = (lane & 1) != 0;
    const int cnt = (lane == 3) ? (nf >> 1) : nv;
    const int ncommon = min(nv, nf >> 1);
    float s = 0.0f;
    int k = 0;
    for (; k + 8 <= ncommon; k += 8, addr += 128u) {
      float2 t[8];
      t[0] = lds_v2<0>(addr); t[1] = lds_v2<16>(addr); t[2] = lds_v2<32>(addr); t[3] = lds_v2<48>(addr);
      t[4] = lds_v2<64>(addr); t[5] = lds_v2<80>(addr); t[6] = lds_v2<96>(addr); t[7] = lds_v2<112>(addr);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        if (useA) s = __fadd_rn(s, t[q].x);
        if (useB) s = __fadd_rn(s, t[q].y);
      }
    }
    for (; k < cnt; k++, addr += 16u) {
      const float2 t = lds_v2<0>(addr);
      if (useA) s = __fadd_rn(s, t.x);
      if (useB) s = __fadd_rn(s, t.y);
    }
    if (lane == 3 && (nf & 1)) s = __fadd_rn(s, sTerm[4 * (nf >> 1)]);
    sSc[lane] = (lane == 3) ? fabsf(s) : __fmul_rn(s, __fdiv_rn(1.0f, (float)nv));
  }
  __syncthreads();
  const float3 com = f3(sSc[0], sSc[1], sSc[2]);
  float r2 = 0.0f, r2min = INFINITY;
  for (int v = tid; v < nv; v += STEP_THREADS) { const float3 q = sub3(sP[v], com); const float qq = dot3(q, q); r2 = fmaxf(r2, qq); r2min = fminf(r2min, qq); }
  r2 = warp_max(r2);
  r2min = warp_min(r2min);
  if (lane == 0) { sRed[warp][10] = r2; sRed[warp][11] = r2min; }
  star = __syncthreads_and(star);
  if (tid == 0) {
    float l[3], h[3], em = 0.f, rr = 0.f, rm = INFINITY;
    for (int d = 0; d < 3; d++) { l[d] = INFINITY; h[d] = -INFINITY; }
    for (int w = 0; w < STEP_WARPS; w++) {
      for (int d = 0; d < 3; d++) { l[d] = fminf(l[d], sRed[w][3 + d]); h[d] = fmaxf(h[d], sRed[w][6 + d]); }
      em = fmaxf(em, sRed[w][9]); rr = fmaxf(rr, sRed[w][10]); rm = fminf(rm, sRed[w][11]);
    }
    const float pad = CONTACT_PAD * sqrtf(em);
    bnd_cell[0] = make_float4(l[0], l[1], l[2], rr);
    bnd_cell[1] = make_float4(h[0], h[1], h[2], pad);
    bnd_cell[2] = make_float4(com.x, com.y, com.z, sSc[3]);
    bnd_cell[3] = make_float4(rm, star ? 1.f : 0.f, vol_prev, *l0_ptr);  // l0 travels with the bounds (ghost cells have no parameters)
    bnd_cell[4] = make_float4(com.x, com.y, com.z, 0.f);                 // kernel point of freshly uploaded positions: their own COM
    if (st) {  // neighbour-list validity
      const float4 bl = *bbox_lo, bh = *bbox_hi;
      if (l[0] < bl.x || l[1] < bl.y || l[2] < bl.z || h[0] > bh.x || h[1] > bh.y || h[2] > bh.z) st->rebuild = 1;
      if (pad > st->range) st->rebuild = 1;  // the candidate lists were built for smaller contact pads
    }
  }
}

static __global__ void __launch_bounds__(STEP_THREADS) dpm3d_bounds_kernel(const float4 *pos, float4 *bnd, unsigned char *flags, int nc,
                                                                           CellTopo T, const float4 *cellB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *sP = reinterpret_cast<float4 *>(smem_raw);
  float4 *sWide = sP + T.nv;  // same carve-up as the step kernel (sP | sF | vertex flags)
  unsigned char *sV = reinterpret_cast<unsigned char *>(sWide + step3d_mid_slots(T.nv, T.nf));
  const int ci = blockIdx.x;
  for (int v = threadIdx.x; v < vflag_stride(T.nv) / 4; v += STEP_THREADS) reinterpret_cast<unsigned *>(sV)[v] = 0u;
  VertPartial vp;
  vp.init();
  for (int v = threadIdx.x; v < T.nv; v += STEP_THREADS) {
    const float4 p = pos[(size_t)ci * T.nv + v];
    sP[v] = p;
    vp.add(p);
  }
  __syncthreads();
  cell_scalars(sP, sWide, sV, T, vp, 0.0f, &cellB[ci].y, bnd + BND * (size_t)ci, flags + (size_t)ci * vflag_stride(T.nv), nullptr, nullptr, nullptr);
}

// Walk-start table of the fast contact evaluation: for the direction of each octahedral texel, the face of cell 0 whose
// spherical triangle (seen from the vertex mean) contains it best.  A hint only: the walk of winding_fast ends on the pierced
// face from any start.  One CTA per texel, faces strided over the threads.
static __global__ void __launch_bounds__(STEP_THREADS) dpm3d_dirtable_kernel(const float4 *pos, const ushort4 *faces, int nv, int nf,
                                                                             unsigned short *tab) {
  __shared__ float sC[3][STEP_THREADS / 32];
  __shared__ float sBestV[STEP_THREADS / 32];
  __shared__ int sBestF[STEP_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (int v = tid; v < nv; v += STEP_THREADS) { const float4 q = pos[v]; cx += q.x; cy += q.y; cz += q.z; }
  cx = warp_sum(cx); cy = warp_sum(cy); cz = warp_sum(cz);
  if (lane == 0) { sC[0][wid] = cx; sC[1][wid] = cy; sC[2][wid] = cz; }
  __syncthreads();
  cx = cy = cz = 0.f;
  for (int w = 0; w < STEP_THREADS / 32; w++) { cx += sC[0][w]; cy += sC[1][w]; cz += sC[2][w]; }
  const float inv = 1.0f / (float)nv;
  const float3 c = f3(cx * inv, cy * inv, cz * inv);
  const int ix = blockIdx.x % DIR_N, iy = blockIdx.x / DIR_N;
  float x = (ix + 0.5f) / DIR_N * 2.0f - 1.0f, y = (iy + 0.5f) / DIR_N * 2.0f - 1.0f;
  const float z = 1.0f - fabsf(x) - fabsf(y);
  if (z < 0.0f) {
    const float tx = (1.0f - fabsf(y)) * (x >= 0.0f ? 1.0f : -1.0f), ty = (1.0f - fabsf(x)) * (y >= 0.0f ? 1.0f : -1.0f);
    x = tx; y = ty;
  }
  const float3 u = f3(x, y, z);
  float bestv = -3.0e38f;
  int best = 0;
  for (int f = tid; f < nf; f += STEP_THREADS) {
    const ushort4 fc = faces[f];
    const float4 q0 = pos[fc.x], q1 = pos[fc.y], q2 = pos[fc.z];
    const float3 A = f3(q0.x - c.x, q0.y - c.y, q0.z - c.z), B = f3(q1.x - c.x, q1.y - c.y, q1.z - c.z),
                 C = f3(q2.x - c.x, q2.y - c.y, q2.z - c.z);
    const float3 n0 = cross3(A, B), n1 = cross3(B, C), n2 = cross3(C, A);
    const float m = fminf(dot3(u, n0) * rsqrtf(dot3(n0, n0) + 1e-30f),
                          fminf(dot3(u, n1) * rsqrtf(dot3(n1, n1) + 1e-30f), dot3(u, n2) * rsqrtf(dot3(n2, n2) + 1e-30f)));
    if (m > bestv) { bestv = m; best = f; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
    const int of = __shfl_xor_sync(0xffffffffu, best, o);
    if (ov > bestv || (ov == bestv && of < best)) { bestv = ov; best = of; }
  }
  if (lane == 0) { sBestV[wid] = bestv; sBestF[wid] = best; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < STEP_THREADS / 32; w++)
      if (sBestV[w] > bestv || (sBestV[w] == bestv && sBestF[w] < best)) { bestv = sBestV[w]; best = sBestF[w]; }
    tab[octa_texel(u.x, u.y, u.z)] = (unsigned short)best;
  }
}

// ---------------------------------------------------------------------------------
// K2: the fused shape-force + integrate kernel.  One CTA of 128 threads per cell, vertices strided over the threads,
// positions and force accumulators in shared memory (small register footprint -> 8 CTAs per SM, which is what
// hides the serial chains of the epilogue).  The cell's positions and face flags arrive by two bulk async copies
// (one instruction each), the new positions leave by one.
// ---------------------------------------------------------------------------------

// Ring gather of one vertex: edge springs (SurfaceAreaForceUpdate :114-178) and the volume gradient (VolumeForceUpdate
// :66-112).  DEG: some ring face of this warp's vertices has a degenerate edge (rare) -> per-edge weights; otherwise
// every edge counts twice.  COMPAT: reference-race mode, the gradient is split by face index and taken about the COM
// like the reference; otherwise it is taken about the vertex itself (same sum over a closed ring, smaller terms).
__device__ __forceinline__ float4 lds_off(const float4 *sP, unsigned off) {
  return *reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(sP) + off);
}
// Degenerate ring faces of one vertex, from the geometry (rare path: some face at this warp's vertices has an edge shorter
// than 1e-12, SurfaceAreaForceUpdate :151): bit i <=> ring face i = (v, n_i, n_{i+1}) has such an edge.  The three squared
// lengths are the same fp32 values the face pass compared, so both sides agree on which faces are skipped.
template <int MAXV, int MINV>
__device__ __forceinline__ unsigned ring_deg_mask(const float4 *sP, const unsigned short (&ro)[MAXV], int val, float4 Pv) {
  unsigned m = 0u;
  float3 E0 = f3(0.f, 0.f, 0.f), Ep = E0;
  float l0 = 0.f, lp = 0.f;
  int last = 0;
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    if (i >= MINV && i >= val) break;
    const float3 E = sub3(lds_off(sP, ro[i]), Pv);
    const float l2 = dot3(E, E);
    if (i == 0) { E0 = E; l0 = l2; }
    else {
      const float3 D = f3(E.x - Ep.x, E.y - Ep.y, E.z - Ep.z);
      if (fminf(lp, fminf(l2, dot3(D, D))) < 1e-24f) m |= 1u << (i - 1);
    }
    Ep = E; lp = l2; last = i;
  }
  const float3 D = f3(E0.x - Ep.x, E0.y - Ep.y, E0.z - Ep.z);
  if (fminf(lp, fminf(l0, dot3(D, D))) < 1e-24f) m |= 1u << last;
  return m;
}
template <int MAXV, int MINV, bool DEG, bool COMPAT>
__device__ __forceinline__ void ring_gather(const float4 *sP, const unsigned short (&ro)[MAXV], const unsigned short (&rf)[MAXV], int val,
                                            unsigned m, float4 Pv, float3 com, float inv_l0, int stale_from, float3 &T, float3 &g,
                                            float3 &gs) {
  float3 Qp = f3(0.f, 0.f, 0.f), Q0 = f3(0.f, 0.f, 0.f);
  int rf_last = 0;  // ring face val-1 (tracked so that no local array is indexed dynamically)
#pragma unroll
  for (int i = 0; i < MAXV; i++) {
    if (i >= MINV && i >= val) break;
    if (COMPAT) rf_last = rf[i];
    const float4 Pn = lds_off(sP, ro[i]);
    const float3 E = sub3(Pn, Pv);
    const float len2 = dot3(E, E);
    const float rl = rsqrt_fast(len2);
    const float dl = len2 * rl * inv_l0 - 1.0f;  // len/l0 - 1  (:156-160)
    float sc = rl * dl;
    if (DEG) {
      // edge (v, n_i) belongs to ring faces i-1 and i; each contributes unit(E)*dl unless degenerate
      const int ip = (i == 0) ? val - 1 : i - 1;
      const int w = 2 - (int)((m >> i) & 1u) - (int)((m >> ip) & 1u);
      sc = (w == 0) ? 0.0f : sc * (float)w;  // an edge of zero length has no direction: both its faces are skipped
    }
    T.x += E.x * sc; T.y += E.y * sc; T.z += E.z * sc;
    const float3 Q = COMPAT ? sub3(Pn, com) : E;
    if (i == 0) Q0 = Q;
    else if (COMPAT) {
      const float3 c = cross3(Qp, Q);  // gradient of ring face i-1 = (v, n_{i-1}, n_i)
      g.x += c.x; g.y += c.y; g.z += c.z;
      if ((int)rf[i - 1] >= stale_from) { gs.x += c.x; gs.y += c.y; gs.z += c.z; }
    } else {  // g += Qp x Q as two chained FMAs per component
      g.x = fmaf(Qp.y, Q.z, g.x); g.x = fmaf(-Qp.z, Q.y, g.x);
      g.y = fmaf(Qp.z, Q.x, g.y); g.y = fmaf(-Qp.x, Q.z, g.y);
      g.z = fmaf(Qp.x, Q.y, g.z); g.z = fmaf(-Qp.y, Q.x, g.z);
    }
    Qp = Q;
  }
  if (COMPAT) {
    const float3 c = cross3(Qp, Q0);  // ring face val-1 = (v, n_{val-1}, n_0)
    g.x += c.x; g.y += c.y; g.z += c.z;
    if (rf_last >= stale_from) { gs.x += c.x; gs.y += c.y; gs.z += c.z; }
  } else {
    g.x = fmaf(Qp.y, Q0.z, g.x); g.x = fmaf(-Qp.z, Q0.y, g.x);
    g.y = fmaf(Qp.z, Q0.x, g.y); g.y = fmaf(-Qp.x, Q0.z, g.y);
    g.z = fmaf(Qp.x, Q0.y, g.z); g.z = fmaf(-Qp.y, Q0.x, g.z);
  }
  if (!DEG) { T.x *= 2.0f; T.y *= 2.0f; T.z *= 2.0f; }
}

// MAXV: ring slots read per vertex (6: valence 5..6, the icospheres; 8; 16 = two 16-byte loads); MINV: ring slots
// known to be occupied for every vertex (no bound check)
// ATT: DPM3D_ATTRACT selected with Kat != 0 (the default instantiation carries none of the attraction code)
template <int MAXV, int MINV, bool COMPAT, bool ATT>
__global__ void __launch_bounds__(STEP_THREADS, DPM_STEP_MINB) dpm3d_step_kernel(Step3DParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long sBar;
  const int nv = P.nv, nf = P.nf;
  float4 *sP = reinterpret_cast<float4 *>(smem_raw);
  float4 *sF = sP + nv;
  const int nvp = vflag_stride(nv);
  unsigned char *sVin = reinterpret_cast<unsigned char *>(sF + step3d_mid_slots(nv, nf));  // per-vertex flags of the current positions
  unsigned char *sVout = sVin + nvp;                                                       // ... of the new positions (face pass below)
  const int tid = threadIdx.x;
  const int ci = blockIdx.x;
  const float4 *gP = P.pos_in + (size_t)ci * nv;

  // ---- stage the vertex ring and the per-vertex flags of the current positions (computed by the previous epilogue) ----
  __shared__ int2 sPush;  // sharded runs: this cell's slots in the neighbouring ranks' inboxes (thread 0's own scratch: read back by thread 0 only)
  int2 ps0 = make_int2(-1, -1);
  if (tid == 0) {
    mbar_init(&sBar, 1);
    const unsigned pb = (unsigned)(sizeof(float4) * nv);
    mbar_expect_tx(&sBar, pb + (unsigned)nvp);
    bulk_g2s(sP, gP, pb, &sBar);
    bulk_g2s(sVin, P.flag_in + (size_t)ci * nvp, (unsigned)nvp, &sBar);
    if (P.push_slot) ps0 = P.push_slot[ci];  // requested now, parked in shared memory once the staging wait below is over
  }
  for (int v = tid; v < nvp / 4; v += STEP_THREADS) reinterpret_cast<unsigned *>(sVout)[v] = 0u;
  const float4 cA = P.cellA[ci], cB = P.cellB[ci];
  const float Kv = cA.x, Ka = cA.y, Ks = cA.z, v0 = cA.w, a0 = cB.x, l0 = cB.y;
  const float4 bi2 = P.bnd_in[BND * (size_t)ci + 2], bi3 = P.bnd_in[BND * (size_t)ci + 3];
  const float3 com = f3(bi2.x, bi2.y, bi2.z);
  const bool doVol = (P.mask & DPM3D_VOLUME) && (Kv != 0.0f);
  // volume of the CURRENT positions was left in the bounds by the previous epilogue (serial-order chain)
  const float coef = doVol ? (-Kv * (bi2.w / v0 - 1.0f)) * (1.0f / 6.0f) : 0.0f;  // :85,:106-108
  // compat mode: faces >= stale_from see the volume of the previous step's start (0 right after an upload)
  const float dcoef = (COMPAT && doVol) ? (-Kv * (bi3.z / v0 - 1.0f)) * (1.0f / 6.0f) - coef : 0.0f;
  const bool doArea = (P.mask & DPM3D_AREA) && !(Ka < 1e-8f);
  const bool doStick = (P.mask & DPM3D_STICK) && !(Ks < 1e-12f);
  const float inv_l0 = 1.0f / l0;
  const float scale = doArea ? Ka * sqrtf(a0) / l0 * 0.3f : 0.0f;  // :162
  const bool doRep = (P.mask & DPM3D_REPEL) && P.Kc != 0.0f;
  __syncthreads();  // the barrier is initialised
  mbar_wait(&sBar, 0);
  if (tid == 0) sPush = ps0;

  // ---- ring pass: every vertex gathers over its constant ring adjacency ---------------------------------
  DPM_UNROLL(DPM_RING_UNROLL)
  for (int r = tid; r < nv; r += STEP_THREADS) {
    const int v = (int)__ldg(P.vorder + r);  // the order the host chose: the 8 lanes of a quarter-warp gather from 8 different bank groups
    const int val = (MINV == MAXV) ? MAXV : (int)__ldg(P.valence + v);
    // ring tables: 8 x uint16 per vertex = one 16-byte load each (valence <= 8; the stride-16 layout takes two); the
    // neighbour table holds byte offsets into sP; the face table is only needed by the reference-race mode
    unsigned short ro[MAXV], rf[MAXV];
    {
      const uint4 a = __ldg(reinterpret_cast<const uint4 *>(P.ring_nbr + (size_t)v * P.ring_stride));
      ro[0] = a.x & 0xffff; ro[1] = a.x >> 16; ro[2] = a.y & 0xffff; ro[3] = a.y >> 16; ro[4] = a.z & 0xffff; ro[5] = a.z >> 16;
      if (MAXV > 6) { ro[6] = a.w & 0xffff; ro[7] = a.w >> 16; }
      if (MAXV > 8) {
        const uint4 c = __ldg(reinterpret_cast<const uint4 *>(P.ring_nbr + (size_t)v * P.ring_stride) + 1);
        ro[MAXV - 8] = c.x & 0xffff; ro[MAXV - 7] = c.x >> 16; ro[MAXV - 6] = c.y & 0xffff; ro[MAXV - 5] = c.y >> 16;
        ro[MAXV - 4] = c.z & 0xffff; ro[MAXV - 3] = c.z >> 16; ro[MAXV - 2] = c.w & 0xffff; ro[MAXV - 1] = c.w >> 16;
      }
#pragma unroll
      for (int i = 0; i < MAXV; i++) rf[i] = 0;
      if (COMPAT) {
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(P.ring_face + (size_t)v * P.ring_stride));
        rf[0] = b.x & 0xffff; rf[1] = b.x >> 16; rf[2] = b.y & 0xffff; rf[3] = b.y >> 16; rf[4] = b.z & 0xffff; rf[5] = b.z >> 16;
        if (MAXV > 6) { rf[6] = b.w & 0xffff; rf[7] = b.w >> 16; }
        if (MAXV > 8) {
          const uint4 d = __ldg(reinterpret_cast<const uint4 *>(P.ring_face + (size_t)v * P.ring_stride) + 1);
          rf[MAXV - 8] = d.x & 0xffff; rf[MAXV - 7] = d.x >> 16; rf[MAXV - 6] = d.y & 0xffff; rf[MAXV - 5] = d.y >> 16;
          rf[MAXV - 4] = d.z & 0xffff; rf[MAXV - 3] = d.z >> 16; rf[MAXV - 2] = d.w & 0xffff; rf[MAXV - 1] = d.w >> 16;
        }
      }
    }
    const unsigned vf = sVin[v];
    const int ndown = (int)(vf & 0x7fu);
    const float4 Pv = sP[v];
    float3 T = f3(0.f, 0.f, 0.f), g = f3(0.f, 0.f, 0.f), gs = f3(0.f, 0.f, 0.f);
    if (__any_sync(__activemask(), (vf & 0x80u) != 0u)) {
      const unsigned m = ring_deg_mask<MAXV, MINV>(sP, ro, val, Pv);
      ring_gather<MAXV, MINV, true, COMPAT>(sP, ro, rf, val, m, Pv, com, inv_l0, P.stale_from, T, g, gs);
    } else
      ring_gather<MAXV, MINV, false, COMPAT>(sP, ro, rf, val, 0u, Pv, com, inv_l0, P.stale_from, T, g, gs);
    float3 F = f3(T.x * scale + coef * g.x, T.y * scale + coef * g.y, T.z * scale + coef * g.z);
    if (COMPAT) { F.x += dcoef * gs.x; F.y += dcoef * gs.y; F.z += dcoef * gs.z; }
    if (doStick && ndown > 0) {
      const float nd = (float)ndown;  // one application per adjacent down-facing face (:226-246)
      const float h = fabsf(Pv.z);
      if (Pv.z < 0.0f) F.z += nd * (Ks * h);
      if (h < l0 * 2.0f) {
        const float3 ctv = f3(Pv.x - com.x, Pv.y - com.y, 0.0f - com.z);
        const float ftmp = Ks * (1.0f - h * inv_l0);
        const float sc = rsqrt_fast(dot3(ctv, ctv)) * ftmp * nd;
        F.x += ctv.x * sc; F.y += ctv.y * sc; F.z += ctv.z * sc;
      }
    }
    sF[v] = make_float4(F.x, F.y, F.z, 0.f);
  }
  // ---- Euler update (EulerPosition :380), outputs ----------------------------------------------------------
  // The contact weights are the only input from this timestep's units / contact kernels: everything above overlaps them.
  // Their per-cell counts are requested before the barrier that ends the ring pass, so the load latency hides behind it.
  griddep_wait();
  constexpr bool doAtt = ATT;
  const int ucnt = (doRep || doAtt) ? P.unit_cnt[ci] : 0;
  const int ubase = (doRep || doAtt) ? P.unit_base[ci] : 0;
  __syncthreads();  // everyone is done reading start-of-step sP
  const float *uw = P.unit_w + ubase;
  const float4 *ua = P.unit_att + ubase;
  // fold the evaluated contact units into the forces of the few vertices that have any (the units kernel's compact list;
  // a vertex's units are ordered by ascending neighbour id: the reference's cj order), before the Euler loop
  if (ucnt > 0) {  // uniform over the CTA
    const int nvl = P.vlist_cnt[ci];
    const uint2 *vl = P.vlist + (size_t)ci * nv;
    for (int i = tid; i < nvl; i += STEP_THREADS) {
      const uint2 e = vl[i];
      const int v = (int)e.x, un = (int)(e.y & 0xffu), uo = (int)(e.y >> 8);
      float4 F = sF[v];
      const float4 np = sP[v];
      float3 dir = f3(0.f, 0.f, 0.f);
      bool have = false;
      if (doAtt)  // AllVertAttraction (:313-364): the units' gathered vertex-vertex terms
        DPM_UNROLL(1)
        for (int u = uo; u < uo + un; u++) { const float4 a = ua[u]; F.x += a.x; F.y += a.y; F.z += a.z; }
      DPM_UNROLL(1)
      for (int u = uo; u < uo + un; u++) {
        const float wn = uw[u];
        if (doRep && !(fabsf(wn) < 1e-6f)) {  // :302-308
          if (!have) {
            const float3 d = f3(com.x - np.x, com.y - np.y, com.z - np.z);
            const float r = rsqrtf(dot3(d, d));
            dir = f3(d.x * r, d.y * r, d.z * r);
            have = true;
          }
          const float mg = fabsf(wn) * 0.5f * P.Kc;
          F.x += mg * dir.x; F.y += mg * dir.y; F.z += mg * dir.z;
        }
      }
      sF[v] = F;
    }
    __syncthreads();
  }
  // The kernel point of the NEW positions: the serial-order COM of the current ones (known since the prologue).
  const float3 kp = com;
  float4 *part = P.part + ((size_t)ci * (STEP_THREADS / 32) + (tid >> 5)) * 3;  // this warp's partial record
  {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float r2 = 0.0f;
    for (int v = tid; v < nv; v += STEP_THREADS) {
      const float4 F = sF[v];
      float4 np = sP[v];
      np.x += F.x * P.dt; np.y += F.y * P.dt; np.z += F.z * P.dt;
      np.w = 0.f;
      if (P.force_out) P.force_out[(size_t)ci * nv + v] = F;
      sP[v] = np;
      lo[0] = fminf(lo[0], np.x); lo[1] = fminf(lo[1], np.y); lo[2] = fminf(lo[2], np.z);
      hi[0] = fmaxf(hi[0], np.x); hi[1] = fmaxf(hi[1], np.y); hi[2] = fmaxf(hi[2], np.z);
      const float3 q = sub3(np, kp);
      r2 = fmaxf(r2, dot3(q, q));
    }
    // the warp's AABB and bounding radius leave now (the group's chain warp folds the partials of the cell's four warps):
    // no register carries them through the face pass, and no thread of this CTA has to wait for a block-wide reduction
#pragma unroll
    for (int d = 0; d < 3; d++) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
    r2 = warp_max(r2);
    if ((tid & 31) == 0) {
      part[0] = make_float4(lo[0], lo[1], lo[2], r2);
      part[1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
  }
  fence_async_smem();  // this thread's sP writes -> visible to the bulk store issued below
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(P.pos_out + (size_t)ci * nv, sP, (unsigned)(sizeof(float4) * nv));
    if (P.push_slot) {  // sharded: a boundary cell also goes straight into the neighbouring ranks' inboxes (peer memory)
      const int2 ps = sPush;
      if (ps.x >= 0) bulk_s2g(P.push_pos[0] + (size_t)ps.x * nv, sP, (unsigned)(sizeof(float4) * nv));
      if (ps.y >= 0) bulk_s2g(P.push_pos[1] + (size_t)ps.y * nv, sP, (unsigned)(sizeof(float4) * nv));
    }
  }
  griddep_launch();  // the next timestep's rebuild kernel may be scheduled; it waits for this grid's completion before it reads anything

  // ---- next step's per-cell scalars from the NEW positions --------------------------------------------------
  // ONE pass over the faces: signed-volume term dot(cross(P0,P1),P2)/6.0f in the reference's operation order, unfused
  // (shaders/Cell3D_Kernel.cl:58-61) -> global memory, for the chain warp of this cell's group; star-shape test about the
  // kernel point; next step's facing-the-substrate (StickToSurface :209-214) and degenerate-edge (:151) flags; longest edge.
  {
    float *gTerm = P.terms + (size_t)ci * terms_stride(nf);
    const float near_z = l0 * 2.0f;
    int star = 1;
    float e2 = 0.0f;
    DPM_UNROLL(DPM_FACE_UNROLL)
    for (int r = tid; r < nf; r += STEP_THREADS) {
      const ushort4 fc = __ldg(P.faces_proc + r);  // .w = face id: a warp's 32 faces are a permutation of one aligned block of 32
      const float4 P0 = sP[fc.x], P1 = sP[fc.y], P2 = sP[fc.z];
      const float cx = __fsub_rn(__fmul_rn(P0.y, P1.z), __fmul_rn(P0.z, P1.y));
      const float cy = __fsub_rn(__fmul_rn(P0.z, P1.x), __fmul_rn(P0.x, P1.z));
      const float cz = __fsub_rn(__fmul_rn(P0.x, P1.y), __fmul_rn(P0.y, P1.x));
      gTerm[fc.w] = div6_rn(__fadd_rn(__fadd_rn(__fmul_rn(cx, P2.x), __fmul_rn(cy, P2.y)), __fmul_rn(cz, P2.z)));
      const float3 A = sub3(P1, P0), B = sub3(P2, P0), C = sub3(P2, P1);
      const float3 n = cross3(A, B);
      const float nn = dot3(n, n);
      star &= face_sees_centre(P0, n, nn, kp) ? 1 : 0;
      const bool down = n.z * rsqrt_fast(nn) < -0.1f;
      const float la = dot3(A, A), lb = dot3(B, B), lc = dot3(C, C);
      e2 = fmaxf(e2, fmaxf(la, fmaxf(lb, lc)));  // every edge is an edge of some face: longest edge for free
      if (down) {  // StickToSurface acts on a corner only below the plane or within 2 l0 of it (:226-246): nothing else is counted
        if (P0.z < near_z) vflag_add(sVout, fc.x, 1u);
        if (P1.z < near_z) vflag_add(sVout, fc.y, 1u);
        if (P2.z < near_z) vflag_add(sVout, fc.z, 1u);
      }
      if (fminf(la, fminf(lb, lc)) < 1e-24f) { vflag_or(sVout, fc.x, 0x80u); vflag_or(sVout, fc.y, 0x80u); vflag_or(sVout, fc.z, 0x80u); }
    }
    e2 = warp_max(e2);
    star = __all_sync(0xffffffffu, star);
    if ((tid & 31) == 0) part[2] = make_float4(e2, star ? 1.f : 0.f, 0.f, 0.f);
    fence_async_smem();  // this thread's flag atomics -> visible to the bulk store issued below
  }
  __syncthreads();  // CTA-scope ordering of every thread's global stores (terms, partials) before thread 0's device-wide fence
  if (tid >= 32) return;  // warps 1..3 are done; warp 0 publishes the cell
  int last = 0;
  if (tid == 0) {
    bulk_s2g(P.flag_out + (size_t)ci * nvp, sVout, (unsigned)nvp);
    bulk_wait_all();   // the new positions and vertex flags have left shared memory and landed
    asm volatile("fence.proxy.async;" ::: "memory");  // async-proxy writes (the bulk store) ordered before the release below
    // (the bulk stores into peer memory, if this cell had any, have completed too.  No system-scope fence here: measured, a
    // fence.sys by 1.5 % of the CTAs slowed the whole kernel by 2 %; the stores are ordered before the arrival flag by the end
    // of this grid and the fence.sys of the push kernel's flag thread, three launches later on the same stream)
    __threadfence();
    // publish this cell; the CTA that completes its group evaluates the group's serial chains and assembles its bounds
    const int grp0 = ci / CHAIN_GROUP;
    const int prev = atomicAdd(P.grp_done + grp0, 1);
    last = (prev == min(CHAIN_GROUP, P.nc - grp0 * CHAIN_GROUP) - 1) ? 1 : 0;
  }
  if (!__shfl_sync(0xffffffffu, last, 0)) return;

  // ---- serial-order COM (:35-44) and signed volume (:46-64) of the group's cells: ONE warp, lane = (cell, chain) ----
  // chains 0/1/2 = COM x/y/z (vertex k: pos[k]), chain 3 = volume (terms 2k, 2k+1).  Every iteration is one 8-byte LDS per
  // lane and two predicated FADDs, as in the single-cell version, but 8 cells share the instruction stream.  The operands
  // are staged from global memory (L2: written by the group's CTAs just now) by cp.async, CHAIN_STAGES chunks in flight.
  __threadfence();  // acquire: the other CTAs' positions / terms / bounds
  asm volatile("fence.proxy.async;" ::: "memory");
  {
    const int grp = ci / CHAIN_GROUP;
    const int gcount = min(CHAIN_GROUP, P.nc - grp * CHAIN_GROUP);
    if (tid == 0) P.grp_done[grp] = 0;  // ready for the next timestep
    const int lane = tid, cell = lane >> 2, chain = lane & 3;
    const bool live = cell < gcount;
    const int c0 = grp * CHAIN_GROUP;
    const int tstride = terms_stride(nf);
    int2 psc = make_int2(-1, -1);  // fused push: the inbox slots of the lane's cell, requested ahead of the chain loop
    if (P.push_slot && live && chain == 0) psc = P.push_slot[c0 + cell];
    const int niter = max(nv, (nf + 1) >> 1), nchunk = (niter + CHAIN_CH - 1) / CHAIN_CH;
    const int cnt = !live ? 0 : (chain == 3 ? ((nf + 1) >> 1) : nv);
    const unsigned sbase = smem_addr(smem_raw);
    const bool useA = chain != 1, useB = (chain & 1) != 0;
    const unsigned lane_off = (chain == 3) ? (unsigned)(CHAIN_GROUP * CHAIN_POS_STRIDE + cell * CHAIN_TERM_STRIDE)
                                           : (unsigned)(cell * CHAIN_POS_STRIDE + (chain == 2 ? 8 : 0));
    const unsigned lane_step = (chain == 3) ? 8u : 16u;
#if DPM_CHAIN_BULK
    // staging by bulk async copies (TMA unit; they do not occupy the LSU data pipe): lane c copies cell c's chunk of
    // positions, lane 8 + c its chunk of terms; one mbarrier per stage counts the bytes
    __shared__ __align__(8) unsigned long long sBarC[CHAIN_STAGES];
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < CHAIN_STAGES; q++) mbar_init(&sBarC[q], 1);
    }
    __syncwarp();
    auto stage_chunk = [&](int ch) {
      if (ch < nchunk) {
        const int stg = ch % CHAIN_STAGES;
        unsigned char *st = smem_raw + stg * CHAIN_STAGE_BYTES;
        const int k0 = ch * CHAIN_CH;
        const unsigned pbytes = (unsigned)(16 * max(0, min(CHAIN_CH, nv - k0)));
        const unsigned tbytes = (2 * k0 < tstride) ? (unsigned)(8 * CHAIN_CH) : 0u;
        if (lane == 0) mbar_expect_tx(&sBarC[stg], (unsigned)gcount * (pbytes + tbytes));
        __syncwarp();
        if (lane < CHAIN_GROUP) {
          if (lane < gcount && pbytes) bulk_g2s(st + lane * CHAIN_POS_STRIDE, P.pos_out + (size_t)(c0 + lane) * nv + k0, pbytes, &sBarC[stg]);
        } else if (lane < 2 * CHAIN_GROUP) {
          const int c = lane - CHAIN_GROUP;
          if (c < gcount && tbytes)
            bulk_g2s(st + CHAIN_GROUP * CHAIN_POS_STRIDE + c * CHAIN_TERM_STRIDE, P.terms + (size_t)(c0 + c) * tstride + 2 * k0, tbytes, &sBarC[stg]);
        }
      }
    };
#pragma unroll
    for (int c = 0; c < CHAIN_STAGES - 1; c++) stage_chunk(c);
    float sacc = 0.0f;
    for (int ch = 0; ch < nchunk; ch++) {
      stage_chunk(ch + CHAIN_STAGES - 1);
      mbar_wait(&sBarC[ch % CHAIN_STAGES], (unsigned)((ch / CHAIN_STAGES) & 1));
#else
    auto stage_chunk = [&](int ch) {
      if (ch < nchunk) {
        const unsigned st = sbase + (unsigned)((ch % CHAIN_STAGES) * CHAIN_STAGE_BYTES);
        const int k0 = ch * CHAIN_CH;
#pragma unroll
        for (int c = 0; c < CHAIN_GROUP; c++) {  // positions: one float4 per lane and cell
          if (c < gcount && k0 + lane < nv) {
            const float4 *src = P.pos_out + (size_t)(c0 + c) * nv + k0 + lane;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + (unsigned)(c * CHAIN_POS_STRIDE + lane * 16)), "l"(src) : "memory");
          }
        }
        if (2 * k0 < tstride) {
#pragma unroll
          for (int c2 = 0; c2 < CHAIN_GROUP; c2 += 2) {  // terms: 2 * CHAIN_CH floats = 16 float4 per cell, two cells per pass
            const int c = c2 + (lane >> 4), q = lane & 15;
            if (c < gcount) {
              const float *src = P.terms + (size_t)(c0 + c) * tstride + 2 * k0 + 4 * q;
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + (unsigned)(CHAIN_GROUP * CHAIN_POS_STRIDE + c * CHAIN_TERM_STRIDE + q * 16)), "l"(src) : "memory");
            }
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");  // (possibly empty) group: keeps the wait counts uniform
    };
#pragma unroll
    for (int c = 0; c < CHAIN_STAGES - 1; c++) stage_chunk(c);
    float sacc = 0.0f;
    for (int ch = 0; ch < nchunk; ch++) {
      stage_chunk(ch + CHAIN_STAGES - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(CHAIN_STAGES - 1) : "memory");
      __syncwarp();
#endif
      const int k0 = ch * CHAIN_CH;
      unsigned addr = sbase + (unsigned)((ch % CHAIN_STAGES) * CHAIN_STAGE_BYTES) + lane_off;
      if (__all_sync(0xffffffffu, !live || k0 + CHAIN_CH <= cnt)) {
        if (live) {
#pragma unroll
          for (int q = 0; q < CHAIN_CH; q += 8, addr += 8u * lane_step) {
            float2 t[8];
            if (chain == 3) {
              t[0] = lds_v2<0>(addr); t[1] = lds_v2<8>(addr); t[2] = lds_v2<16>(addr); t[3] = lds_v2<24>(addr);
              t[4] = lds_v2<32>(addr); t[5] = lds_v2<40>(addr); t[6] = lds_v2<48>(addr); t[7] = lds_v2<56>(addr);
            } else {
              t[0] = lds_v2<0>(addr); t[1] = lds_v2<16>(addr); t[2] = lds_v2<32>(addr); t[3] = lds_v2<48>(addr);
              t[4] = lds_v2<64>(addr); t[5] = lds_v2<80>(addr); t[6] = lds_v2<96>(addr); t[7] = lds_v2<112>(addr);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
              if (useA) sacc = __fadd_rn(sacc, t[j].x);
              if (useB) sacc = __fadd_rn(sacc, t[j].y);
            }
          }
        }
      } else {
        for (int q = 0; q < CHAIN_CH; q++, addr += lane_step) {
          if (k0 + q < cnt) {
            const float2 t = lds_v2<0>(addr);
            if (useA) sacc = __fadd_rn(sacc, t.x);
            if (useB && !(chain == 3 && 2 * (k0 + q) + 1 >= nf)) sacc = __fadd_rn(sacc, t.y);
          }
        }
      }
      __syncwarp();  // every lane is done with this stage before it is refilled
    }
    const float res = (chain == 3) ? fabsf(sacc) : __fmul_rn(sacc, __fdiv_rn(1.0f, (float)nv));
    const int b = lane & ~3;
    const float cx = __shfl_sync(0xffffffffu, res, b), cy = __shfl_sync(0xffffffffu, res, b + 1), cz = __shfl_sync(0xffffffffu, res, b + 2),
                vol = __shfl_sync(0xffffffffu, res, b + 3);
    if (live && chain == 0) {  // one lane per cell: fold the four warps' partials and write the cell's bounds record
      const int cc = c0 + cell;
      const float4 *pr = P.part + (size_t)cc * (STEP_THREADS / 32) * 3;
      float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY}, rr = 0.f, em = 0.f;
      bool st = true;
#pragma unroll
      for (int w = 0; w < STEP_THREADS / 32; w++) {
        const float4 a = __ldcg(pr + 3 * w), b2 = __ldcg(pr + 3 * w + 1), c2 = __ldcg(pr + 3 * w + 2);
        l[0] = fminf(l[0], a.x); l[1] = fminf(l[1], a.y); l[2] = fminf(l[2], a.z); rr = fmaxf(rr, a.w);
        h[0] = fmaxf(h[0], b2.x); h[1] = fmaxf(h[1], b2.y); h[2] = fmaxf(h[2], b2.z);
        em = fmaxf(em, c2.x); st = st && c2.y != 0.0f;
      }
      const float pad = CONTACT_PAD * sqrtf(em);
      const float4 old2 = P.bnd_in[BND * (size_t)cc + 2];  // (COM, volume) of the positions this timestep started from
      float4 *bnd_cell = P.bnd_out + BND * (size_t)cc;
      bnd_cell[0] = make_float4(l[0], l[1], l[2], rr);
      bnd_cell[1] = make_float4(h[0], h[1], h[2], pad);
      bnd_cell[2] = make_float4(cx, cy, cz, vol);
      // previous volume (compat mode only); l0 travels with the bounds (ghost cells have no parameters)
      bnd_cell[3] = make_float4(0.f, st ? 1.f : 0.f, old2.w, P.cellB[cc].y);
      bnd_cell[4] = make_float4(old2.x, old2.y, old2.z, 0.f);  // kernel point of the new positions
      if (P.push_slot) {  // the same record (and the cell's global id) into the inboxes of the ranks that have it as a ghost
        const int2 ps = psc;
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int slot = q == 0 ? ps.x : ps.y;
          if (slot >= 0) {
            float4 *pb = P.push_bnd[q] + BND * (size_t)slot;
            pb[0] = make_float4(l[0], l[1], l[2], rr);
            pb[1] = make_float4(h[0], h[1], h[2], pad);
            pb[2] = make_float4(cx, cy, cz, vol);
            pb[3] = make_float4(0.f, st ? 1.f : 0.f, old2.w, P.cellB[cc].y);
            pb[4] = make_float4(old2.x, old2.y, old2.z, 0.f);
            P.push_gidp[q][slot] = P.push_gid[cc];
          }
        }
      }
      {  // neighbour-list validity
        const float4 bl = P.bbox_lo[cc], bh = P.bbox_hi[cc];
        if (l[0] < bl.x || l[1] < bl.y || l[2] < bl.z || h[0] > bh.x || h[1] > bh.y || h[2] > bh.z) P.st->rebuild = 1;
        if (pad > P.st->range) P.st->rebuild = 1;  // the candidate lists were built for smaller contact pads
      }
    }
  }
}

}  // namespace dpm
