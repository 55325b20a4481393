// dpm3d_kernels.cuh — fused 3D force + integrate step for sm_100a.
//
// Replaces the six per-step OpenCL kernels of shaders/Cell3D_Kernel.cl
// (ClearForces :366, VolumeForceUpdate :66, SurfaceAreaForceUpdate :114,
//  StickToSurface :180, RepellingForces :251, EulerPosition :371) and their
// enqueue sequence (src/Tissue3D.cpp:372-423) by ONE kernel per timestep:
//
//   CTA = one cell.  The cell's vertex ring (float4) is staged in shared memory,
//   shape forces are GATHERED per vertex over a constant ring adjacency (no float
//   atomics, unlike atomic_add_f :10-32), the repulsion evaluates the winding number
//   only for (vertex, neighbour-cell) pairs that survive the cell list + an exact
//   AABB / bounding-sphere cull, one warp per pair with the neighbour's unit vectors
//   staged in shared memory, and the Euler update writes the other position buffer
//   (all forces of a step use start-of-step positions).  The epilogue
//   produces next step's per-cell bounds (AABB, COM, r^2max) and raises the
//   rebuild flag when a cell leaves its build-time box.
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
  const uint16_t *__restrict__ ring_nbr;
  const uint16_t *__restrict__ ring_face;
  const uint8_t *__restrict__ valence;
  int ring_stride;
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
  int nc;  // cells stepped by this launch (owned)
  int nv, nf;
  float dt, Kc;
  int pbc;
  float L;
  unsigned mask;
  int stale_from;  // >= 0: reference-race compatibility mode (dpm3d_set_compat), else -1
};

constexpr int UNIT_CAP_FACTOR = 2;  // unit list capacity = factor * THREADS
constexpr int BND = 4;              // float4 per cell in the bounds arrays:
                                    //   (lo.xyz, r2max) (hi.xyz, contact pad) (com.xyz, volume) (r2min, star flag, 0, 0)
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
constexpr int RING_MAX = 5;              // edge-adjacency rings examined around the radially hit face
constexpr int RING_TAB = 48;             // 1 + 3 + 6 + 9 + 12 + 15 = 46 faces
constexpr int DIR_N = 16;                // octahedral map resolution
constexpr int MAX_WALK = 64;
constexpr int UNIT_LANES = 8;            // lanes cooperating on one (vertex, neighbour) unit: ring faces / literal faces in parallel

__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = UNIT_LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

template <int THREADS>
size_t step3d_smem_bytes(int nv, int nf, int K) {
  size_t b = 0;
  b += sizeof(float4) * nv;                      // sP
  b += sizeof(float) * nf;                       // sTerm
  b += sizeof(float4) * K * 3;                   // per-candidate shift / lo / hi
  b += sizeof(float4) * K;                       // per-candidate sphere (com+shift, r2)
  b += sizeof(int) * K;                          // candidate ids
  b += sizeof(int) * UNIT_CAP_FACTOR * THREADS;  // unit codes
  b += sizeof(float) * UNIT_CAP_FACTOR * THREADS;  // unit winding numbers
  b += ((nf + 15) / 16) * 16;                    // sFlag
  return b + 64;
}



__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 sub3(float4 a, float4 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 sub3(float4 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// Serial, individually rounded COM of nv float4 vertices in shared memory; lanes 0..2 each own
// one component (shaders/Cell3D_Kernel.cl:35-44: sum in index order, then * 1/(float)NV).
__device__ __forceinline__ float com_chain(const float4 *sP, int nv, int comp) {
  const float *base = reinterpret_cast<const float *>(sP) + comp;
  float s = 0.0f;
  int i = 0;
  for (; i + 8 <= nv; i += 8) {
    float t[8];
#pragma unroll
    for (int q = 0; q < 8; q++) t[q] = base[4 * (i + q)];
#pragma unroll
    for (int q = 0; q < 8; q++) s = __fadd_rn(s, t[q]);
  }
  for (; i < nv; i++) s = __fadd_rn(s, base[4 * i]);
  return __fmul_rn(s, __fdiv_rn(1.0f, (float)nv));
}

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
  ra = ra * (1.5f - 0.5f * da * ra * ra);
  rb = rb * (1.5f - 0.5f * db * rb * rb);
  rc = rc * (1.5f - 0.5f * dc * rc * rc);
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
static __device__ __noinline__ float omega_skipped_f64(float3 a, float3 b, float3 c) {
  const double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z, cx = c.x, cy = c.y, cz = c.z;
  const double la = sqrt(ax * ax + ay * ay + az * az), lb = sqrt(bx * bx + by * by + bz * bz), lc = sqrt(cx * cx + cy * cy + cz * cz);
  const double num = ax * (by * cz - bz * cy) + ay * (bz * cx - bx * cz) + az * (bx * cy - by * cx);
  const double den = la * lb * lc + (ax * bx + ay * by + az * bz) * lc + (bx * cx + by * cy + bz * cz) * la + (cx * ax + cy * ay + cz * az) * lb;
  return (float)(2.0 * atan2(num, den));
}

// Returns 0 on success, else the reason the caller must fall back to the literal sum (1: vertex within the pad of
// the neighbour's COM, 2: walk limit, 3: ring limit).
__device__ __forceinline__ int winding_fast(const Step3DParams &P, const float4 *__restrict__ Vj, float4 sh, float4 p, float3 Cs,
                                            float rho, float &w_out, int g, unsigned gmask) {
  const float3 u = f3(p.x - Cs.x, p.y - Cs.y, p.z - Cs.z);
  const float r2 = dot3(u, u);
  if (!(r2 > 1.0201f * rho * rho)) return 1;
  int f = __ldg(P.dir_table + octa_texel(u.x, u.y, u.z));
  float3 A, B, C;
  bool found = false;
  for (int it = 0; it < MAX_WALK; it++) {
    const ushort4 fc = __ldg(P.faces + f);
    const float4 q0 = __ldg(Vj + fc.x), q1 = __ldg(Vj + fc.y), q2 = __ldg(Vj + fc.z);
    A = f3((q0.x + sh.x) - Cs.x, (q0.y + sh.y) - Cs.y, (q0.z + sh.z) - Cs.z);
    B = f3((q1.x + sh.x) - Cs.x, (q1.y + sh.y) - Cs.y, (q1.z + sh.z) - Cs.z);
    C = f3((q2.x + sh.x) - Cs.x, (q2.y + sh.y) - Cs.y, (q2.z + sh.z) - Cs.z);
    const float d0 = dot3(u, cross3(A, B)), d1 = dot3(u, cross3(B, C)), d2 = dot3(u, cross3(C, A));
    const float dm = fminf(d0, fminf(d1, d2));
    if (dm >= 0.0f) { found = true; break; }
    const ushort4 ad = __ldg(P.face_adj + f);
    f = (dm == d0) ? ad.x : (dm == d1 ? ad.y : ad.z);
  }
  if (!found) return 2;
  // inside <=> p on the inner side of the hit face's plane (the COM is on the inner side of every face)
  const float3 n = cross3(f3(B.x - A.x, B.y - A.y, B.z - A.z), f3(C.x - A.x, C.y - A.y, C.z - A.z));
  const float W = (dot3(n, f3(u.x - A.x, u.y - A.y, u.z - A.z)) < 0.0f) ? 1.0f : 0.0f;
  const float rinv = rsqrtf(r2);
  const float sinp = rho * rinv;
  const uint16_t *tab = P.ring_tab + (size_t)f * RING_TAB;
  const uint8_t *rend = P.ring_end + (size_t)f * (RING_MAX + 1);
  float corr = 0.0f;
  int jbeg = 0;
  bool open = true;  // the last examined ring still touched the cone
  for (int ring = 0; ring <= RING_MAX && open; ring++) {  // uniform within the group: all its lanes hold the same unit
    const int jend = __ldg(rend + ring);
    unsigned touched = 0;
    for (int base = jbeg; base < jend; base += UNIT_LANES) {
      const int j = base + g;
      bool hit = false;
      if (j < jend) {
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
      const float s2 = sinp * sinp * r2 * 1.002f, c2 = (1.0f - sinp * sinp) * r2 * 0.998f;
      hit = (e0 >= 0.0f && e1 >= 0.0f && e2 >= 0.0f);
      hit = hit || (ua > 0.0f && ua * ua >= c2 * aa) || (ub > 0.0f && ub * ub >= c2 * bb) || (uc > 0.0f && uc * uc >= c2 * cc);
      hit = hit || (e0 < 0.0f && e0 * e0 <= s2 * dot3(n0, n0) && aa * ub - ab * ua >= 0.0f && ua * bb - ub * ab >= 0.0f);
      hit = hit || (e1 < 0.0f && e1 * e1 <= s2 * dot3(n1, n1) && bb * uc - bc * ub >= 0.0f && ub * cc - uc * bc >= 0.0f);
      hit = hit || (e2 < 0.0f && e2 * e2 <= s2 * dot3(n2, n2) && cc * ua - ca * uc >= 0.0f && uc * aa - ua * ca >= 0.0f);
      if (hit) {
        float den, num;
        solid_angle_terms(a, b, c, den, num);
        if (den < 1e-8f) corr += omega_skipped_f64(a, b, c);
      }
      }
      touched |= __ballot_sync(gmask, hit);
    }
    open = touched != 0;
    jbeg = jend;
  }
  if (open) return 3;  // the outermost tabulated ring still touches the cone
  w_out = W - group_sum(corr, gmask) / (4.0f * 3.14159274101257f);
  return 0;
}

// ---------------------------------------------------------------------------------
// Per-cell bounds of a position array (used once after upload; afterwards the step
// kernel's epilogue keeps them current).  One CTA of 128 threads per cell.
// ---------------------------------------------------------------------------------
// signed volume of the tetrahedron (C, P0, P1, P2) is positive with a margin for every face  <=>  the mesh is
// star-shaped about C (closed, consistently oriented): the precondition of winding_fast
__device__ __forceinline__ bool face_sees_centre(float4 P0, float4 P1, float4 P2, float3 C) {
  const float3 a = sub3(P0, C), b = sub3(P1, C), c = sub3(P2, C);
  const float3 cr = cross3(b, c);
  const float sv = dot3(a, cr);
  return sv > 0.0f && sv * sv > 1e-6f * dot3(a, a) * dot3(cr, cr);
}

static __global__ void dpm3d_bounds_kernel(const float4 *pos, float4 *bnd, int nc, int nv, const uint16_t *ring_nbr,
                                           const uint8_t *valence, int ring_stride, const ushort4 *faces, int nf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *sP = reinterpret_cast<float4 *>(smem_raw);
  __shared__ float sRed[4][8];
  __shared__ float sCom[3];
  const int ci = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int v = tid; v < nv; v += blockDim.x) {
    float4 p = pos[(size_t)ci * nv + v];
    sP[v] = p;
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
  }
  for (int d = 0; d < 3; d++) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
  if (lane == 0) for (int d = 0; d < 3; d++) { sRed[warp][d] = lo[d]; sRed[warp][3 + d] = hi[d]; }
  __syncthreads();
  if (warp == 0 && lane < 3) sCom[lane] = com_chain(sP, nv, lane);
  __syncthreads();
  float3 com = f3(sCom[0], sCom[1], sCom[2]);
  float r2 = 0.0f, e2 = 0.0f, r2min = INFINITY;
  int star = 1;
  for (int f = tid; f < nf; f += blockDim.x) {
    const ushort4 fc = faces[f];
    star &= face_sees_centre(sP[fc.x], sP[fc.y], sP[fc.z], com) ? 1 : 0;
  }
  for (int v = tid; v < nv; v += blockDim.x) {
    float3 q = sub3(sP[v], com);
    r2 = fmaxf(r2, dot3(q, q));
    r2min = fminf(r2min, dot3(q, q));
    const int val = valence[v];
    for (int i = 0; i < val; i++) { float3 e = sub3(sP[ring_nbr[(size_t)v * ring_stride + i]], sP[v]); e2 = fmaxf(e2, dot3(e, e)); }
  }
  r2 = warp_max(r2);
  e2 = warp_max(e2);
  r2min = warp_min(r2min);
  if (lane == 0) { sRed[warp][6] = r2; sRed[warp][7] = e2; }
  star = __syncthreads_and(star);
  __shared__ float sMin[4];
  if (lane == 0) sMin[warp] = r2min;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)blockDim.x / 32; w++) sMin[0] = fminf(sMin[0], sMin[w]);
    bnd[BND * (size_t)ci + 3] = make_float4(sMin[0], star ? 1.f : 0.f, 0.f, 0.f);
    for (int w = 1; w < (int)blockDim.x / 32; w++) {
      for (int d = 0; d < 3; d++) { sRed[0][d] = fminf(sRed[0][d], sRed[w][d]); sRed[0][3 + d] = fmaxf(sRed[0][3 + d], sRed[w][3 + d]); }
      sRed[0][6] = fmaxf(sRed[0][6], sRed[w][6]);
      sRed[0][7] = fmaxf(sRed[0][7], sRed[w][7]);
    }
    bnd[BND * (size_t)ci + 0] = make_float4(sRed[0][0], sRed[0][1], sRed[0][2], sRed[0][6]);
    bnd[BND * (size_t)ci + 1] = make_float4(sRed[0][3], sRed[0][4], sRed[0][5], CONTACT_PAD * sqrtf(sRed[0][7]));
    bnd[BND * (size_t)ci + 2] = make_float4(com.x, com.y, com.z, 0.f);
  }
}

// ---------------------------------------------------------------------------------
// The fused step kernel.
// ---------------------------------------------------------------------------------
template <int THREADS, int VPT>
__global__ void __launch_bounds__(THREADS) dpm3d_step_kernel(Step3DParams P) {
  constexpr int NW = THREADS / 32;
  constexpr int UCAP = UNIT_CAP_FACTOR * THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nv = P.nv, nf = P.nf, K = P.K;
  float4 *sP = reinterpret_cast<float4 *>(smem_raw);
  float4 *sShift = sP + nv;
  float4 *sLo = sShift + K;
  float4 *sHi = sLo + K;
  float4 *sSph = sHi + K;
  float *sTerm = reinterpret_cast<float *>(sSph + K);
  float *sUnitW = sTerm + nf;
  int *sUnit = reinterpret_cast<int *>(sUnitW + UCAP);
  int *sCand = sUnit + UCAP;
  unsigned char *sFlag = reinterpret_cast<unsigned char *>(sCand + K);
  __shared__ float sRed[NW][8];
  __shared__ float sScalar[12];
  __shared__ int sCnt[2][NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ci = blockIdx.x;
  const float4 cA = P.cellA[ci], cB = P.cellB[ci];
  const float Kv = cA.x, Ka = cA.y, Ks = cA.z, v0 = cA.w, a0 = cB.x, l0 = cB.y;
  const float4 bi0 = P.bnd_in[BND * (size_t)ci], bi1 = P.bnd_in[BND * (size_t)ci + 1], bi2 = P.bnd_in[BND * (size_t)ci + 2];
  const float3 com = f3(bi2.x, bi2.y, bi2.z);
  const float4 *gP = P.pos_in + (size_t)ci * nv;

  // ---- phase 0: stage the vertex ring --------------------------------------------
  float4 myP[VPT];
  float3 F[VPT];
#pragma unroll
  for (int j = 0; j < VPT; j++) {
    int v = tid + j * THREADS;
    F[j] = f3(0.f, 0.f, 0.f);
    if (v < nv) { myP[j] = gP[v]; sP[v] = myP[j]; } else myP[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  // ---- phase 1: per-face pass: signed-volume terms, down-facing and degenerate flags -------
  const bool doVol = (P.mask & DPM3D_VOLUME) && (Kv != 0.0f);
  for (int f = tid; f < nf; f += THREADS) {
    const ushort4 fc = __ldg(P.faces + f);
    const float4 P0 = sP[fc.x], P1 = sP[fc.y], P2 = sP[fc.z];
    // dot(cross(P0,P1),P2)/6.0f with the reference's operation order, unfused (:58-61)
    float cx = __fsub_rn(__fmul_rn(P0.y, P1.z), __fmul_rn(P0.z, P1.y));
    float cy = __fsub_rn(__fmul_rn(P0.z, P1.x), __fmul_rn(P0.x, P1.z));
    float cz = __fsub_rn(__fmul_rn(P0.x, P1.y), __fmul_rn(P0.y, P1.x));
    float tp = __fadd_rn(__fadd_rn(__fmul_rn(cx, P2.x), __fmul_rn(cy, P2.y)), __fmul_rn(cz, P2.z));
    sTerm[f] = __fdiv_rn(tp, 6.0f);
    const float3 A = sub3(P1, P0), B = sub3(P2, P0), C = sub3(P2, P1);
    const float3 n = cross3(A, B);
    const bool down = n.z * rsqrtf(dot3(n, n)) < -0.1f;  // StickToSurface :209-214
    const bool deg = dot3(A, A) < 1e-24f || dot3(B, B) < 1e-24f || dot3(C, C) < 1e-24f;  // :151
    sFlag[f] = (unsigned char)((down ? 1 : 0) | (deg ? 2 : 0));
  }
  __syncthreads();

  // ---- phase 2a: serial volume chain (last warp, all lanes redundantly: broadcast LDS) ------
  if (warp == NW - 1) {
    float vol = 0.0f;
    if (doVol) {
      int f = 0;
      for (; f + 8 <= nf; f += 8) {
        float t[8];
#pragma unroll
        for (int q = 0; q < 8; q++) t[q] = sTerm[f + q];
#pragma unroll
        for (int q = 0; q < 8; q++) vol = __fadd_rn(vol, t[q]);
      }
      for (; f < nf; f++) vol = __fadd_rn(vol, sTerm[f]);
    }
    if (lane == 0) sScalar[0] = fabsf(vol);
  }

  // ---- phase 2b: ring pass (edge springs + volume gradient + flag gather) ---------------------
  const float inv_l0 = 1.0f / l0;
  int ndown[VPT];
  float3 G[VPT], Gs[VPT];  // volume gradient: all ring faces / ring faces with index >= stale_from (compat mode)
  float e2max = 0.0f;
#pragma unroll
  for (int j = 0; j < VPT; j++) {
    const int v = tid + j * THREADS;
    ndown[j] = 0;
    G[j] = f3(0.f, 0.f, 0.f);
    Gs[j] = f3(0.f, 0.f, 0.f);
    if (v < nv) {
      const int val = __ldg(P.valence + v);
      const uint16_t *rn = P.ring_nbr + (size_t)v * P.ring_stride;
      const uint16_t *rf = P.ring_face + (size_t)v * P.ring_stride;
      unsigned m = 0;
      for (int i = 0; i < val; i++) m |= (unsigned)sFlag[__ldg(rf + i)] << (2 * i);
      ndown[j] = __popc(m & 0x55555555u);
      const float4 Pv = myP[j];
      float3 T = f3(0.f, 0.f, 0.f), g = f3(0.f, 0.f, 0.f), Qp = f3(0.f, 0.f, 0.f), Q0 = f3(0.f, 0.f, 0.f);
      for (int i = 0; i < val; i++) {
        const float4 Pn = sP[__ldg(rn + i)];
        const float3 E = sub3(Pn, Pv);
        const float len2 = dot3(E, E);
        const float rl = rsqrtf(len2);
        const float dl = len2 * rl * inv_l0 - 1.0f;  // len/l0 - 1  (:156-160)
        const int ip = (i == 0) ? val - 1 : i - 1;
        // edge (v, n_i) belongs to ring faces i-1 and i; each contributes unit(E)*dl unless degenerate
        const float w = (float)(2 - ((m >> (2 * i + 1)) & 1u) - ((m >> (2 * ip + 1)) & 1u));
        const float s = rl * dl * w;
        T.x += E.x * s; T.y += E.y * s; T.z += E.z * s;
        const float3 Q = sub3(Pn, com);
        if (i == 0) Q0 = Q;
        else {
          const float3 c = cross3(Qp, Q);  // gradient of ring face i-1 = (v, n_{i-1}, n_i)
          g.x += c.x; g.y += c.y; g.z += c.z;
          if (P.stale_from >= 0 && (int)__ldg(rf + i - 1) >= P.stale_from) { Gs[j].x += c.x; Gs[j].y += c.y; Gs[j].z += c.z; }
        }
        Qp = Q;
      }
      {
        const float3 c = cross3(Qp, Q0);  // ring face val-1 = (v, n_{val-1}, n_0)
        g.x += c.x; g.y += c.y; g.z += c.z;
        if (P.stale_from >= 0 && (int)__ldg(rf + val - 1) >= P.stale_from) { Gs[j].x += c.x; Gs[j].y += c.y; Gs[j].z += c.z; }
      }
      G[j] = g;
      if ((P.mask & DPM3D_AREA) && !(Ka < 1e-8f)) {
        const float scale = Ka * sqrtf(a0) / l0 * 0.3f;  // :162
        F[j] = f3(T.x * scale, T.y * scale, T.z * scale);
      }
    }
  }
  __syncthreads();  // volume chain done

  // ---- phase 2c: volume force + substrate adhesion --------------------------------------------
  {
    const float volume = sScalar[0];
    const float coef = doVol ? (-Kv * (volume / v0 - 1.0f)) * (1.0f / 6.0f) : 0.0f;  // :85,:106-108
    // compat mode: faces >= stale_from see the volume of the previous step's start (0 right after an upload)
    const float dcoef = (doVol && P.stale_from >= 0) ? (-Kv * (bi2.w / v0 - 1.0f)) * (1.0f / 6.0f) - coef : 0.0f;
    const bool doStick = (P.mask & DPM3D_STICK) && !(Ks < 1e-12f);
#pragma unroll
    for (int j = 0; j < VPT; j++) {
      const int v = tid + j * THREADS;
      if (v < nv) {
        F[j].x += coef * G[j].x + dcoef * Gs[j].x;
        F[j].y += coef * G[j].y + dcoef * Gs[j].y;
        F[j].z += coef * G[j].z + dcoef * Gs[j].z;
        if (doStick && ndown[j] > 0) {
          const float4 Pv = myP[j];
          const float nd = (float)ndown[j];  // one application per adjacent down-facing face (:226-246)
          const float h = fabsf(Pv.z);
          if (Pv.z < 0.0f) F[j].z += nd * (Ks * h);
          if (h < l0 * 2.0f) {
            const float3 ctv = f3(Pv.x - com.x, Pv.y - com.y, 0.0f - com.z);
            const float ftmp = Ks * (1.0f - h / l0);
            const float s = rsqrtf(dot3(ctv, ctv)) * ftmp * nd;
            F[j].x += ctv.x * s; F[j].y += ctv.y * s; F[j].z += ctv.z * s;
          }
        }
      }
    }
  }

  // ---- phase 3: repulsion (winding number) over surviving (vertex, neighbour) units ----------
  if ((P.mask & DPM3D_REPEL) && P.Kc != 0.0f) {
    const int ncand = min(P.cand_count[ci], K);
    for (int k = tid; k < ncand; k += THREADS) {
      const int cj = P.cand[(size_t)ci * K + k];
      const float4 bj0 = P.bnd_in[BND * (size_t)cj], bj1 = P.bnd_in[BND * (size_t)cj + 1], bj2 = P.bnd_in[BND * (size_t)cj + 2];
      const float4 bj3 = P.bnd_in[BND * (size_t)cj + 3];
      float3 sh = f3(0.f, 0.f, 0.f);
      if (P.pbc) {  // shift = L * round((COMi - COMJ) / L)   (:277-281)
        sh.x = P.L * roundf((com.x - bj2.x) / P.L);
        sh.y = P.L * roundf((com.y - bj2.y) / P.L);
        sh.z = P.L * roundf((com.z - bj2.z) / P.L);
      }
      // padded, shifted bounding box of cj: outside it the reference's formula gives exactly zero
      const float pad = bj1.w;  // CONTACT_PAD * (upper bound of cj's longest edge)
      const float4 lo = make_float4((bj0.x + sh.x) - pad, (bj0.y + sh.y) - pad, (bj0.z + sh.z) - pad, pad);
      const float4 hi = make_float4((bj1.x + sh.x) + pad, (bj1.y + sh.y) + pad, (bj1.z + sh.z) + pad, bj3.y);  // .w: star-shaped
      const bool ov = !(lo.x > bi1.x || hi.x < bi0.x || lo.y > bi1.y || hi.y < bi0.y || lo.z > bi1.z || hi.z < bi0.z);
      sCand[k] = ov ? cj : -1;
      sShift[k] = make_float4(sh.x, sh.y, sh.z, 0.f);
      sLo[k] = lo;
      sHi[k] = hi;
      const float rs = sqrtf(bj0.w) + pad;
      sSph[k] = make_float4(bj2.x + sh.x, bj2.y + sh.y, bj2.z + sh.z, rs * rs * 1.0001f + 1e-30f);
    }
    __syncthreads();

    int U = 0, par = 0;
    unsigned long long evals = 0;
    int nlit = 0;
    // processes the U queued units, then folds their forces into the owning threads
    auto flush = [&]() {
      __syncthreads();
      // UNIT_LANES lanes per (vertex, neighbour) unit: the group walks together and splits the ring faces
      const int g = lane & (UNIT_LANES - 1);
      const unsigned gmask = ((1u << UNIT_LANES) - 1u) << (lane & ~(UNIT_LANES - 1));
      for (int u = tid / UNIT_LANES; u < U; u += THREADS / UNIT_LANES) {
        const int code = sUnit[u];
        const int v = code & 0xffff, k = code >> 16;
        const int cj = sCand[k];
        const float4 sh = sShift[k];
        const float4 p = sP[v];
        const float4 sp = sSph[k];
        const float4 *Vj = P.pos_in + (size_t)cj * nv;
        float w;
        int why = -1;  // -1: neighbour not star-shaped about its COM
        if (sHi[k].w != 0.0f) why = winding_fast(P, Vj, sh, p, f3(sp.x, sp.y, sp.z), sLo[k].w, w, g, gmask);
        if (why != 0) {
          w = winding_literal(Vj, P.faces, nf, sh, p, g, gmask);
          if (g == 0) { nlit++; atomicAdd(&P.st->fallback_why[why < 0 ? 0 : why], 1ull); }
        }
        if (g == 0) sUnitW[u] = w;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < VPT; j++) {
        const int v = tid + j * THREADS;
        if (v < nv) {
          float3 dir = f3(0.f, 0.f, 0.f);
          bool have = false;
          for (int u = 0; u < U; u++) {
            if ((sUnit[u] & 0xffff) == v) {
              const float wn = sUnitW[u];
              if (!(fabsf(wn) < 1e-6f)) {  // :302-308
                if (!have) {
                  const float3 d = f3(com.x - myP[j].x, com.y - myP[j].y, com.z - myP[j].z);
                  const float r = rsqrtf(dot3(d, d));
                  dir = f3(d.x * r, d.y * r, d.z * r);
                  have = true;
                }
                const float mg = fabsf(wn) * 0.5f * P.Kc;
                F[j].x += mg * dir.x; F[j].y += mg * dir.y; F[j].z += mg * dir.z;
              }
            }
          }
        }
      }
      evals += U;
      U = 0;
      __syncthreads();
    };

    for (int k = 0; k < ncand; k++) {
      if (sCand[k] < 0) continue;  // uniform
      const float4 lo = sLo[k], hi = sHi[k], sp = sSph[k];
#pragma unroll
      for (int j = 0; j < VPT; j++) {
        if (U + THREADS > UCAP) flush();
        const int v = tid + j * THREADS;
        bool flag = false;
        if (v < nv) {
          const float4 p = myP[j];
          flag = !(p.x < lo.x || p.x > hi.x || p.y < lo.y || p.y > hi.y || p.z < lo.z || p.z > hi.z);
          if (flag) {
            const float dx = p.x - sp.x, dy = p.y - sp.y, dz = p.z - sp.z;
            flag = (dx * dx + dy * dy + dz * dz) <= sp.w;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) sCnt[par][warp] = __popc(bal);
        __syncthreads();
        int off = U, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) { const int c = sCnt[par][w]; if (w < warp) off += c; tot += c; }
        if (flag) sUnit[off + __popc(bal & ((1u << lane) - 1u))] = v | (k << 16);
        U += tot;
        par ^= 1;
      }
    }
    if (U > 0) flush();
    if (tid == 0 && evals) atomicAdd(&P.st->contact_evals, evals);
    if (nlit) atomicAdd(&P.st->literal_evals, (unsigned long long)nlit);
  }

  // ---- phase 4: Euler update, outputs, next-step bounds ------------------------------------------
  __syncthreads();  // everyone is done reading start-of-step sP
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < VPT; j++) {
    const int v = tid + j * THREADS;
    if (v < nv) {
      float4 np = myP[j];
      np.x += F[j].x * P.dt; np.y += F[j].y * P.dt; np.z += F[j].z * P.dt;  // EulerPosition :380
      np.w = 0.f;
      P.pos_out[(size_t)ci * nv + v] = np;
      if (P.force_out) P.force_out[(size_t)ci * nv + v] = make_float4(F[j].x, F[j].y, F[j].z, 0.f);
      sP[v] = np;
      myP[j] = np;
      lo[0] = fminf(lo[0], np.x); lo[1] = fminf(lo[1], np.y); lo[2] = fminf(lo[2], np.z);
      hi[0] = fmaxf(hi[0], np.x); hi[1] = fmaxf(hi[1], np.y); hi[2] = fmaxf(hi[2], np.z);
    }
  }
  for (int d = 0; d < 3; d++) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
  if (lane == 0) for (int d = 0; d < 3; d++) { sRed[warp][d] = lo[d]; sRed[warp][3 + d] = hi[d]; }
  __syncthreads();
  if (warp == NW - 1 && lane < 3) sScalar[1 + lane] = com_chain(sP, nv, lane);
  // longest edge of the NEW positions (sets next step's contact pad); overlaps the serial COM chain
  e2max = 0.0f;
#pragma unroll
  for (int j = 0; j < VPT; j++) {
    const int v = tid + j * THREADS;
    if (v < nv) {
      const int val = __ldg(P.valence + v);
      const uint16_t *rn = P.ring_nbr + (size_t)v * P.ring_stride;
      for (int i = 0; i < val; i++) { const float3 e = sub3(sP[__ldg(rn + i)], myP[j]); e2max = fmaxf(e2max, dot3(e, e)); }
    }
  }
  e2max = warp_max(e2max);
  __syncthreads();
  const float3 ncom = f3(sScalar[1], sScalar[2], sScalar[3]);
  float r2 = 0.0f, r2min = INFINITY;
#pragma unroll
  for (int j = 0; j < VPT; j++) {
    const int v = tid + j * THREADS;
    if (v < nv) { const float3 q = sub3(myP[j], ncom); const float qq = dot3(q, q); r2 = fmaxf(r2, qq); r2min = fminf(r2min, qq); }
  }
  // is the NEW shape star-shaped about its COM?  (precondition of the neighbours' fast contact evaluation next step)
  int star = 1;
  for (int f = tid; f < nf; f += THREADS) {
    const ushort4 fc = __ldg(P.faces + f);
    star &= face_sees_centre(sP[fc.x], sP[fc.y], sP[fc.z], ncom) ? 1 : 0;
  }
  r2 = warp_max(r2);
  r2min = warp_min(r2min);
  if (lane == 0) { sRed[warp][6] = r2; sRed[warp][7] = e2max; sScalar[4 + warp] = r2min; }
  star = __syncthreads_and(star);
  if (tid == 0) {
    float rmin = sScalar[4];
    for (int w = 1; w < NW && w < 8; w++) rmin = fminf(rmin, sScalar[4 + w]);
    P.bnd_out[BND * (size_t)ci + 3] = make_float4(rmin, star ? 1.f : 0.f, 0.f, 0.f);
    float l[3], h[3], rr = sRed[0][6], eb = sRed[0][7];
    for (int d = 0; d < 3; d++) { l[d] = sRed[0][d]; h[d] = sRed[0][3 + d]; }
    for (int w = 1; w < NW; w++) {
      for (int d = 0; d < 3; d++) { l[d] = fminf(l[d], sRed[w][d]); h[d] = fmaxf(h[d], sRed[w][3 + d]); }
      rr = fmaxf(rr, sRed[w][6]);
      eb = fmaxf(eb, sRed[w][7]);
    }
    const float pad = CONTACT_PAD * sqrtf(eb);
    P.bnd_out[BND * (size_t)ci + 0] = make_float4(l[0], l[1], l[2], rr);
    P.bnd_out[BND * (size_t)ci + 1] = make_float4(h[0], h[1], h[2], pad);
    if (pad > P.st->range) P.st->rebuild = 1;  // the candidate lists were built for smaller contact pads
    P.bnd_out[BND * (size_t)ci + 2] = make_float4(ncom.x, ncom.y, ncom.z, sScalar[0]);
    const float4 bl = P.bbox_lo[ci], bh = P.bbox_hi[ci];
    if (l[0] < bl.x || l[1] < bl.y || l[2] < bl.z || h[0] > bh.x || h[1] > bh.y || h[2] > bh.z) P.st->rebuild = 1;
  }
}


}  // namespace dpm
