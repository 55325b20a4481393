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
// The reference skips faces with denom < 1e-8 (shaders/Cell3D_Kernel.cl:293-295): every face that subtends more
// than pi steradians from the vertex.  Its winding number is therefore the true one (0 outside a closed mesh)
// only if no face is that close.  Every point of a face lies within 2/3 of its longest median, hence within
// (2/3) e of its centroid (e = longest edge); a ball of that radius subtends < pi steradians once the distance to
// its centre exceeds (2/sqrt 3)(2/3) e = 0.7698 e.  Centroids lie inside the cell's bounding box/sphere, so a
// vertex farther than CONTACT_PAD * emax from them gets exactly zero repulsion from that cell and is culled.
constexpr float CONTACT_PAD = 0.775f;
constexpr float RANGE_HEADROOM = 1.25f;  // lists are built for pads up to 1.25x the largest current one

template <int THREADS>
size_t step3d_smem_bytes(int nv, int nf, int K) {
  size_t b = 0;
  b += sizeof(float4) * nv;                      // sP
  b += sizeof(float4) * nv * (THREADS / 32);     // sU (per-warp unit vectors)
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

// ---------------------------------------------------------------------------------
// Per-cell bounds of a position array (used once after upload; afterwards the step
// kernel's epilogue keeps them current).  One CTA of 128 threads per cell.
// ---------------------------------------------------------------------------------
static __global__ void dpm3d_bounds_kernel(const float4 *pos, float4 *bnd, int nc, int nv, const uint16_t *ring_nbr,
                                           const uint8_t *valence, int ring_stride) {
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
  float r2 = 0.0f, e2 = 0.0f;
  for (int v = tid; v < nv; v += blockDim.x) {
    float3 q = sub3(sP[v], com);
    r2 = fmaxf(r2, dot3(q, q));
    const int val = valence[v];
    for (int i = 0; i < val; i++) { float3 e = sub3(sP[ring_nbr[(size_t)v * ring_stride + i]], sP[v]); e2 = fmaxf(e2, dot3(e, e)); }
  }
  r2 = warp_max(r2);
  e2 = warp_max(e2);
  if (lane == 0) { sRed[warp][6] = r2; sRed[warp][7] = e2; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)blockDim.x / 32; w++) {
      for (int d = 0; d < 3; d++) { sRed[0][d] = fminf(sRed[0][d], sRed[w][d]); sRed[0][3 + d] = fmaxf(sRed[0][3 + d], sRed[w][3 + d]); }
      sRed[0][6] = fmaxf(sRed[0][6], sRed[w][6]);
      sRed[0][7] = fmaxf(sRed[0][7], sRed[w][7]);
    }
    bnd[3 * (size_t)ci + 0] = make_float4(sRed[0][0], sRed[0][1], sRed[0][2], sRed[0][6]);
    bnd[3 * (size_t)ci + 1] = make_float4(sRed[0][3], sRed[0][4], sRed[0][5], CONTACT_PAD * sqrtf(sRed[0][7]));
    bnd[3 * (size_t)ci + 2] = make_float4(com.x, com.y, com.z, 0.f);
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
  float4 *sU = sP + nv;
  float4 *sShift = sU + (size_t)nv * NW;
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
  const float4 bi0 = P.bnd_in[3 * (size_t)ci], bi1 = P.bnd_in[3 * (size_t)ci + 1], bi2 = P.bnd_in[3 * (size_t)ci + 2];
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
      const float4 bj0 = P.bnd_in[3 * (size_t)cj], bj1 = P.bnd_in[3 * (size_t)cj + 1], bj2 = P.bnd_in[3 * (size_t)cj + 2];
      float3 sh = f3(0.f, 0.f, 0.f);
      if (P.pbc) {  // shift = L * round((COMi - COMJ) / L)   (:277-281)
        sh.x = P.L * roundf((com.x - bj2.x) / P.L);
        sh.y = P.L * roundf((com.y - bj2.y) / P.L);
        sh.z = P.L * roundf((com.z - bj2.z) / P.L);
      }
      // padded, shifted bounding box of cj: outside it the reference's formula gives exactly zero
      const float pad = bj1.w;  // CONTACT_PAD * (upper bound of cj's longest edge)
      const float4 lo = make_float4((bj0.x + sh.x) - pad, (bj0.y + sh.y) - pad, (bj0.z + sh.z) - pad, 0.f);
      const float4 hi = make_float4((bj1.x + sh.x) + pad, (bj1.y + sh.y) + pad, (bj1.z + sh.z) + pad, 0.f);
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
    // processes the U queued units, then folds their forces into the owning threads
    auto flush = [&]() {
      __syncthreads();
      for (int u = warp; u < U; u += NW) {
        const int code = sUnit[u];
        const int v = code & 0xffff, k = code >> 16;
        const int cj = sCand[k];
        const float4 sh = sShift[k];
        const float4 p = sP[v];
        const float4 *Vj = P.pos_in + (size_t)cj * nv;
        float4 *myU = sU + (size_t)warp * nv;
        for (int i = lane; i < nv; i += 32) {
          const float4 q = __ldg(Vj + i);
          // a = V + shift - p ; u = normalize(a)   (:285-291)
          const float ax = (q.x + sh.x) - p.x, ay = (q.y + sh.y) - p.y, az = (q.z + sh.z) - p.z;
          // near-coplanar faces make the triple product cancel: refine MUFU.RSQ (2 ulp) with one Newton step so the
          // unit vectors are as accurate as the reference's a / sqrt(dot(a, a))
          const float d2 = ax * ax + ay * ay + az * az;
          float r = rsqrtf(d2);
          r = r * (1.5f - 0.5f * d2 * r * r);
          myU[i] = make_float4(ax * r, ay * r, az * r, 0.f);
        }
        __syncwarp();
        float om = 0.0f;
        for (int f = lane; f < nf; f += 32) {
          const ushort4 fc = __ldg(P.faces + f);
          const float4 a = myU[fc.x], b = myU[fc.y], c = myU[fc.z];
          const float den = 1.0f + (a.x * b.x + a.y * b.y + a.z * b.z) + (b.x * c.x + b.y * c.y + b.z * c.z) +
                            (c.x * a.x + c.y * a.y + c.z * a.z);
          const float num = a.x * (b.y * c.z - b.z * c.y) + a.y * (b.z * c.x - b.x * c.z) + a.z * (b.x * c.y - b.y * c.x);
          if (!(den < 1e-8f)) om += 2.0f * atan2f(num, den);  // :293-299
        }
        om = warp_sum(om);
        if (lane == 0) sUnitW[u] = om / (4.0f * 3.14159274101257f);
        __syncwarp();
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
  float r2 = 0.0f;
#pragma unroll
  for (int j = 0; j < VPT; j++) {
    const int v = tid + j * THREADS;
    if (v < nv) { const float3 q = sub3(myP[j], ncom); r2 = fmaxf(r2, dot3(q, q)); }
  }
  r2 = warp_max(r2);
  if (lane == 0) { sRed[warp][6] = r2; sRed[warp][7] = e2max; }
  __syncthreads();
  if (tid == 0) {
    float l[3], h[3], rr = sRed[0][6], eb = sRed[0][7];
    for (int d = 0; d < 3; d++) { l[d] = sRed[0][d]; h[d] = sRed[0][3 + d]; }
    for (int w = 1; w < NW; w++) {
      for (int d = 0; d < 3; d++) { l[d] = fminf(l[d], sRed[w][d]); h[d] = fmaxf(h[d], sRed[w][3 + d]); }
      rr = fmaxf(rr, sRed[w][6]);
      eb = fmaxf(eb, sRed[w][7]);
    }
    const float pad = CONTACT_PAD * sqrtf(eb);
    P.bnd_out[3 * (size_t)ci + 0] = make_float4(l[0], l[1], l[2], rr);
    P.bnd_out[3 * (size_t)ci + 1] = make_float4(h[0], h[1], h[2], pad);
    if (pad > P.st->range) P.st->rebuild = 1;  // the candidate lists were built for smaller contact pads
    P.bnd_out[3 * (size_t)ci + 2] = make_float4(ncom.x, ncom.y, ncom.z, sScalar[0]);
    const float4 bl = P.bbox_lo[ci], bh = P.bbox_hi[ci];
    if (l[0] < bl.x || l[1] < bl.y || l[2] < bl.z || h[0] > bh.x || h[1] > bh.y || h[2] > bh.z) P.st->rebuild = 1;
  }
}


}  // namespace dpm
