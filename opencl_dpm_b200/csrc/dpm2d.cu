// dpm2d.cu — fused 2D force + integrate step for sm_100a and its C ABI.
//
// Replaces the six per-step OpenCL kernels of shaders/Cell2D_kernel.cl
// (AreaForceUpdates :13, PerimeterForceUpdates :92, BendingForceUpdates :44,
//  AttractionForceUpdate :222, RepulsionForceUpdate :121, EulerUpdate :270) and their
// enqueue sequence (src/Tissue2D.cpp:215-229) by ONE kernel per timestep.
//
//   one warp = one cell.  The cell's vertex ring (float2) is staged in shared memory;
//   the polygon area and COM are evaluated once per cell (the reference recomputes the
//   area in every work-item, :27-32) in the reference's serial order; perimeter and
//   bending are ring stencils out of shared memory; attraction and repulsion visit only
//   the cells of the sorted cell list's candidate set, whose rings are staged through a
//   second shared-memory tile, with exact per-vertex culls; positions are double-buffered
//   (all forces of a step see start-of-step positions).
//
// The reference's quirks are kept literally: both area-force components use (im1 - ip1)
// (:38-41), the repulsion's point-in-polygon wraps x with floor and y with round and only
// when |d| > L (:178-187), the repulsion force does not depend on which cell
// contains the vertex (:206-218).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <chrono>
#include <vector>

#include "dpm_common.cuh"

using namespace dpm;

namespace dpm {

struct Step2DParams {
  const float2 *__restrict__ pos_in;
  float2 *__restrict__ pos_out;
  float2 *__restrict__ force_out;
  const float4 *__restrict__ bnd_in;
  float4 *__restrict__ bnd_out;
  const int *__restrict__ nv;
  const float4 *__restrict__ cellA;  // (Ka, Kl, Kb, a0)
  const float4 *__restrict__ cellB;  // (l0, r0, 0, 0)
  const int *__restrict__ cand_count;
  const int *__restrict__ cand;
  int K;
  const float4 *__restrict__ bbox_lo;
  const float4 *__restrict__ bbox_hi;
  NbrState *st;
  int nc, S;
  float dt, Kre, Kat;
  int pbc;
  float L;
  unsigned mask;
};

constexpr int T2D = 128;
constexpr int W2D = T2D / 32;

// One edge (V[j] -> V[i]) of the literal even-odd test of RepulsionForceUpdate (:166-196): does it toggle "overlaps"?
// `far` = a |d| > L wrap (the reference's floor/round quirk, :178-187) can fire for this vertex/cell pair at all.
__device__ __forceinline__ bool edge_toggles(float2 p, float2 vi, float2 vj, bool far, float L) {
  float dix = p.x - vi.x, diy = p.y - vi.y, djx = p.x - vj.x, djy = p.y - vj.y;
  if (far) {
    if (fabsf(dix) > L || fabsf(djx) > L) {
      dix = __fsub_rn(dix, __fmul_rn(L, floorf(__fdiv_rn(dix, L))));
      djx = __fsub_rn(djx, __fmul_rn(L, floorf(__fdiv_rn(djx, L))));
    }
    if (fabsf(diy) > L || fabsf(djy) > L) {
      diy = __fsub_rn(diy, __fmul_rn(L, roundf(__fdiv_rn(diy, L))));
      djy = __fsub_rn(djy, __fmul_rn(L, roundf(__fdiv_rn(djy, L))));
    }
  }
  if ((diy > 0.0f) == (djy > 0.0f)) return false;
  // 0 < (djx - dix) * (0 - diy) / (djy - diy) + dix   (:191-194).  The IEEE division (a ~15-instruction subroutine) only
  // decides the sign when the sum nearly cancels: a reciprocal-based quotient within 4 ulp settles every other case.
  const float num = __fmul_rn(__fsub_rn(djx, dix), __fsub_rn(0.0f, diy)), den = __fsub_rn(djy, diy);
  const float qa = num * __frcp_rn(den);
  const float sa = qa + dix;
  if (fabsf(sa) > 1e-5f * (fabsf(qa) + fabsf(dix))) return 0.0f < sa;
  const float xc = __fadd_rn(__fdiv_rn(num, den), dix);
  return 0.0f < xc;
}

// The same test for a vertex that no |d| > L wrap can reach, on an edge held in registers: the straddle condition by direct
// comparison (x - y > 0 <=> x > y in IEEE arithmetic with gradual underflow), the crossing only for the edges that straddle.
__device__ __forceinline__ bool edge_toggles_plain(float2 p, float2 vi, float2 vj) {
  if ((p.y > vi.y) == (p.y > vj.y)) return false;
  const float dix = p.x - vi.x, diy = p.y - vi.y, djx = p.x - vj.x, djy = p.y - vj.y;
  const float num = __fmul_rn(__fsub_rn(djx, dix), __fsub_rn(0.0f, diy)), den = __fsub_rn(djy, diy);
  const float qa = num * __frcp_rn(den);
  const float sa = qa + dix;
  if (fabsf(sa) > 1e-5f * (fabsf(qa) + fabsf(dix))) return 0.0f < sa;
  return 0.0f < __fadd_rn(__fdiv_rn(num, den), dix);
}

// One warp per cell.  Shape forces are per-lane ring stencils; the two contact terms are evaluated WARP-COOPERATIVELY:
// for every (vertex, candidate cell) pair that survives the exact culls, the 32 lanes split the candidate's ring —
// edges of the even-odd test (parity of the ballots) and vertices of the attraction sum (nonzero terms folded in
// ascending vertex order, the reference's order) — instead of each lane looping over the whole ring on its own.
__global__ void __launch_bounds__(T2D, 7) dpm2d_step_kernel(Step2DParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ci = blockIdx.x * W2D + warp;
  griddep_launch();  // the next timestep's rebuild kernel may be scheduled (it waits for this grid's completion first)
  if (ci >= P.nc) return;  // whole warp leaves; no block-level barriers below
  const int S = P.S;
  // per warp: own ring, staged neighbour ring, force accumulators, found flags
  float2 *sV = reinterpret_cast<float2 *>(smem_raw) + (size_t)warp * 3 * S;
  float2 *sN = sV + S;
  float2 *sF = sN + S;
  unsigned char *sFound = reinterpret_cast<unsigned char *>(reinterpret_cast<float2 *>(smem_raw) + (size_t)W2D * 3 * S) + (size_t)warp * S;
  // the candidate's vertices near this cell (attraction), ascending vertex order; after the found flags, 8-byte aligned
  float2 *sNear = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(reinterpret_cast<float2 *>(smem_raw) + (size_t)W2D * 3 * S) +
                                             (((size_t)W2D * S + 7) & ~(size_t)7)) + (size_t)warp * S;
  const int n = P.nv[ci];
  const float4 cA = P.cellA[ci], cB = P.cellB[ci];
  const float Ka = cA.x, Kl = cA.y, Kb = cA.z, a0 = cA.w, l0 = cB.x, r0 = cB.y;
  const float2 *gP = P.pos_in + (size_t)ci * S;
  for (int v = lane; v < n; v += 32) sV[v] = gP[v];
  __syncwarp();

  // ---- per-cell scalars in the reference's serial order (every lane computes the same chain) ----
  float area = 0.0f, sx = 0.0f, sy = 0.0f;
  {
    float2 prev = sV[n - 1];
    for (int vj = 0; vj < n; vj++) {
      const float2 cur = sV[vj];
      // Area += 0.5 * ((x[j-1] + x[j]) * (y[j-1] - y[j]))   (:27-32)
      area = __fadd_rn(area, __fmul_rn(0.5f, __fmul_rn(__fadd_rn(prev.x, cur.x), __fsub_rn(prev.y, cur.y))));
      sx = __fadd_rn(sx, cur.x);  // GetCOM :3-11
      sy = __fadd_rn(sy, cur.y);
      prev = cur;
    }
  }
  if (area < 0.0f) area = -area;
  const float strain = __fsub_rn(__fdiv_rn(area, a0), 1.0f);  // :37
  const float comx = __fdiv_rn(sx, (float)n), comy = __fdiv_rn(sy, (float)n);

  // ---- shape forces: area (:38-41), perimeter (:109-118), bending (:73-89) -------------------------
  for (int vi = lane; vi < n; vi += 32) {
    const int im1 = (vi == 0) ? n - 1 : vi - 1, ip1 = (vi == n - 1) ? 0 : vi + 1;
    const int ip2 = (ip1 == n - 1) ? 0 : ip1 + 1, im2 = (im1 == 0) ? n - 1 : im1 - 1;
    const float2 p = sV[vi], pm1 = sV[im1], pp1 = sV[ip1], pm2 = sV[im2], pp2 = sV[ip2];
    float fx = 0.0f, fy = 0.0f;
    if (P.mask & DPM2D_AREA) {
      const float c = (Ka / sqrtf(a0)) * 0.5f * strain;  // both components use im1 - ip1, sic
      fx += c * (pm1.y - pp1.y);
      fy += c * (pm1.x - pp1.x);
    }
    const float lvx = pp1.x - p.x, lvy = pp1.y - p.y, lmx = p.x - pm1.x, lmy = p.y - pm1.y;
    if (P.mask & DPM2D_PERIMETER) {
      // edge strain len/l0 - 1 cancels to ~1e-2: keep the squared length unfused so it rounds like the reference's dot()
      const float len = sqrtf(__fadd_rn(__fmul_rn(lvx, lvx), __fmul_rn(lvy, lvy)));
      const float lenm = sqrtf(__fadd_rn(__fmul_rn(lmx, lmx), __fmul_rn(lmy, lmy)));
      const float dli = len / l0 - 1.0f, dlim1 = lenm / l0 - 1.0f;
      const float k = Kl * sqrtf(a0 / l0);
      fx += k * (dli * (lvx / len) - dlim1 * (lmx / lenm));
      fy += k * (dli * (lvy / len) - dlim1 * (lmy / lenm));
    }
    if (P.mask & DPM2D_BENDING) {
      const float six = lvx - lmx, siy = lvy - lmy;
      const float sixp = (pp2.x - pp1.x) - lvx, siyp = (pp2.y - pp1.y) - lvy;
      const float sixm = lmx - (pm1.x - pm2.x), siym = lmy - (pm1.y - pm2.y);
      fx += Kb * (2.0f * six - sixm - sixp);
      fy += Kb * (2.0f * siy - siym - siyp);
    }
    sF[vi] = make_float2(fx, fy);
    sFound[vi] = 0;
  }
  __syncwarp();

  // ---- contacts: attraction (:250-267) and repulsion (:163-202) over the candidate cells, ascending id -----
  const int nchunk = (n + 31) >> 5;
  griddep_wait();  // the candidate lists belong to the rebuild kernel ahead; everything above only needs the previous step's state
  const int ncand = min(P.cand_count[ci], P.K);
  const bool doAtt = (P.mask & DPM2D_ATTRACT) && (P.Kat != 0.0f);  // Kat == 0 adds exact zeros in the reference
  const bool doRep = (P.mask & DPM2D_REPEL);
  const float halfL = 0.5f * P.L, invL = 1.0f / P.L;
  // own bounding box of the current positions (exact; written by the previous step's epilogue / the bounds kernel)
  const float4 bi0 = P.bnd_in[3 * (size_t)ci], bi1 = P.bnd_in[3 * (size_t)ci + 1];
  const float hxi = 0.5f * (bi1.x - bi0.x), hyi = 0.5f * (bi1.y - bi0.y), cxi = 0.5f * (bi0.x + bi1.x), cyi = 0.5f * (bi0.y + bi1.y);
  const bool own_cull_ok = !P.pbc || ((hxi + l0 < 0.25f * P.L) && (hyi + l0 < 0.25f * P.L));
  unsigned evals = 0;
  if (doAtt || doRep) {
    // Candidates are taken 32 at a time: every lane loads ONE candidate's id, bounding box and vertex count (one round of
    // global-memory latency for the whole batch instead of two dependent loads per candidate on the warp's critical path)
    // and decides, exactly, whether ANY vertex of this cell can pass the per-vertex culls for it; the survivors are then
    // visited in ascending order (the reference's summation order) with their data broadcast out of the owning lane.
    for (int cbase = 0; cbase < ncand; cbase += 32) {
     const int kk = cbase + lane;
     int my_cj = -1, my_nj = 0;
     float4 mb0 = make_float4(0.f, 0.f, 0.f, 0.f), mb1 = mb0;
     bool pass = false;
     if (kk < ncand) {
       my_cj = P.cand[(size_t)ci * P.K + kk];
       mb0 = P.bnd_in[3 * (size_t)my_cj]; mb1 = P.bnd_in[3 * (size_t)my_cj + 1];
       my_nj = P.nv[my_cj];
       const float hx = 0.5f * (mb1.x - mb0.x), hy = 0.5f * (mb1.y - mb0.y);
       const float cxj = 0.5f * (mb0.x + mb1.x), cyj = 0.5f * (mb0.y + mb1.y);
       const bool att_cull_ok = !P.pbc || ((hx + l0 < 0.25f * P.L) && (hy + l0 < 0.25f * P.L));
       bool repPossible = false, attPossible = false;
       if (doRep) {
         const bool ovx = !(bi1.x < mb0.x || bi0.x > mb1.x), ovy = !(bi1.y < mb0.y || bi0.y > mb1.y);
         const float mx = fmaxf(fmaxf(fabsf(bi0.x - mb0.x), fabsf(bi1.x - mb0.x)), fmaxf(fabsf(bi0.x - mb1.x), fabsf(bi1.x - mb1.x)));
         const float my = fmaxf(fmaxf(fabsf(bi0.y - mb0.y), fabsf(bi1.y - mb0.y)), fmaxf(fabsf(bi0.y - mb1.y), fabsf(bi1.y - mb1.y)));
         repPossible = (ovx || (P.pbc && mx > P.L)) && (ovy || (P.pbc && my > P.L));
       }
       if (doAtt) {
         float dx = cxi - cxj, dy = cyi - cyj;
         if (P.pbc) { dx -= P.L * roundf(dx * invL); dy -= P.L * roundf(dy * invL); }
         const float ax = fmaxf(fabsf(dx) - hx - hxi, 0.0f), ay = fmaxf(fabsf(dy) - hy - hyi, 0.0f);
         attPossible = !att_cull_ok || !own_cull_ok || (ax * ax + ay * ay <= l0 * l0 * 1.001f + 1e-12f);
       }
       pass = repPossible || attPossible;
     }
     unsigned mcand = __ballot_sync(0xffffffffu, pass);
     while (mcand) {
      const int csrc = __ffs(mcand) - 1;
      mcand &= mcand - 1;
      const int cj = __shfl_sync(0xffffffffu, my_cj, csrc);
      const int nj_known = __shfl_sync(0xffffffffu, my_nj, csrc);
      float4 bj0, bj1;
      bj0.x = __shfl_sync(0xffffffffu, mb0.x, csrc); bj0.y = __shfl_sync(0xffffffffu, mb0.y, csrc);
      bj1.x = __shfl_sync(0xffffffffu, mb1.x, csrc); bj1.y = __shfl_sync(0xffffffffu, mb1.y, csrc);
      const float hx = 0.5f * (bj1.x - bj0.x), hy = 0.5f * (bj1.y - bj0.y);
      const float cxj = 0.5f * (bj0.x + bj1.x), cyj = 0.5f * (bj0.y + bj1.y);
      const bool att_cull_ok = !P.pbc || ((hx + l0 < 0.25f * P.L) && (hy + l0 < 0.25f * P.L));
      // rij -= L * round(rij / L) (:254-256) is a no-op for every vertex pair of the two cells unless they can be half a box apart
      const bool wrapPossible = P.pbc && (fmaxf(fabsf(bi0.x - bj1.x), fabsf(bi1.x - bj0.x)) > halfL || fmaxf(fabsf(bi0.y - bj1.y), fabsf(bi1.y - bj0.y)) > halfL);
      int nj = 0, nnear = 0;
      bool staged = false, nearBuilt = false;
      // the candidate's edges i = lane and i = lane + 32 (rings of up to 64 vertices), loaded once per candidate when the first
      // vertex asks for the even-odd test: no shared-memory gathers or address arithmetic per tested vertex
      bool edgesLoaded = false, va = false, vb = false;
      float2 ea_i = make_float2(0.f, 0.f), ea_j = ea_i, eb_i = ea_i, eb_j = ea_i;
      for (int ch = 0; ch < nchunk; ch++) {
        const int vi = lane + 32 * ch;
        const bool act = vi < n;
        const float2 p = sV[act ? vi : 0];
        // per-vertex culls (exact)
        bool wantRep = false, far = false, wantAtt = false;
        if (act && doRep && !sFound[vi]) {
          const float dxl = p.x - bj0.x, dxh = p.x - bj1.x, dyl = p.y - bj0.y, dyh = p.y - bj1.y;
          const bool inx = dxl >= 0.0f && dxh <= 0.0f, iny = dyl >= 0.0f && dyh <= 0.0f;
          const bool farx = P.pbc && (fabsf(dxl) > P.L || fabsf(dxh) > P.L), fary = P.pbc && (fabsf(dyl) > P.L || fabsf(dyh) > P.L);
          wantRep = (inx || farx) && (iny || fary);
          far = farx || fary;
        }
        if (act && doAtt) {
          float dx = p.x - cxj, dy = p.y - cyj;
          // cull only: the quotient by multiplication; its rounding can differ from dx / L only half a box away from the
          // candidate, where either image is far outside l0 (att_cull_ok bounds the box size)
          if (P.pbc) { dx -= P.L * roundf(dx * invL); dy -= P.L * roundf(dy * invL); }
          const float ax = fmaxf(fabsf(dx) - hx, 0.0f), ay = fmaxf(fabsf(dy) - hy, 0.0f);
          wantAtt = !att_cull_ok || (ax * ax + ay * ay <= l0 * l0 * 1.0001f + 1e-12f);
        }
        unsigned mAtt = __ballot_sync(0xffffffffu, wantAtt), mRep = __ballot_sync(0xffffffffu, wantRep);
        const unsigned mFar = __ballot_sync(0xffffffffu, far);
        if ((mAtt | mRep) == 0) continue;
        if (!staged) {  // stage the neighbour's ring once per candidate
          nj = nj_known;
          const float2 *gN = P.pos_in + (size_t)cj * S;
          __syncwarp();
          for (int v = lane; v < nj; v += 32) sN[v] = gN[v];
          __syncwarp();
          staged = true;
        }
        // attraction: a neighbour vertex can only act on this cell if it lies within l0 of the cell's own bounding box
        // (exact: dist < l0[ci] is the reference's test, :258-262), so the neighbour's ring is first compacted to those
        // vertices, in ascending order; then every wanting vertex of the chunk (one per lane) folds the compacted list
        // serially — ascending vj within ascending cj, the reference's summation order — out of broadcast reads.
        if (mAtt) {
          if (!nearBuilt) {
            nnear = 0;
            for (int base = 0; base < nj; base += 32) {
              const int vj = base + lane;
              bool near = false;
              float2 q = make_float2(0.f, 0.f);
              if (vj < nj) {
                q = sN[vj];
                float dx = q.x - cxi, dy = q.y - cyi;
                if (P.pbc) { dx -= P.L * roundf(dx * invL); dy -= P.L * roundf(dy * invL); }  // cull only, see above
                const float ax = fmaxf(fabsf(dx) - hxi, 0.0f), ay = fmaxf(fabsf(dy) - hyi, 0.0f);
                near = !own_cull_ok || (ax * ax + ay * ay <= l0 * l0 * 1.0001f + 1e-12f);
              }
              const unsigned nb = __ballot_sync(0xffffffffu, near);
              if (near) sNear[nnear + __popc(nb & ((1u << lane) - 1u))] = q;
              nnear += __popc(nb);
            }
            __syncwarp();
            nearBuilt = true;
          }
          if (wantAtt && nnear > 0) {
            float2 f = sF[vi];
            // Forces[index] += (Kat / NV * dist / l0) * normalize(rij)   (:263-264): dist cancels, the term is (Kat / NV / l0) * rij
            // (and 0 for rij = 0, as OpenCL's normalize(0) = 0 gives); evaluated in that form, a few ulp from the literal one.
            // The cutoff dist < l0 is the reference's own test: decided on the squared distance, with the sqrt only in the
            // narrow band where the two could disagree.
            const float c = P.Kat / (float)n / l0;
            const float l0sq = l0 * l0, l0sq_lo = l0sq * 0.99999f, l0sq_hi = l0sq * 1.00001f;
#pragma unroll 4
            for (int t = 0; t < nnear; t++) {
              const float2 q = sNear[t];
              float rx = q.x - p.x, ry = q.y - p.y;
              if (wrapPossible) {  // rij -= L * round(rij / L)  (:254-256); a no-op unless |r| > L/2
                if (fabsf(rx) > halfL) rx -= P.L * roundf(rx / P.L);
                if (fabsf(ry) > halfL) ry -= P.L * roundf(ry / P.L);
              }
              const float d2 = rx * rx + ry * ry;
              if (d2 < l0sq_hi) {
                if (d2 < l0sq_lo || sqrtf(d2) < l0) {  // :258-262
                  f.x += c * rx;
                  f.y += c * ry;
                }
              }
            }
            sF[vi] = f;
          }
          __syncwarp();
        }
        // repulsion: lanes split the edges of the even-odd test; parity of all toggles
        while (mRep) {
          const int src = __ffs(mRep) - 1;
          mRep &= mRep - 1;
          const float2 pp = make_float2(__shfl_sync(0xffffffffu, p.x, src), __shfl_sync(0xffffffffu, p.y, src));
          const bool ufar = (mFar >> src) & 1u;
          unsigned par = 0;
          if (!ufar && nj <= 64) {
            if (!edgesLoaded) {
              va = lane < nj; vb = lane + 32 < nj;
              ea_i = sN[va ? lane : 0]; ea_j = sN[va ? (lane == 0 ? nj - 1 : lane - 1) : 0];
              eb_i = sN[vb ? lane + 32 : 0]; eb_j = sN[vb ? lane + 31 : 0];
              edgesLoaded = true;
            }
            const bool tga = va && edge_toggles_plain(pp, ea_i, ea_j), tgb = vb && edge_toggles_plain(pp, eb_i, eb_j);
            par = __popc(__ballot_sync(0xffffffffu, tga)) ^ __popc(__ballot_sync(0xffffffffu, tgb));
          } else {
            for (int base = 0; base < nj; base += 32) {
              const int i = base + lane;
              bool tg = false;
              if (i < nj) tg = edge_toggles(pp, sN[i], sN[i == 0 ? nj - 1 : i - 1], ufar, P.L);
              par ^= __popc(__ballot_sync(0xffffffffu, tg));
            }
          }
          if (lane == 0) sFound[src + 32 * ch] = (unsigned char)(par & 1u);
          evals++;
        }
        __syncwarp();
      }
     }
    }
  }
  __syncwarp();

  // ---- repulsion force (:204-218), Euler (:279), bounds -----------------------------------------------
  float lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
  for (int vi = lane; vi < n; vi += 32) {
    const float2 p = sV[vi];
    float2 f = sF[vi];
    if (sFound[vi]) {
      float dx = comx - p.x, dy = comy - p.y;
      if (P.pbc) { dx -= P.L * roundf(dx / P.L); dy -= P.L * roundf(dy / P.L); }
      const float dist = sqrtf(dx * dx + dy * dy);
      const float xij = dist / (2.0f * r0);
      const float ftmp = P.Kre * (1.0f - xij);
      if (dist != 0.0f) {  // normalize(0) = 0
        f.x += 0.5f * ftmp * (dx / dist);
        f.y += 0.5f * ftmp * (dy / dist);
      }
    }
    const float2 np = make_float2(p.x + f.x * P.dt, p.y + f.y * P.dt);
    P.pos_out[(size_t)ci * S + vi] = np;
    if (P.force_out) P.force_out[(size_t)ci * S + vi] = f;
    lo[0] = fminf(lo[0], np.x); lo[1] = fminf(lo[1], np.y);
    hi[0] = fmaxf(hi[0], np.x); hi[1] = fmaxf(hi[1], np.y);
  }
  for (int d = 0; d < 2; d++) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
  if (lane == 0) {
    P.bnd_out[3 * (size_t)ci + 0] = make_float4(lo[0], lo[1], 0.f, 0.f);
    P.bnd_out[3 * (size_t)ci + 1] = make_float4(hi[0], hi[1], 0.f, 0.f);
    P.bnd_out[3 * (size_t)ci + 2] = make_float4(comx, comy, area, 0.f);
    const float4 bl = P.bbox_lo[ci], bh = P.bbox_hi[ci];
    if (lo[0] < bl.x || lo[1] < bl.y || hi[0] > bh.x || hi[1] > bh.y) P.st->rebuild = 1;
    if (evals) atomicAdd(&P.st->contact_evals, (unsigned long long)evals);
  }
}

// bounds of a freshly uploaded position array; one warp per cell
__global__ void dpm2d_bounds_kernel(const float2 *pos, const int *nv, float4 *bnd, int nc, int S) {
  const int lane = threadIdx.x & 31, ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ci >= nc) return;
  float lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
  for (int v = lane; v < nv[ci]; v += 32) {
    const float2 p = pos[(size_t)ci * S + v];
    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y);
    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y);
  }
  for (int d = 0; d < 2; d++) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
  if (lane == 0) {
    bnd[3 * (size_t)ci + 0] = make_float4(lo[0], lo[1], 0.f, 0.f);
    bnd[3 * (size_t)ci + 1] = make_float4(hi[0], hi[1], 0.f, 0.f);
    bnd[3 * (size_t)ci + 2] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace dpm

struct dpm2d_ctx {
  int device = 0, nc = 0, S = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  float2 *pos[2] = {nullptr, nullptr};
  float2 *force = nullptr;
  float4 *bnd[2] = {nullptr, nullptr};
  int *nv = nullptr;
  float4 *cellA = nullptr, *cellB = nullptr;
  NbrState *st = nullptr;
  float4 *bbox_lo = nullptr, *bbox_hi = nullptr;
  int *bin_id = nullptr, *order = nullptr, *bin_count = nullptr, *bin_start = nullptr, *cand_count = nullptr, *cand = nullptr;
  float *partial = nullptr;
  int *chunk_sum = nullptr;
  int *ext_list = nullptr;
  int cap = 0, K = 32, K_alloc = 0;
  float skin_rel = 0.1f;
  int coop_grid = 0;
  int cur = 0;
  unsigned mask = DPM2D_ALL;
  bool uploaded = false;
  float l0max = 0.0f;
  float last_range = -1.0f, last_L = -1.0f;
  int last_pbc = -1;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float4 *h_cell = nullptr;
  dpm_stats_t stats{};
};

namespace {
// per warp: own ring, staged neighbour ring, force accumulators (float2 each) + found flags (bytes)
inline size_t smem2d_bytes(int S) { return (sizeof(float2) * 4 + 1) * (size_t)S * W2D + 32; }

struct DeviceGuard2 {
  int prev = -1;
  explicit DeviceGuard2(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard2() { if (prev >= 0) cudaSetDevice(prev); }
};

NbrBuffers nbr_buffers2(dpm2d_ctx *h, float range, int pbc, float L) {
  NbrBuffers nb{};
  nb.st = h->st;
  nb.blo = h->bnd[h->cur]; nb.bhi = h->bnd[h->cur] + 1; nb.blo_stride = 3;
  nb.bbox_lo = h->bbox_lo; nb.bbox_hi = h->bbox_hi;
  nb.bin_id = h->bin_id; nb.order = h->order; nb.bin_count = h->bin_count; nb.bin_start = h->bin_start;
  nb.cand_count = h->cand_count; nb.cand = h->cand; nb.partial = h->partial; nb.chunk_sum = h->chunk_sum;
  nb.nc = h->nc; nb.nc_list = h->nc; nb.nd = 2; nb.cap = h->cap; nb.K = h->K;
  nb.pbc = pbc; nb.L = L; nb.skin_rel = h->skin_rel; nb.range = range; nb.far2d = 1; nb.ext_list = h->ext_list;
  return nb;
}

int alloc_cand2(dpm2d_ctx *h) {
  if (h->K_alloc >= h->K) return DPM_OK;
  if (h->cand) cudaFree(h->cand);
  h->cand = nullptr;
  DPM_CUDA_TRY(cudaMalloc(&h->cand, sizeof(int) * (size_t)h->nc * h->K));
  h->K_alloc = h->K;
  return DPM_OK;
}

int mark_rebuild(dpm2d_ctx *h) {
  static const int one = 1;
  DPM_CUDA_TRY(cudaMemcpyAsync(&h->st->rebuild, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  return DPM_OK;
}

int check_flags2(dpm2d_ctx *h) {
  NbrState st;
  DPM_CUDA_TRY(cudaMemcpyAsync(&st, h->st, sizeof(NbrState), cudaMemcpyDeviceToHost, h->stream));
  DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->stats.rebuilds = (uint64_t)st.nbuilds;
  h->stats.contact_evals = st.contact_evals;
  if (st.overflow) return fail(DPM_ERR_RUNTIME, "neighbour candidate list overflow: raise max_candidates (dpm2d_set_neighbor_params)");
  return DPM_OK;
}
}  // namespace

extern "C" {

int dpm2d_create(dpm2d_t **out, int device, int ncells, int max_nv) {
  if (!out) return fail(DPM_ERR_INVALID_ARGUMENT, "handle pointer is NULL");
  *out = nullptr;
  if (ncells <= 0 || max_nv < 3) return fail(DPM_ERR_INVALID_ARGUMENT, "need ncells > 0 and max_nv >= 3");
  if (max_nv > 2048) return fail(DPM_ERR_INVALID_ARGUMENT, "more than 2048 vertices per 2D cell is not supported");
  int ndev = 0;
  DPM_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DPM_ERR_CUDA, "no such CUDA device (there is no CPU fallback)");
  DeviceGuard2 guard(device);
  dpm2d_ctx *h = new dpm2d_ctx();
  h->device = device; h->nc = ncells; h->S = max_nv;
  auto bail = [&](int code) { dpm2d_destroy(h); return code; };
#define TRYB(expr)                                                                                    \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return bail(fail(DPM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
  } while (0)
  TRYB(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  TRYB(cudaEventCreate(&h->ev0));
  TRYB(cudaEventCreate(&h->ev1));
  const size_t nvert = (size_t)ncells * max_nv;
  TRYB(cudaMalloc(&h->pos[0], sizeof(float2) * nvert));
  TRYB(cudaMalloc(&h->pos[1], sizeof(float2) * nvert));
  TRYB(cudaMalloc(&h->force, sizeof(float2) * nvert));
  TRYB(cudaMemset(h->pos[0], 0, sizeof(float2) * nvert));
  TRYB(cudaMemset(h->pos[1], 0, sizeof(float2) * nvert));
  TRYB(cudaMemset(h->force, 0, sizeof(float2) * nvert));
  TRYB(cudaMalloc(&h->bnd[0], sizeof(float4) * 3 * ncells));
  TRYB(cudaMalloc(&h->bnd[1], sizeof(float4) * 3 * ncells));
  TRYB(cudaMalloc(&h->nv, sizeof(int) * ncells));
  TRYB(cudaMalloc(&h->cellA, sizeof(float4) * ncells));
  TRYB(cudaMalloc(&h->cellB, sizeof(float4) * ncells));
  TRYB(cudaMallocHost(&h->h_cell, sizeof(float4) * 2 * ncells));
  h->cap = 4 * ncells + 1024;
  TRYB(cudaMalloc(&h->st, sizeof(NbrState)));
  TRYB(cudaMemset(h->st, 0, sizeof(NbrState)));
  TRYB(cudaMalloc(&h->bbox_lo, sizeof(float4) * ncells));
  TRYB(cudaMalloc(&h->bbox_hi, sizeof(float4) * ncells));
  TRYB(cudaMalloc(&h->bin_id, sizeof(int) * ncells));
  TRYB(cudaMalloc(&h->order, sizeof(int) * ncells));
  TRYB(cudaMalloc(&h->bin_count, sizeof(int) * (h->cap + 1)));
  TRYB(cudaMalloc(&h->bin_start, sizeof(int) * (h->cap + 1)));
  TRYB(cudaMalloc(&h->cand_count, sizeof(int) * ncells));
  h->coop_grid = rebuild_max_grid(device);
  TRYB(cudaMalloc(&h->partial, sizeof(float) * 16 * h->coop_grid));
  TRYB(cudaMalloc(&h->chunk_sum, sizeof(int) * h->coop_grid));
  TRYB(cudaMalloc(&h->ext_list, sizeof(int) * h->nc));
  int rc = alloc_cand2(h);
  if (rc) return bail(rc);
  TRYB(cudaFuncSetAttribute(dpm2d_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2d_bytes(max_nv)));
#undef TRYB
  *out = h;
  return DPM_OK;
}

int dpm2d_destroy(dpm2d_t *h) {
  if (!h) return DPM_OK;
  DeviceGuard2 guard(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  void *ptrs[] = {h->pos[0], h->pos[1], h->force, h->bnd[0], h->bnd[1], h->nv, h->cellA, h->cellB, h->st, h->bbox_lo, h->bbox_hi,
                  h->bin_id, h->order, h->bin_count, h->bin_start, h->cand_count, h->cand, h->partial, h->chunk_sum, h->ext_list};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (h->h_cell) cudaFreeHost(h->h_cell);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return DPM_OK;
}

int dpm2d_set_stream(dpm2d_t *h, void *cuda_stream) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return DPM_OK;
}

int dpm2d_set_neighbor_params(dpm2d_t *h, float skin_rel, int max_candidates) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  if (!(skin_rel >= 0.0f) || max_candidates < 1 || max_candidates > 128)
    return fail(DPM_ERR_INVALID_ARGUMENT, "skin_rel must be >= 0 and 1 <= max_candidates <= 128");
  DeviceGuard2 guard(h->device);
  h->skin_rel = skin_rel;
  h->K = max_candidates;
  int rc = alloc_cand2(h);
  if (rc) return rc;
  h->uploaded = false;
  return DPM_OK;
}

int dpm2d_get_neighbor_params(dpm2d_t *h, float *skin_rel, int *max_candidates) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  if (skin_rel) *skin_rel = h->skin_rel;
  if (max_candidates) *max_candidates = h->K;
  return DPM_OK;
}

int dpm2d_set_force_mask(dpm2d_t *h, unsigned mask) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  h->mask = mask & DPM2D_ALL;
  return DPM_OK;
}

int dpm2d_upload(dpm2d_t *h, const float *verts2, const int32_t *nv, const float *Ka, const float *Kl, const float *Kb,
                 const float *a0, const float *l0, const float *r0) {
  if (!h || !verts2 || !nv || !Ka || !Kl || !Kb || !a0 || !l0 || !r0) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard2 guard(h->device);
  h->l0max = 0.0f;
  for (int c = 0; c < h->nc; c++) {
    if (nv[c] < 3 || nv[c] > h->S) return fail(DPM_ERR_INVALID_ARGUMENT, "each cell needs 3 <= NV <= max_nv");
    h->h_cell[c] = make_float4(Ka[c], Kl[c], Kb[c], a0[c]);
    h->h_cell[h->nc + c] = make_float4(l0[c], r0[c], 0.f, 0.f);
    h->l0max = std::max(h->l0max, l0[c]);
  }
  h->cur = 0;
  DPM_CUDA_TRY(cudaMemcpyAsync(h->pos[0], verts2, sizeof(float2) * (size_t)h->nc * h->S, cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(cudaMemcpyAsync(h->nv, nv, sizeof(int) * h->nc, cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(cudaMemcpyAsync(h->cellA, h->h_cell, sizeof(float4) * h->nc, cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(cudaMemcpyAsync(h->cellB, h->h_cell + h->nc, sizeof(float4) * h->nc, cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(cudaMemsetAsync(h->st, 0, sizeof(NbrState), h->stream));
  dpm2d_bounds_kernel<<<(h->nc + 3) / 4, 128, 0, h->stream>>>(h->pos[0], h->nv, h->bnd[0], h->nc, h->S);
  DPM_CUDA_TRY(cudaGetLastError());
  h->stats.launches += 1;
  h->stats.steps = 0; h->stats.rebuilds = 0; h->stats.contact_evals = 0;
  int rc = mark_rebuild(h);
  if (rc) return rc;
  DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->uploaded = true;
  h->last_range = -1.0f;
  return DPM_OK;
}

int dpm2d_rebuild_neighbors(dpm2d_t *h, float Kat, int pbc, float L) {
  if (!h || !h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "upload first");
  DeviceGuard2 guard(h->device);
  int rc = mark_rebuild(h);
  if (rc) return rc;
  const float range = (Kat != 0.0f) ? h->l0max : 0.0f;
  DPM_CUDA_TRY(launch_rebuild(nbr_buffers2(h, range, pbc, L), h->stream, h->coop_grid));
  h->stats.launches += 1;
  h->last_range = range; h->last_pbc = pbc; h->last_L = L;
  return DPM_OK;
}

int dpm2d_step(dpm2d_t *h, int nsteps, float dt, float Kre, float Kat, int pbc, float L) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  if (nsteps <= 0) return fail(DPM_ERR_INVALID_ARGUMENT, "nsteps must be positive");
  if (!h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "dpm2d_step before dpm2d_upload");
  DeviceGuard2 guard(h->device);
  const float range = ((h->mask & DPM2D_ATTRACT) && Kat != 0.0f) ? h->l0max : 0.0f;
  if (range != h->last_range || pbc != h->last_pbc || L != h->last_L) {
    int rc = mark_rebuild(h);  // the lists depend on the interaction range and the box
    if (rc) return rc;
    h->last_range = range; h->last_pbc = pbc; h->last_L = L;
  }
  Step2DParams p{};
  p.nv = h->nv; p.cellA = h->cellA; p.cellB = h->cellB;
  p.cand_count = h->cand_count; p.cand = h->cand; p.K = h->K;
  p.bbox_lo = h->bbox_lo; p.bbox_hi = h->bbox_hi; p.st = h->st;
  p.nc = h->nc; p.S = h->S; p.dt = dt; p.Kre = Kre; p.Kat = Kat; p.pbc = pbc; p.L = L; p.mask = h->mask;
  const size_t smem = smem2d_bytes(h->S);
  for (int s = 0; s < nsteps; s++) {
    DPM_CUDA_TRY(launch_rebuild(nbr_buffers2(h, range, pbc, L), h->stream, h->coop_grid));
    p.pos_in = h->pos[h->cur]; p.pos_out = h->pos[h->cur ^ 1];
    p.bnd_in = h->bnd[h->cur]; p.bnd_out = h->bnd[h->cur ^ 1];
    p.force_out = (s == nsteps - 1) ? h->force : nullptr;  // src/Tissue2D.cpp:223-227
    {  // programmatic dependent launch: ring staging, serial chains and shape forces overlap the rebuild kernel ahead
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)((h->nc + W2D - 1) / W2D)); cfg.blockDim = dim3(T2D); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      DPM_CUDA_TRY(cudaLaunchKernelEx(&cfg, dpm2d_step_kernel, p));
    }
    h->cur ^= 1;
    if ((s + 1) % 1000 == 0 && s + 1 < nsteps) {  // catch a capacity overflow early (the 3D reference drains its queue every 1000 steps)
      int rc = check_flags2(h);
      if (rc) return rc;
    }
  }
  h->stats.steps += (uint64_t)nsteps;
  h->stats.launches += 2ull * (uint64_t)nsteps;
  return DPM_OK;
}

int dpm2d_sync(dpm2d_t *h) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  DeviceGuard2 guard(h->device);
  return check_flags2(h);
}

int dpm2d_download(dpm2d_t *h, float *verts2, float *forces2) {
  if (!h || !h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "nothing to download");
  DeviceGuard2 guard(h->device);
  const size_t bytes = sizeof(float2) * (size_t)h->nc * h->S;
  if (verts2) DPM_CUDA_TRY(cudaMemcpyAsync(verts2, h->pos[h->cur], bytes, cudaMemcpyDeviceToHost, h->stream));
  if (forces2) DPM_CUDA_TRY(cudaMemcpyAsync(forces2, h->force, bytes, cudaMemcpyDeviceToHost, h->stream));
  return check_flags2(h);
}

int dpm2d_euler_update(dpm2d_t *h, float *verts2, float *forces2, const int32_t *nv, const float *Ka, const float *Kl,
                       const float *Kb, const float *a0, const float *l0, const float *r0, int nsteps, float dt, float Kre,
                       float Kat, int pbc, float L, float *loop_ms) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  if (nsteps <= 0) return fail(DPM_ERR_INVALID_ARGUMENT, "nsteps must be positive");
  DeviceGuard2 guard(h->device);
  static const bool trace = getenv("DPM_TRACE") != nullptr;  // host-side phase times of the call on stderr
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  auto ms_since = [&](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
  double t_up = 0, t_enq = 0, t_run = 0;
  for (int attempt = 0;; attempt++) {
    int rc = dpm2d_upload(h, verts2, nv, Ka, Kl, Kb, a0, l0, r0);
    if (rc) return rc;
    t_up = ms_since(t0);
    DPM_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    rc = dpm2d_step(h, nsteps, dt, Kre, Kat, pbc, L);
    if (rc) return rc;
    DPM_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    t_enq = ms_since(t0);
    rc = check_flags2(h);
    t_run = ms_since(t0);
    if (trace) {
      float dev_ms = 0.f;
      cudaEventElapsedTime(&dev_ms, h->ev0, h->ev1);
      fprintf(stderr, "[dpm2d] euler_update %d steps attempt %d (K=%d): upload %.3f ms, steps enqueued %.3f, steps done %.3f (device loop %.3f) rc=%d\n",
              nsteps, attempt, h->K, t_up, t_enq, t_run, dev_ms, rc);
    }
    if (rc == DPM_ERR_RUNTIME && h->K < 128 && attempt < 3) {
      int rc2 = dpm2d_set_neighbor_params(h, h->skin_rel, std::min(128, h->K * 2));
      if (rc2) return rc2;
      continue;
    }
    if (rc) return rc;
    break;
  }
  if (loop_ms) DPM_CUDA_TRY(cudaEventElapsedTime(loop_ms, h->ev0, h->ev1));
  return dpm2d_download(h, verts2, forces2);
}

int dpm2d_get_neighbor_artifacts(dpm2d_t *h, dpm_grid_t *grid, int32_t *bin_id, int32_t *order, int32_t *bin_start,
                                 int32_t *cand_count, int32_t *cand) {
  if (!h || !h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "no neighbour state");
  DeviceGuard2 guard(h->device);
  DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  NbrState st;
  DPM_CUDA_TRY(cudaMemcpy(&st, h->st, sizeof(NbrState), cudaMemcpyDeviceToHost));
  if (grid) *grid = st.grid;
  if (bin_id) DPM_CUDA_TRY(cudaMemcpy(bin_id, h->bin_id, sizeof(int) * h->nc, cudaMemcpyDeviceToHost));
  if (order) DPM_CUDA_TRY(cudaMemcpy(order, h->order, sizeof(int) * h->nc, cudaMemcpyDeviceToHost));
  if (bin_start) DPM_CUDA_TRY(cudaMemcpy(bin_start, h->bin_start, sizeof(int) * (st.grid.nbins + 1), cudaMemcpyDeviceToHost));
  if (cand_count) DPM_CUDA_TRY(cudaMemcpy(cand_count, h->cand_count, sizeof(int) * h->nc, cudaMemcpyDeviceToHost));
  if (cand) DPM_CUDA_TRY(cudaMemcpy(cand, h->cand, sizeof(int) * (size_t)h->nc * h->K, cudaMemcpyDeviceToHost));
  return DPM_OK;
}

int dpm2d_get_stats(dpm2d_t *h, dpm_stats_t *out) {
  if (!h || !out) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = h->stats;
  return DPM_OK;
}
int dpm2d_reset_stats(dpm2d_t *h) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  memset(&h->stats, 0, sizeof(h->stats));
  return DPM_OK;
}

}  // extern "C"
