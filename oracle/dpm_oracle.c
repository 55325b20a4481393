/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See dpm_oracle_impl.h.
 *
 * Builds liboracle_dpm.so:
 *   oracle{2d,3d}_*_f32  : fp32 restatement of the reference OpenCL kernels
 *   oracle{2d,3d}_*_f64  : same formulas in double ("truth")
 *   oracle_cell_list     : CPU restatement of the cell-list / candidate-list
 *                          specification (DESIGN.md §"Neighbour search"); the
 *                          GPU integer artefacts must match it bit-exactly.
 *   oracle_icosphere / oracle_cell3d_params / oracle_cell2d_init :
 *                          restatement of src/cell.cpp construction.
 * Compile with -ffp-contract=off (no FMA contraction) — see oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define FN(name) CAT(name, _f32)
#define RSQRT sqrtf
#define RABS fabsf
#define RROUND roundf
#define RFLOOR floorf
#define RATAN2 atan2f
#include "dpm_oracle_impl.h"
#undef REAL
#undef FN
#undef RSQRT
#undef RABS
#undef RROUND
#undef RFLOOR
#undef RATAN2

#define REAL double
#define FN(name) CAT(name, _f64)
#define RSQRT sqrt
#define RABS fabs
#define RROUND round
#define RFLOOR floor
#define RATAN2 atan2
#include "dpm_oracle_impl.h"

/* ======================================================================== */
/* Cell-list specification (fp32 + int32 only; every operation is a single   */
/* IEEE op so the GPU can reproduce it bit-exactly with __f*_rn intrinsics). */
/* ======================================================================== */

#define NBMAX 1024

typedef struct {
  int32_t nb[3];
  int32_t periodic[3]; /* bin index wraps on this axis */
  int32_t allpass[3];  /* PBC box too small for bins on this axis: no filtering */
  float origin[3];
  float inv_binw[3];
  float max_ext;
  float margin;
  int32_t nbins;
  int32_t pad;
} oracle_grid_t;

static inline float wrapc(float c, int PBC, float L) { return PBC ? c - L * floorf(c / L) : c; }

/* lo/hi: [nc][3] exact per-cell AABBs (unused axes d>=nd must be 0).
 * Outputs: grid, bin_id[nc], order[nc] (cells sorted by (bin, id)),
 * bin_start[cap+1], cand_count[nc] (true count, may exceed K), cand[nc*K] ascending.
 * far2d: for nd==2 && PBC also add the "|d| > L" partners. */
void oracle_cell_list(int nd, int nc, const float *lo, const float *hi, int PBC, float L, float skin_rel, float range,
                      int cap, int K, int far2d, oracle_grid_t *g, int32_t *bin_id, int32_t *order,
                      int32_t *bin_start, int32_t *cand_count, int32_t *cand) {
  float max_ext = 0.0f;
  for (int c = 0; c < nc; c++) for (int d = 0; d < nd; d++) {
    float e = hi[3 * c + d] - lo[3 * c + d];
    if (e > max_ext) max_ext = e;
  }
  float skin = skin_rel * max_ext;
  float margin = skin + range;
  float binw_req = max_ext + margin;
  g->max_ext = max_ext; g->margin = margin;
  for (int d = 0; d < 3; d++) { g->nb[d] = 1; g->periodic[d] = 0; g->allpass[d] = 0; g->origin[d] = 0.0f; g->inv_binw[d] = 0.0f; }
  for (int d = 0; d < nd; d++) {
    float gmin = 0, gmax = 0;
    for (int c = 0; c < nc; c++) {
      float cw = wrapc(0.5f * (lo[3 * c + d] + hi[3 * c + d]), PBC, L);
      if (c == 0 || cw < gmin) gmin = cw;
      if (c == 0 || cw > gmax) gmax = cw;
    }
    float span = gmax - gmin;
    if (PBC && (span + (binw_req + binw_req)) >= L) {
      int nbp = (int)floorf(L / binw_req);
      if (nbp > NBMAX) nbp = NBMAX;
      if (nbp >= 3) { g->nb[d] = nbp; g->periodic[d] = 1; g->inv_binw[d] = (float)nbp / L; }
      else { g->nb[d] = 1; g->allpass[d] = 1; }
    } else {
      int nb = (int)floorf(span / binw_req) + 1;
      if (nb > NBMAX) nb = NBMAX;
      g->nb[d] = nb; g->origin[d] = gmin; g->inv_binw[d] = 1.0f / binw_req;
      if (nb == NBMAX) g->inv_binw[d] = (float)NBMAX / (span + binw_req);
    }
  }
  while ((long)g->nb[0] * g->nb[1] * g->nb[2] > (long)cap) {
    int d = 0;
    if (g->nb[1] > g->nb[d]) d = 1;
    if (g->nb[2] > g->nb[d]) d = 2;
    if (g->periodic[d]) {
      int nb = g->nb[d] / 2;
      if (nb >= 3) { g->nb[d] = nb; g->inv_binw[d] = (float)nb / L; }
      else { g->nb[d] = 1; g->periodic[d] = 0; g->allpass[d] = 1; g->inv_binw[d] = 0.0f; }
    } else {
      g->nb[d] = (g->nb[d] + 1) / 2;
      g->inv_binw[d] = g->inv_binw[d] * 0.5f;
    }
  }
  g->nbins = g->nb[0] * g->nb[1] * g->nb[2];
  int32_t *ib3 = (int32_t *)malloc(sizeof(int32_t) * 3 * nc);
  for (int c = 0; c < nc; c++) {
    int id3[3] = {0, 0, 0};
    for (int d = 0; d < nd; d++) {
      float cw = wrapc(0.5f * (lo[3 * c + d] + hi[3 * c + d]), PBC, L);
      int ib = (int)floorf((cw - g->origin[d]) * g->inv_binw[d]);
      if (ib < 0) ib = 0;
      if (ib > g->nb[d] - 1) ib = g->nb[d] - 1;
      id3[d] = ib;
    }
    ib3[3 * c] = id3[0]; ib3[3 * c + 1] = id3[1]; ib3[3 * c + 2] = id3[2];
    bin_id[c] = (id3[2] * g->nb[1] + id3[1]) * g->nb[0] + id3[0];
  }
  /* stable counting sort by bin id == std::stable_sort on (bin, id) */
  for (int b = 0; b <= g->nbins; b++) bin_start[b] = 0;
  for (int c = 0; c < nc; c++) bin_start[bin_id[c] + 1]++;
  for (int b = 0; b < g->nbins; b++) bin_start[b + 1] += bin_start[b];
  int32_t *fill = (int32_t *)calloc((size_t)(g->nbins > 0 ? g->nbins : 1), sizeof(int32_t));
  for (int c = 0; c < nc; c++) { int b = bin_id[c]; order[bin_start[b] + fill[b]++] = c; }
  free(fill);
  /* candidates */
  int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * nc);
  for (int i = 0; i < nc; i++) {
    int n = 0;
    int offs[3][3], noff[3];
    for (int d = 0; d < 3; d++) {
      noff[d] = 0;
      int nb = g->nb[d], ib = ib3[3 * i + d];
      for (int o = -1; o <= 1; o++) {
        int b = ib + o;
        if (g->periodic[d]) b = (b + nb) % nb;
        else if (b < 0 || b >= nb) continue;
        int dup = 0;
        for (int q = 0; q < noff[d]; q++) if (offs[d][q] == b) dup = 1;
        if (!dup) offs[d][noff[d]++] = b;
      }
    }
    for (int z = 0; z < noff[2]; z++) for (int y = 0; y < noff[1]; y++) for (int x = 0; x < noff[0]; x++) {
      int b = (offs[2][z] * g->nb[1] + offs[1][y]) * g->nb[0] + offs[0][x];
      for (int s = bin_start[b]; s < bin_start[b + 1]; s++) {
        int j = order[s];
        if (j == i) continue;
        int ok = 1;
        for (int d = 0; d < nd; d++) {
          if (g->allpass[d]) continue;
          float ci_ = 0.5f * (lo[3 * i + d] + hi[3 * i + d]), cj_ = 0.5f * (lo[3 * j + d] + hi[3 * j + d]);
          float hi_ = 0.5f * (hi[3 * i + d] - lo[3 * i + d]), hj_ = 0.5f * (hi[3 * j + d] - lo[3 * j + d]);
          float dd = ci_ - cj_;
          if (PBC) dd = dd - L * roundf(dd / L);
          if (fabsf(dd) > (hi_ + hj_) + margin) { ok = 0; break; }
        }
        if (ok) tmp[n++] = j;
      }
    }
    if (far2d && nd == 2 && PBC) {
      float Lm = L - skin;
      for (int j = 0; j < nc; j++) {
        if (j == i) continue;
        /* The literal even-odd test can only report "inside" if, on EACH axis, the vertex lies within the
         * polygon's extent or a |d| > L wrap can fire on that axis (no y-straddle otherwise; all x-crossings on
         * one side otherwise).  Cell level, with the skin: per axis "overlap or far", and far on >= 1 axis. */
        int farx = (hi[3 * j] - lo[3 * i] > Lm) || (hi[3 * i] - lo[3 * j] > Lm);
        int fary = (hi[3 * j + 1] - lo[3 * i + 1] > Lm) || (hi[3 * i + 1] - lo[3 * j + 1] > Lm);
        if (!(farx || fary)) continue;
        int ovx = !((lo[3 * j] - hi[3 * i] > skin) || (lo[3 * i] - hi[3 * j] > skin));
        int ovy = !((lo[3 * j + 1] - hi[3 * i + 1] > skin) || (lo[3 * i + 1] - hi[3 * j + 1] > skin));
        if (!((ovx || farx) && (ovy || fary))) continue;
        int dup = 0;
        for (int q = 0; q < n; q++) if (tmp[q] == j) { dup = 1; break; }
        if (!dup) tmp[n++] = j;
      }
    }
    /* ascending insertion sort */
    for (int a = 1; a < n; a++) { int v = tmp[a], b = a - 1; while (b >= 0 && tmp[b] > v) { tmp[b + 1] = tmp[b]; b--; } tmp[b + 1] = v; }
    cand_count[i] = n;
    for (int a = 0; a < n && a < K; a++) cand[(size_t)i * K + a] = tmp[a];
  }
  free(tmp); free(ib3);
}

/* exact per-cell AABBs, 3D float4-strided verts */
void oracle_aabb3d(int nc, int nv, const float *verts, float *lo, float *hi) {
  for (int c = 0; c < nc; c++) for (int d = 0; d < 3; d++) {
    float l = verts[4 * ((size_t)c * nv) + d], h = l;
    for (int i = 1; i < nv; i++) { float x = verts[4 * ((size_t)c * nv + i) + d]; if (x < l) l = x; if (x > h) h = x; }
    lo[3 * c + d] = l; hi[3 * c + d] = h;
  }
}
void oracle_aabb2d(int nc, int S, const int32_t *NV, const float *verts, float *lo, float *hi) {
  for (int c = 0; c < nc; c++) {
    for (int d = 0; d < 2; d++) {
      float l = verts[2 * ((size_t)c * S) + d], h = l;
      for (int i = 1; i < NV[c]; i++) { float x = verts[2 * ((size_t)c * S + i) + d]; if (x < l) l = x; if (x > h) h = x; }
      lo[3 * c + d] = l; hi[3 * c + d] = h;
    }
    lo[3 * c + 2] = 0.0f; hi[3 * c + 2] = 0.0f;
  }
}

/* ======================================================================== */
/* Geometry / initial conditions: restatement of src/cell.cpp               */
/* ======================================================================== */

/* src/cell.cpp:160-194 AddMiddlePoint, including the "norm recomputed after each
 * component division" quirk */
static uint32_t add_mid(float *V, int *nv, int32_t *cache, int *ncache, uint32_t p1, uint32_t p2) {
  int key = (int)floor((double)((p1 + p2) * (p1 + p2 + 1) / 2)) + (int)(p1 < p2 ? p1 : p2);
  for (int i = 0; i < *ncache; i++) if (cache[2 * i] == key) return (uint32_t)cache[2 * i + 1];
  float mp[3];
  for (int i = 0; i < 3; i++) {
    mp[i] = V[3 * p2 + i] + V[3 * p1 + i];
    mp[i] = (float)((double)mp[i] * 0.5);
  }
  for (int i = 0; i < 3; i++) {
    float norm = sqrtf(mp[0] * mp[0] + mp[1] * mp[1] + mp[2] * mp[2]);
    mp[i] /= norm;
  }
  int idx = (*nv)++;
  V[3 * idx] = mp[0]; V[3 * idx + 1] = mp[1]; V[3 * idx + 2] = mp[2];
  cache[2 * *ncache] = key; cache[2 * *ncache + 1] = idx; (*ncache)++;
  return (uint32_t)idx;
}

/* src/cell.cpp:62-140. subdiv=2 is the reference (162/320). verts: [nv][3] unit-ish sphere
 * (before scaling by r0), faces [nf][3]. Returns nv. Caller sizes: nv = 10*4^s+2, nf = 20*4^s. */
int oracle_icosphere(int subdiv, float *V, uint32_t *F) {
  float t = (float)((1 + sqrt(5)) / 2);
  const float base[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                             {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
  int nv = 12;
  for (int i = 0; i < 12; i++) {
    float norm = sqrtf(base[i][0] * base[i][0] + base[i][1] * base[i][1] + base[i][2] * base[i][2]);
    for (int d = 0; d < 3; d++) V[3 * i + d] = base[i][d] / norm;
  }
  const uint32_t f0[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                              {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                              {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
  int nf = 20;
  memcpy(F, f0, sizeof f0);
  int nfmax = 20; for (int s = 0; s < subdiv; s++) nfmax *= 4;
  uint32_t *NF_ = (uint32_t *)malloc(sizeof(uint32_t) * 3 * nfmax);
  int32_t *cache = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)nfmax * 2);
  int ncache = 0;
  for (int s = 0; s < subdiv; s++) {
    int k = 0;
    for (int j = 0; j < nf; j++) {
      uint32_t a = add_mid(V, &nv, cache, &ncache, F[3 * j], F[3 * j + 1]);
      uint32_t b = add_mid(V, &nv, cache, &ncache, F[3 * j + 1], F[3 * j + 2]);
      uint32_t c = add_mid(V, &nv, cache, &ncache, F[3 * j + 2], F[3 * j]);
      uint32_t nw[4][3] = {{F[3 * j], a, c}, {F[3 * j + 1], b, a}, {F[3 * j + 2], c, b}, {a, b, c}};
      memcpy(NF_ + 3 * k, nw, sizeof nw); k += 4;
    }
    nf = k;
    memcpy(F, NF_, sizeof(uint32_t) * 3 * nf);
  }
  free(NF_); free(cache);
  return nv;
}

/* src/cell.cpp:142-157: scale/translate, v0, sa0, a0 ; src/Tissue3D.cpp:177: l0.
 * out = {v0, sa0, a0, l0} */
void oracle_cell3d_params(float calA, float r0, int nf, float *out) {
  float v0 = (float)((double)(4.0f / 3.0f) * M_PI * pow((double)r0, 3));
  float sa0 = (float)pow(6 * sqrt(M_PI) * (double)v0 * (double)calA, (double)(2.0f / 3.0f));
  float a0 = sa0 / (float)nf;
  float l0 = (float)(sqrt((double)(4.0f * a0)) / sqrt((double)3.0f));
  out[0] = v0; out[1] = sa0; out[2] = a0; out[3] = l0;
}
void oracle_cell3d_place(int nv, const float *unitV, float r0, const float *start, float *verts4 /*nv*4*/) {
  for (int i = 0; i < nv; i++) {
    for (int d = 0; d < 3; d++) { float x = unitV[3 * i + d]; x *= r0; x += start[d]; verts4[4 * i + d] = x; }
    verts4[4 * i + 3] = 0.0f;
  }
}

/* src/cell.cpp:12-46 Cell2D ctor + GetArea.  out = {calA0, a0, l0}; verts [NV][2] */
void oracle_cell2d_init(float x0, float y0, float calA, int NV, float r0, float *verts, float *out) {
  float calA0 = (float)((double)calA * ((double)NV * tan(M_PI / NV) / M_PI));
  for (int i = 0; i < NV; i++) {
    verts[2 * i] = (float)((double)r0 * cos(2.0 * M_PI * (i + 1.0) / (double)(float)NV) + (double)x0);
    verts[2 * i + 1] = (float)((double)r0 * sin(2.0 * M_PI * (i + 1.0) / (double)(float)NV) + (double)y0);
  }
  float Area = 0.0f;
  int j = NV - 1;
  for (int i = 0; i < NV; i++) {
    Area = (float)((double)Area + 0.5 * (double)((verts[2 * j] + verts[2 * i]) * (verts[2 * j + 1] - verts[2 * i + 1])));
    j = i;
  }
  if (Area < 0.0f) Area = -Area;
  float l0 = (float)(2.0 * sqrt(M_PI * (double)calA0 * (double)Area) / (double)(float)NV);
  out[0] = calA0; out[1] = Area; out[2] = l0;
}
