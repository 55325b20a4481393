/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("port" oracle) of the per-timestep hot path of
 * sudo-shaka/OpenCL_DPM.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 *
 * This header is included twice by dpm_oracle.c, once with REAL=float
 * (the parity oracle: same precision and operation order as the OpenCL
 * kernels, serial summation in index order) and once with REAL=double
 * (a "truth" trajectory used to separate chaos from arithmetic error).
 *
 * Every function cites the reference file:line it restates
 * (paths relative to /root/reference).  Unsuffixed literals that promote an
 * expression to double in OpenCL C are kept as double here (marked "->dbl").
 *
 * Pinning status: the reference ships no golden vectors for these kernels
 * (SURVEY.md §4).  The restatement is pinned against the reference's own
 * OpenCL kernels executed on the GPU box through oracle/_ref (see
 * oracle/README.md and tests/golden/); geometry is pinned against the
 * reference's src/cell.cpp compiled into oracle/_ref.
 */

#ifndef REAL
#error "include from dpm_oracle.c"
#endif

/* ---- small helpers ------------------------------------------------------ */

static inline REAL FN(dot3)(const REAL *a, const REAL *b) {
  /* OpenCL dot(): x*x + y*y + z*z, left to right */
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline void FN(cross3)(const REAL *a, const REAL *b, REAL *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
/* OpenCL normalize(): a zero vector is returned unchanged (measured on the reference's own runtime, NVIDIA OpenCL on the
 * B200 box: oracle/cl_semantics_probe.cpp, profiles/r01_cl_semantics.log) -- reached when a vertex coincides with a vertex
 * of another cell (attraction, solid angle). */
static inline void FN(normalize3)(const REAL *a, REAL *u) {
  REAL n = RSQRT(FN(dot3)(a, a));
  if (n == (REAL)0) { u[0] = u[1] = u[2] = (REAL)0; return; }
  u[0] = a[0] / n; u[1] = a[1] / n; u[2] = a[2] / n;
}

/* shaders/Cell3D_Kernel.cl:35-44  GetCOM: serial sum in index order, then
 * multiply by 1/(float)NV. */
static void FN(com3d)(const REAL *V /*nv*4*/, int nv, REAL *com) {
  REAL s[3] = {0, 0, 0};
  for (int i = 0; i < nv; i++) { s[0] += V[4 * i]; s[1] += V[4 * i + 1]; s[2] += V[4 * i + 2]; }
  REAL inv = (REAL)1.0 / (REAL)nv;
  com[0] = s[0] * inv; com[1] = s[1] * inv; com[2] = s[2] * inv;
}

/* shaders/Cell3D_Kernel.cl:46-64  getVolume */
static REAL FN(volume3d)(const REAL *V, const uint32_t *F, int nf) {
  REAL vol = 0;
  for (int fi = 0; fi < nf; fi++) {
    const REAL *P0 = V + 4 * F[3 * fi], *P1 = V + 4 * F[3 * fi + 1], *P2 = V + 4 * F[3 * fi + 2];
    REAL c[3]; FN(cross3)(P0, P1, c);
    vol += FN(dot3)(c, P2) / (REAL)6.0;
  }
  return RABS(vol);
}

/* ---- 3D kernels, one cell at a time -------------------------------------- */

/* shaders/Cell3D_Kernel.cl:66-112  VolumeForceUpdate (volume = current
 * positions). */
/* stale_from >= 0 emulates the work-group race of the reference as it resolves on NVIDIA's OpenCL:
 * the runtime picks a local size of nf/2 = 160, the barrier is work-group scoped, so faces >= stale_from read
 * cellVolumes[ci] BEFORE work-item 0 of the other group writes it: the previous step's volume (0 on the first
 * step of a call, src/Tissue3D.cpp:180).  stale_from < 0: intended semantics (current volume for every face). */
static void FN(volume_force3d)(const REAL *V, REAL *Fo, const uint32_t *F, int nv, int nf, REAL Kv, REAL v0,
                               int stale_from, REAL *prev_volume) {
  if (Kv == (REAL)0.0) return;
  REAL volume = FN(volume3d)(V, F, nf);
  REAL strain_fresh = (volume / v0) - (REAL)1.0;
  REAL strain_stale = strain_fresh;
  if (stale_from >= 0 && prev_volume) { strain_stale = (*prev_volume / v0) - (REAL)1.0; *prev_volume = volume; }
  REAL com[3]; FN(com3d)(V, nv, com);
  for (int fi = 0; fi < nf; fi++) {
    uint32_t i0 = F[3 * fi], i1 = F[3 * fi + 1], i2 = F[3 * fi + 2];
    REAL A[3], B[3], C[3], g0[3], g1[3], g2[3];
    for (int d = 0; d < 3; d++) {
      A[d] = V[4 * i1 + d] - com[d];
      B[d] = V[4 * i2 + d] - com[d];
      C[d] = V[4 * i0 + d] - com[d];
    }
    FN(cross3)(A, B, g0); FN(cross3)(B, C, g1); FN(cross3)(C, A, g2);
    REAL strain = (stale_from >= 0 && fi >= stale_from) ? strain_stale : strain_fresh;
    REAL coef = -Kv * strain; /* (-Kv*strain)*grad/6 : :106-108 */
    for (int d = 0; d < 3; d++) {
      Fo[4 * i0 + d] += coef * g0[d] / (REAL)6.0;
      Fo[4 * i1 + d] += coef * g1[d] / (REAL)6.0;
      Fo[4 * i2 + d] += coef * g2[d] / (REAL)6.0;
    }
  }
}

/* shaders/Cell3D_Kernel.cl:114-177  SurfaceAreaForceUpdate (edge springs) */
static void FN(area_force3d)(const REAL *V, REAL *Fo, const uint32_t *F, int nf, REAL Ka, REAL a0, REAL l0) {
  if (Ka < (REAL)1e-8f) return;
  for (int fi = 0; fi < nf; fi++) {
    uint32_t i0 = F[3 * fi], i1 = F[3 * fi + 1], i2 = F[3 * fi + 2];
    REAL lv0[3], lv1[3], lv2[3];
    for (int d = 0; d < 3; d++) {
      lv0[d] = V[4 * i1 + d] - V[4 * i0 + d];
      lv1[d] = V[4 * i2 + d] - V[4 * i1 + d];
      lv2[d] = V[4 * i0 + d] - V[4 * i2 + d];
    }
    REAL len0 = RSQRT(FN(dot3)(lv0, lv0)), len1 = RSQRT(FN(dot3)(lv1, lv1)), len2 = RSQRT(FN(dot3)(lv2, lv2));
    if (len0 < (REAL)1e-12f || len1 < (REAL)1e-12f || len2 < (REAL)1e-12f) continue;
    REAL n0[3], n1[3], n2[3];
    for (int d = 0; d < 3; d++) { n0[d] = lv0[d] / len0; n1[d] = lv1[d] / len1; n2[d] = lv2[d] / len2; }
    REAL dl0 = (len0 / l0) - (REAL)1.0, dl1 = (len1 / l0) - (REAL)1.0, dl2 = (len2 / l0) - (REAL)1.0;
    REAL scale = Ka * RSQRT(a0) / l0 * (REAL)0.3f;
    for (int d = 0; d < 3; d++) {
      REAL f0 = (n0[d] * dl0) - (n2[d] * dl2);
      REAL f1 = (n1[d] * dl1) - (n0[d] * dl0);
      REAL f2 = (n2[d] * dl2) - (n1[d] * dl1);
      Fo[4 * i0 + d] += f0 * scale;
      Fo[4 * i1 + d] += f1 * scale;
      Fo[4 * i2 + d] += f2 * scale;
    }
  }
}

/* shaders/Cell3D_Kernel.cl:180-247  StickToSurface */
static void FN(stick_force3d)(const REAL *V, REAL *Fo, const uint32_t *F, int nv, int nf, REAL Ks, REAL l0) {
  if (Ks < (REAL)1e-12f) return;
  REAL com[3]; FN(com3d)(V, nv, com);
  for (int fi = 0; fi < nf; fi++) {
    uint32_t idx[3] = {F[3 * fi], F[3 * fi + 1], F[3 * fi + 2]};
    const REAL *P0 = V + 4 * idx[0], *P1 = V + 4 * idx[1], *P2 = V + 4 * idx[2];
    REAL A[3], B[3], n[3], un[3];
    for (int d = 0; d < 3; d++) { A[d] = P1[d] - P0[d]; B[d] = P2[d] - P0[d]; }
    FN(cross3)(A, B, n);
    FN(normalize3)(n, un);
    if (!(un[2] < (REAL)-0.1f)) continue;
    for (int k = 0; k < 3; k++) {
      const REAL *pos = V + 4 * idx[k];
      if (pos[2] < (REAL)0.0) Fo[4 * idx[k] + 2] += Ks * RABS(pos[2]);
      REAL height = RABS(pos[2]);
      if (height < l0 * (REAL)2.0) {
        REAL ctv[3] = {pos[0] - com[0], pos[1] - com[1], (REAL)0.0 - com[2]};
        /* distance(pos, surface_pos) = sqrt(0+0+z*z) */
        REAL dist = RSQRT(pos[2] * pos[2]);
        /* ftmp = Ks * (1.0 - dist / l0)  ->dbl  (:242) */
        REAL ftmp = (REAL)((double)Ks * (1.0 - (double)(dist / l0)));
        REAL u[3]; FN(normalize3)(ctv, u);
        for (int d = 0; d < 3; d++) Fo[4 * idx[k] + d] += u[d] * ftmp;
      }
    }
  }
}

/* Winding number of point p against cell cj's mesh, shifted.
 * shaders/Cell3D_Kernel.cl:283-303 */
static REAL FN(winding3d)(const REAL *Vj, const uint32_t *F, int nf, const REAL *shift, const REAL *p) {
  REAL total = 0;
  for (int fj = 0; fj < nf; fj++) {
    REAL a[3], b[3], c[3], u[3], v[3], w[3], vw[3];
    for (int d = 0; d < 3; d++) {
      a[d] = Vj[4 * F[3 * fj] + d] + shift[d] - p[d];
      b[d] = Vj[4 * F[3 * fj + 1] + d] + shift[d] - p[d];
      c[d] = Vj[4 * F[3 * fj + 2] + d] + shift[d] - p[d];
    }
    FN(normalize3)(a, u); FN(normalize3)(b, v); FN(normalize3)(c, w);
    REAL denom = (REAL)1.0 + FN(dot3)(u, v) + FN(dot3)(v, w) + FN(dot3)(w, u);
    if (denom < (REAL)1e-8f) continue;
    FN(cross3)(v, w, vw);
    REAL num = FN(dot3)(u, vw);
    total += (REAL)2.0 * RATAN2(num, denom);
  }
  return total / ((REAL)4.0 * (REAL)3.14159274101257f); /* 4.0f * M_PI_F */
}

/* One (vertex, cj) evaluation of RepellingForces. Returns winding number. */
static inline REAL FN(repel_pair3d)(const REAL *Vj, const uint32_t *F, int nf, const REAL *comi, const REAL *comj,
                                    int PBC, REAL L, REAL Kc, const REAL *p, REAL *fo) {
  REAL shift[3] = {0, 0, 0};
  if (PBC) for (int d = 0; d < 3; d++) shift[d] = L * RROUND((comi[d] - comj[d]) / L);
  REAL wn = FN(winding3d)(Vj, F, nf, shift, p);
  if (RABS(wn) < (REAL)1e-6f) return wn;
  REAL dv[3] = {comi[0] - p[0], comi[1] - p[1], comi[2] - p[2]}, dir[3];
  FN(normalize3)(dv, dir);
  REAL m = RABS(wn) * (REAL)0.5 * Kc;
  for (int d = 0; d < 3; d++) fo[d] += m * dir[d];
  return wn;
}

/* ------------------------------------------------------------------------ */
/* Forces of one 3D step.  `which` bit mask: 1 volume, 2 area(edge), 4 stick,
 * 8 repel.  If cand==NULL repulsion is ALL-PAIRS (the reference algorithm,
 * shaders/Cell3D_Kernel.cl:269-309); otherwise only pairs in the candidate
 * lists AND passing the per-vertex exact AABB cull are evaluated (culled
 * form, equal to all-pairs in exact arithmetic: SURVEY A.3).
 * contacts (optional): per vertex, number of cells with |wn|>=1e-3 (force-carrying) and number
 * of noise-level contacts 1e-6<=|wn|<1e-3 are accumulated in contacts[2*vid+{0,1}].
 * Kernel order per step: src/Tissue3D.cpp:372-423. */
static void FN(forces3d_range)(int nc, int nv, int nf, const uint32_t *faces, const REAL *verts, REAL *forces,
                               const REAL *Kv, const REAL *Ka, const REAL *Ks, const REAL *v0, const REAL *a0,
                               const REAL *l0, REAL Kc, int PBC, REAL L, int which, const int32_t *cand_count,
                               const int32_t *cand, int cand_stride, int32_t *contacts, int c0, int c1,
                               int stale_from, REAL *prev_volumes) {
  /* ClearForces :366-369 */
  memset(forces, 0, sizeof(REAL) * 4 * (size_t)nc * nv);
  REAL *coms = (REAL *)malloc(sizeof(REAL) * 3 * nc);
  REAL *lo = (REAL *)malloc(sizeof(REAL) * 3 * nc), *hi = (REAL *)malloc(sizeof(REAL) * 3 * nc);
  REAL *emax = (REAL *)malloc(sizeof(REAL) * nc);
#pragma omp parallel for schedule(static)
  for (int ci = 0; ci < nc; ci++) {
    const REAL *V = verts + 4 * (size_t)ci * nv;
    REAL *Fo = forces + 4 * (size_t)ci * nv;
    FN(com3d)(V, nv, coms + 3 * ci);
    {
      REAL e2 = 0;
      for (int fi = 0; fi < nf; fi++) for (int k = 0; k < 3; k++) {
        const REAL *a = V + 4 * faces[3 * fi + k], *b = V + 4 * faces[3 * fi + (k + 1) % 3];
        REAL dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2], l2 = dx * dx + dy * dy + dz * dz;
        if (l2 > e2) e2 = l2;
      }
      emax[ci] = RSQRT(e2);
    }
    for (int d = 0; d < 3; d++) { lo[3 * ci + d] = V[d]; hi[3 * ci + d] = V[d]; }
    for (int i = 1; i < nv; i++) for (int d = 0; d < 3; d++) {
      REAL x = V[4 * i + d];
      if (x < lo[3 * ci + d]) lo[3 * ci + d] = x;
      if (x > hi[3 * ci + d]) hi[3 * ci + d] = x;
    }
    if (ci < c0 || ci >= c1) continue; /* bounds/COM are needed for every cell, forces only for [c0,c1) */
    if (which & 1) FN(volume_force3d)(V, Fo, faces, nv, nf, Kv[ci], v0[ci], stale_from, prev_volumes ? prev_volumes + ci : NULL);
    if (which & 2) FN(area_force3d)(V, Fo, faces, nf, Ka[ci], a0[ci], l0[ci]);
    if (which & 4) FN(stick_force3d)(V, Fo, faces, nv, nf, Ks[ci], l0[ci]);
  }
  if ((which & 8) && Kc != (REAL)0.0) {
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
    for (int ci = c0; ci < c1; ci++) {
      for (int vi = 0; vi < nv; vi++) {
        const REAL *comi = coms + 3 * ci;
        int ncand = cand ? cand_count[ci] : nc;
        const REAL *p = verts + 4 * ((size_t)ci * nv + vi);
        REAL *fo = forces + 4 * ((size_t)ci * nv + vi);
        for (int k = 0; k < ncand; k++) {
          int cj = cand ? cand[(size_t)ci * cand_stride + k] : k;
          if (cj == ci) continue;
          const REAL *comj = coms + 3 * cj;
          if (cand) {
            /* exact per-vertex cull.  The reference drops every face with denom < 1e-8, i.e. every face that
             * subtends at least pi steradians (:293-295), so w_ref = W - (1/4pi) sum_{skipped} Omega.  A planar
             * triangle subtends >= pi only if the vertex projects inside it and is within R/sqrt(3) <= e/3 of its
             * plane (R <= e/sqrt(3): enclosing circle, e: longest edge).  Hence: p farther than 0.34 * emax(cj) from
             * the (shifted) AABB of cj => no face skipped and W = 0  =>  wn == 0 in exact arithmetic. */
            int out = 0;
            REAL pad = (REAL)0.34f * emax[cj];
            for (int d = 0; d < 3; d++) {
              REAL sh = PBC ? L * RROUND((comi[d] - comj[d]) / L) : (REAL)0.0;
              REAL l = (lo[3 * cj + d] + sh) - pad, h = (hi[3 * cj + d] + sh) + pad;
              if (p[d] < l || p[d] > h) out = 1;
            }
            if (out) continue;
          }
          REAL wn = FN(repel_pair3d)(verts + 4 * (size_t)cj * nv, faces, nf, comi, comj, PBC, L, Kc, p, fo);
          if (contacts) {
            /* [0]: contacts that carry force above noise level (|wn| >= 1e-3); [1]: noise-level ones */
            if (RABS(wn) >= (REAL)1e-3f) contacts[2 * ((size_t)ci * nv + vi)]++;
            else if (RABS(wn) >= (REAL)1e-6f) contacts[2 * ((size_t)ci * nv + vi) + 1]++;
          }
        }
      }
    }
  }
  free(coms); free(lo); free(hi); free(emax);
}

void FN(oracle3d_forces)(int nc, int nv, int nf, const uint32_t *faces, const REAL *verts, REAL *forces,
                         const REAL *Kv, const REAL *Ka, const REAL *Ks, const REAL *v0, const REAL *a0,
                         const REAL *l0, REAL Kc, int PBC, REAL L, int which,
                         const int32_t *cand_count, const int32_t *cand, int cand_stride, int32_t *contacts) {
  FN(forces3d_range)(nc, nv, nf, faces, verts, forces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, which, cand_count, cand,
                     cand_stride, contacts, 0, nc, -1, NULL);
}

/* Same, but forces are evaluated only for cells [c0, c1) (against ALL cells): a bounded sample of the
 * reference's all-pairs work, used to time the CPU baseline on large tissues. */
void FN(oracle3d_forces_range)(int nc, int nv, int nf, const uint32_t *faces, const REAL *verts, REAL *forces,
                               const REAL *Kv, const REAL *Ka, const REAL *Ks, const REAL *v0, const REAL *a0,
                               const REAL *l0, REAL Kc, int PBC, REAL L, int which, int c0, int c1) {
  FN(forces3d_range)(nc, nv, nf, faces, verts, forces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, which, NULL, NULL, 0, NULL,
                     c0, c1, -1, NULL);
}

/* RepellingForces (shaders/Cell3D_Kernel.cl:251-310) for vertices [vi0, vi1) of cell ci against ALL nc cells present
 * (the reference's all-pairs loop over cj, :269-309), nothing else: a bounded sample of the reference algorithm for
 * tissues where even one cell's all-pairs work takes minutes (bench.py --impl reference).  The cj loop is split in
 * NCHUNK chunks so that a handful of vertices still occupies every host thread; a vertex's partial forces are added
 * in chunk order.  out: (vi1 - vi0) x 4. */
void FN(oracle3d_repel_sample)(int nc, int nv, int nf, const uint32_t *faces, const REAL *verts, REAL Kc, int PBC, REAL L,
                               int ci, int vi0, int vi1, REAL *out) {
  enum { NCHUNK = 64 };
  const int nvs = vi1 - vi0;
  REAL *coms = (REAL *)malloc(sizeof(REAL) * 3 * (size_t)nc);
  REAL *part = (REAL *)calloc((size_t)nvs * NCHUNK * 3, sizeof(REAL));
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; c++) FN(com3d)(verts + 4 * (size_t)c * nv, nv, coms + 3 * c);
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int s = 0; s < nvs; s++) {
    for (int ch = 0; ch < NCHUNK; ch++) {
      const REAL *p = verts + 4 * ((size_t)ci * nv + vi0 + s);
      REAL *fo = part + 3 * ((size_t)s * NCHUNK + ch);
      const int j0 = (int)((long long)nc * ch / NCHUNK), j1 = (int)((long long)nc * (ch + 1) / NCHUNK);
      for (int cj = j0; cj < j1; cj++) {
        if (cj == ci) continue;
        FN(repel_pair3d)(verts + 4 * (size_t)cj * nv, faces, nf, coms + 3 * ci, coms + 3 * cj, PBC, L, Kc, p, fo);
      }
    }
  }
  for (int s = 0; s < nvs; s++) {
    REAL f[3] = {0, 0, 0};
    for (int ch = 0; ch < NCHUNK; ch++) for (int d = 0; d < 3; d++) f[d] += part[3 * ((size_t)s * NCHUNK + ch) + d];
    out[4 * s] = f[0]; out[4 * s + 1] = f[1]; out[4 * s + 2] = f[2]; out[4 * s + 3] = 0;
  }
  free(coms); free(part);
}

/* shaders/Cell3D_Kernel.cl:313-364  AllVertAttraction, literally (scatter form: the work-item of vertex (ci,vi)
 * visits every vertex (cj,vj) of every other cell and adds -t to its own force and +t to the OTHER vertex's force,
 * with the rest length l0[ci] of ITS cell).  The reference host never enqueues this kernel, so the
 * default product path does not run it either; it is the §8(f) rank-1 "next" row, enabled by DPM3D_ATTRACT.
 * Serial here (the scatter would race under OpenMP); forces are ACCUMULATED into `forces`.
 * The product evaluates the equivalent gather form  F_i -= sum_j [g(d,l0_i) + g(d,l0_j)] * delta/d  (the second term
 * is what vertex j's work-item scatters onto i: its delta is exactly -delta and its distance exactly d). */
void FN(oracle3d_attract)(int nc, int nv, const REAL *verts, REAL *forces, const REAL *l0, REAL L, int PBC, REAL Kat) {
  if (Kat == (REAL)0.0) return; /* :318-319 */
  for (int ci = 0; ci < nc; ci++) for (int vi = 0; vi < nv; vi++) {
    const size_t i = (size_t)ci * nv + vi;
    const REAL *p1 = verts + 4 * i;
    for (int cj = 0; cj < nc; cj++) {
      if (cj == ci) continue;
      for (int vj = 0; vj < nv; vj++) {
        const size_t j = (size_t)cj * nv + vj;
        const REAL *p2 = verts + 4 * j;
        REAL delta[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        if (PBC) for (int d = 0; d < 3; d++) delta[d] -= L * RROUND(delta[d] / L); /* :338-343 */
        REAL dist = RSQRT(FN(dot3)(delta, delta));
        REAL L0 = l0[ci];
        if (L0 < (REAL)1e-12f) L0 = (REAL)1e-12f;
        if (dist < L0 * (REAL)2.0f && dist > (REAL)1e-12f) { /* :350 */
          for (int d = 0; d < 3; d++) {
            REAL t = Kat * (REAL)0.5f * (dist / L0 - (REAL)1.0f) * (delta[d] / dist);
            forces[4 * i + d] += -t;
            forces[4 * j + d] += t;
          }
        }
      }
    }
  }
}

/* EulerPosition :371-381 */
void FN(oracle3d_euler)(int nc, int nv, REAL *verts, const REAL *forces, REAL dt) {
  size_t n = (size_t)nc * nv;
  for (size_t i = 0; i < n; i++) for (int d = 0; d < 3; d++) verts[4 * i + d] += forces[4 * i + d] * dt;
}

/* nsteps of {forces; euler}; forces holds the last step's forces on return
 * (src/Tissue3D.cpp:372-434). */
void FN(oracle3d_run)(int nc, int nv, int nf, const uint32_t *faces, REAL *verts, REAL *forces, const REAL *Kv,
                      const REAL *Ka, const REAL *Ks, const REAL *v0, const REAL *a0, const REAL *l0, REAL Kc,
                      int PBC, REAL L, int nsteps, REAL dt, int which) {
  for (int s = 0; s < nsteps; s++) {
    FN(oracle3d_forces)(nc, nv, nf, faces, verts, forces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, which, NULL, NULL, 0, NULL);
    FN(oracle3d_euler)(nc, nv, verts, forces, dt);
  }
}

/* One CLEulerUpdate call of the reference AS IT EXECUTES on NVIDIA's OpenCL: faces >= stale_from use the volume
 * of the previous step (zero on the first step of the call) — see volume_force3d. */
void FN(oracle3d_run_compat)(int nc, int nv, int nf, const uint32_t *faces, REAL *verts, REAL *forces, const REAL *Kv,
                             const REAL *Ka, const REAL *Ks, const REAL *v0, const REAL *a0, const REAL *l0, REAL Kc,
                             int PBC, REAL L, int nsteps, REAL dt, int which, int stale_from) {
  REAL *prev = (REAL *)calloc(nc, sizeof(REAL));
  for (int s = 0; s < nsteps; s++) {
    FN(forces3d_range)(nc, nv, nf, faces, verts, forces, Kv, Ka, Ks, v0, a0, l0, Kc, PBC, L, which, NULL, NULL, 0, NULL,
                       0, nc, stale_from, prev);
    FN(oracle3d_euler)(nc, nv, verts, forces, dt);
  }
  free(prev);
}

/* ======================================================================== */
/* 2D  (shaders/Cell2D_kernel.cl).  verts/forces are [nc][S][2], S = maxNV.  */
/* Padding slots (vi >= NV[ci]) never influence real vertices and are not    */
/* copied back by the host (src/Tissue2D.cpp:235-240); they are left alone.  */
/* ======================================================================== */

/* Cell2D_kernel.cl:3-11 GetCOM */
static void FN(com2d)(const REAL *V, int n, REAL *com) {
  REAL sx = 0, sy = 0;
  for (int i = 0; i < n; i++) { sx += V[2 * i]; sy += V[2 * i + 1]; }
  com[0] = sx / (REAL)n; com[1] = sy / (REAL)n;
}

/* point-in-polygon toggle loop of RepulsionForceUpdate, :166-196 (literal PBC quirk :178-187) */
static int FN(inside2d)(const REAL *p, const REAL *Vj, int nj, int PBC, REAL L) {
  int overlaps = 0;
  for (int i = 0, j = nj - 1; i < nj; j = i++) {
    REAL dix = p[0] - Vj[2 * i], diy = p[1] - Vj[2 * i + 1];
    REAL djx = p[0] - Vj[2 * j], djy = p[1] - Vj[2 * j + 1];
    if (PBC) {
      if (RABS(dix) > L || RABS(djx) > L) { dix -= L * RFLOOR(dix / L); djx -= L * RFLOOR(djx / L); }
      if (RABS(diy) > L || RABS(djy) > L) { diy -= L * RROUND(diy / L); djy -= L * RROUND(djy / L); }
    }
    if ((diy > 0) != (djy > 0) && ((REAL)0 < (djx - dix) * ((REAL)0 - diy) / (djy - diy) + dix)) overlaps = !overlaps;
  }
  return overlaps;
}

/* One 2D force evaluation. which: 1 area, 2 perimeter, 4 bending, 8 attraction, 16 repulsion.
 * Kernel order src/Tissue2D.cpp:216-221.  cand==NULL => all-pairs. For the culled form
 * cand lists candidate cells for attraction+repulsion (near set) and far_count/far lists
 * the cells whose |d|>L wrap quirk can fire; per-vertex culls applied inside. */
static void FN(forces2d_range)(int nc, int S, const int32_t *NV, const REAL *verts, REAL *forces, const REAL *Ka,
                         const REAL *Kl, const REAL *Kb, const REAL *a0, const REAL *l0, const REAL *r0, REAL Kre,
                         REAL Kat, int PBC, REAL L, int which, const int32_t *cand_count, const int32_t *cand,
                         int cand_stride, int32_t *inside_flags, int c0, int c1) {
  memset(forces, 0, sizeof(REAL) * 2 * (size_t)nc * S);
  REAL *lo = (REAL *)malloc(sizeof(REAL) * 2 * nc), *hi = (REAL *)malloc(sizeof(REAL) * 2 * nc);
  for (int ci = 0; ci < nc; ci++) {
    const REAL *V = verts + 2 * (size_t)ci * S;
    for (int d = 0; d < 2; d++) { lo[2 * ci + d] = V[d]; hi[2 * ci + d] = V[d]; }
    for (int i = 1; i < NV[ci]; i++) for (int d = 0; d < 2; d++) {
      REAL x = V[2 * i + d];
      if (x < lo[2 * ci + d]) lo[2 * ci + d] = x;
      if (x > hi[2 * ci + d]) hi[2 * ci + d] = x;
    }
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int ci = c0; ci < c1; ci++) {
    const int n = NV[ci];
    const REAL *V = verts + 2 * (size_t)ci * S;
    REAL *Fo = forces + 2 * (size_t)ci * S;
    /* AreaForceUpdates :13-42 ; area recomputed per work-item there, same value */
    REAL Area = 0;
    for (int vj = 0; vj < n; vj++) {
      int pm = (vj == 0) ? n - 1 : vj - 1;
      /* Area += 0.5 * (...)  ->dbl (:30) */
      Area = (REAL)((double)Area + 0.5 * (double)((V[2 * pm] + V[2 * vj]) * (V[2 * pm + 1] - V[2 * vj + 1])));
    }
    if (Area < 0.0) Area = -Area;
    REAL strain = (REAL)((double)(Area / a0[ci]) - 1.0); /* ->dbl :37 */
    REAL com[2]; FN(com2d)(V, n, com);
    for (int vi = 0; vi < n; vi++) {
      int im1 = (vi == 0) ? n - 1 : vi - 1, ip1 = (vi == n - 1) ? 0 : vi + 1;
      REAL fx = 0, fy = 0;
      if (which & 1) {
        /* :38-41 ; both components use (im1 - ip1) (sic) ; ->dbl */
        double c = (double)(Ka[ci] / RSQRT(a0[ci])) * 0.5 * (double)strain;
        fx = (REAL)((double)fx + c * (double)(V[2 * im1 + 1] - V[2 * ip1 + 1]));
        fy = (REAL)((double)fy + c * (double)(V[2 * im1] - V[2 * ip1]));
      }
      if (which & 2) {
        /* PerimeterForceUpdates :92-119 */
        REAL lvx = V[2 * ip1] - V[2 * vi], lvy = V[2 * ip1 + 1] - V[2 * vi + 1];
        REAL lmx = V[2 * vi] - V[2 * im1], lmy = V[2 * vi + 1] - V[2 * im1 + 1];
        REAL len = RSQRT(lvx * lvx + lvy * lvy), lenm = RSQRT(lmx * lmx + lmy * lmy);
        REAL ux = lvx / len, uy = lvy / len, umx = lmx / lenm, umy = lmy / lenm;
        REAL dli = (REAL)((double)(len / l0[ci]) - 1.0), dlim1 = (REAL)((double)(lenm / l0[ci]) - 1.0); /* ->dbl :116-117 */
        REAL k = Kl[ci] * RSQRT(a0[ci] / l0[ci]);
        fx += k * (dli * ux - dlim1 * umx);
        fy += k * (dli * uy - dlim1 * umy);
      }
      if (which & 4) {
        /* BendingForceUpdates :44-90 */
        int ip2 = (ip1 == n - 1) ? 0 : ip1 + 1, im2 = (im1 == 0) ? n - 1 : im1 - 1;
        REAL lvx = V[2 * ip1] - V[2 * vi], lvy = V[2 * ip1 + 1] - V[2 * vi + 1];
        REAL lvxm = V[2 * vi] - V[2 * im1], lvym = V[2 * vi + 1] - V[2 * im1 + 1];
        REAL six = lvx - lvxm, siy = lvy - lvym;
        REAL sixp = (V[2 * ip2] - V[2 * ip1]) - lvx, siyp = (V[2 * ip2 + 1] - V[2 * ip1 + 1]) - lvy;
        REAL sixm = lvxm - (V[2 * im1] - V[2 * im2]), siym = lvym - (V[2 * im1 + 1] - V[2 * im2 + 1]);
        /* Kb * (2.0 * six - sixm - sixp) ->dbl :88-89 */
        fx = (REAL)((double)fx + (double)Kb[ci] * (2.0 * (double)six - (double)sixm - (double)sixp));
        fy = (REAL)((double)fy + (double)Kb[ci] * (2.0 * (double)siy - (double)siym - (double)siyp));
      }
      const REAL *p = V + 2 * vi;
      int ncand = cand ? cand_count[ci] : nc;
      if (which & 8) {
        /* AttractionForceUpdate :222-268 (runs even when Kat == 0) */
        for (int k = 0; k < ncand; k++) {
          int cj = cand ? cand[(size_t)ci * cand_stride + k] : k;
          if (cj == ci) continue;
          const REAL *Vj = verts + 2 * (size_t)cj * S;
          for (int vj = 0; vj < NV[cj]; vj++) {
            REAL rx = Vj[2 * vj] - p[0], ry = Vj[2 * vj + 1] - p[1];
            if (PBC) { rx -= L * RROUND(rx / L); ry -= L * RROUND(ry / L); }
            REAL dist = RSQRT(rx * rx + ry * ry);
            if (dist < l0[ci]) {
              REAL ftmp = Kat / (REAL)n * dist / l0[ci];
              REAL nn = RSQRT(rx * rx + ry * ry);
              if (nn != (REAL)0) { fx += ftmp * (rx / nn); fy += ftmp * (ry / nn); } /* normalize(0) = 0 */
            }
          }
        }
      }
      if (which & 16) {
        /* RepulsionForceUpdate :121-220 */
        int overlaps = 0;
        for (int k = 0; k < ncand && !overlaps; k++) {
          int cj = cand ? cand[(size_t)ci * cand_stride + k] : k;
          if (cj == ci) continue;
          if (cand) {
            /* per-vertex cull, exact: evaluate iff p in AABB(cj) or some |d|>L can occur */
            REAL dxl = p[0] - lo[2 * cj], dxh = p[0] - hi[2 * cj], dyl = p[1] - lo[2 * cj + 1], dyh = p[1] - hi[2 * cj + 1];
            int inx = (dxl >= 0 && dxh <= 0), iny = (dyl >= 0 && dyh <= 0);
            int farx = PBC && (RABS(dxl) > L || RABS(dxh) > L), fary = PBC && (RABS(dyl) > L || RABS(dyh) > L);
            if (!((inx || farx) && (iny || fary))) continue;
          }
          overlaps = FN(inside2d)(p, verts + 2 * (size_t)cj * S, NV[cj], PBC, L);
        }
        if (inside_flags) inside_flags[(size_t)ci * S + vi] = overlaps;
        if (overlaps) {
          REAL dx = com[0] - p[0], dy = com[1] - p[1];
          if (PBC) { dx -= L * RROUND(dx / L); dy -= L * RROUND(dy / L); }
          REAL dist = RSQRT(dx * dx + dy * dy);
          REAL xij = dist / ((REAL)2 * r0[ci]);
          REAL ftmp = Kre * ((REAL)1 - xij);
          REAL nn = RSQRT(dx * dx + dy * dy);
          if (nn != (REAL)0) { /* normalize(0) = 0 */
            fx += (REAL)0.5f * ftmp * (dx / nn);
            fy += (REAL)0.5f * ftmp * (dy / nn);
          }
        }
      }
      Fo[2 * vi] = fx; Fo[2 * vi + 1] = fy;
    }
  }
  free(lo); free(hi);
}

void FN(oracle2d_forces)(int nc, int S, const int32_t *NV, const REAL *verts, REAL *forces, const REAL *Ka,
                         const REAL *Kl, const REAL *Kb, const REAL *a0, const REAL *l0, const REAL *r0, REAL Kre,
                         REAL Kat, int PBC, REAL L, int which, const int32_t *cand_count, const int32_t *cand,
                         int cand_stride, int32_t *inside_flags) {
  FN(forces2d_range)(nc, S, NV, verts, forces, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, which, cand_count, cand, cand_stride,
                     inside_flags, 0, nc);
}

/* Same, all-pairs, forces of cells [c0, c1) only (against ALL cells): a bounded sample of the reference algorithm. */
void FN(oracle2d_forces_range)(int nc, int S, const int32_t *NV, const REAL *verts, REAL *forces, const REAL *Ka,
                               const REAL *Kl, const REAL *Kb, const REAL *a0, const REAL *l0, const REAL *r0, REAL Kre,
                               REAL Kat, int PBC, REAL L, int which, int c0, int c1) {
  FN(forces2d_range)(nc, S, NV, verts, forces, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, which, NULL, NULL, 0, NULL, c0, c1);
}

/* EulerUpdate :270-281 (real vertices only; forces are NOT zeroed here so the
 * caller can read the last step's forces, src/Tissue2D.cpp:223-227) */
void FN(oracle2d_euler)(int nc, int S, const int32_t *NV, REAL *verts, const REAL *forces, REAL dt) {
  for (int ci = 0; ci < nc; ci++)
    for (int vi = 0; vi < NV[ci]; vi++)
      for (int d = 0; d < 2; d++) verts[2 * ((size_t)ci * S + vi) + d] += forces[2 * ((size_t)ci * S + vi) + d] * dt;
}

void FN(oracle2d_run)(int nc, int S, const int32_t *NV, REAL *verts, REAL *forces, const REAL *Ka, const REAL *Kl,
                      const REAL *Kb, const REAL *a0, const REAL *l0, const REAL *r0, REAL Kre, REAL Kat, int PBC,
                      REAL L, int nsteps, REAL dt, int which) {
  for (int s = 0; s < nsteps; s++) {
    FN(oracle2d_forces)(nc, S, NV, verts, forces, Ka, Kl, Kb, a0, l0, r0, Kre, Kat, PBC, L, which, NULL, NULL, 0, NULL);
    FN(oracle2d_euler)(nc, S, NV, verts, forces, dt);
  }
}
