// neighbor.cu — cell-list neighbour search built by an integer (counting) sort.
//
// The reference has NO neighbour search: every contact kernel is all-pairs
// (shaders/Cell3D_Kernel.cl:269-309, shaders/Cell2D_kernel.cl:166-202,:251-267).
// This file supplies the O(N) replacement.  Its integer artefacts (bin ids, sorted
// permutation, bin start table, per-cell candidate lists) are specified operation by
// operation in DESIGN.md and restated on the CPU in oracle/dpm_oracle.c
// (oracle_cell_list); tests require bit-exact equality, so every floating-point
// operation below is an explicitly rounded single IEEE op (__f*_rn), never an FMA.
//
// One kernel does the whole rebuild; it returns immediately when the device-side flag
// NbrState::rebuild is 0, so it is enqueued before every step.  It runs as ONE thread-block
// cluster (8 CTAs x 512 threads, cluster-wide barriers between the phases) instead of a
// cooperative grid: a cluster launch is an ordinary launch, so the per-step no-op costs what
// any small kernel costs and it can be chained to its neighbours in the stream with
// programmatic dependent launch (a cooperative launch cost ~6 us per step and broke the chain).
#include <cooperative_groups.h>

#include "dpm_common.cuh"

namespace cg = cooperative_groups;

namespace dpm {

#define NBMAX 1024
#define RB_THREADS 512
#define RB_CLUSTER 8  // portable cluster size
#define KMAX 128

__device__ __forceinline__ float centre_rn(float lo, float hi) { return __fmul_rn(0.5f, __fadd_rn(lo, hi)); }
__device__ __forceinline__ float half_rn(float lo, float hi) { return __fmul_rn(0.5f, __fsub_rn(hi, lo)); }

// Grid parameters from the global reductions — identical on every CTA.
__device__ void compute_grid(dpm_grid_t &g, int nd, int pbc, float L, float skin_rel, float range, int cap,
                             float max_ext, const float *gmin, const float *gmax) {
  float skin = __fmul_rn(skin_rel, max_ext);
  float margin = __fadd_rn(skin, range);
  float binw_req = __fadd_rn(max_ext, margin);
  g.max_ext = max_ext;
  g.margin = margin;
  for (int d = 0; d < 3; d++) {
    g.nb[d] = 1; g.periodic[d] = 0; g.allpass[d] = 0; g.origin[d] = 0.0f; g.inv_binw[d] = 0.0f;
  }
  for (int d = 0; d < nd; d++) {
    float span = __fsub_rn(gmax[d], gmin[d]);
    if (pbc && __fadd_rn(span, __fadd_rn(binw_req, binw_req)) >= L) {
      int nbp = (int)floorf(__fdiv_rn(L, binw_req));
      if (nbp > NBMAX) nbp = NBMAX;
      if (nbp >= 3) { g.nb[d] = nbp; g.periodic[d] = 1; g.inv_binw[d] = __fdiv_rn((float)nbp, L); }
      else { g.nb[d] = 1; g.allpass[d] = 1; }
    } else {
      int nb = (int)floorf(__fdiv_rn(span, binw_req)) + 1;
      if (nb > NBMAX) nb = NBMAX;
      g.nb[d] = nb; g.origin[d] = gmin[d]; g.inv_binw[d] = __fdiv_rn(1.0f, binw_req);
      if (nb == NBMAX) g.inv_binw[d] = __fdiv_rn((float)NBMAX, __fadd_rn(span, binw_req));
    }
  }
  while ((long long)g.nb[0] * g.nb[1] * g.nb[2] > (long long)cap) {
    int d = 0;
    if (g.nb[1] > g.nb[d]) d = 1;
    if (g.nb[2] > g.nb[d]) d = 2;
    if (g.periodic[d]) {
      int nb = g.nb[d] / 2;
      if (nb >= 3) { g.nb[d] = nb; g.inv_binw[d] = __fdiv_rn((float)nb, L); }
      else { g.nb[d] = 1; g.periodic[d] = 0; g.allpass[d] = 1; g.inv_binw[d] = 0.0f; }
    } else {
      g.nb[d] = (g.nb[d] + 1) / 2;
      g.inv_binw[d] = __fmul_rn(g.inv_binw[d], 0.5f);
    }
  }
  g.nbins = g.nb[0] * g.nb[1] * g.nb[2];
  g.pad = 0;
}

__device__ __forceinline__ int axis_bin(const dpm_grid_t &g, int d, int pbc, float L, float lo, float hi) {
  const float cw = wrap_rn(centre_rn(lo, hi), pbc, L);
  int b = (int)floorf(__fmul_rn(__fsub_rn(cw, g.origin[d]), g.inv_binw[d]));
  if (b < 0) b = 0;
  if (b > g.nb[d] - 1) b = g.nb[d] - 1;
  return b;
}
// bin coordinates of a cell; axes >= nd stay 0.  (No dynamically indexed local arrays in this file: with them
// nvcc 12.9 -O3 produced overlapping stack slots in the rebuild kernel below.)
__device__ __forceinline__ int3 cell_bin3(const dpm_grid_t &g, int nd, int pbc, float L, const float4 lo, const float4 hi) {
  int3 ib = make_int3(0, 0, 0);
  ib.x = axis_bin(g, 0, pbc, L, lo.x, hi.x);
  if (nd > 1) ib.y = axis_bin(g, 1, pbc, L, lo.y, hi.y);
  if (nd > 2) ib.z = axis_bin(g, 2, pbc, L, lo.z, hi.z);
  return ib;
}

// neighbour bin along one axis for offset o in {-1,0,1}; -1 if it does not exist.
// (periodic axes have nb >= 3 by construction, so the three wrapped bins are distinct)
__device__ __forceinline__ int nbr_bin(const dpm_grid_t &g, int d, int ib, int o) {
  int b = ib + o;
  if (g.periodic[d]) { b = (b + g.nb[d]) % g.nb[d]; return b; }
  return (b < 0 || b >= g.nb[d]) ? -1 : b;
}

// per-axis separation test of the candidate filter: true = too far apart on this axis
__device__ __forceinline__ bool axis_far(float li, float hi, float lj, float hj, int pbc, float L, float margin) {
  float dd = __fsub_rn(centre_rn(li, hi), centre_rn(lj, hj));
  if (pbc) dd = minimg_rn(dd, L);
  return fabsf(dd) > __fadd_rn(__fadd_rn(half_rn(li, hi), half_rn(lj, hj)), margin);
}

__global__ void __launch_bounds__(RB_THREADS) nbr_rebuild_kernel(NbrBuffers nb) {
  // Programmatic dependent launch: the kernel ahead in the stream (the previous timestep's step kernel, or the upload's
  // bounds kernel) must be complete before its rebuild flag and bounds are read; only then may the kernels behind be
  // scheduled (they prefetch the state that kernel wrote, and wait for THIS kernel before they touch the lists).
  griddep_wait();
  griddep_launch();
  if (blockIdx.x == 0 && threadIdx.x == 0) nb.st->unit_total = 0;  // first kernel of every step: reset the contact-unit queue
  if (nb.st->rebuild == 0) return;  // uniform across the cluster: nobody reaches a barrier
  if (nb.nc_dev) nb.nc = *nb.nc_dev;
  cg::cluster_group grid = cg::this_cluster();  // the launch is one cluster: gridDim.x == cluster size
  const int tid = threadIdx.x;
  const int gtid = blockIdx.x * blockDim.x + tid;
  const int nthreads = gridDim.x * blockDim.x;
  const int lane = tid & 31, warp = tid >> 5;
  __shared__ float sred[RB_THREADS / 32][16];
  __shared__ dpm_grid_t sg;
  __shared__ int sscan[RB_THREADS / 32];
  __shared__ int sbase;

  // ---- P1: global reductions: max extent, wrapped-centre range, raw AABB range ----
  {
    float mx = 0.0f, padmx = 0.0f;
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    float rlo[3] = {INFINITY, INFINITY, INFINITY}, rhi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int c = gtid; c < nb.nc; c += nthreads) {
      float4 lo = nb.blo[(size_t)c * nb.blo_stride], hi = nb.bhi[(size_t)c * nb.blo_stride];
      const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
      padmx = fmaxf(padmx, hi.w);
      if (nb.att_pad_scale > 0.0f) padmx = fmaxf(padmx, nb.att_pad_scale * nb.blo[(size_t)c * nb.blo_stride + 3].w);
#pragma unroll
      for (int d = 0; d < 3; d++) {
        if (d < nb.nd) {
          mx = fmaxf(mx, __fsub_rn(h[d], l[d]));
          const float cw = wrap_rn(centre_rn(l[d], h[d]), nb.pbc, nb.L);
          cmin[d] = fminf(cmin[d], cw); cmax[d] = fmaxf(cmax[d], cw);
          rlo[d] = fminf(rlo[d], l[d]); rhi[d] = fmaxf(rhi[d], h[d]);
        }
      }
    }
    float vals[14];
    vals[0] = warp_max(mx);
    vals[13] = warp_max(padmx);
    for (int d = 0; d < 3; d++) {
      vals[1 + d] = warp_min(cmin[d]); vals[4 + d] = warp_max(cmax[d]);
      vals[7 + d] = warp_min(rlo[d]); vals[10 + d] = warp_max(rhi[d]);
    }
    if (lane == 0) for (int i = 0; i < 14; i++) sred[warp][i] = vals[i];
    __syncthreads();
    if (tid < 14) {
      float v = sred[0][tid];
      bool ismax = (tid == 0) || (tid >= 4 && tid <= 6) || (tid >= 10);
      for (int w = 1; w < RB_THREADS / 32; w++) v = ismax ? fmaxf(v, sred[w][tid]) : fminf(v, sred[w][tid]);
      nb.partial[blockIdx.x * 16 + tid] = v;
    }
  }
  grid.sync();
  if (tid == 0) {
    float vals[14];
    for (int i = 0; i < 14; i++) {
      bool ismax = (i == 0) || (i >= 4 && i <= 6) || (i >= 10);
      float v = nb.partial[i];
      for (int b = 1; b < (int)gridDim.x; b++) v = ismax ? fmaxf(v, nb.partial[b * 16 + i]) : fminf(v, nb.partial[b * 16 + i]);
      vals[i] = v;
    }
    const float range = nb.range_from_bounds ? __fmul_rn(nb.range_scale, vals[13]) : nb.range;
    compute_grid(sg, nb.nd, nb.pbc, nb.L, nb.skin_rel, range, nb.cap, vals[0], &vals[1], &vals[4]);
    if (blockIdx.x == 0) {
      nb.st->grid = sg;
      nb.st->range = range;
      for (int d = 0; d < 3; d++) { nb.st->glo[d] = vals[7 + d]; nb.st->ghi[d] = vals[10 + d]; }
    }
  }
  __syncthreads();
  const dpm_grid_t g = sg;
  const float skin = __fmul_rn(nb.skin_rel, g.max_ext);

  // ---- P2: zero the histogram ----
  for (int b = gtid; b <= g.nbins; b += nthreads) nb.bin_count[b] = 0;
  if (gtid == 0) nb.st->next = 0;
  grid.sync();
  const float Lm = __fsub_rn(nb.L, skin);
  const bool far2d = nb.far2d && nb.nd == 2 && nb.pbc && nb.ext_list;

  // ---- P3: bin ids + histogram ----
  for (int c = gtid; c < nb.nc; c += nthreads) {
    const int3 ib = cell_bin3(g, nb.nd, nb.pbc, nb.L, nb.blo[(size_t)c * nb.blo_stride], nb.bhi[(size_t)c * nb.blo_stride]);
    const int id = (ib.z * g.nb[1] + ib.y) * g.nb[0] + ib.x;
    nb.bin_id[c] = id;
    atomicAdd(&nb.bin_count[id], 1);
    if (far2d) {
      // cells that can see (or be) a |d| > L wrap partner: only those near the global extremes
      const float4 lo = nb.blo[(size_t)c * nb.blo_stride], hi = nb.bhi[(size_t)c * nb.blo_stride];
      const bool maybe = (__fsub_rn(nb.st->ghi[0], lo.x) > Lm) || (__fsub_rn(hi.x, nb.st->glo[0]) > Lm) ||
                         (__fsub_rn(nb.st->ghi[1], lo.y) > Lm) || (__fsub_rn(hi.y, nb.st->glo[1]) > Lm);
      if (maybe) nb.ext_list[atomicAdd(&nb.st->next, 1)] = c;
    }
  }
  grid.sync();

  // ---- P4: exclusive scan of the histogram -> bin_start (chunk per CTA) ----
  const int chunk = (g.nbins + gridDim.x - 1) / gridDim.x;
  const int b0 = blockIdx.x * chunk, b1 = min(b0 + chunk, g.nbins);
  {
    int s = 0;
    for (int b = b0 + tid; b < b1; b += RB_THREADS) s += nb.bin_count[b];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sscan[warp] = s;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < RB_THREADS / 32; w++) t += sscan[w];
      nb.chunk_sum[blockIdx.x] = t;
    }
  }
  grid.sync();
  {
    if (tid == 0) {
      int t = 0;
      for (int b = 0; b < (int)blockIdx.x; b++) t += nb.chunk_sum[b];
      sbase = t;
    }
    __syncthreads();
    int base = sbase;
    for (int t0 = b0; t0 < b1; t0 += RB_THREADS) {
      int b = t0 + tid;
      int cnt = (b < b1) ? nb.bin_count[b] : 0;
      int incl = cnt;
      for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
      __syncthreads();
      if (lane == 31) sscan[warp] = incl;
      __syncthreads();
      int woff = 0;
      for (int w = 0; w < warp; w++) woff += sscan[w];
      int tot = 0;
      for (int w = 0; w < RB_THREADS / 32; w++) tot += sscan[w];
      if (b < b1) { nb.bin_start[b] = base + woff + incl - cnt; nb.bin_count[b] = 0; }
      base += tot;
    }
    if (gtid == 0) nb.bin_start[g.nbins] = nb.nc;
  }
  grid.sync();

  // ---- P5: scatter (order inside a bin fixed in P6) ----
  for (int c = gtid; c < nb.nc; c += nthreads) {
    int id = nb.bin_id[c];
    int pos = nb.bin_start[id] + atomicAdd(&nb.bin_count[id], 1);
    nb.order[pos] = c;
  }
  grid.sync();

  // ---- P6: ascending cell id inside each bin == stable sort by (bin, id) ----
  for (int b = gtid; b < g.nbins; b += nthreads) {
    int s0 = nb.bin_start[b], s1 = nb.bin_start[b + 1];
    for (int a = s0 + 1; a < s1; a++) {
      int v = nb.order[a], q = a - 1;
      while (q >= s0 && nb.order[q] > v) { nb.order[q + 1] = nb.order[q]; q--; }
      nb.order[q + 1] = v;
    }
  }
  grid.sync();

  // ---- P7: candidate lists (thread per cell) + build-time boxes ----
  const float hskin = 0.5f * skin;
  for (int i = gtid; i < nb.nc; i += nthreads) {
    const float4 loi = nb.blo[(size_t)i * nb.blo_stride], hii = nb.bhi[(size_t)i * nb.blo_stride];
    nb.bbox_lo[i] = make_float4(loi.x - hskin, loi.y - hskin, loi.z - hskin, 0.f);
    nb.bbox_hi[i] = make_float4(hii.x + hskin, hii.y + hskin, hii.z + hskin, 0.f);
    if (i >= nb.nc_list) continue;
    const int3 ib = cell_bin3(g, nb.nd, nb.pbc, nb.L, loi, hii);
    int tmp[KMAX];
    int n = 0;
    for (int oz = -1; oz <= 1; oz++) {
      const int bz = nbr_bin(g, 2, ib.z, oz);
      if (bz < 0 || (g.nb[2] == 1 && oz != 0)) continue;
      for (int oy = -1; oy <= 1; oy++) {
        const int by = nbr_bin(g, 1, ib.y, oy);
        if (by < 0 || (g.nb[1] == 1 && oy != 0)) continue;
        for (int ox = -1; ox <= 1; ox++) {
          const int bx = nbr_bin(g, 0, ib.x, ox);
          if (bx < 0 || (g.nb[0] == 1 && ox != 0)) continue;
          const int b = (bz * g.nb[1] + by) * g.nb[0] + bx;
          const int s1 = nb.bin_start[b + 1];
          for (int s = nb.bin_start[b]; s < s1; s++) {
            const int j = nb.order[s];
            if (j == i) continue;
            const float4 loj = nb.blo[(size_t)j * nb.blo_stride], hij = nb.bhi[(size_t)j * nb.blo_stride];
            bool ok = true;
            if (!g.allpass[0] && axis_far(loi.x, hii.x, loj.x, hij.x, nb.pbc, nb.L, g.margin)) ok = false;
            if (ok && nb.nd > 1 && !g.allpass[1] && axis_far(loi.y, hii.y, loj.y, hij.y, nb.pbc, nb.L, g.margin)) ok = false;
            if (ok && nb.nd > 2 && !g.allpass[2] && axis_far(loi.z, hii.z, loj.z, hij.z, nb.pbc, nb.L, g.margin)) ok = false;
            if (ok) { if (n < KMAX) tmp[n] = j; n++; }
          }
        }
      }
    }
    if (far2d) {
      // a wrap partner j of i has farx or fary below, which puts BOTH cells near opposite global extremes: only the
      // extreme cells look, and only at the extreme cells (their order in ext_list does not matter: tmp is sorted)
      const bool maybe = (__fsub_rn(nb.st->ghi[0], loi.x) > Lm) || (__fsub_rn(hii.x, nb.st->glo[0]) > Lm) ||
                         (__fsub_rn(nb.st->ghi[1], loi.y) > Lm) || (__fsub_rn(hii.y, nb.st->glo[1]) > Lm);
      if (maybe) {
        const int next = nb.st->next;
        for (int e = 0; e < next; e++) {
          const int j = nb.ext_list[e];
          if (j == i) continue;
          const float4 loj = nb.blo[(size_t)j * nb.blo_stride], hij = nb.bhi[(size_t)j * nb.blo_stride];
          // per axis "overlap or far", and far on at least one axis (see oracle_cell_list)
          const bool farx = (__fsub_rn(hij.x, loi.x) > Lm) || (__fsub_rn(hii.x, loj.x) > Lm);
          const bool fary = (__fsub_rn(hij.y, loi.y) > Lm) || (__fsub_rn(hii.y, loj.y) > Lm);
          if (!(farx || fary)) continue;
          const bool ovx = !((__fsub_rn(loj.x, hii.x) > skin) || (__fsub_rn(loi.x, hij.x) > skin));
          const bool ovy = !((__fsub_rn(loj.y, hii.y) > skin) || (__fsub_rn(loi.y, hij.y) > skin));
          if (!((ovx || farx) && (ovy || fary))) continue;
          bool dup = false;
          for (int q = 0; q < min(n, KMAX); q++) dup |= (tmp[q] == j);
          if (!dup) { if (n < KMAX) tmp[n] = j; n++; }
        }
      }
    }
    int m = min(n, KMAX);
    for (int a = 1; a < m; a++) {  // ascending (global) cell id == the reference's cj loop order
      const int v = tmp[a], kv = nb.gid ? nb.gid[v] : v;
      int q = a - 1;
      while (q >= 0 && (nb.gid ? nb.gid[tmp[q]] : tmp[q]) > kv) { tmp[q + 1] = tmp[q]; q--; }
      tmp[q + 1] = v;
    }
    nb.cand_count[i] = n;
    if (n > nb.K) nb.st->overflow = 1;
    for (int a = 0; a < m && a < nb.K; a++) nb.cand[(size_t)i * nb.K + a] = tmp[a];
  }
  grid.sync();
  if (gtid == 0) { nb.st->rebuild = 0; nb.st->nbuilds += 1; }
}

int rebuild_max_grid(int device) {
  (void)device;
  return RB_CLUSTER;  // CTAs of the one cluster (sizes of NbrBuffers::partial / chunk_sum)
}

cudaError_t launch_rebuild(const NbrBuffers &nb, cudaStream_t stream, int coop_grid) {
  (void)coop_grid;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(RB_CLUSTER); cfg.blockDim = dim3(RB_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RB_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, nbr_rebuild_kernel, nb);
}

}  // namespace dpm
