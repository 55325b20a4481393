// dpm3d.cu — C ABI of the 3D path: handle, topology tables, upload / step / download.
// Host-side counterpart of src/Tissue3D.cpp:199-470 in the reference (buffers, kernel
// arguments, the step loop and the two read-backs), over CUDA instead of OpenCL.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include <chrono>
#include "dpm3d_ctx.cuh"

using namespace dpm;

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Ring adjacency of a closed oriented 2-manifold: for vertex v the cyclic list n_0..n_{k-1}
// such that ring face i is (v, n_i, n_{i+1}) in the mesh orientation.
int build_rings(int nv, int nf, const uint32_t *faces, std::vector<uint16_t> &ring_nbr, std::vector<uint16_t> &ring_face,
                std::vector<uint8_t> &valence, int &stride, int &minval, int &maxval) {
  std::vector<std::vector<std::array<uint32_t, 3>>> inc(nv);  // (next, nextnext, face)
  for (int f = 0; f < nf; f++) {
    for (int k = 0; k < 3; k++) {
      uint32_t v = faces[3 * f + k], a = faces[3 * f + (k + 1) % 3], b = faces[3 * f + (k + 2) % 3];
      if (v >= (uint32_t)nv || a >= (uint32_t)nv || b >= (uint32_t)nv)
        return fail(DPM_ERR_RUNTIME, "Invalid face indices");  // src/Tissue3D.cpp:149-154
      if (v == a || a == b || v == b) return fail(DPM_ERR_TOPOLOGY, "degenerate face (repeated vertex index)");
      inc[v].push_back({a, b, (uint32_t)f});
    }
  }
  maxval = 0; minval = 1 << 30;
  for (int v = 0; v < nv; v++) { maxval = std::max(maxval, (int)inc[v].size()); minval = std::min(minval, (int)inc[v].size()); }
  if (maxval > 16) return fail(DPM_ERR_TOPOLOGY, "vertex valence > 16 is not supported");
  stride = maxval <= 8 ? 8 : 16;  // 8 x uint16 = one 16-byte load per vertex in the step kernel
  ring_nbr.assign((size_t)nv * stride, 0);
  ring_face.assign((size_t)nv * stride, 0);
  valence.assign(nv, 0);
  for (int v = 0; v < nv; v++) {
    auto &L = inc[v];
    int k = (int)L.size();
    if (k < 3) return fail(DPM_ERR_TOPOLOGY, "mesh is not closed: a vertex has fewer than 3 incident faces");
    std::vector<char> used(k, 0);
    uint32_t cur = 0;  // start at the lowest-index incident face (inc lists are in face order)
    for (int i = 0; i < k; i++) {
      if (used[cur]) return fail(DPM_ERR_TOPOLOGY, "vertex fan is not a single cycle (non-manifold mesh)");
      used[cur] = 1;
      ring_nbr[(size_t)v * stride + i] = (uint16_t)(16u * L[cur][0]);  // byte offset of the neighbour in the staged float4 ring (nv <= 1024)
      ring_face[(size_t)v * stride + i] = (uint16_t)L[cur][2];
      uint32_t want = L[cur][1];
      int nxt = -1;
      for (int q = 0; q < k; q++)
        if (L[q][0] == want) { if (nxt >= 0) return fail(DPM_ERR_TOPOLOGY, "edge shared by more than two faces"); nxt = q; }
      if (nxt < 0) return fail(DPM_ERR_TOPOLOGY, "mesh is not closed or not consistently oriented");
      cur = (uint32_t)nxt;
    }
    if (cur != 0) return fail(DPM_ERR_TOPOLOGY, "vertex fan does not close");
    valence[v] = (uint8_t)k;
  }
  return DPM_OK;
}

// Static tables of the fast contact evaluation: the face across each edge, and for every face the faces of its
// edge-adjacency rings 0..RING_MAX in BFS order with the cumulative ring ends.
void build_face_tables(int nf, const uint32_t *faces, std::vector<ushort4> &adj, std::vector<uint16_t> &ring_tab,
                       std::vector<uint8_t> &ring_end) {
  std::map<std::pair<uint32_t, uint32_t>, int> edge;
  for (int f = 0; f < nf; f++)
    for (int k = 0; k < 3; k++) edge[{faces[3 * f + k], faces[3 * f + (k + 1) % 3]}] = f;
  adj.resize(nf);
  for (int f = 0; f < nf; f++) {
    unsigned short a[3];
    for (int k = 0; k < 3; k++) a[k] = (unsigned short)edge[{faces[3 * f + (k + 1) % 3], faces[3 * f + k]}];  // closed manifold: exists
    adj[f] = make_ushort4(a[0], a[1], a[2], 0);
  }
  ring_tab.assign((size_t)nf * RING_TAB, 0);
  ring_end.assign((size_t)nf * (RING_MAX + 1), 0);
  std::vector<int> mark(nf, -1);
  for (int f = 0; f < nf; f++) {
    std::vector<int> frontier{f}, order{f};
    mark[f] = f;
    ring_end[(size_t)f * (RING_MAX + 1)] = 1;
    for (int r = 1; r <= RING_MAX; r++) {
      std::vector<int> next;
      for (int g : frontier) {
        const unsigned short nb[3] = {adj[g].x, adj[g].y, adj[g].z};
        for (int k = 0; k < 3; k++)
          if (mark[nb[k]] != f && (int)order.size() < RING_TAB) { mark[nb[k]] = f; next.push_back(nb[k]); order.push_back(nb[k]); }
      }
      ring_end[(size_t)f * (RING_MAX + 1) + r] = (uint8_t)order.size();
      frontier.swap(next);
    }
    for (size_t j = 0; j < order.size(); j++) ring_tab[(size_t)f * RING_TAB + j] = (uint16_t)order[j];
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Bank-conflict-aware processing order of the step kernel (host side, once per mesh).
// The step kernel gathers float4 positions from shared memory by vertex index: an LDS.128 is served 8 lanes at a time and
// needs one wavefront per distinct 16-byte bank group (index mod 8) it touches twice, so the 6 ring gathers of 32 consecutive
// vertices (and the 3 corner gathers of 32 consecutive faces) of the reference's icosphere numbering cost 2.1x (1.7x) the
// conflict-free number of wavefronts — measured with ncu, and reproduced exactly by the count below.  The memory layout
// stays the reference's; what is chosen here is (1) the ORDER in which threads take the vertices (octets of lanes = one
// vertex of every residue class, picked so that their i-th ring neighbours fall into different bank groups), (2) the
// rotation of every vertex's cyclic ring list, (3) the order of the faces inside each aligned block of 32 (so that a warp's
// stores of per-face results still cover one 128-byte line).  Greedy construction + hill climbing, deterministic.
// ---------------------------------------------------------------------------------------------------------------------
struct MeshLayout {
  std::vector<uint16_t> vorder;  // [nv] vertex processed by slot r
  std::vector<int> rot;          // [nv] rotation of the vertex's ring list
  std::vector<uint16_t> forder;  // [nf] face processed by slot r
};
struct LayoutRng {
  uint64_t s;
  uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
  uint32_t below(uint32_t n) { return next() % n; }
};
static inline int lds128_wavefronts(const int *idx, int n) {  // one quarter-warp: distinct addresses per bank group, largest multiplicity
  int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seen[8], ns = 0;
  for (int a = 0; a < n; a++) {
    const int x = idx[a];
    if (x < 0) continue;
    bool dup = false;
    for (int q = 0; q < ns; q++) dup = dup || seen[q] == x;
    if (dup) continue;
    seen[ns++] = x;
    cnt[x & 7]++;
  }
  int m = 0;
  for (int r = 0; r < 8; r++) m = std::max(m, cnt[r]);
  return m;
}
MeshLayout optimise_layout(int nv, int nf, const uint32_t *faces, const std::vector<uint16_t> &ring_vertex, const std::vector<uint8_t> &valence,
                           int stride, int maxval) {
  static std::map<std::vector<uint32_t>, MeshLayout> cache;  // one search per mesh and process
  std::vector<uint32_t> key(faces, faces + 3 * (size_t)nf);
  key.push_back((uint32_t)nv);
  auto hit = cache.find(key);
  if (hit != cache.end()) return hit->second;
  MeshLayout L;
  L.rot.assign(nv, 0);
  LayoutRng rng{12345};
  auto nbr = [&](int v, int i) { const int k = valence[v]; return i < k ? (int)ring_vertex[(size_t)v * stride + (i + L.rot[v]) % k] : -1; };
  auto octet = [&](const int *vs, int n) {
    int c = lds128_wavefronts(vs, n), idx[8];
    for (int i = 0; i < maxval; i++) {
      for (int a = 0; a < n; a++) idx[a] = nbr(vs[a], i);
      c += lds128_wavefronts(idx, n);
    }
    return c;
  };
  std::vector<int> order;
  {
    std::vector<std::vector<int>> pool(8);
    for (int v = 0; v < nv; v++) pool[v & 7].push_back(v);
    for (auto &p : pool) for (int i = (int)p.size() - 1; i > 0; i--) std::swap(p[i], p[rng.below(i + 1)]);
    for (;;) {
      bool any = false;
      for (auto &p : pool) any = any || !p.empty();
      if (!any) break;
      int oc[8], n = 0, cls[8];
      for (int i = 0; i < 8; i++) cls[i] = i;
      for (int i = 7; i > 0; i--) std::swap(cls[i], cls[rng.below(i + 1)]);
      for (int ci = 0; ci < 8; ci++) {
        auto &p = pool[cls[ci]];
        if (p.empty()) continue;
        int bc = 1 << 30, bv = -1, br = 0, bi = 0;
        const int lim = std::min<int>((int)p.size(), 40);
        for (int q = 0; q < lim; q++) {
          const int v = p[q];
          for (int ro = 0; ro < (int)valence[v]; ro++) {
            L.rot[v] = ro;
            oc[n] = v;
            const int c = octet(oc, n + 1);
            if (c < bc) { bc = c; bv = v; br = ro; bi = q; }
          }
          L.rot[v] = 0;
        }
        L.rot[bv] = br;
        oc[n++] = bv;
        p.erase(p.begin() + bi);
      }
      for (int i = 0; i < n; i++) order.push_back(oc[i]);
    }
    const int N = (int)order.size();
    for (long it = 0; it < 300000; it++) {
      const int a = (int)rng.below(N);
      if (rng.next() & 1) {
        const int v = order[a], old = L.rot[v], q = a / 8 * 8, n = std::min(8, N - q);
        const int before = octet(&order[q], n);
        L.rot[v] = (int)rng.below(valence[v]);
        if (octet(&order[q], n) > before) L.rot[v] = old;
      } else {
        const int b = (int)rng.below(N), qa = a / 8 * 8, qb = b / 8 * 8;
        if (qa == qb) continue;
        const int na = std::min(8, N - qa), nb = std::min(8, N - qb);
        const int before = octet(&order[qa], na) + octet(&order[qb], nb);
        std::swap(order[a], order[b]);
        if (octet(&order[qa], na) + octet(&order[qb], nb) > before) std::swap(order[a], order[b]);
      }
    }
  }
  L.vorder.assign(order.begin(), order.end());
  std::vector<int> fo(nf);
  for (int i = 0; i < nf; i++) fo[i] = i;
  auto foct = [&](const int *fs, int n) {
    int c = 0, idx[8];
    for (int k = 0; k < 3; k++) {
      for (int a = 0; a < n; a++) idx[a] = (int)faces[3 * (size_t)fs[a] + k];
      c += lds128_wavefronts(idx, n);
    }
    return c;
  };
  for (int b0 = 0; b0 < nf; b0 += 32) {
    const int n = std::min(32, nf - b0);
    if (n <= 8) continue;
    int *blk = &fo[b0];
    auto btot = [&]() { int t = 0; for (int q = 0; q < n; q += 8) t += foct(blk + q, std::min(8, n - q)); return t; };
    const int ideal = 3 * ((n + 7) / 8);
    int bb = btot();
    std::vector<int> bestb(blk, blk + n);
    for (int rs = 0; rs < 20 && bb > ideal; rs++) {
      for (int i = n - 1; i > 0; i--) std::swap(blk[i], blk[rng.below(i + 1)]);
      int c = btot();
      for (int it = 0; it < 1200 && c > ideal; it++) {
        const int i = (int)rng.below(n), j = (int)rng.below(n);
        if (i / 8 == j / 8) continue;
        const int qa = i / 8 * 8, qb = j / 8 * 8, na = std::min(8, n - qa), nb = std::min(8, n - qb);
        const int before = foct(blk + qa, na) + foct(blk + qb, nb);
        std::swap(blk[i], blk[j]);
        const int after = foct(blk + qa, na) + foct(blk + qb, nb);
        if (after > before) std::swap(blk[i], blk[j]); else c += after - before;
      }
      if (c < bb) { bb = c; bestb.assign(blk, blk + n); }
    }
    std::copy(bestb.begin(), bestb.end(), blk);
  }
  L.forder.assign(fo.begin(), fo.end());
  if (getenv("DPM_TRACE")) {
    int ring0 = 0, ring1 = 0, face0 = 0, face1 = 0;
    std::vector<int> keep = L.rot, idn(nv), idf(nf);
    for (int i = 0; i < nv; i++) idn[i] = i;
    for (int i = 0; i < nf; i++) idf[i] = i;
    for (int q = 0; q < nv; q += 8) ring1 += octet(&order[q], std::min(8, nv - q));
    for (int q = 0; q < nf; q += 8) face1 += foct(&fo[q], std::min(8, nf - q));
    L.rot.assign(nv, 0);
    for (int q = 0; q < nv; q += 8) ring0 += octet(&idn[q], std::min(8, nv - q));
    for (int q = 0; q < nf; q += 8) face0 += foct(&idf[q], std::min(8, nf - q));
    L.rot = keep;
    fprintf(stderr, "[dpm3d] shared-memory wavefronts per cell (LDS.128 gathers): ring pass %d -> %d (conflict-free %d), face pass %d -> %d (conflict-free %d)\n",
            ring0, ring1, (maxval + 1) * ((nv + 7) / 8), face0, face1, 3 * ((nf + 7) / 8));
  }
  cache[key] = L;
  return L;
}

int pick_config(dpm3d_ctx *h) {
  if (h->nv > 1024) return fail(DPM_ERR_INVALID_ARGUMENT, "meshes with more than 1024 vertices per cell are not supported yet");
  h->threads = STEP_THREADS; h->vpt = (h->nv + STEP_THREADS - 1) / STEP_THREADS;
  return DPM_OK;
}

size_t smem_for(dpm3d_ctx *h) {
  const char *pad = getenv("DPM_SMEM_PAD");  // experiments: extra dynamic shared memory per CTA (lowers the CTAs per SM)
  return step3d_smem_bytes(h->nv, h->nf) + (pad ? (size_t)atoi(pad) : 0);
}

// The step kernel variant for this mesh: ring slots read per vertex / slots every vertex has, and the compat mode.
template <bool COMPAT, bool ATT, typename Fn>
static cudaError_t with_step_kernel_mode(dpm3d_ctx *h, Fn &&fn) {
  if (h->ring_stride > 8) return fn(dpm3d_step_kernel<16, 3, COMPAT, ATT>);
  if (h->min_valence >= 5 && h->max_valence <= 6) return fn(dpm3d_step_kernel<6, 5, COMPAT, ATT>);
  return fn(dpm3d_step_kernel<8, 3, COMPAT, ATT>);
}
template <typename Fn>
static cudaError_t with_step_kernel(dpm3d_ctx *h, bool att, Fn &&fn) {
  const bool compat = h->stale_from >= 0;
  if (compat) return att ? with_step_kernel_mode<true, true>(h, fn) : with_step_kernel_mode<true, false>(h, fn);
  return att ? with_step_kernel_mode<false, true>(h, fn) : with_step_kernel_mode<false, false>(h, fn);
}

cudaError_t set_smem(dpm3d_ctx *h) {
  const int smem = (int)h->smem;
  for (int mode = 0; mode < 4; mode++) {  // every mode: dpm3d_set_compat / the attraction may be switched at any time
    const int keep = h->stale_from;
    h->stale_from = (mode & 1) ? 0 : -1;
    cudaError_t e = with_step_kernel(h, (mode & 2) != 0, [&](auto *k) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    h->stale_from = keep;
    if (e != cudaSuccess) return e;
  }
  return cudaFuncSetAttribute(dpm3d_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

// Launch with programmatic stream serialization: the kernel's CTAs may be scheduled while the kernel ahead of it in the
// stream drains, and run up to their griddep_wait() (dpm3d_kernels.cuh).
template <typename K>
static cudaError_t launch_pdl(K *kernel, int grid, int block, size_t smem, cudaStream_t stream, const Step3DParams &p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

cudaError_t launch_step(dpm3d_ctx *h, const Step3DParams &p) {
  return with_step_kernel(h, (p.mask & DPM3D_ATTRACT) != 0, [&](auto *k) { return launch_pdl(k, p.nc, STEP_THREADS, h->smem, h->stream, p); });
}

CellTopo cell_topo(dpm3d_ctx *h) {
  CellTopo T;
  T.faces = h->faces; T.ring_nbr = h->ring_nbr; T.valence = h->valence; T.ring_stride = h->ring_stride; T.nv = h->nv; T.nf = h->nf;
  return T;
}

NbrBuffers nbr_buffers(dpm3d_ctx *h, int pbc, float L) {
  NbrBuffers nb{};
  nb.st = h->st;
  nb.blo = h->bnd[h->cur];
  nb.bhi = h->bnd[h->cur] + 1;
  nb.blo_stride = BND;
  nb.bbox_lo = h->bbox_lo; nb.bbox_hi = h->bbox_hi;
  nb.bin_id = h->bin_id; nb.order = h->order; nb.bin_count = h->bin_count; nb.bin_start = h->bin_start;
  nb.cand_count = h->cand_count; nb.cand = h->cand;
  nb.partial = h->partial; nb.chunk_sum = h->chunk_sum;
  nb.nc = h->nc; nb.nc_list = h->nc; nb.nd = 3; nb.cap = h->cap; nb.K = h->K;
  nb.nc_dev = h->nranks > 1 ? reinterpret_cast<const int *>(h->sd) : nullptr;  // ShardDev::n_total is its first member
  nb.gid = h->gid;
  nb.pbc = pbc; nb.L = L; nb.skin_rel = h->skin_rel; nb.range = 0.0f; nb.far2d = 0;
  nb.range_from_bounds = 1; nb.range_scale = RANGE_HEADROOM;
  nb.att_pad_scale = h->att_active ? ATT_REACH / RANGE_HEADROOM * 1.0001f : 0.0f;  // range >= ATT_REACH * max l0
  return nb;
}

int alloc_units(dpm3d_ctx *h, int per_cell) {
  if (h->unit_rec && per_cell <= h->unit_per_cell) return DPM_OK;
  if (h->unit_rec) cudaFree(h->unit_rec);
  if (h->unit_w) cudaFree(h->unit_w);
  if (h->unit_att) cudaFree(h->unit_att);
  h->unit_rec = nullptr; h->unit_w = nullptr; h->unit_att = nullptr;
  h->unit_per_cell = per_cell;
  h->unit_cap = (int)std::min<long long>((long long)h->nc * per_cell, 1ll << 30);
  DPM_CUDA_TRY(cudaMalloc(&h->unit_rec, sizeof(int2) * (size_t)h->unit_cap));
  DPM_CUDA_TRY(cudaMalloc(&h->unit_w, sizeof(float) * (size_t)h->unit_cap));
  if (h->mask & DPM3D_ATTRACT) DPM_CUDA_TRY(cudaMalloc(&h->unit_att, sizeof(float4) * (size_t)h->unit_cap));
  return DPM_OK;
}

int alloc_cand(dpm3d_ctx *h) {
  if (h->K_alloc >= h->K) return DPM_OK;
  if (h->cand) cudaFree(h->cand);
  h->cand = nullptr;
  DPM_CUDA_TRY(cudaMalloc(&h->cand, sizeof(int) * (size_t)h->nc * h->K));
  h->K_alloc = h->K;
  return DPM_OK;
}

}  // namespace

extern "C" {

int dpm3d_create(dpm3d_t **out, int device, int ncells, int nv, int nf, const uint32_t *faces) {
  if (!out) return fail(DPM_ERR_INVALID_ARGUMENT, "handle pointer is NULL");
  *out = nullptr;
  if (ncells <= 0) return fail(DPM_ERR_INVALID_ARGUMENT, "NCELLS must be positive");  // src/Tissue3D.cpp:132-135
  if (nv < 4 || nf < 4 || !faces) return fail(DPM_ERR_INVALID_ARGUMENT, "bad mesh size");
  int ndev = 0;
  DPM_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(DPM_ERR_CUDA, "no such CUDA device (there is no CPU fallback)");
  std::vector<uint16_t> rn, rf;
  std::vector<uint8_t> val;
  int stride = 0, minval = 0, maxval = 0;
  int rc = build_rings(nv, nf, faces, rn, rf, val, stride, minval, maxval);
  if (rc) return rc;
  // processing order + ring rotations that minimise shared-memory bank conflicts (rn holds byte offsets = 16 * vertex)
  std::vector<uint16_t> rv(rn.size());
  for (size_t i = 0; i < rn.size(); i++) rv[i] = (uint16_t)(rn[i] / 16);
  const MeshLayout lay = getenv("DPM_NO_LAYOUT") ? MeshLayout{} : optimise_layout(nv, nf, faces, rv, val, stride, maxval);
  std::vector<uint16_t> vorder(nv), forder(nf);
  for (int i = 0; i < nv; i++) vorder[i] = lay.vorder.empty() ? (uint16_t)i : lay.vorder[i];
  for (int i = 0; i < nf; i++) forder[i] = lay.forder.empty() ? (uint16_t)i : lay.forder[i];
  if (!lay.rot.empty()) {
    std::vector<uint16_t> rn2 = rn, rf2 = rf;
    for (int v = 0; v < nv; v++)
      for (int i = 0; i < (int)val[v]; i++) {
        rn2[(size_t)v * stride + i] = rn[(size_t)v * stride + (i + lay.rot[v]) % val[v]];
        rf2[(size_t)v * stride + i] = rf[(size_t)v * stride + (i + lay.rot[v]) % val[v]];
      }
    rn.swap(rn2); rf.swap(rf2);
  }
  DeviceGuard guard(device);
  dpm3d_ctx *h = new dpm3d_ctx();
  h->device = device; h->nc = ncells; h->nslots = ncells; h->nv = nv; h->nf = nf; h->ring_stride = stride;
  h->min_valence = minval; h->max_valence = maxval;
  rc = pick_config(h);
  if (rc) { delete h; return rc; }
  auto bail = [&](int code) { dpm3d_destroy(h); return code; };
#define TRYB(expr)                                                                                    \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return bail(fail(DPM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
  } while (0)
  TRYB(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  TRYB(cudaEventCreate(&h->ev0));
  TRYB(cudaEventCreate(&h->ev1));
  TRYB(cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming));
  const size_t nvert = (size_t)ncells * nv;
  TRYB(cudaMalloc(&h->pos[0], sizeof(float4) * nvert));
  TRYB(cudaMalloc(&h->pos[1], sizeof(float4) * nvert));
  TRYB(cudaMalloc(&h->force, sizeof(float4) * nvert));
  TRYB(cudaMemsetAsync(h->force, 0, sizeof(float4) * nvert, h->stream));
  TRYB(cudaMalloc(&h->bnd[0], sizeof(float4) * BND * ncells));
  TRYB(cudaMalloc(&h->bnd[1], sizeof(float4) * BND * ncells));
  TRYB(cudaMalloc(&h->cellA, sizeof(float4) * ncells));
  TRYB(cudaMalloc(&h->cellB, sizeof(float4) * ncells));
  TRYB(cudaMallocHost(&h->h_cell, sizeof(float4) * 2 * ncells));
  TRYB(cudaMalloc(&h->faces, sizeof(ushort4) * nf));
  TRYB(cudaMalloc(&h->ring_nbr, sizeof(uint16_t) * rn.size()));
  TRYB(cudaMalloc(&h->ring_face, sizeof(uint16_t) * rf.size()));
  TRYB(cudaMalloc(&h->valence, nv));
  {
    std::vector<ushort4> f4(nf);
    for (int f = 0; f < nf; f++) f4[f] = make_ushort4((unsigned short)faces[3 * f], (unsigned short)faces[3 * f + 1], (unsigned short)faces[3 * f + 2], 0);
    TRYB(cudaMemcpy(h->faces, f4.data(), sizeof(ushort4) * nf, cudaMemcpyHostToDevice));
    TRYB(cudaMemcpy(h->ring_nbr, rn.data(), sizeof(uint16_t) * rn.size(), cudaMemcpyHostToDevice));
    TRYB(cudaMemcpy(h->ring_face, rf.data(), sizeof(uint16_t) * rf.size(), cudaMemcpyHostToDevice));
    TRYB(cudaMemcpy(h->valence, val.data(), nv, cudaMemcpyHostToDevice));
    TRYB(cudaMalloc(&h->vorder, sizeof(uint16_t) * nv));
    TRYB(cudaMemcpy(h->vorder, vorder.data(), sizeof(uint16_t) * nv, cudaMemcpyHostToDevice));
    std::vector<ushort4> fp(nf);
    for (int r = 0; r < nf; r++) {
      const uint32_t f = forder[r];
      fp[r] = make_ushort4((unsigned short)faces[3 * f], (unsigned short)faces[3 * f + 1], (unsigned short)faces[3 * f + 2], (unsigned short)f);
    }
    TRYB(cudaMalloc(&h->faces_proc, sizeof(ushort4) * nf));
    TRYB(cudaMemcpy(h->faces_proc, fp.data(), sizeof(ushort4) * nf, cudaMemcpyHostToDevice));
    std::vector<ushort4> adj;
    std::vector<uint16_t> rtab;
    std::vector<uint8_t> rend;
    build_face_tables(nf, faces, adj, rtab, rend);
    TRYB(cudaMalloc(&h->face_adj, sizeof(ushort4) * nf));
    TRYB(cudaMalloc(&h->ring_tab, sizeof(uint16_t) * rtab.size()));
    TRYB(cudaMalloc(&h->ring_end, rend.size()));
    TRYB(cudaMalloc(&h->dir_table, sizeof(uint16_t) * DIR_N * DIR_N));
    TRYB(cudaMemcpy(h->face_adj, adj.data(), sizeof(ushort4) * nf, cudaMemcpyHostToDevice));
    TRYB(cudaMemcpy(h->ring_tab, rtab.data(), sizeof(uint16_t) * rtab.size(), cudaMemcpyHostToDevice));
    TRYB(cudaMemcpy(h->ring_end, rend.data(), rend.size(), cudaMemcpyHostToDevice));
    TRYB(cudaMemset(h->dir_table, 0, sizeof(uint16_t) * DIR_N * DIR_N));
    h->h_faces.assign(faces, faces + 3 * (size_t)nf);
  }
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
  rc = alloc_cand(h);
  if (rc) return bail(rc);
  TRYB(cudaMalloc(&h->flag[0], (size_t)ncells * vflag_stride(nv)));  // per-vertex flags (dpm3d_kernels.cuh: vflag_stride)
  TRYB(cudaMalloc(&h->flag[1], (size_t)ncells * vflag_stride(nv)));
  TRYB(cudaMalloc(&h->vlist, sizeof(uint2) * nvert));
  TRYB(cudaMalloc(&h->vlist_cnt, sizeof(int) * ncells));
  TRYB(cudaMemset(h->vlist_cnt, 0, sizeof(int) * ncells));
  TRYB(cudaMalloc(&h->unit_base, sizeof(int) * ncells));
  TRYB(cudaMalloc(&h->unit_cnt, sizeof(int) * ncells));
  TRYB(cudaMemset(h->unit_cnt, 0, sizeof(int) * ncells));
  {
    const size_t tb = sizeof(float) * (size_t)terms_stride(nf) * ncells;
    TRYB(cudaMalloc(&h->terms, tb));
    TRYB(cudaMemset(h->terms, 0, tb));  // the padding behind the last face stays +0.0f: adding it changes no chain
    TRYB(cudaMalloc(&h->part, sizeof(float4) * 3 * STEP_WARPS * (size_t)ncells));
    const size_t gb = sizeof(int) * (size_t)((ncells + CHAIN_GROUP - 1) / CHAIN_GROUP);
    TRYB(cudaMalloc(&h->grp_done, gb));
    TRYB(cudaMemset(h->grp_done, 0, gb));
  }
  h->npatch = (nf + PATCH_F - 1) / PATCH_F;
  TRYB(cudaMalloc(&h->patch_box, sizeof(float4) * 2 * (size_t)h->npatch * ncells));
  rc = alloc_units(h, std::min(4096, std::max(256, 4 * nv)));
  if (rc) return bail(rc);
  {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    // the contact kernel is grid-stride over the unit queue: exactly the CTAs that are resident at once (no second wave)
    int per = 0;
    const char *cg = getenv("DPM_CONTACT_CTAS_PER_SM");  // experiments
    if (cg) per = atoi(cg);
    else if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, dpm3d_contact_kernel<false>, CONTACT_THREADS, 0) != cudaSuccess) per = 0;
    h->contact_grid = sms * (per > 0 ? per : 8);
  }
  h->smem = smem_for(h);
  TRYB(set_smem(h));
#undef TRYB
  *out = h;
  return DPM_OK;
}

int dpm3d_destroy(dpm3d_t *h) {
  if (!h) return DPM_OK;
  DeviceGuard guard(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  shard_free(h);
  void *ptrs[] = {h->pos[0], h->pos[1], h->force, h->bnd[0], h->bnd[1], h->cellA, h->cellB, h->faces, h->ring_nbr, h->ring_face,
                  h->valence, h->face_adj, h->ring_tab, h->ring_end, h->dir_table, h->st, h->bbox_lo, h->bbox_hi, h->bin_id, h->order, h->bin_count, h->bin_start, h->cand_count,
                  h->cand, h->partial, h->chunk_sum, h->unit_rec, h->unit_w, h->unit_att, h->unit_base, h->unit_cnt, h->flag[0], h->flag[1], h->vlist, h->vlist_cnt, h->patch_box, h->terms, h->grp_done, h->part, h->vorder, h->faces_proc};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (h->h_cell) cudaFreeHost(h->h_cell);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_up) cudaEventDestroy(h->ev_up);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return DPM_OK;
}

int dpm3d_set_stream(dpm3d_t *h, void *cuda_stream) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return DPM_OK;
}

int dpm3d_set_neighbor_params(dpm3d_t *h, float skin_rel, int max_candidates) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  if (!(skin_rel >= 0.0f) || max_candidates < 1 || max_candidates > 128)
    return fail(DPM_ERR_INVALID_ARGUMENT, "skin_rel must be >= 0 and 1 <= max_candidates <= 128");
  DeviceGuard guard(h->device);
  h->skin_rel = skin_rel;
  h->K = max_candidates;
  int rc = alloc_cand(h);
  if (rc) return rc;
  h->smem = smem_for(h);
  DPM_CUDA_TRY(set_smem(h));
  h->uploaded = false;  // lists must be rebuilt: require a fresh upload
  return DPM_OK;
}

int dpm3d_get_neighbor_params(dpm3d_t *h, float *skin_rel, int *max_candidates) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  if (skin_rel) *skin_rel = h->skin_rel;
  if (max_candidates) *max_candidates = h->K;
  return DPM_OK;
}

int dpm3d_set_compat(dpm3d_t *h, int stale_volume_from_face) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  h->stale_from = stale_volume_from_face < 0 ? -1 : stale_volume_from_face;
  return DPM_OK;
}

int dpm3d_set_force_mask(dpm3d_t *h, unsigned mask) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  h->mask = mask & (DPM3D_ALL | DPM3D_ATTRACT);
  if ((h->mask & DPM3D_ATTRACT) && !h->unit_att) {  // the attraction term of every unit
    DeviceGuard guard(h->device);
    DPM_CUDA_TRY(cudaMalloc(&h->unit_att, sizeof(float4) * (size_t)h->unit_cap));
  }
  return DPM_OK;
}

static int upload_common(dpm3d_t *h, const float *verts4, bool on_device, const float *Kv, const float *Ka, const float *Ks,
                         const float *v0, const float *a0, const float *l0, bool wait = true) {
  if (!h || !verts4 || !Kv || !Ka || !Ks || !v0 || !a0 || !l0) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL argument");
  DeviceGuard guard(h->device);
  DPM_CUDA_TRY(cudaEventSynchronize(h->ev_up));  // the previous upload's copies out of the pinned staging buffer
  for (int c = 0; c < h->nc; c++) {
    h->h_cell[c] = make_float4(Kv[c], Ka[c], Ks[c], v0[c]);
    h->h_cell[h->nc + c] = make_float4(a0[c], l0[c], 0.f, 0.f);
  }
  h->cur = 0;
  const size_t bytes = sizeof(float4) * (size_t)h->nc * h->nv;
  DPM_CUDA_TRY(cudaMemcpyAsync(h->pos[0], verts4, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(cudaMemcpyAsync(h->cellA, h->h_cell, sizeof(float4) * h->nc, cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(cudaMemcpyAsync(h->cellB, h->h_cell + h->nc, sizeof(float4) * h->nc, cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(cudaMemsetAsync(h->st, 0, sizeof(NbrState), h->stream));
  // walk-start table of the fast contact evaluation, from cell 0's current shape
  dpm3d_dirtable_kernel<<<DIR_N * DIR_N, STEP_THREADS, 0, h->stream>>>(h->pos[0], h->faces, h->nv, h->nf, h->dir_table);
  dpm3d_bounds_kernel<<<h->nc, STEP_THREADS, h->smem, h->stream>>>(h->pos[0], h->bnd[0], h->flag[0], h->nc, cell_topo(h), h->cellB);
  DPM_CUDA_TRY(cudaGetLastError());
  h->stats.launches += 1;
  h->stats.steps = 0; h->stats.rebuilds = 0; h->stats.contact_evals = 0; h->stats.halo_bytes = 0;  // per-upload counters (launches stay cumulative)
  if (h->nranks > 1) shard_reset_counters(h);
  // mark the neighbour lists stale
  static const int one = 1;
  DPM_CUDA_TRY(cudaMemcpyAsync(&h->st->rebuild, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  // the pinned parameter staging buffer is reused by the next upload, which waits on this event; the caller's vertex
  // array is borrowed until the call returns (euler_update keeps it until its download, so it does not wait here)
  DPM_CUDA_TRY(cudaEventRecord(h->ev_up, h->stream));
  if (wait && !on_device) DPM_CUDA_TRY(cudaEventSynchronize(h->ev_up));
  h->uploaded = true;
  return DPM_OK;
}

int dpm3d_upload(dpm3d_t *h, const float *verts4, const float *Kv, const float *Ka, const float *Ks, const float *v0,
                 const float *a0, const float *l0) {
  return upload_common(h, verts4, false, Kv, Ka, Ks, v0, a0, l0);
}
int dpm3d_upload_device(dpm3d_t *h, const float *verts4_dev, const float *Kv, const float *Ka, const float *Ks,
                        const float *v0, const float *a0, const float *l0) {
  return upload_common(h, verts4_dev, true, Kv, Ka, Ks, v0, a0, l0);
}

int dpm3d_rebuild_neighbors(dpm3d_t *h, int pbc, float L) {
  if (!h || !h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "upload first");
  DeviceGuard guard(h->device);
  static const int one = 1;
  DPM_CUDA_TRY(cudaMemcpyAsync(&h->st->rebuild, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  DPM_CUDA_TRY(launch_rebuild(nbr_buffers(h, pbc, L), h->stream, h->coop_grid));
  h->stats.launches += 1;
  h->last_pbc = pbc; h->last_L = L;
  return DPM_OK;
}

static int check_device_flags(dpm3d_t *h);

int dpm3d_step(dpm3d_t *h, int nsteps, float dt, float Kre, float Kat, int pbc, float L) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  // AllVertAttraction is never enqueued by the reference host: Kat only acts under DPM3D_ATTRACT
  const bool attract = (h->mask & DPM3D_ATTRACT) && Kat != 0.0f;  // the kernel returns at once when Kat == 0 (:318-319)
  if (nsteps <= 0) return fail(DPM_ERR_INVALID_ARGUMENT, "nsteps must be positive");                          // src/Tissue3D.cpp:123-126
  if (!(dt > 0.0f) || dt > 0.1f) return fail(DPM_ERR_INVALID_ARGUMENT, "dt must be positive and reasonable");  // :127-131
  if (!h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "dpm3d_step before dpm3d_upload");
  DeviceGuard guard(h->device);
  Step3DParams p{};
  p.cellA = h->cellA; p.cellB = h->cellB; p.faces = h->faces;
  p.ring_nbr = h->ring_nbr; p.ring_face = h->ring_face; p.valence = h->valence; p.ring_stride = h->ring_stride;
  p.face_adj = h->face_adj; p.ring_tab = h->ring_tab; p.ring_end = h->ring_end; p.dir_table = h->dir_table;
  p.cand_count = h->cand_count; p.cand = h->cand; p.K = h->K;
  p.bbox_lo = h->bbox_lo; p.bbox_hi = h->bbox_hi; p.st = h->st;
  p.vlist = h->vlist; p.vlist_cnt = h->vlist_cnt;
  p.patch_box = h->patch_box; p.npatch = h->npatch;
  p.terms = h->terms; p.grp_done = h->grp_done; p.part = h->part;
  p.vorder = h->vorder; p.faces_proc = h->faces_proc;
  p.n_total_dev = h->nranks > 1 ? reinterpret_cast<const int *>(h->sd) : nullptr;  // ShardDev::n_total is its first member
  const int units_grid = h->nranks > 1 ? h->nslots : h->nc;  // sharded: ghost cells refresh their patch boxes too
  p.unit_rec = h->unit_rec; p.unit_w = h->unit_w; p.unit_att = h->unit_att; p.unit_base = h->unit_base; p.unit_cnt = h->unit_cnt; p.unit_cap = h->unit_cap;
  const bool repel = (h->mask & DPM3D_REPEL) && Kre != 0.0f;
  p.nc = h->nc; p.nv = h->nv; p.nf = h->nf; p.dt = dt; p.Kc = Kre; p.pbc = pbc; p.L = L;
  p.mask = attract ? h->mask : (h->mask & ~DPM3D_ATTRACT);
  p.Kat = attract ? Kat : 0.0f;
  p.stale_from = h->stale_from;
  if (pbc != h->last_pbc || L != h->last_L || attract != h->att_active) {  // the lists depend on the box: rebuild when the caller changed it
    static const int one = 1;
    DPM_CUDA_TRY(cudaMemcpyAsync(&h->st->rebuild, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->last_pbc = pbc; h->last_L = L; h->att_active = attract;
  }
  // DPM_TRACE: device time of each kernel of ONE timestep in the middle of the call (events on the stream)
  static const bool trace = getenv("DPM_TRACE") != nullptr;
  const int traced = (trace && nsteps >= 4) ? nsteps / 2 : -1;
  cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int s = 0; s < nsteps; s++) {
    const bool tr = s == traced;
    if (tr) for (auto &e : tev) cudaEventCreate(&e);
    if (h->nranks > 1) {  // ghosts of the current state + the global rebuild decision (dpm_halo.cu)
      int rc = shard_exchange(h, pbc, L);
      if (rc) return rc;
    }
    if (tr) cudaEventRecord(tev[0], h->stream);
    DPM_CUDA_TRY(launch_rebuild(nbr_buffers(h, pbc, L), h->stream, h->coop_grid));
    if (tr) cudaEventRecord(tev[1], h->stream);
    p.pos_in = h->pos[h->cur]; p.pos_out = h->pos[h->cur ^ 1];
    p.bnd_in = h->bnd[h->cur]; p.bnd_out = h->bnd[h->cur ^ 1];
    p.flag_in = h->flag[h->cur]; p.flag_out = h->flag[h->cur ^ 1];
    p.force_out = (s == nsteps - 1) ? h->force : nullptr;  // forces are only read back after the last step (:425-434)
    if (repel || attract) {
      DPM_CUDA_TRY(attract ? launch_pdl(dpm3d_units_kernel<true>, units_grid, UNITS_THREADS, 0, h->stream, p)
                           : launch_pdl(dpm3d_units_kernel<false>, units_grid, UNITS_THREADS, 0, h->stream, p));
      if (tr) cudaEventRecord(tev[2], h->stream);
      DPM_CUDA_TRY(attract ? launch_pdl(dpm3d_contact_kernel<true>, h->contact_grid, CONTACT_THREADS, 0, h->stream, p)
                           : launch_pdl(dpm3d_contact_kernel<false>, h->contact_grid, CONTACT_THREADS, 0, h->stream, p));
    } else if (tr) cudaEventRecord(tev[2], h->stream);
    if (tr) cudaEventRecord(tev[3], h->stream);
    p.push_slot = nullptr;
    if (h->nranks > 1 && shard_fused_targets(h, p.push_pos, p.push_bnd, p.push_gidp)) {
      p.push_slot = reinterpret_cast<const int2 *>(h->push_slot);
      p.push_gid = h->gid;
    }
    DPM_CUDA_TRY(launch_step(h, p));
    if (tr) cudaEventRecord(tev[4], h->stream);
    h->cur ^= 1;
    if ((s + 1) % 1000 == 0 && s + 1 < nsteps) {  // the reference drains its queue every 1000 steps "to catch errors early" (:437-444)
      int rc = check_device_flags(h);
      if (rc) return rc;
    }
  }
  if (traced >= 0) {
    cudaEventSynchronize(tev[4]);
    float t[4];
    for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
    fprintf(stderr, "[dpm3d] timestep %d of %d (us): rebuild %.1f  units %.1f  contact %.1f  step %.1f  total %.1f\n", traced, nsteps,
            t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f, t[3] * 1e3f, (t[0] + t[1] + t[2] + t[3]) * 1e3f);
    for (auto &e : tev) cudaEventDestroy(e);
  }
  h->stats.steps += (uint64_t)nsteps;
  h->stats.launches += ((repel || attract) ? 4ull : 2ull) * (uint64_t)nsteps;
  return DPM_OK;
}

static int check_device_flags(dpm3d_t *h) {
  NbrState st;
  DPM_CUDA_TRY(cudaMemcpyAsync(&st, h->st, sizeof(NbrState), cudaMemcpyDeviceToHost, h->stream));
  DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->stats.rebuilds = (uint64_t)st.nbuilds;
  h->stats.contact_evals = st.contact_evals;
  h->stats.reserved[0] = st.literal_evals;
  h->stats.reserved[1] = st.fallback_why[0] | (st.fallback_why[1] << 32);  // diagnostics: not-star | near-COM
  h->stats.reserved[2] = st.fallback_why[2] | (st.fallback_why[3] << 32);  //              walk limit | ring limit
  if (st.overflow) return fail(DPM_ERR_RUNTIME, "neighbour candidate list overflow: raise max_candidates (dpm3d_set_neighbor_params)");
  if (st.unit_overflow) return fail(DPM_ERR_RUNTIME, "contact unit queue overflow");
  if (h->nranks > 1) return shard_check(h);
  return DPM_OK;
}

int dpm3d_sync(dpm3d_t *h) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  DeviceGuard guard(h->device);
  return check_device_flags(h);
}

int dpm3d_download(dpm3d_t *h, float *verts4, float *forces4) {
  if (!h || !h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "nothing to download");
  DeviceGuard guard(h->device);
  const size_t bytes = sizeof(float4) * (size_t)h->nc * h->nv;
  if (verts4) DPM_CUDA_TRY(cudaMemcpyAsync(verts4, h->pos[h->cur], bytes, cudaMemcpyDeviceToHost, h->stream));
  if (forces4) DPM_CUDA_TRY(cudaMemcpyAsync(forces4, h->force, bytes, cudaMemcpyDeviceToHost, h->stream));
  return check_device_flags(h);
}

int dpm3d_device_state(dpm3d_t *h, float **verts4_dev, float **forces4_dev) {
  if (!h || !h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "no device state");
  if (verts4_dev) *verts4_dev = reinterpret_cast<float *>(h->pos[h->cur]);
  if (forces4_dev) *forces4_dev = reinterpret_cast<float *>(h->force);
  return DPM_OK;
}

int dpm3d_euler_update(dpm3d_t *h, float *verts4, float *forces4, const float *Kv, const float *Ka, const float *Ks,
                       const float *v0, const float *a0, const float *l0, int nsteps, float dt, float Kre, float Kat, int pbc,
                       float L, float *loop_ms) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  if (nsteps <= 0) return fail(DPM_ERR_INVALID_ARGUMENT, "nsteps must be positive");
  if (!(dt > 0.0f) || dt > 0.1f) return fail(DPM_ERR_INVALID_ARGUMENT, "dt must be positive and reasonable");
  DeviceGuard guard(h->device);
  static const bool trace = getenv("DPM_TRACE") != nullptr;  // host-side phase times of the call on stderr
  using clk = std::chrono::steady_clock;
  const auto t0 = clk::now();
  auto ms_since = [&](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
  double t_up = 0, t_enq = 0, t_run = 0;
  for (int attempt = 0;; attempt++) {
    int rc = upload_common(h, verts4, false, Kv, Ka, Ks, v0, a0, l0, /*wait=*/false);
    if (rc) return rc;
    t_up = ms_since(t0);
    DPM_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    rc = dpm3d_step(h, nsteps, dt, Kre, Kat, pbc, L);
    if (rc) return rc;
    DPM_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    t_enq = ms_since(t0);
    rc = check_device_flags(h);
    t_run = ms_since(t0);
    if (rc == DPM_ERR_RUNTIME && attempt < 3) {  // a capacity was exceeded: grow it and redo from the host state
      char msg[256];
      dpm_last_error(msg, sizeof msg);
      if (trace) fprintf(stderr, "[dpm3d] euler_update attempt %d: %s -> growing\n", attempt, msg);
      if (strstr(msg, "contact unit")) {
        int rc2 = alloc_units(h, h->unit_per_cell * 4);
        if (rc2) return rc2;
        continue;
      }
      if (strstr(msg, "candidate list") && h->K < 128) {
        int rc2 = dpm3d_set_neighbor_params(h, h->skin_rel, std::min(128, h->K * 2));
        if (rc2) return rc2;
        continue;
      }
    }
    if (rc) return rc;
    break;
  }
  if (loop_ms) DPM_CUDA_TRY(cudaEventElapsedTime(loop_ms, h->ev0, h->ev1));
  const int rc = dpm3d_download(h, verts4, forces4);
  if (trace) {
    float dev_ms = 0.f;
    cudaEventElapsedTime(&dev_ms, h->ev0, h->ev1);
    fprintf(stderr, "[dpm3d] euler_update %d steps: upload enqueued %.3f ms, steps enqueued %.3f, steps done %.3f (device loop %.3f), downloaded %.3f\n",
            nsteps, t_up, t_enq, t_run, dev_ms, ms_since(t0));
  }
  return rc;
}

int dpm3d_get_neighbor_artifacts(dpm3d_t *h, dpm_grid_t *grid, int32_t *bin_id, int32_t *order, int32_t *bin_start,
                                 int32_t *cand_count, int32_t *cand) {
  if (!h || !h->uploaded) return fail(DPM_ERR_INVALID_ARGUMENT, "no neighbour state");
  DeviceGuard guard(h->device);
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

int dpm3d_get_cell_bounds(dpm3d_t *h, float *bounds12) {
  if (!h || !h->uploaded || !bounds12) return fail(DPM_ERR_INVALID_ARGUMENT, "no state");
  DeviceGuard guard(h->device);
  std::vector<float> tmp((size_t)4 * BND * h->nc);
  DPM_CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->bnd[h->cur], sizeof(float4) * BND * h->nc, cudaMemcpyDeviceToHost, h->stream));
  DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < h->nc; c++) memcpy(bounds12 + 12 * (size_t)c, tmp.data() + 4 * BND * (size_t)c, sizeof(float) * 12);
  return DPM_OK;
}

int dpm3d_get_stats(dpm3d_t *h, dpm_stats_t *out) {
  if (!h || !out) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = h->stats;
  return DPM_OK;
}
int dpm3d_reset_stats(dpm3d_t *h) {
  if (!h) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL handle");
  memset(&h->stats, 0, sizeof(h->stats));
  return DPM_OK;
}

}  // extern "C"
