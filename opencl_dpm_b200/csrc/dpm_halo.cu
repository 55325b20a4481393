// dpm_halo.cu — multi-GPU slab decomposition of the 3D path with a per-step halo exchange over NCCL.
//
// The reference is single-device; this is new work for BASELINE config E (262,144 cells at
// 2/4/8 GPUs).  One process per GPU; rank r owns a slab of cells (ownership is static — any
// assignment is correct, a spatially compact one keeps the halo small).  Shape forces, the substrate force and
// the integration are per-cell; only the repulsion reads other cells' START-OF-STEP vertices, and it writes the
// vertex's own force only (shaders/Cell3D_Kernel.cl:308), so ghosts are read-only and nothing flows back.
//
// Every step, on the compute stream, with NO host synchronisation:
//   1. shard_prepare_kernel   region box + largest extent/contact pad of the owned cells, local rebuild flag
//   2. ncclAllGather          8 floats per rank (the only collective; it also makes the rebuild decision global)
//   3. shard_select_kernel    if any rank wants a rebuild: new send lists (owned cells whose padded box reaches a
//                             peer's region), in ascending cell order
//   4. shard_pack_kernel      gather {global id, bounds, nv float4 vertices} of the listed cells per peer
//   5. ncclSend / ncclRecv    one fixed-capacity message per peer (ring of slabs: left and right neighbour)
//   6. shard_unpack_kernel    ghosts appended behind the owned cells in pos/bnd/gid; total cell count on device
// then the usual rebuild kernel (device-flag driven, now over owned + ghost cells) and the step kernel (owned).
// Candidate lists are ordered by GLOBAL id, so forces are summed in the same order as on one GPU and a sharded
// run reproduces the single-GPU run bit for bit (tests/test_gpu_multi.py).
//
// NCCL is resolved at run time (dlopen "libnccl.so.2"): a single-GPU user needs no NCCL at all, and inside a
// torch process the already-loaded library is reused.
//
// Round 2: the per-step data path no longer goes through NCCL calls.  At shard_init every rank allocates one "inbox" (two
// message buffers per neighbouring slab, arrival flags, a mailbox of per-rank summaries), the CUDA IPC handles are
// all-gathered over the NCCL communicator ONCE, and every rank maps its peers' inboxes (NVLink / NVSwitch peer memory).
// Per timestep:
//   1. shard_prepare_kernel   as before
//   2. shard_mailbox_kernel   stores the rank's 8-float summary into every peer's mailbox and waits for theirs (replaces
//                             the ncclAllGather: one tiny kernel, ~2 NVLink round trips instead of a collective launch)
//   3. shard_select_kernel    on rebuild only: one 8-CTA cluster builds the send lists (and the per-cell slot table)
//   4. shard_push_kernel      gathers the listed cells and stores them DIRECTLY into the peer's inbox over NVLink — the
//                             actual count of cells, not the capacity — then releases an arrival flag (count, epoch)
//   5. shard_unpack_kernel    acquires the flags of its inbox and appends the ghosts behind the owned cells
// Fused push (default; DPM_HALO_NO_FUSED=1 turns it off): the send lists also exist as a per-cell slot table, and the STEP kernel's
// epilogue stores a listed cell's new positions (a second bulk store out of shared memory), bounds and id straight into the
// neighbour's inbox buffer of the NEXT exchange while the rest of the grid is still integrating.  Step 4 then only releases the
// arrival flags (unless the lists were rebuilt in this exchange or the state was uploaded since: then it moves the cells itself).
// Inboxes and mailboxes are double-buffered by the parity of the exchange epoch: a rank can be at most one exchange ahead
// of a neighbour (it needs that neighbour's ghosts of the previous timestep), so a buffer is never overwritten before it
// has been consumed.  All waits are bounded (a peer that died turns into error 3 instead of a hung GPU).
// DPM_HALO_NCCL=1 selects the round-1 path (ncclAllGather + ncclSend/ncclRecv of capacity-sized messages).
#include <cooperative_groups.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dpm3d_ctx.cuh"

namespace cg = cooperative_groups;

namespace dpm {

constexpr int MAXR = 64;      // ranks
constexpr int GATHER = 8;     // floats per rank in the all-gather
constexpr int HDR_INTS = 16;  // message header

struct ShardDev {
  int n_total;  // owned + ghosts (MUST stay the first member: the rebuild kernel reads it as an int*)
  int n_ghost;
  int send_count[2];
  int error;  // 1: a peer message overflowed the ghost capacity; 2: a cell reaches beyond the adjacent slabs; 3: a peer never answered
  int rebuilds_global;
  float margin;
  unsigned push_ticket[2];  // CTAs of the push kernel that have stored their cell (per peer; self-resetting)
  unsigned long long sent_bytes;
  int lists_dirty;  // the send lists were rebuilt in this exchange (select -> push; cleared by the unpack)
};

// One rank's inbox (peer-mapped by its neighbours).  msg[parity][side]: message from the neighbour on that side;
// flag[parity][side] = (epoch, count) released by the sender's last CTA; mail[parity][rank][GATHER] + mail_epoch[parity][rank].
struct InboxLayout {
  size_t msg_bytes, off_flag, off_mail, off_mail_epoch, total;
  __host__ __device__ size_t msg(int parity, int side) const { return (size_t)(parity * 2 + side) * msg_bytes; }
};
static InboxLayout inbox_layout(size_t msg_bytes, int nranks) {
  InboxLayout L;
  L.msg_bytes = (msg_bytes + 255) & ~(size_t)255;
  L.off_flag = 4 * L.msg_bytes;
  L.off_mail = L.off_flag + 256;
  L.off_mail_epoch = L.off_mail + sizeof(float) * GATHER * 2 * (size_t)nranks;
  L.total = L.off_mail_epoch + sizeof(int) * 2 * (size_t)nranks + 256;
  return L;
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// bounded wait: ~2 s at 2 GHz; a peer that never answers becomes an error flag, not a hung GPU
__device__ __forceinline__ bool wait_epoch(const int *p, int epoch) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) != epoch) {
    if (clock64() - t0 > 4000000000ll) return false;
    __nanosleep(64);
  }
  return true;
}

struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  Nccl() {
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
#define SYM(m, n) m = reinterpret_cast<decltype(m)>(dlsym(lib, n))
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    ok = GetUniqueId && CommInitRank && AllGather && Send && Recv && GroupStart && GroupEnd;
  }
};
static Nccl &nccl() {
  static Nccl n;
  return n;
}
#define DPM_NCCL_TRY(expr)                                                                                       \
  do {                                                                                                           \
    ncclResult_t _r = (expr);                                                                                    \
    if (_r != ncclSuccess)                                                                                       \
      return fail(DPM_ERR_NCCL, std::string(#expr) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(_r) : "nccl error")); \
  } while (0)

// message layout for capacity C cells and nv vertices per cell (all sections 16-byte aligned, C % 4 == 0)
__host__ __device__ inline size_t msg_off_gid() { return sizeof(int) * HDR_INTS; }
__host__ __device__ inline size_t msg_off_bnd(int C) { return msg_off_gid() + sizeof(int) * (size_t)C; }
__host__ __device__ inline size_t msg_off_pos(int C) { return msg_off_bnd(C) + sizeof(float4) * BND * (size_t)C; }
__host__ __device__ inline size_t msg_size(int C, int nv) { return msg_off_pos(C) + sizeof(float4) * (size_t)nv * C; }

// ---- 1. per-rank summary -------------------------------------------------------------------------------------
// PREP_CTAS CTAs reduce their share of the owned cells; the last one to finish (ticket counter) folds the partials and
// writes the rank's summary.  (A single CTA took 264 us for 131,072 cells: it was the largest part of the halo step.)
constexpr int PREP_CTAS = 128;
__global__ void __launch_bounds__(256) shard_prepare_kernel(const float4 *bnd, int n_own, const NbrState *st, float *out, float att_pad_scale,
                                                             float *partial /*[PREP_CTAS][4]*/, unsigned *ticket) {
  __shared__ float s[8][4];
  __shared__ bool last;
  float rlo = INFINITY, rhi = -INFINITY, ext = 0.f, pad = 0.f;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_own; c += gridDim.x * blockDim.x) {
    const float4 lo = bnd[BND * (size_t)c], hi = bnd[BND * (size_t)c + 1];
    rlo = fminf(rlo, lo.x); rhi = fmaxf(rhi, hi.x);
    ext = fmaxf(ext, fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z)));
    pad = fmaxf(pad, hi.w);
    if (att_pad_scale > 0.0f) pad = fmaxf(pad, att_pad_scale * bnd[BND * (size_t)c + 3].w);  // attraction reach (see NbrBuffers)
  }
  rlo = warp_min(rlo); rhi = warp_max(rhi); ext = warp_max(ext); pad = warp_max(pad);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s[w][0] = rlo; s[w][1] = rhi; s[w][2] = ext; s[w][3] = pad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)blockDim.x / 32; i++) {
      s[0][0] = fminf(s[0][0], s[i][0]); s[0][1] = fmaxf(s[0][1], s[i][1]);
      s[0][2] = fmaxf(s[0][2], s[i][2]); s[0][3] = fmaxf(s[0][3], s[i][3]);
    }
    for (int k = 0; k < 4; k++) partial[4 * blockIdx.x + k] = s[0][k];
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  // the last CTA folds the partials in parallel (a single thread walking them took 12 of this kernel's 19 us)
  __threadfence();
  rlo = INFINITY; rhi = -INFINITY; ext = 0.f; pad = 0.f;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    const float4 q = __ldcg(reinterpret_cast<const float4 *>(partial) + b);
    rlo = fminf(rlo, q.x); rhi = fmaxf(rhi, q.y); ext = fmaxf(ext, q.z); pad = fmaxf(pad, q.w);
  }
  rlo = warp_min(rlo); rhi = warp_max(rhi); ext = warp_max(ext); pad = warp_max(pad);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s[w][0] = rlo; s[w][1] = rhi; s[w][2] = ext; s[w][3] = pad; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int i = 0; i < (int)blockDim.x / 32; i++) {
    rlo = fminf(rlo, s[i][0]); rhi = fmaxf(rhi, s[i][1]); ext = fmaxf(ext, s[i][2]); pad = fmaxf(pad, s[i][3]);
  }
  *ticket = 0u;
  out[0] = st->rebuild ? 1.f : 0.f;
  out[1] = rlo; out[2] = rhi; out[3] = ext; out[4] = pad;
  out[5] = out[6] = out[7] = 0.f;
}

// does [lo-m, hi+m] reach region [qlo, qhi] under the periodic image that brings them closest?
__device__ __forceinline__ bool reaches(float lo, float hi, float m, float qlo, float qhi, int pbc, float L) {
  float d = 0.5f * (lo + hi) - 0.5f * (qlo + qhi);
  if (pbc) d -= L * roundf(d / L);
  return fabsf(d) <= 0.5f * (hi - lo) + m + 0.5f * (qhi - qlo);
}

// ---- 3. send lists (one cluster of 8 CTAs, deterministic ascending order) -----------------------------------------
// Round 1 walked the owned cells with a single 256-thread CTA (three barriers and a serial scan per 256 cells: 190 us at
// 32,768 owned cells).  Now every warp of one 8-CTA cluster owns a contiguous segment of cells: pass 1 counts the wanted
// cells per peer, the 256 per-warp counts are exchanged through distributed shared memory, and pass 2 scatters at the
// exclusive offsets — the lists come out in the same ascending order as before.
constexpr int SEL_CLUSTER = 8, SEL_THREADS = 1024, SEL_WARPS = SEL_THREADS / 32, SEL_GW = SEL_CLUSTER * SEL_WARPS;

__global__ void __launch_bounds__(SEL_THREADS) shard_select_kernel(const float4 *bnd, int n_own, NbrState *st, ShardDev *sd, const float *all,
                                                                   int rank, int nranks, int npeers, int peer0, int peer1, int *list0,
                                                                   int *list1, int cap, float skin_rel, int pbc, float L, int2 *push_slot) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int s_cnt[SEL_WARPS][2];  // this CTA's per-warp counts (the other CTAs read them through DSMEM)
  __shared__ int s_all[SEL_GW][2];     // every warp's counts
  bool any = false;
  float ext = 0.f, pad = 0.f;
  for (int r = 0; r < nranks; r++) {
    any |= all[r * GATHER] != 0.f;
    ext = fmaxf(ext, all[r * GATHER + 3]);
    pad = fmaxf(pad, all[r * GATHER + 4]);
  }
  if (!any) return;  // uniform over the cluster
  // lists stay valid while every cell stays in its build box (skin/2 each) and pads stay below the build range
  const float margin = skin_rel * ext + RANGE_HEADROOM * pad + 1e-4f * ext;
  const int crank = (int)cluster.block_rank();
  if (crank == 0 && threadIdx.x == 0) { st->rebuild = 1; sd->margin = margin; sd->rebuilds_global += 1; sd->lists_dirty = 1; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gw = crank * SEL_WARPS + w;
  const int seg = ((n_own + SEL_GW - 1) / SEL_GW + 31) & ~31;  // cells per warp
  const int c_begin = min(n_own, gw * seg), c_end = min(n_own, c_begin + seg);
  const int peers[2] = {peer0, peer1};
  float qlo[2] = {0.f, 0.f}, qhi[2] = {0.f, 0.f};
#pragma unroll
  for (int p = 0; p < 2; p++)
    if (p < npeers) { qlo[p] = all[peers[p] * GATHER + 1]; qhi[p] = all[peers[p] * GATHER + 2]; }
  // pass 1: count
  int cnt[2] = {0, 0};
  bool far = false;
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    const int c = c0 + lane;
    bool want[2] = {false, false};
    if (c < c_end) {
      const float4 lo = bnd[BND * (size_t)c], hi = bnd[BND * (size_t)c + 1];
#pragma unroll
      for (int p = 0; p < 2; p++) want[p] = p < npeers && reaches(lo.x, hi.x, margin, qlo[p], qhi[p], pbc, L);
      // a cell that reaches a slab which is not an adjacent one cannot be served by this ring exchange
      for (int r = 0; r < nranks; r++)
        if (r != rank && r != peer0 && r != peer1 && reaches(lo.x, hi.x, margin, all[r * GATHER + 1], all[r * GATHER + 2], pbc, L)) far = true;
    }
#pragma unroll
    for (int p = 0; p < 2; p++) cnt[p] += __popc(__ballot_sync(0xffffffffu, want[p]));
  }
  if (far) atomicMax(&sd->error, 2);
  if (lane == 0) { s_cnt[w][0] = cnt[0]; s_cnt[w][1] = cnt[1]; }
  cluster.sync();
  if (threadIdx.x < SEL_GW) {
    const int *remote = cluster.map_shared_rank(&s_cnt[0][0], threadIdx.x / SEL_WARPS);
    s_all[threadIdx.x][0] = remote[2 * (threadIdx.x % SEL_WARPS)];
    s_all[threadIdx.x][1] = remote[2 * (threadIdx.x % SEL_WARPS) + 1];
  }
  cluster.sync();  // also keeps every CTA's s_cnt alive until it has been read
  int off[2] = {0, 0};
  for (int i = lane; i < gw; i += 32) { off[0] += s_all[i][0]; off[1] += s_all[i][1]; }
#pragma unroll
  for (int p = 0; p < 2; p++)
    for (int o = 16; o > 0; o >>= 1) off[p] += __shfl_xor_sync(0xffffffffu, off[p], o);
  if (gw == SEL_GW - 1 && lane == 0)
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const int tot = off[p] + cnt[p];
      sd->send_count[p] = min(tot, cap);
      if (tot > cap) atomicMax(&sd->error, 1);
    }
  // pass 2: scatter
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    const int c = c0 + lane;
    bool want[2] = {false, false};
    if (c < c_end) {
      const float4 lo = bnd[BND * (size_t)c], hi = bnd[BND * (size_t)c + 1];
#pragma unroll
      for (int p = 0; p < 2; p++) want[p] = p < npeers && reaches(lo.x, hi.x, margin, qlo[p], qhi[p], pbc, L);
    }
    int slot[2] = {-1, -1};
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const unsigned b = __ballot_sync(0xffffffffu, want[p]);
      if (want[p]) {
        const int pos = off[p] + __popc(b & ((1u << lane) - 1u));
        if (pos < cap) { (p == 0 ? list0 : list1)[pos] = c; slot[p] = pos; }
      }
      off[p] += __popc(b);
    }
    if (c < c_end) push_slot[c] = make_int2(slot[0], slot[1]);  // where the step kernel's epilogue stores this cell (fused push)
  }
}

// ---- 4. pack: one CTA per (peer, slot) ----------------------------------------------------------------------------
__global__ void shard_pack_kernel(const float4 *pos, const float4 *bnd, const int *gid, const ShardDev *sd, const int *list0,
                                  const int *list1, unsigned char *buf0, unsigned char *buf1, int cap, int nv) {
  const int p = blockIdx.x / cap, s = blockIdx.x % cap;
  unsigned char *buf = p == 0 ? buf0 : buf1;
  const int cnt = sd->send_count[p];
  if (s == 0 && threadIdx.x < HDR_INTS) reinterpret_cast<int *>(buf)[threadIdx.x] = threadIdx.x == 0 ? cnt : 0;
  if (s >= cnt) return;
  const int c = (p == 0 ? list0 : list1)[s];
  if (threadIdx.x == 0) reinterpret_cast<int *>(buf + msg_off_gid())[s] = gid[c];
  if (threadIdx.x < BND) reinterpret_cast<float4 *>(buf + msg_off_bnd(cap))[BND * s + threadIdx.x] = bnd[BND * (size_t)c + threadIdx.x];
  float4 *dst = reinterpret_cast<float4 *>(buf + msg_off_pos(cap)) + (size_t)s * nv;
  const float4 *src = pos + (size_t)c * nv;
  for (int v = threadIdx.x; v < nv; v += blockDim.x) dst[v] = src[v];
}

// ---- 6. unpack: ghosts appended behind the owned cells -----------------------------------------------------------
__global__ void shard_unpack_kernel(float4 *pos, float4 *bnd, int *gid, ShardDev *sd, const unsigned char *buf0,
                                    const unsigned char *buf1, int npeers, int n_own, int cap, int nv) {
  const int p = blockIdx.x / cap, s = blockIdx.x % cap;
  const int cnt0 = min(reinterpret_cast<const int *>(buf0)[0], cap);
  const int cnt1 = npeers > 1 ? min(reinterpret_cast<const int *>(buf1)[0], cap) : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { sd->n_ghost = cnt0 + cnt1; sd->n_total = n_own + cnt0 + cnt1; }
  const int cnt = p == 0 ? cnt0 : cnt1;
  if (s >= cnt) return;
  const unsigned char *buf = p == 0 ? buf0 : buf1;
  const int c = n_own + (p == 0 ? 0 : cnt0) + s;
  if (threadIdx.x == 0) gid[c] = reinterpret_cast<const int *>(buf + msg_off_gid())[s];
  if (threadIdx.x < BND) bnd[BND * (size_t)c + threadIdx.x] = reinterpret_cast<const float4 *>(buf + msg_off_bnd(cap))[BND * s + threadIdx.x];
  const float4 *src = reinterpret_cast<const float4 *>(buf + msg_off_pos(cap)) + (size_t)s * nv;
  float4 *dst = pos + (size_t)c * nv;
  for (int v = threadIdx.x; v < nv; v += blockDim.x) dst[v] = src[v];
}

// ---- 2'. mailbox: the rank's summary goes to every peer, theirs are awaited (replaces the ncclAllGather) ---------------
struct PeerPtrs { unsigned char *inbox[MAXR]; };
__global__ void shard_mailbox_kernel(const float *mine, float *all, PeerPtrs peers, unsigned char *my_inbox, InboxLayout L, int rank, int nranks,
                                     int epoch, ShardDev *sd) {
  const int r = threadIdx.x, par = epoch & 1;
  if (r < nranks) {
    float *dst = reinterpret_cast<float *>(peers.inbox[r] + L.off_mail) + ((size_t)par * nranks + rank) * GATHER;
#pragma unroll
    for (int k = 0; k < GATHER; k++) dst[k] = mine[k];
    __threadfence_system();
    st_release_sys(reinterpret_cast<int *>(peers.inbox[r] + L.off_mail_epoch) + par * nranks + rank, epoch);
    // ... and wait for rank r's summary in my own mailbox
    const bool ok = wait_epoch(reinterpret_cast<const int *>(my_inbox + L.off_mail_epoch) + par * nranks + r, epoch);
    if (!ok) sd->error = 3;
    const volatile float *src = reinterpret_cast<const volatile float *>(my_inbox + L.off_mail) + ((size_t)par * nranks + r) * GATHER;
#pragma unroll
    for (int k = 0; k < GATHER; k++) all[r * GATHER + k] = src[k];
  }
}

// ---- 4'. push: the listed cells are stored straight into the peer's inbox over NVLink --------------------------------------
// grid (PUSH_CTAS at most, peers): the CTAs of a peer stride over its slots.  When the step kernel's epilogue has already stored
// this epoch's cells (fused push, same lists), only the arrival flag is left to release.
constexpr int PUSH_CTAS = 296;
__global__ void shard_push_kernel(const float4 *pos, const float4 *bnd, const int *gid, ShardDev *sd, const int *list0, const int *list1,
                                  unsigned char *dst0, unsigned char *dst1, int *flag0, int *flag1, int cap, int nv, int epoch, int fused) {
  const int p = blockIdx.y, G = gridDim.x;
  unsigned char *buf = p == 0 ? dst0 : dst1;
  int *flag = p == 0 ? flag0 : flag1;
  const int cnt = sd->send_count[p];
  const unsigned long long bytes = (unsigned long long)cnt * (sizeof(int) + sizeof(float4) * (BND + (size_t)nv)) + 8ull;
  if (fused && !sd->lists_dirty) {
    // the stores of the step kernel are complete (that grid has finished); the release orders them before the flag
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      flag[1] = cnt;
      __threadfence_system();
      st_release_sys(flag, epoch);
      atomicAdd(&sd->sent_bytes, bytes);
    }
    return;
  }
  bool stored = false;
  for (int s = blockIdx.x; s < cnt; s += G) {
    const int c = (p == 0 ? list0 : list1)[s];
    if (threadIdx.x == 0) reinterpret_cast<int *>(buf + msg_off_gid())[s] = gid[c];
    if (threadIdx.x < BND) reinterpret_cast<float4 *>(buf + msg_off_bnd(cap))[BND * s + threadIdx.x] = bnd[BND * (size_t)c + threadIdx.x];
    float4 *dst = reinterpret_cast<float4 *>(buf + msg_off_pos(cap)) + (size_t)s * nv;
    const float4 *src = pos + (size_t)c * nv;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) dst[v] = src[v];
    stored = true;
  }
  // the last CTA of this peer to finish releases the arrival flag (count first, epoch last).  One system-scope fence per CTA,
  // by thread 0 behind the CTA barrier (cumulative over the other threads' stores), and none for CTAs that stored nothing.
  __syncthreads();
  if (threadIdx.x == 0) {
    if (stored) __threadfence_system();
    const unsigned done = atomicAdd(&sd->push_ticket[p], 1u);
    if (done == (unsigned)G - 1u) {
      sd->push_ticket[p] = 0u;
      __threadfence_system();
      flag[1] = cnt;
      __threadfence_system();
      st_release_sys(flag, epoch);
      atomicAdd(&sd->sent_bytes, bytes);
    }
  }
}

// ---- 6'. unpack out of the own inbox: waits for the senders' flags ------------------------------------------------------
__global__ void shard_unpack_p2p_kernel(float4 *pos, float4 *bnd, int *gid, ShardDev *sd, const unsigned char *buf0, const unsigned char *buf1,
                                        const int *flag0, const int *flag1, int npeers, int n_own, int cap, int nv, int epoch) {
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) {
    int c = 0;
    if (threadIdx.x < npeers) {
      const int *flag = threadIdx.x == 0 ? flag0 : flag1;
      if (!wait_epoch(flag, epoch)) sd->error = 3;
      else c = min(reinterpret_cast<const volatile int *>(flag)[1], cap);
    }
    s_cnt[threadIdx.x] = c;
  }
  __syncthreads();
  const int p = blockIdx.y, G = gridDim.x;
  const int cnt0 = s_cnt[0], cnt1 = s_cnt[1];
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { sd->n_ghost = cnt0 + cnt1; sd->n_total = n_own + cnt0 + cnt1; sd->lists_dirty = 0; }
  const int cnt = p == 0 ? cnt0 : cnt1;
  const unsigned char *buf = p == 0 ? buf0 : buf1;
  for (int s = blockIdx.x; s < cnt; s += G) {
    const int c = n_own + (p == 0 ? 0 : cnt0) + s;
    if (threadIdx.x == 0) gid[c] = __ldcg(reinterpret_cast<const int *>(buf + msg_off_gid()) + s);
    if (threadIdx.x < BND) bnd[BND * (size_t)c + threadIdx.x] = __ldcg(reinterpret_cast<const float4 *>(buf + msg_off_bnd(cap)) + BND * s + threadIdx.x);
    const float4 *src = reinterpret_cast<const float4 *>(buf + msg_off_pos(cap)) + (size_t)s * nv;
    float4 *dst = pos + (size_t)c * nv;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) dst[v] = __ldcg(src + v);
  }
}

__global__ void iota_kernel(int *a, int n, int base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = base + i;
}

int shard_exchange(dpm3d_ctx *h, int pbc, float L) {
  Nccl &N = nccl();
  ncclComm_t comm = static_cast<ncclComm_t>(h->comm);
  float4 *pos = h->pos[h->cur], *bnd = h->bnd[h->cur];
  // DPM_TRACE: device time of each phase of an exchange, events on the stream: the 150th exchange of the process (a step that
  // keeps its lists) and the first three exchanges after it that rebuild the send lists (found by reading the device counter back,
  // which synchronises — tracing perturbs those calls, nothing else)
  static const bool trace = getenv("DPM_TRACE") != nullptr;
  static int ncall = 0, nrebuild_traced = 0;
  if (trace) ++ncall;
  const bool tr = trace && (ncall == 150 || (ncall > 150 && nrebuild_traced < 3));
  int rebuilds_before = 0;
  if (tr) cudaMemcpyAsync(&rebuilds_before, &h->sd->rebuilds_global, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  cudaEvent_t tev[7] = {};
  auto mark = [&](int i) { if (tr) { cudaEventCreate(&tev[i]); cudaEventRecord(tev[i], h->stream); } };
  mark(0);
  shard_prepare_kernel<<<PREP_CTAS, 256, 0, h->stream>>>(bnd, h->nc, h->st, h->gather_send,
                                                         h->att_active ? ATT_REACH / RANGE_HEADROOM * 1.0001f : 0.0f, h->prep_partial, h->prep_ticket);
  mark(1);
  const int epoch = ++h->halo_epoch, par = epoch & 1;
  const InboxLayout IL = inbox_layout(h->msg_bytes, h->nranks);
  if (h->halo_p2p) {
    PeerPtrs pp;
    for (int r = 0; r < h->nranks; r++) pp.inbox[r] = h->peer_inbox[r];
    shard_mailbox_kernel<<<1, (h->nranks + 31) / 32 * 32, 0, h->stream>>>(h->gather_send, h->gather_all, pp, h->inbox, IL, h->rank, h->nranks, epoch, h->sd);
  } else {
    DPM_NCCL_TRY(N.AllGather(h->gather_send, h->gather_all, GATHER, ncclFloat, comm, h->stream));
  }
  mark(2);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(SEL_CLUSTER); cfg.blockDim = dim3(SEL_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SEL_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DPM_CUDA_TRY(cudaLaunchKernelEx(&cfg, shard_select_kernel, (const float4 *)bnd, h->nc, h->st, h->sd, (const float *)h->gather_all, h->rank,
                                    h->nranks, h->npeers, h->peer[0], h->npeers > 1 ? h->peer[1] : -1, h->sendlist[0], h->sendlist[1],
                                    h->ghost_cap, h->skin_rel, pbc, L, reinterpret_cast<int2 *>(h->push_slot)));
  }
  mark(3);
  if (h->halo_p2p) {
    // my peer p sees me on its other side (with two ranks the single peer is both neighbours: side 0)
    unsigned char *dst[2] = {nullptr, nullptr};
    int *dflag[2] = {nullptr, nullptr};
    for (int p = 0; p < h->npeers; p++) {
      const int side = h->npeers == 1 ? 0 : 1 - p;
      dst[p] = h->peer_inbox[h->peer[p]] + IL.msg(par, side);
      dflag[p] = reinterpret_cast<int *>(h->peer_inbox[h->peer[p]] + IL.off_flag) + 2 * (par * 2 + side);
    }
    shard_push_kernel<<<dim3(h->ghost_cap < PUSH_CTAS ? h->ghost_cap : PUSH_CTAS, h->npeers), 256, 0, h->stream>>>(pos, bnd, h->gid, h->sd, h->sendlist[0], h->sendlist[1], dst[0], dst[1],
                                                                        dflag[0], dflag[1], h->ghost_cap, h->nv, epoch, h->pushed_epoch == epoch ? 1 : 0);
    DPM_CUDA_TRY(cudaGetLastError());
    mark(4);
    mark(5);
    const int *mflag = reinterpret_cast<const int *>(h->inbox + IL.off_flag);
    shard_unpack_p2p_kernel<<<dim3(h->ghost_cap < PUSH_CTAS ? h->ghost_cap : PUSH_CTAS, h->npeers), 256, 0, h->stream>>>(pos, bnd, h->gid, h->sd, h->inbox + IL.msg(par, 0), h->inbox + IL.msg(par, 1),
                                                                              mflag + 2 * (par * 2 + 0), mflag + 2 * (par * 2 + 1), h->npeers, h->nc,
                                                                              h->ghost_cap, h->nv, epoch);
    DPM_CUDA_TRY(cudaGetLastError());
  } else {
    shard_pack_kernel<<<h->ghost_cap * h->npeers, 128, 0, h->stream>>>(pos, bnd, h->gid, h->sd, h->sendlist[0], h->sendlist[1],
                                                                        h->sendbuf[0], h->sendbuf[1], h->ghost_cap, h->nv);
    DPM_CUDA_TRY(cudaGetLastError());
    mark(4);
    DPM_NCCL_TRY(N.GroupStart());
    for (int p = 0; p < h->npeers; p++) {
      DPM_NCCL_TRY(N.Send(h->sendbuf[p], h->msg_bytes, ncclChar, h->peer[p], comm, h->stream));
      DPM_NCCL_TRY(N.Recv(h->recvbuf[p], h->msg_bytes, ncclChar, h->peer[p], comm, h->stream));
    }
    DPM_NCCL_TRY(N.GroupEnd());
    mark(5);
    shard_unpack_kernel<<<h->ghost_cap * h->npeers, 128, 0, h->stream>>>(pos, bnd, h->gid, h->sd, h->recvbuf[0], h->recvbuf[1], h->npeers,
                                                                          h->nc, h->ghost_cap, h->nv);
    DPM_CUDA_TRY(cudaGetLastError());
  }
  mark(6);
  if (tr) {
    cudaEventSynchronize(tev[6]);
    int rebuilds_after = 0;
    cudaMemcpy(&rebuilds_after, &h->sd->rebuilds_global, sizeof(int), cudaMemcpyDeviceToHost);
    const bool rebuilt = rebuilds_after != rebuilds_before;
    if (rebuilt && ncall != 150) nrebuild_traced++;
    float t[6];
    for (int i = 0; i < 6; i++) cudaEventElapsedTime(&t[i], tev[i], tev[i + 1]);
    if (ncall == 150 || rebuilt)
      fprintf(stderr, "[dpm3d] rank %d halo exchange %d%s, %s (us): prepare %.1f  allgather|mailbox %.1f  select %.1f  pack|push %.1f  send/recv %.1f  unpack %.1f  total %.1f\n",
              h->rank, ncall, rebuilt ? " (rebuilds the send lists)" : "", h->halo_p2p ? "peer-memory path" : "NCCL path", t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f,
              t[3] * 1e3f, t[4] * 1e3f, t[5] * 1e3f, (t[0] + t[1] + t[2] + t[3] + t[4] + t[5]) * 1e3f);
    for (auto &e : tev) cudaEventDestroy(e);
  }
  h->stats.launches += h->halo_p2p ? 5 : 4;
  if (!h->halo_p2p) h->stats.halo_bytes += (uint64_t)h->msg_bytes * h->npeers;
  return DPM_OK;
}

int shard_check(dpm3d_ctx *h) {
  ShardDev sd;
  DPM_CUDA_TRY(cudaMemcpyAsync(&sd, h->sd, sizeof(ShardDev), cudaMemcpyDeviceToHost, h->stream));
  DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->halo_p2p) h->stats.halo_bytes = sd.sent_bytes;  // actual bytes stored into the peers' inboxes since the last upload
  if (sd.error == 3) return fail(DPM_ERR_NCCL, "halo exchange: a neighbouring rank did not answer within the time limit");
  if (sd.error == 1) return fail(DPM_ERR_RUNTIME, "halo overflow: more boundary cells than max_ghost per neighbouring slab");
  if (sd.error == 2) return fail(DPM_ERR_RUNTIME, "slab decomposition too thin: a cell interacts beyond the adjacent slabs");
  return DPM_OK;
}

void shard_reset_counters(dpm3d_ctx *h) {
  if (h->sd) cudaMemsetAsync(&h->sd->sent_bytes, 0, sizeof(unsigned long long), h->stream);
  h->pushed_epoch = -1;  // new positions were uploaded: whatever a step kernel stored into the peers' inboxes is stale
}

// The step kernel launched next integrates the state that the NEXT exchange (epoch + 1) sends: hand it the position / bounds /
// id areas of that exchange's buffers in the neighbours' inboxes.  Safe against overwriting unread data for the same reason as
// the push kernel: this rank has consumed the neighbour's message of the current epoch, which the neighbour sent after it had
// consumed everything of the epoch before — the previous user of the buffer with this parity.
bool shard_fused_targets(dpm3d_ctx *h, float4 *pos[2], float4 *bnd[2], int *gid[2]) {
  if (!h->halo_p2p || !h->halo_fused || h->halo_epoch <= 0) return false;
  const int epoch = h->halo_epoch + 1, par = epoch & 1;
  const InboxLayout IL = inbox_layout(h->msg_bytes, h->nranks);
  for (int p = 0; p < 2; p++) {
    pos[p] = nullptr; bnd[p] = nullptr; gid[p] = nullptr;
    if (p >= h->npeers) continue;
    const int side = h->npeers == 1 ? 0 : 1 - p;
    unsigned char *msg = h->peer_inbox[h->peer[p]] + IL.msg(par, side);
    gid[p] = reinterpret_cast<int *>(msg + msg_off_gid());
    bnd[p] = reinterpret_cast<float4 *>(msg + msg_off_bnd(h->ghost_cap));
    pos[p] = reinterpret_cast<float4 *>(msg + msg_off_pos(h->ghost_cap));
  }
  h->pushed_epoch = epoch;
  return true;
}

void shard_free(dpm3d_ctx *h) {
  if (h->inbox) {
    for (int r = 0; r < h->nranks; r++)
      if (r != h->rank && h->peer_inbox[r]) cudaIpcCloseMemHandle(h->peer_inbox[r]);
    cudaFree(h->inbox);
    h->inbox = nullptr;
  }
  void *ptrs[] = {h->prep_partial, h->prep_ticket, h->gid, h->sd, h->gather_send, h->gather_all, h->sendbuf[0], h->sendbuf[1], h->recvbuf[0], h->recvbuf[1],
                  h->sendlist[0], h->sendlist[1], h->push_slot};
  for (void *p : ptrs) if (p) cudaFree(p);
  h->gid = nullptr; h->sd = nullptr;
  if (h->comm && nccl().CommDestroy) nccl().CommDestroy(static_cast<ncclComm_t>(h->comm));
  h->comm = nullptr;
}

}  // namespace dpm

using namespace dpm;

extern "C" {

int dpm_nccl_unique_id(uint8_t id[128]) {
  if (!id) return fail(DPM_ERR_INVALID_ARGUMENT, "id is NULL");
  if (!nccl().ok) return fail(DPM_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId u;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  DPM_NCCL_TRY(nccl().GetUniqueId(&u));
  memcpy(id, &u, 128);
  return DPM_OK;
}

// Turns a freshly created handle (ncells = cells OWNED by this rank) into one shard of a slab-decomposed tissue.
// Must be called before the first upload.  max_ghost = ghost cells accepted from EACH neighbouring slab.
int dpm3d_shard_init(dpm3d_t *h, int rank, int nranks, const uint8_t id[128], int max_ghost) {
  if (!h || !id) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL argument");
  if (nranks < 2 || nranks > MAXR || rank < 0 || rank >= nranks || max_ghost < 4)
    return fail(DPM_ERR_INVALID_ARGUMENT, "need 2 <= nranks <= 64, 0 <= rank < nranks, max_ghost >= 4");
  if (h->uploaded || h->nranks > 1) return fail(DPM_ERR_INVALID_ARGUMENT, "dpm3d_shard_init must be the first call on a fresh handle");
  if (!nccl().ok) return fail(DPM_ERR_NCCL, "libnccl.so.2 could not be loaded");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
  max_ghost = (max_ghost + 3) & ~3;
  h->rank = rank; h->nranks = nranks; h->ghost_cap = max_ghost;
  const int left = (rank + nranks - 1) % nranks, right = (rank + 1) % nranks;
  h->npeers = (left == right) ? 1 : 2;
  h->peer[0] = left; h->peer[1] = right;
  h->nslots = h->nc + h->npeers * max_ghost;
  // regrow every per-cell array that also holds ghosts
  const size_t nvert = (size_t)h->nslots * h->nv;
  void **grow[] = {(void **)&h->pos[0], (void **)&h->pos[1], (void **)&h->bnd[0], (void **)&h->bnd[1], (void **)&h->bbox_lo,
                   (void **)&h->bbox_hi, (void **)&h->bin_id, (void **)&h->order, (void **)&h->bin_count, (void **)&h->bin_start,
                   (void **)&h->patch_box};
  h->cap = 4 * h->nslots + 1024;
  const size_t sizes[] = {sizeof(float4) * nvert, sizeof(float4) * nvert, sizeof(float4) * BND * h->nslots, sizeof(float4) * BND * h->nslots,
                          sizeof(float4) * h->nslots, sizeof(float4) * h->nslots, sizeof(int) * h->nslots, sizeof(int) * h->nslots,
                          sizeof(int) * (h->cap + 1), sizeof(int) * (h->cap + 1), sizeof(float4) * 2 * (size_t)h->npatch * h->nslots};
  for (int i = 0; i < 11; i++) {
    if (*grow[i]) cudaFree(*grow[i]);
    *grow[i] = nullptr;
    DPM_CUDA_TRY(cudaMalloc(grow[i], sizes[i]));
    DPM_CUDA_TRY(cudaMemset(*grow[i], 0, sizes[i]));
  }
  h->msg_bytes = msg_size(max_ghost, h->nv);
  DPM_CUDA_TRY(cudaMalloc(&h->gid, sizeof(int) * h->nslots));
  iota_kernel<<<(h->nslots + 255) / 256, 256>>>(h->gid, h->nslots, 0);
  DPM_CUDA_TRY(cudaMalloc(&h->sd, sizeof(ShardDev)));
  DPM_CUDA_TRY(cudaMemset(h->sd, 0, sizeof(ShardDev)));
  DPM_CUDA_TRY(cudaMalloc(&h->gather_send, sizeof(float) * GATHER));
  DPM_CUDA_TRY(cudaMalloc(&h->prep_partial, sizeof(float) * 4 * PREP_CTAS));
  DPM_CUDA_TRY(cudaMalloc(&h->prep_ticket, sizeof(unsigned)));
  DPM_CUDA_TRY(cudaMemset(h->prep_ticket, 0, sizeof(unsigned)));
  DPM_CUDA_TRY(cudaMalloc(&h->gather_all, sizeof(float) * GATHER * nranks));
  for (int p = 0; p < h->npeers; p++) {
    DPM_CUDA_TRY(cudaMalloc(&h->sendbuf[p], h->msg_bytes));
    DPM_CUDA_TRY(cudaMalloc(&h->recvbuf[p], h->msg_bytes));
    DPM_CUDA_TRY(cudaMemset(h->sendbuf[p], 0, h->msg_bytes));
    DPM_CUDA_TRY(cudaMemset(h->recvbuf[p], 0, h->msg_bytes));
    DPM_CUDA_TRY(cudaMalloc(&h->sendlist[p], sizeof(int) * max_ghost));
  }
  DPM_CUDA_TRY(cudaMalloc(&h->push_slot, sizeof(int) * 2 * (size_t)h->nc));
  DPM_CUDA_TRY(cudaMemset(h->push_slot, 0xff, sizeof(int) * 2 * (size_t)h->nc));
  DPM_CUDA_TRY(cudaDeviceSynchronize());
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t comm = nullptr;
  DPM_NCCL_TRY(nccl().CommInitRank(&comm, nranks, u, rank));
  h->comm = comm;
  // ---- peer-memory path: one inbox per rank, mapped by every other rank through CUDA IPC (handles all-gathered once) ----
  h->halo_p2p = false;
  if (!getenv("DPM_HALO_NCCL")) {
    const InboxLayout IL = inbox_layout(h->msg_bytes, nranks);
    h->inbox_bytes = IL.total;
    DPM_CUDA_TRY(cudaMalloc(&h->inbox, IL.total));
    DPM_CUDA_TRY(cudaMemset(h->inbox, 0, IL.total));
    cudaIpcMemHandle_t mine;
    bool ok = cudaIpcGetMemHandle(&mine, h->inbox) == cudaSuccess;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    unsigned char *d_send = nullptr, *d_all = nullptr;
    DPM_CUDA_TRY(cudaMalloc(&d_send, 64 + 64));
    DPM_CUDA_TRY(cudaMalloc(&d_all, (size_t)128 * nranks));
    unsigned char blob[128];
    memset(blob, 0, sizeof blob);
    memcpy(blob, &mine, 64);
    blob[64] = ok ? 1 : 0;
    DPM_CUDA_TRY(cudaMemcpy(d_send, blob, 128, cudaMemcpyHostToDevice));
    DPM_NCCL_TRY(nccl().AllGather(d_send, d_all, 128, ncclChar, comm, h->stream));
    std::vector<unsigned char> all((size_t)128 * nranks);
    DPM_CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost, h->stream));
    DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(d_send); cudaFree(d_all);
    bool all_ok = true;
    for (int r = 0; r < nranks; r++) all_ok = all_ok && all[(size_t)128 * r + 64] == 1;
    if (all_ok) {
      for (int r = 0; r < nranks && all_ok; r++) {
        if (r == rank) { h->peer_inbox[r] = h->inbox; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, &all[(size_t)128 * r], 64);
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; cudaGetLastError(); break; }
        h->peer_inbox[r] = static_cast<unsigned char *>(ptr);
      }
    }
    // every rank must take the same path: agree through one more all-gather of the outcome
    {
      unsigned char *d_f = nullptr, *d_fa = nullptr;
      DPM_CUDA_TRY(cudaMalloc(&d_f, 16));
      DPM_CUDA_TRY(cudaMalloc(&d_fa, (size_t)16 * nranks));
      unsigned char f16[16] = {(unsigned char)(all_ok ? 1 : 0)};
      DPM_CUDA_TRY(cudaMemcpy(d_f, f16, 16, cudaMemcpyHostToDevice));
      DPM_NCCL_TRY(nccl().AllGather(d_f, d_fa, 16, ncclChar, comm, h->stream));
      std::vector<unsigned char> fa((size_t)16 * nranks);
      DPM_CUDA_TRY(cudaMemcpyAsync(fa.data(), d_fa, fa.size(), cudaMemcpyDeviceToHost, h->stream));
      DPM_CUDA_TRY(cudaStreamSynchronize(h->stream));
      cudaFree(d_f); cudaFree(d_fa);
      for (int r = 0; r < nranks; r++) all_ok = all_ok && fa[(size_t)16 * r] == 1;
    }
    h->halo_p2p = all_ok;
    h->halo_fused = all_ok && !getenv("DPM_HALO_NO_FUSED");  // DPM_HALO_NO_FUSED=1: the push kernel moves the cells (round-2 first version)
    if (!all_ok && getenv("DPM_TRACE")) fprintf(stderr, "[dpm3d] rank %d: peer-memory halo path unavailable (CUDA IPC), using the NCCL path\n", rank);
  }
  return DPM_OK;
}

// Global ids of the OWNED cells (length ncells).  Candidate lists are ordered by global id, which makes a sharded
// run sum forces in exactly the single-GPU order.
int dpm3d_set_global_ids(dpm3d_t *h, const int32_t *gid) {
  if (!h || !gid) return fail(DPM_ERR_INVALID_ARGUMENT, "NULL argument");
  if (h->nranks < 2) return fail(DPM_ERR_INVALID_ARGUMENT, "dpm3d_set_global_ids needs a sharded handle (dpm3d_shard_init)");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  cudaError_t e = cudaMemcpy(h->gid, gid, sizeof(int) * h->nc, cudaMemcpyHostToDevice);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(DPM_ERR_CUDA, cudaGetErrorString(e));
  return DPM_OK;
}

}  // extern "C"
