// dpm3d_ctx.cuh — the 3D handle, shared by dpm3d.cu (single-GPU path) and dpm_halo.cu (slab sharding).
#pragma once
#include <vector>

#include "dpm3d_kernels.cuh"

namespace dpm {
struct ShardDev;  // device-side sharding state (dpm_halo.cu)
}

using dpm::NbrState;

struct dpm3d_ctx {
  int device = 0;
  int nc = 0, nv = 0, nf = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  float4 *pos[2] = {nullptr, nullptr};
  float4 *force = nullptr;
  float4 *bnd[2] = {nullptr, nullptr};
  float4 *cellA = nullptr, *cellB = nullptr;
  ushort4 *faces = nullptr;
  uint16_t *ring_nbr = nullptr, *ring_face = nullptr;
  uint8_t *valence = nullptr;
  int ring_stride = 0, min_valence = 0, max_valence = 0;
  ushort4 *face_adj = nullptr;     // static topology tables of the fast contact evaluation (dpm3d_kernels.cuh)
  uint16_t *ring_tab = nullptr;
  uint8_t *ring_end = nullptr;
  uint16_t *dir_table = nullptr;   // rebuilt at every upload from cell 0
  std::vector<uint32_t> h_faces;
  unsigned char *flag[2] = {nullptr, nullptr};  // per-face flags, ping-pong with pos/bnd
  uint2 *vlist = nullptr;                       // per cell: its vertices that have units, (vertex, offset << 8 | count)
  int *vlist_cnt = nullptr;
  int2 *unit_rec = nullptr;  // contact-unit queue (dpm3d_units_kernel -> dpm3d_contact_kernel -> dpm3d_step_kernel)
  float *unit_w = nullptr;
  float4 *unit_att = nullptr;  // allocated when DPM3D_ATTRACT is selected
  int *unit_base = nullptr, *unit_cnt = nullptr;
  int unit_cap = 0, unit_per_cell = 0;
  int contact_grid = 0;
  uint16_t *vorder = nullptr;   // [nv] vertex taken by processing slot r of the step kernel's ring pass (bank-conflict-aware order)
  ushort4 *faces_proc = nullptr;  // [nf] (a, b, c, face id) in the processing order of the step kernel's face pass
  float *terms = nullptr;     // [nc][terms_stride(nf)] signed-volume terms of the new positions (step kernel face pass -> group chain warp)
  float4 *part = nullptr;     // [nc][4 warps][3] per-warp partials of the step kernel's epilogue (folded by the group's chain warp)
  int *grp_done = nullptr;    // [ceil(nc / CHAIN_GROUP)] group completion counters of the step kernel (self-resetting)
  float4 *patch_box = nullptr;  // [nslots][npatch][2] boxes of the face patches of cells that are not star-shaped (winding_patches)
  int npatch = 0;
  // neighbour search
  NbrState *st = nullptr;
  float4 *bbox_lo = nullptr, *bbox_hi = nullptr;
  int *bin_id = nullptr, *order = nullptr, *bin_count = nullptr, *bin_start = nullptr, *cand_count = nullptr, *cand = nullptr;
  float *partial = nullptr;
  int *chunk_sum = nullptr;
  int cap = 0, K = 32, K_alloc = 0;
  float skin_rel = 0.1f;
  int coop_grid = 0;
  int cur = 0;
  unsigned mask = DPM3D_ALL;
  bool uploaded = false;
  int threads = 0, vpt = 0;
  size_t smem = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_up = nullptr;
  float4 *h_cell = nullptr;  // pinned staging for per-cell parameters
  dpm_stats_t stats{};
  bool att_active = false;  // the neighbour lists in use were built with the attraction reach
  int last_pbc = -1;
  float last_L = -1.f;
  int stale_from = -1;
  // ---- slab sharding (dpm_halo.cu); nranks == 1 means unsharded -------------------------------------------
  int nslots = 0;  // allocated cell slots: owned cells + ghost capacity (== nc when unsharded)
  int rank = 0, nranks = 1;
  void *comm = nullptr;  // ncclComm_t
  int npeers = 0, peer[2] = {-1, -1};
  int ghost_cap = 0;  // ghost cells accepted from each peer
  int *gid = nullptr;  // [nslots] global cell ids (owned: set by dpm3d_set_global_ids, ghosts: from the halo messages)
  dpm::ShardDev *sd = nullptr;
  float *gather_send = nullptr, *gather_all = nullptr;
  float *prep_partial = nullptr;  // per-CTA partials of the per-rank summary kernel
  unsigned *prep_ticket = nullptr;
  unsigned char *sendbuf[2] = {nullptr, nullptr}, *recvbuf[2] = {nullptr, nullptr};
  int *sendlist[2] = {nullptr, nullptr};
  size_t msg_bytes = 0;
  // peer-memory halo path (dpm_halo.cu): this rank's inbox, the peers' inboxes mapped through CUDA IPC, the exchange epoch
  unsigned char *inbox = nullptr;
  unsigned char *peer_inbox[64] = {};
  size_t inbox_bytes = 0;
  int halo_epoch = 0;
  bool halo_p2p = false;
  // fused push: the step kernel's epilogue stores the cells on the send lists straight into the neighbours' inboxes
  int *push_slot = nullptr;  // [owned cell][2] slot in the message to peer 0 / 1, or -1 (written with the send lists)
  int pushed_epoch = -1;     // exchange epoch whose inbox buffers the last step kernel has filled (-1: none)
  bool halo_fused = false;
};

namespace dpm {
// one halo exchange of the CURRENT state (pos[cur], bnd[cur]) + global rebuild decision; no-op when unsharded
int shard_exchange(dpm3d_ctx *h, int pbc, float L);
int shard_check(dpm3d_ctx *h);   // after a sync: sharding errors (ghost overflow, slabs too thin)
void shard_free(dpm3d_ctx *h);
void shard_reset_counters(dpm3d_ctx *h);  // per-upload statistics
// peer-memory targets of the step kernel about to be launched (the inbox buffers of the NEXT exchange); false: no fused push
bool shard_fused_targets(dpm3d_ctx *h, float4 *pos[2], float4 *bnd[2], int *gid[2]);
}  // namespace dpm


