// scoring.cu - batched pharmacophore graph-match scoring for sm_100a (one warp per ligand, one lane per conformer).
//
// Replaces, for a whole library in one launch, the reference's per-ligand Python/numba path
//   GraphMatcher.setup/run          src/pmnet/scoring/graph_match.py:85-101
//   scoring_matching_pair/_self     src/pmnet/scoring/match_utils_numba.py:12-231
//   ClusterMatchTree.dfs_run        src/pmnet/scoring/tree.py:55-104
//   GraphMatcher._run_average       src/pmnet/scoring/graph_match.py:103-109
//
// Design (DESIGN.md sections 3-4):
//  * persistent grid (one 32-warp CTA per SM for up to 32 conformers), warps pull ligands from a global counter,
//    optionally through a longest-first permutation (pmnet_cost_order): DFS cost varies by 100x between ligands;
//  * the pharmacophore model (edge table as float4 {mu, 1/sigma, w_b/sigma, w_a w_b/sigma}, cluster tables) is pinned
//    in shared memory once per block (read from global memory instead when it exceeds 100 KB); ligand coordinates
//    stream from HBM as 128 B rows (lane = conformer);
//  * phase 1 evaluates every (ligand cluster i, model cluster k) x (j, l) pair score once. Only what the tree
//    needs is kept: a 32-bit conformer-validity word V per pair (pair score > 0) and, for pairs with V != 0, one
//    128 B row of fp32 scores in a per-warp scratch pool;
//  * phase 2 is the DFS with an explicit stack. Candidate sets are bit masks: a child's masks are
//    parent_mask & alive & V, updated 32 candidates at a time (lane = candidate), the "any conformer left" tests
//    are word != 0, per-depth control state lives in lane-indexed registers (lane d = depth d); the (triangular) mask
//    stack and the per-depth totals sit in shared memory; per-conformer totals are only formed for the node being
//    created: total' = total + self + sum of pair rows with the matched ancestors; all leaf children of a node are
//    evaluated in one pass.
// The discrete decisions (prefilter, sigma^2 < 4, fail counts, score > 0, the < 5 rule) use the same fp32
// operations as the reference (no FMA contraction on those paths), so they agree bit for bit; the accumulated
// scores are fp32 here (fp64 in the reference) and agree to ~1e-6 relative.

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/pmnet_b200.h"

namespace {

constexpr int kMaxDepth = PMNET_MAX_DEPTH;      // levels
constexpr int kSlots = kMaxDepth + 1;           // depths 0..20 (root = depth 0)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxClusterNodes = 255;           // matched model nodes per ligand node (M - 1 is stored in 8 bits)

thread_local char g_err[256] = "";

void set_err(const char* msg) {
  int i = 0;
  for (; msg[i] && i < 255; ++i) g_err[i] = msg[i];
  g_err[i] = 0;
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- workspace layout (shared by host and device)
struct WarpLayout {
  int t_cap;        // max entries (level, model cluster) per ligand
  int pair_cap;     // max pair entries
  int rows;         // pair-score rows (128 B each)
  int dn_cap;       // max ligand nodes in the selected levels (distance table is dn_cap x dn_cap rows)
  int rec_cap;      // max node-match records
  size_t off_rows, off_dist, off_v, off_prow, off_masks, off_tot, off_rowbase, off_srow, off_nmoff, off_geo, off_rec,
      off_entmc, off_entlev, off_nmcnt, off_lnode, off_mlist;
  size_t mlist_cap;
  size_t bytes;     // per warp, multiple of 256
};

// W = 32-conformer words per ligand (1, 2 or 4): every per-conformer row is W * 128 B, every mask W words.
__host__ __device__ inline WarpLayout make_layout(int n_model_clusters, int scratch_rows, int W) {
  WarpLayout L;
  // entries: at most 20 levels x Km model clusters. The default / mid configurations cap it at 1024 (more would also
  // overflow their pair table); the roomy re-run configuration (>= 65536 rows) takes the true worst case
  int t = n_model_clusters * kMaxDepth;
  if (t > 1024 && scratch_rows < 65536) t = 1024;
  L.t_cap = (t + 31) / 32 * 32;
  L.rows = scratch_rows;
  L.pair_cap = scratch_rows * 4;
  // distance table: 64 nodes by default (8192 rows); the roomy re-run configuration (>= 65536 rows) takes 255
  L.dn_cap = scratch_rows >= 65536 ? 255 : (scratch_rows >= 4096 ? 64 : 32);
  L.rec_cap = L.t_cap * 8;
  size_t o = 0;
  L.off_rows = o;    o += (size_t)L.rows * 128 * W;
  L.off_dist = o;    o += (size_t)L.dn_cap * L.dn_cap * 128 * W;
  L.off_v = o;       o += (size_t)L.pair_cap * 4 * W;
  L.off_prow = o;    o += (size_t)L.pair_cap * 4;
  L.off_masks = o;   o += (size_t)kSlots * L.t_cap * 4 * W;
  L.off_tot = o;     o += (size_t)kSlots * 128 * W;
  L.off_rowbase = o; o += (size_t)L.t_cap * 4;
  L.off_srow = o;    o += (size_t)L.t_cap * 4;
  L.off_nmoff = o;   o += (size_t)L.t_cap * 4;
  L.off_geo = o;     o += (size_t)kMaxDepth * 4 * 32 * 4 * W;
  L.off_rec = o;     o += (size_t)L.rec_cap * 8;
  L.off_entmc = o;   o += (size_t)L.t_cap;
  L.off_entlev = o;  o += (size_t)L.t_cap;
  L.off_nmcnt = o;   o += (size_t)L.t_cap;
  L.off_lnode = o;   o += 256;
  L.off_mlist = o;   L.mlist_cap = 65536; o += L.mlist_cap;
  L.bytes = align_up(o, 256);
  return L;
}

constexpr size_t kHeaderBytes = 256;
// header words (zeroed by every call): [0] queue of the specialised kernel, [1] number of heavy ligands, [2] queue of
// the general kernel's pass over the deferred ligands, [3] queue over the heavy list (root tasks), [5] number of
// deferred ligands, [9] / [10] / [16] / [24]: the task queue (kHdrTask*)

// ---------------------------------------------------------------- heavy ligands (task-parallel DFS)
// A ligand whose tree exceeds `heavy_budget` nodes gets a slot in a list of heavy ligands (status PMNET_LIG_HEAVY) and
// is walked by MANY warps. A walker - the warp that has the ligand, or a task warp - that has created `heavy_budget`
// nodes since it started (half of it for a task; then every quarter of it) gives away every not yet visited candidate
// of the SHALLOWEST node on its path that may be given away, one task per candidate, into the task queue, and goes on
// with what it keeps. The queue is consumed by the task kernel (TK = the general kernel's code behind a task queue):
//   * its first launch is also the general kernel's pass over the ligand queue. A warp takes, in this order: a ligand
//     the SPECIALISED kernel gave up (it has no donation code: it abandons the ligand, and one "root" task walks the
//     whole tree again); a ready task; the next ligand of the queue; and once the ligands are gone it waits for tasks
//     (nanosleep polling, bounded) while any walker of the launch is still active - work stealing inside one
//     persistent launch. Tasks before ligands: the donation chains of the heaviest ligands bound the launch
//   * a task {heavy slot, depth j, entries chosen at levels 0..j}: the warp recomputes phase 0 and the part of phase 1
//     its subtree can reach, replays the path to the donor's node at depth j with every choice forced (not counted)
//     and walks the subtree below the chosen candidate, donating in turn into the same queue
//   * the launch is repeated (kTaskRounds in all: what a bounded wait or a full queue left over); the last one does
//     not donate and walks everything it gets to the end.
// A node may give its remaining candidates away only when its None child (tree.py:98: nothing matched, or fewer than 5
// matches on the best path through it) is already ruled out, and with it the None children of all its ancestors: that
// holds as soon as a node with >= 5 matches on its path has been created below it (`deep` bit per depth). Then no
// return value of the donated subtrees is needed by anyone: a task only folds its per-conformer best scores (integer
// atomicMax on non-negative floats: exact, order independent) and its node / leaf counts (atomicAdd) into the ligand's
// accumulator, and pmnet_heavy_finish_kernel writes score, status and statistics - identical to the un-split walk.
constexpr int kHeavyCap = 65536;   // heavy ligands split per call (more are walked to the end where they are)
constexpr int kAccWords = 136;     // per heavy ligand: best[128], nodes, leaves, rows, pairs, needs a root task, -
constexpr int kAccNodes = 128, kAccLeaves = 129, kAccRows = 130, kAccPairs = 131, kAccRoot = 132;
constexpr int kTaskWords = 32;     // [0] heavy slot, [1] depth j, [4 + i] entry chosen at level i <= j (~0: the None child)
constexpr int kTaskCap = 1 << 18;  // tasks per call (a full queue: the walker keeps its candidates)
constexpr int kTaskRounds = 3;
constexpr int kTaskMaxSpins = 200000;  // ~0.2 s of polling before an idle warp gives up (the next launch takes over)
constexpr size_t kHeavyBytes =
    ((size_t)kHeavyCap * 4 * (1 + kAccWords) + (size_t)kTaskCap * (kTaskWords + 1) * 4 + 255) / 256 * 256;
constexpr uint32_t kDefaultHeavyBudget = 1u << 16;
constexpr int kHdrTaskCount = 16;   // header word: task slots reserved so far
constexpr int kHdrTaskHead = 24;    // header word: tasks taken so far
constexpr int kHdrTaskActive = 10;  // header word: walkers of the running task launch
constexpr int kHdrRootHead = 3;     // header word: queue position over the heavy list (root tasks)
constexpr int kHdrTaskBad = 9;      // diagnostics: tasks that could not be replayed (must stay 0)

// Claim a slot of the heavy list for `lig` (whole warp; -1: the list is full) and clear its accumulator. `root` = the
// warp gives the ligand up: a root task walks its whole tree (else the warp keeps walking and only donates).
__device__ __forceinline__ int heavy_append(unsigned char* workspace, uint32_t* heavy_list, uint32_t* heavy_acc,
                                            uint32_t lig, int lane, bool root) {
  unsigned hi = 0;
  if (lane == 0) hi = atomicAdd((unsigned int*)workspace + 1, 1u);
  hi = __shfl_sync(0xffffffffu, hi, 0);
  if (hi >= (unsigned)kHeavyCap) return -1;
  if (lane == 0) heavy_list[hi] = lig;
  uint32_t* acc = heavy_acc + (size_t)hi * kAccWords;
  for (int i = lane; i < kAccWords; i += 32) acc[i] = (i == kAccRoot && root) ? 1u : 0u;
  return (int)hi;
}

// Ligands the specialised kernel defers are appended to a list (when the call has at most this many ligands) so that
// the general kernel hands them out one at a time in the same longest-first order; a larger call falls back to the
// status scan (32 queue positions per warp), which is only balanced when deferred ligands are sparse.
constexpr int kDeferCap = 1 << 20;
constexpr size_t kDeferBytes = (size_t)kDeferCap * 4;

// ---------------------------------------------------------------- shared-memory model image
struct SmemModel {
  int nm, km;
  float4* edge;           // [nm*nm] {mu, r = 1/sigma, wr = w[type b] * r, w[type a] * wr}
  float* cdist;           // [km*km]
  float* csize;           // [km*km]
  float* wnode;           // [nm] weight of each model node's type
  uint16_t* cnode_off;    // [km+1]
  uint8_t* cnodes;        // [cnode_off[km]]
  uint8_t* ntype;         // [nm]
  uint8_t* cmask;         // [km]
};

// Large models (TG = tables in global memory): the edge table (built once per launch in the workspace) and the
// cluster-pair tables stay in HBM / L2 / L1; only the small per-node and per-cluster arrays are pinned.
__host__ __device__ inline size_t smem_model_bytes(int nm, int km, int n_cluster_nodes, bool tables_global) {
  size_t o = 0;
  if (!tables_global) {
    o += (size_t)nm * nm * 16;
    o += (size_t)km * km * 4 * 2;
  }
  o += (size_t)nm * 4;
  o += align_up((size_t)(km + 1) * 2, 4);
  o += align_up((size_t)n_cluster_nodes, 4);
  o += align_up((size_t)nm, 4);
  o += align_up((size_t)km, 4);
  return align_up(o, 16);
}
constexpr size_t kSmemTablesMax = 100 * 1024;  // above this the model tables are read from global memory

// Per-warp shared memory of the DFS: the per-depth conformer totals and the candidate-mask stack of the common case
// (the mask stack is triangular: depth s only keeps the entries of levels >= s). Ligands that need more depth or more
// mask words use the same layout in the global workspace instead.
#ifndef PM_SMEM_TOT_SLOTS
#define PM_SMEM_TOT_SLOTS 12
#endif
#ifndef PM_SMEM_MASK_WORDS
#define PM_SMEM_MASK_WORDS 256
#endif
constexpr int kSmemTotFloats = PM_SMEM_TOT_SLOTS * 32;  // 12 depths of 32 conformers (6 of 64, 3 of 128)
constexpr int kSmemMaskWords = PM_SMEM_MASK_WORDS;
template <int W>
struct WarpSmem {
  float tot[kSmemTotFloats];
  uint32_t mk[kSmemMaskWords];
  int lev_start[kMaxDepth + 1];
  int lev_q[kMaxDepth];       // ligand cluster (global CSR index) of each level
  int lev_nbase[kMaxDepth];   // first local node id of each level
  int lev_run[kMaxDepth];     // pair-table index of the level's first entry against the first later entry
  int moff[kMaxDepth];        // mask stack: word offset of depth s, minus lev_start[s] (index with the entry number)
  int pad[2];
};

struct KernelArgs {
  PmModel model;
  PmLigandBatch batch;
  float w[PMNET_NUM_TYPES];
  float* out_scores;
  float* out_conf;
  int32_t* out_status;
  uint32_t* out_stats;
  unsigned char* workspace;
  int scratch_rows;
  int n_cluster_nodes;
  int conf_stride;  // floats per ligand in out_conf (32 * W)
  int only_status;  // -1: score every ligand; else only the ligands whose out_status holds this code (deferred / re-run)
  int counter_word; // 32-bit word of the workspace header that is this launch's queue counter
  const uint32_t* list;  // non-null: the queue runs over this list of ligands (only_status still filters)
  int list_count_word;   // header word holding the list's length
  int list_cap;
  uint32_t heavy_budget;   // tree nodes after which a walker starts giving subtrees away (0: never)
  uint32_t* heavy_list;    // [kHeavyCap] ligand indices (count in header word 1)
  uint32_t* heavy_acc;     // [kHeavyCap][kAccWords]
  uint32_t* task_buf;      // [kTaskCap][kTaskWords] task descriptors
  uint32_t* task_ready;    // [kTaskCap] set (after a fence) once the descriptor is complete; zeroed by every call
  int task_round;          // task kernel: 0 .. kTaskRounds - 1
  int ligands_first;       // task kernel, first launch: serve the ligand queue (like the general kernel) before the tasks
  WarpLayout ly;    // computed once on the host: the kernel reads the offsets from the constant bank
  const float4* edge_g;  // large models: the edge table in the workspace (build_edge_table_kernel)
};

__device__ __forceinline__ float ld_coord(const float* xyz, int stride, int node, int axis, int lane, bool on) {
  return on ? __ldg(xyz + (size_t)(node * 3 + axis) * stride + lane) : 0.0f;
}

// |p1 - p2| exactly as numpy's fp32 norm (ligand.py:349-351): separate multiplies and adds, IEEE sqrt
__device__ __forceinline__ float norm3(float dx, float dy, float dz) {
  float s = __fmul_rn(dx, dx);
  s = __fadd_rn(s, __fmul_rn(dy, dy));
  s = __fadd_rn(s, __fmul_rn(dz, dz));
  return __fsqrt_rn(s);
}

// exp(-0.5 * s2) on the SFU: ex2.approx(s2 * (-0.5 * log2 e)); relative error ~2^-22 + |x| 2^-24
__device__ __forceinline__ float gauss(float s2) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s2 * -0.72134752044448170368f));
  return r;
}

// Node-match record, 8 bytes (one per ligand node of an entry with >= 1 matched model node, graph_match.py:139-172):
//   .x bits 0-7 local ligand node id, 8-15 M - 1 (M = number of matched model nodes), 16-31 offset of the model-node bytes
//      in `mlist` when M > 4;  .y = the first four matched model nodes, one byte each (all of them when M <= 4).
__device__ __forceinline__ int rec_node(uint2 r) { return r.x & 255u; }
__device__ __forceinline__ int rec_m(uint2 r) { return (int)((r.x >> 8) & 255u) + 1; }
__device__ __forceinline__ int rec_model_node(uint2 r, int a, const uint8_t* __restrict__ mlist) {
  return (a < 4) ? (int)((r.y >> (8 * a)) & 255u) : (int)mlist[(r.x >> 16) + a];
}

template <int W>
__device__ __forceinline__ void eval_edge(const float4 e, const float (&d)[W], float (&lik)[W], int (&npass)[W]) {
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const float s = __fmul_rn(__fsub_rn(d[w], e.x), e.y);
    const float s2 = __fmul_rn(s, s);
    lik[w] = fmaf(e.w, gauss(s2), lik[w]);
    npass[w] += (s2 < 4.0f) ? 1 : 0;
  }
}

// One ligand-node pair against two matched model-node lists (match_utils_numba.py:67-86) for the W conformers of a
// lane: adds likelihood / (M*N) to sc[] and counts the conformers failing the "half of the pairs within 2 sigma" test.
template <int W>
__device__ __forceinline__ void pair_term(const SmemModel& sm, const uint8_t* __restrict__ mlist, const float (&d)[W],
                                          uint2 r1, uint2 r2, float (&sc)[W], int (&nfail)[W]) {
  if (((r1.x | r2.x) & 0xff00u) == 0u) {  // M == 1 and N == 1
    const float4 e = sm.edge[(r1.y & 255u) * sm.nm + (r2.y & 255u)];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const float s = __fmul_rn(__fsub_rn(d[w], e.x), e.y);
      const float s2 = __fmul_rn(s, s);
      nfail[w] += (s2 < 4.0f) ? 0 : 1;
      sc[w] = fmaf(e.w, gauss(s2), sc[w]);
    }
    return;
  }
  const int M = rec_m(r1), N = rec_m(r2);
  int npass[W];
  float lik[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    npass[w] = 0;
    lik[w] = 0.0f;
  }
  if (((r1.x | r2.x) & 0xfe00u) == 0u) {  // M <= 2 and N <= 2
    // by far the most common multi-node case (small model clusters): straight-line code, same evaluation order
    const float4* row0 = sm.edge + (r1.y & 255u) * sm.nm;
    const unsigned b0 = r2.y & 255u, b1 = (r2.y >> 8) & 255u;
    eval_edge<W>(row0[b0], d, lik, npass);
    if (N == 2) eval_edge<W>(row0[b1], d, lik, npass);
    if (M == 2) {
      const float4* row1 = sm.edge + ((r1.y >> 8) & 255u) * sm.nm;
      eval_edge<W>(row1[b0], d, lik, npass);
      if (N == 2) eval_edge<W>(row1[b1], d, lik, npass);
    }
    const int mn = M * N;  // 2 or 4
    const float inv = mn == 2 ? 0.5f : 0.25f;
    const int half = mn >> 1;  // (mn + 1) / 2
#pragma unroll
    for (int w = 0; w < W; ++w) {
      nfail[w] += (npass[w] < half) ? 1 : 0;
      sc[w] = fmaf(lik[w], inv, sc[w]);
    }
    return;
  }
  // larger matches (wide model clusters): plain loops, model nodes from the record word or the spill list
#pragma unroll 1
  for (int a = 0; a < M; ++a) {
    const float4* row = sm.edge + rec_model_node(r1, a, mlist) * sm.nm;
#pragma unroll 1
    for (int b = 0; b < N; ++b) eval_edge<W>(row[rec_model_node(r2, b, mlist)], d, lik, npass);
  }
  const int mn = M * N;
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"((float)mn));
#pragma unroll
  for (int w = 0; w < W; ++w) {
    nfail[w] += (npass[w] < ((mn + 1) >> 1)) ? 1 : 0;
    sc[w] = fmaf(lik[w], inv, sc[w]);
  }
}

// Launch shape. Up to 32 conformers (W = 1): one CTA of 32 warps per SM at 64 registers per thread - the kernel is
// bound by instruction issue and by the latency of its scratch reads, and 32 resident warps hide that best
// (measured: 28 / 38 / 47 / 52.8 M conformers/s at 16 / 16 / 24 / 32 warps per SM across the round-1 versions).
// More conformers per lane need more registers: two CTAs of 8 warps.
#ifndef PM_BLOCK_THREADS
#define PM_BLOCK_THREADS 1024
#endif
#ifndef PM_MIN_BLOCKS
#define PM_MIN_BLOCKS 1
#endif
constexpr int block_threads(int W) { return W == 1 ? PM_BLOCK_THREADS : 256; }
constexpr int min_blocks(int W) { return W == 1 ? PM_MIN_BLOCKS : 2; }

// float4 {mu, 1/sigma, w_b/sigma, w_a*w_b/sigma} of edge (a, b): the same arithmetic as the reference's fp32 inputs
__device__ __forceinline__ float4 edge_entry(const PmModel& gm, const float* w, int i, int nm) {
  const int a = i / nm, b = i % nm;
  const float r = __fdiv_rn(1.0f, gm.edge_sigma[i]);
  const float wr = __fmul_rn(w[gm.node_type[b]], r);
  return make_float4(gm.edge_mu[i], r, wr, __fmul_rn(w[gm.node_type[a]], wr));
}

struct EdgeTableArgs {
  PmModel model;
  float w[PMNET_NUM_TYPES];
  float4* out;
};
__global__ void build_edge_table_kernel(const EdgeTableArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.model.n_nodes * a.model.n_nodes) a.out[i] = edge_entry(a.model, a.w, i, a.model.n_nodes);
}

#include "scoring_fast.cuh"

// TK = task kernel: the general kernel's code behind the task queue of the heavy ligands (see kHeavyCap).
template <int W, bool TG, bool TK = false>
__global__ void __launch_bounds__(block_threads(W), min_blocks(W)) pmnet_score_kernel(const KernelArgs args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int CW = 32 * W;  // conformer slots per row
  constexpr bool DON = true;  // (every instantiation can donate subtrees to the task queue)
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const PmModel& gm = args.model;
  const int NM = gm.n_nodes, KM = gm.n_clusters;
  if (TK && !(args.task_round == 0 && args.ligands_first)) {
    // nothing queued for this launch (the normal case): return before the model is loaded
    const unsigned int* hdr = (const unsigned int*)args.workspace;
    if (hdr[1] == 0u || (args.task_round != 0 && hdr[kHdrTaskHead] >= min(hdr[kHdrTaskCount], (unsigned)kTaskCap))) return;
  }

  // ---- carve shared memory and load the model (once per block)
  SmemModel sm;
  sm.nm = NM;
  sm.km = KM;
  {
    unsigned char* p = smem_raw;
    if (TG) {
      sm.edge = const_cast<float4*>(args.edge_g);
      sm.cdist = const_cast<float*>(gm.cluster_dist);
      sm.csize = const_cast<float*>(gm.cluster_size_sum);
    } else {
      sm.edge = (float4*)p;          p += (size_t)NM * NM * 16;
      sm.cdist = (float*)p;          p += (size_t)KM * KM * 4;
      sm.csize = (float*)p;          p += (size_t)KM * KM * 4;
    }
    sm.wnode = (float*)p;            p += (size_t)NM * 4;
    sm.cnode_off = (uint16_t*)p;     p += align_up((size_t)(KM + 1) * 2, 4);
    sm.cnodes = (uint8_t*)p;         p += align_up((size_t)args.n_cluster_nodes, 4);
    sm.ntype = (uint8_t*)p;          p += align_up((size_t)NM, 4);
    sm.cmask = (uint8_t*)p;          p += align_up((size_t)KM, 4);
  }
  WarpSmem<W>* ws_all = (WarpSmem<W>*)(smem_raw + smem_model_bytes(NM, KM, args.n_cluster_nodes, TG));
  WarpSmem<W>& ws = ws_all[warp_in_block];

  if (!TG) {
    for (int i = threadIdx.x; i < NM * NM; i += blockDim.x) sm.edge[i] = edge_entry(gm, args.w, i, NM);
    for (int i = threadIdx.x; i < KM * KM; i += blockDim.x) {
      sm.cdist[i] = gm.cluster_dist[i];
      sm.csize[i] = gm.cluster_size_sum[i];
    }
  }
  for (int i = threadIdx.x; i < NM; i += blockDim.x) {
    sm.ntype[i] = gm.node_type[i];
    sm.wnode[i] = args.w[gm.node_type[i]];
  }
  for (int i = threadIdx.x; i < KM; i += blockDim.x) sm.cmask[i] = gm.cluster_mask[i];
  for (int i = threadIdx.x; i <= KM; i += blockDim.x) sm.cnode_off[i] = (uint16_t)gm.cluster_node_off[i];
  for (int i = threadIdx.x; i < args.n_cluster_nodes; i += blockDim.x) sm.cnodes[i] = gm.cluster_nodes[i];
  __syncthreads();

  // ---- per-warp scratch
  const WarpLayout& LY = args.ly;
  const int gwarp = blockIdx.x * warps_per_block + warp_in_block;
  unsigned char* wbase = args.workspace + kHeaderBytes + (size_t)gwarp * LY.bytes;
  // opaque to the optimiser: keeps the warp's base in registers instead of re-deriving it from the thread and
  // block ids inside the inner loops when registers are tight
  asm volatile("" : "+l"(wbase));
  __builtin_assume(__isGlobal(wbase));
  float* const rows = (float*)(wbase + LY.off_rows);
  uint32_t* const Vt = (uint32_t*)(wbase + LY.off_v);
  int32_t* const prow = (int32_t*)(wbase + LY.off_prow);
  uint32_t* const masks = (uint32_t*)(wbase + LY.off_masks);
  int32_t* const rowbase = (int32_t*)(wbase + LY.off_rowbase);
  int32_t* const srow = (int32_t*)(wbase + LY.off_srow);
  uint32_t* const nmoff = (uint32_t*)(wbase + LY.off_nmoff);
  float* const geo = (float*)(wbase + LY.off_geo);
  float* const dist = (float*)(wbase + LY.off_dist);
  uint2* const rec = (uint2*)(wbase + LY.off_rec);
  uint8_t* const entmc = wbase + LY.off_entmc;
  uint8_t* const entlev = wbase + LY.off_entlev;
  uint8_t* const nmcnt = wbase + LY.off_nmcnt;
  uint8_t* const lnode = wbase + LY.off_lnode;
  uint8_t* const mlist = wbase + LY.off_mlist;
  unsigned int* const counter = (unsigned int*)args.workspace + args.counter_word;

  const PmLigandBatch& B = args.batch;

  // status-driven queue (only_status >= 0): the warp takes 32 queue positions at a time, keeps those whose status
  // matches, and works through them one by one
  unsigned pend = 0;
  unsigned pend_lig = 0;
  bool task_active = false;                 // TK: this warp is counted in the header's walker count
  bool roots_left = args.task_round == 0;   // TK: the heavy list may still hold ligands that need a root task
  // TK, first launch: the ligand queue comes first (this launch then replaces the general kernel's; its warps go on
  // with the tasks as soon as the ligands run out, while others are still walking theirs)
  bool ligands_left = !TK || (args.task_round == 0 && args.ligands_first != 0);
  for (;;) {
    unsigned int lig = 0;
    int task_j = -1;         // TK: depth of the donor's node whose candidate this task walks (-1: the whole tree)
    unsigned task_h = 0;     // TK: slot of the ligand in the heavy list
    unsigned task_word = 0;  // TK: lane l holds word l of the task descriptor
    uint32_t* task_acc = nullptr;
    bool is_task = false;    // TK: this item is a task (else a ligand of the queue)
    bool is_root = false;    // TK: ... the task of a ligand the specialised kernel gave up (its whole tree)
    unsigned task_slot = 0;  // TK: position of the task in its queue
    if (TK && task_active) {  // the previous item of this warp is finished
      if (lane == 0) atomicSub((unsigned int*)args.workspace + kHdrTaskActive, 1u);
      task_active = false;
    }
    if (TK) {
      // tasks first (they are the long donation chains of the heaviest ligands: the earlier they run, the shorter the
      // launch), ligands when no task is ready; once the ligands are gone, wait for tasks while walkers are active
      unsigned int* const hdr = (unsigned int*)args.workspace;
      unsigned t = 0xffffffffu;
      if (roots_left) {
        // the ligands the specialised kernel gave up (first launch only)
        unsigned nh = hdr[1];
        if (nh > (unsigned)kHeavyCap) nh = kHeavyCap;
        for (;;) {
          if (lane == 0) t = atomicAdd(hdr + kHdrRootHead, 1u);
          t = __shfl_sync(kFull, t, 0);
          if (t >= nh) {
            roots_left = false;
            t = 0xffffffffu;
            break;
          }
          if (args.heavy_acc[(size_t)t * kAccWords + kAccRoot] != 0u) {
            is_root = true;
            break;
          }
        }
      }
      if (!is_root) {
        // the task queue; walkers of this launch (heavy_budget != 0) may still add to it
        if (lane == 0) {
          volatile unsigned int* const vh = hdr;
          for (int spins = 0;;) {
            const unsigned h = vh[kHdrTaskHead];
            const unsigned c = min(vh[kHdrTaskCount], (unsigned)kTaskCap);
            if (h < c) {
              if (atomicCAS(hdr + kHdrTaskHead, h, h + 1u) == h) {
                t = h;
                break;
              }
              continue;
            }
            if (ligands_left || args.heavy_budget == 0u || vh[kHdrTaskActive] == 0u || ++spins > kTaskMaxSpins) break;
            __nanosleep(1000);
          }
          if (t != 0xffffffffu) {
            // the donor reserved the slot first and publishes the descriptor right after
            volatile unsigned int* const rdy = args.task_ready + t;
            for (int spins = 0; *rdy == 0u && spins < (1 << 22); ++spins) __nanosleep(100);
            if (*rdy == 0u) {
              atomicAdd(hdr + kHdrTaskBad, 1u);
              t = 0xfffffffeu;  // (cannot happen) skip it
            }
            __threadfence();
          }
        }
        t = __shfl_sync(kFull, t, 0);
        if (t == 0xfffffffeu) continue;
      }
      if (t != 0xffffffffu) {
        is_task = true;
        task_slot = t;
      } else if (!ligands_left) {
        break;
      }
    }
    if (TK && is_task) {
      const unsigned t = task_slot;
      if (is_root) {
        task_h = t;
      } else {
        task_word = __ldcg(args.task_buf + (size_t)t * kTaskWords + lane);
        task_h = __shfl_sync(kFull, task_word, 0);
        task_j = (int)__shfl_sync(kFull, task_word, 1);
      }
      lig = args.heavy_list[task_h];
      task_acc = args.heavy_acc + (size_t)task_h * kAccWords;
    } else if (args.list != nullptr) {
      bool done = false;
      for (;;) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(kFull, t, 0);
        unsigned nl = ((const unsigned int*)args.workspace)[args.list_count_word];
        if (nl > (unsigned)args.list_cap) nl = args.list_cap;
        if (t >= nl) {
          done = true;
          break;
        }
        lig = args.list[t];
        if (args.only_status < 0 || args.out_status[lig] == args.only_status) break;
      }
      if (done) {
        if (!TK) break;
        ligands_left = false;
        continue;
      }
    } else if (args.only_status < 0) {
      if (lane == 0) {
        lig = atomicAdd(counter, 1u);
        if (B.order != nullptr && lig < (unsigned)B.n_ligands) lig = (unsigned)B.order[lig];
      }
      lig = __shfl_sync(kFull, lig, 0);
      if (lig >= (unsigned)B.n_ligands) {
        if (!TK) break;
        ligands_left = false;
        continue;
      }
    } else {
      bool done = false;
      while (pend == 0u) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(counter, 32u);
        base = __shfl_sync(kFull, base, 0);
        if (base >= (unsigned)B.n_ligands) {
          done = true;
          break;
        }
        const unsigned pos = base + lane;
        bool mine = false;
        if (pos < (unsigned)B.n_ligands) {
          pend_lig = B.order != nullptr ? (unsigned)B.order[pos] : pos;
          mine = args.out_status[pend_lig] == args.only_status;
        }
        pend = __ballot_sync(kFull, mine);
      }
      if (done) {
        if (!TK) break;
        ligands_left = false;
        continue;
      }
      const int src = __ffs(pend) - 1;
      pend &= pend - 1;
      lig = __shfl_sync(kFull, pend_lig, src);
    }

    if (TK) {  // a walker of the task launch (it may donate): idle warps wait for it
      if (lane == 0) atomicAdd((unsigned int*)args.workspace + kHdrTaskActive, 1u);
      task_active = true;
    }
    const int C = B.n_conf[lig];
    float score_out = 0.0f;
    int status = PMNET_LIG_OK;
    uint32_t st_nodes = 0, st_leaves = 0, st_rows = 0, st_pairs = 0;
#ifdef PM_TIMING
    unsigned long long tm0 = 0, tm1 = 0, tm2 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm0));
    tm1 = tm0;
#endif
    float best[W];
#pragma unroll
    for (int w = 0; w < W; ++w) best[w] = 0.0f;
    bool task_bad = false;      // TK: the replayed path did not match the donor's (never: diagnostics only)
    bool heavy = false;         // !TK: the ligand has a slot in the heavy list (this warp donates; results go there)
    bool heavy_denied = false;  // !TK: the list of heavy ligands is full: walk the tree to the end here

    if (C < 1 || C > CW) {
      status = PMNET_LIG_UNSUPPORTED;
    } else {
      const int stride = (C + 3) & ~3;
      const float* xyz = B.coords + (B.coord_off[lig] - B.coord_base);
      const uint8_t* tmask = B.node_type_mask + (B.lig_node_off[lig] - B.node_base);
      const int q0 = B.lig_cluster_off[lig] - B.cluster_base, q1 = B.lig_cluster_off[lig + 1] - B.cluster_base;
      const uint8_t* cl_nodes = B.cluster_nodes - B.cnode_base;  // indexed with the stored (un-rebased) offsets
      bool on[W];
      unsigned cfull[W];  // conformer mask words of this ligand
      bool lane_on = false;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        on[w] = lane + 32 * w < C;
        lane_on |= on[w];
        const int rem = C - 32 * w;
        cfull[w] = rem >= 32 ? kFull : (rem > 0 ? ((1u << rem) - 1u) : 0u);
      }

      // ================= phase 0: levels, entries, node-match records (graph_match.py:85-92, 124-172)
      int L = 0, T = 0, NL = 0;
      int mask_need = 0;  // entries of the triangular mask stack
      bool overflow = false;
      for (int q = q0; q < q1 && L < kMaxDepth && !overflow; ++q) {
        const int c0 = B.cluster_node_off[q], c1 = B.cluster_node_off[q + 1];
        unsigned m = 0;
        for (int i = c0 + lane; i < c1; i += 32) m |= tmask[cl_nodes[i]];
        m = __reduce_or_sync(kFull, m);
        int t_level = T;
        for (int k0 = 0; k0 < KM; k0 += 32) {
          const int k = k0 + lane;
          const bool hit = (k < KM) && (sm.cmask[k] & m);
          const unsigned bal = __ballot_sync(kFull, hit);
          if (T + __popc(bal) > LY.t_cap) {
            overflow = true;
            break;
          }
          if (hit) {
            const int e = T + __popc(bal & ((1u << lane) - 1u));
            entmc[e] = (uint8_t)k;
            entlev[e] = (uint8_t)L;
          }
          T += __popc(bal);
        }
        if (overflow) break;
        if (T > t_level) {
          const int n = c1 - c0;
          if (NL + n > LY.dn_cap) {
            overflow = true;
            break;
          }
          if (lane == 0) {
            ws.lev_start[L] = t_level;
            ws.lev_q[L] = q;
            ws.lev_nbase[L] = NL;
          }
          // the level's nodes get consecutive local ids (cluster order, ligand.py:387-395)
          for (int i = lane; i < n; i += 32) lnode[NL + i] = cl_nodes[c0 + i];
          // cluster centre and size per conformer (ligand.py:458-473), fp32 sequential like numpy
          const float fn = (float)n;
#pragma unroll
          for (int w = 0; w < W; ++w) {
            const int c = lane + 32 * w;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            for (int i = c0; i < c1; ++i) {
              const int node = cl_nodes[i];
              const float x = ld_coord(xyz, stride, node, 0, c, on[w]), y = ld_coord(xyz, stride, node, 1, c, on[w]),
                          z = ld_coord(xyz, stride, node, 2, c, on[w]);
              if (i == c0) {
                cx = x; cy = y; cz = z;
              } else {
                cx = __fadd_rn(cx, x); cy = __fadd_rn(cy, y); cz = __fadd_rn(cz, z);
              }
            }
            cx = __fdiv_rn(cx, fn); cy = __fdiv_rn(cy, fn); cz = __fdiv_rn(cz, fn);
            float sz = 0.f;
            for (int i = c0; i < c1; ++i) {
              const int node = cl_nodes[i];
              const float d = norm3(__fsub_rn(ld_coord(xyz, stride, node, 0, c, on[w]), cx),
                                    __fsub_rn(ld_coord(xyz, stride, node, 1, c, on[w]), cy),
                                    __fsub_rn(ld_coord(xyz, stride, node, 2, c, on[w]), cz));
              sz = (i == c0) ? d : fmaxf(sz, d);
            }
            float* g = geo + (size_t)L * 4 * CW + c;
            g[0] = cx; g[CW] = cy; g[2 * CW] = cz; g[3 * CW] = sz;
          }
          NL += n;
          ++L;
        }
      }
      __syncwarp();
      if (!overflow && L > 0) {
        if (lane == 0) ws.lev_start[L] = T;
        __syncwarp();
        // node-match records, one lane per entry: count, exclusive scan for the offsets, then fill
        uint32_t rec_used = 0, ml_used = 0;
        for (int e0 = 0; e0 < T && !overflow; e0 += 32) {
          const int e = e0 + lane;
          uint32_t nrec = 0, nml = 0;
          bool toobig = false;
          int q = 0, k = 0, nb = 0;
          if (e < T) {
            const int lev = entlev[e];
            q = ws.lev_q[lev];
            nb = ws.lev_nbase[lev];
            k = entmc[e];
            for (int i = B.cluster_node_off[q]; i < B.cluster_node_off[q + 1]; ++i) {
              const unsigned tm = tmask[cl_nodes[i]];
              int M = 0;
              for (int j = sm.cnode_off[k]; j < sm.cnode_off[k + 1]; ++j) M += (tm >> sm.ntype[sm.cnodes[j]]) & 1u;
              if (M > kMaxClusterNodes) toobig = true;
              if (M > 0) ++nrec;
              if (M > 4) nml += M;
            }
            if (nrec > 255) toobig = true;
          }
          uint32_t irec = nrec, iml = nml;
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v1 = __shfl_up_sync(kFull, irec, o), v2 = __shfl_up_sync(kFull, iml, o);
            if (lane >= o) {
              irec += v1;
              iml += v2;
            }
          }
          const uint32_t trec = __shfl_sync(kFull, irec, 31), tml = __shfl_sync(kFull, iml, 31);
          if (__any_sync(kFull, toobig) || rec_used + trec > (uint32_t)LY.rec_cap || ml_used + tml > LY.mlist_cap) {
            overflow = true;
            break;
          }
          if (e < T) {
            uint32_t ro = rec_used + irec - nrec, mo = ml_used + iml - nml;
            nmoff[e] = ro;
            nmcnt[e] = (uint8_t)nrec;
            const int c0 = B.cluster_node_off[q], c1 = B.cluster_node_off[q + 1];
            for (int i = c0; i < c1; ++i) {
              const unsigned tm = tmask[cl_nodes[i]];
              int M = 0;
              uint32_t packed = 0;
              for (int j = sm.cnode_off[k]; j < sm.cnode_off[k + 1]; ++j) {
                const int mn = sm.cnodes[j];
                if ((tm >> sm.ntype[mn]) & 1u) {
                  if (M < 4) packed |= (uint32_t)mn << (8 * M);
                  ++M;
                }
              }
              if (M == 0) continue;
              uint32_t x = 0;
              if (M > 4) {
                x = mo;
                for (int j = sm.cnode_off[k]; j < sm.cnode_off[k + 1]; ++j) {
                  const int mn = sm.cnodes[j];
                  if ((tm >> sm.ntype[mn]) & 1u) mlist[mo++] = (uint8_t)mn;
                }
              }
              rec[ro++] = make_uint2((uint32_t)(nb + (i - c0)) | ((uint32_t)(M - 1) << 8) | (x << 16), packed);
            }
          }
          rec_used += trec;
          ml_used += tml;
        }
        // pair-index base of each entry: V/prow rows of e1 cover all entries of later levels
        if (!overflow) {
          int run = 0;  // running pair count, computed level by level (uniform)
          for (int l = 0; l < L; ++l) {
            const int s = ws.lev_start[l], e_end = ws.lev_start[l + 1];
            const int width = T - e_end;
            for (int e = s + lane; e < e_end; e += 32) rowbase[e] = run + (e - s) * width - e_end;
            if (lane == 0) {
              ws.lev_run[l] = run;
              ws.moff[l] = mask_need - s;
            }
            mask_need += T - s;
            run += (e_end - s) * width;
          }
          st_pairs = (uint32_t)run;
          if (run > LY.pair_cap) overflow = true;
        }
        // ligand node-pair distances (LigandEdge.set_distances, ligand.py:349-351), upper triangle of NL x NL
        if (!overflow) {
          for (int i = 0; i < NL - 1; ++i) {
            const int ni = lnode[i];
            float xi[W], yi[W], zi[W];
#pragma unroll
            for (int w = 0; w < W; ++w) {
              const int c = lane + 32 * w;
              xi[w] = ld_coord(xyz, stride, ni, 0, c, on[w]);
              yi[w] = ld_coord(xyz, stride, ni, 1, c, on[w]);
              zi[w] = ld_coord(xyz, stride, ni, 2, c, on[w]);
            }
            float* drow = dist + ((size_t)i * NL) * CW + lane;
            for (int j = i + 1; j < NL; ++j) {
              const int nj = lnode[j];
#pragma unroll
              for (int w = 0; w < W; ++w) {
                const int c = lane + 32 * w;
                drow[(size_t)j * CW + 32 * w] = norm3(__fsub_rn(xi[w], ld_coord(xyz, stride, nj, 0, c, on[w])),
                                                      __fsub_rn(yi[w], ld_coord(xyz, stride, nj, 1, c, on[w])),
                                                      __fsub_rn(zi[w], ld_coord(xyz, stride, nj, 2, c, on[w])));
              }
            }
          }
        }
        __syncwarp();
      }

      if (overflow) {
        status = PMNET_LIG_OVERFLOW;
      } else if (L == 0) {
        status = PMNET_LIG_EMPTY;
      } else {
        // ================= phase 1: self scores and pair table (graph_match.py:222-279)
        // (row / distance indices fit 32 bits: at most 2^17 rows and 255^2 node pairs of 128 conformer slots)
        const float* const dist_l = dist + lane;
        float* const rows_l = rows + lane;
        int nrows = 0;
        // TK: a task only creates the entries of its path (levels <= task_j) and entries of later levels: the rows
        // of every other entry of the path's levels are never read and are not computed (path_entry(l) = the path's
        // entry at level l, -1 for a None child, -2 for "no restriction")
        auto path_entry = [&](int lev) -> int {
          return (TK && lev <= task_j) ? (int)__shfl_sync(kFull, task_word, 4 + lev) : -2;
        };
        for (int e = 0; e < T && !overflow; ++e) {
          const int cnt = nmcnt[e];
          int r = -1;
          if (TK && task_j >= 0) {
            const int pe_ = path_entry(entlev[e]);
            if (pe_ != -2 && pe_ != e) {
              if (lane == 0) srow[e] = -1;
              continue;
            }
          }
          if (cnt >= 2) {
            float sc[W];
            int nf[W];
#pragma unroll
            for (int w = 0; w < W; ++w) {
              sc[w] = 0.0f;
              nf[w] = 0;
            }
            const uint32_t off = nmoff[e];
            for (int i = 0; i < cnt - 1; ++i) {
              const uint2 r1 = rec[off + i];
              const int drow = rec_node(r1) * NL;
              for (int j = i + 1; j < cnt; ++j) {
                const uint2 r2 = rec[off + j];
                float d[W];
#pragma unroll
                for (int w = 0; w < W; ++w) d[w] = dist_l[(drow + rec_node(r2)) * CW + 32 * w];
                pair_term<W>(sm, mlist, d, r1, r2, sc, nf);
              }
            }
            if (nrows >= LY.rows) {
              overflow = true;
              break;
            }
            r = nrows++;
#pragma unroll
            for (int w = 0; w < W; ++w) rows_l[(r * W + w) * 32] = sc[w];
          }
          if (lane == 0) srow[e] = r;
        }
        for (int i = 0; i < L - 1 && !overflow; ++i) {
          const float* gi = geo + (size_t)i * 4 * CW + lane;
          float cix[W], ciy[W], ciz[W], csi[W];
#pragma unroll
          for (int w = 0; w < W; ++w) {
            cix[w] = gi[32 * w]; ciy[w] = gi[CW + 32 * w]; ciz[w] = gi[2 * CW + 32 * w]; csi[w] = gi[3 * CW + 32 * w];
          }
          const int s1 = ws.lev_start[i], e1_end = ws.lev_start[i + 1];
          const int only1 = path_entry(i);
          if (only1 == -1) continue;  // the path takes the None child at this level
          for (int j = i + 1; j < L && !overflow; ++j) {
            const int only2 = path_entry(j);
            if (only2 == -1) continue;
            const float* gj = geo + (size_t)j * 4 * CW + lane;
            float ldist[W], lsize[W];
#pragma unroll
            for (int w = 0; w < W; ++w) {
              ldist[w] = norm3(__fsub_rn(cix[w], gj[32 * w]), __fsub_rn(ciy[w], gj[CW + 32 * w]),
                               __fsub_rn(ciz[w], gj[2 * CW + 32 * w]));
              lsize[w] = __fadd_rn(csi[w], gj[3 * CW + 32 * w]);
            }
            const int s2 = ws.lev_start[j], e2_end = ws.lev_start[j + 1];
            for (int e1 = s1; e1 < e1_end && !overflow; ++e1) {
              if (only1 != -2 && e1 != only1) continue;
              const int k = entmc[e1];
              const int cnt1 = nmcnt[e1];
              const uint32_t off1 = nmoff[e1];
              const int pb = rowbase[e1];
              const float* cd = sm.cdist + k * KM;
              const float* cs = sm.csize + k * KM;
              for (int e2 = s2; e2 < e2_end; ++e2) {
                if (only2 != -2 && e2 != only2) continue;
                const int l = entmc[e2];
                // cluster prefilter (graph_match.py:263-268): min_c(|d_lig - d_mod| - size_lig) > size_mod
                const float cdl = cd[l], csl = cs[l];
                bool far = true;
#pragma unroll
                for (int w = 0; w < W; ++w)
                  far &= !on[w] || (__fsub_rn(fabsf(__fsub_rn(ldist[w], cdl)), lsize[w]) > csl);
                unsigned valid[W];
#pragma unroll
                for (int w = 0; w < W; ++w) valid[w] = 0;
                bool anyvalid = false;
                int r = -1;
                if (!__all_sync(kFull, far)) {
                  const int cnt2 = nmcnt[e2];
                  const uint32_t off2 = nmoff[e2];
                  const int thr2 = cnt1 * cnt2;  // fail <= 0.5*cnt1*cnt2  <=>  2*fail <= cnt1*cnt2
                  float sc[W];
                  int nfail[W];
#pragma unroll
                  for (int w = 0; w < W; ++w) {
                    sc[w] = 0.0f;
                    nfail[w] = 0;
                  }
                  bool dead = false;
                  for (int a = 0; a < cnt1; ++a) {
                    const uint2 r1 = rec[off1 + a];
                    const float* const drow = dist_l + (unsigned)(rec_node(r1) * NL * CW);
                    for (int b = 0; b < cnt2; ++b) {
                      const uint2 r2 = rec[off2 + b];
                      float d[W];
#pragma unroll
                      for (int w = 0; w < W; ++w) d[w] = drow[(unsigned)(rec_node(r2) * CW + 32 * w)];
                      pair_term<W>(sm, mlist, d, r1, r2, sc, nfail);
                    }
                    // every conformer already failed: the pair is invalid whatever follows
                    // (match_utils_numba.py:191-192 tests this after every term; the outcome is the same)
                    bool alldead = true;
#pragma unroll
                    for (int w = 0; w < W; ++w) alldead &= !on[w] || (2 * nfail[w] > thr2);
                    if (__all_sync(kFull, alldead)) {
                      dead = true;
                      break;
                    }
                  }
                  if (!dead) {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                      valid[w] = __ballot_sync(kFull, on[w] && (2 * nfail[w] <= thr2) && (sc[w] > 0.0f));
                      anyvalid |= valid[w] != 0;
                    }
                  }
                  if (anyvalid) {
                    if (nrows >= LY.rows) {
                      overflow = true;
                      break;
                    }
                    r = nrows++;
#pragma unroll
                    for (int w = 0; w < W; ++w) rows_l[(r * W + w) * 32] = sc[w];
                  }
                }
                if (lane == 0) {
#pragma unroll
                  for (int w = 0; w < W; ++w) Vt[(size_t)(pb + e2) * W + w] = valid[w];
                  prow[pb + e2] = r;
                }
              }
            }
          }
        }
        st_rows = (uint32_t)nrows;
        __syncwarp();

        if (overflow) {
          status = PMNET_LIG_OVERFLOW;
        } else {
#ifdef PM_TIMING
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm1));
#endif
          // ================= phase 2: DFS (tree.py:55-104) with an explicit stack; lane d holds depth d's state
          int st_cursor = 0, st_maxm = 0, st_nchild = 0, st_phase = 0, st_nmatch = 0, st_entry = -1, st_mslot = 0,
              st_tslot = 0, st_pbase = 0;
          unsigned st_alive[W];
          // stacks in shared memory when they fit (the common case), else in the workspace; same layout
          uint32_t* const mk = (mask_need * W <= kSmemMaskWords) ? ws.mk : masks;
          float* const tot_l = ((L * CW <= kSmemTotFloats) ? ws.tot : (float*)(wbase + LY.off_tot)) + lane;
#pragma unroll
          for (int w = 0; w < W; ++w) {
            st_alive[w] = 0;
            for (int e = lane; e < T; e += 32) mk[e * W + w] = cfull[w];  // depth 0: moff[0] = 0
            tot_l[32 * w] = 0.0f;
          }
          if (lane == 0) {
            st_cursor = 0;  // lev_start[0]
#pragma unroll
            for (int w = 0; w < W; ++w) st_alive[w] = cfull[w];
          }
          __syncwarp();
          st_nodes = (TK && task_j >= 0) ? 0u : 1u;  // a task counts what it creates below the replayed path
          uint32_t deep = 0;   // TK: bit s = a node with >= 5 matches on its path exists below the node at depth s
          uint32_t mark = 0;   // TK: st_nodes at the last donation
#ifndef PM_TASK_FIRST_SHIFT
#define PM_TASK_FIRST_SHIFT 1  // measured on the dense model: 526 / 499 / 536 ms for 0 / 1 / 2
#endif
          // nodes until the next donation: the budget at first (a donated task: budget >> PM_TASK_FIRST_SHIFT), then a
          // quarter of it
          uint32_t grain = (TK && task_j >= 0) ? (args.heavy_budget >> PM_TASK_FIRST_SHIFT) : args.heavy_budget;
          int d = 0;
          for (;;) {
            // node at depth d; its children live at level y = d
            const int y = d;
            const int phase = __shfl_sync(kFull, st_phase, d);
            // TK: the nodes at depth <= task_j are the donor's path: one forced child each, then the task is over
            bool forced = false;
            int forced_p = 0;
            if (TK && d <= task_j) {
              if (phase != 0 || __shfl_sync(kFull, st_nchild, d) != 0) break;
              forced = true;
              forced_p = (int)__shfl_sync(kFull, task_word, 4 + d);
            }
            const int mslot = __shfl_sync(kFull, st_mslot, d);
            const int tslot = __shfl_sync(kFull, st_tslot, d);
            const uint32_t* pm = mk + ws.moff[mslot] * W;
            // lanes 1..d that hold a matched ancestor (or this node itself). A candidate's mask is the AND of the
            // validity words of its pairs with every matched ancestor, so each of those pairs has a score row.
            const bool is_anc = lane >= 1 && lane <= d && st_entry >= 0;
            bool do_return = false;
            if (phase == 0 && y == L - 1) {
              // ---- every child of this node is a leaf (graph_match.py:103-109): take them all in one pass
              const int end = T;
              float tt[W];
#pragma unroll
              for (int w = 0; w < W; ++w) tt[w] = tot_l[tslot * CW + 32 * w];
              int nleaf = 0;
              for (int cur = ws.lev_start[y]; cur < end; cur += 32) {
                const int idx = cur + lane;
                unsigned mw[W];
                unsigned anyw = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                  mw[w] = (idx < end) ? pm[idx * W + w] : 0u;
                  anyw |= mw[w];
                }
                unsigned bal = __ballot_sync(kFull, anyw != 0u);
                if (bal == 0u) continue;
                // row indices of the next candidate are fetched while the current one is summed
                int nf = cur + __ffs(bal) - 1;
                int myrow_n = is_anc ? prow[st_pbase + nf] : -1;
                int sr_n = srow[nf];
                while (bal) {
                  const int src = __ffs(bal) - 1;
                  bal &= bal - 1;
                  const int myrow = myrow_n, sr = sr_n;
                  if (bal) {
                    nf = cur + __ffs(bal) - 1;
                    myrow_n = is_anc ? prow[st_pbase + nf] : -1;
                    sr_n = srow[nf];
                  }
                  float t[W], self[W];
#pragma unroll
                  for (int w = 0; w < W; ++w) {
                    self[w] = (sr >= 0) ? rows_l[(sr * W + w) * 32] : 0.0f;
                    t[w] = 0.0f;
                  }
                  // lanes 1..d hold the pair-row index of their depth (-1: unmatched); four loads in flight per round
                  for (int d0 = 1; d0 <= d; d0 += 4) {
                    int rr[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) rr[u] = __shfl_sync(kFull, myrow, d0 + u);
                    float vv[4][W];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                      for (int w = 0; w < W; ++w)
                        vv[u][w] = (rr[u] >= 0) ? rows_l[(rr[u] * W + w) * 32] : 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                      for (int w = 0; w < W; ++w) t[w] += vv[u][w];
                  }
#pragma unroll
                  for (int w = 0; w < W; ++w) {
                    const unsigned al = __shfl_sync(kFull, mw[w], src);
                    if ((al >> lane) & 1u) best[w] = fmaxf(best[w], (tt[w] + self[w]) + t[w]);
                  }
                  ++nleaf;
                }
              }
              st_nodes += nleaf;
              st_leaves += nleaf;
              const int nmatch = __shfl_sync(kFull, st_nmatch, d);
              if (DON && nleaf > 0 && nmatch + 1 >= PMNET_MIN_MATCHES) deep |= (2u << d) - 1u;
              if (lane == d) st_maxm = nleaf > 0 ? 1 : 0;
              if (nleaf == 0 || nmatch + 1 < PMNET_MIN_MATCHES) {
                // the None leaf (tree.py:98)
                ++st_nodes;
                ++st_leaves;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                  const unsigned al = __shfl_sync(kFull, st_alive[w], d);
                  if ((al >> lane) & 1u) best[w] = fmaxf(best[w], tt[w]);
                }
              }
              do_return = true;
            } else if (phase == 0) {
              int cur = __shfl_sync(kFull, st_cursor, d);
              const int end = ws.lev_start[y + 1];
              if (forced) cur = forced_p < 0 ? end : forced_p;
              int found = -1;
              unsigned alive2[W];
#pragma unroll
              for (int w = 0; w < W; ++w) alive2[w] = 0;
              while (cur < end) {
                const int idx = cur + lane;
                unsigned mw[W];
                unsigned anyw = 0;
#pragma unroll
                for (int w = 0; w < W; ++w) {
                  mw[w] = (idx < end) ? pm[idx * W + w] : 0u;
                  anyw |= mw[w];
                }
                const unsigned bal = __ballot_sync(kFull, anyw != 0u);
                if (bal) {
                  const int src = __ffs(bal) - 1;
                  found = cur + src;
#pragma unroll
                  for (int w = 0; w < W; ++w) alive2[w] = __shfl_sync(kFull, mw[w], src);
                  break;
                }
                cur += 32;
              }
              if (forced && forced_p >= 0 && found != forced_p) {
                task_bad = true;  // cannot happen: the path is recomputed from the same inputs
                break;
              }
              if (found >= 0) {
                // ---- matched child (y, found): ClusterMatchTree.__init__ (tree.py:33-41); it is not a leaf
                if (!(TK && d < task_j)) ++st_nodes;
                if (lane == d) {
                  st_cursor = forced ? end : found + 1;
                  st_nchild += 1;
                }
                if (DON && __shfl_sync(kFull, st_nmatch, d) + 1 >= PMNET_MIN_MATCHES) deep |= (2u << d) - 1u;
                bool donate = DON && args.heavy_budget != 0u && d > task_j && st_nodes - mark > grain && !heavy_denied;
                if (donate && !is_task && !heavy) {
                  // first time over the budget: the ligand needs a slot in the heavy list (its results go there)
                  const int hi = heavy_append(args.workspace, args.heavy_list, args.heavy_acc, lig, lane, false);
                  if (hi < 0) {
                    heavy_denied = true;
                    donate = false;
                  } else {
                    heavy = true;
                    task_h = (unsigned)hi;
                    task_acc = args.heavy_acc + (size_t)hi * kAccWords;
                  }
                }
                if (donate) {
                  // ---- donate the unvisited candidates of the shallowest node that may give them away
                  mark = st_nodes;
                  const int pe = __shfl_sync(kFull, st_entry, (lane - 3) & 31);  // lane 4 + i: the entry chosen at level i
                  unsigned int* const qcount = (unsigned int*)args.workspace + kHdrTaskCount;
                  uint32_t* const qout = args.task_buf;
                  for (int s = task_j + 1; s <= d; ++s) {
                    if (!((deep >> s) & 1u)) continue;
                    const int cs = __shfl_sync(kFull, st_cursor, s), es = ws.lev_start[s + 1];
                    const uint32_t* pms = mk + ws.moff[__shfl_sync(kFull, st_mslot, s)] * W;
                    int cnt = 0;
                    for (int b = cs; b < es; b += 32) {
                      unsigned anyw = 0;
                      if (b + lane < es)
#pragma unroll
                        for (int w = 0; w < W; ++w) anyw |= pms[(b + lane) * W + w];
                      cnt += __popc(__ballot_sync(kFull, anyw != 0u));
                    }
                    if (cnt == 0) continue;
                    unsigned base = 0xffffffffu;  // reserve cnt queue slots (all or nothing)
                    if (lane == 0) {
                      unsigned old = *(volatile unsigned int*)qcount;
                      while (old + (unsigned)cnt <= (unsigned)kTaskCap) {
                        const unsigned prev = atomicCAS(qcount, old, old + (unsigned)cnt);
                        if (prev == old) {
                          base = old;
                          break;
                        }
                        old = prev;
                      }
                    }
                    base = __shfl_sync(kFull, base, 0);
                    if (base == 0xffffffffu) break;  // the queue is full: keep everything
                    for (int b = cs; b < es; b += 32) {
                      unsigned anyw = 0;
                      if (b + lane < es)
#pragma unroll
                        for (int w = 0; w < W; ++w) anyw |= pms[(b + lane) * W + w];
                      for (unsigned bal = __ballot_sync(kFull, anyw != 0u); bal; bal &= bal - 1) {
                        const int c = b + __ffs(bal) - 1;
                        uint32_t wv = 0;
                        if (lane == 0) wv = task_h;
                        else if (lane == 1) wv = (uint32_t)s;
                        else if (lane >= 4 && lane - 4 < s) wv = (uint32_t)pe;
                        else if (lane - 4 == s) wv = (uint32_t)c;
                        qout[(size_t)base * kTaskWords + lane] = wv;
                        __threadfence();  // the descriptor before its ready flag
                        __syncwarp();
                        if (lane == 0) *(volatile uint32_t*)(args.task_ready + base) = 1u;
                        ++base;
                      }
                    }
                    if (lane == s) st_cursor = es;
                    grain = args.heavy_budget >> 2;
                    break;  // one node per donation
                  }
                }
                // everything that depends only on `found` is requested first: the ancestors' pair-row indices,
                // the self row index, the parent's totals and the first chunk of the child's masks
                const int myrow = is_anc ? prow[st_pbase + found] : -1;
                const int sr = srow[found];
                const int pbc = ws.lev_run[y] + (found - ws.lev_start[y]) * (T - end) - end;
                uint32_t* nm_ = mk + ws.moff[d + 1] * W;
                const int e2a = end + lane;
                unsigned pmv[W], vtv[W];
#pragma unroll
                for (int w = 0; w < W; ++w) {
                  pmv[w] = (e2a < T) ? pm[e2a * W + w] : 0u;
                  vtv[w] = (e2a < T) ? Vt[(pbc + e2a) * W + w] : 0u;
                }
                float t[W];
#pragma unroll
                for (int w = 0; w < W; ++w) {
                  t[w] = tot_l[tslot * CW + 32 * w];
                  if (sr >= 0) t[w] += rows_l[(sr * W + w) * 32];
                }
                // pair rows with the matched ancestors (tree.py:78-82), four independent row loads per round
                float acc[W];
#pragma unroll
                for (int w = 0; w < W; ++w) acc[w] = 0.0f;
                for (int d0 = 1; d0 <= d; d0 += 4) {
                  int rr[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) rr[u] = __shfl_sync(kFull, myrow, d0 + u);
                  float vv[4][W];
#pragma unroll
                  for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int w = 0; w < W; ++w)
                      vv[u][w] = (rr[u] >= 0) ? rows_l[(rr[u] * W + w) * 32] : 0.0f;
#pragma unroll
                  for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int w = 0; w < W; ++w) acc[w] += vv[u][w];
                }
                {
                  // no candidate left at any later level: the child's subtree is the chain of None nodes down to the
                  // None leaf (tree.py:98 with nchild == 0 at every level) - account for it without walking it
                  unsigned any0 = 0;
#pragma unroll
                  for (int w = 0; w < W; ++w) any0 |= pmv[w] & alive2[w] & vtv[w];
                  if (T - end <= 32 && !__any_sync(kFull, any0 != 0u)) {
                    st_nodes += L - d - 1;
                    ++st_leaves;
#pragma unroll
                    for (int w = 0; w < W; ++w)
                      if ((alive2[w] >> lane) & 1u) best[w] = fmaxf(best[w], t[w] + acc[w]);
                    if (lane == d) st_maxm = max(st_maxm, 1);
                    continue;
                  }
                }
                const int nmatch = __shfl_sync(kFull, st_nmatch, d) + 1;
                if (d + 2 == L && T - end <= 32) {
                  // ---- the child's own children are leaves and their masks are in registers (lane = entry of the last
                  // level): evaluate them here, exactly like the leaf pass above with the child as one more ancestor,
                  // instead of pushing the child, storing its masks / totals and coming back for them
                  unsigned nmw[W];
                  unsigned anyw = 0;
                  float tt[W];
#pragma unroll
                  for (int w = 0; w < W; ++w) {
                    nmw[w] = pmv[w] & alive2[w] & vtv[w];
                    anyw |= nmw[w];
                    tt[w] = t[w] + acc[w];
                  }
                  unsigned bal = __ballot_sync(kFull, anyw != 0u);  // not empty: the dead-end case was taken above
                  const bool is_anc2 = is_anc || lane == d + 1;
                  const int pb2 = (lane == d + 1) ? pbc : st_pbase;
                  int nleaf = 0;
                  int nf = end + __ffs(bal) - 1;
                  int myrow_n = is_anc2 ? prow[pb2 + nf] : -1;
                  int sr_n = srow[nf];
                  while (bal) {
                    const int src = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const int myrow2 = myrow_n, sr2 = sr_n;
                    if (bal) {
                      nf = end + __ffs(bal) - 1;
                      myrow_n = is_anc2 ? prow[pb2 + nf] : -1;
                      sr_n = srow[nf];
                    }
                    float tl[W], self[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                      self[w] = (sr2 >= 0) ? rows_l[(sr2 * W + w) * 32] : 0.0f;
                      tl[w] = 0.0f;
                    }
                    for (int d0 = 1; d0 <= d + 1; d0 += 4) {
                      int rr[4];
#pragma unroll
                      for (int u = 0; u < 4; ++u) rr[u] = __shfl_sync(kFull, myrow2, d0 + u);
                      float vv[4][W];
#pragma unroll
                      for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int w = 0; w < W; ++w)
                          vv[u][w] = (rr[u] >= 0) ? rows_l[(rr[u] * W + w) * 32] : 0.0f;
#pragma unroll
                      for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int w = 0; w < W; ++w) tl[w] += vv[u][w];
                    }
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                      const unsigned al = __shfl_sync(kFull, nmw[w], src);
                      if ((al >> lane) & 1u) best[w] = fmaxf(best[w], (tt[w] + self[w]) + tl[w]);
                    }
                    ++nleaf;
                  }
                  st_nodes += nleaf;
                  st_leaves += nleaf;
                  if (DON && nmatch + 1 >= PMNET_MIN_MATCHES) deep |= (2u << d) - 1u;  // (nleaf > 0 here)
                  if (nmatch + 1 < PMNET_MIN_MATCHES) {
                    // the child's None leaf (tree.py:98: too few matches on the path)
                    ++st_nodes;
                    ++st_leaves;
#pragma unroll
                    for (int w = 0; w < W; ++w)
                      if ((alive2[w] >> lane) & 1u) best[w] = fmaxf(best[w], tt[w]);
                  }
                  if (lane == d) st_maxm = max(st_maxm, 2);  // the child returns 1 (a matched leaf) + 1 (itself)
                  continue;
                }
                // the child's candidate masks: parent mask & conformers alive in the child & pair validity.
                // (depth d + 1's slot was last read by other lanes while the previous child's subtree was walked)
                __syncwarp();
                if (e2a < T) {
#pragma unroll
                  for (int w = 0; w < W; ++w) nm_[e2a * W + w] = pmv[w] & alive2[w] & vtv[w];
                }
                for (int e2 = e2a + 32; e2 < T; e2 += 32) {
#pragma unroll
                  for (int w = 0; w < W; ++w)
                    nm_[e2 * W + w] = pm[e2 * W + w] & alive2[w] & Vt[(pbc + e2) * W + w];
                }
#pragma unroll
                for (int w = 0; w < W; ++w) tot_l[(d + 1) * CW + 32 * w] = t[w] + acc[w];
                if (DON) deep = (deep & ((2u << d) - 1u)) | (nmatch >= PMNET_MIN_MATCHES ? (2u << d) : 0u);
                if (lane == d + 1) {
                  st_cursor = end;
                  st_maxm = 0;
                  st_nchild = 0;
                  st_phase = 0;
                  st_nmatch = nmatch;
                  st_entry = found;
                  st_mslot = d + 1;
                  st_tslot = d + 1;
                  st_pbase = pbc;
#pragma unroll
                  for (int w = 0; w < W; ++w) st_alive[w] = alive2[w];
                }
                __syncwarp();
                d = d + 1;
                continue;
              }
              // ---- no matched child left: None child iff nothing matched or too few matches so far (tree.py:98)
              const int nchild = __shfl_sync(kFull, st_nchild, d);
              const int nmatch = __shfl_sync(kFull, st_nmatch, d);
              const int maxm = __shfl_sync(kFull, st_maxm, d);
              // (TK, replayed path: the None child is walked iff the donor's path goes through it)
              if (forced ? forced_p < 0 : (nchild == 0 || nmatch + maxm < PMNET_MIN_MATCHES)) {
                // the None child is not a leaf here (y < L - 1): same masks and totals, one level down
                if (!forced) ++st_nodes;
                if (DON) deep = (deep & ((2u << d) - 1u)) | (nmatch >= PMNET_MIN_MATCHES ? (2u << d) : 0u);
                if (lane == d) st_phase = 1;
                unsigned alive[W];
#pragma unroll
                for (int w = 0; w < W; ++w) alive[w] = __shfl_sync(kFull, st_alive[w], d);
                if (lane == d + 1) {
                  st_cursor = end;
                  st_maxm = 0;
                  st_nchild = 0;
                  st_phase = 0;
                  st_nmatch = nmatch;
                  st_entry = -1;
                  st_mslot = mslot;
                  st_tslot = tslot;
                  st_pbase = 0;
#pragma unroll
                  for (int w = 0; w < W; ++w) st_alive[w] = alive[w];
                }
                d = d + 1;
                continue;
              }
              do_return = true;
            } else {
              do_return = true;  // the None child returned
            }
            if (do_return) {
              if (d == 0) break;
              const int ret = __shfl_sync(kFull, st_maxm, d) + (__shfl_sync(kFull, st_entry, d) >= 0 ? 1 : 0);
              if (lane == d - 1) st_maxm = max(st_maxm, ret);
              d = d - 1;
            }
          }
          // mean over conformers (graph_match.py:109)
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < W; ++w) s += on[w] ? (double)best[w] : 0.0;
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
          score_out = (float)(s / (double)C);
        }
      }
      (void)lane_on;
    }
    if (TK && is_task) {
      // fold the task into the ligand's accumulator (pmnet_heavy_finish_kernel)
      if (status != PMNET_LIG_OK || task_bad) {
        if (lane == 0) atomicAdd((unsigned int*)args.workspace + kHdrTaskBad, 1u);
      } else {
        // scores are >= 0 (best starts at 0): their bit patterns order like integers
#pragma unroll
        for (int w = 0; w < W; ++w) atomicMax((int*)task_acc + 32 * w + lane, __float_as_int(best[w]));
        if (lane == 0) {
          atomicAdd(task_acc + kAccNodes, st_nodes);
          atomicAdd(task_acc + kAccLeaves, st_leaves);
          if (task_j < 0) {  // (a task with a path has computed only the rows it needs; the donor wrote these)
            task_acc[kAccRows] = st_rows;
            task_acc[kAccPairs] = st_pairs;
          }
        }
      }
      __syncwarp();
      continue;
    }
    if (heavy) {
      // this warp donated parts of the tree: its own part is folded into the accumulator like a task's, and
      // pmnet_heavy_finish_kernel writes the ligand's outputs after the last task launch
#pragma unroll
      for (int w = 0; w < W; ++w) atomicMax((int*)task_acc + 32 * w + lane, __float_as_int(best[w]));
      if (lane == 0) {
        atomicAdd(task_acc + kAccNodes, st_nodes);
        atomicAdd(task_acc + kAccLeaves, st_leaves);
        task_acc[kAccRows] = st_rows;
        task_acc[kAccPairs] = st_pairs;
      }
      status = PMNET_LIG_HEAVY;
      score_out = 0.0f;
    }
#ifdef PM_TIMING
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm2));
    st_rows = (uint32_t)((tm1 - tm0) / 1000ull);
    st_pairs = (uint32_t)((tm2 - tm1) / 1000ull);
#endif
    if (lane == 0) {
      args.out_scores[lig] = score_out;
      args.out_status[lig] = status;
      if (args.out_stats) {
        uint32_t* o = args.out_stats + (size_t)lig * 4;
        o[0] = st_nodes;  // tree nodes incl. the root
        o[1] = st_leaves;
        o[2] = st_rows;
        o[3] = st_pairs;
      }
    }
    if (args.out_conf) {
#pragma unroll
      for (int w = 0; w < W; ++w)
        if (32 * w < args.conf_stride)
          args.out_conf[(size_t)lig * args.conf_stride + 32 * w + lane] = (status == PMNET_LIG_OK) ? best[w] : 0.0f;
    }
    __syncwarp();
  }
}


// One warp per heavy ligand, after the last task launch: score, status and statistics from the accumulator.
struct FinishArgs {
  const unsigned char* workspace;
  const uint32_t* heavy_list;
  const uint32_t* heavy_acc;
  const int32_t* n_conf;
  float* out_scores;
  float* out_conf;
  int32_t* out_status;
  uint32_t* out_stats;
  int conf_stride;
};

__global__ void __launch_bounds__(128) pmnet_heavy_finish_kernel(const FinishArgs args) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned nh = ((const unsigned int*)args.workspace)[1];
  if (nh > (unsigned)kHeavyCap) nh = kHeavyCap;
  for (unsigned h = warp; h < nh; h += nwarps) {
    const uint32_t lig = args.heavy_list[h];
    const uint32_t* const acc = args.heavy_acc + (size_t)h * kAccWords;
    const int C = args.n_conf[lig];
    const int nw = args.conf_stride / 32;  // conformer words per lane (1, 2 or 4)
    float best[4];
    double s = 0.0;  // mean over conformers (graph_match.py:109): the same sum, in the same order, as the walkers'
    for (int w = 0; w < nw; ++w) {
      best[w] = __int_as_float((int)acc[32 * w + lane]);
      s += (lane + 32 * w < C) ? (double)best[w] : 0.0;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (lane == 0) {
      args.out_scores[lig] = (float)(s / (double)C);
      args.out_status[lig] = PMNET_LIG_OK;
      if (args.out_stats) {
        uint32_t* o = args.out_stats + (size_t)lig * 4;
        o[0] = acc[kAccNodes];
        o[1] = acc[kAccLeaves];
        o[2] = acc[kAccRows];
        o[3] = acc[kAccPairs];
      }
    }
    if (args.out_conf)
      for (int w = 0; w < nw; ++w) args.out_conf[(size_t)lig * args.conf_stride + 32 * w + lane] = best[w];
  }
}

int sm_count_cached() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// 32-conformer words per ligand for a launch that must handle up to `max_conformers` conformers
int conf_words(int max_conformers) { return max_conformers <= 32 ? 1 : (max_conformers <= 64 ? 2 : 4); }

void resolve_cfg(const PmScoreConfig* in, PmScoreConfig* out, bool query_device) {
  PmScoreConfig c = {};
  if (in) c = *in;
  if (c.max_conformers <= 0) c.max_conformers = 32;
  const int W = conf_words(c.max_conformers);
  const int max_warps = block_threads(W) / 32;
  if (c.warps_per_block <= 0) c.warps_per_block = max_warps;
  if (c.warps_per_block > max_warps) c.warps_per_block = max_warps;
  if (c.blocks <= 0) c.blocks = min_blocks(W) * (query_device ? sm_count_cached() : 148);
  if (c.scratch_rows <= 0) c.scratch_rows = 8192;
  *out = c;
}

// The specialised kernel runs first when the caller left the launch shape to the library, every ligand has at most 32
// conformers, nothing restricts the call to a status and the model tables fit in shared memory next to the per-warp
// tables. PMNET_NO_FAST=1 in the environment forces the general kernel (A/B measurements).
bool use_fast_kernel(const PmScoreConfig* in, int n_nodes, int n_clusters, int n_cluster_nodes) {
  static const bool disabled = [] {
    const char* e = getenv("PMNET_NO_FAST");
    return e && e[0] && e[0] != '0';
  }();
  if (disabled) return false;
  if (in && (in->warps_per_block > 0 || in->blocks > 0 || in->scratch_rows > 0 || in->rescore_status != 0)) return false;
  if (in && in->max_conformers > 32) return false;
  if (n_cluster_nodes < 0) n_cluster_nodes = n_nodes > n_clusters ? 4 * n_nodes : 4 * n_clusters;  // sizing query: a bound
  return fastk::smem_bytes(n_nodes, n_clusters, n_cluster_nodes) <= fastk::kSmemMax;
}

size_t fast_workspace_bytes() {
  return kHeaderBytes + (size_t)sm_count_cached() * fastk::kCtasPerSm * fastk::kWarps * fastk::G_BYTES;
}

// Node budget of the task-parallel walk for this configuration (0: off): not on a status-restricted re-run.
uint32_t heavy_budget_of(const PmScoreConfig* in) {
  if (!in) return kDefaultHeavyBudget;
  if (in->heavy_budget < 0 || in->rescore_status != 0) return 0u;
  return in->heavy_budget > 0 ? (uint32_t)in->heavy_budget : kDefaultHeavyBudget;
}

// bytes of the workspace in front of the heavy-ligand region (header, per-warp scratch, edge table)
size_t scratch_bytes(int32_t n_model_nodes, int32_t n_model_clusters, const PmScoreConfig* cfg) {
  PmScoreConfig c;
  resolve_cfg(cfg, &c, cfg == nullptr || cfg->blocks <= 0);
  const WarpLayout L = make_layout(n_model_clusters, c.scratch_rows, conf_words(c.max_conformers));
  // + room for the edge table of a large model (used when the tables do not fit in shared memory)
  const size_t edge = align_up((size_t)(n_model_nodes > 0 ? n_model_nodes : 0) * (size_t)(n_model_nodes > 0 ? n_model_nodes : 0) * 16, 256);
  size_t need = kHeaderBytes + (size_t)c.blocks * c.warps_per_block * L.bytes + edge;
  // the specialised kernel's scratch aliases the general kernel's (they run one after the other on the stream)
  if (use_fast_kernel(cfg, n_model_nodes, n_model_clusters, -1) && fast_workspace_bytes() > need) need = fast_workspace_bytes();
  return align_up(need, 256);
}

}  // namespace

// shared by the other translation units of the library
void pmnet_set_error(const char* msg) { set_err(msg); }

extern "C" {

int pmnet_abi_version(void) { return PMNET_ABI_VERSION; }

const char* pmnet_last_error_string(void) { return g_err; }

size_t pmnet_score_workspace_bytes(int32_t n_model_nodes, int32_t n_model_clusters, const PmScoreConfig* cfg) {
  // the list of heavy ligands and their task records sit behind everything else
  return scratch_bytes(n_model_nodes, n_model_clusters, cfg) + (heavy_budget_of(cfg) ? kHeavyBytes : 0) +
         (use_fast_kernel(cfg, n_model_nodes, n_model_clusters, -1) ? kDeferBytes : 0);
}

int pmnet_score_batch(const PmModel* model, const PmLigandBatch* batch, const float* weights, float* out_scores,
                      float* out_conf_scores, int32_t* out_status, uint32_t* out_stats, void* workspace,
                      size_t workspace_bytes, const PmScoreConfig* cfg, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!model || !batch || !weights) {
    set_err("pmnet_score_batch: null argument");
    return PMNET_EINVAL;
  }
  if (batch->n_ligands < 0 || model->n_nodes < 0 || model->n_clusters < 0) {
    set_err("pmnet_score_batch: negative size");
    return PMNET_EINVAL;
  }
  if (batch->n_ligands == 0) return PMNET_OK;
  if (!out_scores || !out_status || !workspace) {
    set_err("pmnet_score_batch: null output or workspace pointer");
    return PMNET_EINVAL;
  }
  if (model->n_nodes > 255 || model->n_clusters > 255) {
    set_err("pmnet_score_batch: model has more than 255 nodes or clusters");
    return PMNET_ELIMIT;
  }
  PmScoreConfig c;
  resolve_cfg(cfg, &c, true);
  // sized with the caller's configuration: an all-default one also reserves the specialised kernel's scratch
  const size_t need = pmnet_score_workspace_bytes(model->n_nodes, model->n_clusters, cfg);
  if (workspace_bytes < need) {
    set_err("pmnet_score_batch: workspace too small");
    return PMNET_EWORKSPACE;
  }
  KernelArgs a;
  a.model = *model;
  a.batch = *batch;
  for (int i = 0; i < PMNET_NUM_TYPES; ++i) a.w[i] = weights[i];
  a.out_scores = out_scores;
  a.out_conf = out_conf_scores;
  a.out_status = out_status;
  a.out_stats = out_stats;
  a.workspace = (unsigned char*)workspace;
  a.scratch_rows = c.scratch_rows;
  a.ly = make_layout(model->n_clusters, c.scratch_rows, conf_words(c.max_conformers));
  const int32_t n_cluster_nodes = model->n_cluster_nodes;
  if (n_cluster_nodes < 0 || n_cluster_nodes > 65535) {
    set_err("pmnet_score_batch: n_cluster_nodes out of range");
    return PMNET_ELIMIT;
  }
  cudaError_t e;
  a.n_cluster_nodes = n_cluster_nodes;
  if (c.max_conformers > PMNET_MAX_CONFORMERS) {
    set_err("pmnet_score_batch: max_conformers exceeds PMNET_MAX_CONFORMERS");
    return PMNET_ELIMIT;
  }
  const int W = conf_words(c.max_conformers);
  a.conf_stride = 32 * W;
  const size_t wsm = W == 1 ? sizeof(WarpSmem<1>) : (W == 2 ? sizeof(WarpSmem<2>) : sizeof(WarpSmem<4>));
  // model tables: pinned in shared memory when they fit next to the per-warp DFS stacks, else read from global memory
  const bool tg = smem_model_bytes(model->n_nodes, model->n_clusters, n_cluster_nodes, false) > kSmemTablesMax;
  const size_t smem =
      smem_model_bytes(model->n_nodes, model->n_clusters, n_cluster_nodes, tg) + (size_t)c.warps_per_block * wsm;
  if (smem > 200 * 1024) {
    set_err("pmnet_score_batch: per-node / per-cluster model arrays do not fit in shared memory");
    return PMNET_ELIMIT;
  }
  // ---- the specialised kernel first (csrc/scoring_fast.cuh); what it defers is picked up by the general kernel below
  const bool fast = W == 1 && !tg && use_fast_kernel(cfg, model->n_nodes, model->n_clusters, n_cluster_nodes);
  a.only_status = (cfg && cfg->rescore_status != 0) ? cfg->rescore_status : -1;
  a.counter_word = 0;
  a.heavy_budget = heavy_budget_of(cfg);
  a.heavy_list = (uint32_t*)((unsigned char*)workspace + scratch_bytes(model->n_nodes, model->n_clusters, cfg));
  a.heavy_acc = a.heavy_list + kHeavyCap;
  a.task_buf = a.heavy_acc + (size_t)kHeavyCap * kAccWords;
  a.task_ready = a.task_buf + (size_t)kTaskCap * kTaskWords;
  a.task_round = 0;
  a.list = nullptr;
  a.list_count_word = 0;
  a.list_cap = 0;
  e = cudaMemsetAsync(workspace, 0, kHeaderBytes, stream);
  if (e == cudaSuccess && a.heavy_budget != 0u) e = cudaMemsetAsync(a.task_ready, 0, (size_t)kTaskCap * 4, stream);
  if (e != cudaSuccess) {
    set_err(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  if (fast) {
    fastk::FastArgs fa;
    fa.model = *model;
    fa.batch = *batch;
    for (int i = 0; i < PMNET_NUM_TYPES; ++i) fa.w[i] = weights[i];
    fa.out_scores = out_scores;
    fa.out_conf = out_conf_scores;
    fa.out_status = out_status;
    fa.out_stats = out_stats;
    fa.workspace = (unsigned char*)workspace;
    fa.n_cluster_nodes = n_cluster_nodes;
    fa.heavy_budget = a.heavy_budget;
    fa.heavy_list = a.heavy_list;
    fa.heavy_acc = a.heavy_acc;
    fa.defer_list = nullptr;
    if (batch->n_ligands <= kDeferCap) {
      fa.defer_list = (uint32_t*)((unsigned char*)a.heavy_list + (a.heavy_budget ? kHeavyBytes : 0));
      a.list = fa.defer_list;
      a.list_count_word = 5;
      a.list_cap = kDeferCap;
    }
    const size_t fsmem = fastk::smem_bytes(model->n_nodes, model->n_clusters, n_cluster_nodes);
    e = cudaFuncSetAttribute(fastk::pmnet_score_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
    if (e != cudaSuccess) {
      set_err(cudaGetErrorString(e));
      return PMNET_ECUDA;
    }
    fastk::pmnet_score_fast_kernel<<<sm_count_cached() * fastk::kCtasPerSm, fastk::kWarps * 32, fsmem, stream>>>(fa);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_err(cudaGetErrorString(e));
      return PMNET_ECUDA;
    }
    a.only_status = PMNET_LIG_DEFERRED;
    a.counter_word = 2;
  }
  a.edge_g = nullptr;
  if (tg) {
    // the edge table sits behind the per-warp scratch (pmnet_score_workspace_bytes reserves it)
    const size_t off = kHeaderBytes + (size_t)c.blocks * c.warps_per_block * a.ly.bytes;
    EdgeTableArgs ea;
    ea.model = *model;
    for (int i = 0; i < PMNET_NUM_TYPES; ++i) ea.w[i] = weights[i];
    ea.out = (float4*)((unsigned char*)workspace + off);
    a.edge_g = ea.out;
    const int n2 = model->n_nodes * model->n_nodes;
    build_edge_table_kernel<<<(n2 + 255) / 256, 256, 0, stream>>>(ea);
  }
  const void* fn;
  if (W == 1) fn = tg ? (const void*)pmnet_score_kernel<1, true> : (const void*)pmnet_score_kernel<1, false>;
  else if (W == 2) fn = tg ? (const void*)pmnet_score_kernel<2, true> : (const void*)pmnet_score_kernel<2, false>;
  else fn = tg ? (const void*)pmnet_score_kernel<4, true> : (const void*)pmnet_score_kernel<4, false>;
  a.ligands_first = 0;
  if (a.heavy_budget == 0u) {
    // ---- the general kernel over the ligand queue (all ligands, the deferred ones, or a status-restricted re-run)
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_err(cudaGetErrorString(e));
      return PMNET_ECUDA;
    }
    void* kargs[] = {(void*)&a};
    e = cudaLaunchKernel(fn, dim3(c.blocks), dim3(c.warps_per_block * 32), kargs, smem, stream);
    if (e != cudaSuccess) {
      set_err(cudaGetErrorString(e));
      return PMNET_ECUDA;
    }
  } else {
    // ---- the task kernel (the general kernel's code behind a task queue), kTaskRounds launches: the first serves the
    // ligand queue like the general kernel would, then the ligands the specialised kernel gave up and the task queue
    // (its walkers donate into the queue it is consuming); the others return at once unless something is left over.
    // Then the kernel that turns the accumulators of the heavy ligands into scores
    const void* tfn;
    if (W == 1) tfn = tg ? (const void*)pmnet_score_kernel<1, true, true> : (const void*)pmnet_score_kernel<1, false, true>;
    else if (W == 2) tfn = tg ? (const void*)pmnet_score_kernel<2, true, true> : (const void*)pmnet_score_kernel<2, false, true>;
    else tfn = tg ? (const void*)pmnet_score_kernel<4, true, true> : (const void*)pmnet_score_kernel<4, false, true>;
    e = cudaFuncSetAttribute(tfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_err(cudaGetErrorString(e));
      return PMNET_ECUDA;
    }
    for (int round = 0; round < kTaskRounds; ++round) {
      KernelArgs t = a;
      t.heavy_budget = round == kTaskRounds - 1 ? 0u : a.heavy_budget;  // the last launch walks everything to the end
      t.task_round = round;
      t.ligands_first = round == 0 ? 1 : 0;
      void* targs[] = {(void*)&t};
      e = cudaLaunchKernel(tfn, dim3(c.blocks), dim3(c.warps_per_block * 32), targs, smem, stream);
      if (e != cudaSuccess) {
        set_err(cudaGetErrorString(e));
        return PMNET_ECUDA;
      }
    }
    FinishArgs fa;
    fa.workspace = (const unsigned char*)workspace;
    fa.heavy_list = a.heavy_list;
    fa.heavy_acc = a.heavy_acc;
    fa.n_conf = batch->n_conf;
    fa.out_scores = out_scores;
    fa.out_conf = out_conf_scores;
    fa.out_status = out_status;
    fa.out_stats = out_stats;
    fa.conf_stride = a.conf_stride;
    pmnet_heavy_finish_kernel<<<sm_count_cached() * 2, 128, 0, stream>>>(fa);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_err(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

// ---------------------------------------------------------------- processing order (longest ligands first)
// key[i] = number of (level, model cluster) entries of ligand i, exactly the T of the scoring kernel's phase 0
__global__ void ligand_cost_kernel(const PmModel gm, const PmLigandBatch B, uint32_t* keys, int32_t* idx) {
  __shared__ uint16_t cnt_by_mask[128];  // model clusters sharing a type with a 7-bit ligand cluster mask
  for (int m = threadIdx.x; m < 128; m += blockDim.x) {
    int c = 0;
    for (int k = 0; k < gm.n_clusters; ++k) c += (gm.cluster_mask[k] & m) ? 1 : 0;
    cnt_by_mask[m] = (uint16_t)c;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B.n_ligands) return;
  const uint8_t* tmask = B.node_type_mask + (B.lig_node_off[i] - B.node_base);
  const uint8_t* cl_nodes = B.cluster_nodes - B.cnode_base;
  const int q0 = B.lig_cluster_off[i] - B.cluster_base, q1 = B.lig_cluster_off[i + 1] - B.cluster_base;
  uint32_t t = 0;
  int levels = 0;
  for (int q = q0; q < q1 && levels < kMaxDepth; ++q) {
    unsigned m = 0;
    for (int j = B.cluster_node_off[q]; j < B.cluster_node_off[q + 1]; ++j) m |= tmask[cl_nodes[j]];
    const uint32_t c = cnt_by_mask[m & 127u];
    t += c;
    levels += c ? 1 : 0;
  }
  keys[i] = t > 65535u ? 65535u : t;
  idx[i] = i;
}

static size_t order_sort_bytes(int32_t n) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                            (const int32_t*)nullptr, (int32_t*)nullptr, n, 0, 16);
  return tmp;
}

size_t pmnet_order_workspace_bytes(int32_t n) {
  if (n <= 0) return 256;
  return 3 * align_up((size_t)n * 4, 256) + align_up(order_sort_bytes(n), 256) + 256;
}

int pmnet_cost_order(const PmModel* model, const PmLigandBatch* batch, int32_t* out_order, void* workspace,
                     size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!model || !batch || batch->n_ligands < 0 || (batch->n_ligands > 0 && (!out_order || !workspace))) {
    set_err("pmnet_cost_order: bad argument");
    return PMNET_EINVAL;
  }
  const int32_t n = batch->n_ligands;
  if (n == 0) return PMNET_OK;
  if (workspace_bytes < pmnet_order_workspace_bytes(n)) {
    set_err("pmnet_cost_order: workspace too small");
    return PMNET_EWORKSPACE;
  }
  unsigned char* p = (unsigned char*)workspace;
  uint32_t* keys = (uint32_t*)p;      p += align_up((size_t)n * 4, 256);
  uint32_t* keys_out = (uint32_t*)p;  p += align_up((size_t)n * 4, 256);
  int32_t* idx = (int32_t*)p;         p += align_up((size_t)n * 4, 256);
  ligand_cost_kernel<<<(n + 255) / 256, 256, 0, stream>>>(*model, *batch, keys, idx);
  size_t tmp = order_sort_bytes(n);
  // stable LSD radix sort on the 16 key bits: equal cost keeps index order
  cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(p, tmp, keys, keys_out, idx, out_order, n, 0, 16, stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_err(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

// ---------------------------------------------------------------- top-k of a shard
__global__ void iota_ids_kernel(int64_t* ids, int64_t n, int64_t base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = base + i;
}

__global__ void topk_pad_kernel(const float* ks, const int64_t* vs, int64_t n, int k, float* out_s, int64_t* out_i) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) {
    out_s[i] = (i < n) ? ks[i] : -__int_as_float(0x7f800000);
    out_i[i] = (i < n) ? vs[i] : -1;
  }
}

static size_t topk_sort_bytes(int64_t n) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, (const float*)nullptr, (float*)nullptr,
                                            (const int64_t*)nullptr, (int64_t*)nullptr, n);
  return tmp;
}

size_t pmnet_topk_workspace_bytes(int64_t n, int32_t k) {
  (void)k;
  if (n <= 0) return 256;
  return align_up((size_t)n * 8, 256) * 2 + align_up((size_t)n * 4, 256) + align_up(topk_sort_bytes(n), 256) + 256;
}

int pmnet_topk(const float* scores, int64_t n, int64_t id_base, int32_t k, float* out_scores, int64_t* out_ids,
               void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n < 0 || k <= 0 || !out_scores || !out_ids || (n > 0 && (!scores || !workspace))) {
    set_err("pmnet_topk: bad argument");
    return PMNET_EINVAL;
  }
  if (n > 0x7fffffffLL) {
    set_err("pmnet_topk: more than 2^31-1 items");
    return PMNET_ELIMIT;
  }
  if (workspace_bytes < pmnet_topk_workspace_bytes(n, k)) {
    set_err("pmnet_topk: workspace too small");
    return PMNET_EWORKSPACE;
  }
  float* ks = nullptr;
  int64_t* vs = nullptr;
  if (n > 0) {
    unsigned char* p = (unsigned char*)workspace;
    int64_t* ids = (int64_t*)p;  p += align_up((size_t)n * 8, 256);
    vs = (int64_t*)p;            p += align_up((size_t)n * 8, 256);
    ks = (float*)p;              p += align_up((size_t)n * 4, 256);
    size_t tmp = topk_sort_bytes(n);
    iota_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ids, n, id_base);
    // stable LSD radix sort: equal scores keep ascending ligand id
    cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(p, tmp, scores, ks, ids, vs, n, 0, 32, stream);
    if (e != cudaSuccess) {
      set_err(cudaGetErrorString(e));
      return PMNET_ECUDA;
    }
  }
  topk_pad_kernel<<<(k + 255) / 256, 256, 0, stream>>>(ks, vs, n, k, out_scores, out_ids);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_err(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

}  // extern "C"
