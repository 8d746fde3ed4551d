// scoring_fast.cuh - the specialised scoring kernel for the common case (included by scoring.cu, inside its anonymous
// namespace). Same algorithm, same fp32 operations in the same order and therefore bit-identical results as
// pmnet_score_kernel (the generic kernel, which stays the in-launch fallback), restricted to
//   up to 32 conformers (lane = conformer), model tables pinned in shared memory,
//   per ligand: <= 88 (level, model cluster) entries, <= 32 entries per level, <= 12 levels, <= 224 node-match records
//   with <= 3 matched model nodes each, <= 48 ligand nodes in the levels, <= 384 mask-stack words, <= 4096 pair
//   entries (their <= 4184 pair-score rows always fit).
// A ligand outside these caps gets status PMNET_LIG_DEFERRED and is scored by the generic kernel, which
// pmnet_score_batch enqueues right behind this one on the same stream (status-driven queue, no host round trip).
//
// What the restriction buys (round-1 ncu: 4.9e5 warp instructions per ligand, 11 % of them re-deriving scratch
// pointers from the constant bank, ~105 per pair term, ~125 per tree node):
//  * every per-ligand table (entries, node-match records, pair bases, level tables, DFS stacks) lives in shared memory at
//    compile-time offsets from ONE per-warp base register; the global scratch (pair-score rows, node-pair distances,
//    validity words, row indices, cluster geometry) hangs off ONE 64-bit per-lane base with compile-time region offsets
//    that fit the load / store immediates;
//  * phase 1: an entry is one 32-bit word {model cluster, record count, record offset, all-single-match flag}, a record
//    is {local node, matched model nodes}; pairs of entries whose records all match a single model node run a dedicated
//    term loop (load record, load distance, load edge, 8 fp instructions); validity words / row indices of a whole
//    level are kept lane-indexed and stored coalesced;
//  * phase 2: the current node's control state lives in ordinary (uniform) registers and is saved / restored as one
//    16-byte shared-memory word on push / pop; the candidates of a node are ONE ballot word that is peeled bit by bit.
//
// Reference: src/pmnet/scoring/graph_match.py:85-101, 222-279; match_utils_numba.py:12-231; tree.py:55-104.

namespace fastk {

#ifndef PM_FAST_TC
#define PM_FAST_TC 88
#endif
#ifndef PM_FAST_WARPS
#define PM_FAST_WARPS 32
#endif
#ifndef PM_FAST_PREFETCH
#define PM_FAST_PREFETCH 0  // measured: 71.4 vs 70.0 M conformers/s without / with the sibling prefetch
#endif
#ifndef PM_FAST_HEAVY_CHECK
#define PM_FAST_HEAVY_CHECK 3  // where the node budget is tested: 0 child created, 1 child pushed, 2 never, 3 node popped (measured:
                               // 69.1 / 70.2 / 71.6 / 72.3 M conformers/s), 4 node popped at depth <= 3
#endif
#ifndef PM_FAST_ROW_LD
#define PM_FAST_ROW_LD 0    // how the DFS reads score rows: 0 plain (L1 + L2), 1 ld.global.cg (L2 only), 2 ld.global.cs (streaming)
#endif
#if PM_FAST_ROW_LD == 1
#define PM_ROW_LD(p) __ldcg(p)
#elif PM_FAST_ROW_LD == 2
#define PM_ROW_LD(p) __ldcs(p)
#else
#define PM_ROW_LD(p) (*(p))
#endif
#ifndef PM_FAST_ANC_WIDTH
#define PM_FAST_ANC_WIDTH 2  // independent row loads per round of the ancestor sums
#endif
#ifndef PM_FAST_CTAS
#define PM_FAST_CTAS 1      // CTAs per SM (PM_FAST_WARPS warps each)
#endif
constexpr int TC = PM_FAST_TC;  // (level, model cluster) entries per ligand (88: keeps the CTA under the 164 KB carve-out)
constexpr int RC = 224;     // node-match records per ligand
constexpr int LC = 12;      // levels
constexpr int NLC = 48;     // ligand nodes in the selected levels
constexpr int MKW = 384;    // words of the triangular mask stack
#ifndef PM_FAST_ROWS
#define PM_FAST_ROWS 4224   // >= PC + TC: a ligand within the other caps can never run out of rows. (With 2048 a few
#endif                      // ligands per 262 144 were deferred and the general kernel's pass for them - one warp
                            // each, the GPU idle - cost 2.6 ms per launch)
constexpr int ROWS = PM_FAST_ROWS;  // pair-score rows per warp
constexpr int PC = 4096;    // pair entries per warp
constexpr int kWarps = PM_FAST_WARPS;  // one 1024-thread CTA per SM at 64 registers per thread
constexpr int kNoBase = INT32_MIN;  // lane a of my_pbase: the node at depth a is a None node (pair bases may be negative)

// per-warp global scratch, byte offsets from the warp's base (every per-conformer row is 128 B, lane = conformer)
constexpr uint32_t G_ROWS = 0;
constexpr uint32_t G_DIST = G_ROWS + ROWS * 128;
constexpr uint32_t G_GEO = G_DIST + NLC * NLC * 128;
constexpr uint32_t G_V = G_GEO + LC * 4 * 128;
constexpr uint32_t G_PROW = G_V + PC * 4;
constexpr uint32_t G_BYTES = G_PROW + PC * 4;
static_assert(G_BYTES % 256 == 0, "per-warp scratch must keep 256 B alignment");
static_assert(G_BYTES < (1u << 23), "region offsets must fit the 24-bit signed load/store immediates");

struct __align__(16) WarpS {
  uint4 stack[LC + 1];         // saved node state per depth: {candidate bits, alive word, packed counters, -}
  uint32_t ent[TC];            // model cluster | records << 8 | first record << 16 | all-single-match << 24
  int32_t rowbase[TC];         // pair index of (entry, e2) is rowbase[entry] + e2 for every entry e2 of a later level
  uint16_t srow[TC];           // row of the entry's self score, 0xffff = none
  int32_t lev_start[LC + 2];   // first entry of each level; [L] = T
  int32_t moff[LC + 2];        // mask stack: word offset of depth s minus lev_start[s]
  union {
    struct {                   // phases 0-1
      uint2 rec[RC];           // .x = local ligand node, .y = (M - 1) | model node a0 << 8 | a1 << 16 | a2 << 24
      int32_t lev_q[LC];
      int32_t lev_nbase[LC];
      uint8_t entlev[TC];
      uint8_t lnode[NLC];
    } a;
    struct {                   // phase 2
      uint32_t mk[MKW];
      float tot[LC][32];
    } b;
  } u;
};
static_assert(sizeof(WarpS) % 16 == 0, "WarpS must keep 16 B alignment");

struct FastArgs {
  PmModel model;
  PmLigandBatch batch;
  float w[PMNET_NUM_TYPES];
  float* out_scores;
  float* out_conf;   // [n][32] or null
  int32_t* out_status;
  uint32_t* out_stats;
  unsigned char* workspace;
  int n_cluster_nodes;
  uint32_t heavy_budget;  // tree nodes after which a ligand is abandoned as PMNET_LIG_HEAVY (0: never)
  uint32_t* heavy_list;   // ligands abandoned that way (count in header word 1); see scoring.cu
  uint32_t* heavy_acc;
  uint32_t* defer_list;   // ligands left PMNET_LIG_DEFERRED, in the order met (count in header word 5); null: no list
};

__host__ __device__ inline size_t smem_bytes(int nm, int km, int n_cluster_nodes) {
  return smem_model_bytes(nm, km, n_cluster_nodes, false) + (size_t)kWarps * sizeof(WarpS);
}
constexpr int kCtasPerSm = PM_FAST_CTAS;
constexpr size_t kSmemMax = (227 * 1024) / kCtasPerSm - 1024 * (kCtasPerSm - 1);

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// one ligand-node pair with single matched model nodes on both sides (match_utils_numba.py:67-86, M = N = 1)
// (the fail counter is kept in fp32: small integers are exact and it saves the predicate round trip)
__device__ __forceinline__ void term11(const float4 e, const float d, float& sc, float& nf) {
  const float s = __fmul_rn(__fsub_rn(d, e.x), e.y);
  const float s2 = __fmul_rn(s, s);
  nf += (s2 < 4.0f) ? 0.0f : 1.0f;
  sc = fmaf(e.w, gauss(s2), sc);
}

// 1 / (M N) exactly as the general kernel forms it: 0.5 and 0.25 for the common 2 and 4, rcp.approx otherwise
__device__ __forceinline__ float inv_mn(const int mn) {
  return mn == 2 ? 0.5f : (mn == 4 ? 0.25f : rcp_approx((float)mn));
}

// one Gaussian evaluation of a multi-match term (match_utils_numba.py:67-84)
#define PM_EVAL(rowp, col)                                                    \
  {                                                                           \
    const float4 e_ = *(const float4*)((rowp) + (col));                       \
    const float s_ = __fmul_rn(__fsub_rn(d, e_.x), e_.y);                     \
    const float s2_ = __fmul_rn(s_, s_);                                      \
    lik = fmaf(e_.w, gauss(s2_), lik);                                        \
    npass += (s2_ < 4.0f) ? 1 : 0;                                            \
  }

// A ligand-node pair whose records match up to 3 x 3 model nodes, row-major like the reference. `rows` = shared-memory
// byte addresses of the edge-table rows of the first record's model nodes (hoisted by the caller: they do not depend on
// the second record), M1 = M - 1, y2 = the second record's {N - 1, b0, b1, b2}; inv_tab[mn] = inv_mn(mn).
__device__ __forceinline__ void term_rows(const unsigned char* const (&rows)[3], const unsigned M1, const uint32_t y2,
                                          const float d, const float* __restrict__ inv_tab, float& sc, float& nf) {
  const unsigned N1 = y2 & 255u;
  if ((M1 | N1) == 0u) {
    term11(*(const float4*)(rows[0] + (y2 >> 4)), d, sc, nf);  // y2 = b0 << 8: (y2 >> 4) = 16 b0
    return;
  }
  const unsigned c0 = (y2 >> 4) & 0xff0u, c1 = (y2 >> 12) & 0xff0u, c2 = (y2 >> 20) & 0xff0u;
  float lik = 0.0f;
  int npass = 0;
  PM_EVAL(rows[0], c0)
  if (N1 >= 1u) PM_EVAL(rows[0], c1)
  if (N1 >= 2u) PM_EVAL(rows[0], c2)
  if (M1 >= 1u) {
    PM_EVAL(rows[1], c0)
    if (N1 >= 1u) PM_EVAL(rows[1], c1)
    if (N1 >= 2u) PM_EVAL(rows[1], c2)
    if (M1 >= 2u) {
      PM_EVAL(rows[2], c0)
      if (N1 >= 1u) PM_EVAL(rows[2], c1)
      if (N1 >= 2u) PM_EVAL(rows[2], c2)
    }
  }
  const int mn = (int)((M1 + 1u) * (N1 + 1u));
  nf += (npass < ((mn + 1) >> 1)) ? 1.0f : 0.0f;
  sc = fmaf(lik, inv_tab[mn], sc);
}

// All leaf children of a node in one pass (graph_match.py:103-109): `bal` = candidate bits (bit b = entry base + b),
// `lmw` = lane-indexed conformer masks of those entries, `tt` = totals of the node, lanes 1..dmax with is_anc hold the
// pair base `pb` of a matched ancestor.
__device__ __forceinline__ int leaf_pass(const WarpS& ws, const float* __restrict__ rows_l, const int32_t* __restrict__ prow,
                                         const int base, const unsigned lmw, unsigned bal, const float tt,
                                         const bool is_anc, const int pb, const int dmax, const int lane, float& best) {
  int nleaf = 0;
  // the row indices of the next leaf are requested while the current one is summed (two dependent loads per leaf)
  int nf = base + __ffs(bal) - 1;
  int myrow_n = (is_anc && bal) ? prow[pb + nf] : -1;
  unsigned sr_n = bal ? ws.srow[nf] : 0xffffu;
  while (bal) {
    const int src = __ffs(bal) - 1;
    bal &= bal - 1;
    const int myrow = myrow_n;
    const unsigned sr = sr_n;
    const float self = (sr != 0xffffu) ? PM_ROW_LD(rows_l + sr * 32u) : 0.0f;
    if (bal) {
      nf = base + __ffs(bal) - 1;
      myrow_n = is_anc ? prow[pb + nf] : -1;
      sr_n = ws.srow[nf];
    }
    float t = 0.0f;
#if PM_FAST_ANC_WIDTH == 4
    for (int a0 = 1; a0 <= dmax; a0 += 4) {  // four independent row loads per round (lanes > dmax hold -1)
      const int r0 = __shfl_sync(kFull, myrow, a0), r1 = __shfl_sync(kFull, myrow, a0 + 1);
      const int r2 = __shfl_sync(kFull, myrow, a0 + 2), r3 = __shfl_sync(kFull, myrow, a0 + 3);
      const float v0 = (r0 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r0 * 32u) : 0.0f;
      const float v1 = (r1 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r1 * 32u) : 0.0f;
      const float v2 = (r2 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r2 * 32u) : 0.0f;
      const float v3 = (r3 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r3 * 32u) : 0.0f;
      t = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t, v0), v1), v2), v3);
    }
#else
    for (int a0 = 1; a0 <= dmax; a0 += 2) {  // two independent row loads per round (lanes > dmax hold -1)
      const int r0 = __shfl_sync(kFull, myrow, a0), r1 = __shfl_sync(kFull, myrow, a0 + 1);
      const float v0 = (r0 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r0 * 32u) : 0.0f;
      const float v1 = (r1 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r1 * 32u) : 0.0f;
      t = __fadd_rn(__fadd_rn(t, v0), v1);
    }
#endif
    const unsigned al = __shfl_sync(kFull, lmw, src);
    if ((al >> lane) & 1u) best = fmaxf(best, __fadd_rn(__fadd_rn(tt, self), t));
    ++nleaf;
  }
  return nleaf;
}

__global__ void __launch_bounds__(kWarps * 32, PM_FAST_CTAS) pmnet_score_fast_kernel(const FastArgs args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const PmModel& gm = args.model;
  const int NM = gm.n_nodes, KM = gm.n_clusters;

  // ---- shared memory: the model (same image as the generic kernel) + one WarpS per warp
  SmemModel sm;
  sm.nm = NM;
  sm.km = KM;
  {
    unsigned char* p = smem_raw;
    sm.edge = (float4*)p;            p += (size_t)NM * NM * 16;
    sm.cdist = (float*)p;            p += (size_t)KM * KM * 4;
    sm.csize = (float*)p;            p += (size_t)KM * KM * 4;
    sm.wnode = (float*)p;            p += (size_t)NM * 4;
    sm.cnode_off = (uint16_t*)p;     p += align_up((size_t)(KM + 1) * 2, 4);
    sm.cnodes = (uint8_t*)p;         p += align_up((size_t)args.n_cluster_nodes, 4);
    sm.ntype = (uint8_t*)p;          p += align_up((size_t)NM, 4);
    sm.cmask = (uint8_t*)p;          p += align_up((size_t)KM, 4);
  }
  WarpS& ws = ((WarpS*)(smem_raw + smem_model_bytes(NM, KM, args.n_cluster_nodes, false)))[warp_in_block];
  __shared__ float inv_tab[16];  // inv_tab[M N], M, N <= 3
  if (threadIdx.x < 16) inv_tab[threadIdx.x] = threadIdx.x ? inv_mn((int)threadIdx.x) : 0.0f;
  const unsigned char* const edge_b = (const unsigned char*)sm.edge;
  const unsigned nm16 = (unsigned)NM * 16u;
  // cluster prefilter tables side by side: {distance, size sum} of model cluster pair (k, l)
  float2* const cds = (float2*)sm.cdist;
  for (int i = threadIdx.x; i < NM * NM; i += blockDim.x) sm.edge[i] = edge_entry(gm, args.w, i, NM);
  for (int i = threadIdx.x; i < KM * KM; i += blockDim.x)
    cds[i] = make_float2(gm.cluster_dist[i], gm.cluster_size_sum[i]);  // (the two fp32 tables' space, interleaved)
  for (int i = threadIdx.x; i < NM; i += blockDim.x) {
    sm.ntype[i] = gm.node_type[i];
    sm.wnode[i] = args.w[gm.node_type[i]];
  }
  for (int i = threadIdx.x; i < KM; i += blockDim.x) sm.cmask[i] = gm.cluster_mask[i];
  for (int i = threadIdx.x; i <= KM; i += blockDim.x) sm.cnode_off[i] = (uint16_t)gm.cluster_node_off[i];
  for (int i = threadIdx.x; i < args.n_cluster_nodes; i += blockDim.x) sm.cnodes[i] = gm.cluster_nodes[i];
  __syncthreads();

  // ---- per-warp global scratch: ONE base, compile-time region offsets
  unsigned char* const wbase =
      args.workspace + kHeaderBytes + (size_t)(blockIdx.x * kWarps + warp_in_block) * G_BYTES;
  float* const rows_l = (float*)(wbase + G_ROWS) + lane;
  float* const dist_l = (float*)(wbase + G_DIST) + lane;
  const float* const dist0 = (const float*)(wbase + G_DIST);
  float* const geo_l = (float*)(wbase + G_GEO) + lane;
  uint32_t* const Vt = (uint32_t*)(wbase + G_V);
  int32_t* const prow = (int32_t*)(wbase + G_PROW);
  unsigned int* const counter = (unsigned int*)args.workspace;
  const PmLigandBatch& B = args.batch;
  const uint32_t budget = args.heavy_budget != 0u ? args.heavy_budget : 0xffffffffu;

  for (;;) {
    unsigned int lig = 0;
    if (lane == 0) {
      lig = atomicAdd(counter, 1u);
      if (B.order != nullptr && lig < (unsigned)B.n_ligands) lig = (unsigned)B.order[lig];
    }
    lig = __shfl_sync(kFull, lig, 0);
    if (lig >= (unsigned)B.n_ligands) break;

    const int C = B.n_conf[lig];
    float score_out = 0.0f;
    int status = PMNET_LIG_OK;
    uint32_t st_nodes = 0, st_leaves = 0, st_rows = 0, st_pairs = 0;
    float best = 0.0f;
    bool defer = (C < 1 || C > 32);
    bool heavy = false;
    uint32_t node_limit = budget;  // nodes after which the ligand is handed to the task rounds (~0: never)

    if (!defer) {
      const int stride = (C + 3) & ~3;
      const float* xyz = B.coords + (B.coord_off[lig] - B.coord_base);
      const uint8_t* tmask = B.node_type_mask + (B.lig_node_off[lig] - B.node_base);
      const int q0 = B.lig_cluster_off[lig] - B.cluster_base, q1 = B.lig_cluster_off[lig + 1] - B.cluster_base;
      const uint8_t* cl_nodes = B.cluster_nodes - B.cnode_base;
      const bool on = lane < C;
      const unsigned cfull = C >= 32 ? kFull : ((1u << C) - 1u);

      // ================= phase 0: levels, entries, node-match records (graph_match.py:85-92, 124-172)
      int L = 0, T = 0, NL = 0;
      for (int q = q0; q < q1 && L < kMaxDepth && !defer; ++q) {
        const int c0 = B.cluster_node_off[q], c1 = B.cluster_node_off[q + 1];
        unsigned m = 0;
        for (int i = c0 + lane; i < c1; i += 32) m |= tmask[cl_nodes[i]];
        m = __reduce_or_sync(kFull, m);
        const int t_level = T;
        for (int k0 = 0; k0 < KM; k0 += 32) {
          const int k = k0 + lane;
          const bool hit = (k < KM) && (sm.cmask[k] & m);
          const unsigned bal = __ballot_sync(kFull, hit);
          if (T + __popc(bal) > TC) {
            defer = true;
            break;
          }
          if (hit) {
            const int e = T + __popc(bal & ((1u << lane) - 1u));
            ws.ent[e] = (uint32_t)k;
            ws.u.a.entlev[e] = (uint8_t)L;
          }
          T += __popc(bal);
        }
        if (defer) break;
        if (T > t_level) {
          const int n = c1 - c0;
          if (NL + n > NLC || T - t_level > 32 || L >= LC) {
            defer = true;
            break;
          }
          if (lane == 0) {
            ws.lev_start[L] = t_level;
            ws.u.a.lev_q[L] = q;
            ws.u.a.lev_nbase[L] = NL;
          }
          for (int i = lane; i < n; i += 32) ws.u.a.lnode[NL + i] = cl_nodes[c0 + i];
          // cluster centre and size per conformer (ligand.py:458-473), fp32 sequential like numpy
          {
            const float fn = (float)n;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            for (int i = c0; i < c1; ++i) {
              const int node = cl_nodes[i];
              const float x = ld_coord(xyz, stride, node, 0, lane, on), y = ld_coord(xyz, stride, node, 1, lane, on),
                          z = ld_coord(xyz, stride, node, 2, lane, on);
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
              const float dd = norm3(__fsub_rn(ld_coord(xyz, stride, node, 0, lane, on), cx),
                                     __fsub_rn(ld_coord(xyz, stride, node, 1, lane, on), cy),
                                     __fsub_rn(ld_coord(xyz, stride, node, 2, lane, on), cz));
              sz = (i == c0) ? dd : fmaxf(sz, dd);
            }
            float* g = geo_l + L * 128;
            g[0] = cx; g[32] = cy; g[64] = cz; g[96] = sz;
          }
          NL += n;
          ++L;
        }
      }
      __syncwarp();
      int mask_need = 0;
      if (!defer && L > 0) {
        if (lane == 0) ws.lev_start[L] = T;
        __syncwarp();
        // node-match records, one lane per entry: count, exclusive scan for the offsets, then fill
        uint32_t rec_used = 0;
        for (int e0 = 0; e0 < T && !defer; e0 += 32) {
          const int e = e0 + lane;
          uint32_t nrec = 0;
          bool toobig = false, multi = false;
          int q = 0, k = 0, nb = 0;
          if (e < T) {
            const int lev = ws.u.a.entlev[e];
            q = ws.u.a.lev_q[lev];
            nb = ws.u.a.lev_nbase[lev];
            k = (int)(ws.ent[e] & 255u);
            for (int i = B.cluster_node_off[q]; i < B.cluster_node_off[q + 1]; ++i) {
              const unsigned tm = tmask[cl_nodes[i]];
              int M = 0;
              for (int j = sm.cnode_off[k]; j < sm.cnode_off[k + 1]; ++j) M += (tm >> sm.ntype[sm.cnodes[j]]) & 1u;
              if (M > 3) toobig = true;
              if (M > 1) multi = true;
              if (M > 0) ++nrec;
            }
          }
          uint32_t irec = nrec;
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v1 = __shfl_up_sync(kFull, irec, o);
            if (lane >= o) irec += v1;
          }
          const uint32_t trec = __shfl_sync(kFull, irec, 31);
          if (__any_sync(kFull, toobig) || rec_used + trec > (uint32_t)RC) {
            defer = true;
            break;
          }
          if (e < T) {
            uint32_t ro = rec_used + irec - nrec;
            ws.ent[e] = (uint32_t)k | (nrec << 8) | (ro << 16) | (multi ? 0u : (1u << 24));
            const int c0 = B.cluster_node_off[q], c1 = B.cluster_node_off[q + 1];
            for (int i = c0; i < c1; ++i) {
              const unsigned tm = tmask[cl_nodes[i]];
              int M = 0;
              uint32_t packed = 0;
              for (int j = sm.cnode_off[k]; j < sm.cnode_off[k + 1]; ++j) {
                const int mn = sm.cnodes[j];
                if ((tm >> sm.ntype[mn]) & 1u) {
                  packed |= (uint32_t)mn << (8 * M + 8);
                  ++M;
                }
              }
              if (M == 0) continue;
              ws.u.a.rec[ro++] = make_uint2((uint32_t)(nb + (i - c0)), packed | (uint32_t)(M - 1));
            }
          }
          rec_used += trec;
        }
        // pair-index base of each entry: the pairs of e1 cover all entries of later levels
        if (!defer) {
          int run = 0;
          for (int l = 0; l < L; ++l) {
            const int s = ws.lev_start[l], e_end = ws.lev_start[l + 1];
            const int width = T - e_end;
            for (int e = s + lane; e < e_end; e += 32) ws.rowbase[e] = run + (e - s) * width - e_end;
            if (lane == 0) ws.moff[l] = mask_need - s;
            mask_need += T - s;
            run += (e_end - s) * width;
          }
          st_pairs = (uint32_t)run;
          if (run > PC || mask_need > MKW) defer = true;
        }
        // ligand node-pair distances (LigandEdge.set_distances, ligand.py:349-351), upper triangle of NL x NL
        if (!defer) {
          for (int i = 0; i < NL - 1; ++i) {
            const int ni = ws.u.a.lnode[i];
            const float xi = ld_coord(xyz, stride, ni, 0, lane, on), yi = ld_coord(xyz, stride, ni, 1, lane, on),
                        zi = ld_coord(xyz, stride, ni, 2, lane, on);
            float* drow = dist_l + (unsigned)(i * NL) * 32u;
            for (int j = i + 1; j < NL; ++j) {
              const int nj = ws.u.a.lnode[j];
              drow[(unsigned)j * 32u] = norm3(__fsub_rn(xi, ld_coord(xyz, stride, nj, 0, lane, on)),
                                              __fsub_rn(yi, ld_coord(xyz, stride, nj, 1, lane, on)),
                                              __fsub_rn(zi, ld_coord(xyz, stride, nj, 2, lane, on)));
            }
          }
        }
        __syncwarp();
      }

      if (defer) {
        // handled below
      } else if (L == 0) {
        status = PMNET_LIG_EMPTY;
      } else {
        // ================= phase 1: self scores and pair table (graph_match.py:222-279)
        int nrows = 0;
        for (int e = 0; e < T; ++e) {
          const uint32_t w1 = ws.ent[e];
          const int cnt = (int)((w1 >> 8) & 255u);
          unsigned r = 0xffffu;
          if (cnt >= 2) {
            float sc = 0.0f, nf = 0.0f;
            const uint2* rp = ws.u.a.rec + ((w1 >> 16) & 255u);
            for (int i = 0; i < cnt - 1; ++i) {
              const uint2 r1 = rp[i];
              const unsigned dro = r1.x * (unsigned)NL * 32u + (unsigned)lane;
              const unsigned char* const erows[3] = {edge_b + ((r1.y >> 8) & 255u) * nm16,
                                                     edge_b + ((r1.y >> 16) & 255u) * nm16, edge_b + (r1.y >> 24) * nm16};
              for (int j = i + 1; j < cnt; ++j) {
                const uint2 r2 = rp[j];
                term_rows(erows, r1.y & 255u, r2.y, dist0[dro + r2.x * 32u], inv_tab, sc, nf);
              }
            }
            if (nrows >= ROWS) {
              defer = true;
              break;
            }
            r = (unsigned)nrows++;
            rows_l[r * 32u] = sc;
          }
          if (lane == 0) ws.srow[e] = (uint16_t)r;
        }
        for (int i = 0; i < L - 1 && !defer; ++i) {
          const float* gi = geo_l + i * 128;
          const float cix = gi[0], ciy = gi[32], ciz = gi[64], csi = gi[96];
          const int s1 = ws.lev_start[i], e1_end = ws.lev_start[i + 1];
          for (int j = i + 1; j < L && !defer; ++j) {
            const float* gj = geo_l + j * 128;
            const float ldist = norm3(__fsub_rn(cix, gj[0]), __fsub_rn(ciy, gj[32]), __fsub_rn(ciz, gj[64]));
            const float lsize = __fadd_rn(csi, gj[96]);
            const int s2 = ws.lev_start[j], e2_end = ws.lev_start[j + 1];
            for (int e1 = s1; e1 < e1_end && !defer; ++e1) {
              const uint32_t w1 = ws.ent[e1];
              const int cnt1 = (int)((w1 >> 8) & 255u);
              const uint2* rp1 = ws.u.a.rec + ((w1 >> 16) & 255u);
              const float2* cd = cds + (w1 & 255u) * KM;
              unsigned myV = 0;  // lane b: validity word / row index of the pair (e1, s2 + b)
              int myR = -1;
              for (int e2 = s2; e2 < e2_end; ++e2) {
                const uint32_t w2 = ws.ent[e2];
                const unsigned l = w2 & 255u;
                // cluster prefilter (graph_match.py:263-268): min_c(|d_lig - d_mod| - size_lig) > size_mod
                const float2 cq = cd[l];
                const bool far = (__fsub_rn(fabsf(__fsub_rn(ldist, cq.x)), lsize) > cq.y) | !on;
                if (__all_sync(kFull, far)) continue;  // myV / myR of this lane stay 0 / -1
                const int cnt2 = (int)((w2 >> 8) & 255u);
                const uint2* rp2 = ws.u.a.rec + ((w2 >> 16) & 255u);
                const float thr = 0.5f * (float)(cnt1 * cnt2);  // match_utils_numba.py:191-196 (exact in fp32)
                float sc = 0.0f, nf = 0.0f;
                bool dead = false;
                if ((w1 & w2) >> 24) {
                  // every ligand node of both entries matches a single model node: straight term loop
                  for (int a = 0; a < cnt1; ++a) {
                    const uint2 r1 = rp1[a];
                    const unsigned dro = r1.x * (unsigned)NL * 32u + (unsigned)lane;
                    const unsigned char* erow = edge_b + (r1.y >> 8) * nm16;
#pragma unroll 2
                    for (int b = 0; b < cnt2; ++b) {
                      const uint2 r2 = rp2[b];
                      // r2.y = a0 << 8 here: (r2.y >> 4) is the byte offset of edge (a, a0) in the row
                      term11(*(const float4*)(erow + (r2.y >> 4)), dist0[dro + r2.x * 32u], sc, nf);
                    }
                    // every conformer already failed: the pair is invalid whatever follows
                    // (match_utils_numba.py:191-192 tests this after every term; the outcome is the same)
                    if (__all_sync(kFull, !on || (nf > thr))) {
                      dead = true;
                      break;
                    }
                  }
                } else {
                  for (int a = 0; a < cnt1; ++a) {
                    const uint2 r1 = rp1[a];
                    const unsigned dro = r1.x * (unsigned)NL * 32u + (unsigned)lane;
                    const unsigned char* const erows[3] = {edge_b + ((r1.y >> 8) & 255u) * nm16,
                                                           edge_b + ((r1.y >> 16) & 255u) * nm16,
                                                           edge_b + (r1.y >> 24) * nm16};
                    const unsigned M1 = r1.y & 255u;
                    for (int b = 0; b < cnt2; ++b) {
                      const uint2 r2 = rp2[b];
                      term_rows(erows, M1, r2.y, dist0[dro + r2.x * 32u], inv_tab, sc, nf);
                    }
                    if (__all_sync(kFull, !on || (nf > thr))) {
                      dead = true;
                      break;
                    }
                  }
                }
                if (dead) continue;
                const unsigned valid = __ballot_sync(kFull, on && (nf <= thr) && (sc > 0.0f));
                if (valid == 0u) continue;
                if (nrows >= ROWS) {
                  defer = true;
                  break;
                }
                rows_l[(unsigned)nrows * 32u] = sc;
                if (lane == e2 - s2) {
                  myV = valid;
                  myR = nrows;
                }
                ++nrows;
              }
              if (lane < e2_end - s2) {
                const unsigned p = (unsigned)(ws.rowbase[e1] + s2 + lane);
                Vt[p] = myV;
                prow[p] = myR;
              }
            }
          }
        }
        st_rows = (uint32_t)nrows;
        __syncwarp();

        if (!defer) {
          // ================= phase 2: DFS (tree.py:55-104) with an explicit stack
          // The node being expanded keeps its state in registers (uniform over the warp):
          //   d depth = level of its children, cand = bits of the not yet visited candidate entries of that level,
          //   hadc = it had candidates at all, maxm / nmatch (tree.py:93-102), entry (-1: a None node), slot = depth
          //   whose masks / totals it uses (None nodes share their parent's), alive = its conformers, phase = 1 once
          //   its None child has been walked. Lane a (1..d) keeps the pair base of the matched ancestor at depth a.
          uint32_t* const mk = ws.u.b.mk;
          float* const tot_l = &ws.u.b.tot[0][0] + lane;
          for (int e = lane; e < T; e += 32) mk[e] = cfull;  // depth 0: moff[0] = 0
          tot_l[0] = 0.0f;
          __syncwarp();
          st_nodes = 1;
          int d = 0, slot = 0, nmatch = 0, entry = -1, maxm = 0, phase = 0;
          unsigned alive = cfull;
          int my_pbase = kNoBase;
#if PM_FAST_PREFETCH
          int pf_found = -1, pf_row = -1;  // prefetched row indices of the next sibling (lane a: ancestor at depth a)
#endif
          unsigned cand = (ws.lev_start[1] >= 32) ? kFull : ((1u << ws.lev_start[1]) - 1u);  // every entry of level 0
          bool hadc = true;
#if PM_FAST_HEAVY_CHECK >= 3
          for (;;) {  // (re-entered when the list of heavy ligands has no room for this one)
#endif
          for (;;) {
            const int y = d;
            const uint32_t* pm = mk + ws.moff[slot];
            const bool is_anc = lane >= 1 && lane <= d && my_pbase != kNoBase;
            bool do_return = false;
            if (y == L - 1) {
              // ---- every child of this node is a leaf (graph_match.py:103-109): take them all in one pass
              const int s = ws.lev_start[y];
              const int e = s + lane;
              const unsigned lmw = (e < T) ? pm[e] : 0u;
              const float tt = tot_l[slot * 32];
              const int nleaf = leaf_pass(ws, rows_l, prow, s, lmw, cand, tt, is_anc, my_pbase, d, lane, best);
              st_nodes += nleaf;
              st_leaves += nleaf;
              maxm = nleaf > 0 ? 1 : 0;
              if (nleaf == 0 || nmatch + 1 < PMNET_MIN_MATCHES) {
                ++st_nodes;  // the None leaf (tree.py:98)
                ++st_leaves;
                if ((alive >> lane) & 1u) best = fmaxf(best, tt);
              }
              do_return = true;
            } else if (cand) {
              // ---- matched child (y, found): ClusterMatchTree.__init__ (tree.py:33-41); it is not a leaf
              const int end = ws.lev_start[y + 1];
              const int src = __ffs(cand) - 1;
              cand &= cand - 1;
              const int found = ws.lev_start[y] + src;
              const unsigned alive2 = pm[found];
              ++st_nodes;
#if PM_FAST_HEAVY_CHECK == 0
              if (st_nodes > node_limit) {
                // abandon: the tree is walked by the task rounds (pmnet_score_batch) - if the list has room
                if (heavy_append(args.workspace, args.heavy_list, args.heavy_acc, lig, lane, true) >= 0) {
                  heavy = true;
                  break;
                }
                node_limit = 0xffffffffu;
              }
#endif
#if PM_FAST_PREFETCH
              const int myrow = (pf_found == found) ? pf_row : (is_anc ? prow[my_pbase + found] : -1);
              // the next sibling's row indices are requested now (used unless this child is pushed in between)
              pf_found = -1;
              if (cand) {
                pf_found = ws.lev_start[y] + __ffs(cand) - 1;
                pf_row = is_anc ? prow[my_pbase + pf_found] : -1;
              }
#else
              const int myrow = is_anc ? prow[my_pbase + found] : -1;
#endif
              const unsigned sr = ws.srow[found];
              const int pbc = ws.rowbase[found];
              // the child's candidate masks: parent mask & conformers alive in the child & pair validity
              const int e2a = end + lane;
              const unsigned nmw = (e2a < T) ? (pm[e2a] & alive2 & Vt[pbc + e2a]) : 0u;
              const unsigned balc = __ballot_sync(kFull, nmw != 0u);
              bool anylater = balc != 0u;
              if (T - end > 32) {
                unsigned more = 0;
                for (int e2 = e2a + 32; e2 < T; e2 += 32) more |= pm[e2] & alive2 & Vt[pbc + e2];
                anylater |= __any_sync(kFull, more != 0u);
              }
              float t = tot_l[slot * 32];
              if (sr != 0xffffu) t = __fadd_rn(t, PM_ROW_LD(rows_l + sr * 32u));
              // pair rows with the matched ancestors (tree.py:78-82)
              float acc = 0.0f;
#if PM_FAST_ANC_WIDTH == 4
              for (int a0 = 1; a0 <= d; a0 += 4) {  // four independent row loads per round (lanes > d hold -1)
                const int r0 = __shfl_sync(kFull, myrow, a0), r1 = __shfl_sync(kFull, myrow, a0 + 1);
                const int r2 = __shfl_sync(kFull, myrow, a0 + 2), r3 = __shfl_sync(kFull, myrow, a0 + 3);
                const float v0 = (r0 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r0 * 32u) : 0.0f;
                const float v1 = (r1 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r1 * 32u) : 0.0f;
                const float v2 = (r2 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r2 * 32u) : 0.0f;
                const float v3 = (r3 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r3 * 32u) : 0.0f;
                acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, v0), v1), v2), v3);
              }
#else
              for (int a0 = 1; a0 <= d; a0 += 2) {  // two independent row loads per round (lanes > d hold -1)
                const int r0 = __shfl_sync(kFull, myrow, a0), r1 = __shfl_sync(kFull, myrow, a0 + 1);
                const float v0 = (r0 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r0 * 32u) : 0.0f;
                const float v1 = (r1 >= 0) ? PM_ROW_LD(rows_l + (unsigned)r1 * 32u) : 0.0f;
                acc = __fadd_rn(__fadd_rn(acc, v0), v1);
              }
#endif
              const float tc = __fadd_rn(t, acc);
              if (!anylater) {
                // no candidate left at any later level: the child's subtree is the chain of None nodes down to the
                // None leaf (tree.py:98 with no surviving candidate at every level) - account for it in place
                st_nodes += L - d - 1;
                ++st_leaves;
                if ((alive2 >> lane) & 1u) best = fmaxf(best, tc);
                maxm = max(maxm, 1);
                continue;
              }
              if (d + 2 == L) {
                // the child's own children are leaves and their masks are in registers (lane = entry of the last
                // level): evaluate them here with the child as one more ancestor
                const bool is_anc2 = is_anc || lane == d + 1;
                const int pb2 = (lane == d + 1) ? pbc : my_pbase;
                const int nleaf = leaf_pass(ws, rows_l, prow, end, nmw, balc, tc, is_anc2, pb2, d + 1, lane, best);
                st_nodes += nleaf;
                st_leaves += nleaf;
                if (nmatch + 2 < PMNET_MIN_MATCHES) {
                  ++st_nodes;  // the child's None leaf (tree.py:98: too few matches on the path)
                  ++st_leaves;
                  if ((alive2 >> lane) & 1u) best = fmaxf(best, tc);
                }
                maxm = max(maxm, 2);  // the child returns 1 (a matched leaf) + 1 (itself)
                continue;
              }
#if PM_FAST_HEAVY_CHECK == 1
              if (st_nodes > node_limit) {
                // abandon: the tree is walked by the task rounds (pmnet_score_batch) - if the list has room
                if (heavy_append(args.workspace, args.heavy_list, args.heavy_acc, lig, lane, true) >= 0) {
                  heavy = true;
                  break;
                }
                node_limit = 0xffffffffu;
              }
#endif
              // push the child. (depth d + 1's mask slot was last read by other lanes while the previous child's
              // subtree was walked)
              __syncwarp();
              uint32_t* nm_ = mk + ws.moff[d + 1];
              if (e2a < T) nm_[e2a] = nmw;
              for (int e2 = e2a + 32; e2 < T; e2 += 32) nm_[e2] = pm[e2] & alive2 & Vt[pbc + e2];
              tot_l[(d + 1) * 32] = tc;
              ws.stack[d] = make_uint4(cand, alive,
                                       (uint32_t)maxm | ((uint32_t)nmatch << 8) | ((uint32_t)(entry + 1) << 16) |
                                           ((uint32_t)slot << 24) | (hadc ? (1u << 30) : 0u) | ((uint32_t)phase << 31),
                                       0u);
              if (lane == d + 1) my_pbase = pbc;
#if PM_FAST_PREFETCH
              pf_found = -1;
#endif
              const int width = ws.lev_start[y + 2] - end;
              cand = balc & (width >= 32 ? kFull : ((1u << width) - 1u));
              hadc = cand != 0u;
              maxm = 0;
              nmatch += 1;
              entry = found;
              slot = d + 1;
              alive = alive2;
              phase = 0;
              d += 1;
              __syncwarp();
              continue;
            } else if (phase == 0 && (!hadc || nmatch + maxm < PMNET_MIN_MATCHES)) {
              // ---- None child (tree.py:98), not a leaf here (y < L - 1): same masks and totals, one level down
              ++st_nodes;
              ws.stack[d] = make_uint4(0u, alive,
                                       (uint32_t)maxm | ((uint32_t)nmatch << 8) | ((uint32_t)(entry + 1) << 16) |
                                           ((uint32_t)slot << 24) | (hadc ? (1u << 30) : 0u) | (1u << 31),
                                       0u);
              if (lane == d + 1) my_pbase = kNoBase;
#if PM_FAST_PREFETCH
              pf_found = -1;
#endif
              d += 1;
              entry = -1;
              maxm = 0;
              phase = 0;
              const int e = ws.lev_start[d] + lane;
              const unsigned mw = (e < ws.lev_start[d + 1]) ? pm[e] : 0u;
              cand = __ballot_sync(kFull, mw != 0u);
              hadc = cand != 0u;
              continue;
            } else {
              do_return = true;
            }
            if (do_return) {
              if (d == 0) break;
              const int ret = maxm + (entry >= 0 ? 1 : 0);
              const uint4 sv = ws.stack[d - 1];
              cand = sv.x;
              alive = sv.y;
              maxm = max((int)(sv.z & 255u), ret);
              nmatch = (int)((sv.z >> 8) & 255u);
              entry = (int)((sv.z >> 16) & 255u) - 1;
              slot = (int)((sv.z >> 24) & 63u);
              hadc = (sv.z >> 30) & 1u;
              phase = (int)(sv.z >> 31);
              d -= 1;
#if PM_FAST_HEAVY_CHECK == 3
              if (st_nodes > node_limit) {
                heavy = true;
                break;
              }
#elif PM_FAST_HEAVY_CHECK == 4
              if (d <= 3 && st_nodes > node_limit) {
                heavy = true;
                break;
              }
#endif
            }
          }
#if PM_FAST_HEAVY_CHECK >= 3
            // over the node budget: the tree goes to the task rounds (pmnet_score_batch) - if the list has room
            if (!heavy || heavy_append(args.workspace, args.heavy_list, args.heavy_acc, lig, lane, true) >= 0) break;
            heavy = false;
            node_limit = 0xffffffffu;
          }
#endif
          // mean over conformers (graph_match.py:109)
          double s = on ? (double)best : 0.0;
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
          score_out = (float)(s / (double)C);
        }
      }
    }
    if (defer) {
      status = PMNET_LIG_DEFERRED;
      score_out = 0.0f;
      best = 0.0f;
      st_nodes = st_leaves = st_rows = 0;
      // the general kernel's queue: (roughly) longest first like this kernel's own, one ligand per warp at a time
      if (args.defer_list != nullptr && lane == 0)
        args.defer_list[atomicAdd((unsigned int*)args.workspace + 5, 1u)] = lig;
    } else if (heavy) {
      status = PMNET_LIG_HEAVY;
      score_out = 0.0f;
    }
    if (lane == 0) {
      args.out_scores[lig] = score_out;
      args.out_status[lig] = status;
      if (args.out_stats) {
        uint32_t* o = args.out_stats + (size_t)lig * 4;
        o[0] = st_nodes;
        o[1] = st_leaves;
        o[2] = st_rows;
        o[3] = st_pairs;
      }
    }
    if (args.out_conf) args.out_conf[(size_t)lig * 32 + lane] = (status == PMNET_LIG_OK) ? best : 0.0f;
    __syncwarp();
  }
}

}  // namespace fastk
