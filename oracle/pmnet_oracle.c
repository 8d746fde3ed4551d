/* pmnet_oracle.c - TEST INFRASTRUCTURE. CPU restatement of the reference's scoring path, one ligand at a time.
 *
 * This file is the parity oracle of the CUDA kernel. It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Parity is pinned against outputs of the real reference (GraphMatcher.run imported from /root/reference in the
 * build container, oracle/make_golden.py) stored in tests/golden/; see tests/test_oracle_golden.py.
 *
 * It deliberately follows the reference's own structure - full pair-score table first, then a recursive DFS
 * that carries per-candidate {conformer: accumulated score} maps - and not the kernel's (bit masks + lazy
 * sums), so that the two are independent statements of the same algorithm.
 *
 *   candidates / level order / truncation     graph_match.py:85-92, 124-137
 *   node matches                               graph_match.py:139-172
 *   pair-score table + cluster prefilter       graph_match.py:222-279
 *   pair term, fail counting                   match_utils_numba.py:12-86, 163-197
 *   self term                                  match_utils_numba.py:89-151, 200-231
 *   DFS                                        tree.py:16-43, 55-104, 219-227
 *   final average                              graph_match.py:103-109
 *   ligand edge distance / cluster centre+size ligand.py:349-351, 458-473
 *
 * Numerics follow what numba compiles (checked on the JIT's LLVM IR / x86 assembly): r = 1.0f/std,
 * s = (d - mu)*r, s2 = s*s and w2*r in fp32; exp and all products/sums after it in fp64; each call's result is
 * added to an fp32 score array. Tree sums are fp64 (Python floats). Build with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "../include/pmnet_b200.h"

#define MAXC 128         /* conformers the oracle handles per ligand */
#define MAXLEV PMNET_MAX_DEPTH

typedef struct {
  int node;              /* ligand-local node id */
  int n;                 /* number of matched model nodes */
  int m[32];             /* matched model node indices (cluster order) */
} NodeMatch;

typedef struct {
  int n;                 /* node matches in this (ligand cluster, model cluster) */
  NodeMatch* nm;
} MatchList;

typedef struct {
  const PmModel* model;
  const float* w;        /* 7 weights */
  int C;
  int stride;
  const float* xyz;      /* this ligand's coordinates */
  int L;                 /* levels */
  int lev_cluster[MAXLEV];      /* global cluster index (into batch CSR) of each level */
  int lev_start[MAXLEV + 1];    /* first entry of each level */
  int T;                        /* total entries */
  int* ent_mc;                  /* model cluster of each entry */
  MatchList* ent_match;
  double* self_score;           /* [T][C] */
  double* pair_score;           /* [T][T][C], valid for level(e1) < level(e2); -1 = invalid */
  double best[MAXC];
  uint64_t n_tree_nodes, n_leaves;
} Ctx;

static inline float lig_coord(const Ctx* x, int node, int axis, int c) {
  return x->xyz[(size_t)(node * 3 + axis) * x->stride + c];
}

/* LigandEdge.set_distances (ligand.py:349-351): np.linalg.norm of an fp32 difference, fp32 throughout */
static inline float node_distance(const Ctx* x, int n1, int n2, int c) {
  float dx = lig_coord(x, n1, 0, c) - lig_coord(x, n2, 0, c);
  float dy = lig_coord(x, n1, 1, c) - lig_coord(x, n2, 1, c);
  float dz = lig_coord(x, n1, 2, c) - lig_coord(x, n2, 2, c);
  float s = dx * dx;
  s = s + dy * dy;
  s = s + dz * dz;
  return sqrtf(s);
}

/* __numba_run / __numba_run_self (match_utils_numba.py:12-151) for one ligand-node pair */
static void term(const Ctx* x, const NodeMatch* a, const NodeMatch* b, float* score, short* fail) {
  const PmModel* md = x->model;
  const int M = a->n, N = b->n, nm = md->n_nodes;
  const int num_match = M * N;
  const int pass_threshold = (num_match + 1) / 2;
  double W1 = 0.0, W2 = 0.0;
  for (int i = 0; i < M; ++i) W1 += (double)x->w[md->node_type[a->m[i]]];
  for (int j = 0; j < N; ++j) W2 += (double)x->w[md->node_type[b->m[j]]];
  const double normalize_coeff = 1.0 / (W1 * W2);
  const double score_coeff = (W1 * W2) / (double)num_match;
  for (int c = 0; c < x->C; ++c) {
    const float d = node_distance(x, a->node, b->node, c);
    int num_pass = 0;
    double likelihood = 0.0;
    for (int i = 0; i < M; ++i) {
      const float w1 = x->w[md->node_type[a->m[i]]];
      double l = 0.0;
      for (int j = 0; j < N; ++j) {
        const float w2 = x->w[md->node_type[b->m[j]]];
        const float mu = md->edge_mu[a->m[i] * nm + b->m[j]];
        const float sd = md->edge_sigma[a->m[i] * nm + b->m[j]];
        const float r = 1.0f / sd;
        const float s = (d - mu) * r;
        const float s2 = s * s;
        const float wr = w2 * r;
        l += (double)wr * exp(-0.5 * (double)s2);
        if (s2 < 4.0f) ++num_pass;
      }
      likelihood += (double)w1 * l;
    }
    score[c] = (float)((double)score[c] + likelihood * normalize_coeff * score_coeff);
    if (fail && num_pass < pass_threshold) fail[c] += 1;
  }
}

/* scoring_matching_pair (match_utils_numba.py:163-197) */
static void matching_pair(const Ctx* x, const MatchList* l1, const MatchList* l2, double* out) {
  const int C = x->C;
  const double thr = (double)l1->n * (double)l2->n * 0.5;
  float score[MAXC];
  short fail[MAXC];
  memset(score, 0, sizeof score);
  memset(fail, 0, sizeof fail);
  for (int i = 0; i < l1->n; ++i)
    for (int j = 0; j < l2->n; ++j) {
      term(x, &l1->nm[i], &l2->nm[j], score, fail);
      short mn = fail[0];
      for (int c = 1; c < C; ++c) mn = fail[c] < mn ? fail[c] : mn;
      if ((double)mn > thr) {
        for (int c = 0; c < C; ++c) out[c] = -1.0;
        return;
      }
    }
  for (int c = 0; c < C; ++c) out[c] = ((double)fail[c] <= thr) ? (double)score[c] : -1.0;
}

/* scoring_matching_self (match_utils_numba.py:200-231) */
static void matching_self(const Ctx* x, const MatchList* l, double* out) {
  float score[MAXC];
  memset(score, 0, sizeof score);
  for (int i = 0; i < l->n; ++i)
    for (int j = i + 1; j < l->n; ++j) term(x, &l->nm[i], &l->nm[j], score, NULL);
  for (int c = 0; c < x->C; ++c) out[c] = (double)score[c];
}

/* LigandNodeCluster.center / .size (ligand.py:458-473), fp32 */
static void cluster_geometry(const Ctx* x, const uint8_t* nodes, int n, float* ctr /*[C][3]*/, float* size /*[C]*/) {
  for (int c = 0; c < x->C; ++c) {
    for (int a = 0; a < 3; ++a) {
      float s = lig_coord(x, nodes[0], a, c);
      for (int k = 1; k < n; ++k) s = s + lig_coord(x, nodes[k], a, c);
      ctr[c * 3 + a] = s / (float)n;
    }
    float mx = 0.0f;
    for (int k = 0; k < n; ++k) {
      float dx = lig_coord(x, nodes[k], 0, c) - ctr[c * 3 + 0];
      float dy = lig_coord(x, nodes[k], 1, c) - ctr[c * 3 + 1];
      float dz = lig_coord(x, nodes[k], 2, c) - ctr[c * 3 + 2];
      float s = dx * dx;
      s = s + dy * dy;
      s = s + dz * dz;
      float d = sqrtf(s);
      if (k == 0 || d > mx) mx = d;
    }
    size[c] = mx;
  }
}

/* ClusterMatchTree.dfs_run (tree.py:55-104). acc_sum/acc_has hold `match_dict` for entries of levels > level:
 * acc_has[e*C+c] says conformer c is a key of match_dict[ligand_cluster(e)][model_cluster(e)]. */
static int dfs(Ctx* x, int level, int entry /* -1 = None */, int num_matches, const double* total,
               const unsigned char* alive, const double* acc_sum, const unsigned char* acc_has) {
  const int C = x->C, T = x->T, L = x->L;
  x->n_tree_nodes++;
  double* upd_sum = NULL;
  unsigned char* upd_has = NULL;
  const double* cur_sum = acc_sum;
  const unsigned char* cur_has = acc_has;
  if (entry >= 0) {
    upd_sum = (double*)malloc((size_t)T * C * sizeof(double));
    upd_has = (unsigned char*)calloc((size_t)T * C, 1);
    for (int e = x->lev_start[level + 1]; e < T; ++e) {
      const double* ps = x->pair_score + ((size_t)entry * T + e) * C;
      for (int c = 0; c < C; ++c) {
        if (acc_has[e * C + c] && alive[c] && ps[c] > 0.0) {
          upd_has[e * C + c] = 1;
          upd_sum[e * C + c] = acc_sum[e * C + c] + ps[c];
        }
      }
    }
    cur_sum = upd_sum;
    cur_has = upd_has;
  }
  int ret;
  if (level < L - 1) {
    const int y = level + 1;
    int max_num_matches = 0, n_children = 0;
    for (int e = x->lev_start[y]; e < x->lev_start[y + 1]; ++e) {
      int any = 0;
      for (int c = 0; c < C; ++c) any |= cur_has[e * C + c];
      if (!any) continue;
      ++n_children;
      /* ClusterMatchTree.__init__ (tree.py:33-41) */
      double child_total[MAXC];
      unsigned char child_alive[MAXC];
      for (int c = 0; c < C; ++c) {
        child_alive[c] = cur_has[e * C + c];
        child_total[c] = child_alive[c] ? total[c] + x->self_score[(size_t)e * C + c] + cur_sum[e * C + c] : 0.0;
      }
      int r = dfs(x, y, e, num_matches + 1, child_total, child_alive, cur_sum, cur_has);
      if (r > max_num_matches) max_num_matches = r;
    }
    if (n_children == 0 || num_matches + max_num_matches < PMNET_MIN_MATCHES) {
      int r = dfs(x, y, -1, num_matches, total, alive, cur_sum, cur_has);
      if (r > max_num_matches) max_num_matches = r;
    }
    ret = max_num_matches + (entry >= 0);
  } else {
    /* leaf: GraphMatcher._run_average (graph_match.py:103-109) */
    x->n_leaves++;
    for (int c = 0; c < C; ++c)
      if (alive[c] && total[c] > x->best[c]) x->best[c] = total[c];
    ret = (entry >= 0);
  }
  free(upd_sum);
  free(upd_has);
  return ret;
}

/* One ligand. Returns the status code; *out_score gets GraphMatcher.run()'s value. */
static int score_one(const PmModel* md, const PmLigandBatch* b, const float* w, int lig, double* out_score,
                     double* out_conf, uint64_t* out_stats) {
  Ctx x;
  memset(&x, 0, sizeof x);
  x.model = md;
  x.w = w;
  x.C = b->n_conf[lig];
  *out_score = 0.0;
  if (x.C < 1 || x.C > MAXC) return PMNET_LIG_UNSUPPORTED;
  x.stride = (x.C + 3) & ~3;
  x.xyz = b->coords + (b->coord_off[lig] - b->coord_base);
  const uint8_t* tmask = b->node_type_mask + (b->lig_node_off[lig] - b->node_base);
  const int q0 = b->lig_cluster_off[lig] - b->cluster_base, q1 = b->lig_cluster_off[lig + 1] - b->cluster_base;
  const uint8_t* cl_nodes = b->cluster_nodes - b->cnode_base; /* indexed with the stored (un-rebased) offsets */
  const int Km = md->n_clusters;

  /* levels = clusters with >= 1 candidate model cluster, in priority order, first 20 (graph_match.py:85-88) */
  uint8_t lev_mask[MAXLEV];
  for (int q = q0; q < q1 && x.L < MAXLEV; ++q) {
    uint8_t m = 0;
    for (int i = b->cluster_node_off[q]; i < b->cluster_node_off[q + 1]; ++i) m |= tmask[cl_nodes[i]];
    int any = 0;
    for (int k = 0; k < Km; ++k) any |= (md->cluster_mask[k] & m) != 0;
    if (any) {
      lev_mask[x.L] = m;
      x.lev_cluster[x.L++] = q;
    }
  }
  if (x.L == 0) return PMNET_LIG_EMPTY;

  /* entries = (level, candidate model cluster) in model.node_clusters order (graph_match.py:124-137) */
  int T = 0;
  for (int l = 0; l < x.L; ++l) {
    x.lev_start[l] = T;
    for (int k = 0; k < Km; ++k) T += (md->cluster_mask[k] & lev_mask[l]) != 0;
  }
  x.lev_start[x.L] = T;
  x.T = T;
  x.ent_mc = (int*)malloc(sizeof(int) * T);
  x.ent_match = (MatchList*)calloc(T, sizeof(MatchList));
  int* ent_level = (int*)malloc(sizeof(int) * T);
  {
    int e = 0;
    for (int l = 0; l < x.L; ++l)
      for (int k = 0; k < Km; ++k)
        if (md->cluster_mask[k] & lev_mask[l]) {
          ent_level[e] = l;
          x.ent_mc[e++] = k;
        }
  }
  /* node matches (graph_match.py:139-172) */
  for (int e = 0; e < T; ++e) {
    const int q = x.lev_cluster[ent_level[e]], k = x.ent_mc[e];
    const int nn = b->cluster_node_off[q + 1] - b->cluster_node_off[q];
    MatchList* ml = &x.ent_match[e];
    ml->nm = (NodeMatch*)malloc(sizeof(NodeMatch) * (nn > 0 ? nn : 1));
    for (int i = 0; i < nn; ++i) {
      const int node = cl_nodes[b->cluster_node_off[q] + i];
      NodeMatch t;
      t.node = node;
      t.n = 0;
      for (int j = md->cluster_node_off[k]; j < md->cluster_node_off[k + 1]; ++j) {
        const int mn = md->cluster_nodes[j];
        if ((tmask[node] >> md->node_type[mn]) & 1) {
          if (t.n < 32) t.m[t.n] = mn;
          t.n++;
        }
      }
      if (t.n > 32) { /* outside what the oracle models */
        *out_score = NAN;
        return PMNET_LIG_UNSUPPORTED;
      }
      if (t.n > 0) ml->nm[ml->n++] = t;
    }
  }

  /* pair-score table (graph_match.py:222-279) */
  const int C = x.C;
  x.self_score = (double*)calloc((size_t)T * C, sizeof(double));
  x.pair_score = (double*)malloc((size_t)T * T * C * sizeof(double));
  for (size_t i = 0; i < (size_t)T * T * C; ++i) x.pair_score[i] = -1.0;
  float* ctr = (float*)malloc(sizeof(float) * x.L * C * 3);
  float* size = (float*)malloc(sizeof(float) * x.L * C);
  for (int l = 0; l < x.L; ++l) {
    const int q = x.lev_cluster[l];
    cluster_geometry(&x, cl_nodes + b->cluster_node_off[q], b->cluster_node_off[q + 1] - b->cluster_node_off[q],
                     ctr + (size_t)l * C * 3, size + (size_t)l * C);
  }
  uint64_t n_pair_entries = 0;
  for (int e = 0; e < T; ++e) matching_self(&x, &x.ent_match[e], x.self_score + (size_t)e * C);
  for (int i = 0; i < x.L; ++i)
    for (int j = i + 1; j < x.L; ++j) {
      float ldist[MAXC], lsize[MAXC];
      for (int c = 0; c < C; ++c) {
        const float* a = ctr + ((size_t)i * C + c) * 3;
        const float* bb = ctr + ((size_t)j * C + c) * 3;
        float dx = a[0] - bb[0], dy = a[1] - bb[1], dz = a[2] - bb[2];
        float s = dx * dx;
        s = s + dy * dy;
        s = s + dz * dz;
        ldist[c] = sqrtf(s);
        lsize[c] = size[i * C + c] + size[j * C + c];
      }
      for (int e1 = x.lev_start[i]; e1 < x.lev_start[i + 1]; ++e1)
        for (int e2 = x.lev_start[j]; e2 < x.lev_start[j + 1]; ++e2) {
          ++n_pair_entries;
          const int k = x.ent_mc[e1], l = x.ent_mc[e2];
          const float mdist = md->cluster_dist[k * Km + l];
          const float msize = md->cluster_size_sum[k * Km + l];
          float mn = fabsf(ldist[0] - mdist) - lsize[0];
          for (int c = 1; c < C; ++c) {
            float v = fabsf(ldist[c] - mdist) - lsize[c];
            mn = v < mn ? v : mn;
          }
          double* out = x.pair_score + ((size_t)e1 * T + e2) * C;
          if (mn > msize) continue; /* NO_MATCH_SCORE, already -1 */
          matching_pair(&x, &x.ent_match[e1], &x.ent_match[e2], out);
        }
    }

  /* ClusterMatchTreeRoot.run (tree.py:219-227) */
  double total[MAXC];
  unsigned char alive[MAXC];
  for (int c = 0; c < C; ++c) {
    total[c] = 0.0;
    alive[c] = 1;
    x.best[c] = 0.0;
  }
  double* acc_sum = (double*)calloc((size_t)T * C, sizeof(double));
  unsigned char* acc_has = (unsigned char*)malloc((size_t)T * C);
  memset(acc_has, 1, (size_t)T * C);
  dfs(&x, -1, -1, 0, total, alive, acc_sum, acc_has);

  double s = 0.0;
  for (int c = 0; c < C; ++c) s += x.best[c];
  *out_score = s / (double)C;
  if (out_conf)
    for (int c = 0; c < C; ++c) out_conf[c] = x.best[c];
  if (out_stats) {
    out_stats[0] = x.n_tree_nodes;
    out_stats[1] = x.n_leaves;
    out_stats[2] = (uint64_t)T;
    out_stats[3] = n_pair_entries;
  }
  free(acc_sum);
  free(acc_has);
  free(ctr);
  free(size);
  free(x.pair_score);
  free(x.self_score);
  for (int e = 0; e < T; ++e) free(x.ent_match[e].nm);
  free(x.ent_match);
  free(x.ent_mc);
  free(ent_level);
  return PMNET_LIG_OK;
}

/* Score ligands [begin, end) of a HOST batch with `threads` POSIX threads (0 = all online cores), ligands handed
 * out dynamically in blocks of 4. out_scores [end-begin] fp64; out_conf optional [(end-begin)*conf_stride] fp64;
 * out_status int32; out_stats optional [(end-begin)*4] uint64 {tree nodes, leaves, entries, pair entries}.
 * All pointers inside the structs are HOST pointers here. */
typedef struct {
  const PmModel* model;
  const PmLigandBatch* batch;
  const float* weights;
  int begin, end;
  double* out_scores;
  double* out_conf;
  int conf_stride;
  int32_t* out_status;
  uint64_t* out_stats;
  atomic_int next;
} Job;

static void* worker(void* arg) {
  Job* j = (Job*)arg;
  for (;;) {
    int i0 = atomic_fetch_add(&j->next, 4);
    if (i0 >= j->end) break;
    int i1 = i0 + 4 < j->end ? i0 + 4 : j->end;
    for (int i = i0; i < i1; ++i) {
      double s = 0.0;
      int st = score_one(j->model, j->batch, j->weights, i, &s,
                         j->out_conf ? j->out_conf + (size_t)(i - j->begin) * j->conf_stride : NULL,
                         j->out_stats ? j->out_stats + (size_t)(i - j->begin) * 4 : NULL);
      j->out_scores[i - j->begin] = s;
      if (j->out_status) j->out_status[i - j->begin] = st;
    }
  }
  return NULL;
}

int pmnet_oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int pmnet_oracle_score(const PmModel* model, const PmLigandBatch* batch, const float* weights, int begin, int end,
                       double* out_scores, double* out_conf, int conf_stride, int32_t* out_status,
                       uint64_t* out_stats, int threads) {
  if (!model || !batch || !weights || begin < 0 || end > batch->n_ligands) return PMNET_EINVAL;
  Job job = {model, batch, weights, begin, end, out_scores, out_conf, conf_stride, out_status, out_stats, 0};
  atomic_store(&job.next, begin);
  if (threads <= 0) threads = pmnet_oracle_max_threads();
  if (threads > 256) threads = 256;
  if (threads == 1) {
    worker(&job);
    return PMNET_OK;
  }
  pthread_t tid[256];
  int started = 0;
  for (int t = 0; t < threads; ++t)
    if (pthread_create(&tid[started], NULL, worker, &job) == 0) ++started;
  if (started == 0) worker(&job);
  for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
  return PMNET_OK;
}
