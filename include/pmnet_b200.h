/* pmnet_b200.h - C ABI of the B200-native PharmacoNet screening hot path.
 *
 * The reference (SeonghwanSeo/PharmacoNet) is pure Python and has no FFI; the interface this library replaces
 * is the Python call chain
 *     PharmacophoreModel._scoring(ligand, weights)      src/pmnet/pharmacophore_model.py:101-106
 *       -> GraphMatcher(model, ligand, weights).run()   src/pmnet/scoring/graph_match.py:63-101
 *            -> scoring_matching_pair / _self           src/pmnet/scoring/match_utils_numba.py:163-231
 *            -> ClusterMatchTreeRoot.run / dfs_run      src/pmnet/scoring/tree.py:55-104, 219-227
 *            -> _run_average                            src/pmnet/scoring/graph_match.py:103-109
 * evaluated for a whole library (screening.py:46-68) in one call. The ctypes binding a maintainer of the
 * reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types. All array pointers are DEVICE pointers unless the
 *    name says host. The caller owns every buffer; the library never allocates or frees device memory.
 *  - all work is enqueued on the caller's CUDA stream (passed as void*, i.e. cudaStream_t); no implicit
 *    synchronisation.
 *  - return value 0 = OK, otherwise a PMNET_E* code; pmnet_last_error_string() gives a thread-local message.
 *    Nothing is thrown across the ABI. Per-ligand problems are reported in out_status, not as call failures.
 */
#ifndef PMNET_B200_H
#define PMNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMNET_ABI_VERSION 4

/* pharmacophore types, bit positions of every type mask (graph_match.py:32-40) */
enum {
  PMNET_HYDROPHOBIC = 0,
  PMNET_AROMATIC = 1,
  PMNET_CATION = 2,
  PMNET_ANION = 3,
  PMNET_HBOND_DONOR = 4,
  PMNET_HBOND_ACCEPTOR = 5,
  PMNET_HALOGEN = 6,
  PMNET_NUM_TYPES = 7
};

enum {
  PMNET_OK = 0,
  PMNET_EINVAL = 1,      /* bad argument / shape */
  PMNET_EWORKSPACE = 2,  /* workspace too small */
  PMNET_ELIMIT = 3,      /* model exceeds a compiled limit (nodes, clusters, shared memory) */
  PMNET_ECUDA = 4        /* a CUDA runtime call failed */
};

/* per-ligand status written to out_status */
enum {
  PMNET_LIG_OK = 0,
  PMNET_LIG_EMPTY = 1,    /* no cluster / no candidate: score 0 (graph_match.py:95-99), not an error */
  PMNET_LIG_OVERFLOW = 2, /* per-warp scratch exhausted: score not computed; re-run with a larger scratch */
  PMNET_LIG_UNSUPPORTED = 3, /* more conformers than PMNET_MAX_CONFORMERS */
  PMNET_LIG_DEFERRED = 4,    /* internal: left by the specialised kernel for the general kernel that pmnet_score_batch
                                enqueues behind it in the same call; never visible once the stream has drained */
  PMNET_LIG_HEAVY = 5        /* internal: the tree outgrew PmScoreConfig.heavy_budget and was handed to the task-parallel
                                walk of the same call; never visible once the stream has drained */
};

#define PMNET_MAX_CONFORMERS 128  /* one warp lane per conformer, up to 4 conformers per lane */
#define PMNET_MAX_DEPTH 20        /* graph_match.py:88 */
#define PMNET_MIN_MATCHES 5       /* tree.py:98 */

/* Pharmacophore model tables (pharmacophore_model.py:51-58, 207-365), flat.
 * Edge tables are complete and symmetric and include the self loops (density_map.py:66-72). */
typedef struct PmModel {
  int32_t n_nodes;                 /* Nm <= 255 */
  int32_t n_clusters;              /* Km <= 255 */
  int32_t n_cluster_nodes;         /* = cluster_node_off[Km] (host copy, so that no device read-back is needed) */
  int32_t reserved;
  const uint8_t* node_type;        /* [Nm]     pharmacophore type index 0..6 of each model node */
  const float* edge_mu;            /* [Nm*Nm]  ModelEdge.distance_mean as fp32 */
  const float* edge_sigma;         /* [Nm*Nm]  ModelEdge.distance_std  as fp32 */
  const uint8_t* cluster_mask;     /* [Km]     bit t set iff type t in ModelNodeCluster.node_types */
  const int32_t* cluster_node_off; /* [Km+1] */
  const uint8_t* cluster_nodes;    /* [cluster_node_off[Km]] model node indices, ascending per cluster */
  const float* cluster_dist;       /* [Km*Km]  |centre_k - centre_l|  (graph_match.py:258-260) as fp32 */
  const float* cluster_size_sum;   /* [Km*Km]  size_k + size_l        (graph_match.py:261) as fp32 */
} PmModel;

/* A library (or shard / chunk of one) of typed ligand graphs in CSR form (ligand.py:110-259).
 * Clusters of a ligand are stored in matcher priority order (graph_match.py:43-60, 87).
 * coords: node n, axis a, conformer c of ligand i at
 *     coords[coord_off[i] + (n*3 + a)*stride_i + c],  stride_i = (n_conf[i] + 3) & ~3
 * so that a warp (lane = conformer) reads one contiguous, 16 B-aligned row per (node, axis). */
typedef struct PmLigandBatch {
  int32_t n_ligands;
  const int32_t* lig_node_off;     /* [n+1]  first node of each ligand in node_type_mask */
  const int32_t* lig_cluster_off;  /* [n+1]  first cluster of each ligand */
  const int32_t* cluster_node_off; /* [total clusters + 1] */
  const uint8_t* cluster_nodes;    /* ligand-local node ids, high-priority node first (ligand.py:387-395) */
  const uint8_t* node_type_mask;   /* [total nodes] 7-bit mask of LigandNode.types */
  const int32_t* n_conf;           /* [n] conformers per ligand, 1..PMNET_MAX_CONFORMERS (see PmScoreConfig) */
  const int64_t* coord_off;        /* [n+1] in floats */
  const float* coords;             /* fp32 node coordinates (LigandNode.positions, ligand.py:293-301) */
  /* A chunk of a larger library can be passed without re-basing its CSR offsets: the values stored in the
   * offset arrays are relative to these bases (0 for a self-contained batch), i.e. the data arrays passed here
   * start at node `node_base`, cluster `cluster_base`, cluster-node `cnode_base`, float `coord_base`. */
  int64_t coord_base;
  int32_t node_base;
  int32_t cluster_base;
  int32_t cnode_base;
  int32_t reserved;
  /* Optional processing order: [n_ligands] ligand indices (a permutation of 0..n-1), or NULL for index order.
   * The persistent grid hands ligands to warps in this order; pmnet_cost_order fills it with the ligands sorted by
   * decreasing expected work, which removes most of the end-of-launch tail (the cost of a ligand varies 100x).
   * Results do not depend on it: every output array is indexed by ligand. */
  const int32_t* order;
} PmLigandBatch;

/* Launch configuration; zero-initialise for defaults. */
typedef struct PmScoreConfig {
  int32_t warps_per_block;   /* default 32 (8 when max_conformers > 32) */
  int32_t blocks;            /* default: 1 x SM count (2 x when max_conformers > 32) */
  int32_t scratch_rows;      /* per-warp pair-table capacity in rows; default 8192 */
  int32_t max_conformers;    /* largest n_conf in the batch (default 32): selects 1, 2 or 4 conformers per lane;
                                ligands with more conformers than the launch was sized for get PMNET_LIG_UNSUPPORTED */
  int32_t rescore_status;    /* 0: score every ligand. Otherwise only the ligands whose out_status[i] currently equals
                                this PMNET_LIG_* code are scored (the others keep score and status): re-running the
                                PMNET_LIG_OVERFLOW ligands of a previous call with a larger scratch needs no host copy
                                of the library and no compaction */
  int32_t heavy_budget;      /* tree nodes after which a ligand's tree is shared out: the warp walking it (and every
                                warp that takes a piece) gives unvisited subtrees away to a task queue served inside
                                the same call; 0: default 65536, < 0: never. Not on status-restricted re-runs
                                (rescore_status != 0). Scores, per-conformer scores and tree statistics do not depend
                                on it */
  int32_t reserved[2];
} PmScoreConfig;

int pmnet_abi_version(void);
const char* pmnet_last_error_string(void);

/* Bytes of device workspace pmnet_score_batch needs for this model / launch configuration. */
size_t pmnet_score_workspace_bytes(int32_t n_model_nodes, int32_t n_model_clusters, const PmScoreConfig* cfg);

/* Score every ligand of `batch` against `model` (replaces GraphMatcher.run per ligand).
 *   model, batch : HOST structs holding DEVICE array pointers
 *   weights      : HOST, 7 floats in PMNET_* type order (graph_match.py:32-40, 82-84)
 *   out_scores   : [n_ligands] fp32, mean over conformers of the best leaf score (graph_match.py:103-109)
 *   out_conf_scores : optional [n_ligands * S] per-conformer best leaf score, S = 32 / 64 / 128 for max_conformers
 *                     <= 32 / 64 / 128, or NULL
 *   out_status   : [n_ligands] PMNET_LIG_* code
 *   out_stats    : optional [n_ligands * 4] uint32 {tree nodes, leaves, table rows used, pair entries}, or NULL
 * With an all-default configuration and <= 32 conformers the call enqueues two kernels: the specialised one (all
 * per-ligand tables in shared memory; csrc/scoring_fast.cuh) and, behind it, the general one for the ligands the first
 * left PMNET_LIG_DEFERRED. Both compute the same fp32 operations in the same order: results do not depend on which one
 * scored a ligand. An explicit warps_per_block / blocks / scratch_rows selects the general kernel alone.
 * Unless heavy_budget < 0, the general kernel's work is done by the first of three launches of the task kernel (the
 * same code behind a task queue: trees larger than heavy_budget are walked by many warps - walkers donate subtrees to
 * the queue, idle warps take them; the other two launches normally return at once), followed by the kernel that writes
 * the outputs of those ligands.
 */
int pmnet_score_batch(const PmModel* model, const PmLigandBatch* batch, const float* weights,
                      float* out_scores, float* out_conf_scores, int32_t* out_status, uint32_t* out_stats,
                      void* workspace, size_t workspace_bytes, const PmScoreConfig* cfg, void* stream);

/* Longest-processing-time-first order for pmnet_score_batch: out_order[n_ligands] = ligand indices sorted by
 * decreasing number of (ligand cluster, model cluster) candidate entries over the ligand's first PMNET_MAX_DEPTH
 * matching clusters (graph_match.py:85-92, 124-137) - the size of the pair table and a good predictor of the tree
 * size. Stable: ties keep index order. `batch->order` is ignored. workspace: pmnet_order_workspace_bytes(n). */
size_t pmnet_order_workspace_bytes(int32_t n_ligands);
int pmnet_cost_order(const PmModel* model, const PmLigandBatch* batch, int32_t* out_order, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Keep the k best (score, ligand id) pairs of a shard, descending by score, ties by ascending id
 * (screening.py:70 sorts the whole list; a shard only needs its top k for the final merge).
 *   scores [n] fp32, ids = id_base + index; out_scores [k] fp32, out_ids [k] int64 (padded with -inf / -1).
 *   workspace: pmnet_topk_workspace_bytes(n, k). */
size_t pmnet_topk_workspace_bytes(int64_t n, int32_t k);
int pmnet_topk(const float* scores, int64_t n, int64_t id_base, int32_t k, float* out_scores,
               int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream);

/* 3x3x3 convolution, 96 -> 96 channels, stride 1, zero padding 1, fused per-channel scale/bias (folded
 * BatchNorm3d in eval mode) and optional ReLU: replaces BaseConv3d.forward (src/pmnet/network/nn/layers.py:45-46)
 * for the shape used by FPNDecoder (decoders/fpn_decoder.py:54-66), CavityHead (cavity_head.py:18-37) and the
 * MaskHead decoder (mask_head.py:38-80). tcgen05 implicit GEMM, bf16 operands, fp32 accumulation.
 *   x_c8, y_c8 : bf16 [B][12][D][H][W][8]  (8-channel chunks; channel c = chunk*8 + lane)
 *   w_packed   : bf16 [27][12][96][8]      (tap = (kd*3 + kh)*3 + kw, then C_in chunk, C_out, 8 C_in lanes)
 *   scale,bias : fp32 [96]
 *   head_w/head_b/head_out : optional fused 1x1 conv to ONE channel applied to the activated output
 *                (cavity_head.py:26,36): head_out fp32 [B][D][H][W]; y_c8 may then be NULL to skip the store
 *   planes_per_item : output d-planes per work item (even; 0 = 16), max_ctas : 0 = one per SM
 * D must be even. */
int pmnet_conv3d_k3_c96(const void* x_c8, const void* w_packed, const float* scale, const float* bias, void* y_c8,
                        const float* head_w, float head_b, float* head_out, int32_t B, int32_t D, int32_t H,
                        int32_t W, int32_t relu, int32_t planes_per_item, int32_t max_ctas, void* stream);

/* One pass of a split-precision convolution (same kernel, same layouts). With activations and weights split into bf16
 * pairs x = x_hi + x_lo, w = w_hi + w_lo the product x w = x_hi w_hi + x_lo w_hi + x_hi w_lo (+ 2^-16 relative) is
 * three launches that chain their fp32 accumulators through global memory:
 *   acc_in  : fp32 [B][D][H][W][96] partial sums of the earlier passes (NULL for the first pass)
 *   acc_out : not NULL = store the raw fp32 sum there (may alias acc_in) and skip activation and outputs
 *   the last pass (acc_out NULL) applies scale / bias / ReLU / head to the full sum and writes y_c8 = bf16(y) and, if
 *   y_lo_c8 is given, y_lo_c8 = bf16(y - y_c8): the operand pair of the next layer.
 * This is the opt-in precision mode for outputs that feed a threshold (cavity / segmentation masks, module.py:232-233,
 * 288): with plain bf16 operands a few hundred of 262144 mask voxels flip against the fp32 reference. */
int pmnet_conv3d_k3_c96_pass(const void* x_c8, const void* w_packed, const float* scale, const float* bias, void* y_c8,
                             void* y_lo_c8, const float* acc_in, float* acc_out, const float* head_w, float head_b,
                             float* head_out, int32_t B, int32_t D, int32_t H, int32_t W, int32_t relu,
                             int32_t planes_per_item, int32_t max_ctas, void* stream);

/* FPNDecoder lateral (decoders/fpn_decoder.py:100-111): out = act(scale * (W x) + bias) + nearest_upsample(up).
 *   x      : fp32 [B][C_in][D][H][W] (x_is_c8 = 0) or bf16 c8 [B][C_in/8][D][H][W][8] (x_is_c8 = 1)
 *   w_t    : fp32 [C_in][96] (the 1x1 conv weight, transposed); scale/bias fp32 [96] or NULL for a raw linear map
 *   up_c8  : optional bf16 c8 [B][12][D/2][H/2][W/2][8], added after the activation; out_c8 bf16 c8 [B][12][D][H][W][8] */
int pmnet_lateral_c96(const void* x, int32_t x_is_c8, int32_t c_in, const float* w_t, const float* scale,
                      const float* bias, int32_t relu, const void* up_c8, void* out_c8, int32_t B, int32_t D,
                      int32_t H, int32_t W, void* stream);

/* Same with two-term bf16 splits (x = x + x_lo when x_is_c8, up = up + up_lo; out = bf16(y), out_lo = bf16(y - out));
 * any of the *_lo pointers may be NULL. */
int pmnet_lateral_c96_split(const void* x, const void* x_lo, int32_t x_is_c8, int32_t c_in, const float* w_t,
                            const float* scale, const float* bias, int32_t relu, const void* up_c8, const void* up_lo_c8,
                            void* out_c8, void* out_lo_c8, int32_t B, int32_t D, int32_t H, int32_t W, void* stream);

/* MaskHead.get_box_features (mask_head.py:170-196) pushed through the linear lateral conv: for every box j of a group
 *   out[j] = act(scale * (S + u[j] + [voxel in pvox] pvec[j]) + bias) + nearest_upsample(up[j])
 *   s_c8 : bf16 c8 [12][D][H][W][8] shared by the group (the lateral conv of the pocket's feature map, or the feature
 *          map itself at the top level where the lateral is the identity: then scale = bias = NULL, relu = 0)
 *   u, pvec : fp32 [nbox][96]; pvox int32 [nbox][4] flat voxel ids (-1 = unused) of the token voxels of the box's
 *          group of 4 (every box gets its own pvec at all of them - the reference's broadcasting)
 *   up_c8 : optional bf16 c8 [nbox][12][D/2][H/2][W/2][8]; out_c8 bf16 c8 [nbox][12][D][H][W][8] */
int pmnet_box_combine_c96(const void* s_c8, const float* u, const float* pvec, const int32_t* pvox,
                          const float* scale, const float* bias, int32_t relu, const void* up_c8, void* out_c8,
                          int32_t nbox, int32_t D, int32_t H, int32_t W, void* stream);

int pmnet_box_combine_c96_split(const void* s_c8, const void* s_lo_c8, const float* u, const float* pvec,
                                const int32_t* pvox, const float* scale, const float* bias, int32_t relu,
                                const void* up_c8, const void* up_lo_c8, void* out_c8, void* out_lo_c8, int32_t nbox,
                                int32_t D, int32_t H, int32_t W, void* stream);

/* Density-map post-processing (module.py:277-288): sigmoid(logits) masked to box & protein & cavity, 5^3 Gaussian
 * smoothing with zero padding (utils/smoothing.py), masked again, values < threshold set to 0. The spherical box area of
 * token (x, y, z, type) (data/token_inference.py:118-146) is evaluated in place.
 *   logits fp32 [n][size^3]; tokens int32 [n][4]; masks uint8 [size^3]; taps3 HOST fp32 [3] = 1-D Gaussian taps at
 *   distance 2, 1, 0 (normalised); out fp32 [n][size^3] */
int pmnet_density_post(const float* logits, const int32_t* tokens, const uint8_t* protein_mask,
                       const uint8_t* cavity_mask, const float* taps3, float threshold, float* out, int32_t n,
                       int32_t size, void* stream);

/* Window attention of a 3-D Swin-V2 block (swinv2.py:114-158, 272-298): cyclic shift (first two spatial axes only,
 * like the reference), 4^3 window partition, cosine attention with per-head logit scale, continuous position bias,
 * shift mask, softmax, P.V, window reverse and un-shift in one kernel.
 *   qkv  : [B * res^3][3 * heads * 32] token-major (q | k | v), fp32 or bf16; out : [B * res^3][heads * 32], same type
 *   logit_scale : fp32 [heads] = exp(min(logit_scale, log 100)); rel_bias : fp32 [heads][64][64] =
 *   16 * sigmoid(cpb_mlp(table))[index]; attn_mask : fp32 [(res/4)^3][64][64] or NULL (shift = 0) */
int pmnet_window_attention(const void* qkv, void* out, const float* logit_scale, const float* rel_bias,
                           const float* attn_mask, int32_t B, int32_t res, int32_t shift, int32_t heads,
                           int32_t is_bf16, void* stream);

/* y = shortcut + LayerNorm(h) * gamma + beta (res-post-norm, swinv2.py:300-303); shortcut may be NULL (plain
 * LayerNorm) and may alias y. h fp32 or bf16 [rows][C]; y, shortcut fp32; C in {96, 192, 384, 768}. */
int pmnet_ln_residual(const float* shortcut, const void* h, int32_t h_is_bf16, const float* gamma, const float* beta,
                      float* y, int64_t rows, int32_t C, float eps, void* stream);

/* Same two kernels emitting the two-term bf16 split of their result (hi = bf16(y), lo = bf16(y - hi)), the operand
 * pair of a split-precision pmnet_gemm_bf16: the attention result [B * res^3][heads * 32] from fp32 qkv, and the
 * LayerNorm output (fp32 y as before, plus y_hi and optionally y_lo, [rows][C]). */
int pmnet_window_attention_split(const float* qkv, void* out_hi, void* out_lo, const float* logit_scale,
                                 const float* rel_bias, const float* attn_mask, int32_t B, int32_t res, int32_t shift,
                                 int32_t heads, void* stream);
int pmnet_ln_residual_split(const float* shortcut, const void* h, int32_t h_is_bf16, const float* gamma,
                            const float* beta, float* y, void* y_hi, void* y_lo, int64_t rows, int32_t C, float eps,
                            void* stream);

/* Linear layer y[M][N] = act(a[M][K] . w[N][K]^T + bias) on tcgen05 (csrc/gemm.cu): nn.Linear of the Swin blocks
 * (swinv2.py:114-158, swin.py:19-44), PatchMerging.reduction (swinv2.py:346-363), PatchEmbed.proj and the 4^3 FPN
 * convolution in im2col form, the 1x1 laterals of the small FPN levels, the token- and mask-head MLPs.
 *   a_hi, w_hi : bf16 row-major, K contiguous, 16-byte aligned; K % 8 == 0; N <= 96 or N % 96 == 0, N % 32 == 0
 *   a_lo, w_lo : both NULL = one tensor-core pass over bf16 operands; both given = the two-term split operands
 *                (value = hi + lo) and three passes hi.hi + lo.hi + hi.lo into one fp32 accumulator
 *   bias       : fp32 [N] or NULL;  act: 0 none, 1 GELU (erf), 2 ReLU, 3 SiLU
 *   out_f32    : fp32 [M][N] or NULL; out_hi / out_lo : bf16 [M][N] or NULL, the split of the result */
int pmnet_gemm_bf16(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                    float* out_f32, void* out_hi, void* out_lo, int64_t M, int32_t N, int32_t K, int32_t act,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMNET_B200_H */
