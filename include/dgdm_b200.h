/* dgdm_b200.h — C ABI of the B200-native DGDM forward hot path.
 *
 * The reference (danieleschmidt/dgdm-histopath-lab) is pure Python: it has no FFI/operator layer to mirror, so
 * this header IS the boundary a maintainer would bind (ctypes stubs in INTEGRATION.md).  Each entry point
 * cites the reference code it replaces (paths relative to the reference's dgdm_histopath/).
 *
 * Conventions: every pointer is a DEVICE pointer unless named host_*; fp32 row-major contiguous rows with an
 * explicit leading dimension where given; int32 CSR; the caller owns every buffer (no allocation, no host
 * sync inside unless stated); `stream` is a cudaStream_t passed as void*.  Return 0 on success, <0 on error.
 */
#ifndef DGDM_B200_H_
#define DGDM_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGDM_OK 0
#define DGDM_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define DGDM_ERR_LAUNCH  (-2)   /* CUDA launch or runtime error */
#define DGDM_ERR_NODEV   (-3)   /* no sm_100 device */

#define DGDM_ACT_NONE 0
#define DGDM_ACT_GELU 1   /* exact erf GELU (nn.GELU default) */
#define DGDM_ACT_RELU 2
#define DGDM_ACT_SILU 3
#define DGDM_ACT_TANH 4
#define DGDM_ACT_ELU  5

/* ABI / build info. */
int dgdm_abi_version(void);
/* Returns 0 if device `dev` is compute capability 10.x, DGDM_ERR_NODEV otherwise. */
int dgdm_device_check(int dev);

/* ------------------------------------------------------------------------------------------------ GEMM
 * Node-level linear layers (every nn.Linear at node scale: models/encoders.py:73-89,215;
 * core/graph_layers.py:45,49,138-146,272-274,398; core/attention.py:44-49; core/diffusion.py:87-104) with the
 * row-wise epilogue that follows them in the reference fused in:
 *     v = A @ W^T (+bias) (+r_pre) ; v = act_pre(v) ; [v = LN1(v)] ; [v = LN2(v)] ; v = act_post(v) (+r_post)
 *     C = v ;  optionally dot_out[m] = dot_act(sum_n v[m,n]*dot_w[n] + dot_b)
 * W is supplied pre-split into two fp16 planes (hi = fp16(w), lo = fp16(w - hi)) stored as pre-swizzled K-block images
 * made by dgdm_split_weight_f16: img[kb][plane][N rows x 128 B] with Kpad = round_up(K, 64) zero padded, every
 * [N x 64] tile already in the tcgen05 K-major SWIZZLE_128B canonical layout so that one cp.async.bulk lands it in
 * shared memory. W_hi = image base, W_lo = W_hi + N*128 (the lo plane of K-block 0), ldw = Kpad; the buffer holds
 * N*Kpad*4 bytes. A is fp32 and is split the same way inside the kernel; the product is hi*hi + lo*hi + hi*lo on
 * tcgen05 (kind::f16, fp32 accumulate in TMEM), ~2^-21 relative — fp32-class, not TF32-class.
 * Limits: N % 16 == 0, N >= 16 (N > 512 runs as column groups of 512 + dgdm_rowwise_epilogue and needs C),
 * K % 4 == 0, lda % 4 == 0, pointers 16-byte aligned.
 */
typedef struct {
  const float* A; int lda; int M; int K;
  const void* W_hi; const void* W_lo; int ldw;   /* pre-swizzled images (see above), ldw = Kpad */
  int N;
  float* C; int ldc;                              /* may be NULL when only dot_out is wanted */
  const float* bias;                              /* [N] or NULL */
  const float* r_pre;  int ld_r_pre;              /* [M,N] addend before act_pre/LN, or NULL */
  const float* r_post; int ld_r_post;             /* [M,N] addend after act_post, or NULL */
  int act_pre;
  const float* ln1_g; const float* ln1_b;         /* LayerNorm 1 (NULL = skip) */
  const float* ln2_g; const float* ln2_b;         /* LayerNorm 2 (NULL = skip) */
  float ln_eps;
  int act_post;
  const float* dot_w; float dot_b; int dot_act; float* dot_out;   /* row-dot epilogue (NULL = skip) */
  float a_scale;                                  /* power of two applied to A before the fp16 split (1.0f default) */
  const int* row_bias_index; const float* row_bias; int ld_row_bias; /* optional per-row bias: v += row_bias[row_bias_index[m], :] */
  /* Fused halo exchange for node-partitioned graphs (GEMM -> all-gather in ONE kernel over NVLink peer memory): every
   * output tile is also stored at the same (row, column, ldc) position of n_peers peer buffers — C_peers[i] is the address
   * in THIS process of the peer's copy of C (symmetric memory, P2P mapped). peers_multicast = 1: C_peers[0] is an NVLS
   * multicast address that covers every rank including this one (one multimem.st per tile, the switch replicates it) and
   * the plain store to C is skipped. n_peers = 0 (default): ordinary single-GPU store. */
  float* C_peers[8]; int n_peers; int peers_multicast;
} dgdm_gemm_params;

/* tcgen05 + TMEM path (the product path). */
int dgdm_gemm_f16x2(const dgdm_gemm_params* p, void* stream);
/* Plain fp32 SIMT GEMM with the same epilogue: bring-up / on-device comparator only. W32 is fp32 [N,K]. */
int dgdm_gemm_simt(const dgdm_gemm_params* p, const float* W32, void* stream);
/* Split fp32 W[N,K] (N % 8 == 0) into the fp16 hi/lo K-block images described above; W_lo must be W_hi + N*128. */
int dgdm_split_weight_f16(const float* W, int N, int K, void* W_hi, void* W_lo, int Kpad, void* stream);

/* Stand-alone row-wise epilogue for widths that do not fit TMEM (N > 512, DGDM-Large 768-d layers):
 *   v = X (+bias) (+r_pre); act_pre; LN1; LN2; act_post; (+r_post) -> Y.  Same field meaning as above. */
int dgdm_rowwise_epilogue(const float* X, int ldx, int M, int N, const dgdm_gemm_params* ep, void* stream);

/* ------------------------------------------------------------------------------------- graph structure
 * CSR by destination with the original edge order kept inside each row, GCN normalisation
 * (core/graph_layers.py:76-84: add_self_loops -> degree(col) -> deg^-1/2[row]*deg^-1/2[col]).
 * edge_index is the reference's int64 [2,E] (row 0 = source, row 1 = target).
 * Outputs: rowptr[N+1], col[E] (source of each in-edge), eid[E] (original edge id), norm[E],
 * dinv[N] (deg^-1/2 incl. the self loop).  work: int32 scratch of dgdm_csr_build_work_ints(N,E) ints.
 */
size_t dgdm_csr_build_work_ints(int N, int E);
int dgdm_csr_build(const int64_t* edge_index, int E, int N, int* rowptr, int* col, int* eid, float* norm,
                   float* dinv, int* work, void* stream);

/* GraphConvolution.propagate + bias + activation (core/graph_layers.py:92-110, 233-239) for output rows i in [0, n_rows):
 *   out[i,:] = act( sum_{e in row i, edge order} norm[e]*H[col[e],:] + dinv[i]^2*H[h_row0+i,:] (+addend[i,:]) (+bias) )
 * Rows i >= n_csr_rows (decoder levels of GraphUNet, D8) have only the self loop with weight 1 and no addend
 * (addend holds n_csr_rows rows).
 * Node-partitioned graphs: a rank that owns rows [a,b) passes rowptr+a, dinv+a, h_row0 = a and the all-gathered H
 * (col holds global indices); single GPU: h_row0 = 0.
 * C % 4 == 0, C <= 1024. Deterministic: fixed in-row order, self loop last, no atomics. */
int dgdm_spmm_gcn(const int* rowptr, const int* col, const float* norm, const float* dinv, int n_csr_rows,
                  int n_rows, int h_row0, const float* H, int ldh, int C, const float* addend, int ld_addend,
                  const float* bias, int act, float* out, int ldo, void* stream);

/* sum_{e in row i} norm[e]*edge_attr[eid[e],:]  -> agg[N, D]  (identity (i) of SURVEY §8: the edge_lin term of
 * core/graph_layers.py:103-105 aggregated once per graph; self-loop rows are zero under P2). */
int dgdm_edge_attr_aggregate(const int* rowptr, const int* eid, const float* norm, int N,
                             const float* edge_attr, int ld_attr, int D, float* agg, void* stream);

/* AdaptiveGraphPooling (core/graph_layers.py:285-329): exact top-k (k = max(1,int(ratio*N)) computed by the
 * caller) over tanh scores, ascending `perm`, node_map, pooled features x[perm]*score[perm] with `act`
 * applied (GraphUNet applies self.act right after pooling, :432/:419), and the pooled CSR (order kept).
 * Ties at the k-th value are broken towards the lower node index.
 * work: dgdm_topk_pool_work_ints(N) ints. counts_out (device, 2 ints): [k, E_pooled]. */
size_t dgdm_topk_pool_work_ints(int N);
int dgdm_topk_select(const float* scores, int N, int k, int* perm, int* node_map, int* work, void* stream);
int dgdm_pool_gather(const float* X, int ldx, int C, const float* scores, const int* perm, int k, int act,
                     float* out, int ldo, void* stream);
/* eid / p_eid (both NULL, or both given): the pooled entries keep the LEVEL-0 edge id of the edge they came from, so that
 * `edge_attr[edge_mask]` (:327) never has to be materialised — dgdm_edge_attr_aggregate on the pooled CSR indexes the
 * original edge_attr. */
int dgdm_csr_pool(const int* rowptr, const int* col, const int* eid, int N, const int* perm, const int* node_map, int k,
                  int* p_rowptr, int* p_col, int* p_eid, float* p_norm, float* p_dinv, int* work, void* stream);
/* GraphUNet un-pool + skip + act (core/graph_layers.py:441-452): out[i] = act(skip[i] + (node_map[i]>=0 ? x[node_map[i]] : 0)) */
int dgdm_unpool_add_act(const float* skip, int ld_skip, const float* x, int ldx, const int* node_map, int N,
                        int C, int act, float* out, int ldo, void* stream);

/* DynamicGraphLayer.compute_dynamic_edges (core/graph_layers.py:160-205): per-edge per-head
 * q[src].(k[dst]+ek[e])/sqrt(hd) then PyG softmax over the in-edges of each dst (eps 1e-16).
 * qkv [N,3*hid] (q|k|v), ek [E,hid] = edge_to_key(edge_attr) in ORIGINAL edge order. out [E,heads] original order. */
int dgdm_segment_softmax_scores(const int* rowptr, const int* col, const int* eid, int N, const float* qkv,
                                int ld_qkv, const float* ek, int ld_ek, int hid, int heads, float* out,
                                void* stream);

/* ------------------------------------------------------------------------------------------- attention
 * SpatialAttention (core/attention.py:225-327) per graph segment [seg_ptr[g], seg_ptr[g+1]) of a node-sorted batch:
 *   1. dgdm_spatial_pe_add     x + sinusoidal PE(min-max normalised pos)        (:225-259)
 *   2. (QKV projection: dgdm_gemm_f16x2 with the concatenated [3D, D] weight)
 *   3. dgdm_spatial_attention  softmax(q k^T/sqrt(hd) - |p_i - p_j|/T) v, flash-style, never forming N x N (:261-283,:135-157)
 *   4. (out_proj + residual + LayerNorm: dgdm_gemm_f16x2 epilogue)              (:163,:325)
 */
int dgdm_segment_minmax(const float* pos, const int* seg_ptr, int num_seg, float* minmax /*[num_seg,2]*/, void* stream);
int dgdm_spatial_pe_add(const float* X, int ldx, const float* pos, const int* seg_ptr, int num_seg,
                        const float* minmax, int N, int D, float* out, int ldo, void* stream);
/* qkv: [N, 3*D] fp32 (q|k|v, biases included). out: [N, D] fp32 (heads concatenated, before out_proj).
 * kv_pos/kv_qkv allow the keys/values to come from a different (all-gathered) node set for node-partitioned
 * graphs: queries rows [0,Nq) use key segment [kseg_ptr[g], kseg_ptr[g+1]) of the kv arrays, with g = the
 * query's segment in qseg_ptr.  work: dgdm_spatial_attention_work_bytes(...) bytes. max_seg_keys: key count of the largest
 * segment (0 = unknown, Nkv is assumed); it sizes the key splits that cap the tensor-core accumulation chain.
 * precise: 1 = product path, 3-term split-fp16 products (~fp32 accuracy): the tcgen05/TMEM kernel when D == 128 and heads == 8
 *             (DGDM-Base), else the mma.sync kernel; 2 = force the 3-term mma.sync kernel (comparison baseline);
 *          0 = single-pass fp16 operands on mma.sync (~1e-3, comparison only); 3 = force the exact fp32 CUDA-core kernel.
 * head_dim 16 runs on the tensor cores; head_dim 4/8/32/64 (DGDM-Large: 16 heads x 8) run on the exact fp32 CUDA-core kernel. */
size_t dgdm_spatial_attention_work_bytes(int Nq, int Nkv, int D, int heads, int num_seg, int max_seg_keys);
int dgdm_spatial_attention(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                           const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int Nkv,
                           int num_seg, int max_seg_keys, int D, int heads, float temperature, int precise,
                           float* out, int ldo, void* work, size_t work_bytes, void* stream);
/* Same, additionally writing the log-sum-exp of every row's scores per head (natural-log units, [Nq, heads]) for the backward pass.
 * head_dim 16 only (lse_out must be NULL for the CUDA-core path). */
int dgdm_spatial_attention_lse(const float* q_qkv, int ld_q, const float* q_pos, const int* qseg_ptr, int Nq,
                               const float* kv_qkv, int ld_kv, const float* kv_pos, const int* kseg_ptr, int Nkv,
                               int num_seg, int max_seg_keys, int D, int heads, float temperature, int precise,
                               float* out, int ldo, float* lse_out, void* work, size_t work_bytes, void* stream);
/* Head-averaged attention weights [n,n] of ONE graph segment (return_attention=True; O(n^2) output). */
int dgdm_spatial_attention_weights(const float* qkv, int ld, const float* pos, int n, int D, int heads,
                                   float temperature, float* weights /*[n,n]*/, void* stream);

/* MultiHeadAttention.forward core (core/attention.py:113-172) for the options the flash kernels do not take: additive
 * (`attn_scores += attn_mask`, :142) or boolean (`masked_fill_(-inf)`, :140) masks with torch-broadcast strides in ELEMENTS over
 * (batch, head, query, key) — stride 0 = broadcast —, key_padding_mask [B,Lk] (:145-150), cross attention (Lq != Lk), a value
 * tensor distinct from the key, and weights per head [B*H,Lq,Lk] (average_attn_weights=False, :172) or head-averaged
 * [B,Lq,Lk] (:170). q [B*Lq, D], k / v [B*Lk, D] are the PROJECTED rows. Optional spatial bias -|p_i - p_j| / temperature
 * (SpatialAttention.compute_spatial_bias, :261-283) from pos_q [P,Lq,2] / pos_k [P,Lk,2]; pos_by_head = 1 indexes P by the
 * head instead of the batch element: that is what the reference's `[B,H,L,L] += bias[B,L,L]` broadcast does when B == H > 1
 * (a quirk reproduced as is; for other B > 1 the reference raises). A fully masked row yields NaN like torch's softmax.
 * Exact fp32 on CUDA cores, O(Lq*Lk*D) — the public-API path, not the model's forward. */
int dgdm_mha_dense(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, int B, int Lq, int Lk,
                   int D, int heads, const float* mask, const unsigned char* bool_mask, long long ms_b, long long ms_h,
                   long long ms_q, long long ms_k, const unsigned char* key_padding_mask, const float* pos_q,
                   const float* pos_k, int pos_by_head, float temperature, float* out, int ldo, float* weights,
                   int per_head, void* stream);

/* Backward of dgdm_spatial_attention for the training step (models/dgdm_model.py:447-480 runs this layer under
 * autograd): given qkv [N,3D], positions, the forward output O and its gradient dO, returns dqkv [N,3D] (dq | dk | dv).
 * Self-attention inside each graph segment (seg_of[i] = segment of row i, seg_ptr = segment bounds). lse / delta: [N, heads]
 * (log-sum-exp of the scores and O.dO per row and head). have_lse = 1: lse holds the forward's statistics
 * (dgdm_spatial_attention_lse) and only delta is formed; 0: a recompute pass produces both.
 * head_dim 16 only (DGDM-Base). Exact fp32, deterministic (no atomics). */
int dgdm_spatial_attention_backward(const float* qkv, int ld, const float* pos, const int* seg_of, const int* seg_ptr,
                                    int N, int D, int heads, float temperature, const float* O, int ldo,
                                    const float* dO, int lddo, int have_lse, float* lse, float* delta, float* dqkv, int ldd,
                                    void* stream);

/* Tensor-core version of the same three passes (warp-level mma.sync m16n8k16, single fp16 operand planes, fp32 accumulation; the
 * default of the training path). go_scale: device scalar, a power of two that brings max|dO| to ~8 (gradients of ~1e-6 would
 * otherwise fall below the normal fp16 range); the outputs are un-scaled. Gradients agree with the fp32 path to ~2e-3 (stated
 * tensor-core tolerance), deterministic. */
int dgdm_spatial_attention_backward_tc(const float* qkv, int ld, const float* pos, const int* seg_of, const int* seg_ptr,
                                       int N, int D, int heads, float temperature, const float* O, int ldo,
                                       const float* dO, int lddo, const float* go_scale, int have_lse, float* lse, float* delta,
                                       float* dqkv, int ldd, void* stream);

/* GlobalAttentionPool (models/dgdm_model.py:588-615): one query (q_tok = q_proj(global_token), [D]) per graph
 * over that graph's keys/values kv [N, 2*D] (k|v). out [num_seg, D] before out_proj.
 * work: dgdm_attention_pool_work_floats(N,D,heads,num_seg) floats. Empty segments give zeros. */
size_t dgdm_attention_pool_work_floats(int N, int D, int heads, int num_seg);
int dgdm_attention_pool(const float* q_tok, const float* kv, int ld_kv, const int* seg_ptr, int num_seg, int N,
                        int D, int heads, float* out, float* work, void* stream);

/* ------------------------------------------------------------------------------------------- diffusion
 * DiffusionLayer (core/diffusion.py:123-172): x_t = sqrt(ac[t_g])*x0 + sqrt(1-ac[t_g])*noise per graph segment,
 * GroupNorm(G)+SiLU row-wise on [N,C], and the fused DDPM sample update (:257-273). */
int dgdm_diffusion_add_noise(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac,
                             const int* seg_ptr, const int* seg_t, int num_seg, int N, int C, float* out, void* stream);
int dgdm_groupnorm_silu(const float* X, int ldx, int M, int C, int groups, const float* gamma, const float* beta,
                        float eps, float* out, int ldo, void* stream);
int dgdm_ddpm_update(const float* x, const float* eps_pred, const float* noise /*NULL on the last step*/,
                     float sqrt_1mac, float inv_sqrt_ac, float sqrt_alpha, float sqrt_var, size_t n,
                     float* out, void* stream);

/* ------------------------------------------------------------------------------------------ validation
 * DGDMModel._validate_forward_inputs/_validate_input_features (models/dgdm_model.py:646-710) as ONE pass +
 * one 32-byte readback: flags[0]=nan count>0, [1]=inf, [2]=edge index > N-1, [3]=edge index < 0;
 * stats[0]=abs max, [1]=sum, [2]=sum of squares (for the std<1e-6 warning). out: 4 ints + 4 floats. */
int dgdm_validate_inputs(const float* x, size_t n_x, const int64_t* edge_index, size_t n_ei, int N,
                         int* flags4, float* stats4, void* stream);

/* ---------------------------------------------------------------------------------------- edge builder
 * TissueGraphBuilder._create_morphological_edges (preprocessing/tissue_graph_builder.py:322-352): for query rows
 * [q0,q0+nq) against all n rows of F [n,d] fp32: the k nearest rows by Euclidean distance — sklearn brute-force
 * semantics for float32 input: ranked on fp64 squared distances |x|^2 - 2x.y + |y|^2 clipped at 0, the query row
 * itself included (the reference asks for k+1 neighbours and skips column 0, so callers pass morphological_k+1) —
 * and the cosine similarity of each kept pair (sklearn cosine_similarity: rows normalised in fp32, fp32 result).
 * Ties are broken towards the lower index. nbr_idx [nq,k] int32 ascending distance, nbr_cos [nq,k] fp32,
 * nbr_dist [nq,k] fp64 (sqrt of the clipped squared distance). */
size_t dgdm_knn_cosine_work_bytes(int nq, int n, int d, int k);
int dgdm_knn_cosine(const float* F, int n, int d, int q0, int nq, int k, int* nbr_idx, float* nbr_cos,
                    double* nbr_dist, void* work, size_t work_bytes, void* stream);
/* Tensor-core candidate pass for large n (BASELINE.json configs[2]: 1M x 768 rows, query rows sharded across GPUs by q0/nq): the
 * N^2 x d Gram work runs on tcgen05 through the split-fp16 GEMM, 512 key columns at a time; per query row the 32 nearest
 * candidates by approximate distance are kept, re-ranked with the exact fp64 arithmetic above, and a row is flagged
 * (row_flags[i] = 1) unless it is PROVED that no non-candidate can belong to its top-k — the caller recomputes flagged rows
 * with dgdm_knn_cosine, so results are always exactly the reference's. F: [n_pad, d] (n_pad = round_up(n,16), pad rows zero,
 * d % 4 == 0), W_img: dgdm_split_weight_f16 images of the same matrix, sq: dgdm_row_sqnorm_f64 of it. k <= 32. */
int dgdm_row_sqnorm_f64(const float* F, int n, int d, double* sq, void* stream);
size_t dgdm_knn_cosine_tc_work_bytes(int nq);
int dgdm_knn_cosine_tc(const float* F, const void* W_img, const double* sq, int n, int n_pad, int d, int q0, int nq, int k,
                       int* nbr_idx, float* nbr_cos, double* nbr_dist, int* row_flags, void* work, size_t work_bytes, void* stream);
/* ContrastiveLoss._batch_contrastive_loss (training/losses.py:123-176), never forming the N x N matrix:
 * dgdm_row_normalize = F.normalize(embeddings, dim=1) (:131); the similarity chunks come from dgdm_gemm_f16x2_strided (rows of the
 * normalised matrix against a 512-row key range of its own fp16 hi/lo image); dgdm_contrastive_accumulate folds one chunk
 * S [M, nk] (rows row0.., keys k0..) into sum_all[i] += sum_j exp((S_ij - 1)/T), sum_pos[i] (same graph id, j != i) and
 * cnt_pos[i]. The shift by 1/T cancels in the ratio; the caller adds it back inside the logarithms. batch: int64 graph id per
 * node or NULL (one graph). Deterministic (fixed lane -> key mapping, one update per chunk). */
int dgdm_row_normalize(const float* X, int ldx, int M, int C, float eps, float* out, int ldo, void* stream);
int dgdm_contrastive_accumulate(const float* S, int lds, int M, int nk, int row0, int k0, const int64_t* batch,
                                float temperature, float* sum_all, float* sum_pos, int* cnt_pos, void* stream);

/* dgdm_gemm_f16x2 on a row sub-range of a larger weight image: W_hi points at the first row of the range in the hi plane of
 * K-block 0, w_plane_stride_bytes = (rows of the whole image) * 128; p->N <= 512; plain product (no epilogue fields needed). */
int dgdm_gemm_f16x2_strided(const dgdm_gemm_params* p, long long w_plane_stride_bytes, void* stream);
/* _create_spatial_edges (:290-320): exact k nearest neighbours on 2-D fp64 coords (query row included, callers pass
 * spatial_k+1), ascending distance, ties towards the lower index. mode 0: brute force over fp64 distance rows (small n);
 * mode 1: uniform-grid ring search, O(N), identical results; work = dgdm_knn2d_grid_work_bytes(n) bytes (mode 1 only). */
size_t dgdm_knn2d_grid_work_bytes(int n);
int dgdm_knn2d(const double* coords, int n, int q0, int nq, int k, int* nbr_idx, double* nbr_dist, int mode, void* work, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGDM_B200_H_ */
