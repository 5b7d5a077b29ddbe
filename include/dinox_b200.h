/* dinox_b200 - C ABI of the B200-native (sm_100a) DINO-X loss head.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 * launches asynchronously on the given stream, never allocates persistent device memory except
 * through an explicit plan/workspace object, and returns DINOX_OK or a negative DINOX_E_* code
 * (message via dinox_last_error_string()).  All device pointers must be 16-byte aligned and the
 * tensors contiguous unless a leading dimension is given.  There is no CPU fallback.
 *
 * Each function cites the reference code it replaces (paths relative to timlawrenz/DINO-X).
 * Python binding: dinox_b200/_ext.py (ctypes).  Reference-side stub: INTEGRATION.md.
 */
#ifndef DINOX_B200_H_
#define DINOX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DINOX_API
#define DINOX_API __attribute__((visibility("default")))
#endif

/* cudaStream_t without pulling in cuda_runtime.h for pure-C consumers */
typedef struct CUstream_st* dinox_stream_t;

#define DINOX_OK 0
#define DINOX_E_BADARG (-1)      /* bad shape / null pointer / unsupported size */
#define DINOX_E_ALIGN (-2)       /* pointer or leading dimension not 16-byte aligned */
#define DINOX_E_ARCH (-3)        /* device is not sm_100 (B200); there is no fallback */
#define DINOX_E_CUDA (-4)        /* CUDA runtime/driver error, see dinox_last_error_string */
#define DINOX_E_UNSUPPORTED (-5) /* valid request that this build does not implement */

/* element types of tensors that may arrive in several precisions */
#define DINOX_F32 0
#define DINOX_BF16 1
#define DINOX_F16 2

DINOX_API int dinox_version(void);
DINOX_API const char* dinox_last_error_string(void);
/* DINOX_OK iff the current device is compute capability 10.x */
DINOX_API int dinox_device_check(void);
/* number of kernels this library has launched on the calling thread since the last reset */
DINOX_API int64_t dinox_launch_count(void);
DINOX_API void dinox_launch_count_reset(void);
/* Launch trace (diagnostics, tools/trace_step.py): between trace_begin and trace_end every launch of this library is
 * followed, on its stream, by a one-thread kernel that writes %globaltimer (ns) into device_slots[i], i = launch
 * order; captured into a CUDA graph the stamps are rewritten by every replay.  trace_end returns the number of
 * launches recorded; trace_name / trace_stream describe slot i. */
DINOX_API int dinox_trace_begin(uint64_t* device_slots, int capacity);
DINOX_API int dinox_trace_end(void);
DINOX_API const char* dinox_trace_name(int i);
DINOX_API uint64_t dinox_trace_stream(int i);

/* ------------------------------------------------------------------------------------------
 * a8  EMA teacher update.  Replaces the per-tensor loop
 *     `p_t.data.mul_(m).add_(p_s.data, alpha=1-m)` (scripts/phase5_big_run.py:1798-1802,
 *     scripts/phase3_micro_run.py:152-155) by ONE multi-tensor launch.
 *     A plan uploads a chunk table (pointer pairs, 128-bit vectorised) once; parameters keep
 *     their addresses across steps.  All tensors fp32.
 * ------------------------------------------------------------------------------------------ */
typedef struct dinox_ema_plan dinox_ema_plan;
DINOX_API int dinox_ema_plan_create(const void* const* student, void* const* teacher,
                                    const int64_t* numel, int n_tensors, dinox_ema_plan** out);
DINOX_API int dinox_ema_plan_destroy(dinox_ema_plan* plan);
DINOX_API int64_t dinox_ema_plan_numel(const dinox_ema_plan* plan);
/* p_t <- fl(p_t*m) (+) one_minus_m*p_s  (fma), in place */
DINOX_API int dinox_ema_apply(const dinox_ema_plan* plan, float m, float one_minus_m,
                              dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row/column statistics on MATERIALISED logits (the DINOLoss.forward(student_out, teacher_out)
 * drop-in path, scripts/phase5_big_run.py:692-720).  u[i,k] = x[i,k]*inv_tau - colbias[k]
 * (colbias may be NULL).  ld = row stride in elements.
 * ------------------------------------------------------------------------------------------ */
/* a2/a3: lse[i] = ln sum_k exp(u[i,k] - rowshift) ... returns natural-log LSE per row; optional
 * entropy[i] = -sum_k p log p of softmax(u[i,:]) (a10, scripts/phase5_big_run.py:1843-1853). */
DINOX_API int dinox_rows_lse(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld,
                             float inv_tau, const float* colbias, float* lse, float* entropy,
                             dinox_stream_t stream);
/* E2: per-column LSE over rows of (u[i,k] - rowbias[i]); rowbias may be NULL. out[k] natural log */
DINOX_API int dinox_cols_lse(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld,
                             float inv_tau, const float* rowbias, float* out,
                             dinox_stream_t stream);
/* combine per-rank LSE vectors gathered as (world, K): out[k] = ln sum_r exp(g[r,k]) + add */
DINOX_API int dinox_lse_combine(const float* gathered, int world, int64_t K, float add, float* out,
                                dinox_stream_t stream);
/* a5: column sums of x (rows, K) -> out (K) fp32 (torch.mean(teacher_output, dim=0) numerator,
 * scripts/phase5_big_run.py:688) */
DINOX_API int dinox_cols_sum(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld,
                             float* out, dinox_stream_t stream);
/* out[k] (+)= scale*(*scale_dev) * column sum (accumulate != 0: +=): the bias gradients db1 / db2 summed from their
 * partial rows straight into `.grad` (autograd's Linear backward `grad_output.sum(0)`, zoo/arch.py:252-256) */
/* column sums of one or two row ranges [seg_begin[i], seg_end[i]) of x in ONE launch -> out (nseg, K); optionally
 * counts_out[i] = counts_in[i] rides along (the [activation sums | row counts] all-reduce payload of the centre
 * update, scripts/phase5_big_run.py:686-690).  64-row chunk sums are combined in chunk order by the block that
 * finishes last (ticket counter), so the result is run-to-run identical.  `workspace`
 * (dinox_segment_cols_sum_workspace_bytes) must be ZERO before its first use; every launch leaves its ticket area
 * zero again.  Not to be shared by launches that may run concurrently. */
DINOX_API int64_t dinox_segment_cols_sum_chunks(int64_t rows0, int64_t rows1);
DINOX_API size_t dinox_segment_cols_sum_workspace_bytes(int64_t rows0, int64_t rows1, int64_t K);
DINOX_API int dinox_segment_cols_sum(const void* x, int dtype, int64_t K, int64_t ld, int nseg,
                                     const int64_t* seg_begin, const int64_t* seg_end, float* out,
                                     const float* counts_in, float* counts_out, void* workspace,
                                     dinox_stream_t stream);
DINOX_API int dinox_cols_sum_axpy(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld, float scale,
                                  const float* scale_dev, float* out, int accumulate, dinox_stream_t stream);
/* partial[c, :] = column sums of rows [c*chunk, min((c+1)*chunk, rows)): first phase of a tall-skinny
 * column sum (finish with dinox_cols_sum on the (ceil(rows/chunk), K) fp32 partial matrix) */
DINOX_API int dinox_cols_sum_chunked(const void* x, int dtype, int64_t rows, int64_t K, int64_t ld, int64_t chunk,
                                     float* partial, dinox_stream_t stream);
/* a5: center <- center*m + (colsum*inv_rows)*(1-m)   (scripts/phase5_big_run.py:689) */
DINOX_API int dinox_center_ema(float* center, const float* colsum, float inv_rows, float m,
                               int64_t K, dinox_stream_t stream);
/* vector helper: out[k] = a[k]*alpha + beta (colbias = center*inv_tau etc.) */
DINOX_API int dinox_axpb(const float* a, float alpha, float beta, float* out, int64_t n,
                         dinox_stream_t stream);
/* out[k] = x[k]*alpha*(*alpha_dev) + y[k]*beta   (alpha_dev, y optional; e.g. col offsets
 * (b2 - center)*inv_tau*log2e, or bias.grad += upstream * db) */
DINOX_API int dinox_axpby(const float* x, float alpha, const float* alpha_dev, const float* y, float beta,
                          float* out, int64_t n, dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a4 / E1 / E3: cross-entropy between teacher and student rows on materialised logits.
 * Rows are organised in `groups` (multi-crop: group = image b, student row v*groups+b, teacher
 * row iq*groups+b; iBOT: group = masked token, V = Vg = 1).  Pairs (iq, v) with v != iq are
 * summed when exclude_same != 0, all pairs otherwise:
 *   loss = norm * sum_g w[g] * sum_pairs ( lse_s[v,g] - sum_k q[iq,g,k] * us[v,g,k] )
 *   q[i,k] = exp(t[i,k]*inv_tau_t - colbias_t[k] - rowbias_t[i]),  us = s*inv_tau_s
 * scripts/phase5_big_run.py:703-717 is (V=Vg=2, exclude_same=1, norm=1/(2B), w=NULL).
 * `loss_out` is one fp32 on device.  workspace: dinox_ce_workspace_bytes(groups, K) bytes.
 * ------------------------------------------------------------------------------------------ */
DINOX_API size_t dinox_ce_workspace_bytes(int64_t groups, int64_t K);
DINOX_API int dinox_ce_fwd(const void* student, int s_dtype, const void* teacher, int t_dtype,
                           int64_t groups, int V, int Vg, int64_t K, int64_t ld_s, int64_t ld_t,
                           float inv_tau_s, float inv_tau_t, const float* colbias_t,
                           const float* rowbias_t, const float* lse_s, const float* group_w,
                           float norm, int exclude_same, float* loss_out, void* workspace,
                           dinox_stream_t stream);
/* The same loss with a softmax-centred teacher, every logit read ONCE: softmax((t - c)/tau_t) and
 * log_softmax(s/tau_s) of scripts/phase5_big_run.py:703-706 are never formed - the row LSEs and the cross terms
 * sum_k exp(ut[k]) s[k] are accumulated online (running maximum + rescale) in the pass that reads the logits,
 * merged over K-splits in a fixed order.  By-products for dinox_ce_bwd: lse_s_out (V*groups) = ln sum_k exp(us),
 * rowbias_t_out (Vg*groups) = ln sum_k exp(t*inv_tau_t - colbias_t) (so q sums to 1 per row).  V <=
 * dinox_ce_onepass_max_views(); more views (or a Sinkhorn-Knopp teacher, whose row offsets are inputs) take
 * dinox_rows_lse + dinox_ce_fwd.  workspace: dinox_ce_onepass_workspace_bytes(groups, V, Vg, K) bytes. */
DINOX_API int dinox_ce_onepass_max_views(void);
DINOX_API size_t dinox_ce_onepass_workspace_bytes(int64_t groups, int V, int Vg, int64_t K);
DINOX_API int dinox_ce_fwd_onepass(const void* student, int s_dtype, const void* teacher, int t_dtype,
                                   int64_t groups, int V, int Vg, int64_t K, int64_t ld_s, int64_t ld_t,
                                   float inv_tau_s, float inv_tau_t, const float* colbias_t,
                                   const float* group_w, float norm, int exclude_same, float* loss_out,
                                   float* lse_s_out, float* rowbias_t_out, void* workspace,
                                   dinox_stream_t stream);
/* grad[v,g,k] = (*upstream) * norm * w[g] * inv_tau_s * ( n_q(v) * softmax(us)[k] - sum_{iq} q[iq,g,k] )
 * written in the student's dtype (autograd of log_softmax(s/tau), :706). upstream: device fp32. */
DINOX_API int dinox_ce_bwd(const void* student, int s_dtype, const void* teacher, int t_dtype,
                           int64_t groups, int V, int Vg, int64_t K, int64_t ld_s, int64_t ld_t,
                           float inv_tau_s, float inv_tau_t, const float* colbias_t,
                           const float* rowbias_t, const float* lse_s, const float* group_w,
                           float norm, int exclude_same, const float* upstream, void* grad,
                           int64_t ld_g, dinox_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * a1 and the dense contractions of its autograd: bf16 x bf16 -> fp32 tcgen05 GEMM.
 *   C[M,N] (+)= alpha * (*alpha_dev) * sum_k A[m,k] B[n,k] + bias_n[n]
 * A operand: a_mn_major=0 -> stored (M,K) row-major with ld lda; 1 -> stored (K,M) row-major.
 * B operand likewise with N.  out_dtype DINOX_F32 or DINOX_BF16.  m_fastest picks the tile
 * walk order (which operand stays L2/SMEM-hot).  Replaces nn.Linear forward/backward GEMMs of
 * the projection head (zoo/arch.py:252-256) that PyTorch dispatches to cuBLAS.
 * ------------------------------------------------------------------------------------------ */
DINOX_API int dinox_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K,
                              int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major,
                              int out_dtype, int accumulate, float alpha, const float* alpha_dev,
                              const float* bias_n, int m_fastest, dinox_stream_t stream);

/* The same (fp32 output) with a balanced schedule: when the tile count does not fill whole waves of the persistent
 * grid (dW2: 256 tile pairs on 74 CTA pairs = 3.46 waves; dH: 32 on 74), the tiles of the partial wave - or all
 * tiles when there are fewer tiles than clusters - are cut along K into parts that other clusters compute at the same
 * time.  The parts of a tile accumulate into C in a FIXED order (part j waits for part j-1 through the counters in
 * `flags`), so the result is run-to-run identical.  `flags` (dinox_gemm_bf16_balanced_workspace_bytes): zero before
 * the first launch, left zero by every launch; not to be shared by launches that may overlap. */
DINOX_API size_t dinox_gemm_bf16_balanced_workspace_bytes(int64_t M, int64_t N);
/* host-side views of that schedule (no GPU): the plan - tiles from `first` on are cut into `parts` K ranges, parts = 0:
 * whole tiles only - and the (item, cluster, m_tile, n_tile, kpart, kparts) list the clusters walk; returns the
 * number of items (out receives the first `cap`) or -1 */
DINOX_API int dinox_plan_ordered_split(int64_t tiles, int64_t clusters, int64_t kblocks, int* first, int* parts);
DINOX_API int64_t dinox_debug_walk_ordered(int num_m_super, int num_n_tiles, int m_fastest, int os_first,
                                           int os_parts, int clusters, int cl, int32_t* out, int64_t cap);
DINOX_API int dinox_gemm_bf16_balanced(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K,
                                       int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major,
                                       int accumulate, float alpha, const float* alpha_dev, const float* bias_n,
                                       int m_fastest, void* flags, dinox_stream_t stream);

/* Split-K flavour for GEMMs with few output tiles and a long reduction (dH = G . W2: 67 tiles,
 * K = 65536): split s reduces its share of K into the fp32 slab C_partials + s*split_stride;
 * the caller sums the slabs in fixed order (dinox_gather_sum_rows does), so the result is
 * deterministic.  dinox_gemm_splitk_plan() returns the split count that fills whole waves of
 * the persistent one-CTA-per-SM grid (1 = do not split). */
DINOX_API int dinox_gemm_bf16_splitk(const void* A, const void* B, float* C_partials, int64_t M, int64_t N,
                                     int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t split_stride,
                                     int splits, int a_mn_major, int b_mn_major, float alpha,
                                     const float* alpha_dev, int m_fastest, dinox_stream_t stream);
DINOX_API int dinox_gemm_splitk_plan(int64_t M, int64_t N, int64_t K);
/* GEMM fused with the reduce-scatter of its output rows over the data-parallel ranks (the gradient
 * reduction of dW2 that follows `loss.backward()` at scripts/phase5_big_run.py:1772-1796 in a DDP run of the loop):
 * C = alpha*(*alpha_dev) * A @ B^T is never stored locally; the tile rows [o*rows_per_owner, (o+1)*rows_per_owner)
 * are reduce-ADDED into shard_ptrs[o] ((rows_per_owner, ldc) fp32), which may be memory of a PEER GPU mapped into
 * this process (CUDA IPC / symmetric memory over NVLink).  `shard_ptrs` is a HOST array of `owners` (<= 8) device
 * pointers; rows_per_owner is a multiple of 128.  Every rank calls it with the same table: afterwards (once all
 * ranks' kernels have completed - order them with any collective) shard o holds the sum over ranks of its rows.
 * add_local (optional, (M, ld_local) fp32, scaled by add_scale) is added to every tile before it leaves: the
 * gradient accumulated LOCALLY over the earlier micro-steps of a window travels with the last micro-step's GEMM,
 * so the NVLink traffic is one gradient per window, not one per micro-step. */
DINOX_API int dinox_gemm_bf16_reduce_scatter(const void* A, const void* B, float* const* shard_ptrs, int owners,
                                             int64_t rows_per_owner, int64_t M, int64_t N, int64_t K, int64_t lda,
                                             int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, float alpha,
                                             const float* alpha_dev, const float* add_local, int64_t ld_local,
                                             float add_scale, dinox_stream_t stream);
/* diagnostics: number of co-resident clusters of `cluster_size` CTAs for the pass-1 kernel */
DINOX_API int dinox_debug_max_active_clusters(int cluster_size);

/* ------------------------------------------------------------------------------------------
 * Fused pass 1: prototype logits (H . W2^T, bf16 operands, fp32 TMEM accumulators) with the
 * row-wise online log-sum-exp of  u = logits*inv_tau + col2/log2e  computed in the GEMM epilogue;
 * the (rows x K) logit matrix is never written.  col2[k] is a per-prototype offset in LOG2 units
 * (e.g. (b2[k]-center[k])*inv_tau*log2e).  Outputs natural-log LSE and/or log2 LSE per row.
 * Replaces F.softmax/F.log_softmax statistics of scripts/phase5_big_run.py:703,706.
 * ------------------------------------------------------------------------------------------ */
DINOX_API size_t dinox_head_stats_workspace_bytes(int64_t rows, int64_t K);
DINOX_API int dinox_head_stats(const void* H, const void* W2, int64_t rows, int64_t K, int64_t D,
                               int64_t ldh, int64_t ldw, float inv_tau, const float* col2,
                               float* lse_nat, float* lse2, void* workspace, dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused pass 2: for E (student row, teacher row) entries recompute both logit tiles side by side
 * in TMEM and emit, without materialising logits or probabilities in HBM:
 *   Gt[k,e]  = cw[e]*inv_tau_s*( softmax_s[e,k] - q_t[e,k] )          (bf16, (K, ldg) = dL/dlogits^T)
 *   loss[0..1] (+)= sum_e cw[e] * sum_k q_t[e,k] * (-ln softmax_s[e,k])  (two fp32 on device:
 *              [0] entries < alt_from (CLS pairs), [1] entries >= alt_from (iBOT), [1]=0 if no alt)
 *   db2_partial[dinox_head_grad_db2_rows(E), K]  column partial sums of Gt (reduce with dinox_cols_sum), optional
 * with softmax_s = 2^(S*inv_tau_s*log2e + cs2[k] - lse2[e]), q_t = 2^(T*inv_tau_t*log2e + ct2[k] - rb2[e]);
 * entries >= alt_from (multiple of 128) use ct2_alt (iBOT patch centre).  HsE/HtE: (E, D) bf16
 * gathered head activations of the entry's student / teacher row.  Cross-entropy of
 * scripts/phase5_big_run.py:703-717 and its autograd (grad of log_softmax) in one kernel.
 * ------------------------------------------------------------------------------------------ */
DINOX_API size_t dinox_head_grad_workspace_bytes(int64_t K, int64_t E);
/* rows of db2_partial: one per (128-entry tile, epilogue column group); db2_partial is (rows, K) fp32 */
DINOX_API int64_t dinox_head_grad_db2_rows(int64_t E);
DINOX_API int dinox_head_grad(const void* W2s, const void* W2t, const void* HsE, const void* HtE,
                              int64_t K, int64_t D, int64_t E, int64_t ldw_s, int64_t ldw_t,
                              int64_t ldh_s, int64_t ldh_t, float inv_tau_s, float inv_tau_t,
                              const float* cs2, const float* ct2, const float* ct2_alt,
                              int64_t alt_from, const float* lse2_e, const float* rb2_e,
                              const float* cw_e, void* Gt, int64_t ldg, float* db2_partial,
                              float* loss_out, int loss_accumulate, void* workspace,
                              dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Teacher in ONE pass (default fused path): the statistics of dinox_head_stats PLUS the teacher
 * probabilities themselves in un-normalised 16-bit form, so that pass 2 never recomputes a teacher
 * logit.  For x[i,k] = (H.W2^T)[i,k]*inv_tau*log2e + col2[k] and every granule g of 128 or 64 prototypes
 * (dinox_head_teacher_granules_per_tile()):
 *   refs[g][i] = max_{k in g} x[i,k]                       (log2 units; (granules, ld_refs) fp32)
 *   qt[i,k]    = fp16( 2^(x[i,k] - refs[g(k)][i]) )        ((rows, ldq) fp16, ldq >= 256*ceil(K/256);
 *                                                            columns in [K, ldq) are written as 0)
 *   lse2[i]    = log2 sum_k 2^x[i,k]  (and/or the natural-log lse_nat), from the unrounded values
 * so softmax((t - c)/tau_t)[i,k] = qt[i,k] * 2^(refs[g][i] - lse2[i])  (scripts/phase5_big_run.py:703).
 * Rows of M tiles at or beyond alt_from_row (multiple of 128) use col2_alt (iBOT patch centre); NULL = none.
 * ------------------------------------------------------------------------------------------ */
DINOX_API size_t dinox_head_teacher_workspace_bytes(int64_t rows, int64_t K);
/* granules per 256-prototype tile (2: 128 prototypes each, or 4: 64 each - a build constant); refs has
 * granules_per_tile * ceil(K/256) rows and granule g(k) = k / (256 / granules_per_tile) */
DINOX_API int dinox_head_teacher_granules_per_tile(void);
/* prototypes per tile of the read-back pair (256; 192 in the 12-epilogue-warp build): qt rows are padded to a multiple
 * of it, and wherever this header says ceil(K/256) for qt / refs it means ceil(K / tile_cols) */
DINOX_API int dinox_head_teacher_tile_cols(void);
DINOX_API int dinox_head_teacher(const void* H, const void* W2, int64_t rows, int64_t K, int64_t D,
                                 int64_t ldh, int64_t ldw, float inv_tau, const float* col2,
                                 const float* col2_alt, int64_t alt_from_row, void* qt, int64_t ldq,
                                 float* refs, int64_t ld_refs, float* lse_nat, float* lse2,
                                 void* workspace, dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Pass 2 on top of dinox_head_teacher: for E entries (student row gathered into HsE, teacher row
 * trow_e[e] of qt/refs) recompute only the STUDENT logit tile in TMEM and emit
 *   G[e,k]   = cw[e]*inv_tau_s*( softmax_s[e,k] - q_t[e,k] )            (bf16, (E, ldg) = dL/dlogits)
 *   loss[0..1] (+)= sum_e cw[e] * sum_k q_t[e,k] * (-ln softmax_s[e,k])   ([0]: e < alt_from, [1]: the rest)
 *   loss[2]     = loss[0] + loss[1]                                      (loss_out holds THREE fp32)
 *   db2_partial[dinox_head_grad2_db2_rows(E), K]  column sums of G per 32 entries (reduce with dinox_cols_sum)
 * with softmax_s = 2^(S*inv_tau_s*log2e + cs2[k] - lse2_e[e]) and
 *      q_t       = qt[trow_e[e],k] * 2^(refs[g(k)][trow_e[e]] - rb2_e[e]).
 * srow_e != NULL: the row statistics are NOT gathered per entry - lse2_e is indexed by the entry's student row
 * srow_e[e] (-1 = padding entry) and rb2_e by its teacher row trow_e[e]; entries with cw_e[e] == 0 are dead.
 * ticket != NULL (one uint32, zero before the first launch, left zero by every launch): the kernel itself adds up
 * the per-warp loss partials in a fixed order (the warp that finishes last does it) instead of a follow-up launch.
 * Cross-entropy of scripts/phase5_big_run.py:706-717 and its autograd in one kernel.
 * ------------------------------------------------------------------------------------------ */
DINOX_API size_t dinox_head_grad2_workspace_bytes(int64_t E, int64_t K);
DINOX_API int64_t dinox_head_grad2_db2_rows(int64_t E);
DINOX_API int dinox_head_grad2(const void* HsE, const void* W2s, int64_t E, int64_t K, int64_t D,
                               int64_t ldh, int64_t ldw, float inv_tau_s, const float* cs2,
                               const float* lse2_e, const float* cw_e, const float* rb2_e,
                               const int32_t* trow_e, const int32_t* srow_e, const void* qt, int64_t ldq,
                               const float* refs, int64_t ld_refs, int64_t alt_from, void* G, int64_t ldg,
                               float* db2_partial, float* loss_out, int loss_accumulate,
                               void* workspace, uint32_t* ticket, dinox_stream_t stream);

/* batched variant: `batches` independent problems, element strides between problems.  Used for the
 * per-image Gram backward  dXn[b] = alpha * Delta[b] @ Xn[b]  (autograd of torch.bmm,
 * scripts/phase5_big_run.py:727). */
DINOX_API int dinox_gemm_bf16_batched(const void* A, const void* B, void* C, int64_t batches, int64_t M,
                                      int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                                      int64_t stride_a, int64_t stride_b, int64_t stride_c,
                                      int a_mn_major, int b_mn_major, int out_dtype, int accumulate,
                                      float alpha, const float* alpha_dev, dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a6/a7  Gram anchoring (scripts/phase5_big_run.py:723-739).
 * dinox_normalize_tokens: xn[b,t,:] = bf16(x / max(||x||,1e-12)) for tokens skip..T-1 of feats
 *   (B, T, D) given by element strides (F.normalize, :726); inv_norm kept for the backward.
 * dinox_gram_diff: per image Gs = Xs Xs^T and Gt = Xt Xt^T as two TMEM accumulators of one tile,
 *   loss = loss_scale * sum (Gs-Gt)^2 (mse_loss, :738) and Delta = Gs-Gt as bf16 (B, tokens, ldd)
 *   for the backward GEMM; the Gram matrices themselves are never written.
 * dinox_normalize_tokens_bwd: dX = (dXn - xn <xn,dXn>) * inv_norm * scale into grad (B, T, D).
 * ------------------------------------------------------------------------------------------ */
DINOX_API int dinox_normalize_tokens(const void* feats, int dtype, int64_t batch, int64_t tokens_total,
                                     int64_t D, int64_t stride_b, int64_t stride_t, int skip,
                                     void* xn_bf16, float* inv_norm, dinox_stream_t stream);
DINOX_API int dinox_normalize_tokens_bwd(const void* feats, int dtype, int64_t batch, int64_t tokens_total,
                                         int64_t D, int64_t stride_b, int64_t stride_t, int skip,
                                         const float* dxn, const float* inv_norm, const float* scale_dev,
                                         float scale, float* grad, int64_t gstride_b, int64_t gstride_t,
                                         dinox_stream_t stream);
DINOX_API size_t dinox_gram_diff_workspace_bytes(int64_t batches, int64_t tokens);
DINOX_API int dinox_gram_diff(const void* xn_s, const void* xn_t, int64_t batches, int64_t tokens, int64_t D,
                              void* delta, int64_t ldd, float loss_scale, float* loss_out, void* workspace,
                              dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Staging helpers around the contractions (caller-side glue of scripts/phase5_big_run.py:1746-1747:
 * `feats[:, 0]` slicing, autocast's fp32->bf16 casts, nn.GELU of zoo/arch.py:254).
 * ------------------------------------------------------------------------------------------ */
/* dst[r,:] = bf16(src[idx ? idx[r] : r, :]); idx[r] < 0 writes a zero row (padding entries) */
DINOX_API int dinox_gather_cast_bf16(const void* src, int src_dtype, int64_t ld_src, const int64_t* idx,
                                     int64_t rows, int64_t D, void* dst, int64_t ld_dst,
                                     dinox_stream_t stream);
/* two sources into one destination in one launch: dst rows [0, rows0) from (src0, idx0), rows [rows0, rows0 + rows1)
 * from (src1, idx1) - the [CLS rows | masked-patch rows] staging of one branch (:1746-1747).  Rows must be 16-byte
 * vectors (D % 8 == 0, aligned pitches); both sources share src_dtype. */
DINOX_API int dinox_gather_cast_bf16_2(const void* src0, int64_t ld_src0, const int64_t* idx0, int64_t rows0,
                                       const void* src1, int64_t ld_src1, const int64_t* idx1, int64_t rows1,
                                       int src_dtype, int64_t D, void* dst, int64_t ld_dst, dinox_stream_t stream);
DINOX_API int dinox_gather_f32(const float* src, const int64_t* idx, int64_t n, float fill, float* out,
                               dinox_stream_t stream);
/* dst[idx[r], :] = src[r, :] for r < rows (fp32, D % 4 == 0, unique idx, idx < 0 skipped): the backward of
 * gathering iBOT rows straight out of the backbone's token tensor (`feats[:, 1:][mask]`) */
DINOX_API int dinox_scatter_rows_f32(const float* src, int64_t ld_src, const int64_t* idx, int64_t rows, int64_t D,
                                     float* dst, int64_t ld_dst, dinox_stream_t stream);
/* h = bf16(gelu_erf(a)), n elements */
DINOX_API int dinox_gelu_fwd(const float* a, int64_t n, void* h_bf16, dinox_stream_t stream);
/* da = bf16(dh * (*scale_dev) * gelu'(a)); colsum_partial (dinox_gelu_bwd_workspace_bytes / (4 D) rows, D): partial column sums of da per slab of rows */
DINOX_API size_t dinox_gelu_bwd_workspace_bytes(int64_t rows, int64_t D);
DINOX_API int dinox_gelu_bwd(const float* dh, const float* a, int64_t rows, int64_t D, const float* scale_dev,
                             void* da_bf16, float* colsum_partial, dinox_stream_t stream);
/* the same with dh[r,:] = sum over the row's entries i in [ptr[r], ptr[r+1]) and the split-K slabs s of
 * src[s*slab_stride + ent[i]*ld_src, :] formed on the fly (dinox_gather_sum_rows + dinox_gelu_bwd in one launch,
 * identical summation order) */
DINOX_API int dinox_gelu_bwd_gather(const float* src, int64_t ld_src, int slabs, int64_t slab_stride,
                                    const int64_t* ptr, const int64_t* ent, const float* a, int64_t rows, int64_t D,
                                    const float* scale_dev, void* da_bf16, float* colsum_partial,
                                    dinox_stream_t stream);
/* out[k] = alpha * sum_d W[k,d] x[d] + beta * bias[k]   (W bf16 (K,D), x fp32): batch-mean teacher
 * logits from the mean head activation, used for the centre update of the fused path (:686-690) */
DINOX_API int dinox_gemv_bf16(const void* W, int64_t ldw, const float* x, int64_t K, int64_t D, float alpha,
                              const float* bias, float beta, float* out, dinox_stream_t stream);
/* nvec (1..4) vectors X (nvec, D) against the same matrix in ONE pass over W: out (nvec, K),
 * out[v][k] = alphas_host[v] / divisors_dev[v] * sum_d W[k,d] X[v,d] + beta * bias[k]  (CLS and iBOT centre means
 * together; divisors_dev = optional DEVICE vector, e.g. the all-reduced row counts, NULL = 1) */
DINOX_API int dinox_gemv_bf16_multi(const void* W, int64_t ldw, const float* X, int nvec, int64_t K, int64_t D,
                                    const float* alphas_host, const float* divisors_dev, const float* bias,
                                    float beta, float* out, dinox_stream_t stream);
/* the same GEMV with the centre EMA folded in (scripts/phase5_big_run.py:686-690): instead of storing the nvec result
 * vectors, targets[v][k] = momenta[v]*targets[v][k] + (1-momenta[v])*value[v][k] in place (HOST arrays of nvec
 * device pointers / momenta) - one launch for "mean teacher logits -> centre and patch-centre update" */
DINOX_API int dinox_gemv_bf16_multi_ema(const void* W, int64_t ldw, const float* X, int nvec, int64_t K, int64_t D,
                                        const float* alphas_host, const float* divisors_dev, const float* bias,
                                        float beta, float* const* targets, const float* momenta_host,
                                        dinox_stream_t stream);
/* dst[i] (+)= scale*(*scale_dev) * sum_{s<slabs} src[s*slab_stride + i]: fixed-order reduction of split-K slabs */
DINOX_API int dinox_sum_slabs(const float* src, int slabs, int64_t slab_stride, int64_t n, const float* scale_dev,
                              float scale, float* dst, int accumulate, dinox_stream_t stream);
/* dst[r,:] (+)= scale*(*scale_dev) * sum_{i in [ptr[r],ptr[r+1])} sum_{s<slabs} src[s*slab_stride + ent[i]*ld_src,:]
 * (fp32; slabs > 1 sums the partial outputs of dinox_gemm_bf16_splitk in fixed order) */
DINOX_API int dinox_gather_sum_rows(const float* src, int64_t ld_src, int slabs, int64_t slab_stride,
                                    const int64_t* ptr, const int64_t* ent, int64_t rows, int64_t D,
                                    const float* scale_dev, float scale, float* dst, int64_t ld_dst,
                                    int accumulate, dinox_stream_t stream);
/* out[r,:] = float(src[idx[r],:]) (idx NULL: identity, idx < 0: zero row): index-named token rows as fp32 rows */
DINOX_API int dinox_gather_rows_f32(const void* src, int dtype, int64_t ld_src, const int64_t* idx, int64_t rows,
                                    int64_t D, float* out, int64_t ld_out, dinox_stream_t stream);
/* dst[idx[r],:] += src[r,:] (fp32 rows, unique idx, idx < 0 skipped): adds the gradients of index-gathered rows
 * into a gradient tensor that already holds another term (Gram anchoring + iBOT on the same token tensor) */
DINOX_API int dinox_scatter_add_rows_f32(const float* src, int64_t ld_src, const int64_t* idx, int64_t rows,
                                         int64_t D, float* dst, int64_t ld_dst, dinox_stream_t stream);
/* per-prototype offsets of the fused passes (log2 units) in one launch: cs2 = b2_student/tau_s*log2e,
 * ct2 = (b2_teacher - center)/tau_t*log2e, ct2_patch = (b2_teacher - center_patch)/tau_t*log2e (optional) -
 * the bias add of zoo/arch.py:256 and the centring of scripts/phase5_big_run.py:703 folded into the epilogues */
DINOX_API int dinox_head_offsets(const float* b2_student, const float* b2_teacher, const float* center,
                                 const float* center_patch, float inv_tau_s, float inv_tau_t, float* cs2,
                                 float* ct2, float* ct2_patch, int64_t K, dinox_stream_t stream);
/* entry weights of pass 2: out = base (the CLS pair weights, 0 for padding) with out[offset + i] = mask_weights[i]*scale
 * for the n iBOT entries */
DINOX_API int dinox_entry_weights(const float* base, int64_t total, const float* mask_weights, int64_t offset,
                                  int64_t n, float scale, float* out, dinox_stream_t stream);
DINOX_API int dinox_fill_f32(float* p, int64_t n, float v, dinox_stream_t stream);
/* a9 step glue (scripts/phase5_big_run.py:1749-1772, `loss = L_dino + w_g*L_gram (+ w_k*L_koleo); loss /= accum`):
 *   out[0] = scale * sum_{i<n} weights[i] * terms[i][0]      (n <= 8 device scalars, fixed order)
 *   out_unscaled[0] = the same sum without `scale` (optional, NULL to skip; the logged loss)
 * and for its backward  out[i] = upstream[0] * scale * weights[i].  `terms` / `weights` are HOST arrays. */
DINOX_API int dinox_scalar_combine(const float* const* terms, const float* weights, int n, float scale,
                                   float* out, float* out_unscaled, dinox_stream_t stream);
DINOX_API int dinox_scalar_fanout(const float* upstream, const float* weights, int n, float scale,
                                  float* out, dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * fp32-faithful contraction mode.  The reference without `--amp` (its default, scripts/phase5_big_run.py:1322)
 * runs nn.Linear (zoo/arch.py:252-256) and torch.bmm (:727) in true fp32.  The drop-in reproduces that on the bf16
 * tensor pipe by evaluating every contraction as three GEMMs on hi/lo bf16 splits of the fp32 operands
 * (A_hi.B_hi + A_hi.B_lo + A_lo.B_hi, fp32 accumulation; ~16 mantissa bits per product).  These are the
 * element-wise helpers of that mode; the GEMMs themselves are dinox_gemm_bf16[_batched] with accumulate = 1.
 * ------------------------------------------------------------------------------------------ */
/* hi = bf16(x), lo = bf16(x - hi) for a (rows, cols) matrix with row pitch ld_src; outputs have pitch ld_dst */
DINOX_API int dinox_split_bf16(const void* src, int dtype, int64_t rows, int64_t cols, int64_t ld_src, void* hi,
                               void* lo, int64_t ld_dst, dinox_stream_t stream);
DINOX_API int dinox_gelu_fwd_f32(const float* a, int64_t n, float* h, dinox_stream_t stream);
DINOX_API size_t dinox_gelu_bwd_f32_workspace_bytes(int64_t rows, int64_t D);
/* da = dh * (*scale_dev) * gelu'(a) in fp32; colsum_partial (workspace) holds per-16-row-slab column sums */
DINOX_API int dinox_gelu_bwd_f32(const float* dh, const float* a, int64_t rows, int64_t D, const float* scale_dev,
                                 float* da, float* colsum_partial, dinox_stream_t stream);
DINOX_API int dinox_normalize_tokens_f32(const void* feats, int dtype, int64_t batch, int64_t tokens_total, int64_t D,
                                         int64_t stride_b, int64_t stride_t, int skip, float* xn, float* inv_norm,
                                         dinox_stream_t stream);
DINOX_API size_t dinox_sqdiff_workspace_bytes(void);
/* loss_out = scale * sum (a - b)^2 over n elements (fixed order), delta = a - b (optional)  - F.mse_loss, :738 */
DINOX_API int dinox_sqdiff_f32(const float* a, const float* b, int64_t n, float scale, float* delta, float* loss_out,
                               void* workspace, dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a11 KoLeo regulariser on head outputs (scripts/phase5_big_run.py:742-773, wired at :1764-1766):
 *   x = z / max(||z||, 1e-12);  d_i = min_{j != i} ||x_i - x_j||;  loss = -mean_i log(d_i + eps).
 * dinox_koleo_rownorm: inv_norm[r] (and an optional bf16 copy of z for the ranking GEMM).
 * The caller forms gram = z_bf16 z_bf16^T with dinox_gemm_bf16_splitk (+ dinox_sum_slabs); it only
 * ranks neighbours.  dinox_koleo_fwd picks dinox_koleo_candidates() candidates per row, recomputes
 * their distances exactly in fp32 from z, and reduces the loss; nn/dist feed dinox_koleo_bwd, which
 * writes dL/dz (same dtype as z) = autograd of normalize -> cdist -> min -> log.  rows <= 1024.
 * ------------------------------------------------------------------------------------------ */
DINOX_API int dinox_koleo_candidates(void);
DINOX_API int dinox_koleo_rownorm(const void* z, int dtype, int64_t rows, int64_t K, int64_t ld, float* inv_norm,
                                  void* z_bf16, int64_t ldb, dinox_stream_t stream);
DINOX_API int dinox_koleo_fwd(const void* z, int dtype, int64_t rows, int64_t K, int64_t ld, const float* inv_norm,
                              const float* gram, int64_t ldg, float eps, int* cand, float* d2, int* nn, float* dist,
                              float* loss, dinox_stream_t stream);
DINOX_API int dinox_koleo_bwd(const void* z, int dtype, int64_t rows, int64_t K, int64_t ld, const float* inv_norm,
                              const int* nn, const float* dist, float eps, const float* upstream, void* dz,
                              int64_t ldd, dinox_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * multi-tensor AdamW step + global gradient norm, one launch for every tensor
 * (fp32 params / grads / moments).  Replaces the optimizer step of scripts/phase5_big_run.py:1781-1796:
 * the per-parameter `p.grad.norm(2).item()` loop (a host sync per tensor) and torch.optim.AdamW (:1621).
 * Arithmetic of torch.optim.AdamW's single-tensor path; bias corrections 1 - beta^step are passed in
 * (hyper-parameters are doubles like the python floats torch derives its scalars from).  grad_scale multiplies every gradient first (1/loss-scale; 1 for bf16 training).
 * grad_norm_out: device scalar = || grad_scale * g ||_2 over all tensors (fixed reduction order).
 * ------------------------------------------------------------------------------------------ */
typedef struct dinox_adamw_plan dinox_adamw_plan;
DINOX_API int dinox_adamw_plan_create(void* const* params, const void* const* grads, void* const* exp_avg,
                                      void* const* exp_avg_sq, const int64_t* numel, int n_tensors,
                                      dinox_adamw_plan** out);
DINOX_API int dinox_adamw_plan_destroy(dinox_adamw_plan* plan);
DINOX_API int dinox_adamw_step(const dinox_adamw_plan* plan, double lr, double beta1, double beta2, double eps,
                               double weight_decay, double bias_correction1, double bias_correction2,
                               float grad_scale, float* grad_norm_out, dinox_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DINOX_B200_H_ */
