/*
 * ppnp_b200.h -- C ABI of libppnp_b200.so: the PPNP/APPNP propagation hot path of
 * bkj/ppnp as hand-written CUDA for sm_100a (B200).
 *
 * The reference is pure Python; there is no FFI in it.  Each entry point below
 * replaces the reference expression cited next to it (paths are relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     name ends in _host; the caller owns every buffer (the library never
 *     allocates or frees caller-visible memory and keeps no global device state);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing
 *     synchronises the device unless stated;
 *   - return 0 on success, a negative PPNP_E* code on failure; the message of the
 *     last failure on the calling thread is ppnp_last_error(); nothing throws.
 */
#ifndef PPNP_B200_H
#define PPNP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPNP_OK 0
#define PPNP_EINVAL (-1)   /* bad argument (shape, alignment, null pointer)  */
#define PPNP_ECUDA (-2)    /* a CUDA runtime call or a launch failed          */
#define PPNP_ENOTSUP (-3)  /* configuration not supported by this build       */

/* normalisation modes of helpers.py:58-66 calc_A_hat(adj, mode) */
#define PPNP_MODE_SYM 0    /* D^-1/2 (A+I) D^-1/2   helpers.py:61-63 */
#define PPNP_MODE_RW 1     /* D^-1 (A+I)            helpers.py:64-66 */
/* 'sym' with the input ALREADY scaled, H_in = D^-1/2 H (the fused encoder tail, ppnp_linear_rowscale, writes it
 * that way): every step is value-free, the stored values of A_hat are never read and need not exist. */
#define PPNP_MODE_SYM_Y0 2
/* OR-ed into the `mode` argument of ppnp_appnp_propagate: always run the K steps as per-step launches, never as the
 * one-launch cooperative kernel that small graphs get by default (tests of the per-step epilogues on small graphs,
 * callers that must not use cooperative launches). */
#define PPNP_MODE_PER_STEP 0x100

/* epilogue of one propagation step, out[r] = a(deg_r) * acc_r + b(deg_r) * T[r] */
#define PPNP_EPI_PLAIN 0   /* a = 1-alpha,            b = alpha            stored values, Z-space         */
#define PPNP_EPI_Z2Y 1     /* a = (1-alpha)/sqrt(d),  b = alpha/sqrt(d)    stored values, first step Z->Y  */
#define PPNP_EPI_Y 2       /* a = (1-alpha)/d,        b = alpha/sqrt(d)    value-free, Y = D^-1/2 Z space  */
#define PPNP_EPI_Y2Z 3     /* a = (1-alpha)/sqrt(d),  b = alpha            value-free, last step Y->Z      */
#define PPNP_EPI_RW 4      /* a = (1-alpha)/d,        b = alpha            value-free 'rw' mode; also the inner steps of
                              PPNP_MODE_SYM_Y0 (T = Y0 = D^-1/2 H)                                      */
#define PPNP_EPI_Y02Z 5    /* a = (1-alpha)/sqrt(d),  b = alpha*sqrt(d)    last step of PPNP_MODE_SYM_Y0, T = Y0      */
/* OR-ed onto one of the above: out[r] = a * acc_r + 1 * T[r] with T == Zout, i.e. the step ADDS its
 * contribution to rows an earlier pass already wrote (multi-pass steps of the partitioned form,
 * ppnp_b200/dist.py: local columns first, halo columns once they have arrived). */
#define PPNP_EPI_ACC 16
/* OR-ed onto PPNP_EPI_ACC: Zin may be the output buffer itself.  The caller guarantees that the rows the
 * stream GATHERS are not rows it produces (the hub-combine launch of ppnp_b200/dist.py HybridPushPropagation:
 * it gathers the partial-row slots of the buffer and adds them to hub rows of the same buffer). */
#define PPNP_EPI_INPLACE 32

/* peers addressable by the fused halo push (ppnp_spmm_step_push): ranks of one NVSwitch domain */
#define PPNP_MAX_PEERS 8

/* bit 31 of a stream column: last edge of its segment; of a seg_row entry: partial segment */
#define PPNP_FLAG 0x80000000u

const char* ppnp_last_error(void);
int ppnp_version(void);
/* SM count, compute capability and L2 size of the current device. */
int ppnp_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* l2_bytes);

/* ------------------------------------------------------------------------------------------
 * (1) CSR build + normalisation.                      replaces helpers.py:58-66 calc_A_hat
 *
 * in : canonical CSR of adj (sparsegraph.py:191-222 standardize): indptr int32[n+1],
 *      indices int32[nnz] sorted per row, data fp32[nnz] or NULL (= all ones).
 * out: structure of A = adj + I (helpers.py:59; the diagonal is merged in sorted position,
 *      an existing diagonal entry gets +1), bit-exact with scipy:
 *        out_indptr int32[n+1], out_indices int32[cap >= nnz+n],
 *        out_deg    fp64[n]   D = rowsum(A)                        (helpers.py:60)
 *        out_val64  fp64[cap] (D_i^-1/2 a_ij) D_j^-1/2 in that order, or (1/D_i) a_ij   (nullable)
 *        out_val32  fp32[cap] the same value rounded once to fp32                        (nullable)
 *        out_dinv   fp32[n]   D^-1/2 ('sym') or D^-1 ('rw') rounded to fp32              (nullable)
 * workspace: ppnp_csr_normalize_workspace_bytes(n) bytes.
 * ---------------------------------------------------------------------------------------- */
int64_t ppnp_csr_normalize_workspace_bytes(int64_t n);
int ppnp_csr_normalize(const int32_t* indptr, const int32_t* indices, const float* data,
                       int64_t n, int64_t nnz, int32_t mode,
                       int32_t* out_indptr, int32_t* out_indices, double* out_deg,
                       double* out_val64, float* out_val32, float* out_dinv,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2) APPNP propagation.        Z_{k+1} = (1-alpha) A_hat Z_k + alpha H, Z_0 = H   (north_star;
 *     the reference holds only the K->inf limit, model.py:63 with helpers.py:68-71)
 *
 * The adjacency is consumed as an "edge stream" (ppnp_b200/plan.py builds it from the
 * normalised CSR): the rows of A_hat in processing order, cut into segments that never cross
 * a chunk of `chunk_edges` edges.  cols[e] = column | PPNP_FLAG on the last edge of a
 * segment (padding after the last segment: column 0, no flag, never emitted); seg_row[s] = row the segment finishes, or
 * PPNP_FLAG | slot when the row is split over several segments (the partial sums go to
 * partial[slot] and ppnp_spmm_fixup adds them up in slot order -- deterministic);
 * chunk_seg[c] = index of the first segment of chunk c.  vals (nullable) are the stored
 * values of A_hat in stream order; without them every edge has weight 1 (value-free form,
 * SURVEY.md section 8d) and the row scaling is done by the epilogue from the row degree.
 * ---------------------------------------------------------------------------------------- */
/* plan->flags.
 * PPNP_PLAN_WIDE_CTA: launch one 1024-thread CTA per SM instead of four 256-thread ones, so that all
 *   the warps of an SM walk CONSECUTIVE chunks -- for "carved" streams (plan.py build_carved_plan)
 *   whose leading part lists, column block by column block, the pieces of rows that fall into an
 *   L1-sized block of hot columns: the gathered rows of the block are then re-used out of the SM's
 *   L1 instead of crossing the L2 -> SM fabric once per edge.
 * PPNP_PLAN_LANE_GROUP(flags) = G > 0: cols (and vals) are stored "lane-transposed" for lane groups
 *   of G lanes: the 4 index words one lane consumes over 4 / SR consecutive slabs (SR = max(1, 16/G)
 *   words per lane and slab) are contiguous, so each lane stages them with ONE 16-byte cp.async.
 *   Stored position of logical chunk position p (slab j = p / SE, word r = (p % SE) / G, lane
 *   l = p % G, SE = SR * G, CPS = 4 / SR):  (j / CPS) * CPS * SE + l * 4 + (j % CPS) * SR + r.
 *   A launch whose feature width asks for another group size rejects the plan (PPNP_EINVAL). */
#define PPNP_PLAN_WIDE_CTA 1
#define PPNP_PLAN_LANE_GROUP(flags) (((flags) >> 8) & 0xff)

typedef struct ppnp_plan {
    int64_t n;              /* rows                                                     */
    int64_t n_edges;        /* stream length, multiple of chunk_edges                   */
    int64_t n_chunks;       /* n_edges / chunk_edges, multiple of 32                    */
    int64_t n_segs;
    int64_t n_fix;          /* rows split over several segments                         */
    int64_t n_slots;        /* partial segments                                          */
    int32_t chunk_edges;    /* edges per chunk (multiple of 128)                        */
    int32_t flags;          /* PPNP_PLAN_* bits below; 0 = row-major stream, linear index layout */
    const int32_t* cols;      /* [n_edges]                                              */
    const float* vals;        /* [n_edges] or NULL                                      */
    const int32_t* seg_row;   /* [n_segs + 64] (64 readable spare entries after the last one) */
    const int32_t* chunk_seg; /* [n_chunks]                                             */
    const int32_t* fix_ptr;   /* [n_fix + 1] slot ranges                                */
    const int32_t* fix_row;   /* [n_fix]                                                */
    const float* fix_deg;     /* [n_fix] row degree (edge count incl. self loop)        */
    const float* row_deg;     /* [n] or NULL: degree of every row, for streams that hold only
                                 part of a row's edges (NULL: degree = edges of the segment) */
} ppnp_plan_t;

/* One step: out = a * (A_hat-or-(A+I)) Zin + b * T with the epilogue `epi`.
 * Zin, T, Zout: n x F fp32 row-major with leading dimension ld (floats); Zout must not alias
 * Zin.  partial: n_slots x ld floats (may be NULL when n_slots == 0). use_vals != 0 selects the
 * stored-value form (plan->vals must be set). */
int ppnp_spmm_step(const ppnp_plan_t* plan, const float* Zin, const float* T, float* Zout,
                   float* partial, int64_t ld, int32_t F, float alpha, int32_t epi,
                   int32_t use_vals, void* stream);

/* The same step with the halo push FUSED into the epilogue (partitioned propagation, no counterpart in
 * the single-process reference): every finished row r with push_ptr[r] < push_ptr[r+1] is also
 * written to peer_bases[code >> 28] + (code & 0x0fffffff) * ld for code in push_code[push_ptr[r] ..
 * push_ptr[r+1]) -- the halo slots of that row in the peers' mappings (NVLink peer memory) of the
 * output buffer.  push_first[r] summarises the list (-1: none, >= 0: the single code, <= -2: several)
 * so that the common one-destination case needs no list walk.  peer_bases_host is a HOST array of
 * n_peers device pointers.  The caller orders
 * the next step after these writes with a barrier across the ranks. */
int ppnp_spmm_step_push(const ppnp_plan_t* plan, const float* Zin, const float* T, float* Zout,
                        float* partial, int64_t ld, int32_t F, float alpha, int32_t epi,
                        int32_t use_vals, const int32_t* push_ptr, const int32_t* push_code,
                        const int32_t* push_first, const void* const* peer_bases_host, int32_t n_peers,
                        void* stream);

/* K steps from Z_0 = H.  mode PPNP_MODE_SYM: value-free Y-space iteration when plan->vals is
 * given only for the first step (use_vals == 0), stored values every step when use_vals != 0.
 * Result in Z (n x ld); scratch is a second n x ld buffer; partial as above.
 * Used for the forward AND the backward pass (A_hat symmetric). */
int ppnp_appnp_propagate(const ppnp_plan_t* plan, const float* H, float* Z, float* scratch,
                         float* partial, int64_t ld, int32_t F, int32_t K, float alpha,
                         int32_t mode, int32_t use_vals, void* stream);

/* The same K steps in ONE cooperative launch (grid-wide barriers between the steps; stored-value
 * form).  ppnp_appnp_propagate takes this path by itself for small graphs (<= 4096 chunks) unless
 * PPNP_PERSISTENT=0 is set in the environment. */
int ppnp_appnp_propagate_persistent(const ppnp_plan_t* plan, const float* H, float* Z, float* scratch,
                                    float* partial, int64_t ld, int32_t F, int32_t K, float alpha,
                                    void* stream);

/* ------------------------------------------------------------------------------------------
 * (2b) The same step for the HUB rows of a skewed graph with their accumulators resident in shared
 *      memory (csrc/appnp_tiled.cu; ppnp_b200/plan.py build_tiled_plan).  No counterpart in the reference
 *      (north_star subsystem 2: "gathers of Z rows staged through shared memory").
 *
 * A CTA owns a group of rows ("slots": a row, or one part of a row too long for one warp) and one slice of
 * `slice_width` floats of the feature dimension; its warps own disjoint slots and walk their own edge
 * streams, sorted by (column window, slot, column), in slabs of 32 edges:
 *   cols[e]        column | PPNP_FLAG on the last edge of a piece (the edges of one slot in one window)
 *   vals[e]        stored A_hat values (nullable: value-free form)
 *   slab_meta[2s]  index of the first piece that ends in slab s or later; slab_meta[2s+1] = column window
 *                  of the slab, numbered from 1 (non-decreasing along a warp's stream; paces the warps of a
 *                  CTA), bit 30 set when some slot ends twice inside the slab
 *   piece_slot[p]  CTA-local slot the p-th piece adds to (32 readable spare entries after the last one);
 *                  padding edges (column 0, no flag) only follow the last piece of a warp
 *   warp_slab_ptr  [n_ctas * warps_per_cta + 1] slab range of every warp
 *   cta_slot_ptr   [n_ctas + 1] slot range of every CTA in slot_row
 *   slot_row[s]    row the slot belongs to; PPNP_FLAG | row for the 2nd.. part of a split row (the parts
 *                  of a row are consecutive slots of one CTA)
 *   row_deg[r]     degree (edge count incl. the self loop) of every row, for the epilogue
 * Rows that own no slot are not written: run one of the other kernels over them (ppnp_spmm_step with a plan of
 * the remaining rows, or ppnp_spmm_step_rows); ppnp_appnp_propagate_parts does that for K steps.  Results equal ppnp_spmm_step up
 * to the order of the fp32 additions.
 * ---------------------------------------------------------------------------------------- */
typedef struct ppnp_tiled_plan {
    int64_t n;               /* rows of the matrix                                        */
    int64_t n_slabs;         /* stream length / 32                                        */
    int64_t n_pieces;
    int32_t n_ctas;          /* row groups (grid.x); grid.y = F / slice_width             */
    int32_t warps_per_cta;   /* 1..16                                                     */
    int32_t slots_cap;       /* max slots of a CTA (+1 spare)                             */
    int32_t slack;           /* windows a warp may run ahead of the slowest warp of its CTA */
    const int32_t* cols;
    const float* vals;
    const int32_t* slab_meta;
    const int32_t* piece_slot;
    const int32_t* warp_slab_ptr;
    const int32_t* cta_slot_ptr;
    const int32_t* slot_row;
    const float* row_deg;
} ppnp_tiled_plan_t;

int ppnp_spmm_step_tiled(const ppnp_tiled_plan_t* plan, const float* Zin, const float* T, float* Zout,
                         int64_t ld, int32_t F, int32_t slice_width, float alpha, int32_t epi,
                         int32_t use_vals, void* stream);
/* ------------------------------------------------------------------------------------------
 * (2c) The same step for rows of LOW degree, one lane group per row, straight off the normalised CSR
 *      (csrc/appnp_rows.cu): no edge stream, no flags, no partial sums.  rows[0 .. n_rows) lists the rows to
 *      produce (any subset of [0, n), best in descending degree); indptr / indices / vals are the CSR of
 *      A_hat (vals nullable: value-free form, the degree of a row is its number of stored entries).  The
 *      push arguments are those of ppnp_spmm_step_push (all NULL / 0: no halo push).
 * ---------------------------------------------------------------------------------------- */
int ppnp_spmm_step_rows(const int32_t* indptr, const int32_t* indices, const float* vals,
                        const int32_t* rows, int64_t n_rows, int64_t n, const float* Zin, const float* T,
                        float* Zout, int64_t ld, int32_t F, float alpha, int32_t epi, int32_t use_vals,
                        const int32_t* push_ptr, const int32_t* push_code, const int32_t* push_first,
                        const void* const* peer_bases_host, int32_t n_peers, void* stream);

typedef struct ppnp_rows_plan {
    int64_t n;               /* rows of the matrix                 */
    int64_t n_rows;          /* rows this part produces            */
    const int32_t* indptr;   /* [n + 1] CSR of A_hat               */
    const int32_t* indices;
    const float* vals;       /* nullable                           */
    const int32_t* rows;     /* [n_rows]                           */
} ppnp_rows_plan_t;

/* K steps from Z_0 = H with the rows of the matrix split over up to three kernels by degree: `tiled` (hub
 * rows, shared-memory accumulators), `stream_plan` (row-major edge stream) and `rows` (one lane group per
 * row); each part is nullable, together they must produce every row exactly once.  Same step sequence,
 * epilogues and arguments as ppnp_appnp_propagate (partial: the stream part's partial-sum buffer). */
int ppnp_appnp_propagate_parts(const ppnp_tiled_plan_t* tiled, const ppnp_plan_t* stream_plan,
                               const ppnp_rows_plan_t* rows, const float* H, float* Z, float* scratch,
                               float* partial, int64_t ld, int32_t F, int32_t slice_width, int32_t K,
                               float alpha, int32_t mode, int32_t use_vals, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2d) Encoder tail fused with the row scaling of the propagation.
 *      replaces model.py:51 (the encoder's last nn.Linear) where it feeds model.py:63 in APPNP mode:
 *        out[i, :] = scale_i * (A[i, :] @ W^T + bias)     A: n x hidden row-major (contiguous), W: C x hidden
 *      (nn.Linear weight layout), bias / scale nullable.  With scale = D^-1/2 (ppnp_csr_normalize's out_dinv) the
 *      result is Y0 = D^-1/2 H: ppnp_appnp_propagate(..., mode = PPNP_MODE_SYM_Y0) propagates it value-free in
 *      every step and returns Z = P_K(A_hat) H without H or the stored values of A_hat ever existing.
 *      backward: dA = scale * (dOut @ W) (nullable), dW = sum_i scale_i dOut_i^T A_i, dbias (nullable);
 *      two-stage fixed-order reduction (deterministic); workspace from the _workspace_bytes query.
 *      hidden <= 256, C <= 64.
 * ---------------------------------------------------------------------------------------- */
int ppnp_linear_rowscale(const float* A, int64_t n, int32_t hidden, const float* W, const float* bias,
                         const float* scale, float* out, int64_t ld_out, int32_t C, void* stream);
int64_t ppnp_linear_rowscale_backward_workspace_bytes(int64_t n, int32_t hidden, int32_t C);
int ppnp_linear_rowscale_backward(const float* A, const float* dOut, int64_t ld_dout, int64_t n,
                                  int32_t hidden, int32_t C, const float* W, const float* scale,
                                  float* dA, float* dW, float* dbias, void* workspace,
                                  int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (3) Exact PPNP.                                      replaces helpers.py:68-71 compute_ppr
 *     Pi = alpha (I - (1-alpha) A_hat)^-1 by power iteration on all n right-hand sides:
 *     Pi_0 = I, Pi_{k+1} = (1-alpha) A_hat Pi_k + alpha I.   A_hat: normalised CSR with fp32
 *     values (output of ppnp_csr_normalize).  Pi, scratch: n x n fp32 row-major.  Result in Pi.
 * ---------------------------------------------------------------------------------------- */
int ppnp_ppr_dense(const int32_t* indptr, const int32_t* indices, const float* val,
                   int64_t n, float alpha, int32_t K, float* Pi, float* scratch, void* stream);
/* The same matrix by the Chebyshev-accelerated iteration (the iteration matrix (1-alpha) A_hat has a real
 * spectrum in [-(1-alpha), 1-alpha]): x_{k+1} = w_{k+1} ((1-alpha) A_hat x_k + alpha I) + (1 - w_{k+1}) x_{k-1}.
 * Error ~ sigma^K with sigma = (1 - sqrt(1 - rho^2)) / rho, rho = 1 - alpha (0.627 for alpha = 0.1 against 0.9):
 * K ~ 40 reaches what the plain iteration needs ~150 steps for.  Same buffers, same bytes per step. */
int ppnp_ppr_dense_cheb(const int32_t* indptr, const int32_t* indices, const float* val,
                        int64_t n, float alpha, int32_t K, float* Pi, float* scratch, void* stream);


/* ------------------------------------------------------------------------------------------
 * (3b) Dense apply.                 replaces model.py:63  self.ppr[idx] @ H   (idx != NULL)
 *                                   and      model.py:65  ppr @ H             (idx == NULL)
 *      out[m x C] = op(Pi)[idx, :] @ H[n x C]; Pi is m_pi x n row-major (ld_pi elements).
 *      transpose != 0 computes the autograd adjoint  out[n x C] = Pi[idx, :]^T @ H[m x C].
 *      ppnp_gather_gemm_f32 : fp32 SIMT, the 1e-5 parity path.
 *      ppnp_gather_gemm_bf16: Pi in bf16, H converted to bf16 on the fly, fp32 accumulation
 *                             in TMEM via tcgen05.mma (the 1e-2 path).  workspace from
 *                             ppnp_gather_gemm_bf16_workspace_bytes.
 * ---------------------------------------------------------------------------------------- */
int ppnp_gather_gemm_f32(const float* Pi, int64_t ld_pi, const int64_t* idx, int64_t m, int64_t n,
                         const float* H, int64_t ld_h, int32_t C, float* out, int64_t ld_out,
                         int32_t transpose, void* stream);
int64_t ppnp_gather_gemm_bf16_workspace_bytes(int64_t m, int64_t n, int32_t C);
int ppnp_gather_gemm_bf16(const void* Pi_bf16, int64_t ld_pi, const int64_t* idx, int64_t m,
                          int64_t n, const float* H, int64_t ld_h, int32_t C, float* out,
                          int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream);
/* fp32 -> bf16 (round to nearest even) copy used to build the bf16 shadow of Pi. */
int ppnp_f32_to_bf16(const float* src, void* dst_bf16, int64_t count, void* stream);

/* ------------------------------------------------------------------------------------------
 * (4) batch-main.py path.
 *   ppnp_topk_thresh : batch-main.py:115  thresh, _ = ppr.topk(k, axis=-1); thresh[:, -1]
 *                      k-th largest of every row of the dense n_rows x n_cols fp32 matrix (bit-exact
 *                      selection, no sorting of the row).
 *   ppnp_topk_mask   : batch-main.py:116  ppr[ppr < thresh[:, -1]] = 0   (in place; entry (i, j)
 *                      is compared with thresh[j] -- the reference's broadcast)
 *   ppnp_dense_row_nnz / ppnp_dense_to_csr: compact the masked matrix (entries > 0) to CSR.
 *   ppnp_batch_support: batch-main.py:140-141  sel = (ppr[idx_batch] > 0).any(0) on the compact
 *                      form: mark[j] = 1 for every column in the support of the batch rows.
 *   ppnp_batch_support_colmap: batch-main.py:140-142 in ONE launch: sel[j] (one byte per column, 0/1: the
 *                      bool mask of line 141) and colmap[j] = position of column j inside sel, -1 outside it
 *                      (what line 142's ppr_sub[:, sel] amounts to on the compact form); *m_out = sel.sum()
 *                      (nullable).  One CTA, one byte of shared memory per column: n_cols <= 204800, else
 *                      PPNP_ENOTSUP (use ppnp_batch_support + a prefix sum).
 *   ppnp_batch_propagate: batch-main.py:142-146  logits = ppr[idx_batch][:, sel] @ Hsub where
 *                      Hsub = encoder(X[sel]); colmap[j] = position of column j in sel, or -1 for a
 *                      column outside the mask (skipped, like the reference's column selection).
 *                      transpose != 0: the autograd adjoint dHsub = ppr_sub^T @ dlogits
 *                      (dHsub must be zeroed by the caller; accumulated with atomics).
 * ---------------------------------------------------------------------------------------- */
int ppnp_topk_thresh(const float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t k,
                     float* thresh, void* stream);
int ppnp_topk_mask(float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld, const float* thresh,
                   void* stream);
int ppnp_dense_row_nnz(const float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld,
                       int32_t* row_nnz, void* stream);
int ppnp_dense_to_csr(const float* ppr, int64_t n_rows, int64_t n_cols, int64_t ld,
                      const int64_t* indptr, int32_t* indices, float* val, void* stream);
int ppnp_batch_support(const int64_t* indptr, const int32_t* indices, const int64_t* idx_batch,
                       int64_t B, uint8_t* mark, void* stream);
int ppnp_batch_support_colmap(const int64_t* indptr, const int32_t* indices, const int64_t* idx_batch,
                              int64_t B, int64_t n_cols, uint8_t* sel, int32_t* colmap, int32_t* m_out,
                              void* stream);
int ppnp_batch_propagate(const int64_t* indptr, const int32_t* indices, const float* val,
                         const int64_t* idx_batch, int64_t B, const int32_t* colmap,
                         const float* Hsub, int64_t ld_h, int32_t C, float* out, int64_t ld_out,
                         int32_t transpose, void* stream);

/* ------------------------------------------------------------------------------------------
 * (4b) Edge-stream plan, built on the GPU (csrc/plan_build.cu).  No counterpart in the reference: the stream is a
 *      re-encoding of the CSR that helpers.py:58-63 (calc_A_hat) produces, in a caller-chosen processing order of
 *      the rows.  Host mirror and specification: ppnp_b200/plan.py build_stream_plan / lane_transpose.
 *   ppnp_plan_workspace_bytes: scratch for ppnp_plan_measure over n_listed rows.
 *   ppnp_plan_measure: order[n_listed] (int64, nullable = natural order) lists the rows to stream.  Outputs, each
 *                      [n_listed + 1]: row_start (stream position of a row; last = edges in the stream), seg_first,
 *                      slot_first, fix_first (first segment / partial slot / cut-row index of a row; last = totals);
 *                      totals[5] (device) = {edges, segments, partial slots, cut rows, listed rows without an edge}.
 *                      The caller reads totals, allocates the plan arrays and calls
 *   ppnp_plan_fill   : cols[n_chunks * chunk_edges] (bit 31 = last edge of a segment; lane_group > 0: stored
 *                      lane-transposed, PPNP_PLAN_LANE_GROUP), out_vals (nullable with vals), seg_row[n_segs],
 *                      chunk_seg[n_chunks], fix_ptr[n_fix + 1], fix_row[n_fix], fix_deg[n_fix] (row_deg[row] when
 *                      row_deg is given, else the row's edge count).  Padding edges point at column 0.
 * ---------------------------------------------------------------------------------------- */
int64_t ppnp_plan_workspace_bytes(int64_t n_listed);
int ppnp_plan_measure(const int64_t* indptr, const int64_t* order, int64_t n_listed, int32_t chunk_edges,
                      int64_t* row_start, int32_t* seg_first, int32_t* slot_first, int32_t* fix_first,
                      int64_t* totals, void* workspace, int64_t workspace_bytes, void* stream);
int ppnp_plan_fill(const int64_t* indptr, const int32_t* indices, const float* vals, const int64_t* order,
                   int64_t n_listed, int32_t chunk_edges, int32_t lane_group, const int64_t* row_start,
                   const int32_t* seg_first, const int32_t* slot_first, const int32_t* fix_first,
                   const float* row_deg, int64_t nnz, int64_t n_chunks, int64_t n_segs, int64_t n_fix,
                   int32_t* cols, float* out_vals, int32_t* seg_row, int32_t* chunk_seg, int32_t* fix_ptr,
                   int32_t* fix_row, float* fix_deg, void* stream);

/* ------------------------------------------------------------------------------------------
 * (5) Partitioned propagation helper (BASELINE config 5; no counterpart in the single-process
 *     reference): dst[i, :] = src[idx[i], :] for i < n_rows.  src may be a peer GPU's buffer mapped
 *     over NVLink (halo pull) or the local Z (pack in front of an NCCL send).
 * ---------------------------------------------------------------------------------------- */
int ppnp_gather_rows(const float* src, int64_t ld_src, const int64_t* idx, int64_t n_rows, int32_t F,
                     float* dst, int64_t ld_dst, void* stream);

/* ------------------------------------------------------------------------------------------
 * Synthetic-workload plumbing (BASELINE.json configs 4/5): R-MAT raw draws e0..e1 of stream
 * `seed` (include/ppnp_rmat.h), written as 64-bit keys (src << 32 | dst), both directions,
 * loops and ids >= n replaced by the key -1.  out_keys has 2 * (e1 - e0) entries.
 * ---------------------------------------------------------------------------------------- */
int ppnp_rmat_keys(uint64_t seed, int32_t scale, int64_t n, int64_t e0, int64_t e1,
                   int64_t* out_keys, void* stream);

/* ------------------------------------------------------------------------------------------
 * (7) Graph standardisation -- the step right before the path.
 *     replaces ppnp/data/sparsegraph.py:191-222 SparseGraph.standardize for the unit-weight
 *     pipeline that main.py:75 / batch-main.py:76 run (make_unweighted=True):
 *       PPNP_STD_UNDIRECTED     to_undirected :127-148   pattern union of A and A^T
 *       PPNP_STD_NO_SELF_LOOPS  remove_self_loops :381-395
 *       PPNP_STD_LCC            largest_connected_components :355-379 + create_subgraph :300-352
 *                               (weak components; ties between largest components: the one whose
 *                               smallest node id is largest, = np.argsort(sizes)[::-1][0] of a stable sort)
 * in : CSR pattern of the raw adjacency, indptr int64[n+1], indices int32[nnz] (any order inside a
 *      row, duplicates allowed; stored weights are never read: every entry counts as 1).
 * out: canonical CSR of the standardised graph, bit-exact with the reference:
 *        out_indptr  int64[n+1]   (n_keep + 1 entries used)
 *        out_indices int32[cap]   cap = 2 * nnz with PPNP_STD_UNDIRECTED, else nnz
 *        out_keep    int32[n]     original ids of the kept nodes, ascending (n_keep entries used)
 *        out_counts  int64[3]     DEVICE: {n_keep, nnz_out, status}; status != 0: a column index
 *                                 was outside [0, n) and the result is invalid.
 * Everything is stream-ordered; the caller reads out_counts after synchronising.
 * workspace: ppnp_graph_standardize_workspace_bytes(n, nnz, flags) bytes (~34 B per key + 40 B per node).
 * ---------------------------------------------------------------------------------------- */
#define PPNP_STD_UNDIRECTED 1
#define PPNP_STD_NO_SELF_LOOPS 2
#define PPNP_STD_LCC 4
int64_t ppnp_graph_standardize_workspace_bytes(int64_t n, int64_t nnz, int32_t flags);
int ppnp_graph_standardize(const int64_t* indptr, const int32_t* indices, int64_t n, int64_t nnz, int32_t flags,
                           int64_t* out_indptr, int32_t* out_indices, int32_t* out_keep, int64_t* out_counts,
                           void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPNP_B200_H */
