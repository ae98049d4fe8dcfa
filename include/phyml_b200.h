/* include/phyml_b200.h -- C ABI of the B200 likelihood engine (libphyml_b200.so).
 *
 * This is the drop-in boundary for PhyML's likelihood hot path (BASELINE.json north_star,
 * SURVEY.md section 8b).  PhyML has no plugin ABI: the boundary is the set of C functions in
 * src/lk.h that ~400 call sites link against.  A replacement lk.c keeps those symbols and
 * forwards to the entry points below (INTEGRATION.md shows the binding; the precedent is the
 * reference's own `#ifdef BEAGLE` hooks, src/beagle_utils.h:64-70, src/lk.c:585-587,1300-1302).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; all arrays passed in are HOST memory, copied by the call;
 *   - fp64 everywhere (phydbl is double, src/utilities.h:462), scalers are int (SCALE_FAST);
 *   - buffers are named by integer handles, like BEAGLE's indices (src/utilities.h:745-772
 *     `p_lk_left_idx`, `Pij_rr_idx`): the host may swap which edge owns which handle
 *     (Prune_Subtree/Graft_Subtree swap pointers, src/utilities.c:6247-6430) at no device cost;
 *   - layouts are the reference's: CLV [site][catg][state],
 *     P [catg][from][to], tip vector [site][state];
 *   - every function returns PLK_OK (0) or a negative error code; plk_last_error() gives the text.
 *     The reference's convention for a failing accelerator call is print + Exit()
 *     (src/beagle_utils.c:246-249); the shim does that with this text;
 *   - calls are asynchronous on the instance's CUDA stream; functions that return a scalar
 *     (plk_edge_lnl, plk_edge_lnl_dlnl, plk_edge_lnl_eigen) and the plk_get_* read-backs
 *     synchronise.  One instance per t_tree; an instance is not thread-safe (neither is PhyML).
 *   - upload calls (plk_set_*) are stream-ordered copies: pageable host arrays are consumed before
 *     the call returns; PINNED host arrays are read asynchronously and must stay unchanged until the
 *     next synchronising call;
 *   - there is NO CPU fallback: without a CUDA device plk_create fails with PLK_ERR_CUDA.
 */
#ifndef PHYML_B200_H
#define PHYML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLK_OK 0
#define PLK_ERR_ARG (-1)      /* bad argument / handle out of range */
#define PLK_ERR_CUDA (-2)     /* CUDA runtime error (text in plk_last_error) */
#define PLK_ERR_STATE (-3)    /* call order violated (e.g. dLk before eigen_lr) */
#define PLK_ERR_NCCL (-4)     /* NCCL error / NCCL not loadable */
#define PLK_ERR_UNSUPPORTED (-5)

typedef struct plk_instance plk_instance;

/* Sizes fixed at creation.  Reference: Make_Tree_For_Lk (src/make.c:17-291) knows n_otu,
 * n_pattern, ns, n_catg and the number of CLV / P-matrix buffers it carves from its arena
 * (src/make.c:96-104); the device instance is created right after it and destroyed in
 * Free_Tree_Lk (src/free.c:387-391), as BEAGLE's was at src/main.c:272,336. */
typedef struct plk_config
{
  int n_tips;     /* n_otu */
  int n_patterns; /* data->n_pattern owned by THIS instance (its shard when site-sharded) */
  int ns;         /* mod->ns: 4 and 20 have specialised fused kernels, other 2..32 run the generic kernel */
  int ncatg;      /* mod->ras->n_catg (1..16) */
  int n_clv;      /* number of CLV handles (2 per edge: p_lk_left, p_lk_rght); storage is lazy */
  int n_pmat;     /* number of P-matrix handles (1 per edge: Pij_rr) */
  int device;     /* CUDA device ordinal */
  int flags;      /* PLK_FLAG_* */
} plk_config;

#define PLK_FLAG_NO_SCALING 1 /* tree->apply_lk_scaling == NO */

/* One operand: a tip (tip >= 0, clv < 0) or an internal CLV buffer (clv >= 0, tip < 0). */
typedef struct plk_side
{
  int tip;
  int clv;
} plk_side;

/* One Update_Partial_Lk(tree,b,d) resolved by Set_All_Partial_Lk (src/lk.c:2922-3195). */
typedef struct plk_op
{
  int      dst;   /* CLV handle written (p_lk_left|p_lk_rght of b on d's side) + its scaler */
  plk_side c1;    /* n_v1 side: far-end CLV / tip of the first other edge of d */
  int      pmat1; /* Pij1 */
  plk_side c2;
  int      pmat2;
} plk_op;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int  plk_create(const plk_config *cfg, plk_instance **out);
void plk_destroy(plk_instance *inst);
const char *plk_last_error(const plk_instance *inst); /* inst may be NULL: last creation error */
int  plk_sync(plk_instance *inst);

/* ---- data upload (host -> device) ---------------------------------------------------------- */
/* data->wght[] (double, src/utilities.h:1155) and data->invar[] (short, >=0: constant state). */
int plk_set_pattern_weights(plk_instance *inst, const double *wght, const short *invar);
/* code -> 0/1 state vector table (n_codes x ns doubles, n_codes <= 256).  Replaces the tables of
 * Init_Tips_At_One_Site_{Nucleotides,AA}_Float (src/lk.c:26-69,122-161). */
int plk_set_tip_table(plk_instance *inst, int n_codes, const double *vectors);
/* one byte per pattern for tip `tip`, indexing the table above. */
int plk_set_tip_codes(plk_instance *inst, int tip, const uint8_t *codes);
/* all tips at once: row i (n_patterns bytes) of a [n_tips][host_stride] byte matrix is tip i. */
int plk_set_all_tip_codes(plk_instance *inst, const uint8_t *codes, size_t host_stride);
/* the same with 4-bit codes, two patterns per byte (low nibble = the even pattern): row i
 * ((n_patterns + 1) / 2 bytes) of a [n_tips][host_stride] byte matrix is tip i.  For tip tables of at
 * most 16 codes (nucleotide masks, src/lk.c:26-69): half the host-to-device bytes of a full upload. */
int plk_set_all_tip_codes_packed4(plk_instance *inst, const uint8_t *packed, size_t host_stride);
/* reference-format tip: the fp64 0/1 vectors b->p_lk_tip_r [n_patterns][ns] written by
 * Init_Partial_Lk_Tips_Double (src/lk.c:2060-2118); codes and table rows are derived here. */
int plk_set_tip_vectors(plk_instance *inst, int tip, const double *tip_vectors);
/* Results of Update_Eigen / Update_Efrq / Update_RAS / Update_Boundaries (src/lk.c:489-495):
 * U = eigen->r_e_vect [ns][ns], V = eigen->l_e_vect, lambda = eigen->e_val, pi = e_frq->pi->v,
 * rates = ras->gamma_rr->v, rate_probs = ras->gamma_r_proba->v, pinv = ras->pinvar->v,
 * invar_flag = ras->invar, l_min/l_max = mod->l_min/l_max, br_len_mult = mod->br_len_mult->v. */
int plk_set_model(plk_instance *inst, const double *U, const double *V, const double *lambda,
                  const double *pi, const double *rates, const double *rate_probs, double pinv,
                  int invar_flag, double l_min, double l_max, double br_len_mult);

/* ---- K0: transition matrices ----------------------------------------------------------------
 * replaces Update_PMat_At_Given_Edge (src/lk.c:2238-2325) + PMat/PMat_Empirical
 * (src/models.c:353-373, 257-326) for n edges at once: P = U diag(exp(lambda len_c)) V per rate
 * category, floored at 1e-100, rows renormalised.  l[i] is b->l->v (unclamped). */
int plk_update_pmats(plk_instance *inst, int n, const int *pmat, const double *l);
/* host-computed P (has_zero_br_len identity, PMat_MGF_Gamma ...): [ncatg][ns][ns] */
int plk_set_pmat(plk_instance *inst, int pmat, const double *P);
int plk_get_pmat(plk_instance *inst, int pmat, double *P);

/* ---- K1: conditional-likelihood updates -----------------------------------------------------
 * replaces the compute of Update_Partial_Lk (src/lk.c:1282) = AVX_Update_Partial_Lk
 * (src/avx.c:301-522) / SSE_ (src/sse.c:254) / Core_Default_ (src/lk.c:1659-1768).
 * `ops` must be in dependency order (post-order / pre-order as the host recursion issues them);
 * independent ops are batched into one launch per dependency level. */
int plk_update_partials(plk_instance *inst, int n_ops, const plk_op *ops);

/* ---- K2: log-likelihood at an edge ----------------------------------------------------------
 * replaces the site loop of Lk (src/lk.c:605-645) + Lk_Core (:767-861) +
 * Pull_Scaling_Factors (:2696-2803) + Invariant_Lk (:1226-1273).  Per-site by-products
 * (c_lnL_sorted, cur_site_lk, unscaled_site_lk_cat, fact_sum_scale) stay on the device until
 * plk_get_site_lnl.  *lnl is THIS instance's partial sum over its patterns.
 * *numerical_warning (may be NULL) mirrors tree->numerical_warning. */
int plk_edge_lnl(plk_instance *inst, plk_side left, plk_side rght, int pmat, double *lnl,
                 int *numerical_warning);

/* ---- K1 + K2: a traversal and the edge reduction in one call --------------------------------
 * replaces the body of Lk(NULL) after the P-matrix loop: Post_Order_Lk (src/lk.c:562-564) followed by
 * the site loop at the root edge (src/lk.c:578-645).  Same results and by-products as
 * plk_update_partials + plk_edge_lnl; for 4-state data with 4 rate categories the edge reduction
 * runs as the epilogue of the traversal kernel (one launch per full-tree evaluation). */
int plk_traverse_edge_lnl(plk_instance *inst, int n_ops, const plk_op *ops, plk_side left,
                          plk_side rght, int pmat, double *lnl, int *numerical_warning);

/* the whole of Lk(NULL) after the host's model update (Update_RAS / Update_Efrq / Update_Eigen,
 * src/lk.c:489-495) as one call: the P-matrix loop over all edges (src/lk.c:500-505), Post_Order_Lk
 * (:562-564) and the site loop at the root edge (:578-645) = plk_update_pmats + plk_traverse_edge_lnl. */
int plk_lk_full(plk_instance *inst, int n_pmat, const int *pmat, const double *lengths, int n_ops,
                const plk_op *ops, plk_side left, plk_side rght, int edge_pmat, double *lnl,
                int *numerical_warning);

/* plk_lk_full split in two for hosts that pipeline evaluations: _begin enqueues everything and returns, _wait returns
 * the result.  Between the two only uploads (plk_set_*) may be issued: they are ordered behind the evaluation on the
 * device, and the host-to-device copy of plk_set_all_tip_codes_packed4 overlaps it (own copy stream). */
int plk_lk_full_begin(plk_instance *inst, int n_pmat, const int *pmat, const double *lengths, int n_ops,
                      const plk_op *ops, plk_side left, plk_side rght, int edge_pmat);
int plk_lk_wait(plk_instance *inst, double *lnl, int *numerical_warning);

/* ---- K3: eigen-basis projection -------------------------------------------------------------
 * replaces Update_Eigen_Lr (src/lk.c:1038-1114, src/avx.c:21-105): tree->dot_prod stays on the
 * device; also latches fact_sum_scale = scale(left)+scale(rght) for K4. */
int plk_eigen_lr(plk_instance *inst, plk_side left, plk_side rght);

/* ---- K4: lnL and d lnL / dl at a trial length ------------------------------------------------
 * replaces dLk (src/lk.c:655-753) + Lk_dLk_Core_Eigen_Lr (:955-1032).  *l is clamped to
 * [l_min,l_max] in place like the reference (lk.c:673-674). */
int plk_edge_lnl_dlnl(plk_instance *inst, double *l, double *lnl, double *dlnl,
                      int *numerical_warning);
/* the use_eigen_lr == YES branch of Lk (src/lk.c:592-603,625-629 + Lk_Core_Eigen_Lr :866-950) */
int plk_edge_lnl_eigen(plk_instance *inst, double l, double *lnl, int *numerical_warning);

/* ---- read-backs (device -> host), for host code that reads engine state ----------------------
 * (io.c:Print_Site_Lk, alrt.c, ancestral.c, Optimiz_Alpha_And_Pinv ...) and for tests. */
int plk_get_clv(plk_instance *inst, int clv, double *clv_out, int *scale_out);
int plk_set_clv(plk_instance *inst, int clv, const double *clv_in, const int *scale_in);
/* any pointer may be NULL */
int plk_get_site_lnl(plk_instance *inst, double *site_lnl, double *site_lk, double *site_lk_cat,
                     int *fact_sum_scale);
int plk_get_dot_prod(plk_instance *inst, double *dot_prod);

/* ---- site sharding across GPUs (one process per GPU) ------------------------------------------
 * Each rank creates an instance over its own block of patterns; the only exchange is one
 * ncclAllReduce(sum) of 1 (lnL) or 2 (lnL, dlnL) doubles per evaluation, enqueued on the
 * instance stream right behind the reduction kernel.  NCCL is dlopen()ed on first use.
 * plk_comm_unique_id: rank 0 fills a 128-byte id that the host broadcasts to the other ranks. */
int plk_comm_unique_id(void *id128);
int plk_comm_init(plk_instance *inst, int rank, int world, const void *id128);
/* Fused alternative (default of bench.py): the reduction kernels themselves post their partial sums into
 * every rank's mailbox over NVLink peer memory (CUDA IPC) and add all ranks' partials in rank order --
 * no extra launch, bitwise identical result on every rank.  Each rank exports a 64-byte IPC handle; the
 * host gathers the `world` handles (rank order) and hands the array to every rank. */
int plk_comm_p2p_export(plk_instance *inst, int world, void *handle64);
int plk_comm_p2p_init(plk_instance *inst, int rank, int world, const void *handles);
/* after plk_comm_init, plk_edge_lnl / _dlnl / _eigen return the ALL-RANK sum when enabled */
int plk_comm_set_allreduce(plk_instance *inst, int enable);

/* ---- site sharding across GPUs inside ONE process (the reference is one process, one t_tree) -----
 * SURVEY.md section 8(e) "one process, 8 devices": cfg->n_patterns is the whole alignment
 * (cfg->device is ignored); device devices[i] (NULL: 0..n_gpus-1) owns the i-th contiguous block of
 * patterns of every per-site buffer, P-matrices and the model are replicated.  The returned instance
 * takes every call of this header with whole-alignment arrays; scalar results are the all-shard
 * sums (the partial sums are exchanged inside the reduction kernels over NVLink peer memory when
 * the devices can map each other, else added on the host in shard order).  This is what lets
 * lk.c's single t_tree (src/lk.c:443, src/make.c:17) drive all GPUs of a box.
 * The same device may be listed several times (shards then share it; used by the 1-GPU tests). */
int plk_create_sharded(const plk_config *cfg, int n_gpus, const int *devices, plk_instance **out);
int plk_n_shards(const plk_instance *inst);

/* ---- batched SPR candidates ------------------------------------------
 * Test_One_Spr_Target (src/spr.c:589-650) scores ONE regraft position of a pruned subtree: Graft_Subtree,
 * Update_PMat_At_Given_Edge on the two halves of the target edge, Update_Partial_Lk(tree, b_arrow, n_link),
 * Lk(b_arrow).  Once the subtree is pruned and the remaining tree's CLVs are up to date on both sides, every such
 * score depends only on existing buffers, so the scores of ALL regraft positions come from one call:
 *   a, l_a : CLV (or tip) seen from one end of the target edge, looking away from the regraft point, and the
 *            length of that half after Graft_Subtree (b_target->l->v);  b, l_b : the other end (b_residual);
 *   prune, l_prune : the pruned subtree's CLV (or tip) and the length of the edge that carries it (b_arrow);
 *   link_on_left   : non-zero when the new node n_link is b_arrow->left (always the case when `prune` is a tip).
 * lnl[i] = what Lk(b_arrow) would return for candidate i (the all-shard sum on a sharded instance); the CLV of
 * n_link is never written and no per-site by-product is stored.  Using it needs a caller that enumerates the
 * regraft positions itself (Test_One_Spr_Target_Recur, src/spr.c:525-586, interleaves enumeration and scoring),
 * i.e. it is outside the symbol-for-symbol drop-in; INTEGRATION.md section 6 shows the change to spr.c. */
typedef struct plk_spr_cand
{
  plk_side a;
  double   l_a;
  plk_side b;
  double   l_b;
} plk_spr_cand;
int plk_spr_candidates(plk_instance *inst, plk_side prune, double l_prune, int link_on_left, int n_cand,
                       const plk_spr_cand *cand, double *lnl, int *numerical_warning);

/* ---- parsimony: the SPR pre-filter (src/pars.c) ----------------------
 * Every edge side owns a parsimony buffer (b->ui_l/pars_l/p_pars_l and ..._r, src/make.c:455-475; the host
 * swaps the three pointers together in Prune_Subtree/Graft_Subtree, src/utilities.c:6268-6278), named here by
 * an integer handle.  A buffer holds the Fitch state set `ui` (bit j = state j) and step count `pars` per
 * pattern and, for the step-matrix variant (`general_pars`, src/init.c:777), p_pars[pattern][state].
 * Tip buffers are uploaded once (Init_Ui_Tips / Init_Partial_Pars_Tips, src/pars.c:111-233).  All results are
 * integers and bit-identical to the reference's. */
typedef struct plk_pars_op
{
  int dst; /* buffer written: ui/pars (or p_pars) of b_fcus on n's side (src/pars.c:271-351) */
  int c1;  /* far-end buffers of the two other edges of n */
  int c2;
} plk_pars_op;
/* allocate the handle table; step_mat = tree->step_mat [ns][ns] (Get_Step_Mat, src/pars.c:498) or NULL when
 * only Fitch parsimony is used.  Replaces Make_Tree_For_Pars' buffers (src/make.c:334-372) on the device. */
int plk_pars_create(plk_instance *inst, int n_buffers, const int *step_mat);
/* upload / read back one buffer (host layouts of the reference: int[n_patterns], p_pars int[n_patterns][ns]);
 * ui and pars go together, p_pars may be NULL (and vice versa). */
int plk_pars_set_buffer(plk_instance *inst, int buf, const int *ui, const int *pars, const int *p_pars);
int plk_pars_get_buffer(plk_instance *inst, int buf, int *ui, int *pars, int *p_pars);
/* replaces Update_Partial_Pars (src/pars.c:239-391) for a dependency-ordered list of updates (the order of
 * Post_Order_Pars / Pre_Order_Pars, src/pars.c:56-93) in ONE launch.  general != 0: the step-matrix branch. */
int plk_pars_update(plk_instance *inst, int general, int n_ops, const plk_pars_op *ops);
/* replaces the site loop of Pars (src/pars.c:40-48) + Pars_Core (src/pars.c:397-439) at the edge whose two
 * sides are `left` and `rght`: *c_pars = tree->c_pars over this instance's patterns (the all-shard total on a
 * plk_create_sharded instance; like lnL, the all-rank total once the instance is wired into a cross-GPU exchange);
 * site_pars stays on the device. */
int plk_pars_edge(plk_instance *inst, int general, int left, int rght, int *c_pars);
/* the updates and the site loop in one launch: the whole of Pars(NULL) (n - 2 or 3n - 6 updates), or the one
 * update + Pars(b) of an SPR candidate scored by parsimony (src/spr.c:636-640). */
int plk_pars_traverse_edge(plk_instance *inst, int general, int n_ops, const plk_pars_op *ops, int left, int rght,
                           int *c_pars);
/* tree->site_pars of the last plk_pars_edge / plk_pars_traverse_edge */
int plk_get_site_pars(plk_instance *inst, int *site_pars);

/* ---- introspection ---------------------------------------------------------------------------- */
/* kernels launched so far by this instance (bench.py's gpu_launches) */
long long plk_launch_count(const plk_instance *inst);
/* device bytes currently allocated by the instance */
size_t plk_device_bytes(const plk_instance *inst);
/* CUDA stream of the instance as an opaque pointer (cudaStream_t), for event timing */
void *plk_stream(plk_instance *inst);
const char *plk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PHYML_B200_H */
