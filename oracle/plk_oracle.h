/* oracle/plk_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain scalar C, fp64, no SIMD, no FMA contraction) of the arithmetic of
 * PhyML's likelihood hot path, on raw arrays with the reference's own layouts
 * (SURVEY.md Appendix A).  It is the checker for the CUDA engine: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  The product
 * (phyml_b200/) never links, imports or calls anything in this directory.
 *
 * Parity pinning: tests/test_oracle_golden.py checks every likelihood function below against arrays
 * dumped from the unmodified reference (oracle/ref_driver.c --dump -> tests/golden/NAME.npz), and
 * tests/test_pars_cpu.py the parsimony functions against the reference's Pars / Update_Partial_Pars
 * state (oracle/ref_driver.c --dump_pars -> tests/golden/pars/NAME.npz).
 *
 * Layouts:  CLV  [site][catg][state]      tip vector [site][state] (0/1 doubles)
 *           P    [catg][from][to]         scalers    int[site]
 */
#ifndef PLK_ORACLE_H
#define PLK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* One child of a CLV update / one side of an edge.
 * Internal side: clv != NULL (scale may be NULL => 0).  Tip side: clv == NULL and
 * tipvec/d_state/is_ambigu are the reference's p_lk_tip_r, c_seq->d_state, c_seq->is_ambigu. */
typedef struct
{
  const double *clv;       /* [npat][ncatg][ns] or NULL for a tip */
  const int    *scale;     /* [npat] or NULL */
  const double *tipvec;    /* [npat][ns], tips only */
  const short  *d_state;   /* [npat], tips only */
  const short  *is_ambigu; /* [npat], tips only */
} plk_oracle_side;

/* K0: transition matrices of one edge for all rate categories.
 * follows lk.c:2276-2325 (length clamp) and models.c:257-326,353-373 (PMat, PMat_Empirical). */
void plk_oracle_pmat(int ns, int ncatg, double l, const double *rates, double br_len_mult,
                     double l_min, double l_max, const double *U, const double *V,
                     const double *lambda, double *P);

/* K1: one conditional-likelihood update. follows lk.c:1659-1768 (Core_Default_Update_Partial_Lk)
 * and lk.c:3328-3406 (Partial_Lk_Inin/Exex/Exin). */
void plk_oracle_update_partial(int ns, int ncatg, int npat, const double *wght, int apply_scaling,
                               double *dst, int *dst_scale, const plk_oracle_side *c1,
                               const double *P1, const plk_oracle_side *c2, const double *P2);

/* K2: log-likelihood at an edge. follows lk.c:608-645 (site loop of Lk), lk.c:767-861 (Lk_Core),
 * lk.c:1185-1218 (one class), lk.c:1226-1273 (Invariant_Lk), lk.c:2777-2801 (SCALE_FAST).
 * outputs (any may be NULL): site_lnl=c_lnL_sorted, site_lk=cur_site_lk,
 * site_lk_cat=unscaled_site_lk_cat [npat][ncatg], fact_sum_scale [npat]. Returns lnL. */
double plk_oracle_edge_lnl(int ns, int ncatg, int npat, const double *wght, const short *invar,
                           int invar_flag, double pinv, const double *pi, const double *rate_probs,
                           const plk_oracle_side *left, const plk_oracle_side *rght, const double *P,
                           double *site_lnl, double *site_lk, double *site_lk_cat,
                           int *fact_sum_scale, int *numerical_warning);

/* K3: projection of both sides of an edge on the eigenbasis. follows lk.c:1038-1114. */
void plk_oracle_eigen_lr(int ns, int ncatg, int npat, const double *wght, const double *U,
                         const double *V, const double *pi, const plk_oracle_side *left,
                         const plk_oracle_side *rght, double *dot_prod);

/* K4: lnL and d lnL / d l at a trial length from dot_prod. follows lk.c:655-753 (dLk),
 * lk.c:955-1032 (Lk_dLk_Core_Eigen_Lr), lk.c:1170-1180. *l is clamped in place (lk.c:673-674).
 * with_derivative==0 follows the use_eigen_lr branch of Lk (lk.c:592-603, 866-950). */
double plk_oracle_lnl_dlnl(int ns, int ncatg, int npat, const double *wght, const short *invar,
                           int invar_flag, double pinv, const double *pi, const double *rates,
                           const double *rate_probs, double br_len_mult, double l_min, double l_max,
                           const double *lambda, const double *dot_prod, const int *fact_sum_scale,
                           double *l, int with_derivative, double *dlnl, int *numerical_warning);

/* Parsimony (src/pars.c).  Fitch: ui = state-set bit mask, pars = steps below the node, int[npat] each.
 * follows pars.c:374-388 (Update_Partial_Pars, Fitch branch). */
void plk_oracle_pars_update(int npat, int *ui, int *pars, const int *ui_v1, const int *pars_v1,
                            const int *ui_v2, const int *pars_v2);
/* follows pars.c:355-372 (general_pars branch): p_pars [npat][ns], step_mat [ns][ns]. */
void plk_oracle_pars_update_general(int ns, int npat, const int *step_mat, int *p_pars, const int *p_pars_v1,
                                    const int *p_pars_v2);
/* follows pars.c:20-51 (site loop of Pars) and pars.c:397-439 (Pars_Core); returns c_pars, fills site_pars. */
int plk_oracle_pars_edge(int general, int ns, int npat, const double *wght, const int *step_mat, const int *ui_l,
                         const int *pars_l, const int *p_pars_l, const int *ui_r, const int *pars_r,
                         const int *p_pars_r, int *site_pars);

#ifdef __cplusplus
}
#endif
#endif
