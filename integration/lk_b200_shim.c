/* integration/lk_b200_shim.c -- the reference-side binding of the B200 likelihood engine.
 *
 * PhyML has no plugin ABI: the boundary is the set of C functions of src/lk.h that ~400 call sites
 * link against.  This file re-defines exactly those entry points with the
 * reference's own signatures and forwards them to the C ABI of include/phyml_b200.h:
 *
 *     Lk                          src/lk.c:443    -> plk_set_model / plk_update_pmats / plk_update_partials / plk_edge_lnl
 *     dLk                         src/lk.c:655    -> plk_edge_lnl_dlnl
 *     Update_Partial_Lk           src/lk.c:1282   -> queued plk_op, flushed as ONE plk_update_partials
 *     Update_PMat_At_Given_Edge   src/lk.c:2238   -> plk_update_pmats
 *     Update_Eigen_Lr             src/lk.c:1038   -> plk_eigen_lr
 *     Make_Tree_For_Lk            src/make.c:17   -> original, then plk_create + uploads
 *     Free_Tree_Lk                src/free.c:387  -> plk_destroy, then original
 *
 * It is linked into an executable together with the UNMODIFIED reference built as a shared library
 * (oracle/_ref/libphyml_ref.so + main.o): definitions in the executable take precedence over the
 * library's for every caller, including calls made from inside the library (spr.c, optimiz.c, the
 * Post_Order_Lk / Pre_Order_Lk recursions of lk.c), so the SPR search and the branch-length
 * optimiser drive the GPU engine unchanged.  The originals stay reachable through
 * dlsym(RTLD_NEXT, ...) for the lifecycle hooks.  With source access the same bodies go under
 * `#ifdef PHYML_B200` at the `#ifdef BEAGLE` sites of lk.c.
 *
 * Handles: the host swaps CLV *pointers* between edges (Prune_Subtree / Graft_Subtree,
 * src/utilities.c:6247-6430), so device buffers are keyed by the host pointer value
 * (p_lk_left / p_lk_rght / Pij_rr), never by edge number.  Tips are keyed by node number.
 *
 * Unsupported configurations abort like the BEAGLE hooks did (src/main.c:240-253): mixture trees
 * (is_mixt_tree), rooted trees (n_root), SCALE_RATE_SPECIFIC, M4, gamma_mgf_bl, ns > 32.
 * Host-side readers of engine state: c_lnL_sorted / cur_site_lk / unscaled_site_lk_cat /
 * fact_sum_scale are mirrored after every Lk(NULL); CLVs and P-matrices stay on the device.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "utilities.h"
#include "lk.h"
#include "models.h"
#include "optimiz.h"
#include "make.h"
#include "free.h"

#include "../include/phyml_b200.h"

/* ------------------------------------------------------------------------------------------------ */
typedef struct
{
  const void *key;
  int         val;
} slot_t;

typedef struct
{
  t_tree       *tree;
  plk_instance *inst;
  slot_t       *clv_map, *pm_map;
  int           map_cap, n_clv, n_pm, clv_cap, pm_cap;
  plk_op       *queue;
  int           n_queue, queue_cap;
  int           pm_h[64];  /* deferred P-matrix updates: handle, branch length */
  double        pm_l[64];
  int           n_pm_queue;
  double        model_print[8 + 2 * 32 + 2 * 16]; /* fingerprint of the uploaded model */
  int           model_valid;
  long long     n_lk, n_dlk, n_partial, n_flush, n_pmat;
} shim_t;

#define MAX_SHIMS 16
static shim_t g_shims[MAX_SHIMS];

static void die(const char *what, plk_instance *inst)
{
  PhyML_Fprintf(stderr, "\n. phyml_b200: %s: %s\n", what, plk_last_error(inst));
  Exit("\n");
}
#define CK(call, sh)                         \
  do                                         \
  {                                          \
    if ((call) != PLK_OK) die(#call, (sh)->inst); \
  } while (0)

static shim_t *shim_of(t_tree *tree)
{
  int i;
  for (i = 0; i < MAX_SHIMS; ++i)
    if (g_shims[i].tree == tree) return &g_shims[i];
  return NULL;
}

static int map_get(slot_t *map, int cap, const void *key, int *counter, int limit, const char *what)
{
  uintptr_t h = ((uintptr_t)key >> 4) * 0x9E3779B97F4A7C15ULL;
  int       i = (int)(h % (uintptr_t)cap);
  while (map[i].key && map[i].key != key) i = (i + 1) % cap;
  if (!map[i].key)
  {
    if (*counter >= limit)
    {
      PhyML_Fprintf(stderr, "\n. phyml_b200: out of %s handles (%d)\n", what, limit);
      Exit("\n");
    }
    map[i].key = key;
    map[i].val = (*counter)++;
  }
  return map[i].val;
}
static int clv_handle(shim_t *sh, const phydbl *p) { return map_get(sh->clv_map, sh->map_cap, p, &sh->n_clv, sh->clv_cap, "CLV"); }
static int pm_handle(shim_t *sh, const phydbl *p) { return map_get(sh->pm_map, sh->map_cap, p, &sh->n_pm, sh->pm_cap, "P-matrix"); }

/* ------------------------------------------------------------------------------------------------ */
static void upload_model_if_changed(shim_t *sh)
{
  t_mod *mod = sh->tree->mod;
  double fp[8 + 2 * 32 + 2 * 16];
  int    ns = mod->ns, nc = mod->ras->n_catg, k = 0, i;
  memset(fp, 0, sizeof(fp));
  fp[k++] = mod->ras->pinvar->v;
  fp[k++] = (double)mod->ras->invar;
  fp[k++] = mod->l_min;
  fp[k++] = mod->l_max;
  fp[k++] = mod->br_len_mult->v;
  k = 8;
  for (i = 0; i < ns; ++i) fp[k++] = mod->eigen->e_val[i];
  for (i = 0; i < ns; ++i) fp[k++] = mod->e_frq->pi->v[i];
  k = 8 + 64;
  for (i = 0; i < nc; ++i) fp[k++] = mod->ras->gamma_rr->v[i];
  for (i = 0; i < nc; ++i) fp[k++] = mod->ras->gamma_r_proba->v[i];
  /* eigenvectors change only together with the eigenvalues (Update_Eigen, models.c:881) plus U[0..] as a guard */
  fp[5] = mod->eigen->r_e_vect[1];
  fp[6] = mod->eigen->l_e_vect[ns > 1 ? ns : 0];
  if (sh->model_valid && !memcmp(fp, sh->model_print, sizeof(fp))) return;
  if (sh->n_pm_queue > 0)
  { /* P-matrices queued under the previous parameter values are computed with them, as the reference did */
    CK(plk_update_pmats(sh->inst, sh->n_pm_queue, sh->pm_h, sh->pm_l), sh);
    sh->n_pm_queue = 0;
  }
  CK(plk_set_model(sh->inst, mod->eigen->r_e_vect, mod->eigen->l_e_vect, mod->eigen->e_val, mod->e_frq->pi->v,
                   mod->ras->gamma_rr->v, mod->ras->gamma_r_proba->v, mod->ras->pinvar->v, mod->ras->invar,
                   mod->l_min, mod->l_max, mod->br_len_mult->v),
     sh);
  memcpy(sh->model_print, fp, sizeof(fp));
  sh->model_valid = 1;
}

/* Deferred work is executed as: all queued P-matrix updates in ONE batched launch, then all queued CLV
   updates in ONE fused launch.  That order is valid because a P-matrix update is only queued when no
   CLV update queued before it reads that matrix (otherwise everything is flushed first). */
static void flush(shim_t *sh)
{
  if (sh->n_pm_queue > 0)
  {
    CK(plk_update_pmats(sh->inst, sh->n_pm_queue, sh->pm_h, sh->pm_l), sh);
    sh->n_pm_queue = 0;
  }
  if (sh->n_queue == 0) return;
  CK(plk_update_partials(sh->inst, sh->n_queue, sh->queue), sh);
  sh->n_queue = 0;
  sh->n_flush++;
}

static void check_supported(t_tree *tree)
{
  if (tree->is_mixt_tree == YES || tree->n_root != NULL || tree->mod->use_m4mod == YES ||
      tree->mod->gamma_mgf_bl == YES || tree->scaling_method != SCALE_FAST || tree->mod->ns > 32 ||
      (tree->io && tree->io->do_alias_subpatt == YES))
  {
    PhyML_Fprintf(stderr, "\n. phyml_b200: unsupported configuration (mixture / rooted tree / M4 / MGF branch lengths /"
                          "\n. rate-specific scaling / sub-pattern aliasing): use the CPU build.\n");
    Exit("\n");
  }
}

static plk_side side_of(shim_t *sh, t_node *n, const phydbl *p_lk)
{
  plk_side s;
  if (n->tax)
  {
    s.tip = n->num;
    s.clv = -1;
  }
  else
  {
    s.tip = -1;
    s.clv = clv_handle(sh, p_lk);
  }
  return s;
}

/* ------------------------------------------------------------------------------------------------ */
/* lifecycle: src/make.c:17-291, src/free.c:387-391 (BEAGLE did the same at src/main.c:272,336)     */
void Make_Tree_For_Lk(t_tree *tree)
{
  static void (*orig)(t_tree *) = NULL;
  shim_t     *sh = NULL;
  plk_config  cfg;
  int         i;
  if (!orig) orig = (void (*)(t_tree *))dlsym(RTLD_NEXT, "Make_Tree_For_Lk");
  orig(tree);
  if (tree->is_mixt_tree == YES) return;
  check_supported(tree);
  for (i = 0; i < MAX_SHIMS; ++i)
    if (!g_shims[i].tree)
    {
      sh = &g_shims[i];
      break;
    }
  if (!sh)
  {
    PhyML_Fprintf(stderr, "\n. phyml_b200: too many live trees\n");
    Exit("\n");
  }
  memset(sh, 0, sizeof(*sh));
  sh->tree = tree;
  sh->clv_cap = 8 * tree->n_otu + 16;
  sh->pm_cap = 4 * tree->n_otu + 16;
  sh->map_cap = 4 * (sh->clv_cap + sh->pm_cap) + 7;
  sh->clv_map = (slot_t *)calloc(sh->map_cap, sizeof(slot_t));
  sh->pm_map = (slot_t *)calloc(sh->map_cap, sizeof(slot_t));
  sh->queue_cap = 8 * tree->n_otu + 64;
  sh->queue = (plk_op *)malloc(sizeof(plk_op) * sh->queue_cap);

  memset(&cfg, 0, sizeof(cfg));
  cfg.n_tips = tree->n_otu;
  cfg.n_patterns = tree->data->n_pattern;
  cfg.ns = tree->mod->ns;
  cfg.ncatg = tree->mod->ras->n_catg;
  cfg.n_clv = sh->clv_cap;
  cfg.n_pmat = sh->pm_cap;
  cfg.device = getenv("PLK_DEVICE") ? atoi(getenv("PLK_DEVICE")) : 0;
  cfg.flags = (tree->apply_lk_scaling == YES) ? 0 : PLK_FLAG_NO_SCALING;
  if (plk_create(&cfg, &sh->inst) != PLK_OK)
  {
    PhyML_Fprintf(stderr, "\n. phyml_b200: plk_create failed: %s\n", plk_last_error(NULL));
    Exit("\n");
  }
  CK(plk_set_pattern_weights(sh->inst, tree->data->wght, tree->data->invar), sh);
  /* tips: the fp64 0/1 vectors written by Init_Partial_Lk_Tips_Double (make.c:264) */
  for (i = 0; i < tree->n_otu; ++i)
  {
    t_node *tip = tree->a_nodes[i];
    t_edge *b = tip->b[0];
    CK(plk_set_tip_vectors(sh->inst, tip->num, (b->rght == tip) ? b->p_lk_tip_r : b->p_lk_tip_l), sh);
  }
  if (getenv("PLK_SHIM_VERBOSE"))
    PhyML_Printf("\n. phyml_b200: %s, instance for %d taxa x %d patterns, ns=%d ncatg=%d", plk_version(), tree->n_otu,
                 tree->data->n_pattern, tree->mod->ns, tree->mod->ras->n_catg);
}

void Free_Tree_Lk(t_tree *tree)
{
  static void (*orig)(t_tree *) = NULL;
  shim_t *sh = shim_of(tree);
  if (!orig) orig = (void (*)(t_tree *))dlsym(RTLD_NEXT, "Free_Tree_Lk");
  if (sh)
  {
    if (getenv("PLK_SHIM_VERBOSE"))
      PhyML_Printf("\n. phyml_b200: Lk %lld  dLk %lld  Update_Partial_Lk %lld (in %lld launches)  Update_PMat %lld  "
                   "kernels %lld\n",
                   sh->n_lk, sh->n_dlk, sh->n_partial, sh->n_flush, sh->n_pmat, plk_launch_count(sh->inst));
    plk_destroy(sh->inst);
    free(sh->clv_map);
    free(sh->pm_map);
    free(sh->queue);
    memset(sh, 0, sizeof(*sh));
  }
  orig(tree);
}

/* ------------------------------------------------------------------------------------------------ */
/* src/lk.c:2238-2325                                                                               */
void Update_PMat_At_Given_Edge(t_edge *b_fcus, t_tree *tree)
{
  shim_t *sh = shim_of(tree);
  int     h;
  double  l;
  if (!sh)
  { /* trees without an instance (e.g. distance-based starting trees) keep the CPU path */
    static void (*orig)(t_edge *, t_tree *) = NULL;
    if (!orig) orig = (void (*)(t_edge *, t_tree *))dlsym(RTLD_NEXT, "Update_PMat_At_Given_Edge");
    orig(b_fcus, tree);
    return;
  }
  assert(b_fcus && b_fcus->Pij_rr);
  h = pm_handle(sh, b_fcus->Pij_rr);
  sh->n_pmat++;
  { /* a queued CLV update must see this matrix as it was when Update_Partial_Lk was called */
    int i, conflict = 0;
    for (i = 0; i < sh->n_queue; ++i)
      if (sh->queue[i].pmat1 == h || sh->queue[i].pmat2 == h) conflict = 1;
    if (conflict || sh->n_pm_queue == 64) flush(sh);
  }
  if (b_fcus->has_zero_br_len == YES)
  { /* identity matrices (PMat_Zero_Br_Len, models.c:331): host-computed, uploaded */
    int     ns = tree->mod->ns, nc = tree->mod->ras->n_catg, c, i;
    double *P = (double *)calloc((size_t)nc * ns * ns, sizeof(double));
    flush(sh);
    for (c = 0; c < nc; ++c)
      for (i = 0; i < ns; ++i) P[(size_t)c * ns * ns + i * ns + i] = 1.0;
    CK(plk_set_pmat(sh->inst, h, P), sh);
    free(P);
    return;
  }
  upload_model_if_changed(sh);
  l = (tree->mod->log_l == YES) ? exp(b_fcus->l->v) : b_fcus->l->v; /* lk.c:2278 */
  { /* defer: the latest length of a handle wins */
    int i;
    for (i = 0; i < sh->n_pm_queue; ++i)
      if (sh->pm_h[i] == h)
      {
        sh->pm_l[i] = l;
        return;
      }
    sh->pm_h[sh->n_pm_queue] = h;
    sh->pm_l[sh->n_pm_queue] = l;
    sh->n_pm_queue++;
  }
}

/* ------------------------------------------------------------------------------------------------ */
/* src/lk.c:1282-1325: flags, then the operands resolved by the reference's own Set_All_Partial_Lk   */
void Update_Partial_Lk(t_tree *tree, t_edge *b, t_node *d)
{
  shim_t *sh = shim_of(tree);
  t_node *n_v1 = NULL, *n_v2 = NULL;
  phydbl *p_lk = NULL, *p_lk_v1 = NULL, *p_lk_v2 = NULL, *Pij1 = NULL, *Pij2 = NULL, *tPij1 = NULL, *tPij2 = NULL;
  int    *sum_scale = NULL, *sum_scale_v1 = NULL, *sum_scale_v2 = NULL, *p_lk_loc = NULL;
  plk_op *op;
  if (!sh)
  {
    static void (*orig)(t_tree *, t_edge *, t_node *) = NULL;
    if (!orig) orig = (void (*)(t_tree *, t_edge *, t_node *))dlsym(RTLD_NEXT, "Update_Partial_Lk");
    orig(tree, b, d);
    return;
  }
  if (b->left == d && b->update_partial_lk_left == NO) return;
  if (b->rght == d && b->update_partial_lk_rght == NO) return;
  if (d->tax) return;
  Set_All_Partial_Lk(&n_v1, &n_v2, &p_lk, &sum_scale, &p_lk_loc, &Pij1, &tPij1, &p_lk_v1, &sum_scale_v1, &Pij2,
                     &tPij2, &p_lk_v2, &sum_scale_v2, d, b, tree);
  if (sh->n_queue == sh->queue_cap) flush(sh);
  op = &sh->queue[sh->n_queue++];
  op->dst = clv_handle(sh, p_lk);
  op->c1 = side_of(sh, n_v1, p_lk_v1);
  op->pmat1 = pm_handle(sh, Pij1);
  op->c2 = side_of(sh, n_v2, p_lk_v2);
  op->pmat2 = pm_handle(sh, Pij2);
  sh->n_partial++;
}

/* ------------------------------------------------------------------------------------------------ */
/* src/lk.c:1038-1114                                                                               */
void Update_Eigen_Lr(t_edge *b, t_tree *tree)
{
  shim_t *sh = shim_of(tree);
  if (!sh)
  {
    static void (*orig)(t_edge *, t_tree *) = NULL;
    if (!orig) orig = (void (*)(t_edge *, t_tree *))dlsym(RTLD_NEXT, "Update_Eigen_Lr");
    orig(b, tree);
    return;
  }
  flush(sh);
  upload_model_if_changed(sh);
  CK(plk_eigen_lr(sh->inst, side_of(sh, b->left, b->p_lk_left), side_of(sh, b->rght, b->p_lk_rght)), sh);
}

/* ------------------------------------------------------------------------------------------------ */
/* src/lk.c:443-649                                                                                 */
phydbl Lk(t_edge *b, t_tree *tree)
{
  shim_t      *sh = shim_of(tree);
  unsigned int br;
  int          warn = 0;
  double       lnl = 0.0;
  const int    full = (b == NULL);
  if (!sh)
  {
    static phydbl (*orig)(t_edge *, t_tree *) = NULL;
    if (!orig) orig = (phydbl(*)(t_edge *, t_tree *))dlsym(RTLD_NEXT, "Lk");
    return orig(b, tree);
  }
  tree->numerical_warning = NO;
  if (b == NULL && tree->mod->s_opt->curr_opt_free_rates == YES)
  { /* lk.c:458-463 */
    tree->mod->s_opt->curr_opt_free_rates = NO;
    Optimize_Free_Rate_Weights(tree, YES, YES);
    tree->mod->s_opt->curr_opt_free_rates = YES;
  }
  tree->old_lnL = tree->c_lnL;
  if (tree->rates && tree->io && tree->io->lk_approx == NORMAL)
  {
    PhyML_Fprintf(stderr, "\n. phyml_b200: Lk_Normal_Approx is not supported.\n");
    Exit("\n");
  }

  if (b == NULL)
  { /* lk.c:489-495: host model maths stay on the host */
    Update_Boundaries(tree->mod);
    Update_RAS(tree->mod);
    Update_Efrq(tree->mod);
    Update_Eigen(tree->mod);
  }
  upload_model_if_changed(sh);

  /* skip_tree_traversal (Optimiz_Alpha_And_Pinv, optimiz.c:2215-2222) only saves work in the
     reference; recomputing gives the same value, so it is ignored here */
  if (!b)
  { /* lk.c:500-505: all P-matrices in one batched launch */
    flush(sh);
    const int n = 2 * tree->n_otu - 3;
    int      *h = (int *)malloc(sizeof(int) * n);
    double   *l = (double *)malloc(sizeof(double) * n);
    int       n_batched = 0;
    for (br = 0; br < (unsigned int)n; ++br)
    {
      t_edge *e = tree->a_edges[br];
      if (e->has_zero_br_len == YES)
        Update_PMat_At_Given_Edge(e, tree);
      else
      {
        h[n_batched] = pm_handle(sh, e->Pij_rr);
        l[n_batched] = (tree->mod->log_l == YES) ? exp(e->l->v) : e->l->v;
        n_batched++;
      }
    }
    CK(plk_update_pmats(sh->inst, n_batched, h, l), sh);
    sh->n_pmat += n_batched;
    free(h);
    free(l);
    /* lk.c:560-565: the reference's own recursion; every visit lands in Update_Partial_Lk above */
    Post_Order_Lk(tree->a_nodes[tree->tip_root], tree->a_nodes[tree->tip_root]->v[0], tree);
    if (tree->both_sides == YES) Pre_Order_Lk(tree->a_nodes[tree->tip_root], tree->a_nodes[tree->tip_root]->v[0], tree);
    b = tree->a_nodes[tree->tip_root]->b[0]; /* lk.c:578-579 */
  }
  else if (tree->use_eigen_lr == NO)
    Update_PMat_At_Given_Edge(b, tree); /* lk.c:515-527 */

  tree->c_lnL = .0;
  tree->sum_min_sum_scale = .0;
  flush(sh);
  if (tree->update_eigen_lr == YES) Update_Eigen_Lr(b, tree); /* lk.c:590 */
  if (tree->use_eigen_lr == YES)
    CK(plk_edge_lnl_eigen(sh->inst, b->l->v, &lnl, &warn), sh); /* lk.c:592-603,625-629 */
  else
    CK(plk_edge_lnl(sh->inst, side_of(sh, b->left, b->p_lk_left), side_of(sh, b->rght, b->p_lk_rght),
                    pm_handle(sh, b->Pij_rr), &lnl, &warn),
       sh);
  tree->c_lnL = lnl;
  if (warn) tree->numerical_warning = YES;
  if (full && !getenv("PLK_SHIM_NO_SITE_READBACK"))
    CK(plk_get_site_lnl(sh->inst, tree->c_lnL_sorted, tree->cur_site_lk, tree->unscaled_site_lk_cat,
                        tree->fact_sum_scale),
       sh);
  sh->n_lk++;
  return tree->c_lnL;
}

/* ------------------------------------------------------------------------------------------------ */
/* src/lk.c:655-753                                                                                 */
phydbl dLk(phydbl *l, t_edge *b, t_tree *tree)
{
  shim_t *sh = shim_of(tree);
  double  lnl = 0.0, dlnl = 0.0;
  int     warn = 0;
  if (!sh)
  {
    static phydbl (*orig)(phydbl *, t_edge *, t_tree *) = NULL;
    if (!orig) orig = (phydbl(*)(phydbl *, t_edge *, t_tree *))dlsym(RTLD_NEXT, "dLk");
    return orig(l, b, tree);
  }
  tree->numerical_warning = NO;
  assert(isnan(*l) == FALSE);
  assert(b != NULL);
  if (tree->update_eigen_lr == YES) Update_Eigen_Lr(b, tree);
  upload_model_if_changed(sh);
  CK(plk_edge_lnl_dlnl(sh->inst, l, &lnl, &dlnl, &warn), sh); /* clamps *l like lk.c:673-674 */
  tree->c_dlnL = dlnl;
  tree->c_lnL = lnl;
  if (warn) tree->numerical_warning = YES;
  sh->n_dlk++;
  return tree->c_lnL;
}
