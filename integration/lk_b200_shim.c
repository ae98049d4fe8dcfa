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
 *     Pars / Pars_At_Given_Edge   src/pars.c:20,468 -> plk_pars_traverse_edge (queued updates + the site loop, ONE launch)
 *     Update_Partial_Pars         src/pars.c:239  -> queued plk_pars_op
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
 * Rooted trees (n_root, Add_Root src/utilities.c:8426): the operands of every update are resolved by the
 * reference's own Set_All_Partial_Lk (src/lk.c:2988-3194 are its rooted cases) and the two root edges
 * (a_edges[2n-3], [2n-2]) own P-matrix / CLV handles like any other edge; Lk() follows lk.c:504-576.
 * Unsupported configurations abort like the BEAGLE hooks did (src/main.c:240-253): mixture trees
 * (is_mixt_tree), SCALE_RATE_SPECIFIC, M4, gamma_mgf_bl, ns > 32, and any
 * likelihood call on a tree that has no device instance (bootstrap replicates share the main tree's
 * structures, src/utilities.c:4042) -- there is no silent CPU fallback.
 * Host-side readers of engine state: c_lnL_sorted / cur_site_lk / unscaled_site_lk_cat /
 * fact_sum_scale are mirrored after every Lk(NULL); inside aLRT() (src/alrt.c:172), whose
 * NNI_Neigh_BL reads c_lnL_sorted after edge-level calls (alrt.c:453,555,682), c_lnL_sorted is
 * mirrored after every Lk(b) as well.  CLVs and P-matrices stay on the device.
 *
 * No host likelihood arena: Make_Tree_For_Lk asks posix_memalign for (3n-2)*P*ncatg*ns doubles
 * (src/make.c:96-104; 3.8 GB at 100 taxa x 100k sites, 192 GB at 500 x 1M, and the size is computed in
 * int arithmetic, so the reference itself aborts beyond 2^31 elements).  The shim interposes that one
 * request and hands out an address-space reservation (mmap PROT_NONE | MAP_NORESERVE): the CLV and
 * P-matrix pointers the reference carves from it (make.c:516-522,572-576,681-685) keep their role as
 * buffer NAMES (they are the keys of the device handle tables) but no host memory backs them.
 * Make_Edge_Lk is wrapped so that every edge carves from its own sub-range: the reference's int
 * bump index (utilities.h:886) then never exceeds one edge's worth, whatever the alignment size.
 *
 * PLK_GPUS=N shards the instance over N devices of the box inside this one process
 * (plk_create_sharded): the single t_tree of lk.c drives all of them.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>

#include "utilities.h"
#include "lk.h"
#include "models.h"
#include "optimiz.h"
#include "make.h"
#include "free.h"
#include "alrt.h"
#include "pars.h"
#include "ancestral.h"

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
  double        model_print[8 + 2 * 32 + 2 * 16]; /* scalars, eigenvalues, pi, rates of the uploaded model */
  double       *model_uv;                          /* full copies of U and V (2 * ns * ns) */
  int           model_valid;
  double       *wght_print;                        /* data->wght as uploaded (bootstrap-style in-place edits) */
  long long     n_lk, n_dlk, n_partial, n_flush, n_pmat;
  /* parsimony (src/pars.c): buffers keyed by the host pointer ui_l / ui_r (swapped together with pars_* and
     p_pars_* by Prune_Subtree / Graft_Subtree, src/utilities.c:6268-6278) */
  slot_t       *pars_map;
  int           n_pars, pars_cap, pars_created;
  unsigned char *pars_on_dev[2];                   /* per handle: the device holds the Fitch [0] / step-matrix [1] data */
  plk_pars_op  *pars_queue;
  int           n_pars_queue, pars_queue_cap, pars_queue_general;
  long long     n_pars_calls, n_pars_upd;
  char         *arena;      /* address-space reservation standing in for tree->big_lk_array */
  size_t        arena_bytes, edge_span;
  double        t_create, t_engine; /* wall-clock: instance creation time stamp, seconds spent inside the hooks */
} shim_t;

#define MAX_SHIMS 16
static shim_t g_shims[MAX_SHIMS];
static int    g_mirror_sites = 0; /* inside aLRT(): c_lnL_sorted is read after edge-level Lk() calls */

/* wall-clock accounting of the time spent inside the hooks (outermost call only) */
static int    g_depth = 0;
static double g_t_enter = 0.0;
static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
#define HOOK_ENTER()                         \
  do                                         \
  {                                          \
    if (g_depth++ == 0) g_t_enter = now_s(); \
  } while (0)
#define HOOK_LEAVE(sh)                                             \
  do                                                               \
  {                                                                \
    if (--g_depth == 0 && (sh)) (sh)->t_engine += now_s() - g_t_enter; \
  } while (0)

/* ------------------------------------------------------------------------------------------------ */
/* Address-only likelihood arena.  The reference calls posix_memalign(&tree->big_lk_array, ...) once  */
/* per tree (src/make.c:96-104); that request, and only that one, is answered with a reservation.    */
static t_tree *g_arena_tree = NULL;
static char   *g_arena_base = NULL;
static size_t  g_arena_bytes = 0, g_arena_edge_span = 0;
static int     g_arena_edge = 0;

int posix_memalign(void **memptr, size_t alignment, size_t size)
{
  static int (*real)(void **, size_t, size_t) = NULL;
  if (g_arena_tree && memptr == (void **)&g_arena_tree->big_lk_array)
  {
    const t_tree *tree = g_arena_tree;
    const size_t  nc = (size_t)MAX(tree->mod->ras->n_catg, tree->mod->n_mixt_classes);
    const size_t  ns = (size_t)tree->mod->ns;
    /* per edge: Pij_rr, tPij_rr, p_lk_left, p_lk_rght (make.c:516-522,572-576,681-685), 64-bit arithmetic */
    size_t span = (2 * (size_t)tree->mod->ras->n_catg * ns * ns + 2 * (size_t)tree->data->n_pattern * nc * ns) * sizeof(phydbl);
    span = (span + 4095) & ~(size_t)4095;
    g_arena_edge_span = span;
    g_arena_bytes = span * (size_t)(2 * tree->n_otu);
    g_arena_base = (char *)mmap(NULL, g_arena_bytes, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_arena_base == (char *)MAP_FAILED)
    {
      g_arena_base = NULL;
      return 12; /* ENOMEM */
    }
    g_arena_edge = 0;
    *memptr = g_arena_base;
    (void)alignment;
    (void)size; /* the reference's own figure overflows int beyond 2^31 elements: not used */
    return 0;
  }
  if (!real) real = (int (*)(void **, size_t, size_t))dlsym(RTLD_NEXT, "posix_memalign");
  return real(memptr, alignment, size);
}

/* src/make.c:480-525: every edge carves its four buffers from its own sub-range of the reservation */
void Make_Edge_Lk(t_edge *b, t_tree *tree)
{
  static void (*orig)(t_edge *, t_tree *) = NULL;
  if (!orig) orig = (void (*)(t_edge *, t_tree *))dlsym(RTLD_NEXT, "Make_Edge_Lk");
  if (tree == g_arena_tree && g_arena_base)
  {
    if (g_arena_edge >= 2 * tree->n_otu)
    {
      PhyML_Fprintf(stderr, "\n. phyml_b200: more Make_Edge_Lk calls than edges\n");
      Exit("\n");
    }
    tree->big_lk_array = (phydbl *)(g_arena_base + (size_t)g_arena_edge * g_arena_edge_span);
    tree->big_lk_array_pos = 0;
    g_arena_edge++;
  }
  orig(b, tree);
}

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
/* look-up only: -1 when the device has never seen this buffer name */
static int map_find(const slot_t *map, int cap, const void *key)
{
  uintptr_t h = ((uintptr_t)key >> 4) * 0x9E3779B97F4A7C15ULL;
  int       i = (int)(h % (uintptr_t)cap);
  while (map[i].key && map[i].key != key) i = (i + 1) % cap;
  return map[i].key ? map[i].val : -1;
}
static int clv_handle(shim_t *sh, const phydbl *p) { return map_get(sh->clv_map, sh->map_cap, p, &sh->n_clv, sh->clv_cap, "CLV"); }
static int pm_handle(shim_t *sh, const phydbl *p) { return map_get(sh->pm_map, sh->map_cap, p, &sh->n_pm, sh->pm_cap, "P-matrix"); }

static void flush(shim_t *sh);

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
  if (sh->model_valid && !memcmp(fp, sh->model_print, sizeof(fp)) &&
      !memcmp(sh->model_uv, mod->eigen->r_e_vect, sizeof(double) * ns * ns) &&
      !memcmp(sh->model_uv + ns * ns, mod->eigen->l_e_vect, sizeof(double) * ns * ns))
    return;
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
  memcpy(sh->model_uv, mod->eigen->r_e_vect, sizeof(double) * ns * ns);
  memcpy(sh->model_uv + ns * ns, mod->eigen->l_e_vect, sizeof(double) * ns * ns);
  sh->model_valid = 1;
}

/* data->wght is edited in place by some callers (e.g. the bootstrap resampling of src/stats.c:2093-2133) */
static void upload_weights_if_changed(shim_t *sh)
{
  const size_t n = sizeof(double) * (size_t)sh->tree->data->n_pattern;
  if (!memcmp(sh->wght_print, sh->tree->data->wght, n)) return;
  flush(sh);
  CK(plk_set_pattern_weights(sh->inst, sh->tree->data->wght, sh->tree->data->invar), sh);
  memcpy(sh->wght_print, sh->tree->data->wght, n);
}

static void no_instance(const char *fn)
{
  PhyML_Fprintf(stderr, "\n. phyml_b200: %s() was called on a tree without a device instance (trees that share another"
                        "\n. tree's likelihood structures, e.g. bootstrap replicates, are not supported): use the CPU build.\n", fn);
  Exit("\n");
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
  if (tree->is_mixt_tree == YES || tree->mixt_tree != NULL || tree->mod->use_m4mod == YES ||
      tree->mod->gamma_mgf_bl == YES || tree->scaling_method != SCALE_FAST || tree->mod->ns > 32 ||
      (tree->io && tree->io->do_alias_subpatt == YES))
  {
    PhyML_Fprintf(stderr, "\n. phyml_b200: unsupported configuration (mixture / M4 / MGF branch lengths /"
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
  int         i, n_gpus;
  if (!orig) orig = (void (*)(t_tree *))dlsym(RTLD_NEXT, "Make_Tree_For_Lk");
  check_supported(tree);
  /* the reference's own allocation sequence, with the arena request answered by a reservation */
  g_arena_tree = tree;
  g_arena_base = NULL;
  orig(tree);
  g_arena_tree = NULL;
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
  sh->arena = g_arena_base;
  sh->arena_bytes = g_arena_bytes;
  sh->edge_span = g_arena_edge_span;
  if (sh->arena) tree->big_lk_array = (phydbl *)sh->arena;
  sh->t_create = now_s();
  sh->model_uv = (double *)calloc((size_t)2 * tree->mod->ns * tree->mod->ns, sizeof(double));
  sh->wght_print = (double *)malloc(sizeof(double) * (size_t)tree->data->n_pattern);
  memcpy(sh->wght_print, tree->data->wght, sizeof(double) * (size_t)tree->data->n_pattern);
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
  n_gpus = getenv("PLK_GPUS") ? atoi(getenv("PLK_GPUS")) : 1;
  if ((n_gpus > 1 ? plk_create_sharded(&cfg, n_gpus, NULL, &sh->inst) : plk_create(&cfg, &sh->inst)) != PLK_OK)
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
    PhyML_Printf("\n. phyml_b200: %s, instance for %d taxa x %d patterns, ns=%d ncatg=%d on %d GPU(s); host likelihood "
                 "arena: %.1f MB of address space reserved, 0 bytes committed",
                 plk_version(), tree->n_otu, tree->data->n_pattern, tree->mod->ns, tree->mod->ras->n_catg,
                 plk_n_shards(sh->inst), (double)sh->arena_bytes / 1e6);
}

void Free_Tree_Lk(t_tree *tree)
{
  static void (*orig)(t_tree *) = NULL;
  shim_t *sh = shim_of(tree);
  if (!orig) orig = (void (*)(t_tree *))dlsym(RTLD_NEXT, "Free_Tree_Lk");
  if (sh)
  {
    if (getenv("PLK_SHIM_VERBOSE"))
    {
      const double total = now_s() - sh->t_create;
      PhyML_Printf("\n. phyml_b200: Lk %lld  dLk %lld  Update_Partial_Lk %lld (in %lld launches)  Update_PMat %lld  "
                   "kernels %lld\n",
                   sh->n_lk, sh->n_dlk, sh->n_partial, sh->n_flush, sh->n_pmat, plk_launch_count(sh->inst));
      if (sh->n_pars_calls)
        PhyML_Printf(". phyml_b200: Pars %lld  Update_Partial_Pars %lld\n", sh->n_pars_calls, sh->n_pars_upd);
      PhyML_Printf(". phyml_b200: wall-clock since the instance was created %.2f s: %.2f s inside the likelihood hooks "
                   "(engine + binding), %.2f s in the untouched host code (spr.c, optimiz.c, ...)\n",
                   total, sh->t_engine, total - sh->t_engine);
    }
    plk_destroy(sh->inst);
    if (sh->arena)
    { /* free.c:391 frees big_lk_array: give it something free() accepts, then drop the reservation */
      munmap(sh->arena, sh->arena_bytes);
      tree->big_lk_array = (phydbl *)malloc(8);
    }
    free(sh->clv_map);
    free(sh->pm_map);
    free(sh->pars_map);
    free(sh->pars_on_dev[0]);
    free(sh->pars_on_dev[1]);
    free(sh->pars_queue);
    free(sh->queue);
    free(sh->model_uv);
    free(sh->wght_print);
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
  if (!sh) no_instance("Update_PMat_At_Given_Edge");
  assert(b_fcus && b_fcus->Pij_rr);
  HOOK_ENTER();
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
    HOOK_LEAVE(sh);
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
        HOOK_LEAVE(sh);
        return;
      }
    sh->pm_h[sh->n_pm_queue] = h;
    sh->pm_l[sh->n_pm_queue] = l;
    sh->n_pm_queue++;
  }
  HOOK_LEAVE(sh);
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
  if (!sh) no_instance("Update_Partial_Lk");
  if (b->left == d && b->update_partial_lk_left == NO) return;
  if (b->rght == d && b->update_partial_lk_rght == NO) return;
  if (d->tax) return;
  Set_All_Partial_Lk(&n_v1, &n_v2, &p_lk, &sum_scale, &p_lk_loc, &Pij1, &tPij1, &p_lk_v1, &sum_scale_v1, &Pij2,
                     &tPij2, &p_lk_v2, &sum_scale_v2, d, b, tree);
  if (!n_v1 || !n_v2)
  { /* d == n_root with ignore_root == NO (lk.c:3010-3050): a one-child update.  The reference's own AVX / SSE /
       default kernels dereference n_v1 there (avx.c:450, lk.c:1736), only its generic kernel survives it */
    PhyML_Fprintf(stderr, "\n. phyml_b200: one-child update at the root node (ignore_root == NO) is not supported.\n");
    Exit("\n");
  }
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
  if (!sh) no_instance("Update_Eigen_Lr");
  HOOK_ENTER();
  flush(sh);
  upload_model_if_changed(sh);
  CK(plk_eigen_lr(sh->inst, side_of(sh, b->left, b->p_lk_left), side_of(sh, b->rght, b->p_lk_rght)), sh);
  HOOK_LEAVE(sh);
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
  if (!sh) no_instance("Lk");
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
  HOOK_ENTER();
  if (b == NULL) upload_weights_if_changed(sh);
  upload_model_if_changed(sh);

  /* skip_tree_traversal (Optimiz_Alpha_And_Pinv, optimiz.c:2215-2222) only saves work in the
     reference; recomputing gives the same value, so it is ignored here */
  if (!b)
  { /* lk.c:500-511: all P-matrices in one batched launch (+ the two root edges of a rooted tree) */
    flush(sh);
    const int rooted = (tree->n_root != NULL && tree->ignore_root == NO);
    const int n = 2 * tree->n_otu - 3 + (rooted ? 2 : 0);
    int      *h = (int *)malloc(sizeof(int) * n);
    double   *l = (double *)malloc(sizeof(double) * n);
    int       n_batched = 0;
    for (br = 0; br < (unsigned int)n; ++br)
    {
      t_edge *e = (br < (unsigned int)(2 * tree->n_otu - 3)) ? tree->a_edges[br] : tree->n_root->b[br - (2 * tree->n_otu - 3) + 1];
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
    /* lk.c:529-566: the reference's own recursions; every visit lands in Update_Partial_Lk above */
    if (tree->n_root != NULL)
    {
      if (tree->ignore_root == NO)
      { /* rooted evaluation (PhyTime / PhyREX trees): both subtrees of the root, then the root's two edges */
        Post_Order_Lk(tree->n_root, tree->n_root->v[1], tree);
        Post_Order_Lk(tree->n_root, tree->n_root->v[2], tree);
        Update_Partial_Lk(tree, tree->n_root->b[1], tree->n_root);
        Update_Partial_Lk(tree, tree->n_root->b[2], tree->n_root);
        if (tree->both_sides == YES)
        {
          Pre_Order_Lk(tree->n_root, tree->n_root->v[2], tree);
          Pre_Order_Lk(tree->n_root, tree->n_root->v[1], tree);
        }
        b = (tree->n_root->v[1]->tax == NO) ? (tree->n_root->b[2]) : (tree->n_root->b[1]); /* lk.c:572-573 */
      }
      else
      { /* the root is ignored: the traversal starts from both ends of the edge that carries it */
        Post_Order_Lk(tree->e_root->rght, tree->e_root->left, tree);
        Post_Order_Lk(tree->e_root->left, tree->e_root->rght, tree);
        if (tree->both_sides == YES)
        {
          Pre_Order_Lk(tree->e_root->rght, tree->e_root->left, tree);
          Pre_Order_Lk(tree->e_root->left, tree->e_root->rght, tree);
        }
        b = tree->e_root; /* lk.c:575 */
      }
    }
    else
    {
      Post_Order_Lk(tree->a_nodes[tree->tip_root], tree->a_nodes[tree->tip_root]->v[0], tree);
      if (tree->both_sides == YES) Pre_Order_Lk(tree->a_nodes[tree->tip_root], tree->a_nodes[tree->tip_root]->v[0], tree);
      b = tree->a_nodes[tree->tip_root]->b[0]; /* lk.c:578-579 */
    }
  }
  else if (tree->use_eigen_lr == NO && tree->n_root && (b == tree->n_root->b[1] || b == tree->n_root->b[2]) &&
           tree->ignore_root == YES)
    Update_PMat_At_Given_Edge(tree->e_root, tree); /* lk.c:517-522 */
  else if (tree->use_eigen_lr == NO)
    Update_PMat_At_Given_Edge(b, tree); /* lk.c:515-527 */

  tree->c_lnL = .0;
  tree->sum_min_sum_scale = .0;
  if (tree->update_eigen_lr == NO && tree->use_eigen_lr == NO && sh->n_queue > 0)
  { /* the queued CLV updates and the site loop at b as ONE engine call: a full traversal (Lk(NULL)) or the one to
       three updates of an SPR / NNI candidate followed by Lk(b) become a single launch (lk.c:562-645) */
    if (sh->n_pm_queue > 0)
    {
      CK(plk_update_pmats(sh->inst, sh->n_pm_queue, sh->pm_h, sh->pm_l), sh);
      sh->n_pm_queue = 0;
    }
    CK(plk_traverse_edge_lnl(sh->inst, sh->n_queue, sh->queue, side_of(sh, b->left, b->p_lk_left),
                             side_of(sh, b->rght, b->p_lk_rght), pm_handle(sh, b->Pij_rr), &lnl, &warn),
       sh);
    sh->n_queue = 0;
    sh->n_flush++;
    goto have_lnl;
  }
  flush(sh);
  if (tree->update_eigen_lr == YES) Update_Eigen_Lr(b, tree); /* lk.c:590 */
  if (tree->use_eigen_lr == YES)
    CK(plk_edge_lnl_eigen(sh->inst, b->l->v, &lnl, &warn), sh); /* lk.c:592-603,625-629 */
  else
    CK(plk_edge_lnl(sh->inst, side_of(sh, b->left, b->p_lk_left), side_of(sh, b->rght, b->p_lk_rght),
                    pm_handle(sh, b->Pij_rr), &lnl, &warn),
       sh);
have_lnl:
  tree->c_lnL = lnl;
  if (warn) tree->numerical_warning = YES;
  if (full && !getenv("PLK_SHIM_NO_SITE_READBACK"))
    CK(plk_get_site_lnl(sh->inst, tree->c_lnL_sorted, tree->cur_site_lk, tree->unscaled_site_lk_cat,
                        tree->fact_sum_scale),
       sh);
  else if (g_mirror_sites) /* alrt.c:453,555,682 read c_lnL_sorted after Br_Len_Opt / Lk(b) */
    CK(plk_get_site_lnl(sh->inst, tree->c_lnL_sorted, NULL, NULL, NULL), sh);
  sh->n_lk++;
  HOOK_LEAVE(sh);
  return tree->c_lnL;
}

/* ------------------------------------------------------------------------------------------------ */
/* src/lk.c:655-753                                                                                 */
phydbl dLk(phydbl *l, t_edge *b, t_tree *tree)
{
  shim_t *sh = shim_of(tree);
  double  lnl = 0.0, dlnl = 0.0;
  int     warn = 0;
  if (!sh) no_instance("dLk");
  tree->numerical_warning = NO;
  assert(isnan(*l) == FALSE);
  assert(b != NULL);
  HOOK_ENTER();
  if (tree->update_eigen_lr == YES) Update_Eigen_Lr(b, tree);
  upload_model_if_changed(sh);
  CK(plk_edge_lnl_dlnl(sh->inst, l, &lnl, &dlnl, &warn), sh); /* clamps *l like lk.c:673-674 */
  tree->c_dlnL = dlnl;
  tree->c_lnL = lnl;
  if (warn) tree->numerical_warning = YES;
  sh->n_dlk++;
  HOOK_LEAVE(sh);
  return tree->c_lnL;
}

/* ------------------------------------------------------------------------------------------------ */
/* src/alrt.c:172: the branch-support tests read tree->c_lnL_sorted right after edge-level likelihood  */
/* calls (NNI_Neigh_BL, alrt.c:453,555,682): mirror it after every Lk() while aLRT() runs             */
void aLRT(t_tree *tree)
{
  static void (*orig)(t_tree *) = NULL;
  if (!orig) orig = (void (*)(t_tree *))dlsym(RTLD_NEXT, "aLRT");
  g_mirror_sites = 1;
  orig(tree);
  g_mirror_sites = 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Parsimony: src/pars.c.  The SPR search scores every regraft position by Fitch parsimony before any  */
/* likelihood is computed (spr_pars, src/init.c:786; Test_One_Spr_Target src/spr.c:636-640).  The       */
/* reference's recursions (Post_Order_Pars / Pre_Order_Pars, src/pars.c:56-93) run unchanged and land  */
/* in Update_Partial_Pars below; updates are queued and executed together with the site loop of Pars   */
/* as ONE launch.  Device buffers are created lazily: a buffer that the device has never written is     */
/* uploaded from the host arrays the first time it is read (tips: Init_Ui_Tips, src/pars.c:164).       */
static void pars_init(shim_t *sh)
{
  t_tree *tree = sh->tree;
  if (sh->pars_created) return;
  if (!tree->step_mat || !tree->site_pars)
  {
    PhyML_Fprintf(stderr, "\n. phyml_b200: parsimony call before Make_Tree_For_Pars\n");
    Exit("\n");
  }
  sh->pars_cap = 8 * tree->n_otu + 16;
  sh->pars_map = (slot_t *)calloc(sh->map_cap, sizeof(slot_t));
  sh->pars_on_dev[0] = (unsigned char *)calloc(sh->pars_cap, 1);
  sh->pars_on_dev[1] = (unsigned char *)calloc(sh->pars_cap, 1);
  sh->pars_queue_cap = 8 * tree->n_otu + 64;
  sh->pars_queue = (plk_pars_op *)malloc(sizeof(plk_pars_op) * sh->pars_queue_cap);
  CK(plk_pars_create(sh->inst, sh->pars_cap, tree->step_mat), sh);
  sh->pars_created = 1;
}

static void pars_flush(shim_t *sh)
{
  if (sh->n_pars_queue == 0) return;
  CK(plk_pars_update(sh->inst, sh->pars_queue_general, sh->n_pars_queue, sh->pars_queue), sh);
  sh->n_pars_queue = 0;
}

/* src/make.c:334-372, src/free.c:330-353: when the host (re)allocates or frees its parsimony arrays the pointer
   keys change meaning: forget them, the device buffers are refilled on first sight */
static void pars_forget(shim_t *sh)
{
  if (!sh || !sh->pars_created) return;
  sh->n_pars_queue = 0;
  memset(sh->pars_map, 0, sizeof(slot_t) * sh->map_cap);
  memset(sh->pars_on_dev[0], 0, sh->pars_cap);
  memset(sh->pars_on_dev[1], 0, sh->pars_cap);
  sh->n_pars = 0;
}
void Make_Tree_For_Pars(t_tree *tree)
{
  static void (*orig)(t_tree *) = NULL;
  if (!orig) orig = (void (*)(t_tree *))dlsym(RTLD_NEXT, "Make_Tree_For_Pars");
  orig(tree);
  pars_forget(shim_of(tree));
}
void Free_Tree_Pars(t_tree *tree)
{
  static void (*orig)(t_tree *) = NULL;
  if (!orig) orig = (void (*)(t_tree *))dlsym(RTLD_NEXT, "Free_Tree_Pars");
  pars_forget(shim_of(tree));
  orig(tree);
}

/* handle of a buffer that is about to be READ: first sight => the host arrays are the current contents */
static int pars_src(shim_t *sh, int general, int *ui, int *pars, int *p_pars)
{
  const int h = map_get(sh->pars_map, sh->map_cap, ui, &sh->n_pars, sh->pars_cap, "parsimony");
  if (!sh->pars_on_dev[general][h])
  {
    CK(plk_pars_set_buffer(sh->inst, h, general ? NULL : ui, general ? NULL : pars, general ? p_pars : NULL), sh);
    sh->pars_on_dev[general][h] = 1;
  }
  return h;
}

static int pars_dst(shim_t *sh, int general, int *ui)
{
  const int h = map_get(sh->pars_map, sh->map_cap, ui, &sh->n_pars, sh->pars_cap, "parsimony");
  sh->pars_on_dev[general][h] = 1;
  return h;
}

/* src/pars.c:239-391 */
void Update_Partial_Pars(t_tree *tree, t_edge *b_fcus, t_node *n)
{
  shim_t      *sh = shim_of(tree);
  t_edge      *b1, *b2;
  plk_pars_op *op;
  int          general, left;
  if (!sh) no_instance("Update_Partial_Pars");
  if (n->tax) return; /* pars.c:268 */
  HOOK_ENTER();
  pars_init(sh);
  general = tree->mod->s_opt->general_pars ? 1 : 0;
  if (sh->n_pars_queue > 0 && (sh->pars_queue_general != general || sh->n_pars_queue == sh->pars_queue_cap)) pars_flush(sh);
  sh->pars_queue_general = general;
  left = (n == b_fcus->left);
  b1 = n->b[left ? b_fcus->l_v1 : b_fcus->r_v1]; /* pars.c:277-351: the far-end buffers of n's two other edges */
  b2 = n->b[left ? b_fcus->l_v2 : b_fcus->r_v2];
  op = &sh->pars_queue[sh->n_pars_queue];
  op->c1 = (n == b1->left) ? pars_src(sh, general, b1->ui_r, b1->pars_r, b1->p_pars_r)
                           : pars_src(sh, general, b1->ui_l, b1->pars_l, b1->p_pars_l);
  op->c2 = (n == b2->left) ? pars_src(sh, general, b2->ui_r, b2->pars_r, b2->p_pars_r)
                           : pars_src(sh, general, b2->ui_l, b2->pars_l, b2->p_pars_l);
  op->dst = pars_dst(sh, general, left ? b_fcus->ui_l : b_fcus->ui_r);
  sh->n_pars_queue++;
  sh->n_pars_upd++;
  HOOK_LEAVE(sh);
}

/* the site loop of Pars / Pars_At_Given_Edge (src/pars.c:39-50,468-484) behind the queued updates */
static int pars_at_edge(shim_t *sh, t_edge *b)
{
  t_tree   *tree = sh->tree;
  const int general = tree->mod->s_opt->general_pars ? 1 : 0;
  int       c_pars = 0, hl, hr;
  pars_init(sh);
  if (sh->n_pars_queue > 0 && sh->pars_queue_general != general) pars_flush(sh);
  hl = pars_src(sh, general, b->ui_l, b->pars_l, b->p_pars_l);
  hr = pars_src(sh, general, b->ui_r, b->pars_r, b->p_pars_r);
  CK(plk_pars_traverse_edge(sh->inst, general, sh->n_pars_queue, sh->pars_queue, hl, hr, &c_pars), sh);
  sh->n_pars_queue = 0;
  sh->n_pars_calls++;
  if (getenv("PLK_SHIM_SITE_PARS")) CK(plk_get_site_pars(sh->inst, tree->site_pars), sh); /* no host reader needs it */
  tree->c_pars = c_pars;
  return c_pars;
}

/* src/pars.c:20-51 */
int Pars(t_edge *b, t_tree *tree)
{
  shim_t *sh = shim_of(tree);
  int     c_pars;
  if (!sh) no_instance("Pars");
  HOOK_ENTER();
  if (b == NULL)
  { /* pars.c:33-39: the reference's own recursions; every visit lands in Update_Partial_Pars above */
    Post_Order_Pars(tree->a_nodes[0], tree->a_nodes[0]->v[0], tree);
    if (tree->both_sides == YES) Pre_Order_Pars(tree->a_nodes[0], tree->a_nodes[0]->v[0], tree);
    b = tree->a_nodes[0]->b[0];
  }
  c_pars = pars_at_edge(sh, b);
  HOOK_LEAVE(sh);
  return c_pars;
}

/* src/pars.c:468-484 */
int Pars_At_Given_Edge(t_edge *b, t_tree *tree)
{
  shim_t *sh = shim_of(tree);
  int     c_pars;
  if (!sh) no_instance("Pars_At_Given_Edge");
  HOOK_ENTER();
  c_pars = pars_at_edge(sh, b);
  HOOK_LEAVE(sh);
  return c_pars;
}

/* src/pars.c:104,443: host-side readers of the per-edge Fitch sets; no caller in the reference */
void Site_Pars(t_tree *tree)
{
  (void)tree;
  PhyML_Fprintf(stderr, "\n. phyml_b200: Site_Pars() reads host parsimony buffers that live on the device.\n");
  Exit("\n");
}
int One_Pars_Step(t_edge *b, t_tree *tree)
{
  (void)b;
  (void)tree;
  PhyML_Fprintf(stderr, "\n. phyml_b200: One_Pars_Step() reads host parsimony buffers that live on the device.\n");
  Exit("\n");
  return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Host readers of CLVs / scalers / P-matrices: Ancestral_Sequences       */
/* (src/ancestral.c:527, called by main() after Lk(NULL) with both sides, main.c:289) reads            */
/* b->p_lk_left / p_lk_rght, sum_scale_* and Pij_rr of every edge on the host.  While it runs, the     */
/* address-only arena is backed by real pages and filled from the device; afterwards the pages are     */
/* given back.  (The host memory this needs is the reference's own requirement for this analysis.)     */
static void mirror_side(shim_t *sh, t_node *n, phydbl *p_lk, int *sum_scale)
{
  int h;
  if (!n || n->tax || !p_lk) return; /* tips keep their host p_lk_tip_r (make.c:687-705) */
  h = map_find(sh->clv_map, sh->map_cap, p_lk);
  if (h < 0) return;                 /* never computed on the device: the reference would read zeros as well */
  CK(plk_get_clv(sh->inst, h, p_lk, sum_scale), sh);
}

static void mirror_edge(shim_t *sh, t_edge *b)
{
  const int ns = sh->tree->mod->ns, nc = sh->tree->mod->ras->n_catg;
  int       h, c, i, j;
  if (!b) return;
  mirror_side(sh, b->left, b->p_lk_left, b->sum_scale_left);
  mirror_side(sh, b->rght, b->p_lk_rght, b->sum_scale_rght);
  if (b->Pij_rr && (h = map_find(sh->pm_map, sh->map_cap, b->Pij_rr)) >= 0)
  {
    CK(plk_get_pmat(sh->inst, h, b->Pij_rr), sh);
    if (b->tPij_rr)
      for (c = 0; c < nc; ++c)
        for (i = 0; i < ns; ++i)
          for (j = 0; j < ns; ++j) b->tPij_rr[c * ns * ns + j * ns + i] = b->Pij_rr[c * ns * ns + i * ns + j]; /* models.c:313 */
  }
}

static void host_mirror_begin(shim_t *sh)
{
  t_tree *tree = sh->tree;
  int     e;
  flush(sh);
  if (sh->arena && mprotect(sh->arena, sh->arena_bytes, PROT_READ | PROT_WRITE) != 0)
  {
    PhyML_Fprintf(stderr, "\n. phyml_b200: cannot back the likelihood arena with host memory (%.1f MB)\n", (double)sh->arena_bytes / 1e6);
    Exit("\n");
  }
  for (e = 0; e < 2 * tree->n_otu - 3; ++e) mirror_edge(sh, tree->a_edges[e]);
  if (tree->n_root)
  {
    mirror_edge(sh, tree->n_root->b[1]);
    mirror_edge(sh, tree->n_root->b[2]);
  }
}

static void host_mirror_end(shim_t *sh)
{
  if (!sh->arena) return;
  madvise(sh->arena, sh->arena_bytes, MADV_DONTNEED);
  mprotect(sh->arena, sh->arena_bytes, PROT_NONE);
}

/* src/ancestral.c:527-604 */
void Ancestral_Sequences(t_tree *tree, int print)
{
  static void (*orig)(t_tree *, int) = NULL;
  shim_t *sh = shim_of(tree);
  if (!orig) orig = (void (*)(t_tree *, int))dlsym(RTLD_NEXT, "Ancestral_Sequences");
  if (!sh) no_instance("Ancestral_Sequences");
  host_mirror_begin(sh);
  orig(tree, print);
  host_mirror_end(sh);
}
