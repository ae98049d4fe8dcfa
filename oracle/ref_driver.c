/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Library-style driver around the UNMODIFIED reference (linked as _ref/libphyml_ref.so, built by
 * oracle/Makefile from the sources under /root/reference, never copied into this repo).
 * It replays the set-up sequence of the reference's own main() (src/main.c:73-260):
 *   Get_Input -> Get_Seq -> Make_Model_Complete -> Compact_Data -> Init_Model ->
 *   Set_Model_Parameters -> Dist_And_BioNJ | Read_User_Tree -> Connect_CSeqs_To_Nodes ->
 *   Make_Tree_For_Pars -> Make_Tree_For_Lk -> Make_Spr -> Set_Update_Eigen(YES); Lk(NULL,tree)
 * and then either
 *   --dump FILE : writes every input the likelihood hot path consumes (patterns, weights, tip
 *                 vectors, topology, branch lengths, eigen system, rates) and every array it
 *                 produces (P-matrices, CLVs, scalers, per-site lnL, dot_prod, lnL/dlnL probes)
 *                 as tagged binary records (read by tests/golden/make_golden.py), or
 *   --time N    : times N calls of Lk(NULL,tree) (both_sides as given) and prints one JSON line
 *                 (the "reference" CPU baseline of bench.py).
 *
 *   --summary FILE : after Lk(NULL) writes only the small things a full-size parity pin needs (sizes, the
 *                 model's eigen system and rates, lnL, pattern weights, per-pattern lnL) -- used by
 *                 tests/golden/make_golden_big.py for the BASELINE.json configurations.
 *
 *   --dump_pars FILE : parsimony (src/pars.c): after Pars(NULL) with both_sides == YES writes the Fitch sets
 *                 and step counts of both sides of every edge (ui_l/ui_r, pars_l/pars_r), site_pars, c_pars,
 *                 Pars(b) at every edge, and the same for the general (step-matrix) variant (p_pars_l/r)
 *                 -- read by tests/golden/make_golden_pars.py.
 *
 *   --rooted E  : Add_Root on edge E (src/utilities.c:8426) after the set-up, then Lk(NULL) with both_sides == YES
 *                 (the tree->n_root branches of src/lk.c:529-576 with ignore_root == YES, the default) and Lk(b)
 *                 at every 5th edge; prints one REF_ROOTED JSON line.  Host arrays are not read, so this mode also
 *                 runs when the driver is linked against the B200 binding (integration/_build/ref_driver_b200).
 *
 * usage: ref_driver [--dump FILE] [--summary FILE] [--dump_pars FILE] [--rooted E] [--time N] [--warmup W]
 *                   [--both_sides 0|1] [--dlk N_EDGES] -- <phyml args>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "utilities.h"
#include "lk.h"
#include "models.h"
#include "make.h"
#include "free.h"
#include "spr.h"
#include "pars.h"
#include "cl.h"
#include "io.h"
#include "bionj.h"
#include "optimiz.h"

static FILE *g_out = NULL;

static void rec(const char *name, char dtype, long long n, const void *data)
{
  /* record := name[48] | dtype (1 char: d=f64 i=i32 h=i16 B=u8) | pad[7] | n (i64) | payload */
  char hdr[48];
  char pad[8] = {0};
  size_t sz = (dtype == 'd') ? 8 : (dtype == 'i') ? 4 : (dtype == 'h') ? 2 : 1;
  memset(hdr, 0, sizeof(hdr));
  strncpy(hdr, name, sizeof(hdr) - 1);
  pad[0] = dtype;
  fwrite(hdr, 1, sizeof(hdr), g_out);
  fwrite(pad, 1, 8, g_out);
  fwrite(&n, sizeof(long long), 1, g_out);
  if (n > 0) fwrite(data, sz, (size_t)n, g_out);
}

static void rec_d(const char *name, long long n, const double *d) { rec(name, 'd', n, d); }
static void rec_i(const char *name, long long n, const int *d) { rec(name, 'i', n, d); }
static void rec_h(const char *name, long long n, const short *d) { rec(name, 'h', n, d); }
static void rec_1d(const char *name, double v) { rec(name, 'd', 1, &v); }
static void rec_1i(const char *name, int v) { rec(name, 'i', 1, &v); }

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void dump_all(t_tree *tree, int n_dlk_edges)
{
  const int n_otu = tree->n_otu;
  const int P = tree->data->n_pattern;
  const int ns = tree->mod->ns;
  const int ncatg = tree->mod->ras->n_catg;
  const int n_edges = 2 * n_otu - 3;
  char nm[64];
  int i, e;

  rec_1i("n_otu", n_otu);
  rec_1i("n_pattern", P);
  rec_1i("ns", ns);
  rec_1i("ncatg", ncatg);
  rec_1i("tip_root", tree->tip_root);
  rec_1i("root_edge", tree->a_nodes[tree->tip_root]->b[0]->num);
  rec_1i("invar_flag", tree->mod->ras->invar);
  rec_1d("pinvar", tree->mod->ras->pinvar->v);
  rec_1d("l_min", tree->mod->l_min);
  rec_1d("l_max", tree->mod->l_max);
  rec_1d("br_len_mult", tree->mod->br_len_mult->v);
  rec_1i("scaling_method", tree->scaling_method);
  rec_1i("apply_lk_scaling", tree->apply_lk_scaling);

  rec_d("wght", P, tree->data->wght);
  rec_h("invar", P, tree->data->invar);

  rec_d("U", ns * ns, tree->mod->eigen->r_e_vect);
  rec_d("V", ns * ns, tree->mod->eigen->l_e_vect);
  rec_d("lambda", ns, tree->mod->eigen->e_val);
  rec_d("pi", ns, tree->mod->e_frq->pi->v);
  rec_d("rates", ncatg, tree->mod->ras->gamma_rr->v);
  rec_d("rate_probs", ncatg, tree->mod->ras->gamma_r_proba->v);
  rec_1d("alpha", tree->mod->ras->alpha->v);
  if (tree->mod->r_mat && tree->mod->r_mat->qmat) rec_d("qmat", ns * ns, tree->mod->r_mat->qmat->v);

  /* tips: node numbers 0..n_otu-1; fp64 0/1 vectors live on the tip's only edge (make.c:687-705) */
  for (i = 0; i < n_otu; ++i)
  {
    t_node *tip = tree->a_nodes[i];
    t_edge *b = tip->b[0];
    sprintf(nm, "tip%d.d_state", i);
    rec_h(nm, P, tip->c_seq->d_state);
    sprintf(nm, "tip%d.is_ambigu", i);
    rec_h(nm, P, tip->c_seq->is_ambigu);
    sprintf(nm, "tip%d.vec", i);
    rec_d(nm, (long long)P * ns, (b->rght == tip) ? b->p_lk_tip_r : b->p_lk_tip_l);
    sprintf(nm, "tip%d.edge", i);
    rec_1i(nm, b->num);
    sprintf(nm, "tip%d.name", i);
    rec(nm, 'B', (long long)strlen(tip->name), tip->name);
  }

  /* full evaluation with both directions so that every CLV is up to date */
  Set_Both_Sides(YES, tree);
  Set_Update_Eigen(YES, tree->mod);
  Lk(NULL, tree);
  Set_Update_Eigen(NO, tree->mod);

  rec_1d("lnL", tree->c_lnL);
  rec_d("site_lnl", P, tree->c_lnL_sorted);
  rec_d("site_lk", P, tree->cur_site_lk);
  rec_d("site_lk_cat", (long long)P * ncatg, tree->unscaled_site_lk_cat);
  rec_i("fact_sum_scale", P, tree->fact_sum_scale);

  for (e = 0; e < n_edges; ++e)
  {
    t_edge *b = tree->a_edges[e];
    int lr[2];
    lr[0] = b->left->num;
    lr[1] = b->rght->num;
    sprintf(nm, "edge%d.nodes", e);
    rec_i(nm, 2, lr);
    sprintf(nm, "edge%d.l", e);
    rec_1d(nm, b->l->v);
    sprintf(nm, "edge%d.P", e);
    rec_d(nm, (long long)ncatg * ns * ns, b->Pij_rr);
    if (!b->left->tax)
    {
      sprintf(nm, "edge%d.clv_left", e);
      rec_d(nm, (long long)P * ncatg * ns, b->p_lk_left);
      sprintf(nm, "edge%d.scale_left", e);
      rec_i(nm, P, b->sum_scale_left);
    }
    if (!b->rght->tax)
    {
      sprintf(nm, "edge%d.clv_rght", e);
      rec_d(nm, (long long)P * ncatg * ns, b->p_lk_rght);
      sprintf(nm, "edge%d.scale_rght", e);
      rec_i(nm, P, b->sum_scale_rght);
    }
  }

  /* lnL evaluated at every edge (pulley principle, cf. Check_Lk_At_Given_Edge lk.c:2642) */
  {
    double *lnl_e = (double *)malloc(sizeof(double) * n_edges);
    for (e = 0; e < n_edges; ++e) lnl_e[e] = Lk(tree->a_edges[e], tree);
    rec_d("edge_lnl", n_edges, lnl_e);
    free(lnl_e);
  }

  /* eigen-basis path used by Br_Len_Opt (optimiz.c:607-664): dot_prod, then lnL/dlnL probes */
  {
    static const double mult[5] = {0.1, 0.5, 1.0, 2.0, 10.0};
    int step = (n_dlk_edges > 0) ? (n_edges / n_dlk_edges) : n_edges;
    int cnt = 0;
    if (step < 1) step = 1;
    for (e = 0; e < n_edges && cnt < n_dlk_edges; e += step, ++cnt)
    {
      t_edge *b = tree->a_edges[e];
      double probes[5 * 4];
      double l0 = b->l->v;
      int k;
      Set_Update_Eigen_Lr(YES, tree);
      Set_Use_Eigen_Lr(NO, tree);
      Lk(b, tree);
      Set_Update_Eigen_Lr(NO, tree);
      Set_Use_Eigen_Lr(YES, tree);
      sprintf(nm, "dlk%d.edge", cnt);
      rec_1i(nm, e);
      sprintf(nm, "dlk%d.dot_prod", cnt);
      rec_d(nm, (long long)P * ncatg * ns, tree->dot_prod);
      for (k = 0; k < 5; ++k)
      {
        double l = l0 * mult[k];
        dLk(&l, b, tree);
        probes[4 * k + 0] = l; /* possibly clamped (lk.c:673-674) */
        probes[4 * k + 1] = tree->c_lnL;
        probes[4 * k + 2] = tree->c_dlnL;
        b->l->v = l;
        probes[4 * k + 3] = Lk(b, tree); /* use_eigen_lr==YES branch (lk.c:592-603,625-629) */
        b->l->v = l0;
      }
      sprintf(nm, "dlk%d.probes", cnt);
      rec_d(nm, 20, probes);
      Set_Use_Eigen_Lr(NO, tree);
      Update_PMat_At_Given_Edge(b, tree);
    }
    rec_1i("n_dlk", cnt);
  }
}

/* parsimony state of the reference: Pars (pars.c:20), Update_Partial_Pars (pars.c:239), Pars_Core (pars.c:397) */
static void dump_pars(t_tree *tree)
{
  const int n_otu = tree->n_otu;
  const int P = tree->data->n_pattern;
  const int ns = tree->mod->ns;
  const int n_edges = 2 * n_otu - 3;
  const int init_general = tree->mod->s_opt->general_pars;
  int *edge_pars = (int *)malloc(sizeof(int) * n_edges);
  char nm[64];
  int e, g;

  rec_1i("n_otu", n_otu);
  rec_1i("n_pattern", P);
  rec_1i("ns", ns);
  rec_i("step_mat", ns * ns, tree->step_mat);
  rec_d("wght", P, tree->data->wght);
  for (e = 0; e < n_edges; ++e)
  {
    int lr[2];
    lr[0] = tree->a_edges[e]->left->num;
    lr[1] = tree->a_edges[e]->rght->num;
    sprintf(nm, "edge%d.nodes", e);
    rec_i(nm, 2, lr);
  }
  Set_Both_Sides(YES, tree);
  for (g = 0; g < 2; ++g)
  {
    const char *sfx = g ? "_general" : "";
    tree->mod->s_opt->general_pars = g ? YES : NO;
    Pars(NULL, tree);
    sprintf(nm, "c_pars%s", sfx);
    rec_1i(nm, tree->c_pars);
    sprintf(nm, "site_pars%s", sfx);
    rec_i(nm, P, tree->site_pars);
    for (e = 0; e < n_edges; ++e)
    {
      t_edge *b = tree->a_edges[e];
      if (!g)
      {
        sprintf(nm, "edge%d.ui_l", e);
        rec_i(nm, P, b->ui_l);
        sprintf(nm, "edge%d.ui_r", e);
        rec_i(nm, P, b->ui_r);
        sprintf(nm, "edge%d.pars_l", e);
        rec_i(nm, P, b->pars_l);
        sprintf(nm, "edge%d.pars_r", e);
        rec_i(nm, P, b->pars_r);
      }
      else
      {
        sprintf(nm, "edge%d.p_pars_l", e);
        rec_i(nm, (long long)P * ns, b->p_pars_l);
        sprintf(nm, "edge%d.p_pars_r", e);
        rec_i(nm, (long long)P * ns, b->p_pars_r);
      }
    }
    for (e = 0; e < n_edges; ++e) edge_pars[e] = Pars(tree->a_edges[e], tree);
    sprintf(nm, "edge_pars%s", sfx);
    rec_i(nm, n_edges, edge_pars);
  }
  tree->mod->s_opt->general_pars = init_general;
  free(edge_pars);
}

static void dump_summary(t_tree *tree)
{
  const int P = tree->data->n_pattern;
  const int ns = tree->mod->ns;
  const int ncatg = tree->mod->ras->n_catg;
  rec_1i("n_otu", tree->n_otu);
  rec_1i("n_pattern", P);
  rec_1i("ns", ns);
  rec_1i("ncatg", ncatg);
  rec_1i("invar_flag", tree->mod->ras->invar);
  rec_1d("pinvar", tree->mod->ras->pinvar->v);
  rec_1d("l_min", tree->mod->l_min);
  rec_1d("l_max", tree->mod->l_max);
  rec_1d("br_len_mult", tree->mod->br_len_mult->v);
  rec_1d("alpha", tree->mod->ras->alpha->v);
  rec_d("U", ns * ns, tree->mod->eigen->r_e_vect);
  rec_d("V", ns * ns, tree->mod->eigen->l_e_vect);
  rec_d("lambda", ns, tree->mod->eigen->e_val);
  rec_d("pi", ns, tree->mod->e_frq->pi->v);
  rec_d("rates", ncatg, tree->mod->ras->gamma_rr->v);
  rec_d("rate_probs", ncatg, tree->mod->ras->gamma_r_proba->v);
  rec_1d("lnL", tree->c_lnL);
  rec_d("wght", P, tree->data->wght);
  rec_h("invar", P, tree->data->invar);
  rec_d("site_lnl", P, tree->c_lnL_sorted);
  rec_i("fact_sum_scale", P, tree->fact_sum_scale);
}

int main(int argc, char **argv)
{
  const char *dump_file = NULL, *summary_file = NULL, *pars_file = NULL;
  int n_time = 0, n_warm = 0, both_sides = 0, n_dlk = 4, rooted_edge = -1;
  int i, split = -1;
  option *io;
  calign *cdata;
  t_mod *mod;
  t_tree *tree;

  for (i = 1; i < argc; ++i)
  {
    if (!strcmp(argv[i], "--"))
    {
      split = i;
      break;
    }
    else if (!strcmp(argv[i], "--dump") && i + 1 < argc)
      dump_file = argv[++i];
    else if (!strcmp(argv[i], "--summary") && i + 1 < argc)
      summary_file = argv[++i];
    else if (!strcmp(argv[i], "--dump_pars") && i + 1 < argc)
      pars_file = argv[++i];
    else if (!strcmp(argv[i], "--rooted") && i + 1 < argc)
      rooted_edge = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--time") && i + 1 < argc)
      n_time = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--warmup") && i + 1 < argc)
      n_warm = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--both_sides") && i + 1 < argc)
      both_sides = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--dlk") && i + 1 < argc)
      n_dlk = atoi(argv[++i]);
    else
    {
      fprintf(stderr, "ref_driver: unknown option %s\n", argv[i]);
      return 2;
    }
  }
  if (split < 0)
  {
    fprintf(stderr, "usage: ref_driver [--dump FILE] [--time N] [--both_sides 0|1] [--dlk N] -- <phyml args>\n");
    return 2;
  }
  argv[split] = argv[0];
  io = (option *)Get_Input(argc - split, argv + split);
  if (!io) return 1;
  srand(io->r_seed < 0 ? 1 : io->r_seed);
  if (io->in_tree == 2)
    Test_Multiple_Data_Set_Format(io);
  else
    io->n_trees = 1;

  Get_Seq(io);
  Make_Model_Complete(io->mod);
  Set_Model_Name(io->mod);
  mod = io->mod;
  cdata = Compact_Data(io->data, io);
  Free_Seq(io->data, cdata->n_otu);
  Init_Model(cdata, mod, io);
  Set_Model_Parameters(mod);
  tree = (io->in_tree == 2) ? Read_User_Tree(cdata, mod, io) : Dist_And_BioNJ(cdata, mod, io);
  if (!tree) return 1;
  tree->mod = mod;
  tree->io = io;
  tree->data = cdata;
  tree->n_root = NULL;
  tree->e_root = NULL;
  tree->n_tot_bl_opt = 0;
  Set_Both_Sides(YES, tree);
  Connect_CSeqs_To_Nodes(tree->data, tree->io, tree);
  Make_Tree_For_Pars(tree);
  Make_Tree_For_Lk(tree);
  Make_Spr(tree);
  Br_Len_Not_Involving_Invar(tree);
  Unscale_Br_Len_Multiplier_Tree(tree);

  Set_Both_Sides(both_sides ? YES : NO, tree);
  Set_Update_Eigen(YES, tree->mod);
  Lk(NULL, tree);
  Set_Update_Eigen(NO, tree->mod);
  fprintf(stderr, "\nref_driver: n_otu=%d n_pattern=%d ns=%d ncatg=%d lnL=%.17g\n", tree->n_otu,
          tree->data->n_pattern, tree->mod->ns, tree->mod->ras->n_catg, tree->c_lnL);

  if (n_time > 0)
  {
    double *t = (double *)malloc(sizeof(double) * n_time);
    double tot = 0.0, best = 1e300;
    for (i = 0; i < n_warm; ++i) Lk(NULL, tree);
    for (i = 0; i < n_time; ++i)
    {
      double t0 = now_s();
      Lk(NULL, tree);
      t[i] = now_s() - t0;
      tot += t[i];
      if (t[i] < best) best = t[i];
    }
    printf("\nREF_TIMING {\"n_otu\": %d, \"n_pattern\": %d, \"ns\": %d, \"ncatg\": %d, \"both_sides\": %d, "
           "\"n_evals\": %d, \"mean_s\": %.9g, \"min_s\": %.9g, \"lnL\": %.17g}\n",
           tree->n_otu, tree->data->n_pattern, tree->mod->ns, tree->mod->ras->n_catg, both_sides, n_time,
           tot / n_time, best, tree->c_lnL);
    free(t);
  }

  if (rooted_edge >= 0)
  {
    const int n_edges = 2 * tree->n_otu - 3;
    const double unrooted = tree->c_lnL;
    int e, first = 1;
    Add_Root(tree->a_edges[rooted_edge % n_edges], tree);
    Set_Both_Sides(YES, tree);
    Lk(NULL, tree);
    printf("\nREF_ROOTED {\"root_edge\": %d, \"ignore_root\": %d, \"lnL_unrooted\": %.17g, \"lnL\": %.17g, \"edge_lnl\": [",
           tree->e_root->num, tree->ignore_root, unrooted, tree->c_lnL);
    for (e = 0; e < n_edges; e += 5)
    {
      printf("%s%.17g", first ? "" : ", ", Lk(tree->a_edges[e], tree));
      first = 0;
    }
    printf("]}\n");
  }

  if (summary_file)
  {
    g_out = fopen(summary_file, "wb");
    if (!g_out)
    {
      perror(summary_file);
      return 1;
    }
    dump_summary(tree);
    fclose(g_out);
    g_out = NULL;
  }

  if (pars_file)
  {
    g_out = fopen(pars_file, "wb");
    if (!g_out)
    {
      perror(pars_file);
      return 1;
    }
    dump_pars(tree);
    fclose(g_out);
    g_out = NULL;
  }

  if (dump_file)
  {
    g_out = fopen(dump_file, "wb");
    if (!g_out)
    {
      perror(dump_file);
      return 1;
    }
    dump_all(tree, n_dlk);
    fclose(g_out);
  }
  fflush(NULL);
  return 0;
}
