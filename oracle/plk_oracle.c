/* oracle/plk_oracle.c -- TEST INFRASTRUCTURE ONLY (see plk_oracle.h).
 *
 * Scalar fp64 restatement of PhyML's likelihood hot path.  Written from the semantics of the
 * reference (each function cites the file:line it follows), not copied from it; compiled with
 * -ffp-contract=off so every product and sum rounds exactly once, like the reference's scalar
 * (-DDISABLE_NATIVE) build.  Parity with the unmodified reference is pinned by
 * tests/test_oracle_golden.py against tests/golden/ (arrays dumped by oracle/ref_driver.c).
 */
#include "plk_oracle.h"

#include <float.h>
#include <math.h>
#include <stddef.h>

/* constants: src/utilities.h:267,476-478,507-520 */
#define ORC_LOG2 0.69314718055994528623
#define ORC_SMALL DBL_MIN
#define ORC_SMALL_PIJ 1.E-100
#define ORC_LARGE 256
static double two_pow_large(void) { return ldexp(1.0, ORC_LARGE); }

/* ------------------------------------------------------------------------------------------ */
/* K0  lk.c:2276-2325 + models.c:257-326                                                       */
void plk_oracle_pmat(int ns, int ncatg, double l, const double *rates, double br_len_mult,
                     double l_min, double l_max, const double *U, const double *V,
                     const double *lambda, double *P)
{
  double expt[64], uexpt[64 * 64];
  int c, i, j, k;
  for (c = 0; c < ncatg; ++c)
  {
    double *Pc = P + (size_t)c * ns * ns;
    /* lk.c:2296-2300: len = MAX(0,l)*rate; len *= mult; clamp to [l_min,l_max] */
    double len = (l > 0.0 ? l : 0.0) * rates[c];
    len *= br_len_mult;
    if (len < l_min)
      len = l_min;
    else if (len > l_max)
      len = l_max;
    /* models.c:275-279 */
    for (k = 0; k < ns; ++k) expt[k] = exp(lambda[k] * len);
    for (i = 0; i < ns; ++i)
      for (k = 0; k < ns; ++k) uexpt[i * ns + k] = U[i * ns + k] * expt[k];
    /* models.c:284-301: product, floor at SMALL_PIJ, row renormalisation */
    for (i = 0; i < ns; ++i)
    {
      double sum = 0.0;
      for (j = 0; j < ns; ++j)
      {
        double acc = 0.0;
        for (k = 0; k < ns; ++k) acc += uexpt[i * ns + k] * V[k * ns + j];
        if (acc < ORC_SMALL_PIJ) acc = ORC_SMALL_PIJ;
        Pc[i * ns + j] = acc;
      }
      for (j = 0; j < ns; ++j) sum += Pc[i * ns + j];
      for (j = 0; j < ns; ++j) Pc[i * ns + j] /= sum;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* K1  lk.c:1659-1768, 3328-3406                                                               */

/* matrix-vector: u[i] = sum_j P[i][j] v[j], accumulated in j order from 0.0 (lk.c:3344-3351) */
static void matvec(int ns, const double *P, const double *v, double *u)
{
  int i, j;
  for (i = 0; i < ns; ++i)
  {
    double acc = 0.0;
    for (j = 0; j < ns; ++j) acc += P[i * ns + j] * v[j];
    u[i] = acc;
  }
}

void plk_oracle_update_partial(int ns, int ncatg, int npat, const double *wght, int apply_scaling,
                               double *dst, int *dst_scale, const plk_oracle_side *c1,
                               const double *P1, const plk_oracle_side *c2, const double *P2)
{
  const int ncns = ncatg * ns;
  const double big = two_pow_large();
  const double inv_big = 1.0 / big;
  double u1[64], u2[64];
  int s, c, i;

  for (s = 0; s < npat; ++s)
  {
    double *out = dst + (size_t)s * ncns;
    int st1 = -1, st2 = -1, amb1 = 1, amb2 = 1; /* internal nodes count as "ambiguous" (lk.c:1685) */
    double largest;

    if (!(wght[s] > ORC_SMALL)) continue; /* lk.c:1682,1761-1766: zero-weight site left untouched */

    if (!c1->clv)
    {
      amb1 = c1->is_ambigu[s];
      if (!amb1) st1 = c1->d_state[s];
    }
    if (!c2->clv)
    {
      amb2 = c2->is_ambigu[s];
      if (!amb2) st2 = c2->d_state[s];
    }

    for (c = 0; c < ncatg; ++c)
    {
      const double *p1 = P1 + (size_t)c * ns * ns;
      const double *p2 = P2 + (size_t)c * ns * ns;
      const double *v1 = c1->clv ? c1->clv + (size_t)s * ncns + c * ns : c1->tipvec + (size_t)s * ns;
      const double *v2 = c2->clv ? c2->clv + (size_t)s * ncns + c * ns : c2->tipvec + (size_t)s * ns;
      double *o = out + c * ns;

      if (!amb1 && !amb2)
      { /* Exex lk.c:3377-3382 */
        for (i = 0; i < ns; ++i) o[i] = p1[i * ns + st1] * p2[i * ns + st2];
      }
      else if (amb1 && !amb2)
      { /* Exin lk.c:3398-3405 */
        matvec(ns, p1, v1, u1);
        for (i = 0; i < ns; ++i) o[i] = p2[i * ns + st2] * u1[i];
      }
      else if (!amb1 && amb2)
      {
        matvec(ns, p2, v2, u2);
        for (i = 0; i < ns; ++i) o[i] = p1[i * ns + st1] * u2[i];
      }
      else
      { /* Inin lk.c:3338-3361: all-ones inputs short-circuit to all-ones output */
        int all_one = 1;
        for (i = 0; i < ns; ++i)
          if (v1[i] != 1.0 || v2[i] != 1.0)
          {
            all_one = 0;
            break;
          }
        if (all_one)
          for (i = 0; i < ns; ++i) o[i] = 1.0;
        else
        {
          matvec(ns, p1, v1, u1);
          matvec(ns, p2, v2, u2);
          for (i = 0; i < ns; ++i) o[i] = u1[i] * u2[i];
        }
      }
    }

    /* lk.c:1744-1758: scaler bookkeeping and power-of-two rescue */
    dst_scale[s] = (c1->scale ? c1->scale[s] : 0) + (c2->scale ? c2->scale[s] : 0);
    largest = -DBL_MAX;
    for (i = 0; i < ncns; ++i)
      if (out[i] > largest) largest = out[i];
    if (largest < inv_big && apply_scaling)
    {
      for (i = 0; i < ncns; ++i) out[i] *= big;
      dst_scale[s] += ORC_LARGE;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* +I term  lk.c:1226-1273.  Returns pi[invar]*2^fact in <=63-bit steps, sets *overflow on inf. */
static double invariant_lk(int fact, int invar_state, const double *pi, int *overflow)
{
  double v = 0.0;
  *overflow = 0;
  if (invar_state > -1)
  {
    int e = fact;
    v = pi[invar_state];
    do
    {
      int piece = e < 63 ? e : 63;
      v *= (double)((unsigned long long)1 << piece);
      e -= piece;
    } while (e != 0);
    if (isinf(v)) *overflow = 1;
  }
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* K2  lk.c:608-645, 767-861, 1185-1218, 2777-2801                                             */
double plk_oracle_edge_lnl(int ns, int ncatg, int npat, const double *wght, const short *invar,
                           int invar_flag, double pinv, const double *pi, const double *rate_probs,
                           const plk_oracle_side *left, const plk_oracle_side *rght, const double *P,
                           double *site_lnl, double *site_lk_out, double *site_lk_cat,
                           int *fact_sum_scale, int *numerical_warning)
{
  const int ncns = ncatg * ns;
  double lnL = 0.0;
  int s, c, k, l;
  if (numerical_warning) *numerical_warning = 0;

  for (s = 0; s < npat; ++s)
  {
    double cat_lk[64];
    double site_lk = 0.0, log_site_lk;
    int fact;
    int amb = 1, st = -1;

    if (!(wght[s] > ORC_SMALL)) continue; /* lk.c:632 */

    if (!rght->clv)
    { /* lk.c:614-621 */
      amb = rght->is_ambigu[s];
      if (!amb) st = rght->d_state[s];
    }

    for (c = 0; c < ncatg; ++c)
    {
      const double *Pc = P + (size_t)c * ns * ns;
      const double *L = left->clv ? left->clv + (size_t)s * ncns + c * ns : left->tipvec + (size_t)s * ns;
      const double *R = rght->clv ? rght->clv + (size_t)s * ncns + c * ns : rght->tipvec + (size_t)s * ns;
      double lk = 0.0;
      if (!amb)
      { /* lk.c:1192-1199 */
        double sum = 0.0;
        for (l = 0; l < ns; ++l) sum += Pc[st * ns + l] * L[l];
        lk += sum * pi[st];
      }
      else
      { /* lk.c:1202-1214 */
        for (k = 0; k < ns; ++k)
        {
          if (R[k] > 0.0)
          {
            double sum = 0.0;
            for (l = 0; l < ns; ++l) sum += Pc[k * ns + l] * L[l];
            lk += sum * pi[k] * R[k];
          }
        }
      }
      cat_lk[c] = lk;
    }

    /* lk.c:2781-2791 */
    fact = (left->scale ? left->scale[s] : 0) + (rght->scale ? rght->scale[s] : 0);

    /* lk.c:816-818 */
    for (c = 0; c < ncatg; ++c) site_lk += cat_lk[c] * rate_probs[c];

    /* lk.c:820-842 */
    if (invar_flag)
    {
      int overflow;
      double inv = invariant_lk(fact, invar[s], pi, &overflow);
      if (overflow)
      {
        fact = 0;
        inv = invariant_lk(0, invar[s], pi, &overflow);
        site_lk = inv * pinv;
      }
      else
        site_lk = site_lk * (1. - pinv) + inv * pinv;
    }

    /* lk.c:847-857 */
    if (site_lk < ORC_SMALL)
    {
      site_lk = ORC_SMALL;
      if (numerical_warning) *numerical_warning = 1;
    }
    log_site_lk = log(site_lk) - (double)ORC_LOG2 * fact;
    lnL += wght[s] * log_site_lk;

    if (site_lnl) site_lnl[s] = log_site_lk;
    if (site_lk_out) site_lk_out[s] = exp(log_site_lk);
    if (fact_sum_scale) fact_sum_scale[s] = fact;
    if (site_lk_cat)
      for (c = 0; c < ncatg; ++c) site_lk_cat[(size_t)s * ncatg + c] = cat_lk[c];
  }
  return lnL;
}

/* ------------------------------------------------------------------------------------------ */
/* K3  lk.c:1038-1114                                                                          */
void plk_oracle_eigen_lr(int ns, int ncatg, int npat, const double *wght, const double *U,
                         const double *V, const double *pi, const plk_oracle_side *left,
                         const plk_oracle_side *rght, double *dot_prod)
{
  const int ncns = ncatg * ns;
  int s, c, i, j;
  for (s = 0; s < npat; ++s)
  {
    if (!(wght[s] > ORC_SMALL)) continue; /* lk.c:1082,1103-1112 */
    for (c = 0; c < ncatg; ++c)
    {
      const double *L = left->clv ? left->clv + (size_t)s * ncns + c * ns : left->tipvec + (size_t)s * ns;
      const double *R = rght->clv ? rght->clv + (size_t)s * ncns + c * ns : rght->tipvec + (size_t)s * ns;
      double *o = dot_prod + (size_t)s * ncns + c * ns;
      for (i = 0; i < ns; ++i)
      { /* lk.c:1086-1095 */
        double a = 0.0, b = 0.0;
        for (j = 0; j < ns; ++j)
        {
          a += U[j * ns + i] * L[j] * pi[j];
          b += V[i * ns + j] * R[j];
        }
        o[i] = a * b;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* K4  lk.c:655-753, 955-1032, 1170-1180 (with derivative); lk.c:592-603, 866-950 (without)    */
double plk_oracle_lnl_dlnl(int ns, int ncatg, int npat, const double *wght, const short *invar,
                           int invar_flag, double pinv, const double *pi, const double *rates,
                           const double *rate_probs, double br_len_mult, double l_min, double l_max,
                           const double *lambda, const double *dot_prod, const int *fact_sum_scale,
                           double *l, int with_derivative, double *dlnl_out, int *numerical_warning)
{
  const int ncns = ncatg * ns;
  double E[64 * 8], D[64 * 8];
  double lnL = 0.0, dlnL = 0.0;
  int s, c, i;
  if (numerical_warning) *numerical_warning = 0;

  if (with_derivative)
  { /* lk.c:673-674 */
    if (*l < l_min)
      *l = l_min;
    else if (*l > l_max)
      *l = l_max;
  }

  for (c = 0; c < ncatg; ++c)
  {
    double len, rr = rates[c];
    if (with_derivative)
    { /* lk.c:690-705 */
      rr *= br_len_mult;
      len = (*l) * rr;
    }
    else
    { /* lk.c:596-600 */
      len = ((*l) > 0.0 ? (*l) : 0.0) * rates[c];
      len *= br_len_mult;
    }
    if (len < l_min)
      len = l_min;
    else if (len > l_max)
      len = l_max;
    for (i = 0; i < ns; ++i)
    { /* lk.c:712-725 / lk.c:601 */
      double e = exp(lambda[i] * len);
      E[c * ns + i] = e;
      D[c * ns + i] = e * lambda[i] * rr;
    }
  }

  for (s = 0; s < npat; ++s)
  {
    const double *dp = dot_prod + (size_t)s * ncns;
    double lk = 0.0, dlk = 0.0;
    int fact = fact_sum_scale[s];
    if (!(wght[s] > ORC_SMALL)) continue;

    for (c = 0; c < ncatg; ++c)
    { /* lk.c:1170-1180, 996-997 (derivative) ; lk.c:1157-1164, 908 (no derivative) */
      double cl = 0.0, cd = 0.0;
      for (i = 0; i < ns; ++i)
      {
        cl += dp[c * ns + i] * E[c * ns + i];
        cd += dp[c * ns + i] * D[c * ns + i];
      }
      lk += cl * rate_probs[c];
      dlk += cd * rate_probs[c];
    }

    if (invar_flag)
    {
      int overflow;
      double inv = invariant_lk(fact, invar[s], pi, &overflow);
      if (with_derivative)
      { /* lk.c:1005-1025 */
        if (overflow)
        {
          lk = inv * pinv;
          dlk = 0.0;
        }
        else
        {
          lk = lk * (1. - pinv) + inv * pinv;
          dlk = dlk * (1. - pinv);
        }
      }
      else
      { /* lk.c:910-931 */
        if (overflow)
        {
          fact = 0;
          inv = invariant_lk(0, invar[s], pi, &overflow);
          lk = inv * pinv;
        }
        else
          lk = lk * (1. - pinv) + inv * pinv;
      }
    }

    if (lk < ORC_SMALL)
    { /* lk.c:1027-1031, 933-937 */
      lk = ORC_SMALL;
      if (numerical_warning) *numerical_warning = 1;
    }

    /* lk.c:742-745 / lk.c:944-946 */
    dlnL += wght[s] * (dlk / lk);
    lnL += wght[s] * (log(lk) - (double)ORC_LOG2 * fact);
  }
  if (dlnl_out) *dlnl_out = with_derivative ? dlnL : 0.0;
  return lnL;
}

/* ------------------------------------------------------------------------------------------ */
/* Parsimony: Fitch sets / Sankoff step matrices of src/pars.c.     */

/* pars.c:374-388: pars = pars_v1 + pars_v2; ui = ui_v1 & ui_v2; empty intersection => one more step, union */
void plk_oracle_pars_update(int npat, int *ui, int *pars, const int *ui_v1, const int *pars_v1,
                            const int *ui_v2, const int *pars_v2)
{
  int s;
  for (s = 0; s < npat; ++s)
  {
    int p = pars_v1[s] + pars_v2[s];
    int u = ui_v1[s] & ui_v2[s];
    if (!u)
    {
      p++;
      u = ui_v1[s] | ui_v2[s];
    }
    pars[s] = p;
    ui[s] = u;
  }
}

/* pars.c:355-372 (general_pars): p_pars[s][i] = min_j(p_pars_v1[s][j] + step[i][j]) + min_j(p_pars_v2[s][j] + step[i][j]) */
void plk_oracle_pars_update_general(int ns, int npat, const int *step_mat, int *p_pars, const int *p_pars_v1,
                                    const int *p_pars_v2)
{
  int s, i, j;
  for (s = 0; s < npat; ++s)
    for (i = 0; i < ns; ++i)
    {
      int m1 = 1000000000, m2 = 1000000000; /* MAX_PARS, utilities.h:366 */
      for (j = 0; j < ns; ++j)
      {
        int v = p_pars_v1[s * ns + j] + step_mat[i * ns + j];
        if (v < m1) m1 = v;
      }
      for (j = 0; j < ns; ++j)
      {
        int v = p_pars_v2[s * ns + j] + step_mat[i * ns + j];
        if (v < m2) m2 = v;
      }
      p_pars[s * ns + i] = m1 + m2;
    }
}

/* pars.c:20-51 (site loop of Pars) + pars.c:397-439 (Pars_Core).  c_pars is an int that accumulates
 * int * double products (`tree->c_pars += site_pars * wght`): converted back to int after every site. */
int plk_oracle_pars_edge(int general, int ns, int npat, const double *wght, const int *step_mat, const int *ui_l,
                         const int *pars_l, const int *p_pars_l, const int *ui_r, const int *pars_r,
                         const int *p_pars_r, int *site_pars)
{
  int c_pars = 0, s, i, j;
  for (s = 0; s < npat; ++s)
  {
    int sp;
    if (general)
    {
      sp = 1000000000;
      for (i = 0; i < ns; ++i)
      {
        int ml = 1000000000, mr = 1000000000;
        for (j = 0; j < ns; ++j)
        {
          int v = p_pars_l[s * ns + j] + step_mat[i * ns + j];
          if (v < ml) ml = v;
        }
        for (j = 0; j < ns; ++j)
        {
          int v = p_pars_r[s * ns + j] + step_mat[i * ns + j];
          if (v < mr) mr = v;
        }
        if (ml + mr < sp) sp = ml + mr;
      }
    }
    else
    {
      sp = pars_l[s] + pars_r[s];
      if (!(ui_l[s] & ui_r[s])) sp++;
    }
    if (site_pars) site_pars[s] = sp;
    c_pars = (int)((double)c_pars + (double)sp * wght[s]);
  }
  return c_pars;
}
