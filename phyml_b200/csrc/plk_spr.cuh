// phyml_b200/csrc/plk_spr.cuh -- batched evaluation of SPR regraft candidates.
//
// The reference scores the regraft positions of one pruned subtree one at a time (Test_One_Spr_Target,
// src/spr.c:589-650): Graft_Subtree, two Update_PMat_At_Given_Edge, one Update_Partial_Lk at the new node n_link,
// one Lk(b_arrow) -- a host round trip per candidate.  Every one of those scores depends only on state that exists
// once the subtree has been pruned (the CLVs seen from both ends of every target edge and the CLV of the pruned
// subtree), so all candidates can be scored in ONE launch:
//
//   k_spr_candidates   grid = candidates x site blocks.  Per (candidate, pattern): the CLV of the new node
//                      X = (P(l_a) . A) o (P(l_b) . B) with the 2^256 rescaling of avx.c:460-513, then the site
//                      likelihood across the pruned edge, sum_k pi_k R_k sum_l P(l_prune)_kl L_l (+I, scalers), exactly
//                      the arithmetic of one K1 update followed by one K2 site loop -- X never goes to memory.
//   k_spr_finish       per candidate: adds the per-block partial sums in block order (deterministic).
//
// Per candidate the kernel reads two CLVs (+ the pruned subtree's CLV, shared by all candidates and L2-resident) and
// writes a handful of doubles: HBM-bound, 2 x 8 x ns x ncatg bytes per (candidate, pattern).
#pragma once
#include "plk_kernels.cuh"

namespace plk
{

struct SprCandDev
{
  SideDev       a, b;    // the two ends of the target edge, each looking away from the regraft point
  const double *Pa, *Pb; // P(l_a), P(l_b): [ncatg][ns][ns]
};

constexpr int kSprThreads = 128;

// vector seen through P from one child at (site, category c): an internal CLV (any layout) or a tip mask
__device__ __forceinline__ void spr_child(const SideDev &s, const double *__restrict__ Pc, uint32_t mask, int site, int c,
                                          int ns, int ncatg, int blocked, double *__restrict__ u, bool *ones)
{
  double v[kMaxNs];
  if (s.clv)
  {
    bool o = true;
    for (int j = 0; j < ns; ++j)
    {
      v[j] = s.clv[clv_off(site, c, j, ncatg, ns, blocked)];
      o = o && (v[j] == 1.0);
    }
    *ones = o;
  }
  else
    *ones = (mask == ((ns >= 32) ? 0xffffffffu : ((1u << ns) - 1u)));
  for (int i = 0; i < ns; ++i) u[i] = child_dot(Pc + i * ns, v, mask, s.clv != nullptr, ns);
}

__global__ void __launch_bounds__(kSprThreads)
    k_spr_candidates(const SprCandDev *__restrict__ cands, int blocks_per_cand, SideDev prune,
                     const double *__restrict__ Pp, int link_on_left, const ModelDev *__restrict__ mod, int npat, int ns,
                     int ncatg, const double *__restrict__ wght, const short *__restrict__ invar,
                     const uint32_t *__restrict__ tipmask, int apply_scaling, int blocked,
                     double *__restrict__ partials, int *__restrict__ warn_out)
{
  const int        cand = blockIdx.x / blocks_per_cand, blk = blockIdx.x % blocks_per_cand;
  const SprCandDev cd = cands[cand];
  const int        nn = ns * ns;
  const double    *pi = mod->pi;
  const double     big = two_to_large(), small = inv_two_to_large();
  double           acc = 0.0;
  int              warn = 0;

  for (int site = blk * blockDim.x + threadIdx.x; site < npat; site += blocks_per_cand * blockDim.x)
  {
    const double w = wght[site];
    if (!(w > DBL_MIN)) continue;  // lk.c:632, avx.c:399
    const uint32_t ma = cd.a.clv ? 0u : tipmask[cd.a.tip[site]];
    const uint32_t mb = cd.b.clv ? 0u : tipmask[cd.b.tip[site]];
    const uint32_t mp = prune.clv ? 0u : tipmask[prune.tip[site]];
    double         ua[kMaxNs], ub[kMaxNs], x[kMaxNs];
    bool           oa, ob;

    // pass 1: the largest entry of the new node's CLV decides the 2^256 rescaling (avx.c:460-513)
    double largest = -DBL_MAX;
    for (int c = 0; c < ncatg; ++c)
    {
      spr_child(cd.a, cd.Pa + (size_t)c * nn, ma, site, c, ns, ncatg, blocked, ua, &oa);
      spr_child(cd.b, cd.Pb + (size_t)c * nn, mb, site, c, ns, ncatg, blocked, ub, &ob);
      for (int i = 0; i < ns; ++i) largest = fmax(largest, (oa && ob) ? 1.0 : ua[i] * ub[i]);
    }
    const bool rescale = (largest < small) && apply_scaling;
    const int  sx = (cd.a.scale ? cd.a.scale[site] : 0) + (cd.b.scale ? cd.b.scale[site] : 0) + (rescale ? kLarge : 0);

    // pass 2: per category, the new node's CLV again and the site likelihood across the pruned edge
    const bool right_is_tip = link_on_left && !prune.clv;
    const bool unamb = right_is_tip && (__popc(mp) == 1);  // lk.c:614-621
    const int  st = unamb ? (__ffs(mp) - 1) : -1;
    double     site_lk = 0.0;
    for (int c = 0; c < ncatg; ++c)
    {
      spr_child(cd.a, cd.Pa + (size_t)c * nn, ma, site, c, ns, ncatg, blocked, ua, &oa);
      spr_child(cd.b, cd.Pb + (size_t)c * nn, mb, site, c, ns, ncatg, blocked, ub, &ob);
      for (int i = 0; i < ns; ++i)
      {
        double o = (oa && ob) ? 1.0 : ua[i] * ub[i];
        if (rescale) o *= big;
        x[i] = o;
      }
      const double *Pc = Pp + (size_t)c * nn;
      auto          PV = [&](int k) -> double {
        return prune.clv ? prune.clv[clv_off(site, c, k, ncatg, ns, blocked)] : (double)((mp >> k) & 1u);
      };
      auto LV = [&](int l) -> double { return link_on_left ? x[l] : PV(l); };
      auto RV = [&](int k) -> double { return link_on_left ? PV(k) : x[k]; };
      double lk;
      if ((ns & 3) == 0)
      {  // order of AVX_Lk_Core_One_Class_No_Eigen_Lr (avx.c:110-215), as k_edge_lnl
        if (unamb)
        {
          double q[4] = {0.0, 0.0, 0.0, 0.0};
          for (int b4 = 0; b4 < ns; b4 += 4)
#pragma unroll
            for (int t = 0; t < 4; ++t) q[t] = q[t] + Pc[st * ns + b4 + t] * LV(b4 + t);
          lk = pi[st] * hsum4(q[0], q[1], q[2], q[3]);
        }
        else
        {
          lk = 0.0;
          for (int b4 = 0; b4 < ns; b4 += 4)
          {
            double y[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
            {
              const int    k = b4 + t;
              const double rv = RV(k);
              double       a = 0.0;
              for (int l = 0; l < ns; ++l) a = fma(Pc[k * ns + l], LV(l), a);
              y[t] = a * (rv * pi[k]);
            }
            lk = lk + hsum4(y[0], y[1], y[2], y[3]);
          }
        }
      }
      else
      {  // scalar order, lk.c:1185-1218
        lk = 0.0;
        if (unamb)
        {
          double sum = 0.0;
          for (int l = 0; l < ns; ++l) sum = sum + Pc[st * ns + l] * LV(l);
          lk = sum * pi[st];
        }
        else
          for (int k = 0; k < ns; ++k)
          {
            const double rv = RV(k);
            if (rv > 0.0)
            {
              double sum = 0.0;
              for (int l = 0; l < ns; ++l) sum = sum + Pc[k * ns + l] * LV(l);
              lk = lk + sum * pi[k] * rv;
            }
          }
      }
      site_lk = site_lk + lk * mod->probs[c];  // lk.c:818
    }
    int fact = sx + (prune.scale ? prune.scale[site] : 0);  // lk.c:2781-2791
    if (mod->invar_flag)
    {  // lk.c:820-842
      bool   ovf;
      double inv = invariant_lk(fact, invar[site], pi, &ovf);
      if (ovf)
      {
        fact = 0;
        inv = invariant_lk(0, invar[site], pi, &ovf);
        site_lk = inv * mod->pinv;
      }
      else
        site_lk = site_lk * (1. - mod->pinv) + inv * mod->pinv;
    }
    if (site_lk < DBL_MIN)
    {
      site_lk = DBL_MIN;
      warn = 1;
    }
    acc += w * (log(site_lk) - kLog2 * fact);  // lk.c:854-856
  }

  // block sum: fixed shuffle tree, fixed warp order
  __shared__ double sred[kSprThreads / 32];
  __shared__ int    swarn;
  if (threadIdx.x == 0) swarn = 0;
  __syncthreads();
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  if (warn) atomicOr(&swarn, 1);
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double s = 0.0;
    for (int wq = 0; wq < kSprThreads / 32; ++wq) s += sred[wq];
    partials[(size_t)cand * blocks_per_cand + blk] = s;
    if (swarn) atomicOr(&warn_out[cand], 1);
  }
}

// matvec4 / tipvec4 of plk_kernels.cuh with the matrix read straight from shared memory (short register lifetimes)
__device__ __forceinline__ void spr_matvec4(const double *__restrict__ p, const double4a &v, double (&u)[4])
{
#pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    double a = p[i * 4 + 0] * v.x;
    a = fma(p[i * 4 + 1], v.y, a);
    a = fma(p[i * 4 + 2], v.z, a);
    a = fma(p[i * 4 + 3], v.w, a);
    u[i] = a;
  }
}
__device__ __forceinline__ void spr_tipvec4(const double *__restrict__ p, uint32_t m, double (&u)[4])
{
#pragma unroll
  for (int i = 0; i < 4; ++i)
  {
    double a = (m & 1u) ? p[i * 4 + 0] : 0.0;
    if (m & 2u) a = a + p[i * 4 + 1];
    if (m & 4u) a = a + p[i * 4 + 2];
    if (m & 8u) a = a + p[i * 4 + 3];
    u[i] = a;
  }
}

// 4 states, 4 categories on the blocked layout: thread per (pattern, category) with the lane map of the traversal
// kernels (lane = category * 8 + pattern % 8), so a warp reads 1 KB contiguous of every CLV (256-bit loads); the three
// P-matrices of the block's candidate sit in shared memory (quarter-warp broadcasts); the new node's conditional vector
// (4 doubles per lane) never leaves registers; the per-pattern maximum (rescaling) and the sum over categories (in
// category order, as lk.c:818) are shuffles over the 4 lanes of a pattern.  Same products, FMA chains and summation
// order as k_spr_candidates, i.e. the same per-pattern values bit for bit.
__global__ void __launch_bounds__(kSprThreads)
    k_spr_candidates_dna4(const SprCandDev *__restrict__ cands, int blocks_per_cand, SideDev prune,
                          const double *__restrict__ Pp, int link_on_left, const ModelDev *__restrict__ mod, int npat,
                          const double *__restrict__ wght, const short *__restrict__ invar,
                          const uint32_t *__restrict__ tipmask, int apply_scaling, double *__restrict__ partials,
                          int *__restrict__ warn_out)
{
  constexpr int NCATG = 4;
  __shared__ double sP[3][NCATG * 16];
  __shared__ double sPi[4], sProb[NCATG];
  const int        cand = blockIdx.x / blocks_per_cand, blk = blockIdx.x % blocks_per_cand;
  const SprCandDev cd = cands[cand];
  for (int t = threadIdx.x; t < NCATG * 16; t += blockDim.x)
  {
    sP[0][t] = cd.Pa[t];
    sP[1][t] = cd.Pb[t];
    sP[2][t] = Pp[t];
  }
  if (threadIdx.x < 4) sPi[threadIdx.x] = mod->pi[threadIdx.x];
  if (threadIdx.x < NCATG) sProb[threadIdx.x] = mod->probs[threadIdx.x];
  __syncthreads();
  const double big = two_to_large(), small = inv_two_to_large();
  const int    lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int    c = lane >> 3, r = lane & 7;
  const int    n_groups = (npat + 7) >> 3;
  double       acc = 0.0;
  int          warn = 0;

  for (int grp = blk * wpb + warp; grp < n_groups; grp += blocks_per_cand * wpb)
  {
    const int    site = grp * 8 + r;
    const bool   in = site < npat;
    const double w = in ? wght[site] : 0.0;
    const bool   live = in && (w > DBL_MIN);  // lk.c:632, avx.c:399 (the 4 lanes of a pattern agree)
    const size_t off = ((size_t)grp * NCATG + c) * 32 + r * 4;
    double       x[4] = {0.0, 0.0, 0.0, 0.0};
    int          sx = 0;
    uint32_t     mp = 0u;
    if (live)
    {
      double ua[4], ub[4];
      bool   oa, ob;
      if (cd.a.clv)
      {
        const double4a v = ldg256(cd.a.clv + off);
        oa = all_one(v);
        spr_matvec4(sP[0] + c * 16, v, ua);
        sx += cd.a.scale[site];
      }
      else
      {
        const uint32_t ma = tipmask[cd.a.tip[site]];
        oa = (ma == 15u);
        spr_tipvec4(sP[0] + c * 16, ma, ua);
      }
      if (cd.b.clv)
      {
        const double4a v = ldg256(cd.b.clv + off);
        ob = all_one(v);
        spr_matvec4(sP[1] + c * 16, v, ub);
        sx += cd.b.scale[site];
      }
      else
      {
        const uint32_t mb = tipmask[cd.b.tip[site]];
        ob = (mb == 15u);
        spr_tipvec4(sP[1] + c * 16, mb, ub);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = (oa && ob) ? 1.0 : ua[i] * ub[i];
      if (!prune.clv) mp = tipmask[prune.tip[site]];
    }
    // largest entry of the new node's CLV over its 4 states x 4 categories (avx.c:460-513)
    double largest = fmax(fmax(x[0], x[1]), fmax(x[2], x[3]));
    largest = fmax(largest, __shfl_xor_sync(0xffffffffu, largest, 8));
    largest = fmax(largest, __shfl_xor_sync(0xffffffffu, largest, 16));
    const bool rescale = (largest < small) && apply_scaling;
    if (rescale)
    {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] *= big;
      sx += kLarge;
    }
    double term = 0.0;
    if (live)
    {
      double pv[4];
      if (prune.clv)
      {
        const double4a v = ldg256(prune.clv + off);
        pv[0] = v.x, pv[1] = v.y, pv[2] = v.z, pv[3] = v.w;
      }
      else
      {
#pragma unroll
        for (int k = 0; k < 4; ++k) pv[k] = (double)((mp >> k) & 1u);
      }
      double L[4], R[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
      {
        L[k] = link_on_left ? x[k] : pv[k];
        R[k] = link_on_left ? pv[k] : x[k];
      }
      const double *Pc = sP[2] + c * 16;
      const bool    unamb = link_on_left && !prune.clv && (__popc(mp) == 1);  // lk.c:614-621
      double        lk;
      if (unamb)
      {
        const int st = __ffs(mp) - 1;
        double    q[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) q[t] = 0.0 + Pc[st * 4 + t] * L[t];
        lk = sPi[st] * hsum4(q[0], q[1], q[2], q[3]);
      }
      else
      {
        double y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
          double a = 0.0;
#pragma unroll
          for (int l = 0; l < 4; ++l) a = fma(Pc[k * 4 + l], L[l], a);
          y[k] = a * (R[k] * sPi[k]);
        }
        lk = 0.0 + hsum4(y[0], y[1], y[2], y[3]);
      }
      term = lk * sProb[c];
    }
    // site_lk = ((0 + t0) + t1) + t2) + t3, lk.c:818
    const double t1 = __shfl_sync(0xffffffffu, term, r + 8);
    const double t2 = __shfl_sync(0xffffffffu, term, r + 16);
    const double t3 = __shfl_sync(0xffffffffu, term, r + 24);
    if (live && c == 0)
    {
      double site_lk = 0.0 + term;
      site_lk = site_lk + t1;
      site_lk = site_lk + t2;
      site_lk = site_lk + t3;
      int fact = sx + (prune.scale ? prune.scale[site] : 0);
      if (mod->invar_flag)
      {
        bool   ovf;
        double inv = invariant_lk(fact, invar[site], mod->pi, &ovf);
        if (ovf)
        {
          fact = 0;
          inv = invariant_lk(0, invar[site], mod->pi, &ovf);
          site_lk = inv * mod->pinv;
        }
        else
          site_lk = site_lk * (1. - mod->pinv) + inv * mod->pinv;
      }
      if (site_lk < DBL_MIN)
      {
        site_lk = DBL_MIN;
        warn = 1;
      }
      acc += w * (log(site_lk) - kLog2 * fact);
    }
  }

  __shared__ double sred[kSprThreads / 32];
  __shared__ int    swarn;
  if (threadIdx.x == 0) swarn = 0;
  __syncthreads();
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  if (warn) atomicOr(&swarn, 1);
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double s = 0.0;
    for (int wq = 0; wq < kSprThreads / 32; ++wq) s += sred[wq];
    partials[(size_t)cand * blocks_per_cand + blk] = s;
    if (swarn) atomicOr(&warn_out[cand], 1);
  }
}

// child_dot with the state count known at compile time (the vector stays in registers)
template <int NS>
__device__ __forceinline__ double spr_dot(const double *__restrict__ Prow, const double (&v)[NS], uint32_t mask, bool internal)
{
  if (internal)
  {
    double a = Prow[0] * v[0];
#pragma unroll
    for (int j = 1; j < NS; ++j) a = fma(Prow[j], v[j], a);
    return a;
  }
  double a = (mask & 1u) ? Prow[0] : 0.0;
#pragma unroll
  for (int j = 1; j < NS; ++j)
    if ((mask >> j) & 1u) a = a + Prow[j];
  return a;
}

// NS states (a multiple of 4; 20 in practice), 4 categories on the blocked layout: thread per (pattern, category), lane
// = category * 8 + pattern % 8 as above; the thread's NS-state vectors are NS/4 256-bit loads (a quarter-warp reads a
// contiguous 256-byte block per load), both conditional vectors and the new node's vector stay in registers, the three
// P-matrices of the block's candidate sit in shared memory with a 2-double pad per category (the 4 quarter-warps of a
// broadcast load hit different banks).  Same arithmetic as k_spr_candidates.
template <int NS>
__global__ void __launch_bounds__(kSprThreads)
    k_spr_candidates_reg4(const SprCandDev *__restrict__ cands, int blocks_per_cand, SideDev prune,
                          const double *__restrict__ Pp, int link_on_left, const ModelDev *__restrict__ mod, int npat,
                          const double *__restrict__ wght, const short *__restrict__ invar,
                          const uint32_t *__restrict__ tipmask, int apply_scaling, double *__restrict__ partials,
                          int *__restrict__ warn_out)
{
  constexpr int NCATG = 4, NN = NS * NS, PS = NN + 2, KB = NS / 4;
  extern __shared__ double spr_sm[];  // [3][NCATG][PS] | pi[NS] | probs[NCATG]
  double *sP = spr_sm, *sPi = spr_sm + 3 * NCATG * PS, *sProb = sPi + NS;
  const int        cand = blockIdx.x / blocks_per_cand, blk = blockIdx.x % blocks_per_cand;
  const SprCandDev cd = cands[cand];
  for (int t = threadIdx.x; t < NCATG * NN; t += blockDim.x)
  {
    const int cc = t / NN, e = t % NN;
    sP[(0 * NCATG + cc) * PS + e] = cd.Pa[t];
    sP[(1 * NCATG + cc) * PS + e] = cd.Pb[t];
    sP[(2 * NCATG + cc) * PS + e] = Pp[t];
  }
  if (threadIdx.x < NS) sPi[threadIdx.x] = mod->pi[threadIdx.x];
  if (threadIdx.x < NCATG) sProb[threadIdx.x] = mod->probs[threadIdx.x];
  __syncthreads();
  const double   big = two_to_large(), small = inv_two_to_large();
  const uint32_t full = (NS >= 32) ? 0xffffffffu : ((1u << NS) - 1u);
  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int      c = lane >> 3, r = lane & 7;
  const int      n_groups = (npat + 7) >> 3;
  double         acc = 0.0;
  int            warn = 0;

  for (int grp = blk * wpb + warp; grp < n_groups; grp += blocks_per_cand * wpb)
  {
    const int    site = grp * 8 + r;
    const bool   in = site < npat;
    const double w = in ? wght[site] : 0.0;
    const bool   live = in && (w > DBL_MIN);
    const size_t off = (((size_t)grp * NCATG + c) * KB * 8 + r) * 4;  // clv_off(site, c, 0): + kb * 32 per 4 states
    double       x[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) x[i] = 0.0;
    int      sx = 0;
    uint32_t mp = 0u;
    if (live)
    {
      double   v[NS];
      bool     oa = true, ob = true;
      uint32_t m = 0u;
      // child a
      if (cd.a.clv)
      {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
        {
          const double4a q = ldg256(cd.a.clv + off + kb * 32);
          v[kb * 4 + 0] = q.x, v[kb * 4 + 1] = q.y, v[kb * 4 + 2] = q.z, v[kb * 4 + 3] = q.w;
          oa = oa && all_one(q);
        }
        sx += cd.a.scale[site];
      }
      else
      {
        m = tipmask[cd.a.tip[site]];
        oa = (m == full);
      }
      {
        const double *P = sP + (0 * NCATG + c) * PS;
#pragma unroll
        for (int i = 0; i < NS; ++i) x[i] = spr_dot<NS>(P + i * NS, v, m, cd.a.clv != nullptr);
      }
      // child b
      m = 0u;
      if (cd.b.clv)
      {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
        {
          const double4a q = ldg256(cd.b.clv + off + kb * 32);
          v[kb * 4 + 0] = q.x, v[kb * 4 + 1] = q.y, v[kb * 4 + 2] = q.z, v[kb * 4 + 3] = q.w;
          ob = ob && all_one(q);
        }
        sx += cd.b.scale[site];
      }
      else
      {
        m = tipmask[cd.b.tip[site]];
        ob = (m == full);
      }
      {
        const double *P = sP + (1 * NCATG + c) * PS;
        const bool    ones = oa && ob;
#pragma unroll
        for (int i = 0; i < NS; ++i)
        {
          const double ub = spr_dot<NS>(P + i * NS, v, m, cd.b.clv != nullptr);
          x[i] = ones ? 1.0 : x[i] * ub;
        }
      }
      if (!prune.clv) mp = tipmask[prune.tip[site]];
    }
    double largest = x[0];
#pragma unroll
    for (int i = 1; i < NS; ++i) largest = fmax(largest, x[i]);
    largest = fmax(largest, __shfl_xor_sync(0xffffffffu, largest, 8));
    largest = fmax(largest, __shfl_xor_sync(0xffffffffu, largest, 16));
    const bool rescale = (largest < small) && apply_scaling;
    if (rescale)
    {
#pragma unroll
      for (int i = 0; i < NS; ++i) x[i] *= big;
      sx += kLarge;
    }
    double term = 0.0;
    if (live)
    {
      double pv[NS];
      if (prune.clv)
      {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
        {
          const double4a q = ldg256(prune.clv + off + kb * 32);
          pv[kb * 4 + 0] = q.x, pv[kb * 4 + 1] = q.y, pv[kb * 4 + 2] = q.z, pv[kb * 4 + 3] = q.w;
        }
      }
      else
      {
#pragma unroll
        for (int k = 0; k < NS; ++k) pv[k] = (double)((mp >> k) & 1u);
      }
      if (!link_on_left)
      {  // the new node is the right-hand side: swap roles (L = pruned subtree, R = new node)
#pragma unroll
        for (int k = 0; k < NS; ++k)
        {
          const double tmp = x[k];
          x[k] = pv[k];
          pv[k] = tmp;
        }
      }
      // from here: L = x, R = pv
      const double *Pc = sP + (2 * NCATG + c) * PS;
      const bool    unamb = link_on_left && !prune.clv && (__popc(mp) == 1);
      double        lk;
      if (unamb)
      {  // avx.c:110-215
        const int st = __ffs(mp) - 1;
        double    q[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int b4 = 0; b4 < NS; b4 += 4)
#pragma unroll
          for (int t = 0; t < 4; ++t) q[t] = q[t] + Pc[st * NS + b4 + t] * x[b4 + t];
        lk = sPi[st] * hsum4(q[0], q[1], q[2], q[3]);
      }
      else
      {
        lk = 0.0;
#pragma unroll
        for (int b4 = 0; b4 < NS; b4 += 4)
        {
          double y[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
          {
            const int k = b4 + t;
            double    a = 0.0;
#pragma unroll
            for (int l = 0; l < NS; ++l) a = fma(Pc[k * NS + l], x[l], a);
            y[t] = a * (pv[k] * sPi[k]);
          }
          lk = lk + hsum4(y[0], y[1], y[2], y[3]);
        }
      }
      term = lk * sProb[c];
    }
    const double t1 = __shfl_sync(0xffffffffu, term, r + 8);
    const double t2 = __shfl_sync(0xffffffffu, term, r + 16);
    const double t3 = __shfl_sync(0xffffffffu, term, r + 24);
    if (live && c == 0)
    {
      double site_lk = 0.0 + term;
      site_lk = site_lk + t1;
      site_lk = site_lk + t2;
      site_lk = site_lk + t3;
      int fact = sx + (prune.scale ? prune.scale[site] : 0);
      if (mod->invar_flag)
      {
        bool   ovf;
        double inv = invariant_lk(fact, invar[site], mod->pi, &ovf);
        if (ovf)
        {
          fact = 0;
          inv = invariant_lk(0, invar[site], mod->pi, &ovf);
          site_lk = inv * mod->pinv;
        }
        else
          site_lk = site_lk * (1. - mod->pinv) + inv * mod->pinv;
      }
      if (site_lk < DBL_MIN)
      {
        site_lk = DBL_MIN;
        warn = 1;
      }
      acc += w * (log(site_lk) - kLog2 * fact);
    }
  }

  __shared__ double sred[kSprThreads / 32];
  __shared__ int    swarn;
  if (threadIdx.x == 0) swarn = 0;
  __syncthreads();
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  if (warn) atomicOr(&swarn, 1);
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double s = 0.0;
    for (int wq = 0; wq < kSprThreads / 32; ++wq) s += sred[wq];
    partials[(size_t)cand * blocks_per_cand + blk] = s;
    if (swarn) atomicOr(&warn_out[cand], 1);
  }
}

__global__ void k_spr_finish(const double *__restrict__ partials, int blocks_per_cand, int n_cand, double *__restrict__ lnl)
{
  const int cand = blockIdx.x * blockDim.x + threadIdx.x;
  if (cand >= n_cand) return;
  double s = 0.0;
  for (int b = 0; b < blocks_per_cand; ++b) s += partials[(size_t)cand * blocks_per_cand + b];
  lnl[cand] = s;
}

}  // namespace plk
