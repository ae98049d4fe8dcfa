// phyml_b200/csrc/plk_pars.cuh -- sm_100a kernels for the parsimony scores of src/pars.c (SURVEY.md section 8f,
// row 4: the SPR pre-filter), integer work, bit-exact.
//
//   k_pars_fitch     Update_Partial_Pars, Fitch branch (pars.c:374-388) for a whole dependency-ordered list of
//                    updates in ONE launch + optionally the site loop of Pars / Pars_Core (pars.c:40-48,434-436)
//   k_pars_sankoff   the same for the step-matrix variant (general_pars, pars.c:355-372,409-431)
//   k_pars_chain     c_pars for non-integral pattern weights (the reference truncates to int after every site)
//
// Like the likelihood traversal kernels, a thread keeps the same patterns for every update of the list and only
// re-reads words it wrote itself, so the list needs no grid-wide synchronisation between tree levels.  A Fitch
// buffer is one int2 {ui = state-set bit mask, pars = steps below} per pattern: an (int,int) update moves 24 bytes
// per pattern, and the result of the previous update is forwarded in registers (in post-order it is the most
// recently written child).  HBM/L2-latency bound: ~50 000 patterns x 24 B is far below one wave of the machine,
// the cost of a list is its dependent-latency chain, which is why the whole list is one launch.
#pragma once
#include "plk_kernels.cuh"

namespace plk
{

constexpr int kParsOpsSmem = 1024;        // update descriptors staged per chunk (24 KB of shared memory)
constexpr int kParsThreads = 256;
constexpr int kMaxPars = 1000000000;      // MAX_PARS, utilities.h:366

struct ParsOpDev
{
  void       *dst;  // Fitch: int2[npat]; Sankoff: int[ns][pstride]
  const void *c1;
  const void *c2;
};

// the site loop of Pars at one edge, run as the epilogue of the traversal kernels
struct ParsEdgeDev
{
  const void   *left, *rght;
  const double *wght;
  int          *site_pars;
  int           mode;  // 0: none; 1: site_pars + weighted sum (exact for integral weights); 2: site_pars only
  ReduceOut     ro;
};

__device__ __forceinline__ int2 fitch_join(int2 a, int2 b)
{
  int2 r;
  r.y = a.y + b.y;
  r.x = a.x & b.x;
  if (!r.x)
  {
    r.y++;
    r.x = a.x | b.x;
  }
  return r;
}

// U = patterns per thread (independent dependency chains in flight)
template <int U>
__global__ void __launch_bounds__(kParsThreads) k_pars_fitch(const ParsOpDev *__restrict__ ops, int n_ops, int npat,
                                                             ParsEdgeDev edge)
{
  __shared__ ParsOpDev s_ops[kParsOpsSmem];
  const int T = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  int       site[U];
  bool      ok[U];
  int2      fwd[U];
#pragma unroll
  for (int k = 0; k < U; ++k)
  {
    site[k] = t0 + k * T;
    ok[k] = site[k] < npat;
    fwd[k] = make_int2(0, 0);
  }
  const void *fwd_ptr = nullptr;
  for (int base = 0; base < n_ops; base += kParsOpsSmem)
  {
    const int n = min(kParsOpsSmem, n_ops - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_ops[i] = ops[base + i];
    __syncthreads();
    for (int o = 0; o < n; ++o)
    {
      const ParsOpDev op = s_ops[o];
      const bool      f1 = (op.c1 == fwd_ptr), f2 = (op.c2 == fwd_ptr);
      int2            a[U], b[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
      {
        a[k] = fwd[k];
        b[k] = fwd[k];
        if (ok[k] && !f1) a[k] = ((const int2 *)op.c1)[site[k]];
        if (ok[k] && !f2) b[k] = ((const int2 *)op.c2)[site[k]];
      }
#pragma unroll
      for (int k = 0; k < U; ++k)
      {
        fwd[k] = fitch_join(a[k], b[k]);
        if (ok[k]) ((int2 *)op.dst)[site[k]] = fwd[k];
      }
      fwd_ptr = op.dst;
    }
  }
  if (edge.mode == 0) return;
  double acc[1] = {0.0};
#pragma unroll
  for (int k = 0; k < U; ++k)
    if (ok[k])
    {
      const int2 l = (edge.left == fwd_ptr) ? fwd[k] : ((const int2 *)edge.left)[site[k]];
      const int2 r = (edge.rght == fwd_ptr) ? fwd[k] : ((const int2 *)edge.rght)[site[k]];
      const int  sp = l.y + r.y + ((l.x & r.x) ? 0 : 1);
      edge.site_pars[site[k]] = sp;
      acc[0] += (double)sp * edge.wght[site[k]];
    }
  if (edge.mode == 1) block_reduce_finish<1>(acc, 0, edge.ro);
}

// Step-matrix parsimony: buffers are state-major, p[j * pstride + pattern].  NS_T = 0: ns given at run time (<= 32).
template <int NS_T>
__global__ void __launch_bounds__(128) k_pars_sankoff(const ParsOpDev *__restrict__ ops, int n_ops, int npat,
                                                      size_t pstride, int ns_rt, const int *__restrict__ step_mat,
                                                      ParsEdgeDev edge)
{
  constexpr int NSM = NS_T ? NS_T : kMaxNs;
  __shared__ int s_step[kMaxNs * kMaxNs];
  __shared__ ParsOpDev s_ops[256];
  const int ns = NS_T ? NS_T : ns_rt;
  const int T = gridDim.x * blockDim.x;
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = threadIdx.x; i < ns * ns; i += blockDim.x) s_step[i] = step_mat[i];
  for (int base = 0; base < n_ops; base += 256)
  {
    const int n = min(256, n_ops - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_ops[i] = ops[base + i];
    __syncthreads();
    for (int o = 0; o < n; ++o)
    {
      const int *c1 = (const int *)s_ops[o].c1, *c2 = (const int *)s_ops[o].c2;
      int       *dst = (int *)s_ops[o].dst;
      for (int s = t0; s < npat; s += T)
      {
        int v1[NSM], v2[NSM];
#pragma unroll
        for (int j = 0; j < NSM; ++j)
          if (j < ns)
          {
            v1[j] = c1[(size_t)j * pstride + s];
            v2[j] = c2[(size_t)j * pstride + s];
          }
        for (int i = 0; i < ns; ++i)
        {
          int m1 = kMaxPars, m2 = kMaxPars;
#pragma unroll
          for (int j = 0; j < NSM; ++j)
            if (j < ns)
            {
              const int st = s_step[i * ns + j];
              m1 = min(m1, v1[j] + st);
              m2 = min(m2, v2[j] + st);
            }
          dst[(size_t)i * pstride + s] = m1 + m2;
        }
      }
    }
  }
  if (edge.mode == 0) return;
  __syncthreads();
  double acc[1] = {0.0};
  const int *pl = (const int *)edge.left, *pr = (const int *)edge.rght;
  for (int s = t0; s < npat; s += T)
  {
    int v1[NSM], v2[NSM];
#pragma unroll
    for (int j = 0; j < NSM; ++j)
      if (j < ns)
      {
        v1[j] = pl[(size_t)j * pstride + s];
        v2[j] = pr[(size_t)j * pstride + s];
      }
    int sp = kMaxPars;
    for (int i = 0; i < ns; ++i)
    {
      int m1 = kMaxPars, m2 = kMaxPars;
#pragma unroll
      for (int j = 0; j < NSM; ++j)
        if (j < ns)
        {
          const int st = s_step[i * ns + j];
          m1 = min(m1, v1[j] + st);
          m2 = min(m2, v2[j] + st);
        }
      sp = min(sp, m1 + m2);
    }
    edge.site_pars[s] = sp;
    acc[0] += (double)sp * edge.wght[s];
  }
  if (edge.mode == 1) block_reduce_finish<1>(acc, 0, edge.ro);
}

// c_pars exactly as the reference accumulates it (pars.c:46: an int += int * double, i.e. truncated after every
// pattern), needed only when some pattern weight is not an integer.  One warp: coalesced loads, lane 0 runs the chain.
__global__ void k_pars_chain(const int *__restrict__ site_pars, const double *__restrict__ wght, int npat, int c0,
                             ReduceOut ro)
{
  const int lane = threadIdx.x;
  int       c = c0;
  for (int base = 0; base < npat; base += 32)
  {
    const int    s = base + lane;
    const int    sp = (s < npat) ? site_pars[s] : 0;
    const double w = (s < npat) ? wght[s] : 0.0;
    const int    n = min(32, npat - base);
    for (int k = 0; k < n; ++k)
    {
      const int    spk = __shfl_sync(0xffffffffu, sp, k);
      const double wk = __shfl_sync(0xffffffffu, w, k);
      c = (int)((double)c + (double)spk * wk);
    }
  }
  if (lane == 0)
  {
    ro.dev_out[0] = (double)c;
    ro.dev_out[1] = 0.0;
    ro.dev_out[2] = 0.0;
    ro.host_out->val[0] = (double)c;
    ro.host_out->val[1] = 0.0;
    ro.host_out->warn = 0;
    __threadfence_system();
    ro.host_out->seq = ro.seq;
  }
}

}  // namespace plk
