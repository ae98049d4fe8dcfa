// phyml_b200/csrc/plk_kernels.cuh -- sm_100a kernels of the likelihood engine.
//
// K0  k_pmat            batched P(t) = U diag(exp(lambda t r_c)) V        (models.c:257-326, lk.c:2238-2325)
// K1  k_traverse_dna    fused CLV updates for a whole op list, 4 states: 256-bit LDG/STG, op descriptors and
//                       P-matrices streamed by a producer warp through a TMA/mbarrier ring (avx.c:301-522)
//     k_partial_generic any ns <= 32 / any ncatg                          (lk.c:1659-1768)
// K2  k_edge_lnl        edge log-likelihood + per-site by-products        (lk.c:605-645, 767-861, 2777-2801)
// K3  k_eigen_lr        eigen-basis projection dot_prod                   (lk.c:1038-1114, avx.c:21-105)
// K4  k_lnl_dlnl        lnL and dlnL/dl from dot_prod                     (lk.c:655-753, 955-1032)
//     (K2/K4 end with a deterministic last-block reduction that publishes to mapped pinned memory)
//
// Arithmetic order: where it is free, sums are accumulated in the order of the reference's AVX+FMA
// kernels (first term a plain product, then an FMA chain in ascending state order; horizontal sums
// as (x0+x2)+(x1+x3)), so CLVs and scalers come out bit-identical to the reference's AVX build for
// identical inputs.  The file is compiled with -fmad=false: every fma() below is explicit.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace plk
{

constexpr int kMaxNs = 32;
constexpr int kMaxCatg = 16;
constexpr double kLog2 = 0.69314718055994528623;  // utilities.h:267
constexpr double kSmallPij = 1.E-100;              // utilities.h:478
constexpr int kLarge = 256;                        // utilities.h:507

// 2^256 and 2^-256 exactly (utilities.h:508-520)
__device__ __forceinline__ double two_to_large() { return __longlong_as_double(0x4FF0000000000000LL); }
__device__ __forceinline__ double inv_two_to_large() { return __longlong_as_double(0x2FF0000000000000LL); }

// CLV addressing.  Plain layout (any ns): [site][catg][state], the reference's (lk.c:1474).
// Blocked layout (ns = 4 or 20, internal to the engine; plk_get_clv/plk_set_clv convert):
//   [tile = site/8][catg][kb = state/4][r = site%8][t = state%4]
// so the 8 sites x 4 states that one MMA fragment load/store touches are one contiguous 256-byte block
// and a (site, catg) item of a 4-state CLV is still one 32-byte word.
__host__ __device__ __forceinline__ size_t clv_off(int site, int c, int i, int ncatg, int ns, int blocked)
{
  if (blocked)
    return (((((size_t)(site >> 3) * ncatg + c) * (ns >> 2) + (i >> 2)) * 8 + (site & 7)) << 2) + (i & 3);
  return ((size_t)site * ncatg + c) * ns + i;
}

// plain <-> blocked conversion of one CLV (read-back / upload paths)
__global__ void k_clv_convert(const double *__restrict__ src, double *__restrict__ dst, int npat, int ncatg, int ns,
                              int to_blocked)
{
  const size_t total = (size_t)npat * ncatg * ns;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
  {
    const int    i = (int)(e % ns), c = (int)((e / ns) % ncatg), site = (int)(e / ((size_t)ns * ncatg));
    const size_t b = clv_off(site, c, i, ncatg, ns, 1);
    if (to_blocked)
      dst[b] = src[e];
    else
      dst[e] = src[b];
  }
}

struct __align__(32) double4a
{
  double x, y, z, w;
};

// Model block in device memory (uploaded by plk_set_model)
struct ModelDev
{
  double U[kMaxNs * kMaxNs];
  double V[kMaxNs * kMaxNs];
  double lambda[kMaxNs];
  double pi[kMaxNs];
  double rates[kMaxCatg];
  double probs[kMaxCatg];
  double pinv, l_min, l_max, br_len_mult;
  int    invar_flag;
  int    pad;
};

// One CLV update with every handle resolved to device pointers.
struct OpDev
{
  double        *dst;
  int           *dst_scale;
  const double  *c1;  // internal child CLV or nullptr
  const int     *s1;
  const uint8_t *t1;  // tip codes or nullptr
  const double  *c2;
  const int     *s2;
  const uint8_t *t2;
  const double  *P1;
  const double  *P2;
  int            flags;  // fused kernel: operand sources, kSrc* of child 1 | kSrc* of child 2 << 2
  int            pad[3];
};
static_assert(sizeof(OpDev) == 96, "OpDev must be a multiple of 16 bytes for cp.async.bulk");
// Operand sources of the fused traversal kernel.  The host canonicalises every update (the product
// of the two children commutes exactly) to one of (FWD,TIP) (FWD,SLOT) (SLOT,TIP) (SLOT,LATE) (TIP,TIP):
//   TIP  tip: P1/P2 points at the edge's tip table TP[cat][mask][state] written by k_pmat
//   FWD  the destination of the previous update of the list, still in registers
//   SLOT a CLV in global memory, prefetched one update ahead
//   LATE a second CLV in global memory, loaded inside the update
constexpr int kSrcTip = 0, kSrcFwd = 1, kSrcSlot = 2, kSrcLate = 3;

// 4-state tip tables are stored by ROW = kTipRow[mask]: the unambiguous masks 1,2,4,8 get rows 0..3 so
// that the 32-byte rows read by the lanes of a quarter-warp fall into different shared-memory banks.
__host__ __device__ __forceinline__ int tip_row4(int mask)
{
  // rows:      mask 0->15, 1->0, 2->1, 3->4, 4->2, 5->5, 6->6, 7->7, 8->3, 9..14 -> 8..13, 15->14
  return (int)((0xEDCBA9837652410FULL >> (4 * (mask & 15))) & 15ULL);
}
constexpr int kTipRowAllOnes = 14;

// One side of an edge (K2/K3)
struct SideDev
{
  const double  *clv;
  const int     *scale;
  const uint8_t *tip;
};

// ------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk-copy helpers (cp.async.bulk -> SASS UBLKCP): stage the P-matrices of an op
// in shared memory.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// producer-side wait: backs off between probes so the spinning lane does not eat issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity)
{
  uint32_t done = 0;
  while (true)
  {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// shared-memory progress counters between warps of a block (release / acquire at CTA scope)
__device__ __forceinline__ int lds32_volatile(uint32_t a)
{
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// K0: one block per (P-matrix, rate category); one thread per matrix entry.
struct PmatJob
{
  double *P;  // [ncatg][ns][ns]
  double  l;  // b->l->v
};

__device__ __forceinline__ void pmat_body(double *__restrict__ Pout, double l, const ModelDev *__restrict__ mod, int ns,
                                          int ncatg, int with_tip_table)
{
  extern __shared__ double sm[];  // expt[ns] | raw / normalised P of this category: [ns*ns]
  double *expt = sm;
  double *raw = sm + kMaxNs;
  const int c = blockIdx.x % ncatg;
  const int e = threadIdx.x;
  const int nn = ns * ns;

  // lk.c:2296-2300
  double len = fmax(0.0, l) * mod->rates[c];
  len = len * mod->br_len_mult;
  if (len < mod->l_min)
    len = mod->l_min;
  else if (len > mod->l_max)
    len = mod->l_max;

  if (e < ns) expt[e] = exp(mod->lambda[e] * len);  // models.c:275
  __syncthreads();
  double acc = 0.0;
  if (e < nn)
  {
    const int i = e / ns, j = e % ns;
    for (int k = 0; k < ns; ++k) acc = fma(mod->U[i * ns + k] * expt[k], mod->V[k * ns + j], acc);  // models.c:279,291
    if (acc < kSmallPij) acc = kSmallPij;                                                           // models.c:293
    raw[e] = acc;
  }
  __syncthreads();
  double pn = 0.0;
  if (e < nn)
  {
    const int i = e / ns;
    double    sum = 0.0;
    for (int j = 0; j < ns; ++j) sum = sum + raw[i * ns + j];  // models.c:296-297
    pn = raw[e] / sum;                                          // models.c:298
    Pout[(size_t)c * nn + e] = pn;
  }
  if (with_tip_table == 2)
  {  // ns == 20: tPx[c][s][i] = P[c][i][s] for s < 20 (what an unambiguous tip in state s contributes) and
     // tPx[c][20][i] = sum_j P[c][i][j] in ascending j (a fully ambiguous tip: gap / X / ?)
    __syncthreads();
    if (e < nn) raw[e] = pn;
    __syncthreads();
    double *TX = Pout + (size_t)ncatg * nn + (size_t)c * 420;
    if (e < nn) TX[(e % ns) * ns + (e / ns)] = pn;
    if (e < ns)
    {
      double a = raw[e * ns];
      for (int j = 1; j < ns; ++j) a = a + raw[e * ns + j];
      TX[20 * ns + e] = a;
    }
    // Pf[c][j = n*5+kk][lane = g*4+t] = P[c][8n+g][4kk+t] (0 for rows >= 20): the B fragments of
    // k_traverse_aa in the order its lanes read them (conflict-free LDS.64, no address arithmetic)
    double *PF = Pout + (size_t)ncatg * (nn + 420) + (size_t)c * 480;
    for (int q = e; q < 480; q += blockDim.x)
    {
      const int j = q >> 5, ln = q & 31, n = j / 5, kk = j % 5, gg = ln >> 2, tt = ln & 3;
      PF[q] = (n * 8 + gg < 20) ? raw[(n * 8 + gg) * 20 + kk * 4 + tt] : 0.0;
    }
  }
  else if (with_tip_table)
  {  // ns == 4: TP[c][tip_row4(mask)][i] = sum_{j in mask} P[c][i][j], ascending j (a tip child's vector)
    __syncthreads();
    if (e < nn) raw[e] = pn;
    __syncthreads();
    double *TP = Pout + (size_t)ncatg * nn + (size_t)c * 64;
    for (int t = e; t < 64; t += blockDim.x)
    {
      const int m = t >> 2, i = t & 3;
      double    a = (m & 1) ? raw[i * 4 + 0] : 0.0;
      if (m & 2) a = a + raw[i * 4 + 1];
      if (m & 4) a = a + raw[i * 4 + 2];
      if (m & 8) a = a + raw[i * 4 + 3];
      TP[tip_row4(m) * 4 + i] = a;
    }
  }
}

__global__ void k_pmat(const PmatJob *__restrict__ jobs, const ModelDev *__restrict__ mod, int ns, int ncatg,
                       int with_tip_table)
{
  const int job = blockIdx.x / ncatg;
  pmat_body(jobs[job].P, jobs[job].l, mod, ns, ncatg, with_tip_table);
}

// the same with the job list in the kernel's parameter block (no staging copy in front of the launch): up to
// kPmatInline matrices, the common case of one full-tree evaluation (2n - 3 edges)
constexpr int kPmatInline = 1024;
struct PmatJobsInline
{
  double *base;               // P-matrix record 0
  unsigned stride;            // doubles between records
  int     h[kPmatInline];     // record index
  double  l[kPmatInline];     // branch length
};
__global__ void k_pmat_inline(const __grid_constant__ PmatJobsInline jobs, const ModelDev *__restrict__ mod, int ns,
                              int ncatg, int with_tip_table)
{
  const int job = blockIdx.x / ncatg;
  pmat_body(jobs.base + (size_t)jobs.h[job] * jobs.stride, jobs.l[job], mod, ns, ncatg, with_tip_table);
}

// the same for a handful of matrices (one SPR candidate / one Lk(b): 1 to 3): a 112-byte parameter block instead of 12 KB
constexpr int kPmatInlineSmall = 8;
struct PmatJobsInlineSmall
{
  double  *base;
  unsigned stride;
  int      h[kPmatInlineSmall];
  double   l[kPmatInlineSmall];
};
__global__ void k_pmat_inline_small(const __grid_constant__ PmatJobsInlineSmall jobs, const ModelDev *__restrict__ mod,
                                    int ns, int ncatg, int with_tip_table)
{
  const int job = blockIdx.x / ncatg;
  pmat_body(jobs.base + (size_t)jobs.h[job] * jobs.stride, jobs.l[job], mod, ns, ncatg, with_tip_table);
}

// ------------------------------------------------------------------------------------------------
// K1, 4 states.  Thread <-> (site, category): its 4 states are one 32-byte vector, so consecutive
// threads stream consecutive 32-byte words of each CLV (256-bit LDG/STG, fully coalesced).
// The per-site max over all ncatg*4 entries is a shuffle over the NCATG neighbouring lanes.
__device__ __forceinline__ double4a ld256(const double *p) { return *reinterpret_cast<const double4a *>(p); }
__device__ __forceinline__ void     st256(double *p, const double4a &v) { *reinterpret_cast<double4a *>(p) = v; }

// u = P . v in the order of AVX_Matrix_Vect_Prod (avx.c:593-616): first column a product, then FMAs
__device__ __forceinline__ void matvec4(const double (&p)[16], const double4a &v, double (&u)[4])
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

// tip child: v is a 0/1 vector given as a bit mask; same order, products with 0/1 are exact
__device__ __forceinline__ void tipvec4(const double (&p)[16], uint32_t m, double (&u)[4])
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

// ------------------------------------------------------------------------------------------------
// K1 fused traversal, 4 states: ONE launch executes a whole dependency-ordered list of updates
// (a post-order / pre-order traversal, or a single Update_Partial_Lk).  Site patterns are
// independent, so each thread keeps the same (site, category) items for every update of the list:
// an update only ever reads CLV words that the same thread wrote earlier in the list, i.e. no
// grid-wide synchronisation is needed between tree levels, and freshly written children are
// re-read from L2 instead of HBM.  Scalers are written by the category-0 lane of a site and read by
// its sibling lanes: __syncwarp() after every update orders that.
//
// Warp roles: 8 compute warps + 1 producer warp.  The producer streams, S updates ahead, the
// 80-byte update descriptor and the two ncatg x 4 x 4 P-matrices of each update into a
// shared-memory ring with TMA bulk copies (cp.async.bulk -> UBLKCP) signalled on full[]
// mbarriers; compute warps release a slot by arriving on empty[].
__device__ __forceinline__ double4a ldg256(const double *p)
{
  double4a v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg256(double *p, const double4a &v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kTravStages = 8;
constexpr int kTravComputeWarps = 7;  // + 1 producer warp = 256 threads, 2 blocks/SM at 128 registers
constexpr int kTravThreads = (kTravComputeWarps + 1) * 32;

template <int NCATG>
struct __align__(128) TravStage
{
  OpDev  op;  // 96 bytes
  char   pad[128 - sizeof(OpDev)];
  double M[2][NCATG * 64];  // per child: P[cat][i][j] (NCATG*16) or TP[cat][mask][i] (NCATG*64)
};

__device__ __forceinline__ bool all_one(const double4a &v)
{
  const long long one = 0x3FF0000000000000LL;
  return (((__double_as_longlong(v.x) ^ one) | (__double_as_longlong(v.y) ^ one) | (__double_as_longlong(v.z) ^ one) |
           (__double_as_longlong(v.w) ^ one)) == 0);
}

// operands of one update -> per-child conditional vectors uA, uB, the all-ones flag and the scaler sum
template <int NCATG, int U, int KA, int KB>
__device__ __forceinline__ void trav_operands(const TravStage<NCATG> &stg, int cat, const long long (&off4)[U],
                                              const int (&sidx)[U], const double4a (&prev_o)[U],
                                              const int (&prev_sc)[U], const double4a (&slot_v)[U],
                                              const int (&slot_sc)[U], const uint32_t (&mA)[U],
                                              const uint32_t (&mB)[U], double (&uA)[U][4], double (&uB)[U][4],
                                              bool (&ones)[U], int (&sc)[U])
{
  double4a late_v[U];
  if (KB == kSrcLate)
  {
    const double *c2 = stg.op.c2;
    const int    *s2 = stg.op.s2;
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      late_v[u] = ldg256(c2 + off4[u]);
      sc[u] = s2[sidx[u]];
    }
  }
  else
  {
#pragma unroll
    for (int u = 0; u < U; ++u) sc[u] = 0;
  }
  // ---- child A
  if (KA == kSrcTip)
  {
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double *t = stg.M[0] + (cat * 16 + (int)mA[u]) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) uA[u][i] = t[i];
      ones[u] = (mA[u] == (uint32_t)kTipRowAllOnes);
    }
  }
  else
  {
    double p[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = stg.M[0][cat * 16 + q];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double4a &v = (KA == kSrcFwd) ? prev_o[u] : slot_v[u];
      ones[u] = all_one(v);
      matvec4(p, v, uA[u]);
      sc[u] += (KA == kSrcFwd) ? prev_sc[u] : slot_sc[u];
    }
  }
  // ---- child B
  if (KB == kSrcTip)
  {
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double *t = stg.M[1] + (cat * 16 + (int)mB[u]) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) uB[u][i] = t[i];
      ones[u] = ones[u] && (mB[u] == (uint32_t)kTipRowAllOnes);
    }
  }
  else
  {
    double p[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = stg.M[1][cat * 16 + q];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double4a &v = (KB == kSrcSlot) ? slot_v[u] : late_v[u];
      ones[u] = ones[u] && all_one(v);
      matvec4(p, v, uB[u]);
      if (KB == kSrcSlot) sc[u] += slot_sc[u];
    }
  }
}

template <int NCATG, int UMAX>
__global__ void __launch_bounds__(kTravThreads, 2)
    k_traverse_dna(const OpDev *__restrict__ ops, int n_ops, int npat, int tile_sites, int n_tiles,
                   const double *__restrict__ wght, int apply_scaling)
{
  static_assert(NCATG == 1 || NCATG == 2 || NCATG == 4 || NCATG == 8, "NCATG must divide the warp");
  constexpr int      S = (NCATG >= 8) ? kTravStages / 2 : kTravStages;  // static shared memory <= 48 KB
  constexpr uint32_t PB = NCATG * 16 * sizeof(double);   // P
  constexpr uint32_t TB = NCATG * 64 * sizeof(double);   // tip table
  __shared__ TravStage<NCATG>       st[S];
  __shared__ __align__(8) uint64_t full[S], empty[S];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
  {
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTravComputeWarps);
    }
  }
  __syncthreads();

  const int       rounds = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total_it = (long long)rounds * n_ops;

  if (warp == kTravComputeWarps)
  {  // ---------------- producer warp: descriptor + the two matrices of every update, S updates ahead.
    // The 32 lanes read the matrix pointers / kinds of 32 consecutive updates in one go, so the
    // issuing lane never waits on a dependent global load per update.
    for (long long base = 0; base < total_it; base += 32)
    {
      const long long    my = base + lane;
      unsigned long long m1 = 0, m2 = 0;
      int                kd = 0;
      if (my < total_it)
      {
        const OpDev *o = ops + (my % n_ops);
        m1 = (unsigned long long)o->P1;
        m2 = (unsigned long long)o->P2;
        kd = o->flags;
      }
      const int cnt = (int)min((long long)32, total_it - base);
      for (int j = 0; j < cnt; ++j)
      {
        const unsigned long long a1 = __shfl_sync(0xffffffffu, m1, j), a2 = __shfl_sync(0xffffffffu, m2, j);
        const int                kind = __shfl_sync(0xffffffffu, kd, j);
        if (lane == 0)
        {
          const long long it = base + j;
          const int       s = (int)(it % S);
          const uint32_t  ph = (uint32_t)((it / S) & 1);
          mbar_wait_backoff(&empty[s], ph ^ 1u);
          const uint32_t b1 = ((kind & 3) == kSrcTip) ? TB : PB;
          const uint32_t b2 = ((kind >> 2) == kSrcTip) ? TB : PB;
          mbar_expect_tx(&full[s], (uint32_t)sizeof(OpDev) + b1 + b2);
          tma_bulk_g2s(&st[s].op, ops + (it % n_ops), (uint32_t)sizeof(OpDev), &full[s]);
          tma_bulk_g2s(st[s].M[0], (const void *)a1, b1, &full[s]);
          tma_bulk_g2s(st[s].M[1], (const void *)a2, b2, &full[s]);
        }
        __syncwarp();
      }
    }
    return;
  }

  // ---------------- compute warps
  // lane -> (site, category): the SW = 32/NCATG lanes of a category group read the same P rows
  // (shared-memory broadcast), the NCATG groups of a warp cover the same SW sites.
  constexpr int SW = 32 / NCATG;
  const int     cat = lane / SW;
  const double big = two_to_large(), small = inv_two_to_large();
  long long    it = 0;
  for (int r = 0; r < rounds; ++r)
  {
    const int tile = (int)blockIdx.x + r * (int)gridDim.x;
    const int base_site = tile * tile_sites;
    const int n_sites = min(tile_sites, npat - base_site);
    const int n_items = n_sites * NCATG;
    int       sidx[UMAX];
    long long off4[UMAX];  // element offset of the item's 4 states
    bool      live[UMAX];
#pragma unroll
    for (int u = 0; u < UMAX; ++u)
    {
      const int ls = (u * kTravComputeWarps + warp) * SW + (lane % SW);  // site within the tile
      const int lc = ls < n_sites ? ls : n_sites - 1;
      sidx[u] = base_site + lc;
      off4[u] = ((((long long)(sidx[u] >> 3) * NCATG + cat) << 3) + (sidx[u] & 7)) * 4;  // blocked layout
      live[u] = (ls < n_sites) && (wght[sidx[u]] > DBL_MIN);  // avx.c:399
    }
    (void)n_items;

    double4a prev_o[UMAX];  // result of the previous update (FWD operand of the next one)
    int      prev_sc[UMAX];
    double4a slot_v[UMAX];  // prefetched CLV operand
    int      slot_sc[UMAX];
    uint32_t mA[UMAX], mB[UMAX];  // prefetched tip masks
#pragma unroll
    for (int u = 0; u < UMAX; ++u)
    {
      prev_o[u].x = prev_o[u].y = prev_o[u].z = prev_o[u].w = 0.0;
      slot_v[u] = prev_o[u];
      prev_sc[u] = slot_sc[u] = 0;
      mA[u] = mB[u] = 0u;
    }

    // operand prefetch for the update staged in `sn`
#define PLK_FETCH(sn)                                                                  \
  {                                                                                    \
    const OpDev &on = st[sn].op;                                                       \
    const int    kn = on.flags, ka = kn & 3, kb = kn >> 2;                             \
    nx_kind = kn;                                                                      \
    nx_dst = on.dst;                                                                   \
    nx_dst_scale = on.dst_scale;                                                       \
    if (ka == kSrcSlot || kb == kSrcSlot)                                              \
    {                                                                                  \
      const double *lp = (ka == kSrcSlot) ? on.c1 : on.c2;                             \
      const int    *ls = (ka == kSrcSlot) ? on.s1 : on.s2;                             \
      _Pragma("unroll") for (int u = 0; u < UMAX; ++u)                                 \
      {                                                                                \
        slot_v[u] = ldg256(lp + off4[u]);                                              \
        slot_sc[u] = ls[sidx[u]];                                                      \
      }                                                                                \
    }                                                                                  \
    if (ka == kSrcTip)                                                                 \
    {                                                                                  \
      const uint8_t *tp = on.t1;                                                       \
      _Pragma("unroll") for (int u = 0; u < UMAX; ++u) mA[u] = tp[sidx[u]];            \
    }                                                                                  \
    if (kb == kSrcTip)                                                                 \
    {                                                                                  \
      const uint8_t *tp = on.t2;                                                       \
      _Pragma("unroll") for (int u = 0; u < UMAX; ++u) mB[u] = tp[sidx[u]];            \
    }                                                                                  \
  }

    int     nx_kind = 0;
    double *nx_dst = nullptr;
    int    *nx_dst_scale = nullptr;
    {
      const int      s0 = (int)(it % S);
      const uint32_t ph0 = (uint32_t)((it / S) & 1);
      mbar_wait(&full[s0], ph0);
      PLK_FETCH(s0)
    }

    // one update, fully specialised on its operand kinds: operands -> prefetch of the next update's
    // operands -> products / per-site max / rescaling / store.  Nothing is merged across kinds, so no
    // register shuffling between the specialisations.
#define PLK_UPDATE(KA, KB)                                                                                    \
  {                                                                                                           \
    double uA[UMAX][4], uB[UMAX][4];                                                                          \
    bool   ones[UMAX];                                                                                        \
    int    sc[UMAX];                                                                                          \
    trav_operands<NCATG, UMAX, KA, KB>(stg, cat, off4, sidx, prev_o, prev_sc, slot_v, slot_sc, mA, mB, uA, uB, \
                                       ones, sc);                                                             \
    if (k + 1 < n_ops)                                                                                        \
    {                                                                                                         \
      const int      sn = (int)((it + 1) % S);                                                                \
      const uint32_t phn = (uint32_t)(((it + 1) / S) & 1);                                                    \
      mbar_wait(&full[sn], phn);                                                                              \
      PLK_FETCH(sn)                                                                                           \
    }                                                                                                         \
    _Pragma("unroll") for (int u = 0; u < UMAX; ++u)                                                          \
    {                                                                                                         \
      double4a o;                                                                                             \
      if (ones[u])                                                                                            \
      {                                                                                                       \
        o.x = o.y = o.z = o.w = 1.0;                                                                          \
      }                                                                                                       \
      else                                                                                                    \
      {                                                                                                       \
        o.x = uA[u][0] * uB[u][0];                                                                            \
        o.y = uA[u][1] * uB[u][1];                                                                            \
        o.z = uA[u][2] * uB[u][2];                                                                            \
        o.w = uA[u][3] * uB[u][3];                                                                            \
      }                                                                                                       \
      /* avx.c:498-510: is the largest entry of the site below 2^-256?  All entries are >= 0, so the   */     \
      /* test is an integer compare of the exponent words (NaN counts as large, like the reference).   */     \
      const unsigned hi = (unsigned)max(max(__double2hiint(o.x), __double2hiint(o.y)),                        \
                                        max(__double2hiint(o.z), __double2hiint(o.w)));                       \
      unsigned       hmax = hi;                                                                               \
      _Pragma("unroll") for (int d = SW; d < 32; d <<= 1) hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, d)); \
      int sco = sc[u];                                                                                        \
      if (hmax < 0x2FF00000u && apply_scaling)                                                                \
      {                                                                                                       \
        o.x *= big;                                                                                           \
        o.y *= big;                                                                                           \
        o.z *= big;                                                                                           \
        o.w *= big;                                                                                           \
        sco += kLarge;                                                                                        \
      }                                                                                                       \
      if (live[u])                                                                                            \
      {                                                                                                       \
        stg256(dst + off4[u], o);                                                                             \
        if (cat == 0) dst_scale[sidx[u]] = sco;                                                               \
      }                                                                                                       \
      prev_o[u] = o;                                                                                          \
      prev_sc[u] = sco;                                                                                       \
    }                                                                                                         \
  }

    for (int k = 0; k < n_ops; ++k, ++it)
    {
      const int                s = (int)(it % S);
      const TravStage<NCATG> &stg = st[s];
      const int                kind = nx_kind;
      double *const            dst = nx_dst;
      int *const               dst_scale = nx_dst_scale;
      switch (kind)
      {
      case (kSrcFwd | (kSrcTip << 2)): PLK_UPDATE(kSrcFwd, kSrcTip) break;
      case (kSrcFwd | (kSrcSlot << 2)): PLK_UPDATE(kSrcFwd, kSrcSlot) break;
      case (kSrcTip | (kSrcTip << 2)): PLK_UPDATE(kSrcTip, kSrcTip) break;
      case (kSrcSlot | (kSrcTip << 2)): PLK_UPDATE(kSrcSlot, kSrcTip) break;
      default: PLK_UPDATE(kSrcSlot, kSrcLate) break;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#undef PLK_UPDATE
#undef PLK_FETCH
  }
}

// ------------------------------------------------------------------------------------------------
// K1 fused traversal, 4 states, second generation ("op-major").  Same contract, lane map, TMA/mbarrier
// descriptor ring and arithmetic as k_traverse_dna; what changes is the loop nest.  k_traverse_dna keeps
// U items per thread in registers and walks the whole op list once per register-resident tile, so every
// update pays its fixed cost (barrier wait, descriptor, 16 LDS.128 of P, kind dispatch) per tile pass and a
// pattern count that is not a multiple of the resident capacity wastes a whole pass (measured: the fixed
// cost is ~70 % of a U = 1 pass).  Here a block owns ONE contiguous tile of 32-item chunks for the whole
// launch, and every compute warp loops over its chunks INSIDE each update:
//   * fixed per-update cost paid once per warp, P held in registers across the chunk loop;
//   * the previous update's result (the FWD operand) is forwarded through thread-private shared memory
//     (two conflict-free 16-byte planes + the scaler per lane and chunk) instead of registers, so no value is
//     live across updates and the five operand-kind specialisations share no register state;
//   * global operands of chunk j+1 are loaded while chunk j is computed (L2 latency ~250 cycles);
//   * quantisation is one chunk per warp instead of one pass per block.
// Shared memory per block: kT2Stages ring stages + tile_chunks x 1152 bytes (dynamic).
constexpr int kT2Stages = 4;
constexpr int kT2ChunkBytes = 2 * 512 + 128;
template <int NCATG>
__host__ __device__ inline size_t t2_smem_bytes(int tile_chunks)
{
  return (size_t)kT2Stages * sizeof(TravStage<NCATG>) + (size_t)tile_chunks * kT2ChunkBytes;
}

// Data-path accessors of the op-major kernel: volatile (never dropped, never reordered among themselves) but
// without a "memory" clobber, so that the surrounding scalars stay in registers.  Ordering against the
// C++-level accesses is provided by the mbarrier wait / arrive asm statements, which do clobber memory.
__device__ __forceinline__ void lds128(uint32_t a, double &x, double &y)
{
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void sts128(uint32_t a, double x, double y)
{
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y));
}
__device__ __forceinline__ int lds32(uint32_t a)
{
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ double4a ldg256q(const double *p)
{
  double4a v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256q(double *p, const double4a &v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w));
}
__device__ __forceinline__ int ldg32q(const int *p)
{
  int v;
  asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ldg8q(const uint8_t *p)
{
  uint32_t v;
  asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg32q(int *p, int v) { asm volatile("st.global.s32 [%0], %1;" ::"l"(p), "r"(v)); }

// operands of one chunk.  `fwa` = shared-memory address of this thread's forwarding slot of the chunk,
// `off` = element offset of its 4 states in a blocked CLV (32-bit: a CLV of one shard stays below 2^31
// doubles), `sidx` = site pattern.
#define T2_LOAD(va, vb, scA, scB, rowA, rowB, fwa, off, sidx)    \
  {                                                              \
    if (KA == kSrcFwd)                                           \
    {                                                            \
      lds128((fwa), va.x, va.y);                                 \
      lds128((fwa) + 512, va.z, va.w);                           \
      scA = lds32(fws_of(fwa));                                  \
    }                                                            \
    else if (KA == kSrcSlot)                                     \
    {                                                            \
      va = ldg256q(c1 + (off));                                  \
      scA = ldg32q(s1 + (sidx));                                 \
    }                                                            \
    else                                                         \
      rowA = ldg8q(t1 + (sidx));                                 \
    if (KB == kSrcSlot)                                          \
    {                                                            \
      vb = ldg256q(c2 + (off));                                  \
      scB = ldg32q(s2 + (sidx));                                 \
    }                                                            \
    else                                                         \
      rowB = ldg8q(t2 + (sidx));                                 \
  }

// one chunk: products, all-ones short cut, per-site maximum, rescaling, store + forward
template <int NCATG, int KA, int KB>
__device__ __forceinline__ void t2_chunk(const double4a &va, const double4a &vb, int scA, int scB, uint32_t rowA,
                                         uint32_t rowB, const double (&pA)[16], const double (&pB)[16], uint32_t MAa,
                                         uint32_t MBa, double *dst, int *dst_scale, uint32_t fwa, uint32_t fwsa, int cat,
                                         int off, int sidx, bool live, int apply_scaling)
{
  constexpr int SW = 32 / NCATG;
  // avx.c:575-587: both children all ones (only below fully ambiguous tips) -> the result is exactly 1.0.
  // Cheap filter (tip row / first state); the full compare and the select only run when some lane passes it.
  bool ones = (KA == kSrcTip) ? (rowA == (uint32_t)kTipRowAllOnes) : (__double2hiint(va.x) == 0x3FF00000);
  ones = ones && ((KB == kSrcTip) ? (rowB == (uint32_t)kTipRowAllOnes) : (__double2hiint(vb.x) == 0x3FF00000));
  const bool any_ones = __any_sync(0xffffffffu, ones);
  if (any_ones)
  {
    if (KA != kSrcTip) ones = ones && all_one(va);
    if (KB != kSrcTip) ones = ones && all_one(vb);
  }
  double uA[4], uB[4];
  if (KA == kSrcTip)
  {
    const uint32_t t = MAa + (uint32_t)((cat * 16 + (int)rowA) * 32);
    lds128(t, uA[0], uA[1]);
    lds128(t + 16, uA[2], uA[3]);
  }
  else
    matvec4(pA, va, uA);
  if (KB == kSrcTip)
  {
    const uint32_t t = MBa + (uint32_t)((cat * 16 + (int)rowB) * 32);
    lds128(t, uB[0], uB[1]);
    lds128(t + 16, uB[2], uB[3]);
  }
  else
    matvec4(pB, vb, uB);
  double4a o;
  o.x = uA[0] * uB[0];
  o.y = uA[1] * uB[1];
  o.z = uA[2] * uB[2];
  o.w = uA[3] * uB[3];
  if (any_ones)
  {
    if (ones) o.x = o.y = o.z = o.w = 1.0;
  }
  // avx.c:498-510: is the largest entry of the site below 2^-256?  All entries are >= 0, so the test is an
  // integer compare of the exponent words (NaN counts as large, like the reference).
  unsigned hmax = (unsigned)max(max(__double2hiint(o.x), __double2hiint(o.y)), max(__double2hiint(o.z), __double2hiint(o.w)));
#pragma unroll
  for (int d = SW; d < 32; d <<= 1) hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, d));
  int        sco = scA + scB;
  const bool resc = (hmax < 0x2FF00000u) && apply_scaling;
  if (__any_sync(0xffffffffu, resc))
  {
    if (resc)
    {
      const double big = two_to_large();
      o.x *= big;
      o.y *= big;
      o.z *= big;
      o.w *= big;
      sco += kLarge;
    }
  }
  if (live)
  {
    stg256q(dst + off, o);
    if (cat == 0) stg32q(dst_scale + sidx, sco);
  }
  sts128(fwa, o.x, o.y);
  sts128(fwa + 512, o.z, o.w);
  sts32(fwsa, sco);
}

template <int NCATG, int W, int KA, int KB>
__device__ __forceinline__ void t2_run_op(const TravStage<NCATG> &stg, uint32_t fwa0, uint32_t fwsa0, int nch,
                                          unsigned live_mask, int off0, int sidx0, int cat, int apply_scaling)
{
  constexpr int  SW = 32 / NCATG;
  constexpr int  kFwdStride = W * kT2ChunkBytes;
  constexpr int  kOffStride = W * 128;  // one chunk = 32 items x 4 doubles
  constexpr int  kSiteStride = W * SW;
  const double  *c1 = stg.op.c1, *c2 = stg.op.c2;
  const int     *s1 = stg.op.s1, *s2 = stg.op.s2;
  const uint8_t *t1 = stg.op.t1, *t2 = stg.op.t2;
  double        *dst = stg.op.dst;
  int           *dst_scale = stg.op.dst_scale;
  const uint32_t MAa = smem_u32(stg.M[0]), MBa = smem_u32(stg.M[1]);
  const uint32_t fws_delta = fwsa0 - fwa0;
  auto           fws_of = [fws_delta](uint32_t fwa) { return fwa + fws_delta; };
  double         pA[16], pB[16];
  if (KA != kSrcTip)
  {
#pragma unroll
    for (int q = 0; q < 8; ++q) lds128(MAa + (uint32_t)(cat * 128 + q * 16), pA[2 * q], pA[2 * q + 1]);
  }
  if (KB != kSrcTip)
  {
#pragma unroll
    for (int q = 0; q < 8; ++q) lds128(MBa + (uint32_t)(cat * 128 + q * 16), pB[2 * q], pB[2 * q + 1]);
  }
  // two operand buffers used alternately: the loads of chunk j+1 are in flight while chunk j is computed
  double4a va0, vb0, va1, vb1;
  int      scA0 = 0, scB0 = 0, scA1 = 0, scB1 = 0;
  uint32_t rowA0 = 0, rowB0 = 0, rowA1 = 0, rowB1 = 0;
  va0.x = va0.y = va0.z = va0.w = 0.0;
  vb0 = va1 = vb1 = va0;
  uint32_t fwa = fwa0;
  int      off = off0, sidx = sidx0;
  if (nch > 0) T2_LOAD(va0, vb0, scA0, scB0, rowA0, rowB0, fwa, off, sidx)
  for (int j = 0; j < nch; j += 2)
  {
    const bool has1 = (j + 1 < nch);
    if (has1) T2_LOAD(va1, vb1, scA1, scB1, rowA1, rowB1, fwa + kFwdStride, off + kOffStride, sidx + kSiteStride)
    t2_chunk<NCATG, KA, KB>(va0, vb0, scA0, scB0, rowA0, rowB0, pA, pB, MAa, MBa, dst, dst_scale, fwa, fws_of(fwa), cat, off,
                            sidx, (live_mask >> j) & 1u, apply_scaling);
    if (has1)
    {
      if (j + 2 < nch)
        T2_LOAD(va0, vb0, scA0, scB0, rowA0, rowB0, fwa + 2 * kFwdStride, off + 2 * kOffStride, sidx + 2 * kSiteStride)
      t2_chunk<NCATG, KA, KB>(va1, vb1, scA1, scB1, rowA1, rowB1, pA, pB, MAa, MBa, dst, dst_scale, fwa + kFwdStride,
                              fws_of(fwa + kFwdStride), cat, off + kOffStride, sidx + kSiteStride,
                              (live_mask >> (j + 1)) & 1u, apply_scaling);
    }
    fwa += 2 * kFwdStride;
    off += 2 * kOffStride;
    sidx += 2 * kSiteStride;
  }
}
#undef T2_LOAD

template <int NCATG, int W, int MINB>
__global__ void __launch_bounds__((W + 1) * 32, MINB)
    k_traverse_dna2(const OpDev *__restrict__ ops, int n_ops, int total_chunks, int tile_chunks, int n_tiles,
                    const double *__restrict__ wght, int apply_scaling)
{
  static_assert(NCATG == 1 || NCATG == 2 || NCATG == 4 || NCATG == 8, "NCATG must divide the warp");
  static_assert(NCATG != 8 || (W % 2) == 0, "8 categories: a chunk is half an 8-site block, W must be even");
  constexpr int      S = kT2Stages;
  constexpr int      SW = 32 / NCATG;
  constexpr uint32_t PB = NCATG * 16 * sizeof(double);
  constexpr uint32_t TB = NCATG * 64 * sizeof(double);
  extern __shared__ __align__(128) unsigned char t2_smem[];
  __shared__ __align__(8) uint64_t               full[S], empty[S];
  TravStage<NCATG> *st = reinterpret_cast<TravStage<NCATG> *>(t2_smem);
  unsigned char    *fwd = t2_smem + (size_t)S * sizeof(TravStage<NCATG>);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], W);
    }
  __syncthreads();
  const int       rounds = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total_it = (long long)rounds * n_ops;

  if (warp == W)
  {  // ---------------- producer warp (same protocol as k_traverse_dna)
    for (long long base = 0; base < total_it; base += 32)
    {
      const long long    my = base + lane;
      unsigned long long m1 = 0, m2 = 0;
      int                kd = 0;
      if (my < total_it)
      {
        const OpDev *o = ops + (my % n_ops);
        m1 = (unsigned long long)o->P1;
        m2 = (unsigned long long)o->P2;
        kd = o->flags;
      }
      const int cnt = (int)min((long long)32, total_it - base);
      for (int j = 0; j < cnt; ++j)
      {
        const unsigned long long a1 = __shfl_sync(0xffffffffu, m1, j), a2 = __shfl_sync(0xffffffffu, m2, j);
        const int                kind = __shfl_sync(0xffffffffu, kd, j);
        if (lane == 0)
        {
          const long long it = base + j;
          const int       s = (int)(it % S);
          const uint32_t  ph = (uint32_t)((it / S) & 1);
          mbar_wait_backoff(&empty[s], ph ^ 1u);
          const uint32_t b1 = ((kind & 3) == kSrcTip) ? TB : PB;
          const uint32_t b2 = ((kind >> 2) == kSrcTip) ? TB : PB;
          mbar_expect_tx(&full[s], (uint32_t)sizeof(OpDev) + b1 + b2);
          tma_bulk_g2s(&st[s].op, ops + (it % n_ops), (uint32_t)sizeof(OpDev), &full[s]);
          tma_bulk_g2s(st[s].M[0], (const void *)a1, b1, &full[s]);
          tma_bulk_g2s(st[s].M[1], (const void *)a2, b2, &full[s]);
        }
        __syncwarp();
      }
    }
    return;
  }

  // ---------------- compute warps: lane -> (site within the chunk, category)
  const int cat = lane / SW, ls = lane % SW;
  long long it = 0;
  for (int r = 0; r < rounds; ++r)
  {
    const int tile = (int)blockIdx.x + r * (int)gridDim.x;
    const int chunk0 = tile * tile_chunks;
    const int n_chunks = min(tile_chunks, total_chunks - chunk0);
    const int nch = (n_chunks > warp) ? (n_chunks - 1 - warp) / W + 1 : 0;  // chunks warp, warp + W, ... of the tile
    const int sidx0 = (chunk0 + warp) * SW + ls;
    const int      off0 = ((((sidx0 >> 3) * NCATG + cat) << 3) + (sidx0 & 7)) * 4;  // blocked layout, < 2^31 doubles
    const uint32_t fwa0 = smem_u32(fwd + (size_t)warp * kT2ChunkBytes + lane * 16);  // this thread's slot, plane 0
    const uint32_t fwsa0 = smem_u32(fwd + (size_t)warp * kT2ChunkBytes + 1024 + lane * 4);
    unsigned        live_mask = 0u;
    for (int j = 0; j < nch; ++j)
      if (wght[sidx0 + j * (W * SW)] > DBL_MIN) live_mask |= 1u << j;  // avx.c:399 (arrays are padded with zero weights)

    for (int k = 0; k < n_ops; ++k, ++it)
    {
      const int                s = (int)(it % S);
      const TravStage<NCATG> &stg = st[s];
      mbar_wait(&full[s], (uint32_t)((it / S) & 1));
      const int kind = stg.op.flags, ka = kind & 3, kb = kind >> 2;
      if (ka == kSrcFwd)
      {
        if (kb == kSrcTip)
          t2_run_op<NCATG, W, kSrcFwd, kSrcTip>(stg, fwa0, fwsa0, nch, live_mask, off0, sidx0, cat, apply_scaling);
        else
          t2_run_op<NCATG, W, kSrcFwd, kSrcSlot>(stg, fwa0, fwsa0, nch, live_mask, off0, sidx0, cat, apply_scaling);
      }
      else if (ka == kSrcTip)
        t2_run_op<NCATG, W, kSrcTip, kSrcTip>(stg, fwa0, fwsa0, nch, live_mask, off0, sidx0, cat, apply_scaling);
      else if (kb == kSrcTip)
        t2_run_op<NCATG, W, kSrcSlot, kSrcTip>(stg, fwa0, fwsa0, nch, live_mask, off0, sidx0, cat, apply_scaling);
      else
        t2_run_op<NCATG, W, kSrcSlot, kSrcSlot>(stg, fwa0, fwsa0, nch, live_mask, off0, sidx0, cat, apply_scaling);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

// tip codes -> rows of the 20-state tip tables: state for one-hot codes, 20 for all-ones, 255 = general mask
__global__ void k_codes_to_rows20(const uint8_t *__restrict__ codes, uint8_t *__restrict__ rows, size_t n,
                                  const uint32_t *__restrict__ tipmask)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const uint32_t m = tipmask[codes[i]] & 0xFFFFFu;
    rows[i] = (__popc(m) == 1) ? (uint8_t)(__ffs(m) - 1) : (m == 0xFFFFFu ? (uint8_t)20 : (uint8_t)255);
  }
}

// 4-bit packed tip codes (two patterns per byte, low nibble first) -> 1-byte codes and, for the fused kernels,
// the tip-table rows (mode 1: 4 states, mode 2: 20 states); 8 packed bytes per thread
__global__ void k_unpack_codes4(const uint8_t *__restrict__ packed, uint8_t *__restrict__ codes, uint8_t *__restrict__ rows,
                                size_t n_packed, const uint32_t *__restrict__ tipmask, int mode)
{
  __shared__ uint8_t row_of[16];
  if (threadIdx.x < 16)
  {
    const uint32_t m = tipmask[threadIdx.x];
    uint8_t        r = 0;
    if (mode == 1)
      r = (uint8_t)tip_row4((int)m);
    else if (mode == 2)
    {
      const uint32_t mm = m & 0xFFFFFu;
      r = (__popc(mm) == 1) ? (uint8_t)(__ffs(mm) - 1) : (mm == 0xFFFFFu ? (uint8_t)20 : (uint8_t)255);
    }
    row_of[threadIdx.x] = r;
  }
  __syncthreads();
  const size_t n8 = n_packed / 8;  // n_packed is a multiple of 64 (rows are padded to 128 codes)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x)
  {
    const uint2 v = reinterpret_cast<const uint2 *>(packed)[i];
    uint32_t    c[4], r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
    {
      const uint32_t h = (q < 2) ? (v.x >> (16 * q)) & 0xFFFFu : (v.y >> (16 * (q - 2))) & 0xFFFFu;  // 2 packed bytes
      const uint32_t n0 = h & 15u, n1 = (h >> 4) & 15u, n2 = (h >> 8) & 15u, n3 = (h >> 12) & 15u;
      c[q] = n0 | (n1 << 8) | (n2 << 16) | (n3 << 24);
      r[q] = (uint32_t)row_of[n0] | ((uint32_t)row_of[n1] << 8) | ((uint32_t)row_of[n2] << 16) | ((uint32_t)row_of[n3] << 24);
    }
    reinterpret_cast<uint4 *>(codes)[i] = make_uint4(c[0], c[1], c[2], c[3]);
    if (rows) reinterpret_cast<uint4 *>(rows)[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

// tip codes -> rows of the 4-state tip tables (run once per tip upload; t1/t2 of fused ops point here)
__global__ void k_codes_to_rows(const uint8_t *__restrict__ codes, uint8_t *__restrict__ rows, size_t n,
                                const uint32_t *__restrict__ tipmask)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    rows[i] = (uint8_t)tip_row4((int)tipmask[codes[i]]);
}

// ------------------------------------------------------------------------------------------------
// K1 fused traversal, 4 states, tensor-pipe variant.  With 4 states the matrix-vector product of one
// child for 8 sites of one category is exactly ONE DMMA.8x8x4:  D[site][i] = sum_j x[site][j] P[i][j]
// (A = 8 sites x 4 states of the child CLV, B[k=j][n=i] = P[i][j], columns 4..7 zero).  On B200 the
// DMMA accumulates as an ascending-k FMA chain from C (tools/probes/dmma_order.cu), i.e. with C = 0 it
// rounds exactly like the reference's AVX_Matrix_Vect_Prod, so this variant is bit-identical to
// k_traverse_dna.  What it buys: P costs one register pair per (child, category) instead of 32
// registers, a CLV word is ONE double per thread, so a warp carries U*8 sites x all categories
// (4x the sites of the FMA kernel at the same register budget) and issues ~half the instructions.
// Blocked CLV layout: every fragment load (LDG.64 x 32 lanes) / store (STG.128 x 16 lanes) is one
// contiguous 256-byte block.  The previous update's result is forwarded in registers: its C fragment
// (lane t<2 holds states 2t,2t+1) is turned into the next A fragment (lane t holds state t) with two
// quad shuffles per category.
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  // not volatile: independent accumulator chains may be interleaved by the scheduler
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double ldg64(const double *p)
{
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg128(double *p, double x, double y)
{
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void ldg128(const double *p, double &x, double &y)
{
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(p) : "memory");
}

template <int NCATG, int U>
__global__ void __launch_bounds__(kTravThreads, 2)
    k_traverse_dna_mma(const OpDev *__restrict__ ops, int n_ops, int npat, int tile_sites, int n_tiles,
                       const double *__restrict__ wght, int apply_scaling)
{
  constexpr int      S = (NCATG >= 8) ? kTravStages / 2 : kTravStages;
  constexpr uint32_t PB = NCATG * 16 * sizeof(double);
  constexpr uint32_t TB = NCATG * 64 * sizeof(double);
  __shared__ TravStage<NCATG>       st[S];
  __shared__ __align__(8) uint64_t full[S], empty[S];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTravComputeWarps);
    }
  __syncthreads();
  const int       rounds = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total_it = (long long)rounds * n_ops;

  if (warp == kTravComputeWarps)
  {  // ---------------- producer warp (same protocol as k_traverse_dna)
    for (long long base = 0; base < total_it; base += 32)
    {
      const long long    my = base + lane;
      unsigned long long m1 = 0, m2 = 0;
      int                kd = 0;
      if (my < total_it)
      {
        const OpDev *o = ops + (my % n_ops);
        m1 = (unsigned long long)o->P1;
        m2 = (unsigned long long)o->P2;
        kd = o->flags;
      }
      const int cnt = (int)min((long long)32, total_it - base);
      for (int j = 0; j < cnt; ++j)
      {
        const unsigned long long a1 = __shfl_sync(0xffffffffu, m1, j), a2 = __shfl_sync(0xffffffffu, m2, j);
        const int                kind = __shfl_sync(0xffffffffu, kd, j);
        if (lane == 0)
        {
          const long long it = base + j;
          const int       s = (int)(it % S);
          const uint32_t  ph = (uint32_t)((it / S) & 1);
          mbar_wait_backoff(&empty[s], ph ^ 1u);
          const uint32_t b1 = ((kind & 3) == kSrcTip) ? TB : PB;
          const uint32_t b2 = ((kind >> 2) == kSrcTip) ? TB : PB;
          mbar_expect_tx(&full[s], (uint32_t)sizeof(OpDev) + b1 + b2);
          tma_bulk_g2s(&st[s].op, ops + (it % n_ops), (uint32_t)sizeof(OpDev), &full[s]);
          tma_bulk_g2s(st[s].M[0], (const void *)a1, b1, &full[s]);
          tma_bulk_g2s(st[s].M[1], (const void *)a2, b2, &full[s]);
        }
        __syncwarp();
      }
    }
    return;
  }

  // ---------------- compute warps: lane = (g = site within the 8-site block, t = state / column pair)
  const int      g = lane >> 2, t = lane & 3;
  const int      fwd_src = (lane & ~3) | (t >> 1);  // lane holding state t in the C-fragment layout
  const unsigned qshift = lane & ~3;
  const double   big = two_to_large();
  long long      it = 0;
  for (int r = 0; r < rounds; ++r)
  {
    const int tile = (int)blockIdx.x + r * (int)gridDim.x;
    const int base_site = tile * tile_sites;
    const int n_sites = min(tile_sites, npat - base_site);
    int       sidx[U];
    long long boff[U];  // element offset of (8-site block, category 0)
    bool      live[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int ls = (u * kTravComputeWarps + warp) * 8 + g;
      const int lc = ls < n_sites ? ls : n_sites - 1;
      sidx[u] = base_site + lc;
      boff[u] = (long long)(sidx[u] >> 3) * NCATG * 32 + (sidx[u] & 7) * 4;
      live[u] = (ls < n_sites) && (wght[sidx[u]] > DBL_MIN);
    }
    double   prevA[U][NCATG], slotA[U][NCATG];
    int      prev_sc[U], slot_sc[U];
    uint32_t rA[U], rB[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
#pragma unroll
      for (int c = 0; c < NCATG; ++c) prevA[u][c] = slotA[u][c] = 0.0;
      prev_sc[u] = slot_sc[u] = 0;
      rA[u] = rB[u] = 0u;
    }

    // operand prefetch of the update staged in `sn`, for m-group u
#define MMA_FETCH(sn, u)                                                                  \
  {                                                                                       \
    const OpDev &on = st[sn].op;                                                          \
    const int    kn = on.flags, ka = kn & 3, kb = kn >> 2;                                \
    if (ka == kSrcSlot || kb == kSrcSlot)                                                 \
    {                                                                                     \
      const double *lp = ((ka == kSrcSlot) ? on.c1 : on.c2) + boff[u] + t;                \
      const int    *ls = (ka == kSrcSlot) ? on.s1 : on.s2;                                \
      _Pragma("unroll") for (int c = 0; c < NCATG; ++c) slotA[u][c] = ldg64(lp + c * 32); \
      slot_sc[u] = ls[sidx[u]];                                                           \
    }                                                                                     \
    if (ka == kSrcTip) rA[u] = on.t1[sidx[u]];                                            \
    if (kb == kSrcTip) rB[u] = on.t2[sidx[u]];                                            \
  }

    {
      const int      s0 = (int)(it % S);
      const uint32_t ph0 = (uint32_t)((it / S) & 1);
      mbar_wait(&full[s0], ph0);
#pragma unroll
      for (int u = 0; u < U; ++u) MMA_FETCH(s0, u)
    }

    for (int k = 0; k < n_ops; ++k, ++it)
    {
      const int                s = (int)(it % S);
      const TravStage<NCATG> &stg = st[s];
      const int                kind = stg.op.flags, ka = kind & 3, kb = kind >> 2;
      double *const            dst = stg.op.dst;
      int *const               dst_scale = stg.op.dst_scale;
      const bool               has_next = (k + 1 < n_ops);
      const int                sn = (int)((it + 1) % S);
      if (has_next) mbar_wait(&full[sn], (uint32_t)(((it + 1) / S) & 1));

      // B fragments: B[k = t][n = g] = P[c][i = g][j = t], columns 4..7 zero
      double bA[NCATG], bB[NCATG];
#pragma unroll
      for (int c = 0; c < NCATG; ++c)
      {
        bA[c] = (ka != kSrcTip && g < 4) ? stg.M[0][c * 16 + g * 4 + t] : 0.0;
        bB[c] = (kb != kSrcTip && g < 4) ? stg.M[1][c * 16 + g * 4 + t] : 0.0;
      }

#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        double   o0[NCATG], o1[NCATG];
        unsigned notone = 0u;  // bit c: some child value of (site, c) held by this lane differs from 1.0
        int      sc = 0;
        // ---- child A
        if (ka == kSrcTip)
        {
          if (rA[u] != (uint32_t)kTipRowAllOnes) notone = (1u << NCATG) - 1u;
        }
        else
          sc += (ka == kSrcFwd) ? prev_sc[u] : slot_sc[u];
        double lateB[NCATG];
        if (kb == kSrcLate)
        {
          const double *lp = stg.op.c2 + boff[u] + t;
#pragma unroll
          for (int c = 0; c < NCATG; ++c) lateB[c] = ldg64(lp + c * 32);
          sc += stg.op.s2[sidx[u]];
        }
        else if (kb == kSrcSlot)
          sc += slot_sc[u];
        else if (rB[u] != (uint32_t)kTipRowAllOnes)
          notone = (1u << NCATG) - 1u;

#pragma unroll
        for (int c = 0; c < NCATG; ++c)
        {
          double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
          if (ka == kSrcTip)
          {
            if (t < 2)
            {
              const double *tp = stg.M[0] + (c * 16 + (int)rA[u]) * 4 + 2 * t;
              a0 = tp[0];
              a1 = tp[1];
            }
          }
          else
          {
            const double x = (ka == kSrcFwd) ? prevA[u][c] : slotA[u][c];
            notone |= (((unsigned)__double2hiint(x) ^ 0x3FF00000u) | (unsigned)__double2loint(x)) ? (1u << c) : 0u;
            dmma884(a0, a1, x, bA[c]);
          }
          if (kb == kSrcTip)
          {
            if (t < 2)
            {
              const double *tp = stg.M[1] + (c * 16 + (int)rB[u]) * 4 + 2 * t;
              b0 = tp[0];
              b1 = tp[1];
            }
          }
          else
          {
            const double x = (kb == kSrcSlot) ? slotA[u][c] : lateB[c];
            notone |= (((unsigned)__double2hiint(x) ^ 0x3FF00000u) | (unsigned)__double2loint(x)) ? (1u << c) : 0u;
            dmma884(b0, b1, x, bB[c]);
          }
          o0[c] = a0 * b0;
          o1[c] = a1 * b1;
        }
        // the operands of this m-group are consumed: prefetch those of the next update
        if (has_next) MMA_FETCH(sn, u)

        // all-ones short cut (avx.c:575-587): quad-wide OR of the "differs from 1.0" bits
        notone |= __shfl_xor_sync(0xffffffffu, notone, 1);
        notone |= __shfl_xor_sync(0xffffffffu, notone, 2);
        if (notone != (1u << NCATG) - 1u)
        {
#pragma unroll
          for (int c = 0; c < NCATG; ++c)
            if (!((notone >> c) & 1u)) o0[c] = o1[c] = 1.0;
        }
        // avx.c:498-510: per-site maximum over all categories and states, as an exponent compare
        int hmax = 0;
#pragma unroll
        for (int c = 0; c < NCATG; ++c) hmax = max(hmax, max(__double2hiint(o0[c]), __double2hiint(o1[c])));
        const unsigned bal = __ballot_sync(0xffffffffu, (t < 2) && ((unsigned)hmax >= 0x2FF00000u));
        if ((((bal >> qshift) & 0xFu) == 0u) && apply_scaling)
        {
#pragma unroll
          for (int c = 0; c < NCATG; ++c)
          {
            o0[c] *= big;
            o1[c] *= big;
          }
          sc += kLarge;
        }
        if (live[u])
        {
          if (t < 2)
          {
            double *op_ = dst + boff[u] + 2 * t;
#pragma unroll
            for (int c = 0; c < NCATG; ++c) stg128(op_ + c * 32, o0[c], o1[c]);
          }
          if (t == 0) dst_scale[sidx[u]] = sc;
        }
        // forward: C-fragment layout -> A-fragment layout of the next update
#pragma unroll
        for (int c = 0; c < NCATG; ++c)
        {
          const double x0 = __shfl_sync(0xffffffffu, o0[c], fwd_src);
          const double x1 = __shfl_sync(0xffffffffu, o1[c], fwd_src);
          prevA[u][c] = (t & 1) ? x1 : x0;
        }
        prev_sc[u] = sc;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#undef MMA_FETCH
  }
}

// ------------------------------------------------------------------------------------------------
// K1 fused traversal, 4 states, third generation: op-major like k_traverse_dna2, arithmetic on the FP64
// tensor pipe like k_traverse_dna_mma.  ncu of k_traverse_dna2 (profiles/ncu_r2_dna2.md): 166 registers
// (two 4x4 P matrices = 64 of them) allow 12 warps per SM, and with one chunk of look-ahead per warp the
// kernel waits on L2 (long-scoreboard stalls 4.1 per issued instruction, issue slots 30 % busy).  With one
// DMMA.8x8x4 per (8 sites, category, child) a P matrix costs ONE register pair per thread and a CLV word is
// one double per thread, so the whole update state of a chunk is ~3*NCATG doubles: 3x the warps per SM, each
// with two chunks in flight, and about half the instructions.
//   lane = (g = lane>>2: site within the 8-site block, t = lane&3: state);  chunk = one 8-site block x NCATG
//   A fragment: child CLV word (site g, state t), LDG.64 - 32 lanes read one contiguous 256-byte block
//   B fragment: B[k=t][n=g] = P[c][g][t] for g < 4, else 0
//   C fragment: lanes t < 2 hold states 2t, 2t+1 of site g (columns 4..7 are padding)
// B200's DMMA accumulates as an ascending-k FMA chain from C = 0 (tools/probes/dmma_order.cu), i.e. it rounds
// exactly like the reference's AVX_Matrix_Vect_Prod: bit-identical to the FMA kernels.
// The previous update's result is forwarded through shared memory in C-fragment order (STS.128 by the
// t < 2 lanes) and read back in A-fragment order (LDS.64, conflict-free): the layout change that cost the
// register-forwarding variant two shuffles per category is free here.
// Tip operands: the 1-byte tip-table rows of the block's tile (8 bytes per chunk) are part of what the
// producer warp stages per update with a TMA bulk copy, several updates ahead.  (First version: LDG.U8 in the
// compute warps one chunk ahead -- those loads miss L2, the 1 GB write stream evicts the 8 MB of tip rows
// between evaluations, and their DRAM latency was the largest stall of the kernel, profiles/ncu_r2_dna3.md.)
constexpr int kT3Stages = 4;
constexpr int kT3MaxTileChunks = 256;  // 2 KB of tip rows per operand and stage
template <int NCATG>
struct __align__(128) T3Stage
{
  OpDev   op;  // 96 bytes
  char    pad[128 - sizeof(OpDev)];
  double  M[2][NCATG * 64];               // per child: P[cat][i][j] (NCATG*16) or TP[cat][mask][i] (NCATG*64)
  uint8_t rows[2][kT3MaxTileChunks * 8];  // per tip child: tip-table row of every site of the tile
};
template <int NCATG>
__host__ __device__ constexpr int t3_chunk_bytes()
{
  return NCATG * 256 + 128;
}
template <int NCATG>
__host__ __device__ inline size_t t3_smem_bytes(int tile_chunks)
{
  return (size_t)kT3Stages * sizeof(T3Stage<NCATG>) + (size_t)tile_chunks * t3_chunk_bytes<NCATG>();
}
__device__ __forceinline__ uint32_t lds8(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds64(uint32_t a)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double ldg64q(const double *p)
{
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128q(double *p, double x, double y)
{
  asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(x), "d"(y));
}

// operands of one chunk: fwa = this lane's 8-byte slot in the chunk's forwarding block (category 0),
// fws = its scaler slot, goff = element offset of (8-site block, category 0) + lane, sidx = site pattern
#define T3_LOAD(xA, xB, scA, scB, rowA, rowB, fwa, fws, goff, sidx)                  \
  {                                                                                  \
    if (KA == kSrcFwd)                                                               \
    {                                                                                \
      _Pragma("unroll") for (int c = 0; c < NCATG; ++c) xA[c] = lds64((fwa) + c * 256); \
      scA = lds32(fws);                                                              \
    }                                                                                \
    else if (KA == kSrcSlot)                                                         \
    {                                                                                \
      _Pragma("unroll") for (int c = 0; c < NCATG; ++c) xA[c] = ldg64q(c1 + (goff) + c * 32); \
      scA = ldg32q(s1 + (sidx));                                                     \
    }                                                                                \
    else                                                                             \
      rowA = lds8(rwA + (sidx));                                                     \
    if (KB == kSrcSlot)                                                              \
    {                                                                                \
      _Pragma("unroll") for (int c = 0; c < NCATG; ++c) xB[c] = ldg64q(c2 + (goff) + c * 32); \
      scB = ldg32q(s2 + (sidx));                                                     \
    }                                                                                \
    else                                                                             \
      rowB = lds8(rwB + (sidx));                                                     \
  }

template <int NCATG, int KA, int KB>
__device__ __forceinline__ void t3_chunk(const double (&xA)[NCATG], const double (&xB)[NCATG], int scA, int scB,
                                         uint32_t rowA, uint32_t rowB, const double (&bA)[NCATG],
                                         const double (&bB)[NCATG], uint32_t MAa, uint32_t MBa, double *dst,
                                         int *dst_scale, uint32_t fwa, uint32_t fws, int lane, int goff, int sidx,
                                         bool live, int apply_scaling)
{
  const int t = lane & 3;
  // avx.c:575-587: both children all ones (only below fully ambiguous tips) -> the result is exactly 1.0.
  // Conservative filter on child A (tip row / exponent word of any category); exact test in the slow path.
  bool maybe = false;
  if (KA == kSrcTip)
    maybe = (rowA == (uint32_t)kTipRowAllOnes);
  else
  {
#pragma unroll
    for (int c = 0; c < NCATG; ++c) maybe = maybe || (__double2hiint(xA[c]) == 0x3FF00000);
  }
  const bool any_ones = __any_sync(0xffffffffu, maybe);
  double     o0[NCATG], o1[NCATG];
#pragma unroll
  for (int c = 0; c < NCATG; ++c)
  {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    if (KA == kSrcTip)
    {
      if (t < 2) lds128(MAa + (uint32_t)(((c * 16 + (int)rowA) * 4 + 2 * t) * 8), a0, a1);
    }
    else
      dmma884(a0, a1, xA[c], bA[c]);
    if (KB == kSrcTip)
    {
      if (t < 2) lds128(MBa + (uint32_t)(((c * 16 + (int)rowB) * 4 + 2 * t) * 8), b0, b1);
    }
    else
      dmma884(b0, b1, xB[c], bB[c]);
    o0[c] = a0 * b0;
    o1[c] = a1 * b1;
  }
  if (any_ones)
  {
#pragma unroll
    for (int c = 0; c < NCATG; ++c)
    {
      const bool     pa = (KA == kSrcTip) ? (rowA == (uint32_t)kTipRowAllOnes) : (xA[c] == 1.0);
      const bool     pb = (KB == kSrcTip) ? (rowB == (uint32_t)kTipRowAllOnes) : (xB[c] == 1.0);
      const unsigned bal = __ballot_sync(0xffffffffu, pa && pb);
      if (((bal >> (lane & ~3)) & 0xFu) == 0xFu) o0[c] = o1[c] = 1.0;
    }
  }
  // avx.c:498-510: per-site maximum over all categories and states as an exponent-word compare (all entries
  // are >= 0; NaN counts as large, like the reference); the t >= 2 lanes hold padding zeros
  int hmax = 0;
#pragma unroll
  for (int c = 0; c < NCATG; ++c) hmax = max(hmax, max(__double2hiint(o0[c]), __double2hiint(o1[c])));
  hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, 1));
  int        sco = scA + scB;
  const bool resc = (t < 2) && ((unsigned)hmax < 0x2FF00000u) && apply_scaling;
  if (__any_sync(0xffffffffu, resc))
  {
    if (resc)
    {
      const double big = two_to_large();
#pragma unroll
      for (int c = 0; c < NCATG; ++c)
      {
        o0[c] *= big;
        o1[c] *= big;
      }
      sco += kLarge;
    }
  }
  if (t < 2)
  {
    if (live)
    {
#pragma unroll
      for (int c = 0; c < NCATG; ++c) stg128q(dst + goff + t + c * 32, o0[c], o1[c]);
      if (t == 0) stg32q(dst_scale + sidx, sco);
    }
#pragma unroll
    for (int c = 0; c < NCATG; ++c) sts128(fwa + t * 8 + c * 256, o0[c], o1[c]);
    sts32(fws, sco);  // only the t == 0 copy is ever consumed (scalers are per site)
  }
}

template <int NCATG, int W, int KA, int KB>
__device__ __forceinline__ void t3_run_op(const T3Stage<NCATG> &stg, uint32_t fwa0, uint32_t fws0, int nch,
                                          unsigned live_mask, int goff0, int sidx0, int tile_site0, int lane,
                                          int apply_scaling)
{
  constexpr int  kFwdStride = W * t3_chunk_bytes<NCATG>();
  constexpr int  kOffStride = W * NCATG * 32;  // one chunk = 8 sites x NCATG x 4 doubles
  constexpr int  kSiteStride = W * 8;
  const double  *c1 = stg.op.c1, *c2 = stg.op.c2;
  const int     *s1 = stg.op.s1, *s2 = stg.op.s2;
  // staged tip rows, biased so that they are indexed by the global site pattern like the CLV operands
  const uint32_t rwA = smem_u32(stg.rows[0]) - (uint32_t)tile_site0, rwB = smem_u32(stg.rows[1]) - (uint32_t)tile_site0;
  double        *dst = stg.op.dst;
  int           *dst_scale = stg.op.dst_scale;
  const uint32_t MAa = smem_u32(stg.M[0]), MBa = smem_u32(stg.M[1]);
  const int      g = lane >> 2, t = lane & 3;
  double         bA[NCATG], bB[NCATG];
#pragma unroll
  for (int c = 0; c < NCATG; ++c)
  {
    bA[c] = (KA != kSrcTip && g < 4) ? lds64(MAa + (uint32_t)((c * 16 + g * 4 + t) * 8)) : 0.0;
    bB[c] = (KB != kSrcTip && g < 4) ? lds64(MBa + (uint32_t)((c * 16 + g * 4 + t) * 8)) : 0.0;
  }
  double   xA0[NCATG], xB0[NCATG], xA1[NCATG], xB1[NCATG];
  int      scA0 = 0, scB0 = 0, scA1 = 0, scB1 = 0;
  uint32_t rowA0 = 0, rowB0 = 0, rowA1 = 0, rowB1 = 0;
#pragma unroll
  for (int c = 0; c < NCATG; ++c) xA0[c] = xB0[c] = xA1[c] = xB1[c] = 0.0;
  uint32_t fwa = fwa0, fws = fws0;
  int      goff = goff0, sidx = sidx0;
  if (nch > 0) T3_LOAD(xA0, xB0, scA0, scB0, rowA0, rowB0, fwa, fws, goff, sidx)
  for (int j = 0; j < nch; j += 2)
  {
    const bool has1 = (j + 1 < nch);
    if (has1)
      T3_LOAD(xA1, xB1, scA1, scB1, rowA1, rowB1, fwa + kFwdStride, fws + kFwdStride, goff + kOffStride, sidx + kSiteStride)
    t3_chunk<NCATG, KA, KB>(xA0, xB0, scA0, scB0, rowA0, rowB0, bA, bB, MAa, MBa, dst, dst_scale, fwa, fws, lane, goff, sidx,
                            (live_mask >> j) & 1u, apply_scaling);
    if (has1)
    {
      if (j + 2 < nch)
        T3_LOAD(xA0, xB0, scA0, scB0, rowA0, rowB0, fwa + 2 * kFwdStride, fws + 2 * kFwdStride, goff + 2 * kOffStride,
                sidx + 2 * kSiteStride)
      t3_chunk<NCATG, KA, KB>(xA1, xB1, scA1, scB1, rowA1, rowB1, bA, bB, MAa, MBa, dst, dst_scale, fwa + kFwdStride,
                              fws + kFwdStride, lane, goff + kOffStride, sidx + kSiteStride, (live_mask >> (j + 1)) & 1u,
                              apply_scaling);
    }
    fwa += 2 * kFwdStride;
    fws += 2 * kFwdStride;
    goff += 2 * kOffStride;
    sidx += 2 * kSiteStride;
  }
}
#undef T3_LOAD

// tile -> chunk range.  Tiles are whole PAIRS of chunks (the TMA copies of the staged tip rows start on 16 bytes)
// and differ by at most one pair, so that every block of the persistent grid has work (148 tiles of 66 or 68
// chunks instead of 145 tiles of 68 for 9 811 chunks).
__device__ __forceinline__ void t3_tile_range(int tile, int total_chunks, int n_tiles, int &chunk0, int &n_chunks)
{
  const int pairs = (total_chunks + 1) >> 1;
  const int base = pairs / n_tiles, rem = pairs - base * n_tiles;
  const int p0 = tile * base + min(tile, rem);
  chunk0 = 2 * p0;
  n_chunks = min(2 * (base + (tile < rem ? 1 : 0)), total_chunks - chunk0);
}

template <int NCATG, int W, int MINB>
__global__ void __launch_bounds__((W + 1) * 32, MINB)
    k_traverse_dna3(const OpDev *__restrict__ ops, int n_ops, int total_chunks, int tile_chunks, int n_tiles,
                    const double *__restrict__ wght, int apply_scaling)
{
  constexpr int      S = kT3Stages;
  constexpr uint32_t PB = NCATG * 16 * sizeof(double);
  constexpr uint32_t TB = NCATG * 64 * sizeof(double);
  extern __shared__ __align__(128) unsigned char t3_smem[];
  __shared__ __align__(8) uint64_t               full[S], empty[S];
  T3Stage<NCATG> *st = reinterpret_cast<T3Stage<NCATG> *>(t3_smem);
  unsigned char  *fwd = t3_smem + (size_t)S * sizeof(T3Stage<NCATG>);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], W);
    }
  __syncthreads();
  const int       rounds = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total_it = (long long)rounds * n_ops;

  if (warp == W)
  {  // ---------------- producer warp (same protocol as k_traverse_dna)
    for (long long base = 0; base < total_it; base += 32)
    {
      const long long    my = base + lane;
      unsigned long long m1 = 0, m2 = 0, r1 = 0, r2 = 0;
      int                kd = 0;
      if (my < total_it)
      {
        const OpDev *o = ops + (my % n_ops);
        m1 = (unsigned long long)o->P1;
        m2 = (unsigned long long)o->P2;
        r1 = (unsigned long long)o->t1;
        r2 = (unsigned long long)o->t2;
        kd = o->flags;
      }
      const int cnt = (int)min((long long)32, total_it - base);
      for (int j = 0; j < cnt; ++j)
      {
        const unsigned long long a1 = __shfl_sync(0xffffffffu, m1, j), a2 = __shfl_sync(0xffffffffu, m2, j);
        const unsigned long long q1 = __shfl_sync(0xffffffffu, r1, j), q2 = __shfl_sync(0xffffffffu, r2, j);
        const int                kind = __shfl_sync(0xffffffffu, kd, j);
        if (lane == 0)
        {
          const long long it = base + j;
          const int       s = (int)(it % S);
          const uint32_t  ph = (uint32_t)((it / S) & 1);
          const int       tile = (int)blockIdx.x + (int)(it / n_ops) * (int)gridDim.x;
          int             chunk0, tile_n;
          t3_tile_range(tile, total_chunks, n_tiles, chunk0, tile_n);
          // tip rows of the tile: whole 16-byte units (tiles start on even chunks, rows are padded)
          const uint32_t rb = (uint32_t)((tile_n * 8 + 15) & ~15);
          const bool     tipA = (kind & 3) == kSrcTip, tipB = (kind >> 2) == kSrcTip;
          mbar_wait_backoff(&empty[s], ph ^ 1u);
          const uint32_t b1 = tipA ? TB : PB;
          const uint32_t b2 = tipB ? TB : PB;
          mbar_expect_tx(&full[s], (uint32_t)sizeof(OpDev) + b1 + b2 + (tipA ? rb : 0u) + (tipB ? rb : 0u));
          tma_bulk_g2s(&st[s].op, ops + (it % n_ops), (uint32_t)sizeof(OpDev), &full[s]);
          tma_bulk_g2s(st[s].M[0], (const void *)a1, b1, &full[s]);
          tma_bulk_g2s(st[s].M[1], (const void *)a2, b2, &full[s]);
          if (tipA) tma_bulk_g2s(st[s].rows[0], (const void *)(q1 + (unsigned long long)chunk0 * 8ull), rb, &full[s]);
          if (tipB) tma_bulk_g2s(st[s].rows[1], (const void *)(q2 + (unsigned long long)chunk0 * 8ull), rb, &full[s]);
        }
        __syncwarp();
      }
    }
    return;
  }

  // ---------------- compute warps
  long long it = 0;
  for (int r = 0; r < rounds; ++r)
  {
    const int tile = (int)blockIdx.x + r * (int)gridDim.x;
    int       chunk0, n_chunks;
    t3_tile_range(tile, total_chunks, n_tiles, chunk0, n_chunks);
    const int nch = (n_chunks > warp) ? (n_chunks - 1 - warp) / W + 1 : 0;  // chunks warp, warp + W, ... of the tile
    const int sidx0 = (chunk0 + warp) * 8 + (lane >> 2);
    const int goff0 = (chunk0 + warp) * NCATG * 32 + lane;  // blocked layout, < 2^31 doubles
    const uint32_t cbase = smem_u32(fwd + (size_t)warp * t3_chunk_bytes<NCATG>());
    const uint32_t fwa0 = cbase + lane * 8, fws0 = cbase + NCATG * 256 + lane * 4;
    unsigned       live_mask = 0u;
    for (int j = 0; j < nch; ++j)
      if (wght[sidx0 + j * (W * 8)] > DBL_MIN) live_mask |= 1u << j;  // avx.c:399 (arrays are padded with zero weights)

    for (int k = 0; k < n_ops; ++k, ++it)
    {
      const int                s = (int)(it % S);
      const T3Stage<NCATG> &stg = st[s];
      mbar_wait(&full[s], (uint32_t)((it / S) & 1));
      const int kind = stg.op.flags, ka = kind & 3, kb = kind >> 2;
      if (ka == kSrcFwd)
      {
        if (kb == kSrcTip)
          t3_run_op<NCATG, W, kSrcFwd, kSrcTip>(stg, fwa0, fws0, nch, live_mask, goff0, sidx0, chunk0 * 8, lane, apply_scaling);
        else
          t3_run_op<NCATG, W, kSrcFwd, kSrcSlot>(stg, fwa0, fws0, nch, live_mask, goff0, sidx0, chunk0 * 8, lane, apply_scaling);
      }
      else if (ka == kSrcTip)
        t3_run_op<NCATG, W, kSrcTip, kSrcTip>(stg, fwa0, fws0, nch, live_mask, goff0, sidx0, chunk0 * 8, lane, apply_scaling);
      else if (kb == kSrcTip)
        t3_run_op<NCATG, W, kSrcSlot, kSrcTip>(stg, fwa0, fws0, nch, live_mask, goff0, sidx0, chunk0 * 8, lane, apply_scaling);
      else
        t3_run_op<NCATG, W, kSrcSlot, kSrcSlot>(stg, fwa0, fws0, nch, live_mask, goff0, sidx0, chunk0 * 8, lane, apply_scaling);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1 fused traversal, 20 states, on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
// The per-category update  dst[s][i] = (sum_j P1[i][j] x1[s][j]) * (sum_j P2[i][j] x2[s][j])  is the
// dense contraction (sites x 20) . (20 x 20): sites are the M dimension (8 per m-tile), output states
// the N dimension (3 n-tiles, 20 padded to 24), input states the K dimension (5 k-steps of 4).
//   A fragment (row = lane>>2 = site, col = lane&3): child CLV words, 5 LDG.64 per m-tile and child
//   B fragment (k = lane&3, n = lane>>2): P[n0+n][k0+k], 15 doubles per child, held in REGISTERS for
//     all m-tiles of the warp (loaded once per update and category from the TMA-staged copy in smem)
//   C fragment (row = site, cols 2*(lane&3), +1): 6 doubles per child; the product of the two
//     children's fragments is the output tile, stored as three 16-byte words per thread.
// A warp owns kAaU m-tiles (8 sites each) and loops over the categories; a tip child needs no MMA:
// its fragment is read from the transposed tip table tPx (row = state, or the all-ones row).
// The per-site maximum for the 2^256 rescaling (avx.c:498-510) is tracked while the categories
// stream through; the (rare) rescaling re-reads the thread's own stores.
// Same warp roles / TMA-mbarrier ring as k_traverse_dna; CLVs written by an update are re-read by
// later updates of the same warp (different lanes: __syncwarp() after every update).
constexpr int kAaStages = 3;
constexpr int kAaU = 1;  // m-tiles (of 8 sites) per warp (2 spills at 128 registers and is 40% slower)
constexpr int kAaComputeWarps = 15;  // + 1 producer warp = 512 threads, ONE block per SM (128 registers)
constexpr int kAaThreads = (kAaComputeWarps + 1) * 32;
constexpr int kAaTileCap = kAaComputeWarps * 8 * kAaU;
__host__ __device__ inline size_t aa_stage_bytes(int ncatg) { return 128 + 2 * (size_t)ncatg * 480 * sizeof(double); }

// C fragment of a tip child: u[i] = sum_{j in mask} P[i][j] for the thread's output states
__device__ __forceinline__ void aa_tip_frag(const double *tpx /* [21][20] of this category */, int row, uint32_t mask,
                                            int t, double (&cf)[6])
{
  if (row <= 20)
  {
#pragma unroll
    for (int n = 0; n < 3; ++n)
    {
      const int i0 = n * 8 + 2 * t;
      if (i0 < 20)
      {
        cf[2 * n] = tpx[row * 20 + i0];
        cf[2 * n + 1] = tpx[row * 20 + i0 + 1];
      }
      else
        cf[2 * n] = cf[2 * n + 1] = 0.0;
    }
  }
  else
  {  // general ambiguity code: ascending-j sum of the selected columns (same order as the reference)
#pragma unroll
    for (int n = 0; n < 3; ++n)
    {
      const int i0 = n * 8 + 2 * t;
      double    a0 = 0.0, a1 = 0.0;
      bool      first = true;
      if (i0 < 20)
        for (int j = 0; j < 20; ++j)
          if ((mask >> j) & 1u)
          {
            a0 = first ? tpx[j * 20 + i0] : a0 + tpx[j * 20 + i0];
            a1 = first ? tpx[j * 20 + i0 + 1] : a1 + tpx[j * 20 + i0 + 1];
            first = false;
          }
      cf[2 * n] = a0;
      cf[2 * n + 1] = a1;
    }
  }
}

// general ambiguity code (rare): only instantiated in the non-unrolled copy of the category loop of k_traverse_aa3.  `tpx_a` = shared-memory address of tPx[21][20] of the category.
__device__ __forceinline__ void aa_tip_frag_general(uint32_t tpx_a, uint32_t mask, int t, double (&cf)[6])
{
#pragma unroll
  for (int n = 0; n < 3; ++n)
  {
    const int i0 = n * 8 + 2 * t;
    double    a0 = 0.0, a1 = 0.0;
    bool      first = true;
    if (i0 < 20)
      for (int j = 0; j < 20; ++j)
        if ((mask >> j) & 1u)
        {  // ascending-j sum of the selected columns (same order as the reference)
          double x, y;
          lds128(tpx_a + (uint32_t)(j * 20 + i0) * 8, x, y);
          a0 = first ? x : a0 + x;
          a1 = first ? y : a1 + y;
          first = false;
        }
    cf[2 * n] = a0;
    cf[2 * n + 1] = a1;
  }
}

__global__ void __launch_bounds__(kAaThreads, 1)
    k_traverse_aa(const OpDev *__restrict__ ops, int n_ops, int npat, int ncatg, int tile_sites, int n_tiles,
                  const double *__restrict__ wght, const uint32_t *__restrict__ tipmask,
                  const uint8_t *__restrict__ rows_base, const uint8_t *__restrict__ codes_base, int apply_scaling)
{
  constexpr int                      S = kAaStages;
  extern __shared__ __align__(128) unsigned char aa_smem[];
  __shared__ __align__(8) uint64_t   full[S], empty[S];
  const size_t   stage_bytes = aa_stage_bytes(ncatg);
  const uint32_t PB = (uint32_t)(ncatg * 480 * sizeof(double));  // fragment-ordered P
  const uint32_t TB = (uint32_t)(ncatg * 420 * sizeof(double));  // transposed tip table

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kAaComputeWarps);
    }
  __syncthreads();

  const int       rounds = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long long total_it = (long long)rounds * n_ops;

  if (warp == kAaComputeWarps)
  {  // ---------------- producer warp
    for (long long base = 0; base < total_it; base += 32)
    {
      const long long    my = base + lane;
      unsigned long long m1 = 0, m2 = 0;
      int                kd = 0;
      if (my < total_it)
      {
        const OpDev *o = ops + (my % n_ops);
        m1 = (unsigned long long)o->P1;
        m2 = (unsigned long long)o->P2;
        kd = o->flags;
      }
      const int cnt = (int)min((long long)32, total_it - base);
      for (int j = 0; j < cnt; ++j)
      {
        const unsigned long long a1 = __shfl_sync(0xffffffffu, m1, j), a2 = __shfl_sync(0xffffffffu, m2, j);
        const int                kind = __shfl_sync(0xffffffffu, kd, j);
        if (lane == 0)
        {
          const long long it = base + j;
          const int       s = (int)(it % S);
          const uint32_t  ph = (uint32_t)((it / S) & 1);
          unsigned char  *stg = aa_smem + (size_t)s * stage_bytes;
          mbar_wait_backoff(&empty[s], ph ^ 1u);
          const uint32_t b1 = (kind & 1) ? TB : PB;
          const uint32_t b2 = (kind & 2) ? TB : PB;
          mbar_expect_tx(&full[s], (uint32_t)sizeof(OpDev) + b1 + b2);
          tma_bulk_g2s(stg, ops + (it % n_ops), (uint32_t)sizeof(OpDev), &full[s]);
          tma_bulk_g2s(stg + 128, (const void *)a1, b1, &full[s]);
          tma_bulk_g2s(stg + 128 + (size_t)ncatg * 480 * sizeof(double), (const void *)a2, b2, &full[s]);
        }
        __syncwarp();
      }
    }
    return;
  }

  // ---------------- compute warps
  const int    g = lane >> 2, t = lane & 3;
  const double big = two_to_large(), small = inv_two_to_large();
  long long    it = 0;
  for (int r = 0; r < rounds; ++r)
  {
    const int tile = (int)blockIdx.x + r * (int)gridDim.x;
    const int base_site = tile * tile_sites;
    const int n_sites = min(tile_sites, npat - base_site);
    int       site[kAaU];
    bool      valid[kAaU], live[kAaU];
#pragma unroll
    for (int u = 0; u < kAaU; ++u)
    {
      const int ls = (u * kAaComputeWarps + warp) * 8 + g;
      valid[u] = ls < n_sites;
      site[u] = base_site + (valid[u] ? ls : n_sites - 1);
      live[u] = valid[u] && (wght[site[u]] > DBL_MIN);
    }

    for (int k = 0; k < n_ops; ++k, ++it)
    {
      const int            s = (int)(it % S);
      const uint32_t       ph = (uint32_t)((it / S) & 1);
      const unsigned char *stg = aa_smem + (size_t)s * stage_bytes;
      mbar_wait(&full[s], ph);
      const OpDev  &op = *reinterpret_cast<const OpDev *>(stg);
      const double *M1 = reinterpret_cast<const double *>(stg + 128);
      const double *M2 = M1 + (size_t)ncatg * 480;
      const double *c1 = op.c1, *c2 = op.c2;
      const bool    tip1 = (c1 == nullptr), tip2 = (c2 == nullptr);
      double *const dst = op.dst;

      int      sc[kAaU], row1[kAaU], row2[kAaU];
      uint32_t msk1[kAaU], msk2[kAaU];
      double   mx[kAaU];
#pragma unroll
      for (int u = 0; u < kAaU; ++u)
      {
        sc[u] = 0;
        mx[u] = -DBL_MAX;
        row1[u] = row2[u] = 0;
        msk1[u] = msk2[u] = 0u;
        if (tip1)
        {
          row1[u] = op.t1[site[u]];
          if (row1[u] > 20) msk1[u] = tipmask[codes_base[(op.t1 - rows_base) + site[u]]] & 0xFFFFFu;
        }
        else
          sc[u] += op.s1[site[u]];
        if (tip2)
        {
          row2[u] = op.t2[site[u]];
          if (row2[u] > 20) msk2[u] = tipmask[codes_base[(op.t2 - rows_base) + site[u]]] & 0xFFFFFu;
        }
        else
          sc[u] += op.s2[site[u]];
      }

      for (int c = 0; c < ncatg; ++c)
      {
        double cf1[kAaU][6], cf2[kAaU][6];
        bool   one1[kAaU], one2[kAaU];
        // ---- child 1
        if (!tip1)
        {
          double        bf[15];
          const double *Pf = M1 + (size_t)c * 480 + lane;
#pragma unroll
          for (int j = 0; j < 15; ++j) bf[j] = Pf[j * 32];
#pragma unroll
          for (int u = 0; u < kAaU; ++u)
          {
            const double *row = c1 + (((size_t)(site[u] >> 3) * ncatg + c) * 5) * 32 + (site[u] & 7) * 4 + t;
            double        a[5];
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) a[kk] = ldg64(row + kk * 32);
            const bool mine = (a[0] == 1.0) && (a[1] == 1.0) && (a[2] == 1.0) && (a[3] == 1.0) && (a[4] == 1.0);
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            one1[u] = ((bal >> (lane & ~3)) & 0xFu) == 0xFu;
#pragma unroll
            for (int q = 0; q < 6; ++q) cf1[u][q] = 0.0;
            // k outermost: the three n-tile accumulators advance together (independent MMA chains)
#pragma unroll
            for (int kk = 0; kk < 5; ++kk)
#pragma unroll
              for (int n = 0; n < 3; ++n) dmma884(cf1[u][2 * n], cf1[u][2 * n + 1], a[kk], bf[n * 5 + kk]);
          }
        }
        else
        {
#pragma unroll
          for (int u = 0; u < kAaU; ++u)
          {
            aa_tip_frag(M1 + (size_t)c * 420, row1[u], msk1[u], t, cf1[u]);
            one1[u] = (row1[u] == 20);
          }
        }
        // ---- child 2
        if (!tip2)
        {
          double        bf[15];
          const double *Pf = M2 + (size_t)c * 480 + lane;
#pragma unroll
          for (int j = 0; j < 15; ++j) bf[j] = Pf[j * 32];
#pragma unroll
          for (int u = 0; u < kAaU; ++u)
          {
            const double *row = c2 + (((size_t)(site[u] >> 3) * ncatg + c) * 5) * 32 + (site[u] & 7) * 4 + t;
            double        a[5];
#pragma unroll
            for (int kk = 0; kk < 5; ++kk) a[kk] = ldg64(row + kk * 32);
            const bool mine = (a[0] == 1.0) && (a[1] == 1.0) && (a[2] == 1.0) && (a[3] == 1.0) && (a[4] == 1.0);
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            one2[u] = ((bal >> (lane & ~3)) & 0xFu) == 0xFu;
#pragma unroll
            for (int q = 0; q < 6; ++q) cf2[u][q] = 0.0;
            // k outermost: the three n-tile accumulators advance together (independent MMA chains)
#pragma unroll
            for (int kk = 0; kk < 5; ++kk)
#pragma unroll
              for (int n = 0; n < 3; ++n) dmma884(cf2[u][2 * n], cf2[u][2 * n + 1], a[kk], bf[n * 5 + kk]);
          }
        }
        else
        {
#pragma unroll
          for (int u = 0; u < kAaU; ++u)
          {
            aa_tip_frag(M2 + (size_t)c * 420, row2[u], msk2[u], t, cf2[u]);
            one2[u] = (row2[u] == 20);
          }
        }
        // ---- product, running maximum, store
#pragma unroll
        for (int u = 0; u < kAaU; ++u)
        {
          double      *out = dst + (((size_t)(site[u] >> 3) * ncatg + c) * 5 + (t >> 1)) * 32 + (site[u] & 7) * 4 + 2 * (t & 1);
          const bool   ones = one1[u] && one2[u];  // avx.c:575-587
#pragma unroll
          for (int n = 0; n < 3; ++n)
          {
            if (n * 8 + 2 * t < 20)
            {
              const double o0 = ones ? 1.0 : cf1[u][2 * n] * cf2[u][2 * n];
              const double o1 = ones ? 1.0 : cf1[u][2 * n + 1] * cf2[u][2 * n + 1];
              mx[u] = fmax(mx[u], fmax(o0, o1));
              if (live[u]) stg128(out + n * 64, o0, o1);
            }
          }
        }
      }

      // ---- per-site maximum over all categories and states, rescaling (avx.c:498-510)
#pragma unroll
      for (int u = 0; u < kAaU; ++u)
      {
        double m = mx[u];
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
        int sco = sc[u];
        if (m < small && apply_scaling)
        {
          sco += kLarge;
          if (live[u])
            for (int c = 0; c < ncatg; ++c)
            {
              double *out = dst + (((size_t)(site[u] >> 3) * ncatg + c) * 5 + (t >> 1)) * 32 + (site[u] & 7) * 4 + 2 * (t & 1);
#pragma unroll
              for (int n = 0; n < 3; ++n)
                if (n * 8 + 2 * t < 20)
                {
                  double x, y;
                  ldg128(out + n * 64, x, y);
                  stg128(out + n * 64, x * big, y * big);
                }
            }
        }
        if (live[u] && t == 0) op.dst_scale[site[u]] = sco;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
}

__device__ __forceinline__ const double *lds_ptr(uint32_t a)
{
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return reinterpret_cast<const double *>(v);
}
// byte offsets inside a T4Stage (shared-memory addresses are kept as 32-bit integers: no generic -> shared
// conversion and no 64-bit arithmetic in the item loop)
constexpr uint32_t kT4OffDst = 0, kT4OffDstScale = 8, kT4OffC1 = 16, kT4OffS1 = 24, kT4OffC2 = 40, kT4OffS2 = 48,
                   kT4OffFlags = 80, kT4OffZero = 96, kT4OffM = 128;
static_assert(offsetof(OpDev, c1) == kT4OffC1 && offsetof(OpDev, s1) == kT4OffS1 && offsetof(OpDev, c2) == kT4OffC2 &&
                  offsetof(OpDev, s2) == kT4OffS2 && offsetof(OpDev, flags) == kT4OffFlags &&
                  offsetof(OpDev, dst_scale) == kT4OffDstScale,
              "T4 stage offsets");


// ------------------------------------------------------------------------------------------------
// K1 fused traversal, 20 states, third generation (default for 1, 2, 4 or 8 rate categories): the loop nest
// and the DMMA arithmetic of k_traverse_aa with its overheads removed.  ncu of k_traverse_aa
// (profiles/ncu_r2_aa_v1.md): on B200 DMMA.8x8x4 executes on the FP64 pipe (4 cycles per SM), and that pipe was
// also fed 0.8 DSETP/DMNMX per DMMA by the all-ones test (avx.c:575-587) and the running maximum; there
// were 19 warp instructions per DMMA (integer address arithmetic with a run-time category count, selects,
// predicate logic) against 16 issue slots per DMMA and scheduler at full pipe rate; tip rows were 1-byte
// global loads at the head of every update.  Here:
//  * NCATG is a template parameter, the category loop is unrolled: every fragment address is one base
//    pointer per (update, m-tile) plus an immediate; the A fragments of category c + 1 are in flight while
//    category c is multiplied;
//  * the all-ones rule is a one-compare filter per child and category (exponent word of the first state)
//    in front of a rarely taken exact path; the per-site maximum is an integer maximum of exponent words:
//    the FP64 pipe only sees DMMA and the 6 DMUL per category;
//  * the tile's tip rows are staged per update by the producer warp (TMA); general ambiguity masks are looked
//    up through explicit bases.
// The loop nest stays tile-major (a warp walks the whole update list for its 8 sites): measured, the
// update-major order of k_traverse_dna4 loses here (2.38 vs 2.11 ms) -- a 20-state CLV is 5x larger, one
// update of all m-tiles no longer fits the L2 window and the children come back from DRAM (1.55 vs 0.48 GB read).
// Same fragments, same accumulation order: bit-identical to k_traverse_aa and to the reference.
constexpr int kAa3Stages = 4;
template <int NCATG>
struct __align__(128) Aa3Stage
{
  OpDev   op;
  char    pad[128 - sizeof(OpDev)];
  double  M[2][NCATG * 480];           // per child: fragment-ordered P (480 per category) or tPx (420)
  uint8_t rows[2][kAaTileCap + 24];    // per tip child: tip-table rows of the tile's sites (+ alignment slack)
};

// the category loop of one k_traverse_aa3 item: products, stores, running per-site maximum (returned as the
// largest exponent word).  GEN = a general ambiguity code is present at some lane (rare): not unrolled.
template <int NCATG, bool GEN>
__device__ __forceinline__ int aa3_categories(const double *a1p, const double *a2p, double *outp, uint32_t M1a, uint32_t M2a,
                                              bool tip1, bool tip2, int row1, int row2, uint32_t msk1, uint32_t msk2,
                                              bool live, int lane)
{
  const int t = lane & 3;
    int hm = 0;
          // software pipeline over the categories: the A fragments of category c + 1 are in flight while c is multiplied
          double a1n[5], a2n[5];
  #pragma unroll
          for (int kk = 0; kk < 5; ++kk) a1n[kk] = a2n[kk] = 0.0;
          if (!tip1)
          {
  #pragma unroll
            for (int kk = 0; kk < 5; ++kk) a1n[kk] = ldg64q(a1p + kk * 32);
          }
          if (!tip2)
          {
  #pragma unroll
            for (int kk = 0; kk < 5; ++kk) a2n[kk] = ldg64q(a2p + kk * 32);
          }
          auto one_category = [&](const int c) {
            double a1[5], a2[5], cf1[6], cf2[6];
            bool   f1, f2;  // all-ones filters (avx.c:575-587)
  #pragma unroll
            for (int kk = 0; kk < 5; ++kk)
            {
              a1[kk] = a1n[kk];
              a2[kk] = a2n[kk];
            }
            if (c + 1 < NCATG)
            {
              if (!tip1)
              {
  #pragma unroll
                for (int kk = 0; kk < 5; ++kk) a1n[kk] = ldg64q(a1p + (c + 1) * 160 + kk * 32);
              }
              if (!tip2)
              {
  #pragma unroll
                for (int kk = 0; kk < 5; ++kk) a2n[kk] = ldg64q(a2p + (c + 1) * 160 + kk * 32);
              }
            }
            if (!tip1)
            {
              const uint32_t Pf = M1a + (uint32_t)(c * 480 + lane) * 8;
  #pragma unroll
              for (int q = 0; q < 6; ++q) cf1[q] = 0.0;
              // k outermost: the three n-tile accumulators advance together (independent MMA chains)
  #pragma unroll
              for (int kk = 0; kk < 5; ++kk)
  #pragma unroll
                for (int n = 0; n < 3; ++n) dmma884(cf1[2 * n], cf1[2 * n + 1], a1[kk], lds64(Pf + (n * 5 + kk) * 256));
              f1 = __double2hiint(a1[0]) == 0x3FF00000;
            }
            else
            {
              if (row1 <= 20)
              {
                const uint32_t T = M1a + (uint32_t)(c * 420 + row1 * 20 + 2 * t) * 8;
                lds128(T, cf1[0], cf1[1]);
                lds128(T + 64, cf1[2], cf1[3]);
                cf1[4] = cf1[5] = 0.0;
                if (t < 2) lds128(T + 128, cf1[4], cf1[5]);
              }
              else if constexpr (GEN)
                aa_tip_frag_general(M1a + (uint32_t)(c * 420) * 8, msk1, t, cf1);
              f1 = (row1 == 20);
            }
            if (!tip2)
            {
              const uint32_t Pf = M2a + (uint32_t)(c * 480 + lane) * 8;
  #pragma unroll
              for (int q = 0; q < 6; ++q) cf2[q] = 0.0;
  #pragma unroll
              for (int kk = 0; kk < 5; ++kk)
  #pragma unroll
                for (int n = 0; n < 3; ++n) dmma884(cf2[2 * n], cf2[2 * n + 1], a2[kk], lds64(Pf + (n * 5 + kk) * 256));
              f2 = __double2hiint(a2[0]) == 0x3FF00000;
            }
            else
            {
              if (row2 <= 20)
              {
                const uint32_t T = M2a + (uint32_t)(c * 420 + row2 * 20 + 2 * t) * 8;
                lds128(T, cf2[0], cf2[1]);
                lds128(T + 64, cf2[2], cf2[3]);
                cf2[4] = cf2[5] = 0.0;
                if (t < 2) lds128(T + 128, cf2[4], cf2[5]);
              }
              else if constexpr (GEN)
                aa_tip_frag_general(M2a + (uint32_t)(c * 420) * 8, msk2, t, cf2);
              f2 = (row2 == 20);
            }
            double o[6];
  #pragma unroll
            for (int q = 0; q < 6; ++q) o[q] = cf1[q] * cf2[q];
            if (__any_sync(0xffffffffu, f1 && f2))
            {  // exact test: all 20 states of both children are 1.0 at this (site, category)
              bool one1 = f1, one2 = f2;
              if (!tip1)
              {
                const bool mine = (a1[0] == 1.0) && (a1[1] == 1.0) && (a1[2] == 1.0) && (a1[3] == 1.0) && (a1[4] == 1.0);
                one1 = ((__ballot_sync(0xffffffffu, mine) >> (lane & ~3)) & 0xFu) == 0xFu;
              }
              if (!tip2)
              {
                const bool mine = (a2[0] == 1.0) && (a2[1] == 1.0) && (a2[2] == 1.0) && (a2[3] == 1.0) && (a2[4] == 1.0);
                one2 = ((__ballot_sync(0xffffffffu, mine) >> (lane & ~3)) & 0xFu) == 0xFu;
              }
              if (one1 && one2)
              {
  #pragma unroll
                for (int q = 0; q < 6; ++q) o[q] = 1.0;
              }
            }
            // states n*8 + 2t, +1: the third n-tile only holds states 16..19 (t < 2)
            hm = max(hm, max(max(__double2hiint(o[0]), __double2hiint(o[1])), max(__double2hiint(o[2]), __double2hiint(o[3]))));
            if (t < 2) hm = max(hm, max(__double2hiint(o[4]), __double2hiint(o[5])));
            if (live)
            {
              stg128q(outp + c * 160, o[0], o[1]);
              stg128q(outp + c * 160 + 64, o[2], o[3]);
              if (t < 2) stg128q(outp + c * 160 + 128, o[4], o[5]);
            }
          };
          if constexpr (GEN)
          {
#pragma unroll 1
            for (int c = 0; c < NCATG; ++c) one_category(c);
          }
          else
          {
#pragma unroll
            for (int c = 0; c < NCATG; ++c) one_category(c);
          }
  return hm;
}

template <int NCATG>
__global__ void __launch_bounds__(kAaThreads, 1)
    k_traverse_aa3(const OpDev *__restrict__ ops, int n_ops, int npat, int tile_sites, int n_tiles,
                   const double *__restrict__ wght, const uint32_t *__restrict__ tipmask,
                   const uint8_t *__restrict__ rows_base, const uint8_t *__restrict__ codes_base, int apply_scaling)
{
  constexpr int      S = kAa3Stages, W = kAaComputeWarps;
  constexpr uint32_t PB = NCATG * 480 * sizeof(double);  // fragment-ordered P
  constexpr uint32_t TB = NCATG * 420 * sizeof(double);  // transposed tip table
  constexpr uint32_t kStageB = (uint32_t)sizeof(Aa3Stage<NCATG>);
  constexpr uint32_t kOffM = 128, kOffRows = 128 + 2 * NCATG * 480 * 8, kRowsB = kAaTileCap + 24;
  extern __shared__ __align__(128) unsigned char aa3_smem[];
  __shared__ __align__(8) uint64_t               full[S], empty[S];
  Aa3Stage<NCATG> *st = reinterpret_cast<Aa3Stage<NCATG> *>(aa3_smem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], W);
    }
  __syncthreads();
  const int      rounds = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const unsigned total_it = (unsigned)rounds * (unsigned)n_ops;

  if (warp == W)
  {  // ---------------- producer warp
    for (unsigned base = 0; base < total_it; base += 32)
    {
      const unsigned     my = base + lane;
      unsigned long long m1 = 0, m2 = 0, r1 = 0, r2 = 0;
      int                kd = 0;
      if (my < total_it)
      {
        const OpDev *o = ops + (my % (unsigned)n_ops);
        m1 = (unsigned long long)o->P1;
        m2 = (unsigned long long)o->P2;
        r1 = (unsigned long long)o->t1;
        r2 = (unsigned long long)o->t2;
        kd = o->flags;
      }
      const int cnt = (int)min(32u, total_it - base);
      for (int j = 0; j < cnt; ++j)
      {
        const unsigned long long a1 = __shfl_sync(0xffffffffu, m1, j), a2 = __shfl_sync(0xffffffffu, m2, j);
        const unsigned long long q1 = __shfl_sync(0xffffffffu, r1, j), q2 = __shfl_sync(0xffffffffu, r2, j);
        const int                kind = __shfl_sync(0xffffffffu, kd, j);
        if (lane == 0)
        {
          const unsigned it = base + j;
          const int      s = (int)(it % S);
          const uint32_t ph = (it / S) & 1u;
          const int      tile = (int)blockIdx.x + (int)(it / (unsigned)n_ops) * (int)gridDim.x;
          const unsigned base_site = (unsigned)tile * (unsigned)tile_sites;
          const unsigned off0 = base_site & ~15u;  // 16-byte units from the aligned-down start (rows are padded)
          const uint32_t rb = (base_site - off0 + (unsigned)tile_sites + 15u) & ~15u;
          const bool     tipA = (kind & 1) != 0, tipB = (kind & 2) != 0;
          mbar_wait_backoff(&empty[s], ph ^ 1u);
          const uint32_t b1 = tipA ? TB : PB;
          const uint32_t b2 = tipB ? TB : PB;
          mbar_expect_tx(&full[s], (uint32_t)sizeof(OpDev) + b1 + b2 + (tipA ? rb : 0u) + (tipB ? rb : 0u));
          tma_bulk_g2s(&st[s].op, ops + (it % (unsigned)n_ops), (uint32_t)sizeof(OpDev), &full[s]);
          tma_bulk_g2s(st[s].M[0], (const void *)a1, b1, &full[s]);
          tma_bulk_g2s(st[s].M[1], (const void *)a2, b2, &full[s]);
          if (tipA) tma_bulk_g2s(st[s].rows[0], (const void *)(q1 + off0), rb, &full[s]);
          if (tipB) tma_bulk_g2s(st[s].rows[1], (const void *)(q2 + off0), rb, &full[s]);
        }
        __syncwarp();
      }
    }
    return;
  }

  // ---------------- compute warps: one m-tile (8 sites) per warp and round; lane -> (site g, t)
  const int      g = lane >> 2, t = lane & 3;
  const uint32_t st_a = smem_u32(st), full_a = smem_u32(full), empty_a = smem_u32(empty);
  unsigned       it = 0;
  for (int r = 0; r < rounds; ++r)
  {
    const int  tile = (int)blockIdx.x + r * (int)gridDim.x;
    const int  base_site = tile * tile_sites;
    const int  n_sites = min(tile_sites, npat - base_site);
    const bool active = warp * 8 < n_sites;              // warps beyond the tile only keep the ring protocol
    const int  site = base_site + warp * 8 + g;          // arrays are padded: sites past the end are readable
    const bool live = active && (site < npat) && (wght[site] > DBL_MIN);
    const int  row_off = site - (int)((unsigned)base_site & ~15u);
    const size_t mo = (size_t)(site >> 3) * (NCATG * 160);  // doubles: [m-tile][catg][5][8][4]

    for (int k = 0; k < n_ops; ++k, ++it)
    {
      const uint32_t s = it % S, sa = st_a + s * kStageB;
      {
        const uint32_t fa = full_a + s * 8, par = (it / S) & 1u;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "AA3_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra AA3_DONE;\n\t"
            "bra AA3_WAIT;\n\t"
            "AA3_DONE:\n\t"
            "}" ::"r"(fa),
            "r"(par)
            : "memory");
      }
      if (active)
      {
        const int     kind = lds32(sa + kT4OffFlags);
        const bool    tip1 = (kind & 1) != 0, tip2 = (kind & 2) != 0;
        const double *a1p = lds_ptr(sa + kT4OffC1) + mo + lane, *a2p = lds_ptr(sa + kT4OffC2) + mo + lane;
        double       *outp = const_cast<double *>(lds_ptr(sa + kT4OffDst)) + mo + (t >> 1) * 32 + g * 4 + 2 * (t & 1);
        const uint32_t M1a = sa + kOffM, M2a = M1a + NCATG * 480 * 8;
        int      sc = 0, row1 = 0, row2 = 0;
        uint32_t msk1 = 0u, msk2 = 0u;
        if (tip1)
        {
          row1 = (int)lds8(sa + kOffRows + (uint32_t)row_off);
          if (row1 > 20)
            msk1 = tipmask[codes_base[(reinterpret_cast<const uint8_t *>(lds_ptr(sa + 32)) - rows_base) + site]] & 0xFFFFFu;
        }
        else
          sc += ldg32q(reinterpret_cast<const int *>(lds_ptr(sa + kT4OffS1)) + site);
        if (tip2)
        {
          row2 = (int)lds8(sa + kOffRows + kRowsB + (uint32_t)row_off);
          if (row2 > 20)
            msk2 = tipmask[codes_base[(reinterpret_cast<const uint8_t *>(lds_ptr(sa + 56)) - rows_base) + site]] & 0xFFFFFu;
        }
        else
          sc += ldg32q(reinterpret_cast<const int *>(lds_ptr(sa + kT4OffS2)) + site);

        // general ambiguity codes (rare) take a copy of the category loop that is not unrolled: inlined into the
        // unrolled loop their code (2 children x NCATG copies) quadrupled the kernel
        const bool gen = __any_sync(0xffffffffu, (tip1 && row1 > 20) || (tip2 && row2 > 20));
        const int  hm0 = gen ? aa3_categories<NCATG, true>(a1p, a2p, outp, M1a, M2a, tip1, tip2, row1, row2, msk1, msk2, live, lane)
                             : aa3_categories<NCATG, false>(a1p, a2p, outp, M1a, M2a, tip1, tip2, row1, row2, msk1, msk2, live, lane);
        int        hm = hm0;
        // ---- per-site maximum over all categories and states (exponent words: all entries >= 0; NaN counts as
        // large like the reference), rescaling (avx.c:498-510)
        hm = max(hm, __shfl_xor_sync(0xffffffffu, hm, 1));
        hm = max(hm, __shfl_xor_sync(0xffffffffu, hm, 2));
        const bool resc = ((unsigned)hm < 0x2FF00000u) && apply_scaling;
        if (__any_sync(0xffffffffu, resc))
        {
          if (resc)
          {
            sc += kLarge;
            if (live)
            {
              const double big = two_to_large();
#pragma unroll 1
              for (int c = 0; c < NCATG; ++c)
#pragma unroll
                for (int n = 0; n < 3; ++n)
                  if (n * 8 + 2 * t < 20)
                  {
                    double x, y;
                    ldg128(outp + c * 160 + n * 64, x, y);
                    stg128(outp + c * 160 + n * 64, x * big, y * big);
                  }
            }
          }
        }
        if (live && t == 0) stg32q(reinterpret_cast<int *>(const_cast<double *>(lds_ptr(sa + kT4OffDstScale))) + site, sc);
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a + s * 8) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1 generic: thread per site, any ns <= 32, any ncatg <= 16.  Same arithmetic order.
__device__ __forceinline__ double child_dot(const double *__restrict__ Prow, const double *__restrict__ v, uint32_t mask,
                                            bool internal, int ns)
{
  if (internal)
  {
    double a = Prow[0] * v[0];
    for (int j = 1; j < ns; ++j) a = fma(Prow[j], v[j], a);
    return a;
  }
  double a = (mask & 1u) ? Prow[0] : 0.0;
  for (int j = 1; j < ns; ++j)
    if ((mask >> j) & 1u) a = a + Prow[j];
  return a;
}

__global__ void __launch_bounds__(128)
    k_partial_generic(const OpDev *__restrict__ ops, int blocks_per_op, int npat, int ns, int ncatg,
                      const double *__restrict__ wght, const uint32_t *__restrict__ tipmask, int apply_scaling)
{
  const OpDev    op = ops[blockIdx.x / blocks_per_op];
  const int      blk = blockIdx.x % blocks_per_op;
  const int      ncns = ncatg * ns, nn = ns * ns;
  const uint32_t full = (ns >= 32) ? 0xffffffffu : ((1u << ns) - 1u);
  const double   big = two_to_large(), small = inv_two_to_large();

  for (int site = blk * blockDim.x + threadIdx.x; site < npat; site += blocks_per_op * blockDim.x)
  {
    if (!(wght[site] > DBL_MIN)) continue;
    const double  *v1 = op.c1 ? op.c1 + (size_t)site * ncns : nullptr;
    const double  *v2 = op.c2 ? op.c2 + (size_t)site * ncns : nullptr;
    const uint32_t m1 = op.c1 ? 0u : tipmask[op.t1[site]];
    const uint32_t m2 = op.c2 ? 0u : tipmask[op.t2[site]];
    double        *out = op.dst + (size_t)site * ncns;
    double         largest = -DBL_MAX;
    for (int c = 0; c < ncatg; ++c)
    {
      const double *a1 = v1 ? v1 + c * ns : nullptr;
      const double *a2 = v2 ? v2 + c * ns : nullptr;
      bool          ones = (v1 ? true : m1 == full) && (v2 ? true : m2 == full);
      if (ones && v1)
        for (int j = 0; j < ns; ++j) ones = ones && (a1[j] == 1.0);
      if (ones && v2)
        for (int j = 0; j < ns; ++j) ones = ones && (a2[j] == 1.0);
      for (int i = 0; i < ns; ++i)
      {
        double o;
        if (ones)
          o = 1.0;
        else
          o = child_dot(op.P1 + c * nn + i * ns, a1, m1, v1 != nullptr, ns) *
              child_dot(op.P2 + c * nn + i * ns, a2, m2, v2 != nullptr, ns);
        out[c * ns + i] = o;
        largest = fmax(largest, o);
      }
    }
    int sc = (op.s1 ? op.s1[site] : 0) + (op.s2 ? op.s2[site] : 0);
    if (largest < small && apply_scaling)
    {
      for (int k = 0; k < ncns; ++k) out[k] *= big;
      sc += kLarge;
    }
    op.dst_scale[site] = sc;
  }
}

// ------------------------------------------------------------------------------------------------
// result block shared with the host through mapped pinned memory
struct ResultHost
{
  double            val[2];
  int               warn;
  int               pad;
  unsigned long long seq;
};

// where a reduction kernel delivers its result
struct ReduceOut
{
  double              *partials;  // [gridDim.x][NV]
  unsigned int        *ticket;    // last-block detection
  int                 *warn_flag;
  double              *dev_out;   // [3]: values, warning as double (for the all-reduce)
  volatile ResultHost *host_out;
  unsigned long long   seq;
  unsigned long long   coll_seq;  // sequence number of the cross-GPU exchange (reset when the mailboxes are wired)
  int                  publish;   // 1: write host_out here; 0: an all-reduce + k_publish follow
  // fused cross-GPU sum over NVLink peer memory (world > 1, p2p mode): every rank owns a mailbox
  // [2 parities][world senders] of P2pSlot; peers[q] is rank q's mailbox mapped into this process.
  struct P2pSlot     **peers;
  int                  rank, world;
};

constexpr unsigned long long kP2pTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s
constexpr int                kWarnPeerTimeout = 2;  // bit 1 of ResultHost::warn: the cross-GPU exchange timed out

struct __align__(32) P2pSlot
{
  double             v[3];
  unsigned long long seq;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Deterministic reduction of up to 2 running sums in ONE launch: fixed shuffle tree per warp, fixed
// warp order per block, per-block partials to global memory; the block that finishes last (atomic
// ticket) adds the partials in index order and publishes the result.  The order of the additions
// does not depend on which block happens to be last.
template <int NV>
__device__ __forceinline__ void block_reduce_finish(double (&v)[NV], int warn, const ReduceOut &ro)
{
  __shared__ double sred[NV][32];
  __shared__ int    swarn;
  __shared__ bool   is_last;
  if (threadIdx.x == 0) swarn = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], d);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) sred[k][wid] = v[k];
  if (warn) atomicOr(&swarn, 1);
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
    {
      double sacc = 0.0;
      for (int w = 0; w < nw; ++w) sacc += sred[k][w];
      ro.partials[(size_t)blockIdx.x * NV + k] = sacc;
    }
    if (swarn) atomicOr(ro.warn_flag, 1);
    __threadfence();
    const unsigned int t = atomicAdd(ro.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last block: the first (up to) 128 threads add a strided subset in index order, thread 0 adds the per-thread sums
  __shared__ double sfin[NV][128];
  const int         nfin = min((int)blockDim.x, 128);
  if ((int)threadIdx.x < nfin)
    for (int k = 0; k < NV; ++k)
    {
      double a = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += nfin) a += ro.partials[(size_t)b * NV + k];
      sfin[k][threadIdx.x] = a;
    }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double r[2] = {0.0, 0.0};
    for (int k = 0; k < NV; ++k)
      for (int t = 0; t < nfin; ++t) r[k] += sfin[k][t];
    int w = *ro.warn_flag;
    *ro.warn_flag = 0;
    *ro.ticket = 0u;
    if (ro.world > 1 && ro.peers)
    {
      // reduction + collective in one kernel: post this rank's partial into every rank's mailbox
      // (remote stores over NVLink), wait for all ranks, add in RANK ORDER (bitwise identical result
      // on every rank).  Two parity slots: a rank can be at most one evaluation ahead of its readers.
      const int par = (int)(ro.coll_seq & 1ull);
      for (int q = 0; q < ro.world; ++q)
      {
        P2pSlot *slot = ro.peers[q] + par * ro.world + ro.rank;
        slot->v[0] = r[0];
        slot->v[1] = r[1];
        slot->v[2] = (double)w;
        st_release_sys_u64(&slot->seq, ro.coll_seq);
      }
      double acc[3] = {0.0, 0.0, 0.0};
      bool   timed_out = false;
      unsigned long long t_start = 0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
      for (int q = 0; q < ro.world && !timed_out; ++q)
      {
        const P2pSlot *slot = ro.peers[ro.rank] + par * ro.world + q;
        unsigned       polls = 0;
        while (ld_acquire_sys_u64(&slot->seq) != ro.coll_seq)
        {
          __nanosleep(40);
          if ((++polls & 0x3ffu) == 0)
          {  // a peer that died or fell out of step must not hang this GPU: give up after kP2pTimeoutNs
            unsigned long long now = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t_start > kP2pTimeoutNs)
            {
              timed_out = true;
              break;
            }
          }
        }
        acc[0] += slot->v[0];
        acc[1] += slot->v[1];
        acc[2] += slot->v[2];
      }
      r[0] = acc[0];
      r[1] = acc[1];
      w = (acc[2] > 0.0 ? 1 : 0) | (timed_out ? kWarnPeerTimeout : 0);
    }
    ro.dev_out[0] = r[0];
    ro.dev_out[1] = r[1];
    ro.dev_out[2] = (double)w;
    if (ro.publish)
    {
      ro.host_out->val[0] = r[0];
      ro.host_out->val[1] = r[1];
      ro.host_out->warn = w;
      __threadfence_system();
      ro.host_out->seq = ro.seq;
    }
  }
}

// publishes an (all-reduced) device result to the host
__global__ void k_publish(const double *dev_out, volatile ResultHost *host_out, unsigned long long seq)
{
  host_out->val[0] = dev_out[0];
  host_out->val[1] = dev_out[1];
  host_out->warn = dev_out[2] > 0.0 ? 1 : 0;
  __threadfence_system();
  host_out->seq = seq;
}

// +I term, lk.c:1226-1273
__device__ __forceinline__ double invariant_lk(int fact, int invar_state, const double *pi, bool *overflow)
{
  double v = 0.0;
  *overflow = false;
  if (invar_state > -1)
  {
    int e = fact;
    v = pi[invar_state];
    do
    {
      const int piece = e < 63 ? e : 63;
      v *= (double)(1ULL << piece);
      e -= piece;
    } while (e != 0);
    if (isinf(v)) *overflow = true;
  }
  return v;
}

// horizontal sum in the order of AVX_Vect_Norm (avx.c:281-289)
__device__ __forceinline__ double hsum4(double x0, double x1, double x2, double x3) { return (x0 + x2) + (x1 + x3); }

// ------------------------------------------------------------------------------------------------
// K2: thread per site.
__global__ void __launch_bounds__(128)
    k_edge_lnl(SideDev left, SideDev rght, const double *__restrict__ P, const ModelDev *__restrict__ mod, int npat,
               int ns, int ncatg, const double *__restrict__ wght, const short *__restrict__ invar,
               const uint32_t *__restrict__ tipmask, double *__restrict__ site_lnl, double *__restrict__ site_lk_out,
               double *__restrict__ site_lk_cat, int *__restrict__ fact_sum_scale, ReduceOut ro, int blocked)
{
  const int    nn = ns * ns;
  const double *pi = mod->pi;
  double       acc[1] = {0.0};
  int          warn = 0;

  for (int site = blockIdx.x * blockDim.x + threadIdx.x; site < npat; site += gridDim.x * blockDim.x)
  {
    const double w = wght[site];
    if (!(w > DBL_MIN)) continue;  // lk.c:632
    const uint32_t lm = left.clv ? 0u : tipmask[left.tip[site]];
    const uint32_t rm = rght.clv ? 0u : tipmask[rght.tip[site]];
    const bool     unamb = (!rght.clv) && (__popc(rm) == 1);  // lk.c:614-621
    const int      st = unamb ? (__ffs(rm) - 1) : -1;
    double         site_lk = 0.0;

    for (int c = 0; c < ncatg; ++c)
    {
      const double *Pc = P + (size_t)c * nn;
      // operand accessors: a CLV in its (plain or blocked) layout, or the 0/1 vector of a tip mask
      auto LV = [&](int l) -> double {
        return left.clv ? left.clv[clv_off(site, c, l, ncatg, ns, blocked)] : (double)((lm >> l) & 1u);
      };
      auto RV = [&](int k) -> double {
        return rght.clv ? rght.clv[clv_off(site, c, k, ncatg, ns, blocked)] : (double)((rm >> k) & 1u);
      };
      double lk;
      if ((ns & 3) == 0)
      {  // order of AVX_Lk_Core_One_Class_No_Eigen_Lr (avx.c:110-215)
        if (unamb)
        {
          double q[4] = {0.0, 0.0, 0.0, 0.0};
          for (int b = 0; b < ns; b += 4)
#pragma unroll
            for (int t = 0; t < 4; ++t) q[t] = q[t] + Pc[st * ns + b + t] * LV(b + t);
          lk = pi[st] * hsum4(q[0], q[1], q[2], q[3]);
        }
        else
        {
          lk = 0.0;
          for (int b = 0; b < ns; b += 4)
          {
            double x[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
            {
              const int    k = b + t;
              const double rv = RV(k);
              double       a = 0.0;
              for (int l = 0; l < ns; ++l) a = fma(Pc[k * ns + l], LV(l), a);
              x[t] = a * (rv * pi[k]);
            }
            lk = lk + hsum4(x[0], x[1], x[2], x[3]);
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
      site_lk_cat[(size_t)site * ncatg + c] = lk;  // lk.c:2801
      site_lk = site_lk + lk * mod->probs[c];        // lk.c:818
    }

    int fact = (left.scale ? left.scale[site] : 0) + (rght.scale ? rght.scale[site] : 0);  // lk.c:2781-2791
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
    {  // lk.c:847-851
      site_lk = DBL_MIN;
      warn = 1;
    }
    const double lsl = log(site_lk) - kLog2 * fact;  // lk.c:854
    site_lnl[site] = lsl;
    site_lk_out[site] = exp(lsl);  // lk.c:857
    fact_sum_scale[site] = fact;
    acc[0] += w * lsl;  // lk.c:856
  }
  block_reduce_finish<1>(acc, warn, ro);
}

// ------------------------------------------------------------------------------------------------
// K2, 4 states: thread per (site, category) on the blocked layout (the warp reads 1 KB contiguous);
// lane map as in k_traverse_dna (8 lanes per category), category terms gathered to the category-0
// lane with shuffles and added in category order, i.e. the same arithmetic as k_edge_lnl.
// Everything the edge reduction reads and writes (also the argument block of the traversal kernel's fused
// epilogue: Post_Order_Lk + the site loop of Lk in ONE launch, lk.c:562-645).
struct EdgeDev
{
  SideDev         left, rght;
  const double   *P;  // P-matrix record of the edge
  const ModelDev *mod;
  const double   *wght;
  const short    *invar;
  const uint32_t *tipmask;
  double         *site_lnl, *site_lk_out, *site_lk_cat;
  int            *fact_sum_scale;
  int             npat;
  int             enabled;
  ReduceOut       ro;
};

// per-thread constants of the edge reduction: this lane's category of P, pi, the category weight
template <int NCATG>
struct EdgeLaneConst
{
  double p[16], pi[4], wc;
  __device__ __forceinline__ void load(const EdgeDev &e, int lane)
  {
    const int cat = lane / (32 / NCATG);
#pragma unroll
    for (int q = 0; q < 16; ++q) p[q] = e.P[cat * 16 + q];
#pragma unroll
    for (int q = 0; q < 4; ++q) pi[q] = e.mod->pi[q];
    wc = e.mod->probs[cat];
  }
};

// one group of 32 / NCATG sites x NCATG categories (one warp)
template <int NCATG>
__device__ __forceinline__ void edge_lnl_dna_group(const EdgeDev &e, const EdgeLaneConst<NCATG> &k, int grp, int lane,
                                                   double &acc, int &warn)
{
  constexpr int   SW = 32 / NCATG;
  const int       cat = lane / SW;
  const SideDev  &left = e.left, &rght = e.rght;
  const int       npat = e.npat;
  const double (&p)[16] = k.p;
  const double (&pi)[4] = k.pi;
  {
    const int    site0 = grp * SW + (lane % SW);
    const bool   valid = site0 < npat;
    const int    site = valid ? site0 : npat - 1;
    const double w = e.wght[site];
    const bool   live = valid && (w > DBL_MIN);  // lk.c:632
    const size_t off = ((((size_t)(site >> 3) * NCATG + cat) << 3) + (site & 7)) * 4;
    double       L[4], R[4];
    if (left.clv)
    {
      const double4a v = ldg256(left.clv + off);
      L[0] = v.x; L[1] = v.y; L[2] = v.z; L[3] = v.w;
    }
    else
    {
      const uint32_t m = e.tipmask[left.tip[site]];
#pragma unroll
      for (int q = 0; q < 4; ++q) L[q] = (double)((m >> q) & 1u);
    }
    uint32_t rm = 0u;
    if (rght.clv)
    {
      const double4a v = ldg256(rght.clv + off);
      R[0] = v.x; R[1] = v.y; R[2] = v.z; R[3] = v.w;
    }
    else
    {
      rm = e.tipmask[rght.tip[site]];
#pragma unroll
      for (int q = 0; q < 4; ++q) R[q] = (double)((rm >> q) & 1u);
    }
    const bool unamb = (!rght.clv) && (__popc(rm) == 1);  // lk.c:614-621
    double     lk;
    if (unamb)
    {  // avx.c:119-124
      const int st = __ffs(rm) - 1;
      double    q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, ps = 0.0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        if (kk == st)
        {
          q0 = p[kk * 4 + 0] * L[0];
          q1 = p[kk * 4 + 1] * L[1];
          q2 = p[kk * 4 + 2] * L[2];
          q3 = p[kk * 4 + 3] * L[3];
          ps = pi[kk];
        }
      lk = ps * hsum4(q0, q1, q2, q3);
    }
    else
    {  // avx.c:125-149
      double x[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
      {
        double a = 0.0;
#pragma unroll
        for (int l = 0; l < 4; ++l) a = fma(p[kk * 4 + l], L[l], a);
        x[kk] = a * (R[kk] * pi[kk]);
      }
      lk = hsum4(x[0], x[1], x[2], x[3]);
    }
    if (live) e.site_lk_cat[(size_t)site * NCATG + cat] = lk;  // lk.c:2801
    // lk.c:816-818: site_lk = sum_c lk_c w_c in category order, on the category-0 lane
    double term = lk * k.wc;
    double site_lk = 0.0;
#pragma unroll
    for (int c = 0; c < NCATG; ++c) site_lk = site_lk + __shfl_sync(0xffffffffu, term, c * SW + (lane % SW));
    if (cat == 0 && live)
    {
      const ModelDev *mod = e.mod;
      int fact = (left.scale ? left.scale[site] : 0) + (rght.scale ? rght.scale[site] : 0);  // lk.c:2781-2791
      if (mod->invar_flag)
      {  // lk.c:820-842
        bool   ovf;
        double inv = invariant_lk(fact, e.invar[site], mod->pi, &ovf);
        if (ovf)
        {
          fact = 0;
          inv = invariant_lk(0, e.invar[site], mod->pi, &ovf);
          site_lk = inv * mod->pinv;
        }
        else
          site_lk = site_lk * (1. - mod->pinv) + inv * mod->pinv;
      }
      if (site_lk < DBL_MIN)
      {  // lk.c:847-851
        site_lk = DBL_MIN;
        warn = 1;
      }
      const double lsl = log(site_lk) - kLog2 * fact;  // lk.c:854
      e.site_lnl[site] = lsl;
      e.site_lk_out[site] = exp(lsl);  // lk.c:857
      e.fact_sum_scale[site] = fact;
      acc += w * lsl;  // lk.c:856
    }
  }
}

template <int NCATG>
__global__ void __launch_bounds__(512) k_edge_lnl_dna(const __grid_constant__ EdgeDev e)
{
  constexpr int        SW = 32 / NCATG;
  const int            lane = threadIdx.x & 31;
  const int            warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5;
  EdgeLaneConst<NCATG> k;
  k.load(e, lane);
  double    acc[1] = {0.0};
  int       warn = 0;
  const int groups = (e.npat + SW - 1) / SW;
  for (int grp = blockIdx.x * warps_per_block + warp; grp < groups; grp += gridDim.x * warps_per_block)
    edge_lnl_dna_group<NCATG>(e, k, grp, lane, acc[0], warn);
  block_reduce_finish<1>(acc, warn, e.ro);
}

// ------------------------------------------------------------------------------------------------
// K1 fused traversal, 4 states, fourth generation (default).  ncu of k_traverse_dna3
// (profiles/ncu_r2_dna3.md): 132 warp instructions per (8 sites x 4 categories x update) of which ~35 are
// loads / DMMA / DMUL / stores; half of the lanes idle in the epilogue (a C fragment has 8 columns, 4 states
// fill 4); and a block's chunks are dealt to its warps once for the whole launch, so with 4.4 chunks per warp
// every update waits for the warps that hold 5.  Three changes:
//  * a chunk is TWO 8-site blocks X and Y.  Per category and child the update is two chained DMMAs:
//    C = x_X . B_lo, then C += x_Y . B_hi, where B_lo holds P^T in columns 0..3 (zeros in 4..7) and B_hi holds
//    it in columns 4..7.  Columns 0..3 of C are then exactly P.x of the X sites (the second DMMA adds
//    +0.0 to them) and columns 4..7 exactly P.x of the Y sites (ascending-k FMA chain from +0.0, i.e. the
//    rounding of the reference's AVX_Matrix_Vect_Prod, tools/probes/dmma_order.cu): lanes t < 2 hold
//    states 2t, 2t+1 of X site g, lanes t >= 2 states 2(t-2), 2(t-2)+1 of Y site g.  All 32 lanes
//    multiply, take the per-site maximum, rescale and store 16 bytes: half the epilogue instructions
//    per site, and a store instruction writes two contiguous 256-byte blocks.
//  * work items (update k, chunk i) are dealt round-robin over the compute warps ACROSS updates
//    (item n = k * C + i goes to warp n mod W): every warp gets the same number of items whatever C is.
//    An item needs the previous update of the same chunk, done by another warp: a per-chunk progress counter
//    in shared memory (release / acquire at CTA scope) orders them.  The previous result is forwarded
//    through the chunk's shared-memory block as before, now between warps.
//  * no operand double buffering (the dependent wait is hidden by the other warps): fewer registers, and
//    the specialised item bodies are straight-line code.
// Shared memory per block: kT4Stages ring stages + per chunk NCATG x 640 + 64 bytes of forwarding block,
// a 4-byte progress counter and 16 live flags.
constexpr int kT4Stages = 4;
constexpr int kT4MaxTileChunks = 64;  // 1024 sites: 1 KB of tip rows per operand and stage
template <int NCATG>
struct __align__(128) T4Stage
{
  OpDev   op;  // 96 bytes
  char    pad[128 - sizeof(OpDev)];
  double  M[2][NCATG * 64];                // per child: P[cat][i][j] (NCATG*16) or TP[cat][mask][i] (NCATG*64)
  uint8_t rows[2][kT4MaxTileChunks * 16];  // per tip child: tip-table row of every site of the tile
};
template <int NCATG>
__host__ __device__ constexpr int t4_chunk_bytes()
{
  return NCATG * 640 + 64;  // per category: X block at +0, Y block at +320 (conflict-free STS.128); 16 scalers
}
template <int NCATG>
__host__ __device__ inline size_t t4_smem_bytes(int tile_chunks)
{
  return (size_t)kT4Stages * sizeof(T4Stage<NCATG>) + (size_t)tile_chunks * (t4_chunk_bytes<NCATG>() + 4 + 16);
}

// one item: the update staged at shared address `sa` applied to the 16 sites of chunk `chunkg` (global chunk
// index); `fw` = the chunk's forwarding block, `rw` = staged tip rows biased by the tile's first site
template <int NCATG, int KA, int KB>
__device__ __forceinline__ void t4_item(uint32_t sa, const double (&bAlo)[NCATG], const double (&bAhi)[NCATG],
                                        const double (&bBlo)[NCATG], const double (&bBhi)[NCATG], uint32_t fw,
                                        int chunkg, uint32_t rw, bool live, int lane, int apply_scaling)
{
  constexpr uint32_t kRowsB = kT4MaxTileChunks * 16;
  const int          g = lane >> 2, t = lane & 3, hi = t >> 1, os = hi * 8 + g;
  const int          site = chunkg * 16 + os;                  // this lane's own site (epilogue)
  const int          goff = chunkg * (2 * NCATG * 32) + lane;  // A fragments: (block X, category 0) + lane, < 2^31 doubles
  const uint32_t     fws = fw + NCATG * 640 + os * 4;
  double             xAX[NCATG], xAY[NCATG], xBX[NCATG], xBY[NCATG];
  int                scA = 0, scB = 0;
  uint32_t           rowA = 0, rowB = 0;
  if (KA == kSrcSlot)
  {
    const double *c1 = lds_ptr(sa + kT4OffC1) + goff;
#pragma unroll
    for (int c = 0; c < NCATG; ++c)
    {
      xAX[c] = ldg64q(c1 + c * 32);
      xAY[c] = ldg64q(c1 + (NCATG + c) * 32);
    }
    scA = ldg32q(reinterpret_cast<const int *>(lds_ptr(sa + kT4OffS1)) + site);
  }
  if (KB == kSrcSlot)
  {
    const double *c2 = lds_ptr(sa + kT4OffC2) + goff;
#pragma unroll
    for (int c = 0; c < NCATG; ++c)
    {
      xBX[c] = ldg64q(c2 + c * 32);
      xBY[c] = ldg64q(c2 + (NCATG + c) * 32);
    }
    scB = ldg32q(reinterpret_cast<const int *>(lds_ptr(sa + kT4OffS2)) + site);
  }
  if (KA == kSrcFwd)
  {
#pragma unroll
    for (int c = 0; c < NCATG; ++c)
    {
      xAX[c] = lds64(fw + c * 640 + lane * 8);
      xAY[c] = lds64(fw + c * 640 + 320 + lane * 8);
    }
    scA = lds32(fws);
  }
  if (KA == kSrcTip) rowA = lds8(rw + (uint32_t)site);
  if (KB == kSrcTip) rowB = lds8(rw + kRowsB + (uint32_t)site);

  const uint32_t MAa = sa + kT4OffM + (uint32_t)(2 * (t & 1)) * 8, MBa = MAa + NCATG * 64 * 8;
  double         o0[NCATG], o1[NCATG];
#pragma unroll
  for (int c = 0; c < NCATG; ++c)
  {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    if (KA == kSrcTip)
      lds128(MAa + (uint32_t)(c * 16 * 32) + rowA * 32, a0, a1);
    else
    {
      dmma884(a0, a1, xAX[c], bAlo[c]);
      dmma884(a0, a1, xAY[c], bAhi[c]);
    }
    if (KB == kSrcTip)
      lds128(MBa + (uint32_t)(c * 16 * 32) + rowB * 32, b0, b1);
    else
    {
      dmma884(b0, b1, xBX[c], bBlo[c]);
      dmma884(b0, b1, xBY[c], bBhi[c]);
    }
    o0[c] = a0 * b0;
    o1[c] = a1 * b1;
  }
  // avx.c:575-587: both children all ones at a (site, category) (only below fully ambiguous tips) -> the
  // result is exactly 1.0.  Conservative filter on child A; exact test in the (rare) slow path.
  bool maybe = false;
  if (KA == kSrcTip)
    maybe = (rowA == (uint32_t)kTipRowAllOnes);
  else
  {
#pragma unroll
    for (int c = 0; c < NCATG; ++c)
      maybe = maybe || (__double2hiint(xAX[c]) == 0x3FF00000) || (__double2hiint(xAY[c]) == 0x3FF00000);
  }
  if (__any_sync(0xffffffffu, maybe))
  {
#pragma unroll
    for (int c = 0; c < NCATG; ++c)
    {
      bool pa, pb;
      if (KA == kSrcTip)
        pa = (rowA == (uint32_t)kTipRowAllOnes);
      else
      {
        const unsigned bx = __ballot_sync(0xffffffffu, xAX[c] == 1.0), by = __ballot_sync(0xffffffffu, xAY[c] == 1.0);
        pa = ((((hi ? by : bx) >> (g * 4)) & 0xFu) == 0xFu);
      }
      if (KB == kSrcTip)
        pb = (rowB == (uint32_t)kTipRowAllOnes);
      else
      {
        const unsigned bx = __ballot_sync(0xffffffffu, xBX[c] == 1.0), by = __ballot_sync(0xffffffffu, xBY[c] == 1.0);
        pb = ((((hi ? by : bx) >> (g * 4)) & 0xFu) == 0xFu);
      }
      if (pa && pb) o0[c] = o1[c] = 1.0;
    }
  }
  // avx.c:498-510: per-site maximum over all categories and states as an exponent-word compare (all entries
  // are >= 0; NaN counts as large, like the reference); the two lanes of a site are neighbours
  int hmax = 0;
#pragma unroll
  for (int c = 0; c < NCATG; ++c) hmax = max(hmax, max(__double2hiint(o0[c]), __double2hiint(o1[c])));
  hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, 1));
  int        sco = scA + scB;
  const bool resc = ((unsigned)hmax < 0x2FF00000u) && apply_scaling;
  if (__any_sync(0xffffffffu, resc))
  {
    if (resc)
    {
      const double big = two_to_large();
#pragma unroll
      for (int c = 0; c < NCATG; ++c)
      {
        o0[c] *= big;
        o1[c] *= big;
      }
      sco += kLarge;
    }
  }
  if (live)
  {
    double *dst = const_cast<double *>(lds_ptr(sa + kT4OffDst)) + ((chunkg * 2 + hi) * NCATG * 32 + g * 4 + 2 * (t & 1));
#pragma unroll
    for (int c = 0; c < NCATG; ++c) stg128q(dst + c * 32, o0[c], o1[c]);
    if ((t & 1) == 0) stg32q(reinterpret_cast<int *>(const_cast<double *>(lds_ptr(sa + kT4OffDstScale))) + site, sco);
  }
  const uint32_t fwo = fw + hi * 320 + g * 32 + (t & 1) * 16;
#pragma unroll
  for (int c = 0; c < NCATG; ++c) sts128(fwo + c * 640, o0[c], o1[c]);
  if ((t & 1) == 0) sts32(fws, sco);
}

template <int NCATG, int W, bool PF>
__global__ void __launch_bounds__((W + 1) * 32, 1)
    k_traverse_dna4(const OpDev *__restrict__ ops, int n_ops, int total_chunks, int max_tile_chunks, int n_tiles,
                    const double *__restrict__ wght, int apply_scaling, const __grid_constant__ EdgeDev edge)
{
  constexpr int      S = kT4Stages;
  static_assert((S & (S - 1)) == 0, "ring depth must be a power of two");
  constexpr uint32_t PB = NCATG * 16 * sizeof(double);
  constexpr uint32_t TB = NCATG * 64 * sizeof(double);
  constexpr uint32_t kStageB = (uint32_t)sizeof(T4Stage<NCATG>);
  extern __shared__ __align__(128) unsigned char t4_smem[];
  __shared__ __align__(8) uint64_t               full[S], empty[S];
  T4Stage<NCATG> *st = reinterpret_cast<T4Stage<NCATG> *>(t4_smem);
  unsigned char  *fwd = t4_smem + (size_t)S * sizeof(T4Stage<NCATG>);
  int            *done = reinterpret_cast<int *>(fwd + (size_t)max_tile_chunks * t4_chunk_bytes<NCATG>());
  unsigned char  *livef = reinterpret_cast<unsigned char *>(done + max_tile_chunks);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    for (int s = 0; s < S; ++s)
    {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], W);
    }
  if (tid < S * 4) reinterpret_cast<double *>(st[tid >> 2].pad)[tid & 3] = 0.0;  // the zero B-fragment words
  __syncthreads();
  const int rounds = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // tile -> chunks: tiles differ by at most one chunk
  const int tbase = total_chunks / n_tiles, trem = total_chunks - tbase * n_tiles;

  if (warp == W)
  {  // ---------------- producer warp: descriptor, both matrices and the tile's tip rows of every update
    const unsigned total_it = (unsigned)rounds * (unsigned)n_ops;
    for (unsigned base = 0; base < total_it; base += 32)
    {
      const unsigned     my = base + lane;
      unsigned long long m1 = 0, m2 = 0, r1 = 0, r2 = 0;
      int                kd = 0;
      if (my < total_it)
      {
        const OpDev *o = ops + (my % (unsigned)n_ops);
        m1 = (unsigned long long)o->P1;
        m2 = (unsigned long long)o->P2;
        r1 = (unsigned long long)o->t1;
        r2 = (unsigned long long)o->t2;
        kd = o->flags;
      }
      const int cnt = (int)min(32u, total_it - base);
      for (int j = 0; j < cnt; ++j)
      {
        const unsigned long long a1 = __shfl_sync(0xffffffffu, m1, j), a2 = __shfl_sync(0xffffffffu, m2, j);
        const unsigned long long q1 = __shfl_sync(0xffffffffu, r1, j), q2 = __shfl_sync(0xffffffffu, r2, j);
        const int                kind = __shfl_sync(0xffffffffu, kd, j);
        if (lane == 0)
        {
          const unsigned it = base + j;
          const int      s = (int)(it & (S - 1));
          const uint32_t ph = (it / S) & 1u;
          const int      tile = (int)blockIdx.x + (int)(it / (unsigned)n_ops) * (int)gridDim.x;
          const int      chunk0 = tile * tbase + min(tile, trem);
          const uint32_t rb = (uint32_t)((tbase + (tile < trem ? 1 : 0)) * 16);  // one 16-byte unit per chunk
          const bool     tipA = (kind & 3) == kSrcTip, tipB = (kind >> 2) == kSrcTip;
          mbar_wait_backoff(&empty[s], ph ^ 1u);
          const uint32_t b1 = tipA ? TB : PB;
          const uint32_t b2 = tipB ? TB : PB;
          mbar_expect_tx(&full[s], (uint32_t)sizeof(OpDev) + b1 + b2 + (tipA ? rb : 0u) + (tipB ? rb : 0u));
          tma_bulk_g2s(&st[s].op, ops + (it % (unsigned)n_ops), (uint32_t)sizeof(OpDev), &full[s]);
          tma_bulk_g2s(st[s].M[0], (const void *)a1, b1, &full[s]);
          tma_bulk_g2s(st[s].M[1], (const void *)a2, b2, &full[s]);
          if (tipA) tma_bulk_g2s(st[s].rows[0], (const void *)(q1 + (unsigned long long)chunk0 * 16ull), rb, &full[s]);
          if (tipB) tma_bulk_g2s(st[s].rows[1], (const void *)(q2 + (unsigned long long)chunk0 * 16ull), rb, &full[s]);
        }
        __syncwarp();
      }
    }
  }
  else
  {

  // ---------------- compute warps
  const int      g = lane >> 2, t = lane & 3, os = (t >> 1) * 8 + g;
  const uint32_t st_a = smem_u32(st), fwd_a = smem_u32(fwd), done_a = smem_u32(done), live_a = smem_u32(livef) + os;
  const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
  // B fragments: lane (g, t) holds P[c][g & 3][t] in B_lo if g < 4 and in B_hi otherwise; the other one is read
  // from a zero word of the stage (no select instructions)
  const uint32_t bsel = (uint32_t)(((g & 3) * 4 + t) * 8);
  const uint32_t lo_off = (g < 4) ? (kT4OffM + bsel) : kT4OffZero, hi_off = (g < 4) ? kT4OffZero : (kT4OffM + bsel);
  const uint32_t lo_step = (g < 4) ? 128u : 0u, hi_step = (g < 4) ? 0u : 128u;  // per category
  const uint32_t lo_b = (g < 4) ? (uint32_t)(NCATG * 64 * 8) : 0u, hi_b = (g < 4) ? 0u : (uint32_t)(NCATG * 64 * 8);  // child B
  for (int r = 0; r < rounds; ++r)
  {
    const int tile = (int)blockIdx.x + r * (int)gridDim.x;
    const int chunk0 = tile * tbase + min(tile, trem);
    const int C = tbase + (tile < trem ? 1 : 0);
    // per-tile state: progress counters and live flags (avx.c:399: zero-weight patterns are never stored; the
    // arrays are padded with zero weights)
    if (r > 0) asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");
    for (int q = tid; q < C * 16; q += W * 32) livef[q] = (wght[chunk0 * 16 + q] > DBL_MIN) ? 1 : 0;
    for (int q = tid; q < C; q += W * 32) done[q] = 0;
    asm volatile("bar.sync 1, %0;" ::"n"(W * 32) : "memory");

    const unsigned it0 = (unsigned)r * (unsigned)n_ops;
    int            cur = -1;  // update whose ring stage this warp holds
    int            k = 0, i = warp;
    while (i >= C)
    {
      i -= C;
      ++k;
    }
    double bAlo[NCATG], bAhi[NCATG], bBlo[NCATG], bBhi[NCATG];
#pragma unroll
    for (int c = 0; c < NCATG; ++c) bAlo[c] = bAhi[c] = bBlo[c] = bBhi[c] = 0.0;
    int      kind = 0;
    uint32_t sa = st_a;
    while (true)
    {
      const int kk = min(k, n_ops - 1);
      // walk the ring up to update kk: every warp waits for and releases EVERY update (also those in which it
      // has no item), so that no warp can run more than the ring depth ahead of another
      while (cur < kk)
      {
        if (cur >= 0)
        {
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a + ((it0 + (unsigned)cur) & (S - 1)) * 8)
                         : "memory");
        }
        ++cur;
        const unsigned it = it0 + (unsigned)cur;
        const uint32_t fa = full_a + (it & (S - 1)) * 8, par = (it / S) & 1u;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "T4_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra T4_DONE;\n\t"
            "bra T4_WAIT;\n\t"
            "T4_DONE:\n\t"
            "}" ::"r"(fa),
            "r"(par)
            : "memory");
      }
      if (k >= n_ops) break;
      sa = st_a + ((it0 + (unsigned)k) & (S - 1)) * kStageB;
      if ((kind >> 8) != k + 1)
      {  // first item of this warp in update k: operand kinds and B fragments (`kind` = flags | (k + 1) << 8)
        kind = lds32(sa + kT4OffFlags) | ((k + 1) << 8);
#pragma unroll
        for (int c = 0; c < NCATG; ++c)
        {
          bAlo[c] = lds64(sa + lo_off + c * lo_step);
          bAhi[c] = lds64(sa + hi_off + c * hi_step);
          bBlo[c] = lds64(sa + lo_b + lo_off + c * lo_step);
          bBhi[c] = lds64(sa + hi_b + hi_off + c * hi_step);
        }
      }
      // the previous update of this chunk (another warp's item) must be complete
      const uint32_t da = done_a + (uint32_t)i * 4;
      while (lds32_volatile(da) < k) {}
      fence_cta();
      const uint32_t fw = fwd_a + (uint32_t)i * t4_chunk_bytes<NCATG>();
      const bool     live = lds8(live_a + (uint32_t)(i * 16)) != 0;
      const uint32_t rw = sa + kT4OffM + 2 * NCATG * 64 * 8 - (uint32_t)(chunk0 * 16);
      const int      ka = kind & 3, kb = (kind >> 2) & 3;
#define T4_ITEM(KA, KB) t4_item<NCATG, KA, KB>(sa, bAlo, bAhi, bBlo, bBhi, fw, chunk0 + i, rw, live, lane, apply_scaling)
      if (ka == kSrcFwd)
      {
        if (kb == kSrcTip)
          T4_ITEM(kSrcFwd, kSrcTip);
        else
          T4_ITEM(kSrcFwd, kSrcSlot);
      }
      else if (ka == kSrcTip)
        T4_ITEM(kSrcTip, kSrcTip);
      else if (kb == kSrcTip)
        T4_ITEM(kSrcSlot, kSrcTip);
      else
        T4_ITEM(kSrcSlot, kSrcSlot);
#undef T4_ITEM
      __syncwarp();
      if (lane == 0)
      {
        fence_cta();
        asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(da), "r"(k + 1) : "memory");
      }
      i += W;
      while (i >= C)
      {
        i -= C;
        ++k;
      }
      if (PF && k < n_ops)
      {  // pull the global operands of this warp's NEXT item into L1 (one 2 KB chunk per operand: 16 lines).
         // Its descriptor is read if its ring stage is already full (the producer runs ahead); never waits.
        const unsigned itn = it0 + (unsigned)k;
        const uint32_t san = st_a + (itn & (S - 1)) * kStageB;
        bool           ready = (san == sa) && ((kind >> 8) == k + 1);
        if (!ready)
        {
          uint32_t ok;
          asm volatile(
              "{\n\t"
              ".reg .pred p;\n\t"
              "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
              "selp.u32 %0, 1, 0, p;\n\t"
              "}"
              : "=r"(ok)
              : "r"(full_a + (itn & (S - 1)) * 8), "r"((itn / S) & 1u)
              : "memory");
          ready = ok != 0;
        }
        if (ready && lane < 16)
        {
          const int    fl = lds32(san + kT4OffFlags);
          const size_t po = (size_t)(chunk0 + i) * (2 * NCATG * 32) + lane * 16;
          if ((fl & 3) == kSrcSlot) prefetch_l1(lds_ptr(san + kT4OffC1) + po);
          if (((fl >> 2) & 3) == kSrcSlot) prefetch_l1(lds_ptr(san + kT4OffC2) + po);
        }
      }
    }
    // release the last update of the round
    __syncwarp();
    if (lane == 0)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_a + ((it0 + (unsigned)(n_ops - 1)) & (S - 1)) * 8)
                   : "memory");
  }
  }  // compute warps

  // ---------------- fused epilogue: the site loop of Lk() at one edge over this block's tiles + the
  // deterministic grid reduction (lk.c:605-645).  Every CLV this block reads here was written by its own
  // warps: the block barrier orders those stores.
  if (edge.enabled)
  {
    if constexpr (NCATG == 4)
    {
      __syncthreads();
      EdgeLaneConst<NCATG> ek;
      ek.load(edge, lane);
      double acc[1] = {0.0};
      int    warn = 0;
      for (int r = 0; r < rounds; ++r)
      {
        const int tile = (int)blockIdx.x + r * (int)gridDim.x;
        const int chunk0 = tile * tbase + min(tile, trem);
        const int C = tbase + (tile < trem ? 1 : 0);
        for (int grp = chunk0 * 2 + warp; grp < (chunk0 + C) * 2; grp += W + 1)  // groups of 8 sites
          edge_lnl_dna_group<NCATG>(edge, ek, grp, lane, acc[0], warn);
      }
      block_reduce_finish<1>(acc, warn, edge.ro);
    }
  }
}

// K4, 4 states: thread per (site, category); dot_prod is [site][catg][4] (plain layout).
template <int NCATG>
__global__ void __launch_bounds__(512)
    k_lnl_dlnl_dna(const double *__restrict__ dot_prod, const int *__restrict__ fact_sum_scale,
                   const ModelDev *__restrict__ mod, double l, int with_derivative, int npat,
                   const double *__restrict__ wght, const short *__restrict__ invar, double *__restrict__ site_lnl,
                   ReduceOut ro)
{
  constexpr int SW = 32 / NCATG;
  const int     lane = threadIdx.x & 31, cat = lane / SW;
  const int     warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5;
  double        E[4], D[4];
  {
    double len, rr = mod->rates[cat];
    if (with_derivative)
    {  // lk.c:690-705
      rr = rr * mod->br_len_mult;
      len = l * rr;
    }
    else
    {  // lk.c:596-600
      len = fmax(0.0, l) * mod->rates[cat];
      len = len * mod->br_len_mult;
    }
    if (len < mod->l_min)
      len = mod->l_min;
    else if (len > mod->l_max)
      len = mod->l_max;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
      const double e = exp(mod->lambda[i] * len);
      E[i] = e;
      D[i] = e * mod->lambda[i] * rr;
    }
  }
  const double wc = mod->probs[cat];
  double       acc[2] = {0.0, 0.0};
  int          warn = 0;
  const int    groups = (npat + SW - 1) / SW;
  for (int grp = blockIdx.x * warps_per_block + warp; grp < groups; grp += gridDim.x * warps_per_block)
  {
    const int    site0 = grp * SW + (lane % SW);
    const bool   valid = site0 < npat;
    const int    site = valid ? site0 : npat - 1;
    const double w = wght[site];
    const bool   live = valid && (w > DBL_MIN);
    const double4a dp = ldg256(dot_prod + ((size_t)site * NCATG + cat) * 4);
    double         cl, cd = 0.0;
    if (with_derivative)
    {  // avx.c:250-276
      double le = fma(dp.x, E[0], 0.0), lo = fma(dp.y, E[1], 0.0);
      double de = fma(dp.x, D[0], 0.0), dd = fma(dp.y, D[1], 0.0);
      le = fma(dp.z, E[2], le);
      lo = fma(dp.w, E[3], lo);
      de = fma(dp.z, D[2], de);
      dd = fma(dp.w, D[3], dd);
      cl = le + lo;
      cd = de + dd;
    }
    else
    {  // avx.c:220-245
      cl = hsum4(0.0 + dp.x * E[0], 0.0 + dp.y * E[1], 0.0 + dp.z * E[2], 0.0 + dp.w * E[3]);
    }
    const double tl = cl * wc, td = cd * wc;
    double       lk = 0.0, dlk = 0.0;
#pragma unroll
    for (int c = 0; c < NCATG; ++c)
    {
      lk = lk + __shfl_sync(0xffffffffu, tl, c * SW + (lane % SW));
      dlk = dlk + __shfl_sync(0xffffffffu, td, c * SW + (lane % SW));
    }
    if (cat == 0 && live)
    {
      int fact = fact_sum_scale[site];
      if (mod->invar_flag)
      {
        bool   ovf;
        double inv = invariant_lk(fact, invar[site], mod->pi, &ovf);
        if (with_derivative)
        {  // lk.c:1005-1025
          if (ovf)
          {
            lk = inv * mod->pinv;
            dlk = 0.0;
          }
          else
          {
            lk = lk * (1. - mod->pinv) + inv * mod->pinv;
            dlk = dlk * (1. - mod->pinv);
          }
        }
        else
        {  // lk.c:910-931
          if (ovf)
          {
            fact = 0;
            inv = invariant_lk(0, invar[site], mod->pi, &ovf);
            lk = inv * mod->pinv;
          }
          else
            lk = lk * (1. - mod->pinv) + inv * mod->pinv;
        }
      }
      if (lk < DBL_MIN)
      {
        lk = DBL_MIN;
        warn = 1;
      }
      const double lsl = log(lk) - kLog2 * fact;
      if (!with_derivative) site_lnl[site] = lsl;
      acc[0] += w * lsl;
      acc[1] += w * (dlk / lk);
    }
  }
  block_reduce_finish<2>(acc, warn, ro);
}

// ------------------------------------------------------------------------------------------------
// K3: thread per (site, category).
__global__ void __launch_bounds__(128)
    k_eigen_lr(SideDev left, SideDev rght, const ModelDev *__restrict__ mod, int npat, int ns, int ncatg,
               const double *__restrict__ wght, const uint32_t *__restrict__ tipmask, double *__restrict__ dot_prod,
               int *__restrict__ fact_sum_scale, int blocked)
{
  const int       ncns = ncatg * ns;
  const long long total = (long long)npat * ncatg;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x)
  {
    const int site = (int)(g / ncatg), c = (int)(g % ncatg);
    if (c == 0) fact_sum_scale[site] = (left.scale ? left.scale[site] : 0) + (rght.scale ? rght.scale[site] : 0);
    if (!(wght[site] > DBL_MIN)) continue;  // lk.c:1082
    const uint32_t lm = left.clv ? 0u : tipmask[left.tip[site]];
    const uint32_t rm = rght.clv ? 0u : tipmask[rght.tip[site]];
    auto           LV = [&](int j) -> double {
      return left.clv ? left.clv[clv_off(site, c, j, ncatg, ns, blocked)] : (double)((lm >> j) & 1u);
    };
    auto RV = [&](int j) -> double {
      return rght.clv ? rght.clv[clv_off(site, c, j, ncatg, ns, blocked)] : (double)((rm >> j) & 1u);
    };
    double *o = dot_prod + (size_t)site * ncns + c * ns;
    for (int i = 0; i < ns; ++i)
    {  // avx.c:79-84: left_i = sum_j U[j][i] (L_j pi_j), rght_i = sum_j V[i][j] R_j, first term a product
      double a = mod->U[i] * (LV(0) * mod->pi[0]);
      double b = mod->V[i * ns] * RV(0);
      for (int j = 1; j < ns; ++j)
      {
        a = fma(mod->U[j * ns + i], LV(j) * mod->pi[j], a);
        b = fma(mod->V[i * ns + j], RV(j), b);
      }
      o[i] = a * b;
    }
  }
}

// K3 with the state count known at compile time (20 states): thread per (site, category) as above, but both
// conditional vectors live in registers (read from global memory once instead of once per output state) and U
// (transposed, so that the j-loop walks contiguous words) and V come from shared memory as warp-wide broadcasts.
// Same products and the same FMA chains as k_eigen_lr, i.e. bit-identical dot_prod.
template <int NS>
__global__ void __launch_bounds__(128)
    k_eigen_lr_reg(SideDev left, SideDev rght, const ModelDev *__restrict__ mod, int npat, int ncatg,
                   const double *__restrict__ wght, const uint32_t *__restrict__ tipmask, double *__restrict__ dot_prod,
                   int *__restrict__ fact_sum_scale, int blocked)
{
  __shared__ double sUt[NS * NS], sV[NS * NS], sPi[NS];
  for (int t = threadIdx.x; t < NS * NS; t += blockDim.x)
  {
    sUt[(t % NS) * NS + (t / NS)] = mod->U[t];  // sUt[i][j] = U[j][i]
    sV[t] = mod->V[t];
  }
  if (threadIdx.x < NS) sPi[threadIdx.x] = mod->pi[threadIdx.x];
  __syncthreads();
  const int       ncns = ncatg * NS;
  const long long total = (long long)npat * ncatg;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x)
  {
    const int site = (int)(g / ncatg), c = (int)(g % ncatg);
    if (c == 0) fact_sum_scale[site] = (left.scale ? left.scale[site] : 0) + (rght.scale ? rght.scale[site] : 0);
    if (!(wght[site] > DBL_MIN)) continue;  // lk.c:1082
    double lv[NS], rv[NS];
    if (left.clv)
    {
#pragma unroll
      for (int j = 0; j < NS; ++j) lv[j] = left.clv[clv_off(site, c, j, ncatg, NS, blocked)] * sPi[j];
    }
    else
    {
      const uint32_t lm = tipmask[left.tip[site]];
#pragma unroll
      for (int j = 0; j < NS; ++j) lv[j] = (double)((lm >> j) & 1u) * sPi[j];
    }
    if (rght.clv)
    {
#pragma unroll
      for (int j = 0; j < NS; ++j) rv[j] = rght.clv[clv_off(site, c, j, ncatg, NS, blocked)];
    }
    else
    {
      const uint32_t rm = tipmask[rght.tip[site]];
#pragma unroll
      for (int j = 0; j < NS; ++j) rv[j] = (double)((rm >> j) & 1u);
    }
    double *o = dot_prod + (size_t)site * ncns + c * NS;
#pragma unroll 2
    for (int i = 0; i < NS; ++i)
    {  // avx.c:79-84
      double a = sUt[i * NS] * lv[0];
      double b = sV[i * NS] * rv[0];
#pragma unroll
      for (int j = 1; j < NS; ++j)
      {
        a = fma(sUt[i * NS + j], lv[j], a);
        b = fma(sV[i * NS + j], rv[j], b);
      }
      o[i] = a * b;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4: thread per site.  expl = (E,D) interleaved per category like tree->expl (lk.c:717-725).
__global__ void __launch_bounds__(128)
    k_lnl_dlnl(const double *__restrict__ dot_prod, const int *__restrict__ fact_sum_scale,
               const ModelDev *__restrict__ mod, double l, int with_derivative, int npat, int ns, int ncatg,
               const double *__restrict__ wght, const short *__restrict__ invar, double *__restrict__ site_lnl,
               ReduceOut ro)
{
  __shared__ double sE[kMaxCatg * kMaxNs], sD[kMaxCatg * kMaxNs];
  const int         ncns = ncatg * ns;
  for (int t = threadIdx.x; t < ncns; t += blockDim.x)
  {
    const int c = t / ns, i = t % ns;
    double    len, rr = mod->rates[c];
    if (with_derivative)
    {  // lk.c:690-705
      rr = rr * mod->br_len_mult;
      len = l * rr;
    }
    else
    {  // lk.c:596-600
      len = fmax(0.0, l) * mod->rates[c];
      len = len * mod->br_len_mult;
    }
    if (len < mod->l_min)
      len = mod->l_min;
    else if (len > mod->l_max)
      len = mod->l_max;
    const double e = exp(mod->lambda[i] * len);
    sE[t] = e;
    sD[t] = e * mod->lambda[i] * rr;
  }
  __syncthreads();

  double acc[2] = {0.0, 0.0};
  int    warn = 0;
  for (int site = blockIdx.x * blockDim.x + threadIdx.x; site < npat; site += gridDim.x * blockDim.x)
  {
    const double w = wght[site];
    if (!(w > DBL_MIN)) continue;
    const double *dp = dot_prod + (size_t)site * ncns;
    int           fact = fact_sum_scale[site];
    double        lk = 0.0, dlk = 0.0;
    for (int c = 0; c < ncatg; ++c)
    {
      double cl, cd = 0.0;
      if ((ns & 3) == 0)
      {
        if (with_derivative)
        {  // avx.c:250-276: even/odd states accumulate separately with FMAs, then add
          double le = 0.0, lo = 0.0, de = 0.0, dd = 0.0;
          for (int i = 0; i < ns; i += 2)
          {
            le = fma(dp[c * ns + i], sE[c * ns + i], le);
            de = fma(dp[c * ns + i], sD[c * ns + i], de);
            lo = fma(dp[c * ns + i + 1], sE[c * ns + i + 1], lo);
            dd = fma(dp[c * ns + i + 1], sD[c * ns + i + 1], dd);
          }
          cl = le + lo;
          cd = de + dd;
        }
        else
        {  // avx.c:220-245
          double q[4] = {0.0, 0.0, 0.0, 0.0};
          for (int b = 0; b < ns; b += 4)
#pragma unroll
            for (int t = 0; t < 4; ++t) q[t] = q[t] + dp[c * ns + b + t] * sE[c * ns + b + t];
          cl = hsum4(q[0], q[1], q[2], q[3]);
        }
      }
      else
      {  // lk.c:1157-1180
        cl = 0.0;
        for (int i = 0; i < ns; ++i)
        {
          cl = cl + dp[c * ns + i] * sE[c * ns + i];
          cd = cd + dp[c * ns + i] * sD[c * ns + i];
        }
      }
      lk = lk + cl * mod->probs[c];
      dlk = dlk + cd * mod->probs[c];
    }
    if (mod->invar_flag)
    {
      bool   ovf;
      double inv = invariant_lk(fact, invar[site], mod->pi, &ovf);
      if (with_derivative)
      {  // lk.c:1005-1025
        if (ovf)
        {
          lk = inv * mod->pinv;
          dlk = 0.0;
        }
        else
        {
          lk = lk * (1. - mod->pinv) + inv * mod->pinv;
          dlk = dlk * (1. - mod->pinv);
        }
      }
      else
      {  // lk.c:910-931
        if (ovf)
        {
          fact = 0;
          inv = invariant_lk(0, invar[site], mod->pi, &ovf);
          lk = inv * mod->pinv;
        }
        else
          lk = lk * (1. - mod->pinv) + inv * mod->pinv;
      }
    }
    if (lk < DBL_MIN)
    {
      lk = DBL_MIN;
      warn = 1;
    }
    const double lsl = log(lk) - kLog2 * fact;
    if (!with_derivative) site_lnl[site] = lsl;  // Lk_Core_Eigen_Lr writes c_lnL_sorted (lk.c:945); dLk does not
    acc[0] += w * lsl;
    acc[1] += w * (dlk / lk);
  }
  block_reduce_finish<2>(acc, warn, ro);
}

}  // namespace plk
