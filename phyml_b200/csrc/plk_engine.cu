// phyml_b200/csrc/plk_engine.cu -- instance management, launch scheduling and the C ABI
// (include/phyml_b200.h) of the B200 likelihood engine.  No CPU fallback: every entry point runs
// CUDA kernels on the instance's stream or fails with PLK_ERR_CUDA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/phyml_b200.h"
#include "plk_kernels.cuh"
#include "plk_pars.cuh"
#include "plk_spr.cuh"

using namespace plk;

namespace
{
std::string g_create_error;

struct NcclApi
{
  void *handle = nullptr;
  struct UniqueId
  {
    char internal[128];
  };
  int (*GetUniqueId)(UniqueId *) = nullptr;
  int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load(std::string &err)
  {
    if (handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
    {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle)
    {
      err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy)
    {
      err = "libnccl is missing required symbols";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;

constexpr int    kStageSlots = 8;
constexpr size_t kStageBytes = 256 * 1024;
constexpr int    kMaxReduceBlocks = 4096;
constexpr size_t kSitePad = 32;  // per-site arrays are padded so that the last 32-item chunk can be read unpredicated
}  // namespace

struct plk_instance
{
  plk_config   cfg{};
  cudaStream_t stream = nullptr;
  std::string  err;
  int          apply_scaling = 1;
  int          num_sms = 148;

  double   *d_wght = nullptr;
  short    *d_invar = nullptr;
  uint32_t *d_tipmask = nullptr;
  uint8_t  *d_tipcodes = nullptr;
  uint8_t  *d_tippacked = nullptr;  // staging of plk_set_all_tip_codes_packed4
  uint8_t  *d_tiprows = nullptr;   // ns == 4: codes translated to tip-table rows for the fused kernel
  bool      tiprows_dirty = true;
  size_t    tip_stride = 0;
  ModelDev *d_model = nullptr;
  double   *d_pmat = nullptr;
  size_t    pmat_elems = 0;   // ncatg*ns*ns doubles of P
  size_t    pmat_stride = 0;  // per-handle record: P followed (ns == 4) by the tip table TP[cat][16][4]

  std::vector<double *> clv;
  std::vector<int *>    scale;

  double *d_site_lnl = nullptr, *d_site_lk = nullptr, *d_site_lk_cat = nullptr;
  int    *d_fact = nullptr;
  double *d_dot_prod = nullptr;
  double *d_partials = nullptr;
  int    *d_warn = nullptr;
  unsigned int *d_ticket = nullptr;
  double *d_result = nullptr;

  ResultHost        *h_result = nullptr;
  ResultHost        *h_result_dev = nullptr;
  unsigned long long seq = 0;
  unsigned long long coll_seq = 0;  // cross-GPU exchange counter: reset whenever the mailboxes are (re)wired

  // pinned staging ring for small descriptor uploads
  char       *h_stage[kStageSlots] = {};
  cudaEvent_t stage_ev[kStageSlots] = {};
  int         stage_next = 0;
  char       *d_stage = nullptr;  // device mirror, one region per slot

  std::vector<double>                   tip_table;  // [n_codes][ns]
  std::vector<uint32_t>                 masks;
  std::unordered_map<uint32_t, int>     mask_to_code;
  bool                                  eigen_ready = false;
  bool                                  site_valid = false;

  long long launches = 0;
  size_t    bytes = 0;

  void *comm = nullptr;
  bool  allreduce = false;
  // fused P2P exchange (CUDA IPC mailboxes over NVLink)
  P2pSlot  *d_mbox = nullptr;        // this rank's mailbox [2][world]
  P2pSlot **d_peers = nullptr;       // device array: every rank's mailbox as mapped here
  std::vector<void *> ipc_opened;
  bool      p2p = false;
  int   rank = 0, world = 1;
  double l_min = 1e-8, l_max = 100.0;  // host copy of mod->l_min / l_max (plk_set_model)
  int    trav_umax = 2;                // items per thread of the fused traversal kernel
  int    trav_blocks_per_sm = 2;
  int    trav_v1 = 0;            // PLK_TRAV_V1=1: first-generation kernel (k_traverse_dna), kept for A/B measurements
  // Post_Order_Lk + edge reduction in one launch (plk_traverse_edge_lnl): request consumed by launch_traverse4_t
  bool     pending_edge = false, edge_fused = false;
  plk_side edge_left{}, edge_rght{};
  int      edge_pmat = 0;
  int    t2_variant = 20;        // PLK_T2_VARIANT: < 10: k_traverse_dna2, 10..19: k_traverse_dna3, >= 20: k_traverse_dna4 (default)
  bool   aa_attr_set = false;
  int    aa_v1 = 0;              // PLK_AA_V1=1: first-generation 20-state kernel (k_traverse_aa; also used for ncatg 3, 5, 6, 7)
  int    blocked = 0;            // CLVs in the blocked layout (ns = 4, 20), see clv_off()
  int    dna_mma = 0;            // use k_traverse_dna_mma (tensor-pipe variant) for ns = 4
  int    mma_u = 2;
  double *d_tmp_clv = nullptr;   // plain-layout staging for plk_get_clv / plk_set_clv

  // single-process site sharding (plk_create_sharded): this object is then only a dispatcher over one
  // ordinary instance per device, each owning the contiguous pattern block [shard_lo[i], shard_lo[i+1])
  std::vector<plk_instance *> shards;
  std::vector<int>            shard_lo;
  bool                        inproc_p2p = false;

  // parsimony (plk_pars_*, src/pars.c): one Fitch buffer int2{ui, pars}[P] per handle and, once a step matrix
  // has been given, one state-major step-matrix buffer int[ns][pars_pstride] per handle; storage is lazy
  int                 pars_n = 0;
  std::vector<int2 *> pars_fitch;
  std::vector<int *>  pars_sank;
  size_t              pars_pstride = 0;
  int                *d_step_mat = nullptr;
  int                *d_site_pars = nullptr;
  bool                pars_site_valid = false;
  bool                wght_integral = true;  // every pattern weight is an integer (plk_set_pattern_weights)

  // host-to-device copies of the tip codes run on their own stream so that they overlap compute that was enqueued
  // earlier (plk_lk_full_begin ... upload of the next inputs ... plk_lk_wait)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t  ev_copy = nullptr, ev_unpack = nullptr;

  // batched SPR candidates (plk_spr_candidates): scratch P-matrices, per-block partial sums, results
  double *d_spr_pmat = nullptr, *d_spr_partials = nullptr, *d_spr_lnl = nullptr;
  int    *d_spr_warn = nullptr;
  int     spr_cap = 0;
  bool    spr_attr_set = false;

  // scheduling scratch
  std::vector<int> lvl_write, lvl_read, op_level;
  bool fused_dna = false, fused_aa = false;
  // descriptor cache: a repeated identical op list (Lk(NULL) on an unchanged topology) is launched
  // straight from the resolved descriptors of the previous call
  std::vector<plk_op> cache_ops;
  OpDev             *d_ops_cache = nullptr;
  unsigned long long alloc_epoch = 0, cache_epoch = ~0ull;
};

#define CU_TRY(inst, call)                                                                         \
  do                                                                                               \
  {                                                                                                \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
    {                                                                                              \
      (inst)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      (void)cudaGetLastError(); /* a non-sticky error must not surface again in a later call */    \
      return PLK_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define ARG_CHECK(inst, cond, msg) \
  do                               \
  {                                \
    if (!(cond))                   \
    {                              \
      (inst)->err = (msg);         \
      return PLK_ERR_ARG;          \
    }                              \
  } while (0)

namespace
{
bool g_multi_device = false;  // set once a sharded instance exists: every entry point then selects its device

inline int use_device(plk_instance *inst)
{
  if (g_multi_device) CU_TRY(inst, cudaSetDevice(inst->cfg.device));
  return PLK_OK;
}
#define USE_DEVICE(inst)          \
  do                              \
  {                               \
    int rc_ = use_device(inst);   \
    if (rc_) return rc_;          \
  } while (0)

// run `call` (an expression using `sh`, the shard, and `lo`, the first pattern of its block) on every shard
#define FOR_SHARDS(inst, call)                                                   \
  do                                                                             \
  {                                                                              \
    for (size_t i_ = 0; i_ < (inst)->shards.size(); ++i_)                        \
    {                                                                            \
      plk_instance *sh = (inst)->shards[i_];                                     \
      const size_t  lo = (size_t)(inst)->shard_lo[i_];                           \
      (void)lo;                                                                  \
      int rc_ = (call);                                                          \
      if (rc_)                                                                   \
      {                                                                          \
        (inst)->err = "shard " + std::to_string(i_) + ": " + sh->err;            \
        return rc_;                                                              \
      }                                                                          \
    }                                                                            \
  } while (0)

template <typename T>
int dev_alloc(plk_instance *inst, T **p, size_t n)
{
  const size_t b = std::max<size_t>(n * sizeof(T), 256);
  CU_TRY(inst, cudaMalloc((void **)p, b));
  inst->bytes += b;
  return PLK_OK;
}

// elements of one CLV in the caller's (plain) layout / in device memory (blocked layout pads to 8 sites)
size_t clv_elems_plain(const plk_instance *inst)
{
  return (size_t)inst->cfg.n_patterns * inst->cfg.ncatg * inst->cfg.ns;
}
size_t clv_elems(const plk_instance *inst)
{
  // blocked layout: whole 32-site chunks (a warp of the fused kernels may read up to 31 sites past the last pattern)
  const size_t sites = inst->blocked ? (((size_t)inst->cfg.n_patterns + 31) & ~(size_t)31) : (size_t)inst->cfg.n_patterns;
  return sites * inst->cfg.ncatg * inst->cfg.ns;
}

int ensure_clv(plk_instance *inst, int h)
{
  if (inst->clv[h]) return PLK_OK;
  int rc = dev_alloc(inst, &inst->clv[h], clv_elems(inst) + 4);
  if (rc) return rc;
  inst->alloc_epoch++;
  rc = dev_alloc(inst, &inst->scale[h], (size_t)inst->cfg.n_patterns + kSitePad);
  if (rc) return rc;
  // zero-weight patterns are never written by K1 (avx.c:515-520): start from defined contents
  CU_TRY(inst, cudaMemsetAsync(inst->clv[h], 0, clv_elems(inst) * sizeof(double), inst->stream));
  CU_TRY(inst, cudaMemsetAsync(inst->scale[h], 0, ((size_t)inst->cfg.n_patterns + kSitePad) * sizeof(int), inst->stream));
  return PLK_OK;
}

// copy a small host block to the device through the pinned ring; returns the device address
int stage_upload(plk_instance *inst, const void *src, size_t bytes, void **dev_out)
{
  if (bytes > kStageBytes)
  {
    inst->err = "internal: staging block too large";
    return PLK_ERR_ARG;
  }
  const int s = inst->stage_next;
  inst->stage_next = (s + 1) % kStageSlots;
  CU_TRY(inst, cudaEventSynchronize(inst->stage_ev[s]));
  memcpy(inst->h_stage[s], src, bytes);
  char *dst = inst->d_stage + (size_t)s * kStageBytes;
  CU_TRY(inst, cudaMemcpyAsync(dst, inst->h_stage[s], bytes, cudaMemcpyHostToDevice, inst->stream));
  CU_TRY(inst, cudaEventRecord(inst->stage_ev[s], inst->stream));
  *dev_out = dst;
  return PLK_OK;
}

int check_side(plk_instance *inst, const plk_side &s, bool need_data)
{
  const bool tip = s.tip >= 0;
  const bool in = s.clv >= 0;
  ARG_CHECK(inst, tip != in, "operand must be exactly one of tip / clv");
  if (tip) ARG_CHECK(inst, s.tip < inst->cfg.n_tips, "tip index out of range");
  if (in)
  {
    ARG_CHECK(inst, s.clv < inst->cfg.n_clv, "clv handle out of range");
    if (need_data) ARG_CHECK(inst, inst->clv[s.clv] != nullptr, "clv handle read before it was ever written");
  }
  return PLK_OK;
}

SideDev side_dev(plk_instance *inst, const plk_side &s)
{
  SideDev d;
  if (s.tip >= 0)
  {
    d.clv = nullptr;
    d.scale = nullptr;
    d.tip = inst->d_tipcodes + (size_t)s.tip * inst->tip_stride;
  }
  else
  {
    d.clv = inst->clv[s.clv];
    d.scale = inst->scale[s.clv];
    d.tip = nullptr;
  }
  return d;
}

// blocks per SM of the reduction kernels (K2 / K4).  Every block ends with a fence + ticket atomic + barrier, and the
// last block adds one partial per block: on the latency path (one Lk(b) / dLk per host round trip) fewer, longer
// blocks finish sooner than one block per 128 work items.  Measured at 100 taxa x 50 000 sites (profiles/latency_r2.md):
// 2 blocks of 256 threads per SM: dLk 19.6 -> 18.1 us, Br_Len_Opt 572 -> 537 us against 16 (then 4) blocks of 128.
// PLK_REDUCE_BLOCKS_PER_SM / PLK_REDUCE_THREADS override them for A/B runs.
int reduce_blocks_per_sm()
{
  static const int v = [] {
    const char *e = getenv("PLK_REDUCE_BLOCKS_PER_SM");
    const int   k = e ? atoi(e) : 2;
    return k < 1 ? 1 : (k > 16 ? 16 : k);
  }();
  return v;
}

// threads per block of the two 4-state reduction kernels (thread per (pattern, category): a warp covers 8 patterns)
int reduce_threads_dna()
{
  static const int v = [] {
    const char *e = getenv("PLK_REDUCE_THREADS");
    const int   k = e ? atoi(e) : 256;
    return (k == 128 || k == 512) ? k : 256;
  }();
  return v;
}
int reduce_grid_dna(const plk_instance *inst)
{
  const int groups = (inst->cfg.n_patterns + 7) / 8, wpb = reduce_threads_dna() / 32;
  return std::max(1, std::min((groups + wpb - 1) / wpb, std::min(kMaxReduceBlocks, inst->num_sms * reduce_blocks_per_sm())));
}

// the generic (thread per pattern, 128-thread blocks) reduction kernels: 4 blocks per SM (2 is slower at 20 states)
int reduce_grid(const plk_instance *inst, int threads)
{
  static const int bps = getenv("PLK_REDUCE_BLOCKS_PER_SM") ? reduce_blocks_per_sm() : 4;
  const int        need = (inst->cfg.n_patterns + threads - 1) / threads;
  return std::max(1, std::min(need, std::min(kMaxReduceBlocks, inst->num_sms * bps)));
}

// where the reduction kernel about to be launched delivers its result
ReduceOut make_reduce_out(plk_instance *inst)
{
  ReduceOut ro;
  ro.partials = inst->d_partials;
  ro.ticket = inst->d_ticket;
  ro.warn_flag = inst->d_warn;
  ro.dev_out = inst->d_result;
  ro.host_out = inst->h_result_dev;
  ro.seq = ++inst->seq;
  ro.coll_seq = inst->p2p ? ++inst->coll_seq : 0;
  ro.publish = inst->allreduce ? 0 : 1;
  ro.peers = inst->p2p ? inst->d_peers : nullptr;
  ro.rank = inst->rank;
  ro.world = inst->p2p ? inst->world : 1;
  return ro;
}

// finish a reduction: optional all-reduce + publish, then wait for the mapped result
int finish_reduction(plk_instance *inst, double *out0, double *out1, int *warn)
{
  USE_DEVICE(inst);
  const unsigned long long seq = inst->seq;
  if (inst->allreduce)
  {
    // the warning flag travels as a third double so that one sum all-reduce carries everything
    const int rc = g_nccl.AllReduce(inst->d_result, inst->d_result, 3, /*ncclDouble*/ 8, /*ncclSum*/ 0, inst->comm,
                                    inst->stream);
    if (rc != 0)
    {
      inst->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
      return PLK_ERR_NCCL;
    }
    k_publish<<<1, 1, 0, inst->stream>>>(inst->d_result, inst->h_result_dev, seq);
    inst->launches++;
    CU_TRY(inst, cudaGetLastError());
  }
  // spin on the mapped result; fall back to the stream status to catch launch failures
  volatile ResultHost *h = inst->h_result;
  unsigned long long   spins = 0;
  while (h->seq != seq)
  {
    if ((++spins & 0x3fff) == 0)
    {
      cudaError_t q = cudaStreamQuery(inst->stream);
      if (q == cudaSuccess)
      {
        if (h->seq == seq) break;
        CU_TRY(inst, cudaStreamSynchronize(inst->stream));
        if (h->seq != seq)
        {
          inst->err = "reduction result was not published";
          return PLK_ERR_CUDA;
        }
        break;
      }
      if (q != cudaErrorNotReady) CU_TRY(inst, q);
    }
  }
  if (h->warn & kWarnPeerTimeout)
  {
    inst->err = "cross-GPU exchange timed out: a peer rank did not post its partial sum";
    return PLK_ERR_STATE;
  }
  if (out0) *out0 = h->val[0];
  if (out1) *out1 = h->val[1];
  if (warn) *warn = h->warn & 1;
  return PLK_OK;
}
}  // namespace

// =================================================================================================
extern "C" {

const char *plk_version(void) { return "phyml_b200 0.1 (sm_100a)"; }

const char *plk_last_error(const plk_instance *inst) { return inst ? inst->err.c_str() : g_create_error.c_str(); }

int plk_create(const plk_config *cfg, plk_instance **out)
{
  if (!cfg || !out)
  {
    g_create_error = "null argument";
    return PLK_ERR_ARG;
  }
  *out = nullptr;
  if (cfg->n_tips < 1 || cfg->n_patterns < 1 || cfg->ns < 2 || cfg->ns > kMaxNs || cfg->ncatg < 1 ||
      cfg->ncatg > kMaxCatg || cfg->n_clv < 1 || cfg->n_pmat < 1)
  {
    g_create_error = "plk_create: sizes out of range (2 <= ns <= 32, 1 <= ncatg <= 16)";
    return PLK_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
  {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this engine has no CPU fallback)";
    return PLK_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev)
  {
    g_create_error = "plk_create: device ordinal out of range";
    return PLK_ERR_ARG;
  }
  plk_instance *inst = new plk_instance();
  inst->cfg = *cfg;
  inst->apply_scaling = (cfg->flags & PLK_FLAG_NO_SCALING) ? 0 : 1;
  {  // the blocked layout goes with the fused kernels; the generic kernel keeps the reference's layout
    const int nc = cfg->ncatg;
    inst->fused_dna = (cfg->ns == 4) && (nc == 1 || nc == 2 || nc == 4 || nc == 8);
    inst->fused_aa = (cfg->ns == 20) && nc <= 8 && !getenv("PLK_AA_GENERIC");
    inst->blocked = (inst->fused_dna || inst->fused_aa) ? 1 : 0;
  }
  auto fail = [&](int rc) {
    g_create_error = inst->err;
    plk_destroy(inst);
    return rc;
  };
#define CREATE_TRY(call)                                             \
  do                                                                 \
  {                                                                  \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess)                                           \
    {                                                                \
      inst->err = std::string(#call) + ": " + cudaGetErrorString(e_); \
      return fail(PLK_ERR_CUDA);                                     \
    }                                                                \
  } while (0)
#define CREATE_RC(call)        \
  do                           \
  {                            \
    int rc_ = (call);          \
    if (rc_) return fail(rc_); \
  } while (0)

  CREATE_TRY(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  inst->num_sms = prop.multiProcessorCount;
  if (const char *e = getenv("PLK_TRAV_UMAX")) inst->trav_umax = (atoi(e) == 1) ? 1 : 2;
  if (const char *e = getenv("PLK_DNA_MMA")) inst->dna_mma = atoi(e) != 0;
  if (const char *e = getenv("PLK_TRAV_V1")) inst->trav_v1 = atoi(e) != 0;
  if (const char *e = getenv("PLK_AA_V1")) inst->aa_v1 = atoi(e) != 0;
  if (const char *e = getenv("PLK_T2_VARIANT")) inst->t2_variant = atoi(e);
  if (const char *e = getenv("PLK_TRAV_BLOCKS_PER_SM")) inst->trav_blocks_per_sm = std::max(1, std::min(4, atoi(e)));
  CREATE_TRY(cudaStreamCreateWithFlags(&inst->stream, cudaStreamNonBlocking));

  const size_t P = (size_t)cfg->n_patterns;
  // rows padded to whole 128-byte lines (+ one 32-site chunk): the fused kernels stage them with 16-byte TMA units
  inst->tip_stride = (P + kSitePad + 127) & ~(size_t)127;
  inst->pmat_elems = (size_t)cfg->ncatg * cfg->ns * cfg->ns;
  inst->pmat_stride = inst->pmat_elems + (cfg->ns == 4 ? (size_t)cfg->ncatg * 64 : 0) +
                      (cfg->ns == 20 ? (size_t)cfg->ncatg * (420 + 480) : 0);
  inst->clv.assign(cfg->n_clv, nullptr);
  inst->scale.assign(cfg->n_clv, nullptr);

  CREATE_RC(dev_alloc(inst, &inst->d_wght, P + kSitePad));  // padding keeps weight 0: never stored by K1
  CREATE_RC(dev_alloc(inst, &inst->d_invar, P));
  CREATE_RC(dev_alloc(inst, &inst->d_tipmask, 256));
  CREATE_RC(dev_alloc(inst, &inst->d_tipcodes, inst->tip_stride * cfg->n_tips + kSitePad));
  if (cfg->ns == 4 || cfg->ns == 20)
  {
    CREATE_RC(dev_alloc(inst, &inst->d_tiprows, inst->tip_stride * cfg->n_tips + kSitePad));
    CREATE_TRY(cudaMemsetAsync(inst->d_tiprows, 0, inst->tip_stride * cfg->n_tips + kSitePad, inst->stream));
  }
  CREATE_RC(dev_alloc(inst, &inst->d_model, 1));
  CREATE_RC(dev_alloc(inst, &inst->d_pmat, inst->pmat_stride * cfg->n_pmat));
  CREATE_RC(dev_alloc(inst, &inst->d_site_lnl, P));
  CREATE_RC(dev_alloc(inst, &inst->d_site_lk, P));
  CREATE_RC(dev_alloc(inst, &inst->d_site_lk_cat, P * cfg->ncatg));
  CREATE_RC(dev_alloc(inst, &inst->d_fact, P));
  CREATE_RC(dev_alloc(inst, &inst->d_partials, (size_t)kMaxReduceBlocks * 2));
  CREATE_RC(dev_alloc(inst, &inst->d_warn, 1));
  CREATE_RC(dev_alloc(inst, &inst->d_ticket, 1));
  CREATE_RC(dev_alloc(inst, &inst->d_result, 4));
  CREATE_RC(dev_alloc(inst, &inst->d_stage, (size_t)kStageSlots * kStageBytes));
  CREATE_TRY(cudaMemsetAsync(inst->d_wght, 0, (P + kSitePad) * sizeof(double), inst->stream));
  CREATE_TRY(cudaMemsetAsync(inst->d_invar, 0xff, P * sizeof(short), inst->stream));
  CREATE_TRY(cudaMemsetAsync(inst->d_tipmask, 0, 256 * sizeof(uint32_t), inst->stream));
  CREATE_TRY(cudaMemsetAsync(inst->d_tipcodes, 0, inst->tip_stride * cfg->n_tips + kSitePad, inst->stream));
  CREATE_TRY(cudaMemsetAsync(inst->d_warn, 0, sizeof(int), inst->stream));
  CREATE_TRY(cudaMemsetAsync(inst->d_ticket, 0, sizeof(unsigned int), inst->stream));
  CREATE_TRY(cudaMemsetAsync(inst->d_pmat, 0, inst->pmat_stride * cfg->n_pmat * sizeof(double), inst->stream));
  CREATE_TRY(cudaMemsetAsync(inst->d_fact, 0, P * sizeof(int), inst->stream));

  CREATE_TRY(cudaHostAlloc((void **)&inst->h_result, sizeof(ResultHost), cudaHostAllocMapped));
  memset(inst->h_result, 0, sizeof(ResultHost));
  CREATE_TRY(cudaHostGetDevicePointer((void **)&inst->h_result_dev, inst->h_result, 0));
  for (int s = 0; s < kStageSlots; ++s)
  {
    CREATE_TRY(cudaHostAlloc((void **)&inst->h_stage[s], kStageBytes, cudaHostAllocDefault));
    CREATE_TRY(cudaEventCreateWithFlags(&inst->stage_ev[s], cudaEventDisableTiming));
  }
  CREATE_TRY(cudaStreamSynchronize(inst->stream));
#undef CREATE_TRY
#undef CREATE_RC
  *out = inst;
  return PLK_OK;
}

void plk_destroy(plk_instance *inst)
{
  if (!inst) return;
  if (!inst->shards.empty())
  {
    for (plk_instance *sh : inst->shards) plk_destroy(sh);
    delete inst;
    return;
  }
  cudaSetDevice(inst->cfg.device);
  if (inst->stream) cudaStreamSynchronize(inst->stream);
  if (inst->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(inst->comm);
  for (void *p : inst->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(inst->d_mbox);
  cudaFree(inst->d_peers);
  for (double *p : inst->clv) cudaFree(p);
  for (int *p : inst->scale) cudaFree(p);
  cudaFree(inst->d_wght);
  cudaFree(inst->d_invar);
  cudaFree(inst->d_tipmask);
  cudaFree(inst->d_tipcodes);
  cudaFree(inst->d_tippacked);
  cudaFree(inst->d_tiprows);
  cudaFree(inst->d_model);
  cudaFree(inst->d_pmat);
  cudaFree(inst->d_site_lnl);
  cudaFree(inst->d_site_lk);
  cudaFree(inst->d_site_lk_cat);
  cudaFree(inst->d_fact);
  cudaFree(inst->d_dot_prod);
  cudaFree(inst->d_tmp_clv);
  cudaFree(inst->d_ops_cache);
  cudaFree(inst->d_partials);
  cudaFree(inst->d_warn);
  cudaFree(inst->d_ticket);
  cudaFree(inst->d_result);
  cudaFree(inst->d_stage);
  for (int2 *p : inst->pars_fitch) cudaFree(p);
  for (int *p : inst->pars_sank) cudaFree(p);
  cudaFree(inst->d_step_mat);
  cudaFree(inst->d_site_pars);
  cudaFree(inst->d_spr_pmat);
  cudaFree(inst->d_spr_partials);
  cudaFree(inst->d_spr_lnl);
  cudaFree(inst->d_spr_warn);
  if (inst->h_result) cudaFreeHost(inst->h_result);
  for (int s = 0; s < kStageSlots; ++s)
  {
    if (inst->h_stage[s]) cudaFreeHost(inst->h_stage[s]);
    if (inst->stage_ev[s]) cudaEventDestroy(inst->stage_ev[s]);
  }
  if (inst->copy_stream) cudaStreamDestroy(inst->copy_stream);
  if (inst->ev_copy) cudaEventDestroy(inst->ev_copy);
  if (inst->ev_unpack) cudaEventDestroy(inst->ev_unpack);
  if (inst->stream) cudaStreamDestroy(inst->stream);
  delete inst;
}

int plk_sync(plk_instance *inst)
{
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_sync(sh));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  return PLK_OK;
}

// ---- uploads -----------------------------------------------------------------------------------
int plk_set_pattern_weights(plk_instance *inst, const double *wght, const short *invar)
{
  ARG_CHECK(inst, wght != nullptr, "wght is NULL");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_set_pattern_weights(sh, wght + lo, invar ? invar + lo : nullptr));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const size_t P = inst->cfg.n_patterns;
  inst->wght_integral = true;
  for (size_t s = 0; s < P; ++s)
    if (!(wght[s] >= 0.0 && wght[s] < 9.0e15 && wght[s] == (double)(long long)wght[s])) inst->wght_integral = false;
  CU_TRY(inst, cudaMemcpyAsync(inst->d_wght, wght, P * sizeof(double), cudaMemcpyHostToDevice, inst->stream));
  if (invar)
    CU_TRY(inst, cudaMemcpyAsync(inst->d_invar, invar, P * sizeof(short), cudaMemcpyHostToDevice, inst->stream));
  return PLK_OK;
}

static int upload_masks(plk_instance *inst)
{
  uint32_t tmp[256];
  memset(tmp, 0, sizeof(tmp));
  for (size_t i = 0; i < inst->masks.size(); ++i) tmp[i] = inst->masks[i];
  void *d = nullptr;
  int   rc = stage_upload(inst, tmp, sizeof(tmp), &d);
  if (rc) return rc;
  CU_TRY(inst, cudaMemcpyAsync(inst->d_tipmask, d, sizeof(tmp), cudaMemcpyDeviceToDevice, inst->stream));
  inst->tiprows_dirty = true;
  return PLK_OK;
}

int plk_set_tip_table(plk_instance *inst, int n_codes, const double *vectors)
{
  ARG_CHECK(inst, n_codes >= 1 && n_codes <= 256 && vectors, "tip table: 1..256 codes");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_set_tip_table(sh, n_codes, vectors));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const int ns = inst->cfg.ns;
  inst->tip_table.assign(vectors, vectors + (size_t)n_codes * ns);
  inst->masks.assign(n_codes, 0u);
  inst->mask_to_code.clear();
  for (int c = 0; c < n_codes; ++c)
  {
    uint32_t m = 0;
    for (int k = 0; k < ns; ++k)
      if (vectors[(size_t)c * ns + k] > 0.0) m |= (1u << k);
    inst->masks[c] = m;
    if (!inst->mask_to_code.count(m)) inst->mask_to_code[m] = c;
  }
  return upload_masks(inst);
}

int plk_set_tip_codes(plk_instance *inst, int tip, const uint8_t *codes)
{
  ARG_CHECK(inst, tip >= 0 && tip < inst->cfg.n_tips && codes, "tip index out of range");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_set_tip_codes(sh, tip, codes + lo));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  CU_TRY(inst, cudaMemcpyAsync(inst->d_tipcodes + (size_t)tip * inst->tip_stride, codes, inst->cfg.n_patterns,
                               cudaMemcpyHostToDevice, inst->stream));
  inst->tiprows_dirty = true;
  return PLK_OK;
}

int plk_set_all_tip_codes(plk_instance *inst, const uint8_t *codes, size_t host_stride)
{
  ARG_CHECK(inst, codes && host_stride >= (size_t)inst->cfg.n_patterns, "plk_set_all_tip_codes: bad arguments");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_set_all_tip_codes(sh, codes + lo, host_stride));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  if (host_stride == inst->tip_stride)
    CU_TRY(inst, cudaMemcpyAsync(inst->d_tipcodes, codes, inst->tip_stride * inst->cfg.n_tips, cudaMemcpyHostToDevice,
                                 inst->stream));
  else
    CU_TRY(inst, cudaMemcpy2DAsync(inst->d_tipcodes, inst->tip_stride, codes, host_stride, inst->cfg.n_patterns,
                                   inst->cfg.n_tips, cudaMemcpyHostToDevice, inst->stream));
  inst->tiprows_dirty = true;
  return PLK_OK;
}

// 4-bit tip codes: two patterns per byte (low nibble = the even pattern).  Halves the bytes a full upload moves
// over PCIe (1 byte per (tip, pattern) otherwise holds a 4-bit nucleotide mask); the unpack kernel also writes
// the tip-table rows the fused traversal kernels read, so no separate translation launch follows.
int plk_set_all_tip_codes_packed4(plk_instance *inst, const uint8_t *packed, size_t host_stride)
{
  ARG_CHECK(inst, packed && host_stride >= ((size_t)inst->cfg.n_patterns + 1) / 2, "plk_set_all_tip_codes_packed4: bad arguments");
  if (!inst->shards.empty())
  {
    for (size_t i = 0; i + 1 < inst->shard_lo.size(); ++i)
      ARG_CHECK(inst, (inst->shard_lo[i] & 1) == 0, "plk_set_all_tip_codes_packed4: a shard starts on an odd pattern");
    FOR_SHARDS(inst, plk_set_all_tip_codes_packed4(sh, packed + lo / 2, host_stride));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  ARG_CHECK(inst, inst->masks.size() <= 16, "plk_set_all_tip_codes_packed4: the tip table has more than 16 codes");
  const size_t pstride = inst->tip_stride / 2, row = ((size_t)inst->cfg.n_patterns + 1) / 2;
  if (!inst->d_tippacked)
  {
    int rc = dev_alloc(inst, &inst->d_tippacked, pstride * inst->cfg.n_tips);
    if (rc) return rc;
    CU_TRY(inst, cudaStreamCreateWithFlags(&inst->copy_stream, cudaStreamNonBlocking));
    CU_TRY(inst, cudaEventCreateWithFlags(&inst->ev_copy, cudaEventDisableTiming));
    CU_TRY(inst, cudaEventCreateWithFlags(&inst->ev_unpack, cudaEventDisableTiming));
    CU_TRY(inst, cudaMemsetAsync(inst->d_tippacked, 0, pstride * inst->cfg.n_tips, inst->stream));
    CU_TRY(inst, cudaEventRecord(inst->ev_unpack, inst->stream));
  }
  // the copy only has to wait for the previous unpack (the last reader of the staging buffer), not for the compute
  // enqueued since: it overlaps a traversal that is still running; the unpack that follows is ordered on the
  // instance's stream as before, so every later call sees the new codes
  CU_TRY(inst, cudaStreamWaitEvent(inst->copy_stream, inst->ev_unpack, 0));
  CU_TRY(inst, cudaMemcpy2DAsync(inst->d_tippacked, pstride, packed, host_stride, row, inst->cfg.n_tips,
                                 cudaMemcpyHostToDevice, inst->copy_stream));
  CU_TRY(inst, cudaEventRecord(inst->ev_copy, inst->copy_stream));
  CU_TRY(inst, cudaStreamWaitEvent(inst->stream, inst->ev_copy, 0));
  const size_t n = pstride * inst->cfg.n_tips;
  const int    mode = inst->fused_dna ? 1 : (inst->fused_aa ? 2 : 0);
  k_unpack_codes4<<<(unsigned)std::min<size_t>((n / 8 + 255) / 256 + 1, 8192), 256, 0, inst->stream>>>(
      inst->d_tippacked, inst->d_tipcodes, mode ? inst->d_tiprows : nullptr, n, inst->d_tipmask, mode);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  CU_TRY(inst, cudaEventRecord(inst->ev_unpack, inst->stream));
  inst->tiprows_dirty = false;
  return PLK_OK;
}

int plk_set_tip_vectors(plk_instance *inst, int tip, const double *v)
{
  ARG_CHECK(inst, tip >= 0 && tip < inst->cfg.n_tips && v, "tip index out of range");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_set_tip_vectors(sh, tip, v + lo * (size_t)inst->cfg.ns));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const int            ns = inst->cfg.ns, P = inst->cfg.n_patterns;
  std::vector<uint8_t> codes(P);
  bool                 table_grew = false;
  for (int s = 0; s < P; ++s)
  {
    uint32_t m = 0;
    for (int k = 0; k < ns; ++k)
      if (v[(size_t)s * ns + k] > 0.0) m |= (1u << k);
    auto it = inst->mask_to_code.find(m);
    int  code;
    if (it == inst->mask_to_code.end())
    {
      ARG_CHECK(inst, inst->masks.size() < 256, "more than 256 distinct tip state sets");
      code = (int)inst->masks.size();
      inst->masks.push_back(m);
      inst->mask_to_code[m] = code;
      for (int k = 0; k < ns; ++k) inst->tip_table.push_back((m >> k) & 1u ? 1.0 : 0.0);
      table_grew = true;
    }
    else
      code = it->second;
    codes[s] = (uint8_t)code;
  }
  if (table_grew)
  {
    int rc = upload_masks(inst);
    if (rc) return rc;
  }
  int rc = plk_set_tip_codes(inst, tip, codes.data());
  if (rc) return rc;
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));  // `codes` is a local buffer
  return PLK_OK;
}

int plk_set_model(plk_instance *inst, const double *U, const double *V, const double *lambda, const double *pi,
                  const double *rates, const double *rate_probs, double pinv, int invar_flag, double l_min,
                  double l_max, double br_len_mult)
{
  ARG_CHECK(inst, U && V && lambda && pi && rates && rate_probs, "plk_set_model: NULL array");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_set_model(sh, U, V, lambda, pi, rates, rate_probs, pinv, invar_flag, l_min, l_max, br_len_mult));
    inst->l_min = l_min;
    inst->l_max = l_max;
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const int ns = inst->cfg.ns, nc = inst->cfg.ncatg;
  ModelDev *m = new ModelDev();
  memset(m, 0, sizeof(ModelDev));
  memcpy(m->U, U, sizeof(double) * ns * ns);
  memcpy(m->V, V, sizeof(double) * ns * ns);
  memcpy(m->lambda, lambda, sizeof(double) * ns);
  memcpy(m->pi, pi, sizeof(double) * ns);
  memcpy(m->rates, rates, sizeof(double) * nc);
  memcpy(m->probs, rate_probs, sizeof(double) * nc);
  m->pinv = pinv;
  m->invar_flag = invar_flag;
  m->l_min = l_min;
  m->l_max = l_max;
  m->br_len_mult = br_len_mult;
  void *d = nullptr;
  int   rc = stage_upload(inst, m, sizeof(ModelDev), &d);
  delete m;
  if (rc) return rc;
  CU_TRY(inst, cudaMemcpyAsync(inst->d_model, d, sizeof(ModelDev), cudaMemcpyDeviceToDevice, inst->stream));
  inst->eigen_ready = false;
  inst->l_min = l_min;
  inst->l_max = l_max;
  return PLK_OK;
}

// ---- K0 ------------------------------------------------------------------------------------------
int plk_update_pmats(plk_instance *inst, int n, const int *pmat, const double *l)
{
  ARG_CHECK(inst, n >= 0 && (n == 0 || (pmat && l)), "plk_update_pmats: bad arguments");
  if (!inst->shards.empty())
  {  // P-matrices are replicated (bytes): every shard computes its own copy
    FOR_SHARDS(inst, plk_update_pmats(sh, n, pmat, l));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const int    ns = inst->cfg.ns, nc = inst->cfg.ncatg;
  const int    per_slot = (int)(kStageBytes / sizeof(PmatJob));
  const int    threads = ((ns * ns + 31) / 32) * 32;
  const size_t smem = (size_t)(kMaxNs + ns * ns) * sizeof(double);
  if (n > 0 && n <= kPmatInlineSmall)
  {  // latency path (one candidate / one edge): small parameter block
    PmatJobsInlineSmall jobs;
    jobs.base = inst->d_pmat;
    jobs.stride = (unsigned)inst->pmat_stride;
    for (int i = 0; i < n; ++i)
    {
      ARG_CHECK(inst, pmat[i] >= 0 && pmat[i] < inst->cfg.n_pmat, "pmat handle out of range");
      jobs.h[i] = pmat[i];
      jobs.l[i] = l[i];
    }
    for (int i = n; i < kPmatInlineSmall; ++i) jobs.h[i] = 0, jobs.l[i] = 0.0;
    k_pmat_inline_small<<<n * nc, threads, smem, inst->stream>>>(jobs, inst->d_model, ns, nc, ns == 4 ? 1 : (ns == 20 ? 2 : 0));
    inst->launches++;
    CU_TRY(inst, cudaGetLastError());
    return PLK_OK;
  }
  if (n <= kPmatInline)
  {  // job list in the kernel's parameter block: one launch, nothing staged
    static thread_local PmatJobsInline jobs;
    jobs.base = inst->d_pmat;
    jobs.stride = (unsigned)inst->pmat_stride;
    for (int i = 0; i < n; ++i)
    {
      ARG_CHECK(inst, pmat[i] >= 0 && pmat[i] < inst->cfg.n_pmat, "pmat handle out of range");
      jobs.h[i] = pmat[i];
      jobs.l[i] = l[i];
    }
    if (n > 0)
    {
      k_pmat_inline<<<n * nc, threads, smem, inst->stream>>>(jobs, inst->d_model, ns, nc, ns == 4 ? 1 : (ns == 20 ? 2 : 0));
      inst->launches++;
      CU_TRY(inst, cudaGetLastError());
    }
    return PLK_OK;
  }
  for (int off = 0; off < n; off += per_slot)
  {
    const int             cnt = std::min(per_slot, n - off);
    std::vector<PmatJob> jobs(cnt);
    for (int i = 0; i < cnt; ++i)
    {
      const int h = pmat[off + i];
      ARG_CHECK(inst, h >= 0 && h < inst->cfg.n_pmat, "pmat handle out of range");
      jobs[i].P = inst->d_pmat + (size_t)h * inst->pmat_stride;
      jobs[i].l = l[off + i];
    }
    void *d = nullptr;
    int   rc = stage_upload(inst, jobs.data(), sizeof(PmatJob) * cnt, &d);
    if (rc) return rc;
    k_pmat<<<cnt * nc, threads, smem, inst->stream>>>((const PmatJob *)d, inst->d_model, ns, nc, ns == 4 ? 1 : (ns == 20 ? 2 : 0));
    inst->launches++;
    CU_TRY(inst, cudaGetLastError());
  }
  return PLK_OK;
}

int plk_set_pmat(plk_instance *inst, int pmat, const double *P)
{
  ARG_CHECK(inst, pmat >= 0 && pmat < inst->cfg.n_pmat && P, "pmat handle out of range");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_set_pmat(sh, pmat, P));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  std::vector<double> rec(inst->pmat_stride);
  memcpy(rec.data(), P, inst->pmat_elems * sizeof(double));
  if (inst->cfg.ns == 4)
  {  // tip table TP[c][mask][i] = sum_{j in mask} P[c][i][j], ascending j (same order as k_pmat)
    for (int c = 0; c < inst->cfg.ncatg; ++c)
      for (int m = 0; m < 16; ++m)
        for (int i = 0; i < 4; ++i)
        {
          const double *row = P + (size_t)c * 16 + i * 4;
          double        a = (m & 1) ? row[0] : 0.0;
          if (m & 2) a = a + row[1];
          if (m & 4) a = a + row[2];
          if (m & 8) a = a + row[3];
          rec[inst->pmat_elems + (size_t)c * 64 + tip_row4(m) * 4 + i] = a;
        }
  }
  if (inst->cfg.ns == 20)
  {  // tPx[c][s][i] = P[c][i][s], row 20 = ascending-j row sums (same as k_pmat)
    for (int c = 0; c < inst->cfg.ncatg; ++c)
    {
      const double *Pc = P + (size_t)c * 400;
      double       *TX = rec.data() + inst->pmat_elems + (size_t)c * 420;
      for (int i = 0; i < 20; ++i)
      {
        double a = Pc[i * 20];
        for (int j = 1; j < 20; ++j) a = a + Pc[i * 20 + j];
        TX[20 * 20 + i] = a;
        for (int sidx = 0; sidx < 20; ++sidx) TX[sidx * 20 + i] = Pc[i * 20 + sidx];
      }
      double *PF = rec.data() + inst->pmat_elems + (size_t)inst->cfg.ncatg * 420 + (size_t)c * 480;
      for (int q = 0; q < 480; ++q)
      {
        const int j = q >> 5, ln = q & 31, n = j / 5, kk = j % 5, gg = ln >> 2, tt = ln & 3;
        PF[q] = (n * 8 + gg < 20) ? Pc[(n * 8 + gg) * 20 + kk * 4 + tt] : 0.0;
      }
    }
  }
  const size_t b = inst->pmat_stride * sizeof(double);
  double      *dstp = inst->d_pmat + (size_t)pmat * inst->pmat_stride;
  if (b <= kStageBytes)
  {
    void *d = nullptr;
    int   rc = stage_upload(inst, rec.data(), b, &d);
    if (rc) return rc;
    CU_TRY(inst, cudaMemcpyAsync(dstp, d, b, cudaMemcpyDeviceToDevice, inst->stream));
  }
  else
  {
    CU_TRY(inst, cudaMemcpyAsync(dstp, rec.data(), b, cudaMemcpyHostToDevice, inst->stream));
    CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  }
  return PLK_OK;
}

int plk_get_pmat(plk_instance *inst, int pmat, double *P)
{
  ARG_CHECK(inst, pmat >= 0 && pmat < inst->cfg.n_pmat && P, "pmat handle out of range");
  if (!inst->shards.empty()) return plk_get_pmat(inst->shards[0], pmat, P);
  USE_DEVICE(inst);
  CU_TRY(inst, cudaMemcpyAsync(P, inst->d_pmat + (size_t)pmat * inst->pmat_stride, inst->pmat_elems * sizeof(double),
                               cudaMemcpyDeviceToHost, inst->stream));
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  return PLK_OK;
}

// ---- K1 ------------------------------------------------------------------------------------------
}  // extern "C"

// per-level launch of the generic kernel (ns != 4 or ncatg not a power of two)
static int launch_level_generic(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  const int ns = inst->cfg.ns, nc = inst->cfg.ncatg, P = inst->cfg.n_patterns;
  int       bpo = std::max(1, (P + 127) / 128);
  k_partial_generic<<<(unsigned)(bpo * n_ops), 128, 0, inst->stream>>>(d_ops, bpo, P, ns, nc, inst->d_wght,
                                                                        inst->d_tipmask, inst->apply_scaling);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

template <int NCATG, int UMAX>
static int launch_traverse_t(plk_instance *inst, const OpDev *d_ops, int n_ops, int tile_sites, int n_tiles, int grid)
{
  k_traverse_dna<NCATG, UMAX><<<grid, kTravThreads, 0, inst->stream>>>(d_ops, n_ops, inst->cfg.n_patterns, tile_sites,
                                                                       n_tiles, inst->d_wght, inst->apply_scaling);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

// fused 20-state traversal on the FP64 tensor pipe
// 20-state traversal, third generation (k_traverse_aa3): the tiling of k_traverse_aa, NCATG at compile time
template <int NCATG>
static int launch_traverse_aa3_t(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  const int     P = inst->cfg.n_patterns;
  const size_t  smem = (size_t)kAa3Stages * sizeof(Aa3Stage<NCATG>);
  auto          kern = k_traverse_aa3<NCATG>;
  static size_t smem_set[64] = {};  // per instantiation and device
  if (smem_set[inst->cfg.device & 63] == 0)
  {
    CU_TRY(inst, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[inst->cfg.device & 63] = smem;
  }
  const int       slots = inst->num_sms;  // one 512-thread block per SM
  const long long cap = (long long)slots * kAaTileCap;
  const int       rounds = (int)((P + cap - 1) / cap);
  int             n_tiles = std::max(1, std::min(slots * rounds, (P + 7) / 8));
  int             tile_sites = (P + n_tiles - 1) / n_tiles;
  tile_sites = ((tile_sites + 7) / 8) * 8;
  if (tile_sites > kAaTileCap) tile_sites = kAaTileCap;
  n_tiles = (P + tile_sites - 1) / tile_sites;
  const int grid = std::min(n_tiles, slots);
  kern<<<grid, kAaThreads, smem, inst->stream>>>(d_ops, n_ops, P, tile_sites, n_tiles, inst->d_wght, inst->d_tipmask,
                                                 inst->d_tiprows, inst->d_tipcodes, inst->apply_scaling);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

static int launch_traverse_aa(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  if (!inst->aa_v1)
    switch (inst->cfg.ncatg)
    {
    case 1: return launch_traverse_aa3_t<1>(inst, d_ops, n_ops);
    case 2: return launch_traverse_aa3_t<2>(inst, d_ops, n_ops);
    case 4: return launch_traverse_aa3_t<4>(inst, d_ops, n_ops);
    }
  // other category counts (3, 5..8): the first-generation kernel, whose 3 stages fit shared memory up to ncatg = 8
  // (4 stages of Aa3Stage<8> would need 460 KB)
  const int    nc = inst->cfg.ncatg, P = inst->cfg.n_patterns;
  const size_t smem = (size_t)kAaStages * aa_stage_bytes(nc);
  if (!inst->aa_attr_set)
  {
    CU_TRY(inst, cudaFuncSetAttribute(k_traverse_aa, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    inst->aa_attr_set = true;
  }
  const int slots = inst->num_sms;  // one 512-thread block per SM
  const long long cap = (long long)slots * kAaTileCap;
  const int       rounds = (int)((P + cap - 1) / cap);
  int             n_tiles = std::max(1, std::min(slots * rounds, (P + 7) / 8));
  int             tile_sites = (P + n_tiles - 1) / n_tiles;
  tile_sites = ((tile_sites + 7) / 8) * 8;
  if (tile_sites > kAaTileCap) tile_sites = kAaTileCap;
  n_tiles = (P + tile_sites - 1) / tile_sites;
  const int       grid = std::min(n_tiles, slots);
  k_traverse_aa<<<grid, kAaThreads, smem, inst->stream>>>(d_ops, n_ops, P, nc, tile_sites, n_tiles, inst->d_wght,
                                                          inst->d_tipmask, inst->d_tiprows, inst->d_tipcodes,
                                                          inst->apply_scaling);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

template <int NCATG, int U>
static int launch_traverse_mma_t(plk_instance *inst, const OpDev *d_ops, int n_ops, int tile_sites, int n_tiles, int grid)
{
  k_traverse_dna_mma<NCATG, U><<<grid, kTravThreads, 0, inst->stream>>>(d_ops, n_ops, inst->cfg.n_patterns, tile_sites,
                                                                        n_tiles, inst->d_wght, inst->apply_scaling);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

// tensor-pipe variant of the 4-state traversal: a block tile is 7 warps x U m-groups x 8 sites
static int launch_traverse_mma(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  const int nc = inst->cfg.ncatg, P = inst->cfg.n_patterns;
  const int U = inst->mma_u;
  const int cap_sites = kTravComputeWarps * U * 8;
  const int slots = inst->num_sms * inst->trav_blocks_per_sm;
  const long long cap = (long long)slots * cap_sites;
  const int       rounds = (int)((P + cap - 1) / cap);
  int             n_tiles = std::max(1, std::min(slots * rounds, (P + 7) / 8));
  int             tile_sites = (P + n_tiles - 1) / n_tiles;
  tile_sites = std::min(((tile_sites + 7) / 8) * 8, cap_sites);
  n_tiles = (P + tile_sites - 1) / tile_sites;
  const int grid = std::min(n_tiles, slots);
#define MMA_CASE(NC)                                                                                   \
  case NC:                                                                                             \
    return launch_traverse_mma_t<NC, 2>(inst, d_ops, n_ops, tile_sites, n_tiles, grid);
  switch (nc)
  {
    MMA_CASE(1)
    MMA_CASE(2)
    MMA_CASE(4)
  }
#undef MMA_CASE
  inst->err = "internal: unsupported ncatg for the tensor-pipe traversal kernel";
  return PLK_ERR_ARG;
}

// fused traversal launch: tiles of `tile_sites` patterns, persistent blocks (2 per SM), every block
// gets the same number of equally sized tiles so there is no tail wave
static int launch_traverse(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  const int nc = inst->cfg.ncatg, P = inst->cfg.n_patterns;
  const int CT = kTravComputeWarps * 32;
  const int umax = inst->trav_umax;
  const int slots = inst->num_sms * inst->trav_blocks_per_sm;
  const long long items = (long long)P * nc;
  const long long cap = (long long)slots * CT * umax;  // items resident in one round
  const int       rounds = (int)((items + cap - 1) / cap);
  int             n_tiles = slots * rounds;
  // do not cut tiles below one pass of the block
  const int min_tile = std::max(1, CT / nc);
  n_tiles = std::max(1, std::min(n_tiles, (P + min_tile - 1) / min_tile));
  int tile_sites = (P + n_tiles - 1) / n_tiles;
  tile_sites = std::min(((tile_sites + 7) / 8) * 8, std::max(8, (CT * umax / nc) / 8 * 8));  // whole 8-site blocks
  n_tiles = (P + tile_sites - 1) / tile_sites;
  const int grid = std::min(n_tiles, slots);
  if ((long long)tile_sites * nc > (long long)CT * umax)
  {
    inst->err = "internal: traversal tile exceeds block capacity";
    return PLK_ERR_ARG;
  }
#define TRAV_CASE(NC)                                                                          \
  case NC:                                                                                     \
    return umax == 2 ? launch_traverse_t<NC, 2>(inst, d_ops, n_ops, tile_sites, n_tiles, grid) \
                     : launch_traverse_t<NC, 1>(inst, d_ops, n_ops, tile_sites, n_tiles, grid);
  switch (nc)
  {
    TRAV_CASE(1)
    TRAV_CASE(2)
    TRAV_CASE(4)
    TRAV_CASE(8)
  }
#undef TRAV_CASE
  inst->err = "internal: unsupported ncatg for traversal kernel";
  return PLK_ERR_ARG;
}

// op-major 4-state traversal (k_traverse_dna2): one tile of 32-item chunks per block and launch
template <int NCATG, int W, int MINB>
static int launch_traverse2_t(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  constexpr int    SW = 32 / NCATG;
  constexpr size_t kSmemPerSm = 227 * 1024;
  const int        total_chunks = (inst->cfg.n_patterns + SW - 1) / SW;
  const size_t     budget = kSmemPerSm / MINB - 2048;
  int              cap = (int)((budget - t2_smem_bytes<NCATG>(0)) / kT2ChunkBytes);
  cap = std::min(cap, 32 * W) & ~1;  // live mask is 32 bits per warp; even: tiles start on 8-site blocks for NCATG = 8
  const int       slots = inst->num_sms * MINB;
  const long long per_round = (long long)slots * cap;
  const int       rounds = (int)((total_chunks + per_round - 1) / per_round);
  int             n_tiles = std::max(1, std::min(slots * rounds, total_chunks));
  int             tile_chunks = (total_chunks + n_tiles - 1) / n_tiles;
  tile_chunks = std::min((tile_chunks + 1) & ~1, cap);
  n_tiles = (total_chunks + tile_chunks - 1) / tile_chunks;
  const int    grid = std::min(n_tiles, slots);
  const size_t smem = t2_smem_bytes<NCATG>(tile_chunks);
  auto         kern = k_traverse_dna2<NCATG, W, MINB>;
  static size_t smem_set[64] = {};  // per instantiation and device
  if (smem_set[inst->cfg.device & 63] == 0)
  {
    CU_TRY(inst, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(budget)));
    smem_set[inst->cfg.device & 63] = budget;
  }
  kern<<<grid, (W + 1) * 32, smem, inst->stream>>>(d_ops, n_ops, total_chunks, tile_chunks, n_tiles, inst->d_wght,
                                                    inst->apply_scaling);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

template <int NCATG>
static int launch_traverse2_nc(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  constexpr int kW1 = (NCATG == 8) ? 10 : 11;  // 8 categories: a chunk is half an 8-site block, W must be even
  switch (inst->t2_variant)
  {
  case 5: return launch_traverse2_t<NCATG, (NCATG == 8) ? 4 : 5, 2>(inst, d_ops, n_ops);
  default: return launch_traverse2_t<NCATG, kW1, 1>(inst, d_ops, n_ops);
  }
}

// op-major tensor-pipe 4-state traversal (k_traverse_dna3): chunks of one 8-site block
template <int NCATG, int W, int MINB>
static int launch_traverse3_t(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  constexpr size_t kSmemPerSm = 227 * 1024;
  const int        total_chunks = (inst->cfg.n_patterns + 7) / 8;
  const size_t     budget = kSmemPerSm / MINB - 2048;
  int              cap = (int)((budget - t3_smem_bytes<NCATG>(0)) / t3_chunk_bytes<NCATG>());
  cap = std::min(std::min(cap, 32 * W), kT3MaxTileChunks) & ~1;  // live mask: 32 bits per warp; staged tip rows
  const int       slots = inst->num_sms * MINB;
  const int       pairs = (total_chunks + 1) / 2;  // tiles are whole pairs of chunks (t3_tile_range)
  const long long per_round = (long long)slots * (cap / 2);
  const int       rounds = (int)((pairs + per_round - 1) / per_round);
  const int       n_tiles = (int)std::max<long long>(1, std::min<long long>((long long)slots * rounds, pairs));
  const int       tile_chunks = 2 * ((pairs + n_tiles - 1) / n_tiles);  // the largest tile
  const int       grid = std::min(n_tiles, slots);
  const size_t    smem = t3_smem_bytes<NCATG>(tile_chunks);
  auto            kern = k_traverse_dna3<NCATG, W, MINB>;
  static size_t smem_set[64] = {};  // per instantiation and device
  if (smem_set[inst->cfg.device & 63] == 0)
  {
    CU_TRY(inst, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(budget)));
    smem_set[inst->cfg.device & 63] = budget;
  }
  kern<<<grid, (W + 1) * 32, smem, inst->stream>>>(d_ops, n_ops, total_chunks, tile_chunks, n_tiles, inst->d_wght,
                                                    inst->apply_scaling);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

template <int NCATG>
static int launch_traverse3_nc(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  switch (inst->t2_variant)
  {
  case 13: return launch_traverse3_t<NCATG, 7, 2>(inst, d_ops, n_ops);
  default: return launch_traverse3_t<NCATG, 15, 1>(inst, d_ops, n_ops);
  }
}

static int launch_traverse3(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  switch (inst->cfg.ncatg)
  {
  case 1: return launch_traverse3_nc<1>(inst, d_ops, n_ops);
  case 2: return launch_traverse3_nc<2>(inst, d_ops, n_ops);
  case 4: return launch_traverse3_nc<4>(inst, d_ops, n_ops);
  case 8: return launch_traverse3_nc<8>(inst, d_ops, n_ops);
  }
  inst->err = "internal: unsupported ncatg for traversal kernel";
  return PLK_ERR_ARG;
}

// argument block of the 4-state edge reduction (stand-alone kernel or fused epilogue of the traversal kernel);
// takes the next reduction sequence number
static EdgeDev make_edge_dev(plk_instance *inst, plk_side left, plk_side rght, int pmat)
{
  EdgeDev e;
  e.left = side_dev(inst, left);
  e.rght = side_dev(inst, rght);
  e.P = inst->d_pmat + (size_t)pmat * inst->pmat_stride;
  e.mod = inst->d_model;
  e.wght = inst->d_wght;
  e.invar = inst->d_invar;
  e.tipmask = inst->d_tipmask;
  e.site_lnl = inst->d_site_lnl;
  e.site_lk_out = inst->d_site_lk;
  e.site_lk_cat = inst->d_site_lk_cat;
  e.fact_sum_scale = inst->d_fact;
  e.npat = inst->cfg.n_patterns;
  e.enabled = 1;
  e.ro = make_reduce_out(inst);
  return e;
}

// op-major tensor-pipe 4-state traversal, round-robin items (k_traverse_dna4): chunks of two 8-site blocks
template <int NCATG, int W, bool PF>
static int launch_traverse4_t(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  constexpr size_t kSmemPerSm = 227 * 1024;
  const int        total_chunks = (inst->cfg.n_patterns + 15) / 16;
  const size_t     budget = kSmemPerSm - 4096;  // static shared memory: barriers + the reduction scratch of the fused epilogue
  int              cap = (int)((budget - t4_smem_bytes<NCATG>(0)) / (t4_smem_bytes<NCATG>(1) - t4_smem_bytes<NCATG>(0)));
  cap = std::min(cap, kT4MaxTileChunks);
  const int       slots = inst->num_sms;
  const long long per_round = (long long)slots * cap;
  const int       rounds = (int)((total_chunks + per_round - 1) / per_round);
  const int       n_tiles = (int)std::max<long long>(1, std::min<long long>((long long)slots * rounds, total_chunks));
  const int       tile_chunks = (total_chunks + n_tiles - 1) / n_tiles;  // the largest tile
  const int       grid = std::min(n_tiles, slots);
  const size_t    smem = t4_smem_bytes<NCATG>(tile_chunks);
  auto            kern = k_traverse_dna4<NCATG, W, PF>;
  static size_t   smem_set[64] = {};  // per instantiation and device
  if (smem_set[inst->cfg.device & 63] == 0)
  {
    CU_TRY(inst, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(budget)));
    smem_set[inst->cfg.device & 63] = budget;
  }
  EdgeDev edge;
  memset(&edge, 0, sizeof(edge));
  if (inst->pending_edge && NCATG == 4)
  {  // the edge reduction rides on this launch (its operands exist now: every destination has been allocated)
    int rc = check_side(inst, inst->edge_left, true);
    if (rc) return rc;
    rc = check_side(inst, inst->edge_rght, true);
    if (rc) return rc;
    ARG_CHECK(inst, inst->edge_pmat >= 0 && inst->edge_pmat < inst->cfg.n_pmat, "plk_traverse_edge_lnl: bad arguments");
    edge = make_edge_dev(inst, inst->edge_left, inst->edge_rght, inst->edge_pmat);
    inst->edge_fused = true;
    inst->pending_edge = false;
  }
  kern<<<grid, (W + 1) * 32, smem, inst->stream>>>(d_ops, n_ops, total_chunks, tile_chunks, n_tiles, inst->d_wght,
                                                    inst->apply_scaling, edge);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

template <int NCATG>
static int launch_traverse4_nc(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  switch (inst->t2_variant)
  {
  case 21: return launch_traverse4_t<NCATG, 15, true>(inst, d_ops, n_ops);  // with the L1 look-ahead (measured slower: 0.269 vs 0.249 ms)
  default: return launch_traverse4_t<NCATG, 15, false>(inst, d_ops, n_ops);
  }
}

static int launch_traverse4(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  switch (inst->cfg.ncatg)
  {
  case 1: return launch_traverse4_nc<1>(inst, d_ops, n_ops);
  case 2: return launch_traverse4_nc<2>(inst, d_ops, n_ops);
  case 4: return launch_traverse4_nc<4>(inst, d_ops, n_ops);
  case 8: return launch_traverse4_nc<8>(inst, d_ops, n_ops);
  }
  inst->err = "internal: unsupported ncatg for traversal kernel";
  return PLK_ERR_ARG;
}

static int launch_traverse2(plk_instance *inst, const OpDev *d_ops, int n_ops)
{
  switch (inst->cfg.ncatg)
  {
  case 1: return launch_traverse2_nc<1>(inst, d_ops, n_ops);
  case 2: return launch_traverse2_nc<2>(inst, d_ops, n_ops);
  case 4: return launch_traverse2_nc<4>(inst, d_ops, n_ops);
  case 8: return launch_traverse2_nc<8>(inst, d_ops, n_ops);
  }
  inst->err = "internal: unsupported ncatg for traversal kernel";
  return PLK_ERR_ARG;
}

extern "C" {

int plk_update_partials(plk_instance *inst, int n_ops, const plk_op *ops)
{
  ARG_CHECK(inst, n_ops >= 0 && (n_ops == 0 || ops), "plk_update_partials: bad arguments");
  if (n_ops == 0) return PLK_OK;
  if (!inst->shards.empty())
  {  // asynchronous on every device's stream: the shards run concurrently
    FOR_SHARDS(inst, plk_update_partials(sh, n_ops, ops));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const int  nclv = inst->cfg.n_clv;
  const int  nc = inst->cfg.ncatg;
  const bool fused_dna = inst->fused_dna, fused_aa = inst->fused_aa;
  const bool fused = fused_dna || fused_aa;
  if (fused && inst->tiprows_dirty)
  {
    const size_t n = inst->tip_stride * inst->cfg.n_tips;
    if (fused_dna)
      k_codes_to_rows<<<(unsigned)std::min<size_t>((n + 255) / 256, 4096), 256, 0, inst->stream>>>(
          inst->d_tipcodes, inst->d_tiprows, n, inst->d_tipmask);
    else
      k_codes_to_rows20<<<(unsigned)std::min<size_t>((n + 255) / 256, 4096), 256, 0, inst->stream>>>(
          inst->d_tipcodes, inst->d_tiprows, n, inst->d_tipmask);
    inst->launches++;
    CU_TRY(inst, cudaGetLastError());
    inst->tiprows_dirty = false;
  }
  const int per_slot = (int)(kStageBytes / sizeof(OpDev));
  auto launch_fused = [&](const OpDev *d_ops, int n) {
    return (fused_dna && inst->dna_mma && nc <= 4) ? launch_traverse_mma(inst, d_ops, n)
           : (fused_dna && inst->trav_v1)          ? launch_traverse(inst, d_ops, n)
           : (fused_dna && inst->t2_variant >= 20) ? launch_traverse4(inst, d_ops, n)
           : (fused_dna && inst->t2_variant >= 10) ? launch_traverse3(inst, d_ops, n)
           : fused_dna                             ? launch_traverse2(inst, d_ops, n)
                                                   : launch_traverse_aa(inst, d_ops, n);
  };
  if (fused && inst->d_ops_cache && inst->cache_epoch == inst->alloc_epoch && (int)inst->cache_ops.size() == n_ops &&
      memcmp(inst->cache_ops.data(), ops, sizeof(plk_op) * (size_t)n_ops) == 0)
    return launch_fused(inst->d_ops_cache, n_ops);  // same list as last time: descriptors are still valid

  // dependency levels: an op runs after the last writer of each operand it reads, after the last
  // reader of the buffer it overwrites and after the last writer of that buffer (RAW, WAR, WAW)
  inst->lvl_write.assign(nclv, -1);
  inst->lvl_read.assign(nclv, -1);
  inst->op_level.assign(n_ops, 0);
  int max_level = 0;
  for (int i = 0; i < n_ops; ++i)
  {
    const plk_op &o = ops[i];
    ARG_CHECK(inst, o.dst >= 0 && o.dst < nclv, "dst handle out of range");
    ARG_CHECK(inst, o.pmat1 >= 0 && o.pmat1 < inst->cfg.n_pmat && o.pmat2 >= 0 && o.pmat2 < inst->cfg.n_pmat,
              "pmat handle out of range");
    int rc = check_side(inst, o.c1, false);
    if (rc) return rc;
    rc = check_side(inst, o.c2, false);
    if (rc) return rc;
    int lv = std::max(inst->lvl_write[o.dst], inst->lvl_read[o.dst]) + 1;
    const plk_side *cs[2] = {&o.c1, &o.c2};
    for (const plk_side *c : cs)
      if (c->clv >= 0)
      {
        ARG_CHECK(inst, c->clv != o.dst, "an update cannot read the buffer it writes");
        if (inst->lvl_write[c->clv] < 0)
          ARG_CHECK(inst, inst->clv[c->clv] != nullptr, "operand CLV was never computed");
        lv = std::max(lv, inst->lvl_write[c->clv] + 1);
      }
    if (lv < 0) lv = 0;
    inst->op_level[i] = lv;
    inst->lvl_write[o.dst] = lv;
    for (const plk_side *c : cs)
      if (c->clv >= 0) inst->lvl_read[c->clv] = std::max(inst->lvl_read[c->clv], lv);
    max_level = std::max(max_level, lv);
    rc = ensure_clv(inst, o.dst);
    if (rc) return rc;
  }
  std::vector<int> order(n_ops);
  if (fused)
  {
    // the fused kernel executes the list in program order per thread: the caller's order is valid
    for (int i = 0; i < n_ops; ++i) order[i] = i;
  }
  else
  {  // stable bucket by level
    std::vector<int> cnt(max_level + 2, 0);
    for (int i = 0; i < n_ops; ++i) cnt[inst->op_level[i] + 1]++;
    for (int l = 0; l <= max_level; ++l) cnt[l + 1] += cnt[l];
    for (int i = 0; i < n_ops; ++i) order[cnt[inst->op_level[i]]++] = i;
  }
  std::vector<OpDev> host;
  host.reserve(std::min(n_ops, per_slot));
  int pos = 0;
  while (pos < n_ops)
  {
    host.clear();
    std::vector<std::pair<int, int>> launches;  // (offset in block, count): one per level chunk / fused chunk
    while (pos < n_ops && (int)host.size() < per_slot)
    {
      const int lv = inst->op_level[order[pos]];
      const int start = (int)host.size();
      while (pos < n_ops && (fused || inst->op_level[order[pos]] == lv) && (int)host.size() < per_slot)
      {
        const plk_op &o = ops[order[pos]];
        OpDev         d;
        d.dst = inst->clv[o.dst];
        d.dst_scale = inst->scale[o.dst];
        plk_side x = o.c1, y = o.c2;
        int      px = o.pmat1, py = o.pmat2;
        int      kind = 0;
        if (fused_dna)
        {
          // canonical operand order (the product of the children commutes exactly): a child that is
          // the previous update's destination first (forwarded in registers), CLVs before tips
          const double *prev_dst = ((int)host.size() > start) ? host.back().dst : nullptr;
          const bool    yfwd = y.clv >= 0 && prev_dst && inst->clv[y.clv] == prev_dst;
          const bool    xfwd = x.clv >= 0 && prev_dst && inst->clv[x.clv] == prev_dst;
          if ((yfwd && !xfwd) || (!xfwd && x.tip >= 0 && y.clv >= 0))
          {
            std::swap(x, y);
            std::swap(px, py);
          }
          const bool afwd = x.clv >= 0 && prev_dst && inst->clv[x.clv] == prev_dst;
          const int  ka = afwd ? kSrcFwd : (x.clv >= 0 ? kSrcSlot : kSrcTip);
          const int  kb = (y.tip >= 0) ? kSrcTip : (ka == kSrcSlot ? kSrcLate : kSrcSlot);
          kind = ka | (kb << 2);
        }
        const SideDev a = side_dev(inst, x), b = side_dev(inst, y);
        d.c1 = a.clv;
        d.s1 = a.scale;
        d.t1 = a.tip;
        d.c2 = b.clv;
        d.s2 = b.scale;
        d.t2 = b.tip;
        d.P1 = inst->d_pmat + (size_t)px * inst->pmat_stride;
        d.P2 = inst->d_pmat + (size_t)py * inst->pmat_stride;
        if (fused)
        {  // tip operands read the edge's tip table instead of P, and pre-translated row indices
          if (x.tip >= 0)
          {
            d.P1 += inst->pmat_elems;
            d.t1 = inst->d_tiprows + (size_t)x.tip * inst->tip_stride;
          }
          else if (fused_aa)
            d.P1 += inst->pmat_elems + (size_t)nc * 420;  // fragment-ordered P
          if (y.tip >= 0)
          {
            d.P2 += inst->pmat_elems;
            d.t2 = inst->d_tiprows + (size_t)y.tip * inst->tip_stride;
          }
          else if (fused_aa)
            d.P2 += inst->pmat_elems + (size_t)nc * 420;
        }
        if (fused_aa) kind = (x.tip >= 0 ? 1 : 0) | (y.tip >= 0 ? 2 : 0);
        d.flags = kind;
        d.pad[0] = d.pad[1] = d.pad[2] = 0;
        host.push_back(d);
        ++pos;
      }
      launches.emplace_back(start, (int)host.size() - start);
    }
    void *d = nullptr;
    int   rc = stage_upload(inst, host.data(), host.size() * sizeof(OpDev), &d);
    if (rc) return rc;
    if (fused && n_ops <= per_slot && n_ops >= 8)
    {  // remember the resolved descriptors of a whole-list launch (a traversal) for the next identical call
      if (!inst->d_ops_cache)
      {
        rc = dev_alloc(inst, &inst->d_ops_cache, (size_t)per_slot);
        if (rc) return rc;
      }
      CU_TRY(inst, cudaMemcpyAsync(inst->d_ops_cache, d, host.size() * sizeof(OpDev), cudaMemcpyDeviceToDevice,
                                   inst->stream));
      inst->cache_ops.assign(ops, ops + n_ops);
      inst->cache_epoch = inst->alloc_epoch;
    }
    for (auto &lc : launches)
    {
      rc = fused ? launch_fused((const OpDev *)d + lc.first, lc.second)
                 : launch_level_generic(inst, (const OpDev *)d + lc.first, lc.second);
      if (rc) return rc;
    }
  }
  return PLK_OK;
}

// ---- K2 ------------------------------------------------------------------------------------------
// sum of the shards' partial results in shard order (or, with the in-process P2P exchange, the all-shard sum
// that every shard's reduction kernel computed itself)
static int finish_sharded(plk_instance *inst, double *out0, double *out1, int *warn)
{
  double t0 = 0.0, t1 = 0.0;
  int    w = 0;
  for (size_t i = 0; i < inst->shards.size(); ++i)
  {
    double a = 0.0, b = 0.0;
    int    ww = 0;
    const int rc = finish_reduction(inst->shards[i], &a, &b, &ww);
    if (rc)
    {
      inst->err = "shard " + std::to_string(i) + ": " + inst->shards[i]->err;
      return rc;
    }
    if (inst->inproc_p2p)
    {
      if (i == 0) t0 = a, t1 = b, w = ww;
    }
    else
    {
      t0 += a;
      t1 += b;
      w |= ww;
    }
  }
  if (out0) *out0 = t0;
  if (out1) *out1 = t1;
  if (warn) *warn = w;
  return PLK_OK;
}

static int edge_lnl_launch(plk_instance *inst, plk_side left, plk_side rght, int pmat)
{
  USE_DEVICE(inst);
  int rc = check_side(inst, left, true);
  if (rc) return rc;
  rc = check_side(inst, rght, true);
  if (rc) return rc;
  ARG_CHECK(inst, pmat >= 0 && pmat < inst->cfg.n_pmat, "plk_edge_lnl: bad arguments");
  if (inst->fused_dna && inst->cfg.ncatg == 4)
  {  // coalesced 4-state kernel on the blocked layout: thread per (site, category)
    k_edge_lnl_dna<4><<<reduce_grid_dna(inst), reduce_threads_dna(), 0, inst->stream>>>(make_edge_dev(inst, left, rght, pmat));
  }
  else
  {
    const int grid = reduce_grid(inst, 128);
    k_edge_lnl<<<grid, 128, 0, inst->stream>>>(side_dev(inst, left), side_dev(inst, rght),
                                               inst->d_pmat + (size_t)pmat * inst->pmat_stride, inst->d_model,
                                               inst->cfg.n_patterns, inst->cfg.ns, inst->cfg.ncatg, inst->d_wght,
                                               inst->d_invar, inst->d_tipmask, inst->d_site_lnl, inst->d_site_lk,
                                               inst->d_site_lk_cat, inst->d_fact, make_reduce_out(inst), inst->blocked);
  }
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  inst->site_valid = true;
  return PLK_OK;
}

int plk_edge_lnl(plk_instance *inst, plk_side left, plk_side rght, int pmat, double *lnl, int *warn)
{
  ARG_CHECK(inst, lnl != nullptr, "plk_edge_lnl: bad arguments");
  if (!inst->shards.empty())
  {  // all shards' reduction kernels are in flight before the first result is awaited
    FOR_SHARDS(inst, edge_lnl_launch(sh, left, rght, pmat));
    return finish_sharded(inst, lnl, nullptr, warn);
  }
  const int rc = edge_lnl_launch(inst, left, rght, pmat);
  if (rc) return rc;
  return finish_reduction(inst, lnl, nullptr, warn);
}

// ---- K1 + K2 in one launch ------------------------------------------------------------------------
// Post_Order_Lk followed by the site loop of Lk at one edge (lk.c:562-645).  When the whole list is one launch of
// the 4-state traversal kernel the edge reduction runs as that kernel's epilogue; otherwise the two launches follow
// each other on the stream.  Same results as plk_update_partials + plk_edge_lnl up to the summation order of
// the per-block partial sums (the reduction stays deterministic: fixed partition, fixed order).
static int traverse_edge_launch(plk_instance *inst, int n_ops, const plk_op *ops, plk_side left, plk_side rght, int pmat)
{
  static const bool disabled = getenv("PLK_NO_FUSED_EDGE") != nullptr;
  const int         per_slot = (int)(kStageBytes / sizeof(OpDev));
  const bool        can_fuse = !disabled && inst->fused_dna && inst->cfg.ncatg == 4 && inst->t2_variant >= 20 &&
                        !inst->trav_v1 && !inst->dna_mma && n_ops > 0 && n_ops <= per_slot;
  inst->edge_fused = false;
  if (can_fuse)
  {
    inst->edge_left = left;
    inst->edge_rght = rght;
    inst->edge_pmat = pmat;
    inst->pending_edge = true;
  }
  const int rc = plk_update_partials(inst, n_ops, ops);
  inst->pending_edge = false;
  if (rc) return rc;
  if (inst->edge_fused)
  {
    inst->site_valid = true;
    return PLK_OK;
  }
  return edge_lnl_launch(inst, left, rght, pmat);
}

int plk_traverse_edge_lnl(plk_instance *inst, int n_ops, const plk_op *ops, plk_side left, plk_side rght, int pmat,
                          double *lnl, int *warn)
{
  ARG_CHECK(inst, lnl != nullptr && n_ops >= 0 && (n_ops == 0 || ops), "plk_traverse_edge_lnl: bad arguments");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, traverse_edge_launch(sh, n_ops, ops, left, rght, pmat));
    return finish_sharded(inst, lnl, nullptr, warn);
  }
  const int rc = traverse_edge_launch(inst, n_ops, ops, left, rght, pmat);
  if (rc) return rc;
  return finish_reduction(inst, lnl, nullptr, warn);
}

// Lk(NULL) after the host's model update as ONE call: the P-matrix loop (lk.c:500-505), Post_Order_Lk (:562-564)
// and the site loop at the root edge (:578-645): two launches for 4-state / 4-category data (K0, fused K1 + K2)
int plk_lk_full(plk_instance *inst, int n_pmat, const int *pmat, const double *l, int n_ops, const plk_op *ops,
                plk_side left, plk_side rght, int edge_pmat, double *lnl, int *warn)
{
  const int rc = plk_update_pmats(inst, n_pmat, pmat, l);
  if (rc) return rc;
  return plk_traverse_edge_lnl(inst, n_ops, ops, left, rght, edge_pmat, lnl, warn);
}

// the same evaluation split in two: everything is enqueued by _begin, the result is awaited by _wait; uploads of the
// NEXT evaluation's inputs may be issued in between (they are ordered behind this evaluation on the device, and the
// tip-code copy itself overlaps it)
int plk_lk_full_begin(plk_instance *inst, int n_pmat, const int *pmat, const double *l, int n_ops, const plk_op *ops,
                      plk_side left, plk_side rght, int edge_pmat)
{
  ARG_CHECK(inst, n_ops >= 0 && (n_ops == 0 || ops), "plk_lk_full_begin: bad arguments");
  const int rc = plk_update_pmats(inst, n_pmat, pmat, l);
  if (rc) return rc;
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, traverse_edge_launch(sh, n_ops, ops, left, rght, edge_pmat));
    return PLK_OK;
  }
  return traverse_edge_launch(inst, n_ops, ops, left, rght, edge_pmat);
}

int plk_lk_wait(plk_instance *inst, double *lnl, int *warn)
{
  ARG_CHECK(inst, lnl != nullptr, "plk_lk_wait: bad arguments");
  if (!inst->shards.empty()) return finish_sharded(inst, lnl, nullptr, warn);
  return finish_reduction(inst, lnl, nullptr, warn);
}

// ---- K3 ------------------------------------------------------------------------------------------
int plk_eigen_lr(plk_instance *inst, plk_side left, plk_side rght)
{
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_eigen_lr(sh, left, rght));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  int rc = check_side(inst, left, true);
  if (rc) return rc;
  rc = check_side(inst, rght, true);
  if (rc) return rc;
  if (!inst->d_dot_prod)
  {
    rc = dev_alloc(inst, &inst->d_dot_prod, clv_elems_plain(inst));
    if (rc) return rc;
    CU_TRY(inst, cudaMemsetAsync(inst->d_dot_prod, 0, clv_elems_plain(inst) * sizeof(double), inst->stream));
  }
  const long long work = (long long)inst->cfg.n_patterns * inst->cfg.ncatg;
  const int       grid = (int)std::max<long long>(1, std::min<long long>((work + 127) / 128, inst->num_sms * 32));
  static const bool generic_k3 = getenv("PLK_K3_GENERIC") != nullptr;
  if (inst->cfg.ns == 20 && !generic_k3)
    k_eigen_lr_reg<20><<<grid, 128, 0, inst->stream>>>(side_dev(inst, left), side_dev(inst, rght), inst->d_model,
                                                       inst->cfg.n_patterns, inst->cfg.ncatg, inst->d_wght,
                                                       inst->d_tipmask, inst->d_dot_prod, inst->d_fact, inst->blocked);
  else
    k_eigen_lr<<<grid, 128, 0, inst->stream>>>(side_dev(inst, left), side_dev(inst, rght), inst->d_model,
                                               inst->cfg.n_patterns, inst->cfg.ns, inst->cfg.ncatg, inst->d_wght,
                                               inst->d_tipmask, inst->d_dot_prod, inst->d_fact, inst->blocked);
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  inst->eigen_ready = true;
  return PLK_OK;
}

// ---- K4 ------------------------------------------------------------------------------------------
static int k4_launch(plk_instance *inst, double l, int deriv)
{
  USE_DEVICE(inst);
  if (!inst->eigen_ready)
  {
    inst->err = "plk_edge_lnl_dlnl / _eigen called before plk_eigen_lr (update_eigen_lr)";
    return PLK_ERR_STATE;
  }
  if (inst->cfg.ns == 4 && inst->cfg.ncatg == 4)
  {
    k_lnl_dlnl_dna<4><<<reduce_grid_dna(inst), reduce_threads_dna(), 0, inst->stream>>>(inst->d_dot_prod, inst->d_fact, inst->d_model, l, deriv,
                                                      inst->cfg.n_patterns, inst->d_wght, inst->d_invar,
                                                      inst->d_site_lnl, make_reduce_out(inst));
  }
  else
  {
    const int grid = reduce_grid(inst, 128);
    k_lnl_dlnl<<<grid, 128, 0, inst->stream>>>(inst->d_dot_prod, inst->d_fact, inst->d_model, l, deriv,
                                               inst->cfg.n_patterns, inst->cfg.ns, inst->cfg.ncatg, inst->d_wght,
                                               inst->d_invar, inst->d_site_lnl, make_reduce_out(inst));
  }
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  return PLK_OK;
}

static int run_k4(plk_instance *inst, double l, int deriv, double *lnl, double *dlnl, int *warn)
{
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, k4_launch(sh, l, deriv));
    return finish_sharded(inst, lnl, dlnl, warn);
  }
  const int rc = k4_launch(inst, l, deriv);
  if (rc) return rc;
  return finish_reduction(inst, lnl, dlnl, warn);
}

int plk_edge_lnl_dlnl(plk_instance *inst, double *l, double *lnl, double *dlnl, int *warn)
{
  ARG_CHECK(inst, l && lnl && dlnl, "plk_edge_lnl_dlnl: NULL argument");
  ARG_CHECK(inst, *l == *l, "plk_edge_lnl_dlnl: length is NaN");  // lk.c:671
  // lk.c:673-674
  if (*l < inst->l_min)
    *l = inst->l_min;
  else if (*l > inst->l_max)
    *l = inst->l_max;
  return run_k4(inst, *l, 1, lnl, dlnl, warn);
}

int plk_edge_lnl_eigen(plk_instance *inst, double l, double *lnl, int *warn)
{
  ARG_CHECK(inst, lnl, "plk_edge_lnl_eigen: NULL argument");
  double dummy;
  return run_k4(inst, l, 0, lnl, &dummy, warn);
}

// ---- read-backs ------------------------------------------------------------------------------------
static int ensure_tmp_clv(plk_instance *inst)
{
  if (inst->d_tmp_clv) return PLK_OK;
  return dev_alloc(inst, &inst->d_tmp_clv, clv_elems_plain(inst));
}

int plk_get_clv(plk_instance *inst, int h, double *clv_out, int *scale_out)
{
  ARG_CHECK(inst, h >= 0 && h < inst->cfg.n_clv, "clv handle out of range");
  if (!inst->shards.empty())
  {  // [site][catg][state]: a shard's block of patterns is a contiguous slice
    const size_t per_site = (size_t)inst->cfg.ncatg * inst->cfg.ns;
    FOR_SHARDS(inst, plk_get_clv(sh, h, clv_out ? clv_out + lo * per_site : nullptr, scale_out ? scale_out + lo : nullptr));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  ARG_CHECK(inst, inst->clv[h] != nullptr, "clv handle was never written");
  if (clv_out)
  {
    const double *src = inst->clv[h];
    if (inst->blocked)
    {  // device-side conversion to the reference's [site][catg][state] layout
      int rc = ensure_tmp_clv(inst);
      if (rc) return rc;
      k_clv_convert<<<inst->num_sms * 4, 256, 0, inst->stream>>>(inst->clv[h], inst->d_tmp_clv, inst->cfg.n_patterns,
                                                                   inst->cfg.ncatg, inst->cfg.ns, 0);
      inst->launches++;
      CU_TRY(inst, cudaGetLastError());
      src = inst->d_tmp_clv;
    }
    CU_TRY(inst, cudaMemcpyAsync(clv_out, src, clv_elems_plain(inst) * sizeof(double), cudaMemcpyDeviceToHost,
                                 inst->stream));
  }
  if (scale_out)
    CU_TRY(inst, cudaMemcpyAsync(scale_out, inst->scale[h], (size_t)inst->cfg.n_patterns * sizeof(int),
                                 cudaMemcpyDeviceToHost, inst->stream));
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  return PLK_OK;
}

int plk_set_clv(plk_instance *inst, int h, const double *clv_in, const int *scale_in)
{
  ARG_CHECK(inst, h >= 0 && h < inst->cfg.n_clv && clv_in, "clv handle out of range");
  if (!inst->shards.empty())
  {
    const size_t per_site = (size_t)inst->cfg.ncatg * inst->cfg.ns;
    FOR_SHARDS(inst, plk_set_clv(sh, h, clv_in + lo * per_site, scale_in ? scale_in + lo : nullptr));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  int rc = ensure_clv(inst, h);
  if (rc) return rc;
  if (inst->blocked)
  {
    rc = ensure_tmp_clv(inst);
    if (rc) return rc;
    CU_TRY(inst, cudaMemcpyAsync(inst->d_tmp_clv, clv_in, clv_elems_plain(inst) * sizeof(double), cudaMemcpyHostToDevice,
                                 inst->stream));
    k_clv_convert<<<inst->num_sms * 4, 256, 0, inst->stream>>>(inst->d_tmp_clv, inst->clv[h], inst->cfg.n_patterns,
                                                                 inst->cfg.ncatg, inst->cfg.ns, 1);
    inst->launches++;
    CU_TRY(inst, cudaGetLastError());
  }
  else
    CU_TRY(inst, cudaMemcpyAsync(inst->clv[h], clv_in, clv_elems_plain(inst) * sizeof(double), cudaMemcpyHostToDevice,
                                 inst->stream));
  if (scale_in)
    CU_TRY(inst, cudaMemcpyAsync(inst->scale[h], scale_in, (size_t)inst->cfg.n_patterns * sizeof(int),
                                 cudaMemcpyHostToDevice, inst->stream));
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  return PLK_OK;
}

int plk_get_site_lnl(plk_instance *inst, double *site_lnl, double *site_lk, double *site_lk_cat, int *fact)
{
  if (!inst->shards.empty())
  {
    const size_t nc = (size_t)inst->cfg.ncatg;
    FOR_SHARDS(inst, plk_get_site_lnl(sh, site_lnl ? site_lnl + lo : nullptr, site_lk ? site_lk + lo : nullptr,
                                      site_lk_cat ? site_lk_cat + lo * nc : nullptr, fact ? fact + lo : nullptr));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const size_t P = inst->cfg.n_patterns;
  if (site_lnl)
    CU_TRY(inst, cudaMemcpyAsync(site_lnl, inst->d_site_lnl, P * sizeof(double), cudaMemcpyDeviceToHost, inst->stream));
  if (site_lk)
    CU_TRY(inst, cudaMemcpyAsync(site_lk, inst->d_site_lk, P * sizeof(double), cudaMemcpyDeviceToHost, inst->stream));
  if (site_lk_cat)
    CU_TRY(inst, cudaMemcpyAsync(site_lk_cat, inst->d_site_lk_cat, P * inst->cfg.ncatg * sizeof(double),
                                 cudaMemcpyDeviceToHost, inst->stream));
  if (fact) CU_TRY(inst, cudaMemcpyAsync(fact, inst->d_fact, P * sizeof(int), cudaMemcpyDeviceToHost, inst->stream));
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  return PLK_OK;
}

int plk_get_dot_prod(plk_instance *inst, double *dot_prod)
{
  if (!inst->shards.empty())
  {
    const size_t per_site = (size_t)inst->cfg.ncatg * inst->cfg.ns;
    FOR_SHARDS(inst, plk_get_dot_prod(sh, dot_prod ? dot_prod + lo * per_site : nullptr));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  ARG_CHECK(inst, dot_prod && inst->d_dot_prod, "dot_prod not computed yet");
  CU_TRY(inst, cudaMemcpyAsync(dot_prod, inst->d_dot_prod, clv_elems_plain(inst) * sizeof(double), cudaMemcpyDeviceToHost,
                               inst->stream));
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  return PLK_OK;
}

// ---- NCCL ------------------------------------------------------------------------------------------
int plk_comm_unique_id(void *id128)
{
  if (!id128) return PLK_ERR_ARG;
  if (!g_nccl.load(g_create_error)) return PLK_ERR_NCCL;
  NcclApi::UniqueId id;
  const int         rc = g_nccl.GetUniqueId(&id);
  if (rc != 0)
  {
    g_create_error = "ncclGetUniqueId failed";
    return PLK_ERR_NCCL;
  }
  memcpy(id128, &id, 128);
  return PLK_OK;
}

int plk_comm_init(plk_instance *inst, int rank, int world, const void *id128)
{
  ARG_CHECK(inst, id128 && world >= 1 && rank >= 0 && rank < world, "plk_comm_init: bad arguments");
  if (!g_nccl.load(inst->err)) return PLK_ERR_NCCL;
  CU_TRY(inst, cudaSetDevice(inst->cfg.device));
  NcclApi::UniqueId id;
  memcpy(&id, id128, 128);
  const int rc = g_nccl.CommInitRank(&inst->comm, world, id, rank);
  if (rc != 0)
  {
    inst->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    inst->comm = nullptr;
    return PLK_ERR_NCCL;
  }
  inst->rank = rank;
  inst->world = world;
  inst->allreduce = true;
  return PLK_OK;
}

// ---- fused P2P exchange: mailboxes shared with CUDA IPC, written by the reduction kernels themselves ----
int plk_comm_p2p_export(plk_instance *inst, int world, void *handle64)
{
  ARG_CHECK(inst, handle64 && world >= 1 && world <= 64, "plk_comm_p2p_export: bad arguments");
  if (!inst->d_mbox)
  {
    int rc = dev_alloc(inst, &inst->d_mbox, (size_t)2 * world);
    if (rc) return rc;
    CU_TRY(inst, cudaMemset(inst->d_mbox, 0, sizeof(P2pSlot) * 2 * world));
  }
  cudaIpcMemHandle_t h;
  CU_TRY(inst, cudaIpcGetMemHandle(&h, inst->d_mbox));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  inst->world = world;
  return PLK_OK;
}

int plk_comm_p2p_init(plk_instance *inst, int rank, int world, const void *handles)
{
  ARG_CHECK(inst, handles && inst->d_mbox && world == inst->world && rank >= 0 && rank < world,
            "plk_comm_p2p_init: call plk_comm_p2p_export first");
  std::vector<P2pSlot *> peers(world, nullptr);
  for (int q = 0; q < world; ++q)
  {
    if (q == rank)
    {
      peers[q] = inst->d_mbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + (size_t)q * 64, 64);
    void *p = nullptr;
    CU_TRY(inst, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    inst->ipc_opened.push_back(p);
    peers[q] = (P2pSlot *)p;
  }
  int rc = dev_alloc(inst, &inst->d_peers, (size_t)world);
  if (rc) return rc;
  CU_TRY(inst, cudaMemcpy(inst->d_peers, peers.data(), sizeof(P2pSlot *) * world, cudaMemcpyHostToDevice));
  inst->rank = rank;
  inst->world = world;
  inst->p2p = true;
  inst->coll_seq = 0;       // all ranks wire their mailboxes at the same point of the program
  CU_TRY(inst, cudaMemset(inst->d_mbox, 0, sizeof(P2pSlot) * 2 * world));
  inst->allreduce = false;  // the exchange now happens inside the reduction kernel
  return PLK_OK;
}

// ---- single-process site sharding --------------------------------------------------------------------
// One host thread, n_gpus devices, one ordinary instance per device over a contiguous block of patterns.
// Every entry point fans out; the scalar-returning ones launch all shards' reduction kernels before the
// first result is awaited.  With distinct devices that can reach each other the 24-byte exchange runs inside
// the reduction kernels over NVLink peer memory (same mailbox protocol as plk_comm_p2p_*, peer access
// enabled in-process instead of CUDA IPC); otherwise the host adds the partial sums in shard order.
int plk_create_sharded(const plk_config *cfg, int n_gpus, const int *devices, plk_instance **out)
{
  if (!cfg || !out || n_gpus < 1 || n_gpus > 64)
  {
    g_create_error = "plk_create_sharded: bad arguments";
    return PLK_ERR_ARG;
  }
  *out = nullptr;
  if (cfg->n_patterns < n_gpus)
  {
    g_create_error = "plk_create_sharded: fewer patterns than shards";
    return PLK_ERR_ARG;
  }
  plk_instance *grp = new plk_instance();
  grp->cfg = *cfg;
  grp->shard_lo.assign(n_gpus + 1, 0);
  const long long P = cfg->n_patterns;
  for (int i = 1; i < n_gpus; ++i)
  {
    long long lo = P * i / n_gpus;
    if (P >= 64LL * n_gpus) lo &= ~31LL;  // whole 32-site chunks per shard when there are enough patterns
    grp->shard_lo[i] = (int)lo;
  }
  grp->shard_lo[n_gpus] = (int)P;
  g_multi_device = true;
  bool distinct = n_gpus > 1;
  for (int i = 0; i < n_gpus; ++i)
  {
    plk_config c = *cfg;
    c.n_patterns = grp->shard_lo[i + 1] - grp->shard_lo[i];
    c.device = devices ? devices[i] : i;
    for (int j = 0; j < i; ++j)
      if (grp->shards[j]->cfg.device == c.device) distinct = false;
    plk_instance *sh = nullptr;
    const int     rc = plk_create(&c, &sh);
    if (rc)
    {
      plk_destroy(grp);
      return rc;
    }
    grp->shards.push_back(sh);
  }
  if (distinct && !getenv("PLK_INPROC_HOSTSUM"))
  {  // in-kernel exchange over peer memory if every pair of devices can map each other
    bool ok = true;
    for (int i = 0; i < n_gpus && ok; ++i)
      for (int j = 0; j < n_gpus && ok; ++j)
        if (i != j)
        {
          int can = 0;
          if (cudaDeviceCanAccessPeer(&can, grp->shards[i]->cfg.device, grp->shards[j]->cfg.device) != cudaSuccess || !can)
            ok = false;
        }
    if (ok)
    {
      std::vector<P2pSlot *> boxes(n_gpus, nullptr);
      for (int i = 0; i < n_gpus && ok; ++i)
      {
        plk_instance *sh = grp->shards[i];
        cudaSetDevice(sh->cfg.device);
        for (int j = 0; j < n_gpus; ++j)
          if (j != i)
          {
            const cudaError_t e = cudaDeviceEnablePeerAccess(grp->shards[j]->cfg.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
            (void)cudaGetLastError();
          }
        if (ok && (dev_alloc(sh, &sh->d_mbox, (size_t)2 * n_gpus) ||
                   cudaMemset(sh->d_mbox, 0, sizeof(P2pSlot) * 2 * n_gpus) != cudaSuccess))
          ok = false;
        boxes[i] = sh->d_mbox;
      }
      for (int i = 0; i < n_gpus && ok; ++i)
      {
        plk_instance *sh = grp->shards[i];
        cudaSetDevice(sh->cfg.device);
        if (dev_alloc(sh, &sh->d_peers, (size_t)n_gpus) ||
            cudaMemcpy(sh->d_peers, boxes.data(), sizeof(P2pSlot *) * n_gpus, cudaMemcpyHostToDevice) != cudaSuccess)
          ok = false;
        sh->rank = i;
        sh->world = n_gpus;
      }
      if (ok)
      {
        for (plk_instance *sh : grp->shards) sh->p2p = true, sh->coll_seq = 0;
        grp->inproc_p2p = true;
      }
    }
  }
  *out = grp;
  return PLK_OK;
}

int plk_comm_set_allreduce(plk_instance *inst, int enable)
{
  if (enable && !inst->comm)
  {
    inst->err = "plk_comm_set_allreduce: no communicator";
    return PLK_ERR_STATE;
  }
  inst->allreduce = enable != 0;
  return PLK_OK;
}

// ---- introspection -----------------------------------------------------------------------------------
long long plk_launch_count(const plk_instance *inst)
{
  long long n = inst->launches;
  for (const plk_instance *sh : inst->shards) n += sh->launches;
  return n;
}
size_t plk_device_bytes(const plk_instance *inst)
{
  size_t n = inst->bytes;
  for (const plk_instance *sh : inst->shards) n += sh->bytes;
  return n;
}
void *plk_stream(plk_instance *inst) { return (void *)(inst->shards.empty() ? inst->stream : inst->shards[0]->stream); }
int   plk_n_shards(const plk_instance *inst) { return inst->shards.empty() ? 1 : (int)inst->shards.size(); }

}  // extern "C"

// ---- parsimony (src/pars.c) -------------------------------------------------------------------------------
namespace
{
int pars_check(plk_instance *inst, int h, bool general, bool need_data)
{
  ARG_CHECK(inst, inst->pars_n > 0, "parsimony buffers were not created (plk_pars_create)");
  ARG_CHECK(inst, h >= 0 && h < inst->pars_n, "parsimony buffer handle out of range");
  if (general) ARG_CHECK(inst, inst->d_step_mat != nullptr, "step-matrix parsimony needs the step matrix (plk_pars_create)");
  if (need_data)
    ARG_CHECK(inst, general ? inst->pars_sank[h] != nullptr : inst->pars_fitch[h] != nullptr,
              "parsimony buffer read before it was ever written");
  return PLK_OK;
}

int pars_ensure(plk_instance *inst, int h, bool general)
{
  if (general)
  {
    if (inst->pars_sank[h]) return PLK_OK;
    const size_t n = inst->pars_pstride * inst->cfg.ns;
    int          rc = dev_alloc(inst, &inst->pars_sank[h], n);
    if (rc) return rc;
    CU_TRY(inst, cudaMemsetAsync(inst->pars_sank[h], 0, n * sizeof(int), inst->stream));  // mCalloc, make.c:459
    return PLK_OK;
  }
  if (inst->pars_fitch[h]) return PLK_OK;
  int rc = dev_alloc(inst, &inst->pars_fitch[h], (size_t)inst->cfg.n_patterns);
  if (rc) return rc;
  CU_TRY(inst, cudaMemsetAsync(inst->pars_fitch[h], 0, (size_t)inst->cfg.n_patterns * sizeof(int2), inst->stream));
  return PLK_OK;
}

void *pars_ptr(plk_instance *inst, int h, bool general)
{
  return general ? (void *)inst->pars_sank[h] : (void *)inst->pars_fitch[h];
}

// a result block that is published by the kernel itself, outside the cross-GPU exchange sequence
ReduceOut make_publish_out(plk_instance *inst)
{
  ReduceOut ro;
  ro.partials = inst->d_partials;
  ro.ticket = inst->d_ticket;
  ro.warn_flag = inst->d_warn;
  ro.dev_out = inst->d_result;
  ro.host_out = inst->h_result_dev;
  ro.seq = ++inst->seq;
  ro.coll_seq = 0;
  ro.publish = 1;
  ro.peers = nullptr;
  ro.rank = 0;
  ro.world = 1;
  return ro;
}

// one launch (per 10 922 updates): the whole list, then (edge_mode != 0) the site loop of Pars at (left, rght)
int pars_launch(plk_instance *inst, bool general, int n_ops, const plk_pars_op *ops, int edge_mode, int left, int rght)
{
  USE_DEVICE(inst);
  int rc;
  for (int i = 0; i < n_ops; ++i)
  {
    // sources first: an update may read the buffer it overwrites only if that buffer already holds data
    if ((rc = pars_check(inst, ops[i].c1, general, true))) return rc;
    if ((rc = pars_check(inst, ops[i].c2, general, true))) return rc;
    if ((rc = pars_check(inst, ops[i].dst, general, false))) return rc;
    if ((rc = pars_ensure(inst, ops[i].dst, general))) return rc;
  }
  ParsEdgeDev edge;
  memset(&edge, 0, sizeof(edge));
  if (edge_mode)
  {
    if ((rc = pars_check(inst, left, general, true))) return rc;
    if ((rc = pars_check(inst, rght, general, true))) return rc;
    edge.left = pars_ptr(inst, left, general);
    edge.rght = pars_ptr(inst, rght, general);
    edge.wght = inst->d_wght;
    edge.site_pars = inst->d_site_pars;
  }
  const int P = inst->cfg.n_patterns;
  const int per_slot = (int)(kStageBytes / sizeof(ParsOpDev));
  int       done = 0;
  std::vector<ParsOpDev> dev;
  do
  {
    const int n = std::min(per_slot, n_ops - done);
    const bool last = (done + n == n_ops);
    dev.resize((size_t)std::max(n, 1));
    for (int i = 0; i < n; ++i)
    {
      dev[i].dst = pars_ptr(inst, ops[done + i].dst, general);
      dev[i].c1 = pars_ptr(inst, ops[done + i].c1, general);
      dev[i].c2 = pars_ptr(inst, ops[done + i].c2, general);
    }
    void *d_ops = nullptr;
    if (n > 0 && (rc = stage_upload(inst, dev.data(), sizeof(ParsOpDev) * (size_t)n, &d_ops))) return rc;
    ParsEdgeDev e = edge;
    e.mode = last ? edge_mode : 0;
    if (e.mode == 1) e.ro = make_reduce_out(inst);
    if (general)
    {
      const int grid = std::max(1, std::min((P + 127) / 128, kMaxReduceBlocks));
      const int ns = inst->cfg.ns;
      if (ns == 4)
        k_pars_sankoff<4><<<grid, 128, 0, inst->stream>>>((const ParsOpDev *)d_ops, n, P, inst->pars_pstride, ns,
                                                          inst->d_step_mat, e);
      else if (ns == 20)
        k_pars_sankoff<20><<<grid, 128, 0, inst->stream>>>((const ParsOpDev *)d_ops, n, P, inst->pars_pstride, ns,
                                                           inst->d_step_mat, e);
      else
        k_pars_sankoff<0><<<grid, 128, 0, inst->stream>>>((const ParsOpDev *)d_ops, n, P, inst->pars_pstride, ns,
                                                          inst->d_step_mat, e);
    }
    else
    {
      // patterns per thread: 1 while one pattern per thread still fits 8 blocks per SM, more only for very long
      // alignments (the per-block partial sums of the epilogue bound the grid at kMaxReduceBlocks)
      const long long t1 = (long long)kParsThreads * inst->num_sms * 8;
      int             U = (P <= t1) ? 1 : (P <= 2 * t1) ? 2 : 4;
      while ((long long)U * kParsThreads * kMaxReduceBlocks < P && U < 8) U *= 2;
      ARG_CHECK(inst, (long long)U * kParsThreads * kMaxReduceBlocks >= P, "parsimony: too many patterns for one instance");
      const int grid = std::max(1, (int)(((long long)P + (long long)U * kParsThreads - 1) / ((long long)U * kParsThreads)));
      if (U == 1)
        k_pars_fitch<1><<<grid, kParsThreads, 0, inst->stream>>>((const ParsOpDev *)d_ops, n, P, e);
      else if (U == 2)
        k_pars_fitch<2><<<grid, kParsThreads, 0, inst->stream>>>((const ParsOpDev *)d_ops, n, P, e);
      else if (U == 4)
        k_pars_fitch<4><<<grid, kParsThreads, 0, inst->stream>>>((const ParsOpDev *)d_ops, n, P, e);
      else
        k_pars_fitch<8><<<grid, kParsThreads, 0, inst->stream>>>((const ParsOpDev *)d_ops, n, P, e);
    }
    inst->launches++;
    CU_TRY(inst, cudaGetLastError());
    done += n;
  } while (done < n_ops);
  if (edge_mode) inst->pars_site_valid = true;
  return PLK_OK;
}

// the weighted total of the site_pars just written, accumulated like the reference's int (pars.c:46)
int pars_chain(plk_instance *inst, int carry_in, int *carry_out)
{
  USE_DEVICE(inst);
  k_pars_chain<<<1, 32, 0, inst->stream>>>(inst->d_site_pars, inst->d_wght, inst->cfg.n_patterns, carry_in,
                                           make_publish_out(inst));
  inst->launches++;
  CU_TRY(inst, cudaGetLastError());
  double v = 0.0;
  const int rc = finish_reduction(inst, &v, nullptr, nullptr);
  if (rc) return rc;
  *carry_out = (int)v;
  return PLK_OK;
}

bool pars_all_integral(const plk_instance *inst)
{
  if (inst->shards.empty()) return inst->wght_integral;
  for (const plk_instance *sh : inst->shards)
    if (!sh->wght_integral) return false;
  return true;
}
}  // namespace

extern "C" {

int plk_pars_create(plk_instance *inst, int n_buffers, const int *step_mat)
{
  ARG_CHECK(inst, n_buffers > 0, "plk_pars_create: n_buffers must be positive");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_pars_create(sh, n_buffers, step_mat));
    inst->pars_n = n_buffers;
    return PLK_OK;
  }
  USE_DEVICE(inst);
  ARG_CHECK(inst, inst->pars_n == 0, "plk_pars_create: already created");
  const int ns = inst->cfg.ns;
  inst->pars_n = n_buffers;
  inst->pars_fitch.assign((size_t)n_buffers, nullptr);
  inst->pars_sank.assign((size_t)n_buffers, nullptr);
  inst->pars_pstride = ((size_t)inst->cfg.n_patterns + 31) & ~(size_t)31;
  int rc = dev_alloc(inst, &inst->d_site_pars, (size_t)inst->cfg.n_patterns);
  if (rc) return rc;
  if (step_mat)
  {
    if ((rc = dev_alloc(inst, &inst->d_step_mat, (size_t)ns * ns))) return rc;
    CU_TRY(inst, cudaMemcpyAsync(inst->d_step_mat, step_mat, sizeof(int) * ns * ns, cudaMemcpyHostToDevice, inst->stream));
    CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  }
  return PLK_OK;
}

int plk_pars_set_buffer(plk_instance *inst, int buf, const int *ui, const int *pars, const int *p_pars)
{
  ARG_CHECK(inst, (ui && pars) || p_pars, "plk_pars_set_buffer: nothing to upload");
  ARG_CHECK(inst, (ui == nullptr) == (pars == nullptr), "plk_pars_set_buffer: ui and pars go together");
  if (!inst->shards.empty())
  {
    const int ns = inst->shards[0]->cfg.ns;
    FOR_SHARDS(inst, plk_pars_set_buffer(sh, buf, ui ? ui + lo : nullptr, pars ? pars + lo : nullptr,
                                         p_pars ? p_pars + lo * ns : nullptr));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const size_t P = (size_t)inst->cfg.n_patterns;
  int          rc;
  if (ui)
  {
    if ((rc = pars_check(inst, buf, false, false)) || (rc = pars_ensure(inst, buf, false))) return rc;
    std::vector<int2> tmp(P);
    for (size_t s = 0; s < P; ++s) tmp[s] = make_int2(ui[s], pars[s]);
    CU_TRY(inst, cudaMemcpyAsync(inst->pars_fitch[buf], tmp.data(), P * sizeof(int2), cudaMemcpyHostToDevice, inst->stream));
    CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  }
  if (p_pars)
  {
    if ((rc = pars_check(inst, buf, true, false)) || (rc = pars_ensure(inst, buf, true))) return rc;
    const int        ns = inst->cfg.ns;
    std::vector<int> tmp(inst->pars_pstride * ns, 0);
    for (size_t s = 0; s < P; ++s)
      for (int j = 0; j < ns; ++j) tmp[(size_t)j * inst->pars_pstride + s] = p_pars[s * ns + j];
    CU_TRY(inst, cudaMemcpyAsync(inst->pars_sank[buf], tmp.data(), tmp.size() * sizeof(int), cudaMemcpyHostToDevice, inst->stream));
    CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  }
  return PLK_OK;
}

int plk_pars_get_buffer(plk_instance *inst, int buf, int *ui, int *pars, int *p_pars)
{
  if (!inst->shards.empty())
  {
    const int ns = inst->shards[0]->cfg.ns;
    FOR_SHARDS(inst, plk_pars_get_buffer(sh, buf, ui ? ui + lo : nullptr, pars ? pars + lo : nullptr,
                                         p_pars ? p_pars + lo * ns : nullptr));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  const size_t P = (size_t)inst->cfg.n_patterns;
  int          rc;
  if (ui || pars)
  {
    if ((rc = pars_check(inst, buf, false, true))) return rc;
    std::vector<int2> tmp(P);
    CU_TRY(inst, cudaMemcpyAsync(tmp.data(), inst->pars_fitch[buf], P * sizeof(int2), cudaMemcpyDeviceToHost, inst->stream));
    CU_TRY(inst, cudaStreamSynchronize(inst->stream));
    for (size_t s = 0; s < P; ++s)
    {
      if (ui) ui[s] = tmp[s].x;
      if (pars) pars[s] = tmp[s].y;
    }
  }
  if (p_pars)
  {
    if ((rc = pars_check(inst, buf, true, true))) return rc;
    const int        ns = inst->cfg.ns;
    std::vector<int> tmp(inst->pars_pstride * ns);
    CU_TRY(inst, cudaMemcpyAsync(tmp.data(), inst->pars_sank[buf], tmp.size() * sizeof(int), cudaMemcpyDeviceToHost, inst->stream));
    CU_TRY(inst, cudaStreamSynchronize(inst->stream));
    for (size_t s = 0; s < P; ++s)
      for (int j = 0; j < ns; ++j) p_pars[s * ns + j] = tmp[(size_t)j * inst->pars_pstride + s];
  }
  return PLK_OK;
}

int plk_pars_update(plk_instance *inst, int general, int n_ops, const plk_pars_op *ops)
{
  ARG_CHECK(inst, n_ops >= 0 && (n_ops == 0 || ops), "plk_pars_update: bad arguments");
  if (n_ops == 0) return PLK_OK;
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, pars_launch(sh, general != 0, n_ops, ops, 0, -1, -1));
    return PLK_OK;
  }
  return pars_launch(inst, general != 0, n_ops, ops, 0, -1, -1);
}

int plk_pars_traverse_edge(plk_instance *inst, int general, int n_ops, const plk_pars_op *ops, int left, int rght,
                           int *c_pars)
{
  ARG_CHECK(inst, c_pars != nullptr && n_ops >= 0 && (n_ops == 0 || ops), "plk_pars_traverse_edge: bad arguments");
  const bool integral = pars_all_integral(inst);
  const int  mode = integral ? 1 : 2;
  double     v = 0.0;
  int        rc;
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, pars_launch(sh, general != 0, n_ops, ops, mode, left, rght));
    if (integral)
    {
      if ((rc = finish_sharded(inst, &v, nullptr, nullptr))) return rc;
      *c_pars = (int)v;
      return PLK_OK;
    }
    int c = 0;
    FOR_SHARDS(inst, pars_chain(sh, c, &c));  // pattern order = shard order
    *c_pars = c;
    return PLK_OK;
  }
  if ((rc = pars_launch(inst, general != 0, n_ops, ops, mode, left, rght))) return rc;
  if (integral)
  {
    if ((rc = finish_reduction(inst, &v, nullptr, nullptr))) return rc;
    *c_pars = (int)v;
    return PLK_OK;
  }
  if (inst->world > 1 || inst->allreduce)
  {
    inst->err = "parsimony with non-integral pattern weights is not supported across processes";
    return PLK_ERR_UNSUPPORTED;
  }
  return pars_chain(inst, 0, c_pars);
}

int plk_pars_edge(plk_instance *inst, int general, int left, int rght, int *c_pars)
{
  return plk_pars_traverse_edge(inst, general, 0, nullptr, left, rght, c_pars);
}

int plk_get_site_pars(plk_instance *inst, int *site_pars)
{
  ARG_CHECK(inst, site_pars != nullptr, "plk_get_site_pars: NULL output");
  if (!inst->shards.empty())
  {
    FOR_SHARDS(inst, plk_get_site_pars(sh, site_pars + lo));
    return PLK_OK;
  }
  USE_DEVICE(inst);
  if (!inst->pars_site_valid)
  {
    inst->err = "plk_get_site_pars: no parsimony score has been computed yet";
    return PLK_ERR_STATE;
  }
  CU_TRY(inst, cudaMemcpyAsync(site_pars, inst->d_site_pars, (size_t)inst->cfg.n_patterns * sizeof(int),
                               cudaMemcpyDeviceToHost, inst->stream));
  CU_TRY(inst, cudaStreamSynchronize(inst->stream));
  return PLK_OK;
}

}  // extern "C"

// ---- batched SPR candidates (spr.c:589-650 for many regraft positions at once) ----------------------------
namespace
{
constexpr int kSprChunk = 2048;      // candidates per launch
constexpr int kSprBlocksPerCand = 64;

int spr_ensure(plk_instance *inst)
{
  if (inst->spr_cap) return PLK_OK;
  const size_t pm = (size_t)inst->cfg.ncatg * inst->cfg.ns * inst->cfg.ns;
  int          rc;
  if ((rc = dev_alloc(inst, &inst->d_spr_pmat, (size_t)(2 * kSprChunk + 1) * pm))) return rc;
  if ((rc = dev_alloc(inst, &inst->d_spr_partials, (size_t)kSprChunk * kSprBlocksPerCand))) return rc;
  if ((rc = dev_alloc(inst, &inst->d_spr_lnl, (size_t)kSprChunk))) return rc;
  if ((rc = dev_alloc(inst, &inst->d_spr_warn, (size_t)kSprChunk))) return rc;
  inst->spr_cap = kSprChunk;
  return PLK_OK;
}

// one shard / plain instance: lnl[i] += this instance's partial sum for candidate i, warn[i] |= its warning flag
int spr_candidates_local(plk_instance *inst, plk_side prune, double l_prune, int link_on_left, int n_cand,
                         const plk_spr_cand *cand, double *lnl, int *warn)
{
  USE_DEVICE(inst);
  int rc = check_side(inst, prune, true);
  if (rc) return rc;
  ARG_CHECK(inst, link_on_left || prune.clv >= 0, "plk_spr_candidates: a tip is always the right-hand side of its edge");
  for (int i = 0; i < n_cand; ++i)
    if ((rc = check_side(inst, cand[i].a, true)) || (rc = check_side(inst, cand[i].b, true))) return rc;
  if ((rc = spr_ensure(inst))) return rc;
  const int    ns = inst->cfg.ns, nc = inst->cfg.ncatg, P = inst->cfg.n_patterns;
  const size_t pm = (size_t)nc * ns * ns;
  const int    bpc = std::max(1, std::min(kSprBlocksPerCand, (P + kSprThreads * 4 - 1) / (kSprThreads * 4)));
  const int    threads = ((ns * ns + 31) / 32) * 32;
  const size_t smem = (size_t)(kMaxNs + ns * ns) * sizeof(double);
  std::vector<PmatJob>    jobs;
  std::vector<SprCandDev> dev;
  std::vector<double>     h_lnl;
  std::vector<int>        h_warn;
  for (int off = 0; off < n_cand; off += kSprChunk)
  {
    const int n = std::min(kSprChunk, n_cand - off);
    // K0 for this chunk: P(l_a), P(l_b) per candidate and P(l_prune) (models.c:257-326, lk.c:2296-2300)
    jobs.resize((size_t)2 * n + 1);
    dev.resize((size_t)n);
    for (int i = 0; i < n; ++i)
    {
      jobs[2 * i].P = inst->d_spr_pmat + (size_t)(2 * i) * pm;
      jobs[2 * i].l = cand[off + i].l_a;
      jobs[2 * i + 1].P = inst->d_spr_pmat + (size_t)(2 * i + 1) * pm;
      jobs[2 * i + 1].l = cand[off + i].l_b;
      dev[i].a = side_dev(inst, cand[off + i].a);
      dev[i].b = side_dev(inst, cand[off + i].b);
      dev[i].Pa = jobs[2 * i].P;
      dev[i].Pb = jobs[2 * i + 1].P;
    }
    jobs[2 * n].P = inst->d_spr_pmat + (size_t)(2 * n) * pm;
    jobs[2 * n].l = l_prune;
    void *d_jobs = nullptr, *d_cands = nullptr;
    if ((rc = stage_upload(inst, jobs.data(), sizeof(PmatJob) * jobs.size(), &d_jobs))) return rc;
    k_pmat<<<(unsigned)(jobs.size() * nc), threads, smem, inst->stream>>>((const PmatJob *)d_jobs, inst->d_model, ns, nc, 0);
    inst->launches++;
    CU_TRY(inst, cudaGetLastError());
    if ((rc = stage_upload(inst, dev.data(), sizeof(SprCandDev) * dev.size(), &d_cands))) return rc;
    CU_TRY(inst, cudaMemsetAsync(inst->d_spr_warn, 0, sizeof(int) * (size_t)n, inst->stream));
    static const bool generic_spr = getenv("PLK_SPR_GENERIC") != nullptr;
    if (inst->fused_dna && nc == 4 && !generic_spr)
      k_spr_candidates_dna4<<<(unsigned)(n * bpc), kSprThreads, 0, inst->stream>>>(
          (const SprCandDev *)d_cands, bpc, side_dev(inst, prune), jobs[2 * n].P, link_on_left, inst->d_model, P,
          inst->d_wght, inst->d_invar, inst->d_tipmask, inst->apply_scaling, inst->d_spr_partials, inst->d_spr_warn);
    else if (inst->fused_aa && nc == 4 && !generic_spr)
    {
      const size_t sm20 = (size_t)(3 * 4 * (20 * 20 + 2) + 20 + 4) * sizeof(double);
      if (!inst->spr_attr_set)
      {
        CU_TRY(inst, cudaFuncSetAttribute(k_spr_candidates_reg4<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm20));
        inst->spr_attr_set = true;
      }
      k_spr_candidates_reg4<20><<<(unsigned)(n * bpc), kSprThreads, sm20, inst->stream>>>(
          (const SprCandDev *)d_cands, bpc, side_dev(inst, prune), jobs[2 * n].P, link_on_left, inst->d_model, P,
          inst->d_wght, inst->d_invar, inst->d_tipmask, inst->apply_scaling, inst->d_spr_partials, inst->d_spr_warn);
    }
    else
      k_spr_candidates<<<(unsigned)(n * bpc), kSprThreads, 0, inst->stream>>>(
          (const SprCandDev *)d_cands, bpc, side_dev(inst, prune), jobs[2 * n].P, link_on_left, inst->d_model, P, ns, nc,
          inst->d_wght, inst->d_invar, inst->d_tipmask, inst->apply_scaling, inst->blocked, inst->d_spr_partials,
          inst->d_spr_warn);
    k_spr_finish<<<(n + 127) / 128, 128, 0, inst->stream>>>(inst->d_spr_partials, bpc, n, inst->d_spr_lnl);
    inst->launches += 2;
    CU_TRY(inst, cudaGetLastError());
    h_lnl.resize((size_t)n);
    h_warn.resize((size_t)n);
    CU_TRY(inst, cudaMemcpyAsync(h_lnl.data(), inst->d_spr_lnl, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, inst->stream));
    CU_TRY(inst, cudaMemcpyAsync(h_warn.data(), inst->d_spr_warn, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, inst->stream));
    CU_TRY(inst, cudaStreamSynchronize(inst->stream));
    for (int i = 0; i < n; ++i)
    {
      lnl[off + i] += h_lnl[i];
      if (warn) warn[off + i] |= h_warn[i];
    }
  }
  return PLK_OK;
}
}  // namespace

extern "C" int plk_spr_candidates(plk_instance *inst, plk_side prune, double l_prune, int link_on_left, int n_cand,
                                  const plk_spr_cand *cand, double *lnl, int *numerical_warning)
{
  ARG_CHECK(inst, n_cand >= 0 && (n_cand == 0 || (cand && lnl)), "plk_spr_candidates: bad arguments");
  if (inst->world > 1 || inst->allreduce)
  {
    inst->err = "plk_spr_candidates: not available on an instance that is one rank of a multi-process job";
    return PLK_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < n_cand; ++i)
  {
    lnl[i] = 0.0;
    if (numerical_warning) numerical_warning[i] = 0;
  }
  if (n_cand == 0) return PLK_OK;
  if (!inst->shards.empty())
  {  // all-shard sums, added in shard order
    FOR_SHARDS(inst, spr_candidates_local(sh, prune, l_prune, link_on_left, n_cand, cand, lnl, numerical_warning));
    return PLK_OK;
  }
  return spr_candidates_local(inst, prune, l_prune, link_on_left, n_cand, cand, lnl, numerical_warning);
}
