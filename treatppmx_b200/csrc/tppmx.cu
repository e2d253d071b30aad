// tppmx.cu — C ABI (include/tppmx.h) over the sm_100a kernels.  Host side: argument checks,
// layout conversion between the reference's Armadillo layouts and the device SoA layout,
// device memory ownership, launches and timing.  No compute happens on the host.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "layout.h"
#include "predict.cuh"
#include "rng.cuh"
#include "sim.cuh"
#include "sweep.cuh"
#include "summary.cuh"
#include "sweep_fast_args.h"
#include "update.cuh"
#include "dropin.cuh"

using namespace tppmx;

// Host <-> device copy ordered on the context's own stream, then waited for.  The library never
// uses the legacy default stream (whose copies would wait for, and stall, every other context's
// streams): contexts on different host threads overlap freely.
static cudaError_t cpy_sync(cudaStream_t st, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
  const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, st);
  return e != cudaSuccess ? e : cudaStreamSynchronize(st);
}

// One device allocation + one pinned host mirror of its upload region per batch.  A destroyed
// batch leaves its arena with the context, so create/destroy cycles (one batch per group of
// replicate datasets) cost no cudaMalloc / cudaFree after the first.
struct Arena {
  char *dbase = nullptr;
  size_t dcap = 0;
  char *hbase = nullptr;  // pinned
  size_t hcap = 0;
};
static void arena_free(Arena &a) {
  if (a.dbase) cudaFree(a.dbase);
  if (a.hbase) cudaFreeHost(a.hbase);
  a = Arena();
}

struct tppmx_ctx {
  int device;
  cudaStream_t stream;
  std::string err;
  Arena spare;
  int live_batches = 0;  // tppmx_destroy is deferred until the last batch is gone
  bool closed = false;
  // side stream of tppmx_batch_run: the summaries of a saved iteration (co-clustering counts,
  // predictive step) run beside the updates that do not feed them
  cudaStream_t side = nullptr;
  cudaEvent_t ev_sweep = nullptr, ev_upd = nullptr, ev_side = nullptr;
  // third stream of tppmx_batch_run: the horseshoe update (a ~100 us serial chain per chain that
  // nothing reads before the NEXT iteration's beta step) runs beside the gamma draws, the
  // summaries and the next sweep
  cudaStream_t hs = nullptr;
  cudaEvent_t ev_hs_go = nullptr, ev_hs_done = nullptr;
  bool hs_pending = false;
  // tppmx_set_interrupt_flag
  const volatile int32_t *interrupt = nullptr;
  int check_every = 64;
  // chain groups of tppmx_batch_run (TPPMX_OPT_CHAIN_GROUPS): each group of chains runs its iterations on its
  // own stream triple, so that the sweep of one group overlaps the updates of another -- chains are independent,
  // and at BASELINE's 1024 chains every kernel of the iteration is latency-bound (2.6 warps per scheduler)
  int sm_count = 148;
  std::vector<tppmx_ctx *> lanes;
  // device block of the drop-in's traces (tppmx_dm_ppmx_ct), kept across calls
  void *trace = nullptr;
  size_t trace_bytes = 0;
  cudaEvent_t ev_go = nullptr, ev_done = nullptr;
  bool side_pending = false;  // summaries of a saved iteration still running on `side`
};
static void ctx_release(tppmx_ctx *c) {
  cudaSetDevice(c->device);
  for (tppmx_ctx *l : c->lanes) ctx_release(l);
  if (c->trace) cudaFree(c->trace);
  c->lanes.clear();
  if (c->ev_go) cudaEventDestroy(c->ev_go);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  cudaStreamDestroy(c->stream);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_sweep) cudaEventDestroy(c->ev_sweep);
  if (c->ev_upd) cudaEventDestroy(c->ev_upd);
  if (c->ev_side) cudaEventDestroy(c->ev_side);
  if (c->hs) cudaStreamDestroy(c->hs);
  if (c->ev_hs_go) cudaEventDestroy(c->ev_hs_go);
  if (c->ev_hs_done) cudaEventDestroy(c->ev_hs_done);
  arena_free(c->spare);
  delete c;
}

struct tppmx_batch {
  tppmx_ctx *ctx;
  DevBatch d;  // device pointers + dims
  tppmx_dims dims;
  Arena ar;
  bool dry = false;                // sizing pass of tppmx_batch_create
  size_t up_need = 0, st_need = 0; // bytes: upload region, state region
  size_t up_used = 0, st_used = 0;
  // scratch reserved at creation (no allocation on the call paths)
  int *init_lab = nullptr, *init_ncl = nullptr;
  double *init_beta = nullptr;
  double *pred_pi = nullptr, *pred_y = nullptr, *score_w = nullptr;
  int *pred_c = nullptr, *score_k = nullptr;
  // posterior summaries (summary.cuh); psm_* stay null when the counts would not fit
  int *psm_cnt = nullptr, *psm_n = nullptr, *pi_n = nullptr, *best_arm = nullptr;
  double *pi_sum = nullptr, *psm_mean = nullptr, *pi_mean = nullptr, *utility = nullptr, *util_w = nullptr;
  std::vector<int> arm_n;     // [nd][T]
  std::vector<int> catvec_host;  // [ncat] categories per categorical covariate
  std::vector<int> chain_ds;  // [nc+1]
  cudaEvent_t ev0, ev1;
  double last_ms;
  int last_launches;
  int scratch_chain;  // slot index n_chains
  // fast sweep kernel (sweep_fast.cuh)
  int opt_impl = 0, opt_lpu = 0, opt_ks = 0, opt_groups = 0;
  int last_groups = 1;
  FastArgs fa;
  UpdArgs ua{};
  bool fast_ready = false;
  int last_impl = 0, last_handovers = 0;
  // long-arm kernel: staged capacity per launch from the largest K seen (pinned read-back, no sync)
  int big_ksmax = 0, big_nth = 0, kmax_hint = 0, opt_big_nth = 0, big_hint = 0;
  bool big_two_stage = false;
  int *kmax_host = nullptr;  // pinned
  // special instantiations of the fast kernel (calibration 1, NNIG, categorical, double dipper) stage lanes - CC clusters:
  // the hand-over count and the largest K of the previous launch are read back (pinned, no sync), and the automatic
  // choice is 32 lanes per unit while partitions come close to 16 - CC clusters, 16 lanes otherwise
  int *bail_host = nullptr;  // [2]: hand-overs, largest K at the end of the launch
  int last_lpu = 0;
};

#define CK(ctx, call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                    \
      return TPPMX_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

static int fail(tppmx_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  return code;
}

extern "C" int tppmx_version(void) { return TPPMX_VERSION; }

static tppmx_ctx *ctx_new(int device) {
  tppmx_ctx *c = new tppmx_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_sweep, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_upd, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->hs, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_hs_go, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_hs_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_go, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return nullptr;
  }
  return c;
}

extern "C" int tppmx_create(tppmx_ctx **out, int device) {
  if (!out) return TPPMX_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return TPPMX_ERR_CUDA;
  tppmx_ctx *c = ctx_new(device);
  if (!c) return TPPMX_ERR_CUDA;
  if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->sm_count < 1) c->sm_count = 148;
  *out = c;
  return TPPMX_OK;
}

extern "C" int tppmx_destroy(tppmx_ctx *c) {
  if (!c) return TPPMX_OK;
  c->closed = true;
  if (c->live_batches == 0) ctx_release(c);
  return TPPMX_OK;
}

extern "C" int tppmx_set_interrupt_flag(tppmx_ctx *ctx, const volatile int32_t *flag, int32_t check_every) {
  if (!ctx || (flag && check_every < 1)) return TPPMX_ERR_ARG;
  ctx->interrupt = flag;
  if (flag) ctx->check_every = check_every;
  return TPPMX_OK;
}

extern "C" const char *tppmx_last_error(const tppmx_ctx *c) { return c ? c->err.c_str() : "no context"; }

__global__ void probe_math_kernel(int which, const double *x, int n, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = which == 0 ? fast_log(x[i]) : which == 1 ? fast_exp(x[i]) : which == 2 ? lgamma_pos(x[i]) : which == 3 ? fast_log_tab(x[i]) : lgamma_tab(x[i]);
}

extern "C" int tppmx_probe_math(tppmx_ctx *ctx, int32_t which, const double *x, int32_t n, double *out) {
  if (!ctx || !x || !out || n < 0 || which < 0 || which > 4) return TPPMX_ERR_ARG;
  if (n == 0) return TPPMX_OK;
  CK(ctx, cudaSetDevice(ctx->device));
  double *dx = nullptr, *dy = nullptr;
  CK(ctx, cudaMalloc(&dx, sizeof(double) * n));
  CK(ctx, cudaMalloc(&dy, sizeof(double) * n));
  cudaError_t e = cpy_sync(ctx->stream, dx, x, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    probe_math_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(which, dx, n, dy);
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (e == cudaSuccess) e = cpy_sync(ctx->stream, out, dy, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(dx);
  cudaFree(dy);
  if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("probe_math: ") + cudaGetErrorString(e));
  return TPPMX_OK;
}

// ---- allocation helpers -------------------------------------------------------------------
static inline size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }
// state region (zero-initialised by one memset at the end of tppmx_batch_create)
template <class T>
static int dev_alloc(tppmx_batch *b, T **p, size_t n, bool zero = true) {
  (void)zero;
  if (n == 0) n = 1;
  const size_t bytes = al256(n * sizeof(T));
  if (b->dry) {
    b->st_need += bytes;
    *p = nullptr;
    return TPPMX_OK;
  }
  if (b->st_used + bytes > b->st_need) return TPPMX_ERR_NOMEM;
  *p = (T *)(b->ar.dbase + b->up_need + b->st_used);
  b->st_used += bytes;
  return TPPMX_OK;
}
// upload region: staged in the pinned mirror, sent with one copy
template <class T>
static int dev_upload(tppmx_batch *b, const T **p, const std::vector<T> &h) {
  const size_t bytes = al256((h.empty() ? 1 : h.size()) * sizeof(T));
  if (b->dry) {
    b->up_need += bytes;
    *p = nullptr;
    return TPPMX_OK;
  }
  if (b->up_used + bytes > b->up_need) return TPPMX_ERR_NOMEM;
  if (!h.empty()) memcpy(b->ar.hbase + b->up_used, h.data(), h.size() * sizeof(T));
  *p = (const T *)(b->ar.dbase + b->up_used);
  b->up_used += bytes;
  return TPPMX_OK;
}
// upload region, written in place: returns the host pointer inside the pinned mirror (nullptr in
// the sizing pass) and sets the device pointer
template <class T>
static T *dev_stage(tppmx_batch *b, const T **p, size_t n, int *rc) {
  const size_t bytes = al256((n ? n : 1) * sizeof(T));
  if (b->dry) {
    b->up_need += bytes;
    *p = nullptr;
    return nullptr;
  }
  if (b->up_used + bytes > b->up_need) {
    *rc = TPPMX_ERR_NOMEM;
    return nullptr;
  }
  T *h = (T *)(b->ar.hbase + b->up_used);
  *p = (const T *)(b->ar.dbase + b->up_used);
  b->up_used += bytes;
  return h;
}
static bool fast_eligible(const tppmx_batch *b, int lpu);
static bool big_arms(const DevBatch &d);
static bool fast_special(const DevBatch &d);
static int predict_launch(tppmx_batch *b, int n_chains_launch, uint64_t seed, int iter, double *dpi, double *dy, int *dc,
                          cudaStream_t st = nullptr, double *dsum = nullptr, int *launches = nullptr);
// one thread per new patient of an (chain, arm) block: no more threads than patients
static inline int predict_threads(const DevBatch &d) {
  const int t = ((d.npred + 31) / 32) * 32;
  return t < 32 ? 32 : (t > 128 ? 128 : t);
}
static int fast_reserve(tppmx_batch *b);
#define RC(x)            \
  do {                   \
    int rc__ = (x);      \
    if (rc__) return rc__; \
  } while (0)

// per-n constants of the Normal-Normal similarity (sim.cuh), reference utils_ct.cpp:389-400
static std::vector<NNRow> build_nn_tab(const tppmx_model &m, int ntmax) {
  const double m0 = m.similparam[0], s20 = m.similparam[1], v2 = m.similparam[2];
  std::vector<NNRow> t((size_t)ntmax + 2);
  const double sd0 = sqrt(s20);
  const double z0 = m0 / sd0;
  const double ld2 = -(TPPMX_LN_SQRT_2PI + 0.5 * z0 * z0 + log(sd0));
  for (int n = 0; n <= ntmax + 1; n++) {
    const double s2s = 1 / ((n / v2) + (1 / s20));
    const double s2ss = 1 / ((n / v2) + (1 / s2s));
    const double c1 = -0.5 * n * log(2 * M_PI * v2);
    NNRow r;
    r.A0 = c1 + ld2 + TPPMX_LN_SQRT_2PI + 0.5 * log(s2s);
    r.H = 0.5 * s2s;
    r.B0 = c1 - 0.5 * log(s2s) + 0.5 * log(s2ss);
    r.H2 = 0.5 * s2ss;
    t[n] = r;
  }
  return t;
}

extern "C" int tppmx_batch_create(tppmx_ctx *ctx, const tppmx_model *model, const tppmx_dims *dims,
                                  const tppmx_shared *shared, int32_t n_datasets,
                                  const tppmx_dataset *datasets, int32_t n_chains,
                                  const int32_t *chain_dataset, const int32_t *chain_id,
                                  tppmx_batch **out) {
  if (!ctx || ctx->closed) return TPPMX_ERR_ARG;
  if (!model || !dims || !shared || !datasets || !out || n_datasets < 1 || n_chains < 1)
    return fail(ctx, TPPMX_ERR_ARG, "null or empty argument");
  const tppmx_dims &dm = *dims;
  if (dm.nobs < 1 || dm.T < 1 || dm.T > TPPMX_MAXT || dm.D < 1 || dm.D > TPPMX_MAXD || dm.Q < 0 ||
      dm.Q > TPPMX_MAXQ || dm.P < 0 || dm.ncat < 0 || dm.npred < 0 || dm.G < 1 || dm.ntmax < 1)
    return fail(ctx, TPPMX_ERR_ARG, "dims out of range");
  if (model->CC < 1 || model->CC > TPPMX_MAXCC) return fail(ctx, TPPMX_ERR_ARG, "CC out of range");
  if (model->cohesion != 1 && model->cohesion != 2) return fail(ctx, TPPMX_ERR_ARG, "cohesion must be 1|2");
  if (model->cohesion == 2 && !shared->logV) return fail(ctx, TPPMX_ERR_ARG, "cohesion 2 needs logV");
  if (model->PPMx == 1 && (model->consim < 1 || model->consim > 2 || model->similarity < 1 || model->similarity > 2))
    return fail(ctx, TPPMX_ERR_ARG, "consim/similarity must be 1|2");
  if (dm.ncat > 0 && (dm.maxC < 1 || !shared->catvec)) return fail(ctx, TPPMX_ERR_ARG, "catvec/maxC");
  if (dm.ncat > 0 && dm.npred > dm.nobs)
    return fail(ctx, TPPMX_ERR_ARG, "ncat>0 needs npred<=nobs (reference indexes xcat with pp, dm_ppmx_ct.cpp:1157)");
  CK(ctx, cudaSetDevice(ctx->device));

  tppmx_batch *b = new tppmx_batch();
  b->ctx = ctx;
  ctx->live_batches++;
  b->dims = dm;
  if (b->dims.kcap <= 0 || b->dims.kcap > dm.ntmax) b->dims.kcap = dm.ntmax;
  b->last_ms = 0;
  b->last_launches = 0;
  b->scratch_chain = n_chains;
  cudaEventCreate(&b->ev0);
  cudaEventCreate(&b->ev1);
  DevBatch &d = b->d;
  memset(&d, 0, sizeof(d));
  d.model = *model;
  d.nobs = dm.nobs; d.T = dm.T; d.D = dm.D; d.Q = dm.Q; d.P = dm.P; d.ncat = dm.ncat;
  d.maxC = dm.ncat > 0 ? dm.maxC : 0; d.npred = dm.npred; d.G = dm.G; d.ntmax = dm.ntmax;
  d.kcap = b->dims.kcap; d.CC = model->CC; d.n_chains = n_chains; d.n_datasets = n_datasets;
  const int nobs = d.nobs, T = d.T, D = d.D, Q = d.Q, P = d.P, ncat = d.ncat, npred = d.npred;
  const int nc1 = ncat > 0 ? ncat : 1;
  int rc = TPPMX_OK;
#define TRY(x)                 \
  do {                         \
    rc = (x);                  \
    if (rc) goto bad;          \
  } while (0)

  // pass 0 sizes the two regions, pass 1 fills them (see Arena)
  for (int pass = 0; pass < 2; pass++) {
    b->dry = (pass == 0);
    if (pass == 1) {
      Arena &sp = ctx->spare;
      const size_t need = b->up_need + b->st_need;
      if (sp.dcap >= need && sp.hcap >= b->up_need) {
        b->ar = sp;
        sp = Arena();
      } else {
        arena_free(sp);
        if (cudaMalloc((void **)&b->ar.dbase, need) != cudaSuccess) { rc = fail(ctx, TPPMX_ERR_NOMEM, "cudaMalloc of the batch arena failed"); goto bad; }
        b->ar.dcap = need;
        if (cudaHostAlloc((void **)&b->ar.hbase, b->up_need, cudaHostAllocDefault) != cudaSuccess) { rc = fail(ctx, TPPMX_ERR_NOMEM, "cudaHostAlloc of the staging mirror failed"); goto bad; }
        b->ar.hcap = b->up_need;
      }
    }
    // shared
    std::vector<int> catvec(nc1, 0);
    for (int p = 0; p < ncat; p++) catvec[p] = shared->catvec[p];
    b->catvec_host = catvec;
    std::vector<double> grid(shared->grid, shared->grid + 2 * (size_t)d.G);
    TRY(dev_upload(b, &d.catvec, catvec));
    TRY(dev_upload(b, &d.grid, grid));
    if (model->cohesion == 2) {
      std::vector<double> lv(shared->logV, shared->logV + (size_t)T * 2 * d.ntmax * d.G);
      TRY(dev_upload(b, &d.logV, lv));
      // (sigma,kappa) grid draw (update.cuh, :963): lgamma(n - sigma_g) - lgamma(1 - sigma_g) for
      // every grid point g and cluster size n depends on the inputs only -> tabulated once
      std::vector<double> lgt((size_t)d.G * (d.ntmax + 1), 0.0);
      if (!b->dry)
        for (int g = 0; g < d.G; g++) {
          const double sg = shared->grid[2 * (size_t)g], l1 = lgamma(1.0 - sg);
          for (int n = 1; n <= d.ntmax; n++) lgt[(size_t)g * (d.ntmax + 1) + n] = lgamma((double)n - sg) - l1;
        }
      const double *lgp = nullptr;
      TRY(dev_upload(b, &lgp, lgt));
      b->ua.lgtab = lgp;
    }
    std::vector<NNRow> tab = build_nn_tab(*model, d.ntmax);
    TRY(dev_upload(b, &d.nn_tab, tab));

    // datasets: packed straight into the pinned mirror (patient-major rows); the covariate rows
    // xcon / xconp already have that layout and are block-copied
    {
      int src = TPPMX_OK;
      int *h_arm_n = dev_stage(b, &d.arm_n, (size_t)n_datasets * T, &src);
      int *h_arm_pat = dev_stage(b, &d.arm_pat, (size_t)n_datasets * T * d.ntmax, &src);
      int *h_treat = dev_stage(b, &d.treat, (size_t)n_datasets * nobs, &src);
      double *h_x = dev_stage(b, &d.x, (size_t)n_datasets * nobs * P, &src);
      int *h_xcat = dev_stage(b, &d.xcat, (size_t)n_datasets * nobs * nc1, &src);
      double *h_z = dev_stage(b, &d.z, (size_t)n_datasets * nobs * Q, &src);
      double *h_y = dev_stage(b, &d.y, (size_t)n_datasets * nobs * D, &src);
      double *h_zp = dev_stage(b, &d.zpred, (size_t)n_datasets * npred * Q, &src);
      double *h_xp = dev_stage(b, &d.xp, (size_t)n_datasets * npred * P, &src);
      int *h_xcp = dev_stage(b, &d.xcatp, (size_t)n_datasets * npred * nc1, &src);
      if (src) { rc = src; goto bad; }
      if (!b->dry) {
        memset(h_arm_n, 0, sizeof(int) * (size_t)n_datasets * T);
        memset(h_arm_pat, 0, sizeof(int) * (size_t)n_datasets * T * d.ntmax);
        if (ncat == 0) {
          memset(h_xcat, 0, sizeof(int) * (size_t)n_datasets * nobs * nc1);
          memset(h_xcp, 0, sizeof(int) * (size_t)n_datasets * (npred > 0 ? npred : 0) * nc1);
        }
      }
      for (int ds = 0; ds < n_datasets && !b->dry; ds++) {
        const tppmx_dataset &s = datasets[ds];
        if (!s.treat || !s.y || (Q > 0 && !s.z) || (P > 0 && !s.xcon) || (ncat > 0 && !s.xcat)) {
          rc = fail(ctx, TPPMX_ERR_ARG, "dataset has null arrays");
          goto bad;
        }
        if (P > 0) memcpy(h_x + (size_t)ds * nobs * P, s.xcon, sizeof(double) * (size_t)nobs * P);
        int *an = h_arm_n + (size_t)ds * T;
        for (int i = 0; i < nobs; i++) {
          const int t = s.treat[i];
          if (t < 0 || t >= T) { rc = fail(ctx, TPPMX_ERR_ARG, "treat out of range"); goto bad; }
          if (an[t] >= d.ntmax) { rc = fail(ctx, TPPMX_ERR_ARG, "an arm has more than ntmax patients"); goto bad; }
          h_arm_pat[((size_t)ds * T + t) * d.ntmax + an[t]] = i;
          an[t]++;
          h_treat[(size_t)ds * nobs + i] = t;
          for (int p = 0; p < ncat; p++) {
            const int cv = s.xcat[(size_t)i * ncat + p];
            if (cv < 0 || cv >= d.maxC) { rc = fail(ctx, TPPMX_ERR_ARG, "xcat out of range"); goto bad; }
            h_xcat[((size_t)ds * nobs + i) * nc1 + p] = cv;
          }
          for (int q = 0; q < Q; q++) h_z[((size_t)ds * nobs + i) * Q + q] = s.z[i + (size_t)nobs * q];
          for (int k = 0; k < D; k++) h_y[((size_t)ds * nobs + i) * D + k] = s.y[i + (size_t)nobs * k];
        }
        if (npred > 0 && ((Q > 0 && !s.zpred) || (P > 0 && !s.xconp) || (ncat > 0 && !s.xcatp))) {
          rc = fail(ctx, TPPMX_ERR_ARG, "npred>0 but prediction arrays are null");
          goto bad;
        }
        if (npred > 0 && P > 0) memcpy(h_xp + (size_t)ds * npred * P, s.xconp, sizeof(double) * (size_t)npred * P);
        for (int pp = 0; pp < npred; pp++) {
          for (int q = 0; q < Q; q++) h_zp[((size_t)ds * npred + pp) * Q + q] = s.zpred[pp + (size_t)npred * q];
          for (int p = 0; p < ncat; p++) {
            const int cv = s.xcatp[(size_t)pp * ncat + p];
            if (cv < 0 || cv >= d.maxC) { rc = fail(ctx, TPPMX_ERR_ARG, "xcatp out of range"); goto bad; }
            h_xcp[((size_t)ds * npred + pp) * nc1 + p] = cv;
          }
        }
      }
      if (!b->dry) b->arm_n.assign(h_arm_n, h_arm_n + (size_t)n_datasets * T);
    }

    // chains
    const size_t ns = (size_t)n_chains + 1;
    std::vector<int> cds(ns, 0);
    std::vector<uint32_t> cid(ns, 0);
    for (int c = 0; c < n_chains; c++) {
      cds[c] = chain_dataset ? chain_dataset[c] : 0;
      if (cds[c] < 0 || cds[c] >= n_datasets) { rc = fail(ctx, TPPMX_ERR_ARG, "chain_dataset out of range"); goto bad; }
      cid[c] = chain_id ? (uint32_t)chain_id[c] : (uint32_t)c;
    }
    b->chain_ds = cds;
    TRY(dev_upload(b, &d.chain_ds, cds));
    TRY(dev_upload(b, &d.chain_id, cid));
    TRY(dev_alloc(b, &d.labels, ns * st_labels(d)));
    TRY(dev_alloc(b, &d.nj, ns * st_nj(d)));
    TRY(dev_alloc(b, &d.nclu, ns * T));
    TRY(dev_alloc(b, &d.njc, ns * st_njc(d)));
    TRY(dev_alloc(b, &d.idxs, ns));
    TRY(dev_alloc(b, &d.status, ns));
    TRY(dev_alloc(b, &d.sumx, ns * st_sumx(d)));
    TRY(dev_alloc(b, &d.sumx2, ns * st_sumx(d)));
    TRY(dev_alloc(b, &d.eta, ns * st_eta(d)));
    TRY(dev_alloc(b, &d.etae, ns * st_etae(d)));
    TRY(dev_alloc(b, &d.JJ, ns * st_pat(d)));
    TRY(dev_alloc(b, &d.ss, ns * nobs));
    TRY(dev_alloc(b, &d.lg, ns * st_pat(d)));
    TRY(dev_alloc(b, &d.beta, ns * Q * D));
    TRY(dev_alloc(b, &d.sigma, ns * T));
    TRY(dev_alloc(b, &d.kappa, ns * T));
    TRY(dev_alloc(b, &d.Theta, ns * T * D));
    TRY(dev_alloc(b, &d.Sigma, ns * T * D * D));
    TRY(dev_alloc(b, &d.lambda, ns * Q * D));
    TRY(dev_alloc(b, &d.tau, ns));
    TRY(dev_alloc(b, &d.sigma_beta, ns * Q * D));
    TRY(dev_alloc(b, &d.zb, ns * st_pat(d)));
    TRY(dev_alloc(b, &d.ljj, ns * st_pat(d)));
    TRY(dev_alloc(b, &d.cpat, ns * nobs));
    TRY(dev_alloc(b, &d.wbuf, ns * st_wbuf(d)));
    TRY(dev_alloc(b, &d.eta_flag, ns * T * d.kcap));
    TRY(dev_alloc(b, &d.beta_flag, ns * Q * D));
    TRY(dev_alloc(b, &b->ua.eta2, (size_t)n_chains * T * 2 * D * d.kcap));
    TRY(dev_alloc(b, &b->ua.pat2, (size_t)n_chains * 3 * T * d.ntmax));
    TRY(dev_alloc(b, &b->ua.os, (size_t)n_chains * T * d.G));
    TRY(dev_alloc(b, &b->ua.acc, (size_t)n_chains * T * d.kcap));
    // scratch of the call paths, reserved here so that no call allocates
    TRY(dev_alloc(b, &b->init_lab, (size_t)n_datasets * T * d.ntmax));
    TRY(dev_alloc(b, &b->init_ncl, (size_t)n_datasets * T));
    TRY(dev_alloc(b, &b->init_beta, (size_t)Q * D + 1));
    TRY(dev_alloc(b, &b->score_w, (size_t)2 * (d.kcap + d.CC)));
    TRY(dev_alloc(b, &b->score_k, 1));
    if (npred > 0) {
      TRY(dev_alloc(b, &b->pred_pi, (size_t)n_chains * npred * D * T));
      TRY(dev_alloc(b, &b->pred_y, (size_t)n_chains * npred * D * T));
      TRY(dev_alloc(b, &b->pred_c, (size_t)n_chains * npred * T));
    }
    if (fast_eligible(b, 32)) TRY(fast_reserve(b));
    if (pass == 1 && b->fa.lik && fast_special(d) && !b->bail_host && cudaHostAlloc((void **)&b->bail_host, 2 * sizeof(int), cudaHostAllocDefault) == cudaSuccess)
      b->bail_host[0] = b->bail_host[1] = 0;
    if (pass == 1 && big_arms(d) && !b->kmax_host && cudaHostAlloc((void **)&b->kmax_host, sizeof(int), cudaHostAllocDefault) == cudaSuccess)
      *b->kmax_host = 0;
    {  // posterior summaries; the co-clustering counts are skipped beyond 2 GiB (large-n config)
      const size_t npsm = (size_t)n_datasets * T * d.ntmax * d.ntmax;
      TRY(dev_alloc(b, &b->psm_n, (size_t)n_datasets));
      TRY(dev_alloc(b, &b->pi_n, (size_t)n_datasets));
      TRY(dev_alloc(b, &b->util_w, (size_t)TPPMX_MAXD));
      if (npsm * 12 <= ((size_t)2 << 30)) {
        TRY(dev_alloc(b, &b->psm_cnt, npsm));
        TRY(dev_alloc(b, &b->psm_mean, npsm));
      }
      if (npred > 0) {
        TRY(dev_alloc(b, &b->pi_sum, (size_t)n_datasets * npred * D * T));
        TRY(dev_alloc(b, &b->pi_mean, (size_t)n_datasets * npred * D * T));
        TRY(dev_alloc(b, &b->utility, (size_t)n_datasets * npred * T));
        TRY(dev_alloc(b, &b->best_arm, (size_t)n_datasets * npred));
      }
    }
    if (pass == 1) {
      cudaError_t ce = cudaMemcpyAsync(b->ar.dbase, b->ar.hbase, b->up_used, cudaMemcpyHostToDevice, ctx->stream);
      if (ce == cudaSuccess) ce = cudaMemsetAsync(b->ar.dbase + b->up_need, 0, b->st_used, ctx->stream);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
      if (ce != cudaSuccess) {
        rc = fail(ctx, TPPMX_ERR_CUDA, std::string("batch upload: ") + cudaGetErrorString(ce));
        goto bad;
      }
    }
  }
  *out = b;
  return TPPMX_OK;
bad:
  tppmx_batch_destroy(b);
  return rc;
#undef TRY
}

extern "C" int tppmx_batch_destroy(tppmx_batch *b) {
  if (!b) return TPPMX_OK;
  cudaSetDevice(b->ctx->device);
  // nothing of this batch may still be running when its arena is handed on: the main stream orders
  // itself, the side streams are only ever busy here after an error return out of tppmx_batch_run
  cudaStreamSynchronize(b->ctx->side);
  cudaStreamSynchronize(b->ctx->hs);
  b->ctx->hs_pending = false;
  if (b->ar.dbase) {  // leave the arena with the context for the next batch (keep the larger one)
    Arena &sp = b->ctx->spare;
    if (b->ar.dcap > sp.dcap || (b->ar.dcap == sp.dcap && b->ar.hcap > sp.hcap)) {
      arena_free(sp);
      sp = b->ar;
    } else {
      arena_free(b->ar);
    }
  }
  cudaEventDestroy(b->ev0);
  cudaEventDestroy(b->ev1);
  if (b->kmax_host) cudaFreeHost(b->kmax_host);
  if (b->bail_host) cudaFreeHost(b->bail_host);
  tppmx_ctx *c = b->ctx;
  delete b;
  if (--c->live_batches == 0 && c->closed) ctx_release(c);
  return TPPMX_OK;
}

// ---- state initialisation, reference src/dm_ppmx_ct.cpp:49-236 ------------------------------
__global__ void init_kernel(DevBatch b, uint64_t seed, const int *init_labels, const int *init_nclu,
                            const double *initbeta) {
  const int chain = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int ds = b.chain_ds[chain];
  const Rng rng = make_rng(seed, b.chain_id[chain]);
  const int nobs = b.nobs, T = b.T, D = b.D, Q = b.Q, P = b.P, ks = b.kcap, CC = b.CC;
  const int P1 = P > 0 ? P : 1, nc1 = b.ncat > 0 ? b.ncat * b.maxC : 1;
  int *labels = b.labels + (size_t)chain * st_labels(b);
  int *nj = b.nj + (size_t)chain * st_nj(b);
  double *eta = b.eta + (size_t)chain * st_eta(b);
  double *beta = b.beta + (size_t)chain * Q * D;

  // :75-91 JJ = max(y, 1e-100), ss ~ Gamma(1, scale 1/TT)
  for (int i = tid; i < nobs; i += nth) {
    double TT = 0.0;
    for (int k = 0; k < D; k++) {
      double v = b.y[((size_t)ds * nobs + i) * D + k];
      if (v < 1e-100) v = 1e-100;
      b.JJ[((size_t)chain * nobs + i) * D + k] = v;
      TT += v;
    }
    b.ss[(size_t)chain * nobs + i] = rng_gamma(rng, 0, TPPMX_SITE_INIT_SS, 0, (uint32_t)i, 1.0) * (1.0 / TT);
  }
  // :113-115 partition
  for (int e = tid; e < T * b.ntmax; e += nth) labels[e] = init_labels[(size_t)ds * T * b.ntmax + e];
  for (int e = tid; e < T; e += nth) {
    b.nclu[(size_t)chain * T + e] = init_nclu[(size_t)ds * T + e];
    b.sigma[(size_t)chain * T + e] = b.grid[0];  // :52-55
    b.kappa[(size_t)chain * T + e] = b.grid[1];
  }
  if (tid == 0) {
    b.idxs[chain] = 0;  // :49
    b.tau[chain] = 1.0;
    b.status[chain] = 0;
  }
  // :140 eta_star ~ U(0,1); :210 eta_empty ~ U(0,1)
  for (int e = tid; e < T * D * ks; e += nth) {
    const int t = e / (D * ks), k = (e / ks) % D, j = e % ks;
    double u0, u1;
    rng_u2(rng, 0, TPPMX_SITE_INIT_ETA, t, (uint32_t)(j * D + k), 0, u0, u1);
    eta[e] = u0;
  }
  for (int e = tid; e < T * D * CC; e += nth) {
    const int t = e / (D * CC), k = (e / CC) % D, mm = e % CC;
    double u0, u1;
    rng_u2(rng, 0, TPPMX_SITE_INIT_EMPTY, t, (uint32_t)(mm * D + k), 0, u0, u1);
    b.etae[(size_t)chain * st_etae(b) + e] = u0;
  }
  // :149-155 beta, horseshoe scales; :271-278 Theta = 0, Sigma = I
  for (int e = tid; e < Q * D; e += nth) {
    beta[e] = initbeta[e];
    b.sigma_beta[(size_t)chain * Q * D + e] = 1.0;
    b.lambda[(size_t)chain * Q * D + e] = 1.0;
  }
  for (int e = tid; e < T * D; e += nth) b.Theta[(size_t)chain * T * D + e] = 0.0;
  for (int e = tid; e < T * D * D; e += nth) {
    const int a = e % D, c = (e / D) % D;
    b.Sigma[(size_t)chain * T * D * D + e] = (a == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  // cluster sizes and sufficient statistics of the initial partition (:187-204); patients are
  // accumulated in arm order so the sums are bit-identical to a sequential pass.
  for (int e = tid; e < T * ks; e += nth) {
    const int t = e / ks, j = e % ks;
    const int nt = b.arm_n[(size_t)ds * T + t];
    int cnt = 0;
    for (int it = 0; it < nt; it++) cnt += (labels[t * b.ntmax + it] == j + 1);
    nj[e] = cnt;
  }
  for (int e = tid; e < T * P1 * ks; e += nth) {
    double s1 = 0.0, s2 = 0.0;
    if (b.model.PPMx == 1 && P > 0) {
      const int t = e / (P * ks), p = (e / ks) % P, j = e % ks;
      const int nt = b.arm_n[(size_t)ds * T + t];
      for (int it = 0; it < nt; it++)
        if (labels[t * b.ntmax + it] == j + 1) {
          const int i = b.arm_pat[((size_t)ds * T + t) * b.ntmax + it];
          s1 = __dadd_rn(s1, b.x[((size_t)ds * nobs + i) * P + p]);
          {
            const double xv = b.x[((size_t)ds * nobs + i) * P + p];
            s2 = __dadd_rn(s2, __dmul_rn(xv, xv));
          }
        }
    }
    b.sumx[(size_t)chain * st_sumx(b) + e] = s1;
    b.sumx2[(size_t)chain * st_sumx(b) + e] = s2;
  }
  for (int e = tid; e < T * nc1 * ks; e += nth) {
    int cnt = 0;
    if (b.model.PPMx == 1 && b.ncat > 0) {
      const int t = e / (nc1 * ks), pc = (e / ks) % nc1, j = e % ks;
      const int p = pc / b.maxC, c = pc % b.maxC;
      const int nt = b.arm_n[(size_t)ds * T + t];
      for (int it = 0; it < nt; it++)
        if (labels[t * b.ntmax + it] == j + 1) {
          const int i = b.arm_pat[((size_t)ds * T + t) * b.ntmax + it];
          cnt += (b.xcat[((size_t)ds * nobs + i) * b.ncat + p] == c);
        }
    }
    b.njc[(size_t)chain * st_njc(b) + e] = cnt;
  }
  // :161-173 loggamma (same operation order as calculate_gamma, utils_ct.cpp:514-527)
  for (int e = tid; e < nobs * D; e += nth) {
    const int i = e / D, k = e % D;
    const int t = b.treat[(size_t)ds * nobs + i];
    int it = 0;  // position of i in its arm
    const int *pat = b.arm_pat + ((size_t)ds * T + t) * b.ntmax;
    const int nt = b.arm_n[(size_t)ds * T + t];
    for (int r = 0; r < nt; r++)
      if (pat[r] == i) it = r;
    const int lab = labels[t * b.ntmax + it];
    double lg = eta[((size_t)t * D + k) * ks + lab - 1];
    for (int q = 0; q < Q; q++)
      lg = __dadd_rn(lg, __dmul_rn(beta[k + q * D], b.z[((size_t)ds * nobs + i) * Q + q]));
    b.lg[((size_t)chain * nobs + i) * D + k] = lg;
  }
}

extern "C" int tppmx_batch_init(tppmx_batch *b, uint64_t seed, const int32_t *init_labels,
                                const int32_t *init_nclu, const double *initbeta) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  if (!init_labels || !init_nclu || (!initbeta && b->d.Q > 0)) return fail(ctx, TPPMX_ERR_ARG, "null init argument");
  CK(ctx, cudaSetDevice(ctx->device));
  const DevBatch &d = b->d;
  const int nd = d.n_datasets, T = d.T, nt = d.ntmax;
  b->kmax_hint = 0;
  if (b->kmax_host) *b->kmax_host = 0;
  if (b->bail_host) b->bail_host[0] = b->bail_host[1] = 0;
  std::vector<int> lab((size_t)nd * T * nt), ncl((size_t)nd * T);
  for (int ds = 0; ds < nd; ds++)
    for (int t = 0; t < T; t++) {
      const int K = init_nclu[(size_t)ds * T + t];
      if (K < 1 || K > d.kcap) return fail(ctx, TPPMX_ERR_ARG, "init_nclu out of range (1..kcap)");
      if (K > b->kmax_hint) b->kmax_hint = K;
      ncl[(size_t)ds * T + t] = K;
      const int na = b->arm_n[(size_t)ds * T + t];
      for (int it = 0; it < nt; it++) {
        const int v = init_labels[(size_t)ds * T * nt + t + (size_t)T * it];
        if (it < na && (v < 1 || v > K)) return fail(ctx, TPPMX_ERR_ARG, "init_labels out of range");
        lab[((size_t)ds * T + t) * nt + it] = v;
      }
    }
  int *dl = b->init_lab, *dn = b->init_ncl;
  double *db = b->init_beta;
  CK(ctx, cudaMemcpyAsync(dl, lab.data(), lab.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx, cudaMemcpyAsync(dn, ncl.data(), ncl.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (d.Q > 0) CK(ctx, cudaMemcpyAsync(db, initbeta, (size_t)d.Q * d.D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  init_kernel<<<d.n_chains, 128, 0, ctx->stream>>>(d, seed, dl, dn, db);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("init_kernel: ") + cudaGetErrorString(e));
  CK(ctx, cudaGetLastError());
  return TPPMX_OK;
}

// ---- state upload / download (reference layouts <-> device layouts) --------------------------
template <class T>
static int h2d(tppmx_batch *b, T *dst, const std::vector<T> &h) {
  CK(b->ctx, cpy_sync(b->ctx->stream, dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return TPPMX_OK;
}
template <class T>
static int d2h(tppmx_batch *b, std::vector<T> &h, const T *src) {
  CK(b->ctx, cpy_sync(b->ctx->stream, h.data(), src, h.size() * sizeof(T), cudaMemcpyDeviceToHost));
  return TPPMX_OK;
}

// Device -> caller memory.  Caller buffers are pageable, and a pageable cudaMemcpy crawls (~2 GB/s)
// and holds driver locks other host threads need; large results therefore land in the batch's
// pinned mirror first (free once the upload of tppmx_batch_create has been consumed, which
// stream order guarantees) and are copied out by the host.
static cudaError_t d2h_out(tppmx_batch *b, void *dst, const void *src, size_t bytes) {
  cudaStream_t st = b->ctx->stream;
  if (bytes < (64u << 10) || bytes > b->ar.hcap || !b->ar.hbase) return cpy_sync(st, dst, src, bytes, cudaMemcpyDeviceToHost);
  const cudaError_t e = cpy_sync(st, b->ar.hbase, src, bytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) memcpy(dst, b->ar.hbase, bytes);
  return e;
}

extern "C" int tppmx_batch_set_state(tppmx_batch *b, int32_t chain, const tppmx_state *s) {
  if (!b || !s) return TPPMX_ERR_ARG;
  const DevBatch &d = b->d;
  if (chain < 0 || chain >= d.n_chains) return fail(b->ctx, TPPMX_ERR_ARG, "chain out of range");
  CK(b->ctx, cudaSetDevice(b->ctx->device));
  const int T = d.T, nt = d.ntmax, ks = d.kcap, P = d.P, D = d.D, Q = d.Q, n = d.nobs, CC = d.CC;
  const int ncat = d.ncat, mC = d.maxC;
  if (s->nclu)
    for (int t = 0; t < T; t++)
      if (s->nclu[t] < 0 || s->nclu[t] > ks) return fail(b->ctx, TPPMX_ERR_CAPACITY, "nclu exceeds kcap");
  // the kernels index eta / sumx / nj with label-1 and logV / lgtab with idxs: validate before the upload
  const int ds_of_chain = b->chain_ds[chain];
  if (s->labels) {
    std::vector<int> kcur(T, 0);
    if (s->nclu) {
      for (int t = 0; t < T; t++) kcur[t] = s->nclu[t];
    } else {
      CK(b->ctx, cpy_sync(b->ctx->stream, kcur.data(), d.nclu + (size_t)chain * T, sizeof(int) * T, cudaMemcpyDeviceToHost));
    }
    for (int t = 0; t < T; t++) {
      const int na = b->arm_n[(size_t)ds_of_chain * T + t];
      for (int it = 0; it < na; it++) {
        const int v = s->labels[t + (size_t)T * it];
        if (v < 1 || v > kcur[t]) return fail(b->ctx, TPPMX_ERR_ARG, "set_state: label outside 1..nclu[t]");
      }
    }
  }
  if (s->nj)
    for (int t = 0; t < T; t++)
      for (int j = 0; j < ks; j++)
        if (s->nj[t + (size_t)T * j] < 0) return fail(b->ctx, TPPMX_ERR_ARG, "set_state: negative cluster size");
  if (s->idxs && (s->idxs[0] < 0 || s->idxs[0] >= d.G)) return fail(b->ctx, TPPMX_ERR_ARG, "set_state: idxs outside 0..G-1");
  if (s->labels) {
    std::vector<int> h(st_labels(d));
    for (int t = 0; t < T; t++)
      for (int it = 0; it < nt; it++) h[(size_t)t * nt + it] = s->labels[t + (size_t)T * it];
    RC(h2d(b, d.labels + (size_t)chain * st_labels(d), h));
  }
  if (s->nj) {
    std::vector<int> h(st_nj(d));
    for (int t = 0; t < T; t++)
      for (int j = 0; j < ks; j++) h[(size_t)t * ks + j] = s->nj[t + (size_t)T * j];
    RC(h2d(b, d.nj + (size_t)chain * st_nj(d), h));
  }
  if (s->nclu) {
    std::vector<int> h(s->nclu, s->nclu + T);
    RC(h2d(b, d.nclu + (size_t)chain * T, h));
  }
  for (int which = 0; which < 2; which++) {
    const double *src = which ? s->sumx2 : s->sumx;
    if (!src || P == 0) continue;
    std::vector<double> h(st_sumx(d));
    for (int t = 0; t < T; t++)
      for (int p = 0; p < P; p++)
        for (int j = 0; j < ks; j++) h[((size_t)t * P + p) * ks + j] = src[t + (size_t)T * ((size_t)j * P + p)];
    RC(h2d(b, (which ? d.sumx2 : d.sumx) + (size_t)chain * st_sumx(d), h));
  }
  if (s->njc && ncat > 0) {
    std::vector<int> h(st_njc(d));
    for (int t = 0; t < T; t++)
      for (int p = 0; p < ncat; p++)
        for (int c = 0; c < mC; c++)
          for (int j = 0; j < ks; j++)
            h[(((size_t)t * ncat + p) * mC + c) * ks + j] = s->njc[t + (size_t)T * (((size_t)j * ncat + p) * mC + c)];
    RC(h2d(b, d.njc + (size_t)chain * st_njc(d), h));
  }
  if (s->eta_star) {
    std::vector<double> h(st_eta(d));
    for (int t = 0; t < T; t++)
      for (int k = 0; k < D; k++)
        for (int j = 0; j < ks; j++) h[((size_t)t * D + k) * ks + j] = s->eta_star[k + (size_t)D * (j + (size_t)nt * t)];
    RC(h2d(b, d.eta + (size_t)chain * st_eta(d), h));
  }
  if (s->eta_empty) {
    std::vector<double> h(st_etae(d));
    for (int t = 0; t < T; t++)
      for (int k = 0; k < D; k++)
        for (int m = 0; m < CC; m++) h[((size_t)t * D + k) * CC + m] = s->eta_empty[k + (size_t)D * (m + (size_t)CC * t)];
    RC(h2d(b, d.etae + (size_t)chain * st_etae(d), h));
  }
  for (int which = 0; which < 2; which++) {
    const double *src = which ? s->loggamma : s->JJ;
    if (!src) continue;
    std::vector<double> h(st_pat(d));
    for (int i = 0; i < n; i++)
      for (int k = 0; k < D; k++) h[(size_t)i * D + k] = src[i + (size_t)n * k];
    RC(h2d(b, (which ? d.lg : d.JJ) + (size_t)chain * st_pat(d), h));
  }
#define PUT(field, dptr, count)                                                        \
  if (s->field && (count) > 0) {                                                       \
    CK(b->ctx, cpy_sync(b->ctx->stream, (dptr) + (size_t)chain * (count), s->field, sizeof(*s->field) * (count), cudaMemcpyHostToDevice)); \
  }
  PUT(ss, d.ss, (size_t)n);
  PUT(beta, d.beta, (size_t)Q * D);
  PUT(sigma, d.sigma, (size_t)T);
  PUT(kappa, d.kappa, (size_t)T);
  PUT(idxs, d.idxs, (size_t)1);
  PUT(Theta, d.Theta, (size_t)T * D);
  PUT(Sigma, d.Sigma, (size_t)T * D * D);
  PUT(lambda, d.lambda, (size_t)Q * D);
  PUT(tau, d.tau, (size_t)1);
  PUT(sigma_beta, d.sigma_beta, (size_t)Q * D);
#undef PUT
  return TPPMX_OK;
}

extern "C" int tppmx_batch_get_state(tppmx_batch *b, int32_t chain, tppmx_state *s) {
  if (!b || !s) return TPPMX_ERR_ARG;
  const DevBatch &d = b->d;
  if (chain < 0 || chain >= d.n_chains) return fail(b->ctx, TPPMX_ERR_ARG, "chain out of range");
  CK(b->ctx, cudaSetDevice(b->ctx->device));
  CK(b->ctx, cudaStreamSynchronize(b->ctx->stream));
  const int T = d.T, nt = d.ntmax, ks = d.kcap, P = d.P, D = d.D, Q = d.Q, n = d.nobs, CC = d.CC;
  const int ncat = d.ncat, mC = d.maxC;
  if (s->labels) {
    std::vector<int> h(st_labels(d));
    RC(d2h(b, h, d.labels + (size_t)chain * st_labels(d)));
    for (int t = 0; t < T; t++)
      for (int it = 0; it < nt; it++) s->labels[t + (size_t)T * it] = h[(size_t)t * nt + it];
  }
  if (s->nj) {
    std::vector<int> h(st_nj(d));
    RC(d2h(b, h, d.nj + (size_t)chain * st_nj(d)));
    for (int t = 0; t < T; t++)
      for (int j = 0; j < nt; j++) s->nj[t + (size_t)T * j] = j < ks ? h[(size_t)t * ks + j] : 0;
  }
  if (s->nclu) CK(b->ctx, cpy_sync(b->ctx->stream, s->nclu, d.nclu + (size_t)chain * T, sizeof(int) * T, cudaMemcpyDeviceToHost));
  for (int which = 0; which < 2; which++) {
    double *dst = which ? s->sumx2 : s->sumx;
    if (!dst || P == 0) continue;
    std::vector<double> h(st_sumx(d));
    RC(d2h(b, h, (which ? d.sumx2 : d.sumx) + (size_t)chain * st_sumx(d)));
    for (int t = 0; t < T; t++)
      for (int p = 0; p < P; p++)
        for (int j = 0; j < nt; j++)
          dst[t + (size_t)T * ((size_t)j * P + p)] = j < ks ? h[((size_t)t * P + p) * ks + j] : 0.0;
  }
  if (s->njc && ncat > 0) {
    std::vector<int> h(st_njc(d));
    RC(d2h(b, h, d.njc + (size_t)chain * st_njc(d)));
    for (int t = 0; t < T; t++)
      for (int p = 0; p < ncat; p++)
        for (int c = 0; c < mC; c++)
          for (int j = 0; j < nt; j++)
            s->njc[t + (size_t)T * (((size_t)j * ncat + p) * mC + c)] =
                j < ks ? h[(((size_t)t * ncat + p) * mC + c) * ks + j] : 0;
  }
  if (s->eta_star) {
    std::vector<double> h(st_eta(d));
    RC(d2h(b, h, d.eta + (size_t)chain * st_eta(d)));
    for (int t = 0; t < T; t++)
      for (int k = 0; k < D; k++)
        for (int j = 0; j < nt; j++)
          s->eta_star[k + (size_t)D * (j + (size_t)nt * t)] = j < ks ? h[((size_t)t * D + k) * ks + j] : 0.0;
  }
  if (s->eta_empty) {
    std::vector<double> h(st_etae(d));
    RC(d2h(b, h, d.etae + (size_t)chain * st_etae(d)));
    for (int t = 0; t < T; t++)
      for (int k = 0; k < D; k++)
        for (int m = 0; m < CC; m++) s->eta_empty[k + (size_t)D * (m + (size_t)CC * t)] = h[((size_t)t * D + k) * CC + m];
  }
  for (int which = 0; which < 2; which++) {
    double *dst = which ? s->loggamma : s->JJ;
    if (!dst) continue;
    std::vector<double> h(st_pat(d));
    RC(d2h(b, h, (which ? d.lg : d.JJ) + (size_t)chain * st_pat(d)));
    for (int i = 0; i < n; i++)
      for (int k = 0; k < D; k++) dst[i + (size_t)n * k] = h[(size_t)i * D + k];
  }
#define GET(field, dptr, count)                                                        \
  if (s->field && (count) > 0) {                                                       \
    CK(b->ctx, cpy_sync(b->ctx->stream, s->field, (dptr) + (size_t)chain * (count), sizeof(*s->field) * (count), cudaMemcpyDeviceToHost)); \
  }
  GET(ss, d.ss, (size_t)n);
  GET(beta, d.beta, (size_t)Q * D);
  GET(sigma, d.sigma, (size_t)T);
  GET(kappa, d.kappa, (size_t)T);
  GET(idxs, d.idxs, (size_t)1);
  GET(Theta, d.Theta, (size_t)T * D);
  GET(Sigma, d.Sigma, (size_t)T * D * D);
  GET(lambda, d.lambda, (size_t)Q * D);
  GET(tau, d.tau, (size_t)1);
  GET(sigma_beta, d.sigma_beta, (size_t)Q * D);
#undef GET
  return TPPMX_OK;
}

extern "C" int tppmx_batch_get_partitions(tppmx_batch *b, int32_t *labels, int32_t *nclu) {
  if (!b) return TPPMX_ERR_ARG;
  const DevBatch &d = b->d;
  CK(b->ctx, cudaSetDevice(b->ctx->device));
  CK(b->ctx, cudaStreamSynchronize(b->ctx->stream));
  const int T = d.T, nt = d.ntmax;
  if (labels) {
    const size_t nlab = (size_t)d.n_chains * st_labels(d);
    std::vector<int> hv;
    const int *h;
    if (nlab * sizeof(int) <= b->ar.hcap && b->ar.hbase) {  // through the pinned mirror, transposed from there
      CK(b->ctx, cpy_sync(b->ctx->stream, b->ar.hbase, d.labels, nlab * sizeof(int), cudaMemcpyDeviceToHost));
      h = reinterpret_cast<const int *>(b->ar.hbase);
    } else {
      hv.resize(nlab);
      RC(d2h(b, hv, d.labels));
      h = hv.data();
    }
    for (int c = 0; c < d.n_chains; c++)
      for (int t = 0; t < T; t++)
        for (int it = 0; it < nt; it++)
          labels[(size_t)c * T * nt + t + (size_t)T * it] = h[((size_t)c * T + t) * nt + it];
  }
  if (nclu) CK(b->ctx, cpy_sync(b->ctx->stream, nclu, d.nclu, sizeof(int) * (size_t)d.n_chains * T, cudaMemcpyDeviceToHost));
  return TPPMX_OK;
}

// ---- launches ---------------------------------------------------------------------------------
static int check_status(tppmx_batch *b) {
  std::vector<int> st((size_t)b->d.n_chains);
  CK(b->ctx, cpy_sync(b->ctx->stream, st.data(), b->d.status, st.size() * sizeof(int), cudaMemcpyDeviceToHost));
  for (size_t c = 0; c < st.size(); c++)
    if (st[c] != 0) {
      char buf[160];
      snprintf(buf, sizeof buf, "chain %zu needed more than kcap=%d clusters in one arm", c, b->d.kcap);
      return fail(b->ctx, st[c], buf);
    }
  return TPPMX_OK;
}

extern "C" int tppmx_batch_set_option(tppmx_batch *b, int32_t opt, int32_t value) {
  if (!b) return TPPMX_ERR_ARG;
  if (opt == TPPMX_OPT_SWEEP_IMPL && value >= 0 && value <= 2) {
    b->opt_impl = value;
    return TPPMX_OK;
  }
  if (opt == TPPMX_OPT_LANES_PER_UNIT && (value == 0 || value == 8 || value == 16 || value == 32)) {
    b->opt_lpu = value;
    b->fast_ready = false;  // the staged capacity may depend on it (calibration 1)
    return TPPMX_OK;
  }
  if (opt == TPPMX_OPT_FAST_STAGED_CLUSTERS && value >= 0 && value <= TPPMX_BIG_KSMAX && !b->fast_ready) {
    b->opt_ks = value;
    return TPPMX_OK;
  }
  if (opt == TPPMX_OPT_CHAIN_GROUPS && value >= 0 && value <= TPPMX_MAX_GROUPS) {
    b->opt_groups = value;
    return TPPMX_OK;
  }
  if (opt == TPPMX_OPT_BIG_THREADS && (value == 0 || value == 128 || value == 256)) {
    b->opt_big_nth = value;
    return TPPMX_OK;
  }
  return fail(b->ctx, TPPMX_ERR_ARG, "unknown option or value");
}

extern "C" int tppmx_batch_get_info(const tppmx_batch *b, int32_t what, int32_t *value) {
  if (!b || !value) return TPPMX_ERR_ARG;
  if (what == TPPMX_INFO_LAST_SWEEP_IMPL) *value = b->last_impl;
  else if (what == TPPMX_INFO_LAST_HANDOVERS) *value = b->last_handovers;
  else if (what == TPPMX_INFO_LAST_CHAIN_GROUPS) *value = b->last_groups;
  else return TPPMX_ERR_ARG;
  return TPPMX_OK;
}

// models with their own instantiation of the fast kernel (staged capacity lanes - CC)
static bool fast_special(const DevBatch &d) {
  const tppmx_model &m = d.model;
  return m.calibration == 1 || m.consim == 2 || d.ncat > 0 || m.similarity == 2;
}
// is the model one the fast kernel implements?
static bool fast_eligible(const tppmx_batch *b, int lpu) {
  const DevBatch &d = b->d;
  const tppmx_model &m = d.model;
  // calibration 1 and the NNIG similarity have their own instantiations of sweep_fast.cuh: the package's case only
  // (D <= 3, short arms, 16 | 32 lanes per unit)
  const bool special = fast_special(d);
  if (m.similarity == 2 && (m.calibration == 1 || m.consim != 1 || d.ncat > 0)) return false;
  if (special && (d.D > 3 || m.nolik || big_arms(d) || lpu < 16)) return false;
  if (m.consim == 2 && (m.calibration == 1 || d.P > 32)) return false;
  if (d.ncat > 0 && (m.calibration == 1 || m.consim != 1 || d.ncat > 16 || d.ncat * d.maxC > 256)) return false;
  if (m.PPMx == 0 && (special || big_arms(d))) return false;  // no similarity term: the plain instantiation with a zero factor
  return (m.PPMx == 1 || m.PPMx == 0) && (m.consim == 1 || m.consim == 2) && (m.similarity == 1 || m.similarity == 2) && m.calibration >= 0 && m.calibration <= 2 &&
         (m.reuse == 1 || (m.reuse == 0 && !special && m.PPMx == 1 && d.D <= 3 && !m.nolik && !big_arms(d))) && d.P >= 1 && d.P <= 4 * lpu;
}
// arms too long for the small-arm kernel's shared-memory label / order / cohesion tables (thousands of
// patients, BASELINE config 4) run the long-arm kernel (sweep_big.cuh): one CTA per (chain, arm)
static bool big_arms(const DevBatch &d) {
  const size_t ks = d.kcap < TPPMX_FAST_KSMAX ? d.kcap : TPPMX_FAST_KSMAX, ntp = (d.ntmax + 3) & ~3;
  return fast_tab_bytes(d.ntmax, 0, d.model.consim == 2) + 2 * fast_unit_bytes(d.P, (int)ks, (int)ntp, d.CC, d.D, 0, d.ncat, d.maxC) > 48 * 1024;
}
// largest staged cluster capacity of the long-arm kernel: min(kcap, 256), lowered until one CTA fits an SM
static int big_ks(const DevBatch &d, int opt_ks) {
  int ks = d.kcap < TPPMX_BIG_KSMAX ? d.kcap : TPPMX_BIG_KSMAX;
  if (opt_ks > 0 && opt_ks < ks) ks = opt_ks;
  while (ks > 8 && big_smem_bytes(d.P, ks, d.CC, d.D, 0) > 200 * 1024) ks -= 8;
  return ks;
}

// scratch of the fast sweep kernel, sized for the largest staged capacity (part of the arena)
static int fast_reserve(tppmx_batch *b) {
  const DevBatch &d = b->d;
  FastArgs &fa = b->fa;
  const size_t KS = big_arms(d) ? (size_t)big_ks(d, 0) : (d.kcap < TPPMX_FAST_KSMAX ? d.kcap : TPPMX_FAST_KSMAX);
  const size_t ntp = (d.ntmax + 3) & ~3, units = (size_t)d.n_chains * d.T;
  RC(dev_alloc(b, &fa.lik, units * (KS + d.CC) * ntp + ntp + 8));
  RC(dev_alloc(b, &fa.uq, units * 4 * (ntp + 1) + 8));
  RC(dev_alloc(b, &fa.resume_l, units));
  RC(dev_alloc(b, &fa.resume_it, units));
  RC(dev_alloc(b, &fa.bail_count, TPPMX_MAX_GROUPS));  // one counter per chain group of tppmx_batch_run
  RC(dev_alloc(b, &fa.kmax_dev, TPPMX_MAX_GROUPS));
  fa.lik_unit_stride = (KS + d.CC) * ntp;
  RC(dev_alloc(b, &fa.logng, units * (ntp + 1)));
  RC(dev_alloc(b, &fa.patpad, units * (ntp + 1)));
  {  // the tables over n of sweep_fast.cuh (also read from global memory by the long-arm kernels)
    std::vector<double> tg;
    const double dP = (double)d.P;
    if (d.model.consim == 2) {
      // Normal-Normal-Inverse-Gamma similarity (utils_ct.cpp:427-459 with dN_IG :181-200, dinvgamma :161-176):
      // everything of gsimconNNIG(sumx, sumx2, n) that depends on n only, formed with the reference's own
      // operations (the kernel mirrors its evaluation order: the function cancels terms of order 500 n, so
      // any other association differs from it by ~1e-10 on a log-weight).  Row n (TPPMX_NNIG_ROW doubles):
      //   0 An = (-0.5 n) log(2 pi v)   1 n   2 1/n   3 k0 + n   4 RN(1/(k0+n))   5 (0.5 n) k0   6 sd = sqrt(v/(k0+n))
      //   7 RN(1/sd)   8 log sd   9 a0s = nu0/2 + 0.5 n   10 lgamma(a0s)   11 (a0s + 1) log v   12 (n mu) mu
      // with the function's own evaluation point mu = 10, v = 0.1 (:431-432).
      //   13 ld2 = dN_IG(mu, v; m0, k0, nu0/2, nu0 s20/2) (:451, the same in every row)
      const double mu = 10.0, v = 0.1, m0 = d.model.similparam[0], s20 = d.model.similparam[1], k0 = d.model.similparam[3],
                   nu0 = d.model.similparam[4];
      const double a0 = 0.5 * nu0, b0 = 0.5 * nu0 * s20;
      double ld2;
      {
        const double sd0 = sqrt(v / k0), t = (mu - m0) / sd0;
        const double dnl = -(0.918938533204672741780329736406 + 0.5 * t * t + log(sd0));
        const double ig = a0 * log(b0) - lgamma(a0) - (a0 + 1) * log(v) - (b0 / v);
        ld2 = dnl + ig;
      }
      tg.assign((size_t)TPPMX_NNIG_ROW * (d.ntmax + 2), 0.0);
      for (int n = 0; n <= d.ntmax + 1; n++) {
        const double a0s = a0 + 0.5 * n, kn = k0 + n, sd = sqrt(v / kn);
        double *r = &tg[(size_t)TPPMX_NNIG_ROW * n];
        r[0] = -0.5 * n * log(2 * M_PI * v);
        r[1] = (double)n;
        r[2] = n > 0 ? 1 / (double)n : 0.0;
        r[3] = kn;
        r[4] = 1.0 / kn;
        r[5] = 0.5 * n * k0;
        r[6] = sd;
        r[7] = 1.0 / sd;
        r[8] = log(sd);
        r[9] = a0s;
        r[10] = lgamma(a0s);
        r[11] = (a0s + 1) * log(v);
        r[12] = n * mu * mu;
        r[13] = ld2;
      }
    } else {
      std::vector<NNRow> tab = build_nn_tab(d.model, d.ntmax);
      tg.assign((size_t)4 * (d.ntmax + 2), 0.0);
      for (int n = 0; n <= d.ntmax; n++) {
        tg[(size_t)4 * n + 0] = (n == 0) ? dP * tab[1].A0 : dP * (tab[n + 1].A0 - tab[n].A0);
        tg[(size_t)4 * n + 1] = (n == 0) ? 0.0 : tab[n].H;
        tg[(size_t)4 * n + 2] = tab[n + 1].H;
        tg[(size_t)4 * n + 3] = (n == 0) ? 0.0 : dP * tab[n].A0;  // calibration 1 scores lgN, lgY themselves
        if (d.model.similarity == 2) {  // double dipper: (dB, HN, HY, H2N)
          tg[(size_t)4 * n + 0] = (n == 0) ? dP * tab[1].B0 : dP * (tab[n + 1].B0 - tab[n].B0);
          tg[(size_t)4 * n + 3] = (n == 0) ? 0.0 : tab[n].H2;
        }
      }
    }
    const double *tp = nullptr;
    RC(dev_upload(b, &tp, tg));
    fa.tabg = const_cast<double *>(tp);
    if (d.model.consim == 2) b->d.nnig_tab = tp;  // the generic kernels and the predictive step use it too (sim.cuh)
  }
  fa.cat_lw = fa.cat_ln = nullptr;
  if (d.ncat > 0) {  // gsimcatDM (utils_ct.cpp:468-495) as two logarithms over the integers (sweep_fast.cuh, CAT)
    const double w = d.model.similparam[6];
    std::vector<double> lw((size_t)d.ntmax + 2), ln((size_t)d.ncat * (d.ntmax + 2));
    for (int m = 0; m <= d.ntmax + 1; m++) lw[m] = log((double)m + w);
    for (int p = 0; p < d.ncat; p++)
      for (int n = 0; n <= d.ntmax + 1; n++) ln[(size_t)p * (d.ntmax + 2) + n] = log((double)n + (double)b->catvec_host[p] * w);
    RC(dev_upload(b, &fa.cat_lw, lw));
    RC(dev_upload(b, &fa.cat_ln, ln));
  }
  return TPPMX_OK;
}

static int fast_prepare(tppmx_batch *b, int lpu, size_t *smem_out) {
  DevBatch &d = b->d;
  FastArgs &fa = b->fa;
  if (!fa.lik) return fail(b->ctx, TPPMX_ERR_ARG, "fast sweep scratch was not reserved for this batch");
  if (!b->fast_ready) {
    fa.big = big_arms(d) ? 1 : 0;
    if (fa.big) {
      b->big_ksmax = big_ks(d, 0);
      fa.KS = b->big_ksmax;
    } else {
      fa.KS = d.kcap < TPPMX_FAST_KSMAX ? d.kcap : TPPMX_FAST_KSMAX;
      if (b->opt_ks > 0 && b->opt_ks < fa.KS) fa.KS = b->opt_ks;
      // calibration 1 normalises over all candidates in one pass of the unit's lanes (and the NNIG instantiation
      // has no chunked scoring path): K + CC <= lanes
      if ((d.model.calibration == 1 || d.model.consim == 2 || d.ncat > 0 || d.model.similarity == 2) && fa.KS > lpu - d.CC) fa.KS = lpu - d.CC;
    }
    fa.ntp = (d.ntmax + 3) & ~3;
    fa.ncol = fa.KS + d.CC;
    fa.n_units = d.n_chains * d.T;
    fa.unit_bytes = (int)fast_unit_bytes(d.P, fa.KS, fa.ntp, d.CC, d.D, 0, d.ncat, d.maxC);
    fa.tab_bytes = (int)fast_tab_bytes(d.ntmax, 0, d.model.consim == 2);
    b->fast_ready = true;
  }
  if (fa.big) {
    // staged capacity of THIS launch of the long-arm kernel: room for 1.25 x + 8 the largest K any unit has
    // reached so far (read back from the previous launch without a sync; the initial partition before
    // that), so that the shared-memory footprint -- and with it the number of resident units -- follows the
    // partition.  A unit that outgrows it hands over to the generic kernel: the choice affects speed only.
    int hint = b->kmax_hint;
    if (b->kmax_host && *(volatile int *)b->kmax_host > 0) hint = *(volatile int *)b->kmax_host;
    b->big_hint = hint;
    // two stages while the partitions are small (K + 4 <= 32 everywhere): the one-warp-per-unit kernel
    // (sweep_fast_kernel, big-arm mode, 8 units resident per SM) runs every unit, and only the units that reach
    // 32 clusters are finished by the CTA-per-unit kernel; larger partitions go to the latter directly
    b->big_two_stage = b->opt_ks == 0 && hint > 0 && hint + 4 <= TPPMX_FAST_KSMAX && d.P <= 128;
    int ks = b->big_ksmax;
    if (b->opt_ks > 0) ks = b->opt_ks < ks ? b->opt_ks : ks;
    else if (hint > 0) {
      int want = hint + hint / 4 + 8;
      if (b->big_two_stage && want < 64) want = 64;  // second stage: units that have just outgrown 32 clusters
      want = (want + 7) & ~7;
      if (want < 16) want = 16;
      if (want < ks) ks = want;
    }
    fa.KS = ks;
    fa.ncol = ks + d.CC;
    b->big_nth = b->opt_big_nth ? b->opt_big_nth : (ks <= 128 ? 128 : 256);
    *smem_out = big_smem_bytes(d.P, ks, d.CC, d.D, b->big_nth);
  } else {
    *smem_out = (size_t)fa.tab_bytes + (size_t)(32 / lpu) * fa.unit_bytes;
  }
  return TPPMX_OK;
}

// Launches the sweep kernels for iterations iter0..iter0+n_sweeps-1 on the context's stream.
// sync_handover: read the hand-over count back and launch the generic finisher only if needed
// (one host sync); otherwise the finisher is always launched (it exits at once for chains with
// nothing to resume), so a whole run needs no host round trip.
static int sweep_launch(tppmx_batch *b, uint64_t seed, int32_t iter0, int32_t n_sweeps, bool sync_handover, int *launches) {
  tppmx_ctx *ctx = b->ctx;
  // lanes per (chain, arm): 16 (two units per warp) for the package's shapes, where an arm holds
  // ~4-10 clusters; long arms (config 4) hold tens of clusters -> one unit per warp
  int lpu = b->opt_lpu;
  const bool big = big_arms(b->d);
  if (big) lpu = 32;  // long arms: one CTA per (chain, arm), P <= 128 covariate owners
  // One unit per warp (32 lanes) puts twice the lanes on the sweep's parallel pre-pass; two units per warp (16 lanes)
  // halve the instruction stream per unit.  While every unit still gets a warp of its own within one wave of the
  // machine (12 one-warp CTAs per SM; measured on B200: 600 / 1200 units +30 % / +15 % with 32 lanes, 1800 units
  // -22 %) latency is what counts: 32 lanes up to 10 units per SM, 16 beyond.
  else if (!lpu) {
    lpu = ((long long)b->d.n_chains * b->d.T <= 10LL * ctx->sm_count) ? 32 : 16;
    if (b->bail_host) {
      const volatile int *bh = b->bail_host;
      if ((long long)bh[0] * 50 > (long long)b->d.n_chains * b->d.T || bh[1] + b->d.CC + 2 > 16) lpu = 32;
    }
  }
  if (lpu != b->last_lpu) {  // the staged capacity of the special instantiations follows the lanes
    b->fast_ready = false;
    b->last_lpu = lpu;
  }
  size_t smem = 0;
  bool use_fast = b->opt_impl != 1 && fast_eligible(b, lpu) && b->fa.lik;
  if (use_fast) {
    RC(fast_prepare(b, lpu, &smem));
    if (smem > 200 * 1024) use_fast = false;  // would not fit an SM: generic kernel
  }
  if (b->opt_impl == 2 && !use_fast) return fail(ctx, TPPMX_ERR_ARG, "model/shape not eligible for the fast sweep kernel");
  if (use_fast) {
    CK(ctx, cudaMemsetAsync(b->fa.bail_count, 0, sizeof(int), ctx->stream));
    if (!b->fa.big && b->bail_host) CK(ctx, cudaMemsetAsync(b->fa.kmax_dev, 0, sizeof(int), ctx->stream));
    cudaError_t e;
    if (b->fa.big) {
      CK(ctx, cudaMemsetAsync(b->fa.kmax_dev, 0, sizeof(int), ctx->stream));
      FastArgs f2 = b->fa;
      f2.resume_mode = 0;
      e = cudaSuccess;
      if (b->big_two_stage) {  // first stage: every unit, up to 32 clusters, one warp per unit
        FastArgs f1 = b->fa;
        f1.KS = b->d.kcap < TPPMX_FAST_KSMAX ? b->d.kcap : TPPMX_FAST_KSMAX;
        f1.ncol = f1.KS + b->d.CC;
        f1.unit_bytes = (int)fast_unit_bytes(b->d.P, f1.KS, f1.ntp, b->d.CC, b->d.D, 1);
        f1.tab_bytes = (int)fast_tab_bytes(b->d.ntmax, 1);
        e = fast_launch_lpu32(b->d, f1, (size_t)f1.tab_bytes + f1.unit_bytes, seed, iter0, n_sweeps, ctx->stream);
        (*launches)++;
        CK(ctx, cudaMemsetAsync(b->fa.bail_count, 0, sizeof(int), ctx->stream));  // its hand-overs are resumed next
        f2.resume_mode = 1;
      }
      if (e == cudaSuccess) e = big_launch(b->d, f2, b->big_nth, smem, seed, iter0, n_sweeps, ctx->stream);
    } else {
      e = lpu == 8    ? fast_launch_lpu8(b->d, b->fa, smem, seed, iter0, n_sweeps, ctx->stream)
          : lpu == 32 ? fast_launch_lpu32(b->d, b->fa, smem, seed, iter0, n_sweeps, ctx->stream)
                      : fast_launch_lpu16(b->d, b->fa, smem, seed, iter0, n_sweeps, ctx->stream);
    }
    if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("sweep_fast_kernel: ") + cudaGetErrorString(e));
    (*launches)++;
    if (b->fa.big && b->kmax_host)
      CK(ctx, cudaMemcpyAsync(b->kmax_host, b->fa.kmax_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (!b->fa.big && b->bail_host) {
      CK(ctx, cudaMemcpyAsync(b->bail_host, b->fa.bail_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CK(ctx, cudaMemcpyAsync(b->bail_host + 1, b->fa.kmax_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    }
    int nb = 1;
    if (sync_handover) {
      CK(ctx, cudaMemcpyAsync(&nb, b->fa.bail_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CK(ctx, cudaStreamSynchronize(ctx->stream));
      b->last_handovers = nb;
    }
    if (nb > 0) {  // finish the units that outgrew the staged capacity, from their exact position
      sweep_kernel<<<b->d.n_chains, 32 * b->d.T, 0, ctx->stream>>>(b->d, seed, iter0, n_sweeps, b->fa.resume_l,
                                                                 b->fa.resume_it);
      (*launches)++;
    }
    b->last_impl = 2;
  } else {
    sweep_kernel<<<b->d.n_chains, 32 * b->d.T, 0, ctx->stream>>>(b->d, seed, iter0, n_sweeps, nullptr, nullptr);
    (*launches)++;
    b->last_impl = 1;
  }
  CK(ctx, cudaGetLastError());
  return TPPMX_OK;
}

extern "C" int tppmx_batch_sweep(tppmx_batch *b, uint64_t seed, int32_t iter0, int32_t n_sweeps) {
  if (!b || n_sweeps < 0) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  CK(ctx, cudaSetDevice(ctx->device));
  b->last_handovers = 0;
  int launches = 0;
  CK(ctx, cudaEventRecord(b->ev0, ctx->stream));
  RC(sweep_launch(b, seed, iter0, n_sweeps, true, &launches));
  CK(ctx, cudaEventRecord(b->ev1, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  CK(ctx, cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->last_ms = ms;
  b->last_launches = launches;
  return check_status(b);
}

// Parity probe of the PRODUCTION sweep kernel (include/tppmx.h): one sweep of every chain through the
// PROBE instantiation of sweep_fast_kernel, which records each step's normalised allocation
// probabilities for the units of `chain`.
extern "C" int tppmx_batch_sweep_probe(tppmx_batch *b, uint64_t seed, int32_t iter, int32_t chain, int32_t stride,
                                       double *prob_out, int32_t *K_out) {
  if (!b || !prob_out || !K_out) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (chain < 0 || chain >= d.n_chains) return fail(ctx, TPPMX_ERR_ARG, "chain out of range");
  const int ksmax = big_arms(d) ? big_ks(d, b->opt_ks) : (d.kcap < TPPMX_FAST_KSMAX ? d.kcap : TPPMX_FAST_KSMAX);
  if (stride < ksmax + d.CC) return fail(ctx, TPPMX_ERR_ARG, "stride must be >= the staged cluster capacity + CC");
  if (b->opt_impl == 1 || !b->fa.lik) return fail(ctx, TPPMX_ERR_ARG, "the probe exists for the fast sweep kernel only");
  CK(ctx, cudaSetDevice(ctx->device));
  const int T = d.T, ntp = (d.ntmax + 3) & ~3;
  const size_t np = (size_t)T * ntp * stride, nk = (size_t)T * ntp;
  double *dp = nullptr;
  int *dk = nullptr;
  CK(ctx, cudaMalloc(&dp, sizeof(double) * np));
  if (cudaMalloc(&dk, sizeof(int) * nk) != cudaSuccess) {
    cudaFree(dp);
    return fail(ctx, TPPMX_ERR_NOMEM, "probe buffers");
  }
  cudaMemsetAsync(dp, 0, sizeof(double) * np, ctx->stream);
  cudaMemsetAsync(dk, 0xFF, sizeof(int) * nk, ctx->stream);
  b->fa.probe = dp;
  b->fa.probe_k = dk;
  b->fa.probe_stride = stride;
  b->fa.probe_unit0 = chain * T;
  b->fa.probe_nunits = T;
  const int impl0 = b->opt_impl;
  b->opt_impl = 2;
  b->last_handovers = 0;
  int launches = 0;
  int rc = sweep_launch(b, seed, iter, 1, true, &launches);
  b->opt_impl = impl0;
  b->fa.probe = nullptr;
  b->fa.probe_k = nullptr;
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, "sweep probe: kernel failure");
  if (!rc) {
    std::vector<double> hp(np);
    std::vector<int> hk(nk);
    cudaError_t e = cpy_sync(ctx->stream, hp.data(), dp, sizeof(double) * np, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cpy_sync(ctx->stream, hk.data(), dk, sizeof(int) * nk, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, std::string("sweep probe: ") + cudaGetErrorString(e));
    for (int t = 0; t < T && !rc; t++)
      for (int it = 0; it < d.ntmax; it++) {
        K_out[(size_t)t * d.ntmax + it] = hk[(size_t)t * ntp + it];
        memcpy(prob_out + ((size_t)t * d.ntmax + it) * stride, hp.data() + ((size_t)t * ntp + it) * stride, sizeof(double) * stride);
      }
  }
  cudaFree(dp);
  cudaFree(dk);
  if (!rc) rc = check_status(b);
  return rc;
}

static int accumulate_launch(tppmx_batch *b, uint64_t seed, int32_t iter, int do_psm, int do_pred, int *launches,
                             cudaStream_t st = nullptr);

// ev_after_update (optional): recorded between the update kernel and the gamma draws, so that work
// which only needs eta*, beta, (Theta,Sigma), (sigma,kappa) can start beside the draws
// the main stream rejoins a horseshoe update still running on the third stream
static int hs_join(tppmx_ctx *ctx) {
  if (ctx->hs_pending) {
    CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_hs_done, 0));
    ctx->hs_pending = false;
  }
  return TPPMX_OK;
}

// overlap_hs: run the horseshoe update on the context's third stream (the caller must hs_join()
// before anything but the sweep / summaries touches the batch again); otherwise it runs in the
// first blocks of the gamma-draw kernel.
static int update_launch(tppmx_batch *b, uint64_t seed, int32_t iter, int *launches, cudaEvent_t ev_after_update = nullptr,
                         bool overlap_hs = false) {
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (d.model.cohesion == 2 && d.ncat > 0)
    return fail(ctx, TPPMX_ERR_ARG, "the (sigma,kappa) update needs ncat == 0 (as the reference, dm_ppmx_ct.cpp:942)");
  RC(hs_join(ctx));  // the beta steps read sigma_beta of the previous horseshoe update
  const int nth = 128;  // 3 worker warps + 1 serial warp (update.cuh)
  const size_t smem = sizeof(double) * (64 + (size_t)d.T * d.D * d.D + 16 + 2 * TPPMX_MAXD * TPPMX_MAXQ);
  if (d.D == 3)
    update_kernel<3><<<d.n_chains, nth, smem, ctx->stream>>>(d, b->ua, seed, iter);
  else
    update_kernel<0><<<d.n_chains, nth, smem, ctx->stream>>>(d, b->ua, seed, iter);
  if (ev_after_update) CK(ctx, cudaEventRecord(ev_after_update, ctx->stream));
  const size_t npat = (size_t)d.n_chains * d.nobs;
  const unsigned nhs = (unsigned)((d.n_chains + 127) / 128);  // horseshoe blocks (update.cuh)
  const bool hs_on = d.model.noprog != 1 && d.model.hsp == 1;
  if (overlap_hs && hs_on) {
    CK(ctx, cudaEventRecord(ctx->ev_hs_go, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(ctx->hs, ctx->ev_hs_go, 0));
    horseshoe_kernel<<<nhs, 128, 0, ctx->hs>>>(d, seed, iter);
    CK(ctx, cudaEventRecord(ctx->ev_hs_done, ctx->hs));
    ctx->hs_pending = true;
    jjss_kernel<<<(unsigned)((npat + 127) / 128), 128, 0, ctx->stream>>>(d, seed, iter, 0);
    (*launches) += 1;
  } else {
    jjss_kernel<<<(unsigned)((npat + 127) / 128) + nhs, 128, 0, ctx->stream>>>(d, seed, iter, (int)nhs);
  }
  (*launches) += 2;
  CK(ctx, cudaGetLastError());
  return TPPMX_OK;
}

extern "C" int tppmx_batch_update(tppmx_batch *b, uint64_t seed, int32_t iter) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  CK(ctx, cudaSetDevice(ctx->device));
  int launches = 0;
  CK(ctx, cudaEventRecord(b->ev0, ctx->stream));
  RC(update_launch(b, seed, iter, &launches));
  CK(ctx, cudaEventRecord(b->ev1, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  CK(ctx, cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->last_ms = ms;
  b->last_launches = launches;
  return TPPMX_OK;
}


// chain-major arrays of a DevBatch moved to chain c (layout.h); datasets and tables are shared
static void offset_chain_ptrs(DevBatch &d, const DevBatch &p, size_t c) {
  const size_t T = p.T, D = p.D, Q = p.Q;
  d.chain_ds += c; d.chain_id += c;
  d.labels += c * st_labels(p); d.nj += c * st_nj(p); d.nclu += c * T; d.njc += c * st_njc(p);
  d.idxs += c; d.status += c;
  d.sumx += c * st_sumx(p); d.sumx2 += c * st_sumx(p); d.eta += c * st_eta(p); d.etae += c * st_etae(p);
  d.JJ += c * st_pat(p); d.ss += c * p.nobs; d.lg += c * st_pat(p); d.beta += c * Q * D;
  d.sigma += c * T; d.kappa += c * T; d.Theta += c * T * D; d.Sigma += c * T * D * D;
  d.lambda += c * Q * D; d.tau += c; d.sigma_beta += c * Q * D;
  d.zb += c * st_pat(p); d.ljj += c * st_pat(p); d.cpat += c * p.nobs; d.wbuf += c * st_wbuf(p);
  d.eta_flag += c * T * p.kcap; d.beta_flag += c * Q * D;
}

// A view of chains [c0, c0 + nc) of a batch on the streams of `lane`: every per-chain array is chain-major
// (layout.h), so the view is the same batch with offset pointers; datasets, tables and the per-dataset summary
// accumulators (atomic adds) are shared.  g = group index (its own hand-over counter).
static void batch_view(const tppmx_batch *b, tppmx_batch *v, tppmx_ctx *lane, int g, int c0, int nc) {
  *v = *b;
  v->ctx = lane;
  const DevBatch &p = b->d;
  const size_t c = (size_t)c0, T = p.T, D = p.D;
  offset_chain_ptrs(v->d, p, c);
  v->d.n_chains = nc;
  UpdArgs &u = v->ua;
  u.eta2 += c * T * 2 * D * p.kcap; u.pat2 += c * 3 * T * p.ntmax; u.os += c * T * p.G; u.acc += c * T * p.kcap;
  FastArgs &f = v->fa;
  if (f.lik) {
    const size_t u0 = c * T, ntp = (p.ntmax + 3) & ~3;
    f.lik += u0 * f.lik_unit_stride; f.uq += u0 * 4 * (ntp + 1);
    f.resume_l += u0; f.resume_it += u0; f.logng += u0 * (ntp + 1); f.patpad += u0 * (ntp + 1);
    f.bail_count += g; f.kmax_dev += g;
  }
  v->fast_ready = false;  // n_units follows the view
  v->kmax_host = nullptr;
  v->bail_host = nullptr;
}

// one iteration (sweep, updates, and on a saved iteration the co-clustering counts / predictive step) of the
// batch or chain-group view `b` on the streams of b->ctx
static int run_iteration(tppmx_batch *b, uint64_t seed, int l, bool saved, int do_psm, int do_pred, int *launches) {
  tppmx_ctx *ctx = b->ctx;
  // Saved iterations: the co-clustering counts only need the partition (ready after the sweep) and
  // the predictive step only needs what the update kernel writes, so they run on the side stream
  // beside the update kernel / the JJ-ss gamma draws; the next sweep waits for them.
  if (ctx->side_pending) {
    CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
    ctx->side_pending = false;
  }
  RC(sweep_launch(b, seed, l, 1, false, launches));
  if (!saved) return update_launch(b, seed, l, launches, nullptr, true);
  if (do_psm) {
    CK(ctx, cudaEventRecord(ctx->ev_sweep, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_sweep, 0));
    RC(accumulate_launch(b, seed, l, 1, 0, launches, ctx->side));
  }
  RC(update_launch(b, seed, l, launches, do_pred ? ctx->ev_upd : nullptr, true));
  if (do_pred) {
    CK(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_upd, 0));
    RC(accumulate_launch(b, seed, l, 0, 1, launches, ctx->side));
  }
  CK(ctx, cudaEventRecord(ctx->ev_side, ctx->side));
  ctx->side_pending = true;
  return TPPMX_OK;
}

static int run_groups(const tppmx_batch *b) {
  int g = b->opt_groups;
  if (g == 0) g = 1;  // measured (profiles/r2_chain_groups.txt): at 1024 chains the kernels already fill the issue slots
  if (b->fa.lik && big_arms(b->d)) g = 1;  // the long-arm kernels size their staging per batch
  if (g > b->d.n_chains) g = b->d.n_chains;
  return g < 1 ? 1 : g;
}

extern "C" int tppmx_batch_run(tppmx_batch *b, uint64_t seed, int32_t iter0, int32_t n_iter, int32_t burn,
                               int32_t thin, int32_t do_psm, int32_t do_pred) {
  if (!b || n_iter < 0 || thin < 1) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (do_psm && !b->psm_cnt) return fail(ctx, TPPMX_ERR_CAPACITY, "co-clustering counts were not reserved (ntmax too large)");
  if (do_pred && d.npred < 1) return fail(ctx, TPPMX_ERR_ARG, "batch has npred == 0");
  CK(ctx, cudaSetDevice(ctx->device));
  b->last_handovers = 0;
  int launches = 0;
  // the chain groups: group 0 is the batch itself when there is only one
  const int G = run_groups(b);
  b->last_groups = G;
  std::vector<tppmx_batch> views;
  std::vector<tppmx_batch *> grp;
  if (G > 1) {
    while ((int)ctx->lanes.size() < G) {
      tppmx_ctx *l = ctx_new(ctx->device);
      if (!l) return fail(ctx, TPPMX_ERR_CUDA, "chain-group streams");
      ctx->lanes.push_back(l);
    }
    views.resize(G);
    for (int g = 0; g < G; g++) {
      const int c0 = (int)((long long)d.n_chains * g / G), c1 = (int)((long long)d.n_chains * (g + 1) / G);
      batch_view(b, &views[g], ctx->lanes[g], g, c0, c1 - c0);
      ctx->lanes[g]->side_pending = false;
      ctx->lanes[g]->hs_pending = false;
      grp.push_back(&views[g]);
    }
  } else {
    ctx->side_pending = false;
    grp.push_back(b);
  }
  CK(ctx, cudaEventRecord(b->ev0, ctx->stream));
  if (G > 1) {  // the groups start behind whatever the batch's own stream still holds
    CK(ctx, cudaEventRecord(ctx->ev_go, ctx->stream));
    for (tppmx_batch *v : grp) CK(ctx, cudaStreamWaitEvent(v->ctx->stream, ctx->ev_go, 0));
  }
  int interrupted_at = -1;
  const bool trace = getenv("TPPMX_TRACE") != nullptr;
  const auto t_host0 = std::chrono::steady_clock::now();
  // every error exit below leaves through this: nothing of the batch may still run on the side streams,
  // and hs_pending must not survive (get_state / update of a later call would not join it)
  struct SideGuard {
    tppmx_ctx *c;
    std::vector<tppmx_batch *> *grp;
    bool armed = true;
    ~SideGuard() {
      if (!armed) return;
      for (tppmx_batch *v : *grp) {
        tppmx_ctx *l = v->ctx;
        cudaStreamSynchronize(l->side);
        cudaStreamSynchronize(l->hs);
        cudaStreamSynchronize(l->stream);
        l->hs_pending = false;
        l->side_pending = false;
        if (l != c && !l->err.empty()) c->err = l->err;
      }
    }
  } guard{ctx, &grp};
  for (int l = iter0; l < iter0 + n_iter; l++) {
    if (ctx->interrupt && l > iter0 && (l - iter0) % ctx->check_every == 0) {  // tppmx_set_interrupt_flag
      for (tppmx_batch *v : grp) {
        RC(hs_join(v->ctx));
        CK(v->ctx, cudaStreamSynchronize(v->ctx->stream));
      }
      if (*ctx->interrupt) {
        interrupted_at = l - 1;
        break;
      }
    }
    const bool saved = (do_psm || do_pred) && (l > burn - 1) && ((l + 1) % thin == 0);  // :1060
    for (tppmx_batch *v : grp) RC(run_iteration(v, seed, l, saved, do_psm, do_pred, &launches));
  }
  for (tppmx_batch *v : grp) {  // every group's main stream rejoins its side streams, the batch's stream the groups
    tppmx_ctx *lc = v->ctx;
    if (lc->side_pending) {
      CK(lc, cudaStreamWaitEvent(lc->stream, lc->ev_side, 0));
      lc->side_pending = false;
    }
    RC(hs_join(lc));
    if (lc != ctx) {
      CK(lc, cudaEventRecord(lc->ev_done, lc->stream));
      CK(ctx, cudaStreamWaitEvent(ctx->stream, lc->ev_done, 0));
    }
  }
  if (interrupted_at >= 0) {
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    guard.armed = false;
    char buf[96];
    snprintf(buf, sizeof buf, "interrupted after iteration %d", interrupted_at);
    return fail(ctx, TPPMX_ERR_INTERRUPTED, buf);
  }
  CK(ctx, cudaEventRecord(b->ev1, ctx->stream));
  const auto t_host1 = std::chrono::steady_clock::now();
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (trace)
    fprintf(stderr, "tppmx_batch_run: %d iterations, %d groups, %d launches: host issue %.3f ms, until done %.3f ms\n", n_iter, G,
            launches, std::chrono::duration<double, std::milli>(t_host1 - t_host0).count(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count());
  guard.armed = false;   // the main stream has joined every side stream
  float ms = 0;
  CK(ctx, cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->last_ms = ms;
  b->last_launches = launches;
  if (G > 1) b->last_impl = views[0].last_impl;
  return check_status(b);
}

extern "C" int tppmx_batch_get_acceptance(tppmx_batch *b, int32_t chain, int32_t *eta_acc, int32_t *beta_acc) {
  if (!b) return TPPMX_ERR_ARG;
  const DevBatch &d = b->d;
  if (chain < 0 || chain >= d.n_chains) return fail(b->ctx, TPPMX_ERR_ARG, "chain out of range");
  CK(b->ctx, cudaSetDevice(b->ctx->device));
  CK(b->ctx, cudaStreamSynchronize(b->ctx->stream));
  const int T = d.T, ks = d.kcap, nt = d.ntmax;
  if (eta_acc) {
    std::vector<int> h((size_t)T * ks);
    RC(d2h(b, h, d.eta_flag + (size_t)chain * T * ks));
    for (int t = 0; t < T; t++)
      for (int j = 0; j < nt; j++) eta_acc[t + (size_t)T * j] = j < ks ? h[(size_t)t * ks + j] : 0;
  }
  if (beta_acc && d.Q > 0)
    CK(b->ctx, cpy_sync(b->ctx->stream, beta_acc, d.beta_flag + (size_t)chain * d.Q * d.D, sizeof(int) * d.Q * d.D, cudaMemcpyDeviceToHost));
  return TPPMX_OK;
}

extern "C" int tppmx_batch_last_timing(const tppmx_batch *b, double *ms, int32_t *launches) {
  if (!b) return TPPMX_ERR_ARG;
  if (ms) *ms = b->last_ms;
  if (launches) *launches = b->last_launches;
  return TPPMX_OK;
}

// copy every per-chain field of `src` into the scratch slot
static int copy_chain_to_scratch(tppmx_batch *b, int src) {
  const DevBatch &d = b->d;
  const int dst = b->scratch_chain;
  cudaStream_t st = b->ctx->stream;
#define CP(ptr, count)                                                                         \
  CK(b->ctx, cudaMemcpyAsync((ptr) + (size_t)dst * (count), (ptr) + (size_t)src * (count),      \
                             sizeof(*(ptr)) * (count), cudaMemcpyDeviceToDevice, st))
  CP(d.labels, st_labels(d)); CP(d.nj, st_nj(d)); CP(d.nclu, (size_t)d.T); CP(d.njc, st_njc(d));
  CP(d.idxs, (size_t)1); CP(d.sumx, st_sumx(d)); CP(d.sumx2, st_sumx(d)); CP(d.eta, st_eta(d));
  CP(d.etae, st_etae(d)); CP(d.JJ, st_pat(d)); CP(d.ss, (size_t)d.nobs); CP(d.lg, st_pat(d));
  CP(d.beta, (size_t)d.Q * d.D); CP(d.sigma, (size_t)d.T); CP(d.kappa, (size_t)d.T);
  CP(d.Theta, (size_t)d.T * d.D); CP(d.Sigma, (size_t)d.T * d.D * d.D);
#undef CP
  // the scratch slot reads the same dataset and RNG stream as `src`
  CK(b->ctx, cudaMemcpyAsync((int *)d.chain_ds + dst, d.chain_ds + src, sizeof(int), cudaMemcpyDeviceToDevice, st));
  CK(b->ctx, cudaMemcpyAsync((uint32_t *)d.chain_id + dst, d.chain_id + src, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  return TPPMX_OK;
}

extern "C" int tppmx_score_patient(tppmx_batch *b, int32_t chain, int32_t arm, int32_t it,
                                   uint64_t seed, int32_t iter, double *logw_out, double *prob_out,
                                   int32_t *K_out) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (chain < 0 || chain >= d.n_chains || arm < 0 || arm >= d.T) return fail(ctx, TPPMX_ERR_ARG, "chain/arm out of range");
  const int ds = b->chain_ds[chain];
  if (it < 0 || it >= b->arm_n[(size_t)ds * d.T + arm]) return fail(ctx, TPPMX_ERR_ARG, "it out of range");
  CK(ctx, cudaSetDevice(ctx->device));
  RC(copy_chain_to_scratch(b, chain));
  const int W = d.kcap + d.CC;
  double *dw = b->score_w;
  int *dk = b->score_k;
  score_kernel<<<1, 32, 0, ctx->stream>>>(d, b->scratch_chain, arm, it, seed, iter, dw, dw + W, dk);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  int K = 0;
  if (e == cudaSuccess) e = cpy_sync(ctx->stream, &K, dk, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && logw_out) e = cpy_sync(ctx->stream, logw_out, dw, sizeof(double) * (K + d.CC), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && prob_out) e = cpy_sync(ctx->stream, prob_out, dw + W, sizeof(double) * (K + d.CC), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("score_kernel: ") + cudaGetErrorString(e));
  if (K_out) *K_out = K;
  return TPPMX_OK;
}

extern "C" int tppmx_batch_predict(tppmx_batch *b, uint64_t seed, int32_t iter, double *pipred,
                                   double *ypred, int32_t *predclass) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (d.npred < 1) return fail(ctx, TPPMX_ERR_ARG, "batch has npred == 0");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t n1 = (size_t)d.n_chains * d.npred * d.D * d.T, n2 = (size_t)d.n_chains * d.npred * d.T;
  double *dpi = pipred ? b->pred_pi : nullptr, *dy = ypred ? b->pred_y : nullptr;
  int *dc = predclass ? b->pred_c : nullptr;
  CK(ctx, cudaEventRecord(b->ev0, ctx->stream));
  RC(predict_launch(b, d.n_chains, seed, iter, dpi, dy, dc));
  CK(ctx, cudaEventRecord(b->ev1, ctx->stream));
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e == cudaSuccess && pipred) e = d2h_out(b, pipred, dpi, n1 * sizeof(double));
  if (e == cudaSuccess && ypred) e = d2h_out(b, ypred, dy, n1 * sizeof(double));
  if (e == cudaSuccess && predclass) e = d2h_out(b, predclass, dc, n2 * sizeof(int));
  if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("predict_kernel: ") + cudaGetErrorString(e));
  float ms = 0;
  cudaEventElapsedTime(&ms, b->ev0, b->ev1);
  b->last_ms = ms;
  b->last_launches = 1;
  return TPPMX_OK;
}

// Predictive step.  Default switches (Normal-Normal auxiliary similarity, calibration 0 | 2, any reuse), with or
// without categorical covariates: the closed-form register kernel for arms of at most TPPMX_PRED_KR clusters, and
// behind it -- each block deciding for itself from its arm's K -- the staged closed-form kernel (no categorical
// covariates) or the generic kernel for the arms with more.  Every other switch combination: the generic kernel
// (its NNIG similarity through the per-n table when the batch has one, sim.cuh).  dsum != nullptr: the kernels add
// every probability to its replicate dataset's sum themselves (summary.cuh).
static int predict_launch(tppmx_batch *b, int n_chains_launch, uint64_t seed, int iter, double *dpi, double *dy, int *dc,
                          cudaStream_t st, double *dsum, int *launches) {
  int nl_local = 0;
  if (!launches) launches = &nl_local;
  tppmx_ctx *ctx = b->ctx;
  if (!st) st = ctx->stream;
  const DevBatch &d = b->d;
  const tppmx_model &m = d.model;
  const int nth = predict_threads(d);
  const bool cat = d.ncat > 0;
  // categorical covariates ride on the register kernel (their term is two table logarithms per covariate); the tables
  // are the fast sweep kernel's (reserved when the model is eligible for it)
  const bool eligible = b->opt_impl != 1 && m.PPMx == 1 && m.consim == 1 && m.similarity == 1 &&
                        (m.calibration == 0 || m.calibration == 2) && d.P >= 1 && d.kcap <= 128 &&
                        (!cat || (b->fa.cat_lw && d.ncat * d.maxC <= 256 && d.P <= 16 && d.D <= 3 && d.Q <= TPPMX_MAXQ));
  if (eligible) {
    const int KP = d.kcap;
    const size_t smem = sizeof(double) * ((size_t)(d.P + 3) * KP + (size_t)d.P * nth + (size_t)(KP + 1) * nth);
    if (smem <= 200 * 1024 || cat) {
      // arms with at most TPPMX_PRED_KR clusters (all but a few) go through the register kernel,
      // the rest through the staged one (the generic one with categorical covariates); each block decides for
      // itself and the other exits at once
      int kmin = 0;
      if (d.P <= 16 && d.D <= 3 && d.Q <= TPPMX_MAXQ) {
        const dim3 grid((unsigned)n_chains_launch * d.T);
        const size_t dyn = cat ? sizeof(int) * ((size_t)d.ncat * d.maxC + 1) * TPPMX_PRED_KR : 0;
        const double *lw = cat ? b->fa.cat_lw : nullptr, *ln = cat ? b->fa.cat_ln : nullptr;
        if (cat) {
          if (d.P <= 4) predict_reg_kernel<4, true><<<grid, nth, dyn, st>>>(d, seed, iter, dpi, dy, dc, dsum, lw, ln);
          else if (d.P <= 8) predict_reg_kernel<8, true><<<grid, nth, dyn, st>>>(d, seed, iter, dpi, dy, dc, dsum, lw, ln);
          else if (d.P <= 12) predict_reg_kernel<12, true><<<grid, nth, dyn, st>>>(d, seed, iter, dpi, dy, dc, dsum, lw, ln);
          else predict_reg_kernel<16, true><<<grid, nth, dyn, st>>>(d, seed, iter, dpi, dy, dc, dsum, lw, ln);
        } else {
          if (d.P <= 4) predict_reg_kernel<4, false><<<grid, nth, 0, st>>>(d, seed, iter, dpi, dy, dc, dsum, nullptr, nullptr);
          else if (d.P <= 8) predict_reg_kernel<8, false><<<grid, nth, 0, st>>>(d, seed, iter, dpi, dy, dc, dsum, nullptr, nullptr);
          else if (d.P <= 12) predict_reg_kernel<12, false><<<grid, nth, 0, st>>>(d, seed, iter, dpi, dy, dc, dsum, nullptr, nullptr);
          else predict_reg_kernel<16, false><<<grid, nth, 0, st>>>(d, seed, iter, dpi, dy, dc, dsum, nullptr, nullptr);
        }
        kmin = TPPMX_PRED_KR + 1;
        (*launches)++;
      }
      if (cat) {
        predict_kernel<<<dim3((unsigned)n_chains_launch * d.T), nth, 0, st>>>(d, seed, iter, dpi, dy, dc, dsum, kmin);
      } else {
        cudaError_t e = cudaFuncSetAttribute(predict_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("predict_fast_kernel: ") + cudaGetErrorString(e));
        predict_fast_kernel<<<dim3((unsigned)n_chains_launch * d.T), nth, smem, st>>>(d, seed, iter, KP, kmin, dpi, dy, dc, dsum);
      }
      (*launches)++;
      CK(ctx, cudaGetLastError());
      return TPPMX_OK;
    }
  }
  predict_kernel<<<dim3((unsigned)n_chains_launch * d.T), nth, 0, st>>>(d, seed, iter, dpi, dy, dc, dsum, 0);
  (*launches)++;
  CK(ctx, cudaGetLastError());
  return TPPMX_OK;
}

// ---- posterior summaries ------------------------------------------------------------------------
extern "C" int tppmx_batch_reset_summaries(tppmx_batch *b) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t npsm = (size_t)d.n_datasets * d.T * d.ntmax * d.ntmax;
  CK(ctx, cudaMemsetAsync(b->psm_n, 0, sizeof(int) * d.n_datasets, ctx->stream));
  CK(ctx, cudaMemsetAsync(b->pi_n, 0, sizeof(int) * d.n_datasets, ctx->stream));
  if (b->psm_cnt) CK(ctx, cudaMemsetAsync(b->psm_cnt, 0, sizeof(int) * npsm, ctx->stream));
  if (b->pi_sum) CK(ctx, cudaMemsetAsync(b->pi_sum, 0, sizeof(double) * d.n_datasets * d.npred * d.D * d.T, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  return TPPMX_OK;
}

static int accumulate_launch(tppmx_batch *b, uint64_t seed, int32_t iter, int do_psm, int do_pred, int *launches,
                             cudaStream_t st) {
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (!st) st = ctx->stream;
  if (do_psm) {
    psm_accumulate_kernel<<<dim3(d.n_chains, d.T), 256, 0, st>>>(d, b->psm_cnt);
    (*launches)++;
  }
  if (do_pred) {
    // the predictive kernels add each probability to its replicate dataset's sum themselves (no per-chain cube,
    // no outcome draw: nothing reads ypred here)
    RC(predict_launch(b, d.n_chains, seed, iter, nullptr, nullptr, nullptr, st, b->pi_sum, launches));
  }
  if (do_psm || do_pred) {
    count_saves_kernel<<<(d.n_chains + 255) / 256, 256, 0, st>>>(d, b->psm_n, b->pi_n, do_psm, do_pred);
    (*launches)++;
  }
  CK(ctx, cudaGetLastError());
  return TPPMX_OK;
}

extern "C" int tppmx_batch_accumulate(tppmx_batch *b, uint64_t seed, int32_t iter, int32_t do_psm, int32_t do_pred) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (do_psm && !b->psm_cnt) return fail(ctx, TPPMX_ERR_CAPACITY, "co-clustering counts were not reserved (ntmax too large)");
  if (do_pred && d.npred < 1) return fail(ctx, TPPMX_ERR_ARG, "batch has npred == 0");
  CK(ctx, cudaSetDevice(ctx->device));
  CK(ctx, cudaEventRecord(b->ev0, ctx->stream));
  int launches = 0;
  RC(accumulate_launch(b, seed, iter, do_psm, do_pred, &launches));
  CK(ctx, cudaEventRecord(b->ev1, ctx->stream));
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  CK(ctx, cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->last_ms = ms;
  b->last_launches = launches;
  return TPPMX_OK;
}

extern "C" int tppmx_batch_summaries_device(tppmx_batch *b, const double *util_w, double **psm, double **pi_mean,
                                            double **utility, int32_t **best_arm) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  CK(ctx, cudaSetDevice(ctx->device));
  double w[TPPMX_MAXD] = {0};
  for (int k = 0; k < d.D; k++) w[k] = util_w ? util_w[k] : 0.0;
  CK(ctx, cudaMemcpyAsync(b->util_w, w, sizeof(w), cudaMemcpyHostToDevice, ctx->stream));
  summarise_kernel<<<d.n_datasets, 256, 0, ctx->stream>>>(d, b->psm_cnt, b->psm_n, b->pi_sum, b->pi_n, b->util_w,
                                                         b->psm_cnt ? b->psm_mean : nullptr,
                                                         d.npred > 0 ? b->pi_mean : nullptr, b->utility, b->best_arm);
  CK(ctx, cudaGetLastError());
  CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (psm) *psm = b->psm_cnt ? b->psm_mean : nullptr;
  if (pi_mean) *pi_mean = b->pi_mean;
  if (utility) *utility = b->utility;
  if (best_arm) *best_arm = b->best_arm;
  return TPPMX_OK;
}

extern "C" int tppmx_batch_summaries(tppmx_batch *b, const double *util_w, double *psm, double *pi_mean,
                                     double *utility, int32_t *best_arm) {
  if (!b) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  double *dp = nullptr, *dm = nullptr, *du = nullptr;
  int32_t *db = nullptr;
  RC(tppmx_batch_summaries_device(b, util_w, &dp, &dm, &du, &db));
  const size_t npsm = (size_t)d.n_datasets * d.T * d.ntmax * d.ntmax, n1 = (size_t)d.n_datasets * d.npred * d.D * d.T;
  if (psm) {
    if (!dp) return fail(ctx, TPPMX_ERR_CAPACITY, "co-clustering counts were not reserved (ntmax too large)");
    CK(ctx, d2h_out(b, psm, dp, sizeof(double) * npsm));
  }
  if ((pi_mean || utility || best_arm) && d.npred < 1) return fail(ctx, TPPMX_ERR_ARG, "batch has npred == 0");
  if (pi_mean) CK(ctx, d2h_out(b, pi_mean, dm, sizeof(double) * n1));
  if (utility) CK(ctx, d2h_out(b, utility, du, sizeof(double) * d.n_datasets * d.npred * d.T));
  if (best_arm) CK(ctx, cpy_sync(ctx->stream, best_arm, db, sizeof(int32_t) * d.n_datasets * d.npred, cudaMemcpyDeviceToHost));
  return TPPMX_OK;
}

extern "C" int tppmx_batch_npc(tppmx_batch *b, const int32_t *trtsgn, const int32_t *outcome, int32_t *npc_out) {
  if (!b || !trtsgn || !outcome || !npc_out) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (d.npred < 1) return fail(ctx, TPPMX_ERR_ARG, "batch has npred == 0");
  CK(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)d.n_datasets * d.npred;
  int *dbuf = nullptr;
  CK(ctx, cudaMalloc(&dbuf, sizeof(int) * (2 * n + d.n_datasets)));
  cudaError_t e = cudaMemcpyAsync(dbuf, trtsgn, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dbuf + n, outcome, sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    npc_kernel<<<d.n_datasets, 128, 0, ctx->stream>>>(d, b->pi_mean, dbuf, dbuf + n, dbuf + 2 * n);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(npc_out, dbuf + 2 * n, sizeof(int) * d.n_datasets, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dbuf);
  if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("npc: ") + cudaGetErrorString(e));
  return TPPMX_OK;
}

extern "C" int tppmx_batch_point_estimate(tppmx_batch *b, int32_t loss, int32_t linkage, int32_t max_k, int32_t *labels_out,
                                          int32_t *k_out, double *loss_out) {
  if (!b || !labels_out) return TPPMX_ERR_ARG;
  tppmx_ctx *ctx = b->ctx;
  const DevBatch &d = b->d;
  if (loss < 0 || loss > 1 || linkage < 0 || linkage > 1 || max_k < 0) return fail(ctx, TPPMX_ERR_ARG, "loss 0|1, linkage 0|1, max_k >= 0");
  if (!b->psm_cnt) return fail(ctx, TPPMX_ERR_CAPACITY, "co-clustering counts were not reserved (ntmax too large)");
  if (d.ntmax > TPPMX_PE_MAXN) return fail(ctx, TPPMX_ERR_CAPACITY, "point estimate: arms of more than TPPMX_PE_MAXN patients are not supported");
  CK(ctx, cudaSetDevice(ctx->device));
  const int nth = 256;
  const size_t nu = (size_t)d.n_datasets * d.T, smem = point_estimate_smem(d.ntmax, nth);
  int *dl = nullptr, *dk = nullptr;
  double *dls = nullptr;
  CK(ctx, cudaMalloc(&dl, sizeof(int) * nu * d.ntmax));
  cudaError_t e = cudaMalloc(&dk, sizeof(int) * nu);
  if (e == cudaSuccess) e = cudaMalloc(&dls, sizeof(double) * nu);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(point_estimate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    point_estimate_kernel<<<dim3(d.n_datasets, d.T), nth, smem, ctx->stream>>>(d, b->psm_mean, loss, linkage, max_k, dl, dk, dls);
    e = cudaGetLastError();
  }
  std::vector<int> hl(nu * d.ntmax);
  if (e == cudaSuccess) e = cpy_sync(ctx->stream, hl.data(), dl, sizeof(int) * hl.size(), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && k_out) e = cpy_sync(ctx->stream, k_out, dk, sizeof(int) * nu, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && loss_out) e = cpy_sync(ctx->stream, loss_out, dls, sizeof(double) * nu, cudaMemcpyDeviceToHost);
  cudaFree(dl);
  if (dk) cudaFree(dk);
  if (dls) cudaFree(dls);
  if (e != cudaSuccess) return fail(ctx, TPPMX_ERR_CUDA, std::string("point_estimate: ") + cudaGetErrorString(e));
  for (int ds = 0; ds < d.n_datasets; ds++)   // [ds][T x ntmax col-major], the layout of cl_lab / tppmx_state.labels
    for (int t = 0; t < d.T; t++)
      for (int it = 0; it < d.ntmax; it++)
        labels_out[(size_t)ds * d.T * d.ntmax + t + (size_t)d.T * it] = hl[((size_t)ds * d.T + t) * d.ntmax + it];
  return TPPMX_OK;
}

// ---- drop-in for dm_ppmx_ct (reference src/dm_ppmx_ct.cpp:10-1472), one chain -----------------------
extern "C" int tppmx_dm_ppmx_ct(tppmx_ctx *ctx, const tppmx_ppmxct_args *a, tppmx_ppmxct_out *o) {
  if (!ctx || ctx->closed) return TPPMX_ERR_ARG;
  if (!a || !o) return fail(ctx, TPPMX_ERR_ARG, "null argument");
  if (a->iter < 1 || a->burn < 0 || a->thin < 1 || a->burn > a->iter) return fail(ctx, TPPMX_ERR_ARG, "iter/burn/thin");
  if (!a->treatments || !a->y || !a->grid || !a->similparam || !a->hP0_mu0 || !a->mhtunepar || !a->n_a ||
      !a->curr_cluster || !a->ncluster_curr || (a->Q > 0 && (!a->z || !a->initbeta)) || (a->ncon > 0 && !a->xcon) ||
      (a->ncat > 0 && (!a->xcat || !a->catvec)))
    return fail(ctx, TPPMX_ERR_ARG, "null input array");
  const int nobs = a->nobs, T = a->A, D = a->D, Q = a->Q, P = a->ncon, ncat = a->ncat, npred = a->npred, G = a->G;
  const int ntm = a->ntmax, nout = (a->iter - a->burn) / a->thin;
  if (npred > 0 && ((Q > 0 && !a->zpred) || (P > 0 && !a->xconp) || (ncat > 0 && !a->xcatp)))
    return fail(ctx, TPPMX_ERR_ARG, "npred > 0 but zpred / xconp / xcatp is null");
  if (T < 1 || T > TPPMX_MAXT || D < 1 || D > TPPMX_MAXD || ntm < 1 || G < 1) return fail(ctx, TPPMX_ERR_ARG, "dims out of range");

  tppmx_model M;
  memset(&M, 0, sizeof(M));
  M.PPMx = a->PPMx; M.cohesion = a->cohesion; M.similarity = a->similarity; M.consim = a->consim;
  M.calibration = a->calibration; M.coardegree = a->coardegree; M.reuse = a->reuse; M.CC = a->CC;
  M.upd_hier = a->upd_hier; M.hsp = a->hsp; M.noprog = a->noprog; M.nolik = 0; M.alpha = a->alpha;
  for (int k = 0; k < 7; k++) M.similparam[k] = a->similparam[k];
  for (int k = 0; k < D; k++) M.hP0_mu0[k] = a->hP0_mu0[k];
  M.hP0_nu0 = a->hP0_nu0; M.hP0_s0 = a->hP0_s0; M.hP0_Lambda0 = a->hP0_Lambda0;
  M.mhtune_eta = a->mhtunepar[0]; M.mhtune_beta = a->mhtunepar[1];

  std::vector<int32_t> treat(nobs), catvec(ncat > 0 ? ncat : 1, 0), xcat((size_t)nobs * (ncat > 0 ? ncat : 1), 0),
      xcatp((size_t)(npred > 0 ? npred : 1) * (ncat > 0 ? ncat : 1), 0);
  int maxC = 0;
  for (int i = 0; i < nobs; i++) treat[i] = (int32_t)a->treatments[i];
  for (int p = 0; p < ncat; p++) {
    catvec[p] = (int32_t)a->catvec[p];
    if (catvec[p] > maxC) maxC = catvec[p];  // max_C = catvec.max() (:107)
  }
  for (size_t e = 0; e < (size_t)nobs * ncat; e++) xcat[e] = (int32_t)a->xcat[e];
  for (size_t e = 0; e < (size_t)npred * ncat; e++) xcatp[e] = (int32_t)a->xcatp[e];

  tppmx_dims dm;
  memset(&dm, 0, sizeof(dm));
  dm.nobs = nobs; dm.T = T; dm.D = D; dm.Q = Q; dm.P = P; dm.ncat = ncat; dm.maxC = maxC; dm.npred = npred;
  dm.G = G; dm.ntmax = ntm; dm.kcap = 0;

  std::vector<double> logV;
  const double *logVp = a->logV_compact;
  if (a->cohesion == 2 && a->Vwm) {  // the two rows of the dense cube that are read (:563, :680, :981)
    logV.assign((size_t)T * 2 * ntm * G, -INFINITY);
    for (int t = 0; t < T; t++)
      for (int r = 0; r < 2; r++) {
        const int row = (int)a->n_a[t] - 2 + r;
        if (row < 0 || row >= a->Vwm_rows) continue;
        for (int K = 1; K <= ntm && K <= a->Vwm_cols; K++)
          for (int g = 0; g < G; g++)
            logV[(((size_t)t * 2 + r) * ntm + (K - 1)) * G + g] =
                log(a->Vwm[row + (size_t)a->Vwm_rows * ((K - 1) + (size_t)a->Vwm_cols * g)]);
      }
    logVp = logV.data();
  }
  if (a->cohesion == 2 && !logVp) return fail(ctx, TPPMX_ERR_ARG, "cohesion 2 needs Vwm or logV_compact");
  tppmx_shared sh = {catvec.data(), a->grid, a->cohesion == 2 ? logVp : nullptr};
  tppmx_dataset ds = {treat.data(), a->y, a->z, a->xcon, xcat.data(), a->zpred, a->xconp, xcatp.data()};

  tppmx_batch *B = nullptr;
  RC(tppmx_batch_create(ctx, &M, &dm, &sh, 1, &ds, 1, nullptr, nullptr, &B));
  std::vector<int32_t> lab((size_t)T * ntm), ncl(T);
  for (size_t e = 0; e < lab.size(); e++) lab[e] = (int32_t)a->curr_cluster[e];
  for (int t = 0; t < T; t++) ncl[t] = (int32_t)a->ncluster_curr[t];
  int rc = tppmx_batch_init(B, a->seed, lab.data(), ncl.data(), a->initbeta);

  const DevBatch &d = B->d;
  TraceBufs tr;
  memset(&tr, 0, sizeof(tr));
  tr.nout = nout > 0 ? nout : 1;
  // the traces of all saved iterations: ONE device block kept by the context and reused by later calls (a run of
  // cudaMalloc / cudaFree per call costs up to hundreds of milliseconds on a busy box -- more than the sampler)
  const size_t no = (size_t)tr.nout, n1 = (size_t)npred * D * T;
  {
    struct Seg { double **p; size_t n; };
    const Seg segs[] = {{&tr.eta, no * T * D * ntm}, {&tr.beta, no * Q * D}, {&tr.sigma, no * T}, {&tr.kappa, no * T},
                        {&tr.cl_lab, no * T * ntm}, {&tr.pi, no * nobs * D}, {&tr.pipred, no * n1}, {&tr.yispred, no * nobs * D},
                        {&tr.ypred, no * n1}, {&tr.clupred, no * npred * T}, {&tr.nclus, no * T}, {&tr.lpml, no},
                        {&tr.mnlike, (size_t)nobs}, {&tr.mnllike, (size_t)nobs}, {&tr.sumtotclu, (size_t)T}};
    size_t total = 0;
    for (const Seg &sg : segs) total += al256(sizeof(double) * (sg.n ? sg.n : 1));
    if (!rc && total > ctx->trace_bytes) {
      if (ctx->trace) cudaFree(ctx->trace);
      ctx->trace = nullptr;
      ctx->trace_bytes = 0;
      if (cudaMalloc(&ctx->trace, total) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_NOMEM, "trace buffers");
      else ctx->trace_bytes = total;
    }
    if (!rc) {
      char *q = (char *)ctx->trace;
      for (const Seg &sg : segs) {
        *sg.p = (double *)q;
        q += al256(sizeof(double) * (sg.n ? sg.n : 1));
      }
      if (cudaMemsetAsync(ctx->trace, 0, total, ctx->stream) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, "trace buffers");
    }
  }

  int launches = 0, ll = 0;
  const bool one_stream = getenv("TPPMX_DROPIN_ONE_STREAM") != nullptr;  // A/B switch: everything on the context's main stream
  const auto t_issue0 = std::chrono::steady_clock::now();
  for (int l = 0; l < a->iter && !rc; l++) {
    if (ctx->interrupt && l > 0 && l % ctx->check_every == 0) {  // tppmx_set_interrupt_flag
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, "dm_ppmx_ct: kernel failure");
      if (!rc && *ctx->interrupt) {
        char buf[96];
        snprintf(buf, sizeof buf, "interrupted after iteration %d", l - 1);
        rc = fail(ctx, TPPMX_ERR_INTERRUPTED, buf);
      }
      if (rc) break;
    }
    rc = sweep_launch(B, a->seed, l, 1, false, &launches);
    // the horseshoe update (a ~100 us serial chain that nothing reads before the NEXT iteration's beta step and
    // no output saves) runs on the context's third stream beside the gamma draws, the in-sample fit and the
    // predictive step: for ONE chain it was half of the iteration's latency
    const bool saved = (l > (a->burn - 1)) && ((l + 1) % a->thin == 0) && ll < nout;  // :1060
    const bool pred = saved && npred > 0;
    if (!rc) rc = update_launch(B, a->seed, l, &launches, (pred && !one_stream) ? ctx->ev_upd : nullptr, !one_stream);
    if (rc) break;
    if (pred && one_stream) {
      rc = predict_launch(B, 1, a->seed, l, B->pred_pi, B->pred_y, B->pred_c);
      if (rc) break;
    } else if (pred) {  // the predictive step only reads what the update kernel left: it runs on the side stream beside the gamma draws
      if (cudaStreamWaitEvent(ctx->side, ctx->ev_upd, 0) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, "dm_ppmx_ct: stream wait");
      if (!rc) rc = predict_launch(B, 1, a->seed, l, B->pred_pi, B->pred_y, B->pred_c, ctx->side);
      if (!rc && (cudaEventRecord(ctx->ev_side, ctx->side) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0) != cudaSuccess))
        rc = fail(ctx, TPPMX_ERR_CUDA, "dm_ppmx_ct: stream wait");
      if (rc) break;
    }
    sumtotclu_kernel<<<1, 32, 0, ctx->stream>>>(d, tr);
    if (saved) {
      insample_save_kernel<<<1, 256, 0, ctx->stream>>>(d, tr, a->seed, l, ll, B->pred_pi, B->pred_y, B->pred_c);
      ll++;
    }
  }
  if (!rc) rc = hs_join(ctx);
  const auto t_issue1 = std::chrono::steady_clock::now();
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, "dm_ppmx_ct: kernel failure");
  if (getenv("TPPMX_TRACE"))
    fprintf(stderr, "tppmx_dm_ppmx_ct: %d iterations, %d launches: host issue %.3f ms, until done %.3f ms\n", a->iter, launches,
            std::chrono::duration<double, std::milli>(t_issue1 - t_issue0).count(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_issue0).count());
  if (!rc) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, std::string("dm_ppmx_ct: ") + cudaGetErrorString(e));
  }
  if (!rc) rc = check_status(B);
  auto dl = [&](double *dst, const double *src, size_t n) {
    if (!rc && dst && n && cpy_sync(ctx->stream, dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = fail(ctx, TPPMX_ERR_CUDA, "dm_ppmx_ct: download");
  };
  if (nout > 0) {
    dl(o->eta, tr.eta, no * T * D * ntm); dl(o->beta, tr.beta, no * Q * D); dl(o->sigma_ngg, tr.sigma, no * T);
    dl(o->kappa_ngg, tr.kappa, no * T); dl(o->cl_lab, tr.cl_lab, no * T * ntm); dl(o->pi, tr.pi, no * nobs * D);
    dl(o->pipred, tr.pipred, no * n1); dl(o->yispred, tr.yispred, no * nobs * D); dl(o->ypred, tr.ypred, no * n1);
    dl(o->clupred, tr.clupred, no * npred * T); dl(o->nclus, tr.nclus, no * T); dl(o->lpml, tr.lpml, no);
  }
  if (!rc) {
    std::vector<double> mn(nobs), mnl(nobs), stc(T);
    dl(mn.data(), tr.mnlike, nobs); dl(mnl.data(), tr.mnllike, nobs); dl(stc.data(), tr.sumtotclu, T);
    if (!rc && o->WAIC) {  // :1445-1450
      double el = 0.0;
      for (int i = 0; i < nobs; i++) el += (2 * mnl[i] - log(mn[i]));
      *o->WAIC = -2 * el;
    }
    std::vector<int32_t> ea((size_t)T * ntm), ba((size_t)(Q > 0 ? Q : 1) * D);
    if (!rc) rc = tppmx_batch_get_acceptance(B, 0, ea.data(), ba.data());
    if (!rc && o->eta_acc)
      for (int t = 0; t < T; t++)
        for (int j = 0; j < ntm; j++) o->eta_acc[t + (size_t)T * j] = (double)ea[t + (size_t)T * j] / stc[t];  // :1432-1434
    if (!rc && o->beta_acc)
      for (int e = 0; e < Q * D; e++) o->beta_acc[e] = (double)ba[e];
    if (!rc && o->num_treat)
      for (int t = 0; t < T; t++) o->num_treat[t] = a->n_a[t];
    tppmx_state st;
    memset(&st, 0, sizeof(st));
    st.Theta = o->theta_prior;
    st.Sigma = o->sigma_prior;
    if (!rc) rc = tppmx_batch_get_state(B, 0, &st);
  }
  tppmx_batch_destroy(B);
  return rc;
}

// ---- drop-in for prior_ppmx_core (reference src/prior_ppmx.cpp:9-466) --------------------------------
// saved iteration ll of every chain: nclu(ll) = nclu_curr, njout.col(ll) = nj_curr (:455-459)
__global__ void prior_save_kernel(DevBatch b, int ll, int nout, double *nclu_tr, double *nj_tr) {
  const int chain = blockIdx.x;
  const int K = b.nclu[(size_t)chain * b.T];
  if (threadIdx.x == 0) nclu_tr[(size_t)chain * nout + ll] = (double)K;
  const int *nj = b.nj + (size_t)chain * st_nj(b);
  for (int j = threadIdx.x; j < b.nobs; j += blockDim.x)
    nj_tr[((size_t)chain * nout + ll) * b.nobs + j] = (j < K && j < b.kcap) ? (double)nj[j] : 0.0;
}

extern "C" int tppmx_prior_ppmx_core(tppmx_ctx *ctx, const tppmx_prior_args *a, tppmx_prior_out *o) {
  if (!ctx || ctx->closed) return TPPMX_ERR_ARG;
  if (!a || !o) return fail(ctx, TPPMX_ERR_ARG, "null argument");
  if (a->iter < 1 || a->burn < 0 || a->thin < 1 || a->burn > a->iter || a->nobs < 1 || a->n_chains < 1)
    return fail(ctx, TPPMX_ERR_ARG, "iter/burn/thin/nobs/n_chains");
  if (a->ncat != 0) return fail(ctx, TPPMX_ERR_ARG, "prior_ppmx_core has no categorical covariates (R/prior_ppmx.R:114)");
  if (!a->similparam || !a->curr_cluster || (a->ncon > 0 && !a->xcon)) return fail(ctx, TPPMX_ERR_ARG, "null input array");
  if (a->cohesion == 2 && !a->Vwm) return fail(ctx, TPPMX_ERR_ARG, "cohesion 2 needs Vwm");
  const int nobs = a->nobs, P = a->ncon, nout = (a->iter - a->burn) / a->thin, nc = a->n_chains;

  tppmx_model M;
  memset(&M, 0, sizeof(M));
  M.PPMx = a->PPMx; M.cohesion = a->cohesion; M.similarity = a->similarity; M.consim = a->consim;
  M.calibration = a->calibration; M.coardegree = a->coardegree; M.reuse = 1; M.CC = a->CC;
  M.nolik = 1; M.noprog = 1; M.alpha = a->alpha;
  for (int k = 0; k < 7; k++) M.similparam[k] = a->similparam[k];
  tppmx_dims dm;
  memset(&dm, 0, sizeof(dm));
  dm.nobs = nobs; dm.T = 1; dm.D = 1; dm.Q = 0; dm.P = P; dm.npred = 0; dm.G = 1; dm.ntmax = nobs; dm.kcap = 0;

  // the two rows of V that are read (:239), on the log scale, in the compact layout of tppmx_shared.logV
  std::vector<double> logV((size_t)2 * nobs, -INFINITY), grid = {a->sigma, 0.0};
  if (a->cohesion == 2)
    for (int r = 0; r < 2; r++) {
      const int row = nobs - 2 + r;
      if (row < 0 || row >= a->Vwm_rows) continue;
      for (int K = 1; K <= nobs && K <= a->Vwm_cols; K++)
        logV[(size_t)r * nobs + (K - 1)] = log(a->Vwm[row + (size_t)a->Vwm_rows * (K - 1)]);
    }
  std::vector<int32_t> treat(nobs, 0), lab(nobs), xcat(nobs, 0), catvec(1, 0);
  std::vector<double> y(nobs, 1.0);
  int K0 = 0;
  for (int i = 0; i < nobs; i++) {
    lab[i] = (int32_t)a->curr_cluster[i];
    if (lab[i] > K0) K0 = lab[i];
  }
  if (a->ncluster_curr != K0) return fail(ctx, TPPMX_ERR_ARG, "ncluster_curr does not match curr_cluster");
  tppmx_shared sh = {catvec.data(), grid.data(), a->cohesion == 2 ? logV.data() : nullptr};
  tppmx_dataset ds = {treat.data(), y.data(), nullptr, a->xcon, xcat.data(), nullptr, nullptr, nullptr};

  tppmx_batch *B = nullptr;
  std::vector<int32_t> cds(nc, 0);
  RC(tppmx_batch_create(ctx, &M, &dm, &sh, 1, &ds, nc, cds.data(), nullptr, &B));
  int32_t ncl = K0;
  int rc = tppmx_batch_init(B, a->seed, lab.data(), &ncl, nullptr);
  double *d_nclu = nullptr, *d_nj = nullptr;
  const size_t n1 = (size_t)nc * (nout > 0 ? nout : 1), n2 = n1 * nobs;
  if (!rc && (cudaMalloc(&d_nclu, sizeof(double) * n1) != cudaSuccess || cudaMalloc(&d_nj, sizeof(double) * n2) != cudaSuccess))
    rc = fail(ctx, TPPMX_ERR_NOMEM, "trace buffers");
  int launches = 0, ll = 0;
  for (int l = 0; l < a->iter && !rc; l++) {
    if (ctx->interrupt && l > 0 && l % ctx->check_every == 0) {
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, "prior_ppmx_core: kernel failure");
      if (!rc && *ctx->interrupt) rc = fail(ctx, TPPMX_ERR_INTERRUPTED, "interrupted");
      if (rc) break;
    }
    rc = sweep_launch(B, a->seed, l, 1, false, &launches);
    if (!rc && (l > a->burn - 1) && ((l + 1) % a->thin == 0) && ll < nout) {  // :455
      prior_save_kernel<<<nc, 128, 0, ctx->stream>>>(B->d, ll, nout, d_nclu, d_nj);
      ll++;
    }
  }
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, "prior_ppmx_core: kernel failure");
  if (!rc) rc = check_status(B);
  if (!rc && nout > 0) {
    cudaError_t e = cudaSuccess;
    if (o->nclu) e = cpy_sync(ctx->stream, o->nclu, d_nclu, sizeof(double) * n1, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && o->njout) e = cpy_sync(ctx->stream, o->njout, d_nj, sizeof(double) * n2, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(ctx, TPPMX_ERR_CUDA, std::string("prior_ppmx_core: ") + cudaGetErrorString(e));
  }
  if (d_nclu) cudaFree(d_nclu);
  if (d_nj) cudaFree(d_nj);
  tppmx_batch_destroy(B);
  return rc;
}
