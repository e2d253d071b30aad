// multi.cu — one process, several devices behind the C ABI (include/tppmx.h, tppmx_multi_*).
//
// The path shards by independent units: replicate datasets, with all their chains, go
// to the devices in contiguous blocks; nothing is exchanged while the chains run.  The ONLY
// cross-device step is the final gather of the per-dataset posterior summaries (co-clustering
// matrix, predictive means, utilities, chosen arms), done with NCCL over NVLink: one communicator per
// device from ncclCommInitAll, one group of ncclSend / ncclRecv that lands every device's slab at its
// offset in device 0's buffers, then one copy to the caller.  NCCL is resolved at run time
// (dlopen "libnccl.so.2": the process's already loaded copy if there is one, e.g. torch's), so the
// library has no link-time dependency on it and single-device use never touches it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/tppmx.h"

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err) {
    if (h) return true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) {
      err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror();
      return false;
    }
#define SYM(field, name)                                     \
  field = reinterpret_cast<decltype(field)>(dlsym(h, name)); \
  if (!field) {                                              \
    err = std::string("NCCL symbol missing: ") + name;       \
    return false;                                            \
  }
    SYM(CommInitAll, "ncclCommInitAll");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
  }
};

}  // namespace

struct tppmx_multi {
  std::vector<int> dev;
  std::vector<tppmx_ctx *> ctx;
  std::vector<ncclComm_t> comm;
  std::vector<cudaStream_t> st;
  NcclApi nccl;
  std::string err;
};

static int mfail(tppmx_multi *m, int code, const std::string &msg) {
  if (m) m->err = msg;
  return code;
}

extern "C" const char *tppmx_multi_last_error(const tppmx_multi *m) { return m ? m->err.c_str() : "no handle"; }
extern "C" int tppmx_multi_device_count(const tppmx_multi *m) { return m ? (int)m->dev.size() : 0; }

extern "C" int tppmx_multi_destroy(tppmx_multi *m) {
  if (!m) return TPPMX_OK;
  for (size_t r = 0; r < m->comm.size(); r++)
    if (m->comm[r]) m->nccl.CommDestroy(m->comm[r]);
  for (size_t r = 0; r < m->st.size(); r++) {
    cudaSetDevice(m->dev[r]);
    if (m->st[r]) cudaStreamDestroy(m->st[r]);
  }
  for (tppmx_ctx *c : m->ctx) tppmx_destroy(c);
  delete m;
  return TPPMX_OK;
}

extern "C" int tppmx_multi_create(tppmx_multi **out, const int32_t *device_ids, int32_t ndev) {
  if (!out || ndev < 1 || ndev > 64) return TPPMX_ERR_ARG;
  *out = nullptr;
  int nvis = 0;
  if (cudaGetDeviceCount(&nvis) != cudaSuccess || nvis < 1) return TPPMX_ERR_CUDA;
  tppmx_multi *m = new tppmx_multi();
  for (int r = 0; r < ndev; r++) {
    const int d = device_ids ? device_ids[r] : r;
    if (d < 0 || d >= nvis) {
      tppmx_multi_destroy(m);
      return TPPMX_ERR_ARG;
    }
    for (int q : m->dev)
      if (q == d) {  // NCCL wants distinct devices
        tppmx_multi_destroy(m);
        return TPPMX_ERR_ARG;
      }
    m->dev.push_back(d);
  }
  for (int r = 0; r < ndev; r++) {
    tppmx_ctx *c = nullptr;
    if (tppmx_create(&c, m->dev[r]) != TPPMX_OK) {
      tppmx_multi_destroy(m);
      return TPPMX_ERR_CUDA;
    }
    m->ctx.push_back(c);
    cudaStream_t s = nullptr;
    cudaSetDevice(m->dev[r]);
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    m->st.push_back(s);
  }
  if (ndev > 1) {
    std::string e;
    if (!m->nccl.load(e)) {
      tppmx_multi_destroy(m);
      return TPPMX_ERR_ARG;
    }
    m->comm.assign(ndev, nullptr);
    const ncclResult_t rc = m->nccl.CommInitAll(m->comm.data(), ndev, m->dev.data());
    if (rc != ncclSuccess) {
      m->comm.clear();
      tppmx_multi_destroy(m);
      return TPPMX_ERR_CUDA;
    }
  }
  *out = m;
  return TPPMX_OK;
}

// contiguous blocks, sizes differing by at most one (the same rule as treatppmx_b200/multigpu.py:shard_range)
static void shard_range(int n, int world, int rank, int *lo, int *hi) {
  const int base = n / world, rem = n % world;
  *lo = rank * base + (rank < rem ? rank : rem);
  *hi = *lo + base + (rank < rem ? 1 : 0);
}

extern "C" int tppmx_multi_run_study(tppmx_multi *m, const tppmx_study *in, tppmx_study_out *out) {
  if (!m) return TPPMX_ERR_ARG;
  if (!in || !out || !in->model || !in->dims || !in->shared || !in->datasets || !in->init_labels || !in->init_nclu)
    return mfail(m, TPPMX_ERR_ARG, "null argument");
  if (in->n_datasets < 1 || in->chains_per_dataset < 1 || in->n_iter < 0 || in->thin < 1)
    return mfail(m, TPPMX_ERR_ARG, "n_datasets / chains_per_dataset / n_iter / thin");
  const int ndev = (int)m->dev.size(), nd = in->n_datasets, cpd = in->chains_per_dataset;
  const tppmx_dims &dm = *in->dims;
  const int T = dm.T, nt = dm.ntmax, D = dm.D, npred = dm.npred;
  const size_t s_psm = (size_t)T * nt * nt, s_pi = (size_t)npred * D * T, s_ut = (size_t)T * npred, s_ba = (size_t)npred;
  const bool want_psm = in->do_psm != 0, want_pred = in->do_pred != 0 && npred > 0;

  std::vector<tppmx_batch *> B(ndev, nullptr);
  std::vector<int> rcs(ndev, TPPMX_OK), lo(ndev), hi(ndev);
  std::vector<double> ms(ndev, 0.0);
  std::vector<double *> d_psm(ndev, nullptr), d_pi(ndev, nullptr), d_ut(ndev, nullptr);
  std::vector<int32_t *> d_ba(ndev, nullptr);
  for (int r = 0; r < ndev; r++) shard_range(nd, ndev, r, &lo[r], &hi[r]);

  // ---- the chains: one host thread per device, nothing shared but the read-only inputs
  auto work = [&](int r) {
    const int n = hi[r] - lo[r];
    if (n <= 0) return;
    std::vector<int32_t> cds((size_t)n * cpd), cid((size_t)n * cpd);
    for (int d = 0; d < n; d++)
      for (int c = 0; c < cpd; c++) {
        cds[(size_t)d * cpd + c] = d;
        cid[(size_t)d * cpd + c] = (lo[r] + d) * cpd + c;  // global chain index: results do not depend on ndev
      }
    int rc = tppmx_batch_create(m->ctx[r], in->model, in->dims, in->shared, n, in->datasets + lo[r], n * cpd, cds.data(),
                                cid.data(), &B[r]);
    if (!rc) rc = tppmx_batch_init(B[r], in->seed, in->init_labels + (size_t)lo[r] * T * nt, in->init_nclu + (size_t)lo[r] * T,
                                   in->initbeta);
    if (!rc) rc = tppmx_batch_reset_summaries(B[r]);
    if (!rc) rc = tppmx_batch_run(B[r], in->seed, 0, in->n_iter, in->burn, in->thin, want_psm, want_pred);
    if (!rc) tppmx_batch_last_timing(B[r], &ms[r], nullptr);
    if (!rc) rc = tppmx_batch_summaries_device(B[r], in->util_w, &d_psm[r], &d_pi[r], &d_ut[r], &d_ba[r]);
    rcs[r] = rc;
  };
  {
    std::vector<std::thread> th;
    for (int r = 1; r < ndev; r++) th.emplace_back(work, r);
    work(0);
    for (auto &t : th) t.join();
  }
  int rc = TPPMX_OK;
  for (int r = 0; r < ndev && !rc; r++)
    if (rcs[r]) rc = mfail(m, rcs[r], std::string("device ") + std::to_string(m->dev[r]) + ": " + tppmx_last_error(m->ctx[r]));

  // ---- the gather: every device's slab to its offset on device 0 (NCCL), then one copy to the host
  float gather_ms = 0.f;
  if (!rc) {
    cudaSetDevice(m->dev[0]);
    double *g_psm = nullptr, *g_pi = nullptr, *g_ut = nullptr;
    int32_t *g_ba = nullptr;
    cudaError_t ce = cudaSuccess;
    if (ndev > 1) {
      if (want_psm && d_psm[0]) ce = cudaMalloc(&g_psm, sizeof(double) * s_psm * nd);
      if (ce == cudaSuccess && want_pred) ce = cudaMalloc(&g_pi, sizeof(double) * s_pi * nd);
      if (ce == cudaSuccess && want_pred) ce = cudaMalloc(&g_ut, sizeof(double) * s_ut * nd);
      if (ce == cudaSuccess && want_pred) ce = cudaMalloc(&g_ba, sizeof(int32_t) * s_ba * nd);
      if (ce != cudaSuccess) rc = mfail(m, TPPMX_ERR_NOMEM, "gather buffers on device 0");
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      if (!rc) {
        cudaEventRecord(e0, m->st[0]);
        ncclResult_t nr = m->nccl.GroupStart();
        for (int r = 0; r < ndev && nr == ncclSuccess; r++) {
          const size_t n = (size_t)(hi[r] - lo[r]);
          if (n == 0) continue;
          struct Slab { const void *src; void *dst; size_t count; ncclDataType_t ty; size_t esz; };
          const Slab slabs[4] = {
              {want_psm ? d_psm[r] : nullptr, g_psm ? g_psm + s_psm * lo[r] : nullptr, s_psm * n, ncclDouble, 8},
              {want_pred ? d_pi[r] : nullptr, g_pi ? g_pi + s_pi * lo[r] : nullptr, s_pi * n, ncclDouble, 8},
              {want_pred ? d_ut[r] : nullptr, g_ut ? g_ut + s_ut * lo[r] : nullptr, s_ut * n, ncclDouble, 8},
              {want_pred ? d_ba[r] : nullptr, g_ba ? (void *)(g_ba + s_ba * lo[r]) : nullptr, s_ba * n, ncclInt32, 4}};
          for (const Slab &s : slabs) {
            if (!s.src || !s.dst || nr != ncclSuccess) continue;
            if (r == 0) {
              cudaMemcpyAsync(s.dst, s.src, s.count * s.esz, cudaMemcpyDeviceToDevice, m->st[0]);
            } else {
              nr = m->nccl.Send(s.src, s.count, s.ty, 0, m->comm[r], m->st[r]);
              if (nr == ncclSuccess) nr = m->nccl.Recv(s.dst, s.count, s.ty, r, m->comm[0], m->st[0]);
            }
          }
        }
        const ncclResult_t ne = m->nccl.GroupEnd();
        if (nr == ncclSuccess) nr = ne;
        if (nr != ncclSuccess) rc = mfail(m, TPPMX_ERR_CUDA, std::string("NCCL gather: ") + m->nccl.GetErrorString(nr));
        cudaSetDevice(m->dev[0]);
        cudaEventRecord(e1, m->st[0]);
        for (int r = 0; r < ndev; r++) {
          cudaSetDevice(m->dev[r]);
          if (cudaStreamSynchronize(m->st[r]) != cudaSuccess && !rc) rc = mfail(m, TPPMX_ERR_CUDA, "gather: stream failure");
        }
        cudaSetDevice(m->dev[0]);
        if (!rc) cudaEventElapsedTime(&gather_ms, e0, e1);
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    } else {
      g_psm = d_psm[0]; g_pi = d_pi[0]; g_ut = d_ut[0]; g_ba = d_ba[0];
    }
    cudaSetDevice(m->dev[0]);
    if (!rc) {
      if (out->psm && want_psm) {
        if (!g_psm) rc = mfail(m, TPPMX_ERR_CAPACITY, "co-clustering counts were not reserved (ntmax too large)");
        else ce = cudaMemcpy(out->psm, g_psm, sizeof(double) * s_psm * nd, cudaMemcpyDeviceToHost);
      }
      if (!rc && ce == cudaSuccess && out->pi_mean && want_pred) ce = cudaMemcpy(out->pi_mean, g_pi, sizeof(double) * s_pi * nd, cudaMemcpyDeviceToHost);
      if (!rc && ce == cudaSuccess && out->utility && want_pred) ce = cudaMemcpy(out->utility, g_ut, sizeof(double) * s_ut * nd, cudaMemcpyDeviceToHost);
      if (!rc && ce == cudaSuccess && out->best_arm && want_pred) ce = cudaMemcpy(out->best_arm, g_ba, sizeof(int32_t) * s_ba * nd, cudaMemcpyDeviceToHost);
      if (!rc && ce != cudaSuccess) rc = mfail(m, TPPMX_ERR_CUDA, std::string("gather download: ") + cudaGetErrorString(ce));
    }
    if (ndev > 1) {
      if (g_psm) cudaFree(g_psm);
      if (g_pi) cudaFree(g_pi);
      if (g_ut) cudaFree(g_ut);
      if (g_ba) cudaFree(g_ba);
    }
  }
  double msmax = 0.0;
  for (int r = 0; r < ndev; r++) {
    if (ms[r] > msmax) msmax = ms[r];
    if (B[r]) tppmx_batch_destroy(B[r]);
  }
  out->ms_run = msmax;
  out->ms_gather = gather_ms;
  out->n_devices = ndev;
  out->used_nccl = ndev > 1 ? 1 : 0;
  return rc;
}
