// multi_device.cu — one ensemble over several GPUs of one box, behind the C ABI (pmc_multi_*).
//
// Replaces the fan-out of the reference's launchers (`julia -p N run/interacting_dielectric_study.jl`,
// run/interacting_dielectric_study.jl:37-47: pmap over cases, one single-threaded process per case): chains
// (sweep points × replicas) are independent, so device g owns the contiguous block [g·R/G, (g+1)·R/G) of global
// chain ids (Philox streams are keyed by the global id ⇒ results do not depend on G), runs it from its own host
// thread with NO data-path collective, and only the final per-chain result rows ([R/G][24] doubles) are gathered:
// one ncclAllGather over NVLink, or peer-to-peer copies when no NCCL library can be loaded.
// NCCL is opened with dlopen at run time (libnccl.so.2, e.g. the one a host process already carries), so
// libpolymc_b200.so has no link-time dependency on it.
#include "handle.h"

#include <dlfcn.h>
#include <nccl.h>

#include <thread>

namespace {

int fail(int code, const std::string& msg) { return pmc_fail(code, msg); }

// The result row of one chain, packed on the device (PMC_RESULT_COLS doubles, include/polymc.h).
__global__ void k_pack_results(const ChainDyn* __restrict__ dyn, const ChainDynX* __restrict__ dynx, int nchains,
                               double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchains) return;
  const ChainDyn& d = dyn[c];
  double* o = out + (size_t)c * PMC_RESULT_COLS;
  const double nrm = d.acc[16] + d.comp[16];
  for (int k = 0; k < 16; ++k) o[k] = (d.acc[k] + d.comp[k]) / nrm;  // get_avg, average.jl:38
  o[16] = d.steps_total ? (double)d.nacc_total / (double)d.steps_total : 0.0;
  o[17] = nrm;
  o[18] = d.phi_step;
  o[19] = d.theta_step;
  o[20] = (double)d.steps_total;
  o[21] = d.U;
  o[22] = (dynx[c].acc[0] + dynx[c].comp[0]) / nrm;
  o[23] = (dynx[c].acc[1] + dynx[c].comp[1]) / nrm;
}

struct NcclApi {
  void* dl = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok() const { return dl != nullptr; }
};

NcclApi load_nccl() {
  NcclApi a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    void* dl = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!dl) continue;
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(dl, "ncclCommInitAll"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(dl, "ncclCommDestroy"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(dl, "ncclAllGather"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(dl, "ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(dl, "ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(dl, "ncclGetErrorString"));
    if (a.CommInitAll && a.CommDestroy && a.AllGather && a.GroupStart && a.GroupEnd && a.GetErrorString) {
      a.dl = dl;
      return a;
    }
    dlclose(dl);
  }
  return NcclApi{};
}

}  // namespace

struct pmc_multi {
  std::vector<int> devices;
  std::vector<pmc_handle*> shard;
  std::vector<int64_t> first, count;    // global chain ids [first, first+count) of each shard
  int64_t nchains = 0;
  int n = 0;
  int64_t cmax = 0;                     // largest shard (the all-gather is padded to it)
  std::vector<cudaStream_t> stream;     // gather streams, one per device
  std::vector<double*> res_local;       // [cmax][PMC_RESULT_COLS] on each device
  std::vector<double*> res_all;         // [G·cmax][PMC_RESULT_COLS] on each device (NCCL) / on device 0 (peer copies)
  NcclApi nccl;
  std::vector<ncclComm_t> comm;
  bool use_nccl = false;
  // asynchronous run
  std::vector<std::thread> worker;
  std::vector<int> worker_rc;
  std::vector<std::string> worker_err;
  bool running = false;
};

namespace {

int join_workers(pmc_multi* m) {
  for (auto& t : m->worker)
    if (t.joinable()) t.join();
  m->worker.clear();
  m->running = false;
  for (size_t g = 0; g < m->worker_rc.size(); ++g)
    if (m->worker_rc[g]) return fail(m->worker_rc[g], "device " + std::to_string(m->devices[g]) + ": " + m->worker_err[g]);
  return PMC_OK;
}

// Runs fn(slot) on one host thread per device and waits; the first failing slot's status and message are returned.
template <class F>
int on_every_device(pmc_multi* m, F fn) {
  const size_t G = m->shard.size();
  std::vector<int> rc(G, 0);
  std::vector<std::string> err(G);
  std::vector<std::thread> th;
  for (size_t g = 0; g < G; ++g)
    th.emplace_back([&, g]() {
      rc[g] = fn((int)g);
      if (rc[g]) err[g] = pmc_last_error();
    });
  for (auto& t : th) t.join();
  for (size_t g = 0; g < G; ++g)
    if (rc[g]) return fail(rc[g], "device " + std::to_string(m->devices[g]) + ": " + err[g]);
  return PMC_OK;
}

int check_multi(const pmc_multi* m) {
  if (!m) return fail(PMC_ERR_INVALID, "null multi-device handle");
  if (m->running) return fail(PMC_ERR_INVALID, "a pmc_multi_run_async is in flight: call pmc_multi_wait first");
  return PMC_OK;
}

}  // namespace

extern "C" {

int32_t pmc_multi_create(const pmc_case* cases, int64_t ncases, int32_t replicas_per_case, uint64_t seed,
                         const int32_t* devices, int32_t ndevices, pmc_multi** out) {
  if (!out) return fail(PMC_ERR_INVALID, "null out handle");
  *out = nullptr;
  if (!cases || ncases < 1 || replicas_per_case < 1) return fail(PMC_ERR_INVALID, "need >= 1 case and >= 1 replica");
  // pmc_create checks a shard's cases against the shard's own first case only.  The state offsets below assume one n, and
  // every shard must pick the same kernel family: a handle runs the composite-trial kernels as soon as ONE of its cases has
  // cluster flips or bending (polymc.cu needs_cluster_path), so a mixed list would make a case's kernel — and the rounding
  // of its sums — depend on which other cases share its device.
  auto composite = [](const pmc_case& c) { return c.clustering != 0 || c.kappa != 0.0; };
  for (int64_t c = 1; c < ncases; ++c)
    if (cases[c].n != cases[0].n || cases[c].energy_type != cases[0].energy_type || cases[c].planar != cases[0].planar ||
        composite(cases[c]) != composite(cases[0]))
      return fail(PMC_ERR_INVALID,
                  "all cases of one multi-device ensemble must share num-monomers, energy-type and driver "
                  "(plain / clustering / 2-D); bucket the sweep");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return fail(PMC_ERR_NO_DEVICE, "no CUDA device available (libpolymc_b200 has no CPU fallback)");
  }
  if (ndevices <= 0) ndevices = ndev;  // all devices of the box
  const int64_t R = ncases * (int64_t)replicas_per_case;
  if (R > (int64_t)0x7fffffff) return fail(PMC_ERR_INVALID, "too many chains");
  if ((int64_t)ndevices > R) ndevices = (int)R;
  pmc_multi* m = new (std::nothrow) pmc_multi();
  if (!m) return fail(PMC_ERR_NOMEM, "host allocation failed");
  for (int g = 0; g < ndevices; ++g) {
    const int d = devices ? devices[g] : g;
    if (d < 0 || d >= ndev) {
      delete m;
      return fail(PMC_ERR_NO_DEVICE, "device index out of range");
    }
    m->devices.push_back(d);
  }
  const int G = ndevices;
  m->nchains = R;
  m->n = (int)cases[0].n;
  m->shard.assign((size_t)G, nullptr);
  m->first.resize((size_t)G);
  m->count.resize((size_t)G);
  for (int g = 0; g < G; ++g) {  // contiguous blocks of global chain ids
    m->first[(size_t)g] = R * g / G;
    m->count[(size_t)g] = R * (g + 1) / G - R * g / G;
    m->cmax = std::max(m->cmax, m->count[(size_t)g]);
  }
  // A block may start or end inside a case's replicas, so every shard is created from its own per-chain case list
  // (replicas_per_case = 1); chain_id_base = its first global id keeps the Philox streams those of the whole ensemble.
  int rc = on_every_device(m, [&](int g) -> int {
    std::vector<pmc_case> mine((size_t)m->count[(size_t)g]);
    for (int64_t j = 0; j < m->count[(size_t)g]; ++j)
      mine[(size_t)j] = cases[(m->first[(size_t)g] + j) / replicas_per_case];
    return pmc_create(mine.data(), (int64_t)mine.size(), 1, seed, m->devices[(size_t)g], (uint32_t)m->first[(size_t)g],
                      &m->shard[(size_t)g]);
  });
  if (rc) {
    pmc_multi_destroy(m);
    return rc;
  }
  m->stream.assign((size_t)G, nullptr);
  m->res_local.assign((size_t)G, nullptr);
  m->res_all.assign((size_t)G, nullptr);
  const char* forced = std::getenv("PMC_MULTI_GATHER");  // "peer" forces the copy path (experiments, tests)
  if (G > 1 && !(forced && std::strcmp(forced, "peer") == 0)) m->nccl = load_nccl();
  m->use_nccl = G > 1 && m->nccl.ok();
  auto setup = [&]() -> int {
    for (int g = 0; g < G; ++g) {
      PMC_CU(cudaSetDevice(m->devices[(size_t)g]));
      PMC_CU(cudaStreamCreateWithFlags(&m->stream[(size_t)g], cudaStreamNonBlocking));
      PMC_CU(cudaMalloc(&m->res_local[(size_t)g], (size_t)m->cmax * PMC_RESULT_COLS * sizeof(double)));
      PMC_CU(cudaMemset(m->res_local[(size_t)g], 0, (size_t)m->cmax * PMC_RESULT_COLS * sizeof(double)));
      if (m->use_nccl || g == 0)
        PMC_CU(cudaMalloc(&m->res_all[(size_t)g], (size_t)G * (size_t)m->cmax * PMC_RESULT_COLS * sizeof(double)));
    }
    if (m->use_nccl) {
      m->comm.assign((size_t)G, nullptr);
      const ncclResult_t r = m->nccl.CommInitAll(m->comm.data(), G, m->devices.data());
      if (r != ncclSuccess) {  // keep going on peer copies
        m->comm.clear();
        m->use_nccl = false;
      }
    }
    if (!m->use_nccl && G > 1) {
      PMC_CU(cudaSetDevice(m->devices[0]));
      for (int g = 1; g < G; ++g) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->devices[0], m->devices[(size_t)g]);
        if (can) {
          const cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[(size_t)g], 0);
          if (e != cudaSuccess) cudaGetLastError();  // already enabled is fine; cudaMemcpyPeerAsync stages otherwise
        }
      }
    }
    return PMC_OK;
  };
  rc = setup();
  if (rc) {
    pmc_multi_destroy(m);
    return rc;
  }
  *out = m;
  return PMC_OK;
}

void pmc_multi_destroy(pmc_multi* m) {
  if (!m) return;
  for (auto& t : m->worker)
    if (t.joinable()) t.join();
  for (size_t g = 0; g < m->comm.size(); ++g)
    if (m->comm[g]) m->nccl.CommDestroy(m->comm[g]);
  for (size_t g = 0; g < m->devices.size(); ++g) {
    cudaSetDevice(m->devices[g]);
    if (g < m->res_local.size() && m->res_local[g]) cudaFree(m->res_local[g]);
    if (g < m->res_all.size() && m->res_all[g]) cudaFree(m->res_all[g]);
    if (g < m->stream.size() && m->stream[g]) cudaStreamDestroy(m->stream[g]);
    if (g < m->shard.size() && m->shard[g]) pmc_destroy(m->shard[g]);
  }
  delete m;
}

int32_t pmc_multi_num_devices(const pmc_multi* m) { return m ? (int32_t)m->devices.size() : 0; }
int64_t pmc_multi_num_chains(const pmc_multi* m) { return m ? m->nchains : 0; }
const char* pmc_multi_gather_backend(const pmc_multi* m) {
  return !m ? "" : m->devices.size() < 2 ? "none" : m->use_nccl ? "nccl" : "peer";
}

pmc_handle* pmc_multi_shard(pmc_multi* m, int32_t slot, int32_t* device, int64_t* first_chain, int64_t* nchains) {
  if (!m || slot < 0 || (size_t)slot >= m->shard.size()) {
    fail(PMC_ERR_INVALID, "shard slot out of range");
    return nullptr;
  }
  if (device) *device = m->devices[(size_t)slot];
  if (first_chain) *first_chain = m->first[(size_t)slot];
  if (nchains) *nchains = m->count[(size_t)slot];
  return m->shard[(size_t)slot];
}

int32_t pmc_multi_set_ensemble_hint(pmc_multi* m, int64_t ensemble_chains) {
  int rc = check_multi(m);
  if (rc) return rc;
  for (pmc_handle* h : m->shard)
    if ((rc = pmc_set_ensemble_hint(h, ensemble_chains))) return rc;
  return PMC_OK;
}

int32_t pmc_multi_set_pair_precision(pmc_multi* m, int32_t mode) {
  int rc = check_multi(m);
  if (rc) return rc;
  for (pmc_handle* h : m->shard)
    if ((rc = pmc_set_pair_precision(h, mode))) return rc;
  return PMC_OK;
}

int32_t pmc_multi_begin_stage(pmc_multi* m, double kT_scale) {
  int rc = check_multi(m);
  if (rc) return rc;
  return on_every_device(m, [&](int g) { return pmc_begin_stage(m->shard[(size_t)g], kT_scale); });
}

int32_t pmc_multi_set_state_all(pmc_multi* m, const double* phi, const double* theta) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!phi || !theta) return fail(PMC_ERR_INVALID, "null state pointer");
  return on_every_device(m, [&](int g) {
    const size_t off = (size_t)m->first[(size_t)g] * (size_t)m->n;
    return pmc_set_state_all(m->shard[(size_t)g], phi + off, theta + off);
  });
}

int32_t pmc_multi_get_state_all(pmc_multi* m, double* phi, double* theta) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!phi || !theta) return fail(PMC_ERR_INVALID, "null state pointer");
  return on_every_device(m, [&](int g) {
    const size_t off = (size_t)m->first[(size_t)g] * (size_t)m->n;
    return pmc_get_state_all(m->shard[(size_t)g], phi + off, theta + off);
  });
}

int64_t pmc_multi_rows_for(const pmc_multi* m, int64_t nsteps, int64_t stepout) {
  return (m && !m->shard.empty()) ? pmc_rows_for(m->shard[0], nsteps, stepout) : 0;
}

// ex = 0: pmc_run (17 rolling columns); ex = 1: pmc_run_ex (19 columns + state rows).
static int run_async(pmc_multi* m, int64_t nsteps, int64_t stepout, double* traj, double* roll, double* state, int ex) {
  int rc = check_multi(m);
  if (rc) return rc;
  const size_t G = m->shard.size();
  const int64_t rows = pmc_multi_rows_for(m, nsteps, stepout);
  const int rc_cols = ex ? 19 : 17;
  m->worker_rc.assign(G, 0);
  m->worker_err.assign(G, std::string());
  m->running = true;
  for (size_t g = 0; g < G; ++g)
    m->worker.emplace_back([=]() {
      // every shard writes its own slice of the caller's [chains][rows][cols] buffers
      const size_t c0 = (size_t)m->first[g];
      double* t = traj ? traj + c0 * (size_t)rows * 8 : nullptr;
      double* r = roll ? roll + c0 * (size_t)rows * (size_t)rc_cols : nullptr;
      double* s = state ? state + c0 * (size_t)rows * 2 * (size_t)m->n : nullptr;
      const int w = ex ? pmc_run_ex(m->shard[g], nsteps, stepout, t, r, s) : pmc_run(m->shard[g], nsteps, stepout, t, r);
      m->worker_rc[g] = w;
      if (w) m->worker_err[g] = pmc_last_error();
    });
  return PMC_OK;
}

int32_t pmc_multi_run_async(pmc_multi* m, int64_t nsteps, int64_t stepout, double* traj, double* roll) {
  return run_async(m, nsteps, stepout, traj, roll, nullptr, 0);
}

int32_t pmc_multi_wait(pmc_multi* m) {
  if (!m) return fail(PMC_ERR_INVALID, "null multi-device handle");
  if (!m->running) return PMC_OK;
  return join_workers(m);
}

int32_t pmc_multi_run(pmc_multi* m, int64_t nsteps, int64_t stepout, double* traj, double* roll) {
  int rc = run_async(m, nsteps, stepout, traj, roll, nullptr, 0);
  if (rc) return rc;
  return join_workers(m);
}

int32_t pmc_multi_run_ex(pmc_multi* m, int64_t nsteps, int64_t stepout, double* traj, double* roll19, double* state) {
  int rc = run_async(m, nsteps, stepout, traj, roll19, state, 1);
  if (rc) return rc;
  return join_workers(m);
}

int32_t pmc_multi_last_run_ms(const pmc_multi* m, float* ms_max) {
  if (!m || !ms_max) return fail(PMC_ERR_INVALID, "null argument");
  float mx = 0.f;
  for (pmc_handle* h : m->shard) {
    float t = 0.f;
    pmc_last_run_ms(h, &t);
    mx = std::max(mx, t);
  }
  *ms_max = mx;
  return PMC_OK;
}

int64_t pmc_multi_launch_count(const pmc_multi* m) {
  int64_t s = 0;
  if (m)
    for (pmc_handle* h : m->shard) s += pmc_launch_count(h);
  return s;
}

int32_t pmc_multi_gather(pmc_multi* m, double* table) {
  int rc = check_multi(m);
  if (rc) return rc;
  if (!table) return fail(PMC_ERR_INVALID, "null table");
  const int G = (int)m->shard.size();
  const size_t row = PMC_RESULT_COLS, slab = (size_t)m->cmax * row;
  // pack on every device
  for (int g = 0; g < G; ++g) {
    pmc_handle* h = m->shard[(size_t)g];
    PMC_CU(cudaSetDevice(m->devices[(size_t)g]));
    PMC_CU(cudaStreamSynchronize(h->stream));
    const int tb = 128;
    k_pack_results<<<(unsigned)((h->nchains + tb - 1) / tb), tb, 0, m->stream[(size_t)g]>>>(
        h->dyn, h->dynx, (int)h->nchains, m->res_local[(size_t)g]);
    ++h->launches;
    PMC_CU(cudaGetLastError());
  }
  if (G == 1) {
    PMC_CU(cudaMemcpyAsync(table, m->res_local[0], (size_t)m->nchains * row * sizeof(double), cudaMemcpyDeviceToHost,
                           m->stream[0]));
    PMC_CU(cudaStreamSynchronize(m->stream[0]));
    return PMC_OK;
  }
  if (m->use_nccl) {  // the one collective of the path: final averages only
    ncclResult_t r = m->nccl.GroupStart();
    for (int g = 0; g < G && r == ncclSuccess; ++g)
      r = m->nccl.AllGather(m->res_local[(size_t)g], m->res_all[(size_t)g], slab, ncclDouble, m->comm[(size_t)g],
                            m->stream[(size_t)g]);
    const ncclResult_t r2 = m->nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return fail(PMC_ERR_CUDA, std::string("ncclAllGather: ") + m->nccl.GetErrorString(r));
    for (int g = 0; g < G; ++g) {
      PMC_CU(cudaSetDevice(m->devices[(size_t)g]));
      PMC_CU(cudaStreamSynchronize(m->stream[(size_t)g]));
    }
  } else {  // device-to-device copies into the first device's table
    for (int g = 0; g < G; ++g) {
      PMC_CU(cudaSetDevice(m->devices[(size_t)g]));
      PMC_CU(cudaStreamSynchronize(m->stream[(size_t)g]));
    }
    PMC_CU(cudaSetDevice(m->devices[0]));
    for (int g = 0; g < G; ++g)
      PMC_CU(cudaMemcpyPeerAsync(m->res_all[0] + (size_t)g * slab, m->devices[0], m->res_local[(size_t)g],
                                 m->devices[(size_t)g], slab * sizeof(double), m->stream[0]));
    PMC_CU(cudaStreamSynchronize(m->stream[0]));
  }
  // every device (NCCL) / the first device (copies) now holds the whole table; read it from the first, dropping the padding
  PMC_CU(cudaSetDevice(m->devices[0]));
  for (int g = 0; g < G; ++g)
    PMC_CU(cudaMemcpyAsync(table + (size_t)m->first[(size_t)g] * row, m->res_all[0] + (size_t)g * slab,
                           (size_t)m->count[(size_t)g] * row * sizeof(double), cudaMemcpyDeviceToHost, m->stream[0]));
  PMC_CU(cudaStreamSynchronize(m->stream[0]));
  return PMC_OK;
}

}  // extern "C"
