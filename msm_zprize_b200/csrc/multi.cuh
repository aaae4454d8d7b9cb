// Multi-GPU MSM behind the C ABI (include/msm_b200.h, msm_b200_multi_*): ONE call that fans out over the
// GPUs of a box and returns one point -- the GPU analogue of the reference's SPMD call, in which every pool
// thread takes a static range of the points (`range()`, src/threads/threads.ts:354-359) and the main thread
// adds the partition sums (src/msm-batched-affine.ts:294-322).
//
//   * one msm_b200_ctx per device, one persistent host thread per device (each MSM has one host
//     round-trip -- the bucket totals -- so the devices must not be driven from a single thread);
//   * the point set shards by contiguous range: device g owns [g * per, min(n, (g + 1) * per)), per = ceil(n / G);
//   * every device runs the complete single-GPU pipeline on its range and leaves one partial point
//     (144 / 128 bytes) in its own memory;
//   * the partials are gathered with ONE ncclAllGather over NVLink (single-process communicators from
//     ncclCommInitAll; libnccl.so.2 is opened at run time so that the library still loads where NCCL is
//     absent) or, when NCCL is unavailable, with G - 1 peer copies ordered by events;
//   * device 0 adds the G partials and normalises (k_finalize).
// Included by api.cu.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!handle) return false;
    CommInitAll = (decltype(CommInitAll))dlsym(handle, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
    GroupStart = (decltype(GroupStart))dlsym(handle, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(handle, "ncclGroupEnd");
    GetVersion = (decltype(GetVersion))dlsym(handle, "ncclGetVersion");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetVersion && GetErrorString;
  }
};

// one persistent host thread per device
struct DeviceWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool pending = false, quit = false, idle = true;
  int rc = 0;
  void start() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return pending || quit; });
        if (quit) return;
        std::function<int()> f = std::move(job);
        pending = false;
        lk.unlock();
        int r = f();
        lk.lock();
        rc = r;
        idle = true;
        cv.notify_all();
      }
    });
  }
  void post(std::function<int()> f) {
    std::lock_guard<std::mutex> lk(m);
    job = std::move(f);
    pending = true;
    idle = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [this] { return idle; });
    return rc;
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(m);
      quit = true;
      cv.notify_all();
    }
    if (th.joinable()) th.join();
  }
};

struct msm_b200_multi {
  int curve = 0;
  int n_dev = 0;
  std::vector<int> devices;
  std::vector<msm_b200_ctx*> ctx;
  std::vector<DeviceWorker*> workers;
  std::vector<void*> part;        // per device: its partial
  std::vector<void*> gathered;    // per device: n_dev partials (only device 0's is read)
  std::vector<cudaEvent_t> ready; // per device: partial has reached device 0 (peer-copy path)
  std::vector<size_t> lo, cnt;    // range of the resident bases owned by each device
  size_t n_bases = 0;
  NcclApi nccl;
  std::vector<ncclComm_t> comms;
  bool use_nccl = false;
  std::string gather_kind, err;
  std::vector<msm_b200_timing> last;
  std::mutex call_mutex;  // single caller at a time
};

static int mfail(msm_b200_multi* m, int code, const std::string& msg) {
  if (m) m->err = msg;
  g_err = msg;
  return code;
}

// [lo, lo + cnt) of device g in a set of n: ceil(n / G) items each (src/threads/threads.ts:354-359)
static void shard_range(size_t n, int g, int G, size_t& lo, size_t& cnt) {
  size_t per = (n + G - 1) / G;
  lo = std::min(n, per * (size_t)g);
  cnt = std::min(n, lo + per) - lo;
}

static int multi_run_all(msm_b200_multi* m, const std::function<int(int)>& f) {
  for (int g = 0; g < m->n_dev; g++) m->workers[g]->post([f, g] { return f(g); });
  int rc = 0;
  for (int g = 0; g < m->n_dev; g++) {
    int r = m->workers[g]->wait();
    if (r != 0 && rc == 0) {
      rc = r;
      m->err = "device " + std::to_string(m->devices[g]) + ": " + m->ctx[g]->err;
      g_err = m->err;
    }
  }
  return rc;
}

// partials (already in m->part[g], stream order) -> device 0 -> one point
static int multi_gather_combine(msm_b200_multi* m, msm_b200_point* out) {
  msm_b200_ctx* ctx = m->ctx[0];
  const size_t pb = partial_bytes(m->curve);
  if (m->n_dev == 1) return combine_impl(m->ctx[0], m->part[0], 1, out);
  if (m->use_nccl) {
    ncclResult_t r = m->nccl.GroupStart();
    for (int g = 0; g < m->n_dev && r == ncclSuccess; g++)
      r = m->nccl.AllGather(m->part[g], m->gathered[g], pb, ncclUint8, m->comms[g], m->ctx[g]->stream);
    ncclResult_t r2 = m->nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return mfail(m, MSM_E_CUDA, std::string("ncclAllGather: ") + m->nccl.GetErrorString(r));
  } else {
    CK(cudaSetDevice(m->devices[0]));
    CK(cudaMemcpyAsync(m->gathered[0], m->part[0], pb, cudaMemcpyDeviceToDevice, m->ctx[0]->stream));
    for (int g = 1; g < m->n_dev; g++) {
      CK(cudaSetDevice(m->devices[g]));
      CK(cudaMemcpyPeerAsync((char*)m->gathered[0] + g * pb, m->devices[0], m->part[g], m->devices[g], pb, m->ctx[g]->stream));
      CK(cudaEventRecord(m->ready[g], m->ctx[g]->stream));
    }
    CK(cudaSetDevice(m->devices[0]));
    for (int g = 1; g < m->n_dev; g++) CK(cudaStreamWaitEvent(m->ctx[0]->stream, m->ready[g], 0));
  }
  return combine_impl(m->ctx[0], m->gathered[0], m->n_dev, out);
}

// slowest device = the call's critical path; launches and additions are summed over the devices
static void multi_timing(msm_b200_multi* m, msm_b200_timing* tm, float total_ms) {
  m->last.assign(m->n_dev, msm_b200_timing());
  int slow = 0;
  float slow_ms = -1;
  int launches = 0, hot_launches = 0;
  unsigned long long adds = 0;
  for (int g = 0; g < m->n_dev; g++) {
    msm_b200_last_timing(m->ctx[g], &m->last[g]);
    const msm_b200_timing& t = m->last[g];
    float ms = t.h2d_ms + t.digits_ms + t.sort_ms + t.accumulate_ms + t.reduce_ms;
    if (ms > slow_ms) slow_ms = ms, slow = g;
    launches += t.kernel_launches;
    hot_launches += t.hot_kernel_launches;
    adds += t.n_adds;
  }
  if (!tm) return;
  *tm = m->last[slow];
  tm->kernel_launches = launches;
  tm->hot_kernel_launches = hot_launches;
  tm->n_adds = adds;
  tm->total_ms = total_ms;
}

extern "C" {

int msm_b200_multi_create(msm_b200_multi** out, int curve, const int* devices, int n_dev) {
  if (!out) return mfail(nullptr, MSM_E_INVALID, "null out pointer");
  *out = nullptr;
  if (!devices || n_dev < 1 || n_dev > 64) return mfail(nullptr, MSM_E_INVALID, "bad device list");
  for (int i = 0; i < n_dev; i++)
    for (int j = 0; j < i; j++)
      if (devices[i] == devices[j]) return mfail(nullptr, MSM_E_INVALID, "duplicate device");
  msm_b200_multi* m = new msm_b200_multi();
  m->curve = curve;
  m->n_dev = n_dev;
  m->devices.assign(devices, devices + n_dev);
  const size_t pb = partial_bytes(curve);
  int rc = 0;
  for (int g = 0; g < n_dev && rc == 0; g++) {
    msm_b200_ctx* c = nullptr;
    rc = msm_b200_create(&c, curve, devices[g], nullptr);
    if (rc != 0) break;
    m->ctx.push_back(c);
    void *p = nullptr, *q = nullptr;
    cudaEvent_t e = nullptr;
    if (cudaMalloc(&p, pb) != cudaSuccess || cudaMalloc(&q, pb * n_dev) != cudaSuccess ||
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
      rc = mfail(nullptr, MSM_E_NOMEM, "multi: device allocation failed");
    m->part.push_back(p);
    m->gathered.push_back(q);
    m->ready.push_back(e);
  }
  if (rc != 0) {
    std::string keep = g_err;
    msm_b200_multi_destroy(m);
    g_err = keep;
    return rc;
  }
  for (int g = 0; g < n_dev; g++) {  // direct NVLink access for the peer copies (ignored where unsupported)
    cudaSetDevice(devices[g]);
    for (int h = 0; h < n_dev; h++)
      if (h != g) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[g], devices[h]) == cudaSuccess && can)
          if (cudaDeviceEnablePeerAccess(devices[h], 0) != cudaSuccess) cudaGetLastError();
      }
  }
  m->gather_kind = "none (single device)";
  if (n_dev > 1) {
    const char* want = getenv("MSM_B200_GATHER");  // "peer" forces the copy path
    m->gather_kind = "peer copies (cudaMemcpyPeerAsync)";
    if (!(want && !strcmp(want, "peer")) && m->nccl.load()) {
      m->comms.assign(n_dev, nullptr);
      ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), n_dev, devices);
      if (r == ncclSuccess) {
        int v = 0;
        m->nccl.GetVersion(&v);
        m->use_nccl = true;
        m->gather_kind = "ncclAllGather (NCCL " + std::to_string(v) + ", ncclCommInitAll)";
      } else {
        m->comms.clear();
        cudaGetLastError();
      }
    }
  }
  for (int g = 0; g < n_dev; g++) {
    m->workers.push_back(new DeviceWorker());
    m->workers.back()->start();
  }
  m->lo.assign(n_dev, 0);
  m->cnt.assign(n_dev, 0);
  *out = m;
  return 0;
}

void msm_b200_multi_destroy(msm_b200_multi* m) {
  if (!m) return;
  for (DeviceWorker* w : m->workers) {
    w->stop();
    delete w;
  }
  for (size_t g = 0; g < m->ctx.size(); g++) {
    cudaSetDevice(m->devices[g]);
    cudaStreamSynchronize(m->ctx[g]->stream);
  }
  if (m->use_nccl)
    for (ncclComm_t c : m->comms)
      if (c) m->nccl.CommDestroy(c);
  for (size_t g = 0; g < m->ctx.size(); g++) {
    cudaSetDevice(m->devices[g]);
    if (g < m->part.size() && m->part[g]) cudaFree(m->part[g]);
    if (g < m->gathered.size() && m->gathered[g]) cudaFree(m->gathered[g]);
    if (g < m->ready.size() && m->ready[g]) cudaEventDestroy(m->ready[g]);
    msm_b200_destroy(m->ctx[g]);
  }
  delete m;
}

void msm_b200_multi_shard_range(size_t n, int i, int n_dev, size_t* first, size_t* count) {
  size_t lo = 0, cnt = 0;
  if (n_dev > 0 && i >= 0 && i < n_dev) shard_range(n, i, n_dev, lo, cnt);
  if (first) *first = lo;
  if (count) *count = cnt;
}

const char* msm_b200_multi_last_error(const msm_b200_multi* m) { return m ? m->err.c_str() : g_err.c_str(); }
int msm_b200_multi_devices(const msm_b200_multi* m) { return m ? m->n_dev : 0; }
const char* msm_b200_multi_gather_kind(const msm_b200_multi* m) { return m ? m->gather_kind.c_str() : ""; }
msm_b200_ctx* msm_b200_multi_ctx(msm_b200_multi* m, int i) { return (m && i >= 0 && i < m->n_dev) ? m->ctx[i] : nullptr; }

int msm_b200_multi_set_bases(msm_b200_multi* m, const void* points_host, size_t n, int layout) {
  if (!m) return mfail(nullptr, MSM_E_INVALID, "null context");
  if (n > 0 && !points_host) return mfail(m, MSM_E_INVALID, "null points");
  if (layout != MSM_LAYOUT_LIMB29_MONT && layout != MSM_LAYOUT_LE_BYTES) return mfail(m, MSM_E_INVALID, "bad point layout");
  std::lock_guard<std::mutex> lk(m->call_mutex);
  const size_t pbytes = point_bytes(m->curve, layout);
  for (int g = 0; g < m->n_dev; g++) shard_range(n, g, m->n_dev, m->lo[g], m->cnt[g]);
  m->n_bases = 0;
  int rc = multi_run_all(m, [&](int g) {
    return msm_b200_set_bases(m->ctx[g], (const char*)points_host + m->lo[g] * pbytes, m->cnt[g], layout, 0);
  });
  if (rc == 0) m->n_bases = n;
  return rc;
}

// msm_b200_share_bases for every device: `m` runs over the bases resident in `owner` (same curve, same devices in
// the same order), so that several multi contexts can have MSMs in flight over one point set
int msm_b200_multi_share_bases(msm_b200_multi* m, msm_b200_multi* owner) {
  if (!m || !owner || m == owner) return mfail(m, MSM_E_INVALID, "bad arguments");
  if (m->curve != owner->curve || m->devices != owner->devices) return mfail(m, MSM_E_INVALID, "shared bases need the same curve and devices");
  std::lock_guard<std::mutex> lk(m->call_mutex);
  std::lock_guard<std::mutex> lk2(owner->call_mutex);
  for (int g = 0; g < m->n_dev; g++) {
    int rc = msm_b200_share_bases(m->ctx[g], owner->ctx[g]);
    if (rc != 0) return mfail(m, rc, m->ctx[g]->err);
  }
  m->lo = owner->lo;
  m->cnt = owner->cnt;
  m->n_bases = owner->n_bases;
  return 0;
}

// bases that already live on the devices (benchmarks: the seeded generators): shard g is `points_dev[g]`,
// n_per_dev[g] points in device g's memory; the global order is shard 0, shard 1, ...
int msm_b200_multi_set_bases_sharded(msm_b200_multi* m, const void* const* points_dev, const size_t* n_per_dev, int layout) {
  if (!m || !points_dev || !n_per_dev) return mfail(m, MSM_E_INVALID, "bad arguments");
  std::lock_guard<std::mutex> lk(m->call_mutex);
  size_t at = 0;
  for (int g = 0; g < m->n_dev; g++) {
    m->lo[g] = at;
    m->cnt[g] = n_per_dev[g];
    at += n_per_dev[g];
  }
  m->n_bases = 0;
  int rc = multi_run_all(m, [&](int g) { return msm_b200_set_bases(m->ctx[g], points_dev[g], m->cnt[g], layout, 1); });
  if (rc == 0) m->n_bases = at;
  return rc;
}

static int multi_run_impl(msm_b200_multi* m, const void* scalars_host, const void* const* scalars_dev, const void* points_host,
                          int point_layout, size_t n, int scalar_layout, int form, int window_bits, msm_b200_point* out,
                          msm_b200_timing* timing) {
  if (!m || !out) return mfail(m, MSM_E_INVALID, "bad arguments");
  if (scalar_layout != MSM_LAYOUT_LIMB29_MONT && scalar_layout != MSM_LAYOUT_LE_BYTES)
    return mfail(m, MSM_E_INVALID, "bad scalar layout");
  if (n > 0 && !scalars_host && !scalars_dev) return mfail(m, MSM_E_INVALID, "null scalars");
  std::lock_guard<std::mutex> lk(m->call_mutex);
  auto w0 = std::chrono::steady_clock::now();
  const size_t sb = scalar_bytes(scalar_layout);
  if (points_host) {  // one-shot: the call shards and uploads the points as well
    if (point_layout != MSM_LAYOUT_LIMB29_MONT && point_layout != MSM_LAYOUT_LE_BYTES)
      return mfail(m, MSM_E_INVALID, "bad point layout");
    for (int g = 0; g < m->n_dev; g++) shard_range(n, g, m->n_dev, m->lo[g], m->cnt[g]);
    m->n_bases = n;
  }
  if (n > m->n_bases) return mfail(m, MSM_E_STATE, "more scalars than resident bases (call set_bases first)");
  const size_t pbytes = points_host ? point_bytes(m->curve, point_layout) : 0;
  int rc = multi_run_all(m, [&](int g) {
    msm_b200_ctx* c = m->ctx[g];
    // the first n points of the global order: device g takes what falls inside its range
    const size_t lo = m->lo[g], ng = n > lo ? std::min(m->cnt[g], n - lo) : 0;
    if (points_host) RET_IF(msm_b200_set_bases_async(c, (const char*)points_host + lo * pbytes, ng, point_layout));
    const void* s = scalars_dev ? scalars_dev[g] : (const void*)((const char*)scalars_host + lo * sb);
    return msm_b200_run_partial(c, ng ? s : (const void*)"", ng, scalar_layout, scalars_dev ? 1 : 0, form, window_bits,
                                m->part[g], nullptr);
  });
  RET_IF(rc);
  rc = multi_gather_combine(m, out);  // waits for device 0, which waits for every partial
  if (rc != 0) {
    if (m->err.empty() || m->ctx[0]->err.size()) m->err = m->ctx[0]->err;
    return rc;
  }
  float total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
  multi_timing(m, timing, total);
  return 0;
}

int msm_b200_multi_run(msm_b200_multi* m, const void* scalars_host, size_t n, int scalar_layout, int form, int window_bits,
                       msm_b200_point* out, msm_b200_timing* timing) {
  return multi_run_impl(m, scalars_host, nullptr, nullptr, 0, n, scalar_layout, form, window_bits, out, timing);
}

int msm_b200_multi_run_sharded(msm_b200_multi* m, const void* const* scalars_dev, int scalar_layout, int form, int window_bits,
                               msm_b200_point* out, msm_b200_timing* timing) {
  if (!m || !scalars_dev) return mfail(m, MSM_E_INVALID, "bad arguments");
  return multi_run_impl(m, nullptr, scalars_dev, nullptr, 0, m->n_bases, scalar_layout, form, window_bits, out, timing);
}

int msm_b200_multi_msm(msm_b200_multi* m, const void* scalars_host, int scalar_layout, const void* points_host, int point_layout,
                       size_t n, int form, int window_bits, msm_b200_point* out, msm_b200_timing* timing) {
  if (n > 0 && !points_host) return mfail(m, MSM_E_INVALID, "null points");
  if (n == 0) points_host = "";
  return multi_run_impl(m, scalars_host, nullptr, points_host, point_layout, n, scalar_layout, form, window_bits, out, timing);
}

int msm_b200_multi_last_timings(msm_b200_multi* m, msm_b200_timing* per_device, int count) {
  if (!m || !per_device || count < 0) return mfail(m, MSM_E_INVALID, "bad arguments");
  for (int g = 0; g < count && g < (int)m->last.size(); g++) per_device[g] = m->last[g];
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// Several MSMs in flight over one resident point set, behind plain calls (msm_b200_pipeline_*): `depth` lanes,
// each a multi context of its own (lane 0 owns the bases, the others borrow them) with a dispatcher thread.
// submit() hands a scalar vector to the next lane and returns a ticket at once; wait() blocks until that MSM is
// done and its point is in the caller's buffer.  A caller without threads of its own thus gets the overlap of
// the latency-bound phases of one MSM with the rounds of another (bench.py `pipelined`).
// ------------------------------------------------------------------------------------------
struct PipeJob {
  const void* scalars = nullptr;
  size_t n = 0;
  int layout = 0, form = 0, window_bits = 0;
  msm_b200_point* out = nullptr;
  msm_b200_timing* tm = nullptr;
  int rc = 0;
  bool done = true;
};

struct msm_b200_pipeline {
  int depth = 0;
  std::vector<msm_b200_multi*> lanes;
  std::vector<DeviceWorker*> dispatch;  // one per lane: runs that lane's MSMs in submission order
  std::vector<PipeJob> jobs;            // ring of tickets; ticket t lives in slot t % jobs.size()
  std::mutex m;         // tickets and job states
  std::mutex submit_m;  // one submitter at a time (a lane's dispatcher takes one job at a time)
  std::condition_variable cv;
  long long next_ticket = 0;
  std::string err;
};

extern "C" {

int msm_b200_pipeline_create(msm_b200_pipeline** out, int curve, const int* devices, int n_dev, int depth) {
  if (!out) return mfail(nullptr, MSM_E_INVALID, "null out pointer");
  *out = nullptr;
  if (depth < 1 || depth > 16) return mfail(nullptr, MSM_E_INVALID, "pipeline depth out of range [1,16]");
  msm_b200_pipeline* p = new msm_b200_pipeline();
  p->depth = depth;
  for (int l = 0; l < depth; l++) {
    msm_b200_multi* m = nullptr;
    int rc = msm_b200_multi_create(&m, curve, devices, n_dev);
    if (rc != 0) {
      std::string keep = g_err;
      msm_b200_pipeline_destroy(p);
      g_err = keep;
      return rc;
    }
    p->lanes.push_back(m);
    p->dispatch.push_back(new DeviceWorker());
    p->dispatch.back()->start();
  }
  p->jobs.assign((size_t)depth * 4, PipeJob());
  *out = p;
  return 0;
}

// every submitted MSM has finished
static void pipeline_drain(msm_b200_pipeline* p) {
  for (DeviceWorker* w : p->dispatch) w->wait();
}

void msm_b200_pipeline_destroy(msm_b200_pipeline* p) {
  if (!p) return;
  for (DeviceWorker* w : p->dispatch) {
    w->wait();
    w->stop();
    delete w;
  }
  for (size_t l = p->lanes.size(); l-- > 0;) msm_b200_multi_destroy(p->lanes[l]);  // the borrowers first
  delete p;
}

const char* msm_b200_pipeline_last_error(const msm_b200_pipeline* p) { return p ? p->err.c_str() : g_err.c_str(); }
int msm_b200_pipeline_depth(const msm_b200_pipeline* p) { return p ? p->depth : 0; }

int msm_b200_pipeline_set_bases(msm_b200_pipeline* p, const void* points_host, size_t n, int layout) {
  if (!p) return mfail(nullptr, MSM_E_INVALID, "null pipeline");
  std::lock_guard<std::mutex> submit_lock(p->submit_m);
  pipeline_drain(p);
  int rc = msm_b200_multi_set_bases(p->lanes[0], points_host, n, layout);
  for (int l = 1; l < p->depth && rc == 0; l++) rc = msm_b200_multi_share_bases(p->lanes[l], p->lanes[0]);
  if (rc != 0) p->err = msm_b200_global_error();
  return rc;
}

int msm_b200_pipeline_submit(msm_b200_pipeline* p, const void* scalars_host, size_t n, int scalar_layout, int form, int window_bits,
                             msm_b200_point* out, msm_b200_timing* timing, int* ticket) {
  if (!p || !out || !ticket) return mfail(nullptr, MSM_E_INVALID, "bad arguments");
  std::lock_guard<std::mutex> submit_lock(p->submit_m);
  long long t;
  PipeJob* job;
  {
    std::unique_lock<std::mutex> lk(p->m);
    t = p->next_ticket;
    if (t >= 0x7FFFFFFFll) return mfail(nullptr, MSM_E_STATE, "ticket counter exhausted: create a new pipeline");
    job = &p->jobs[(size_t)(t % (long long)p->jobs.size())];
    if (!job->done) {
      p->err = "too many MSMs in flight without msm_b200_pipeline_wait";
      g_err = p->err;
      return MSM_E_STATE;
    }
    p->next_ticket++;
    *job = PipeJob();
    job->scalars = scalars_host;
    job->n = n;
    job->layout = scalar_layout;
    job->form = form;
    job->window_bits = window_bits;
    job->out = out;
    job->tm = timing;
    job->done = false;
  }
  const int lane = (int)(t % p->depth);
  DeviceWorker* w = p->dispatch[lane];
  w->wait();  // the lane's previous MSM (its caller may not have waited for it yet)
  w->post([p, job, lane] {
    int rc = msm_b200_multi_run(p->lanes[lane], job->scalars, job->n, job->layout, job->form, job->window_bits, job->out, job->tm);
    std::lock_guard<std::mutex> lk(p->m);
    job->rc = rc;
    if (rc != 0) p->err = msm_b200_multi_last_error(p->lanes[lane]);
    job->done = true;
    p->cv.notify_all();
    return rc;
  });
  *ticket = (int)t;
  return 0;
}

int msm_b200_pipeline_wait(msm_b200_pipeline* p, int ticket) {
  if (!p || ticket < 0) return mfail(nullptr, MSM_E_INVALID, "bad arguments");
  std::unique_lock<std::mutex> lk(p->m);
  const long long t = ticket;
  if (t >= p->next_ticket || p->next_ticket - t > (long long)p->jobs.size())
    return mfail(nullptr, MSM_E_INVALID, "unknown or expired ticket");
  PipeJob* job = &p->jobs[(size_t)(t % (long long)p->jobs.size())];
  p->cv.wait(lk, [job] { return job->done; });
  return job->rc;
}

}  // extern "C"
