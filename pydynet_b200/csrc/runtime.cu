// runtime.cu — device/stream/allocator/event/graph plumbing of libpdn_b200.so.
// Replaces what the reference gets from CuPy's runtime + memory pool (reference pydynet/cuda.py:16-32,
// tensor.py:80,90): nothing here is ported, the reference has no native runtime.
#include <cstring>
#include "common.cuh"
#include <nvtx3/nvToolsExt.h>  // header-only: resolves the profiler's injection library at run time, no link dependency

#include <stdarg.h>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace pdn {

static thread_local char g_err[1024] = "";
static uint64_t          g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? PDN_ERR_OOM : PDN_ERR_CUDA;
}

static bool g_nvtx = getenv("PDN_NVTX") != nullptr;  // PDN_NVTX=1: every kernel launch is marked by its entry-point name

static char     g_watch[64] = "";
static uint64_t g_watched = 0;

int after_launch(const char* name) {
  ++g_launches;
  if (g_watch[0] && strcmp(g_watch, name) == 0) ++g_watched;
  if (g_nvtx) nvtxMarkA(name);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("kernel launch `%s` failed: %s", name, cudaGetErrorString(e));
    return PDN_ERR_CUDA;
  }
  return 0;
}

// ---- per-device state -----------------------------------------------------------------------
struct DevState {
  bool         inited = false;
  cudaStream_t compute = nullptr;
  cudaStream_t copy = nullptr;     // H2D prefetch engine of the input pipeline (pdn_prefetch_h2d)
  cudaEvent_t  copy_fence = nullptr;
  cudaStream_t comm = nullptr;
  int          sms = 0;
  // caching allocator: free blocks by size; live blocks by pointer
  struct Block {
    size_t size;
    int    graph;      // 0 = ordinary; else id of the CUDA graph whose kernels reference this block
    bool   user_live;  // still owned by a caller (false: only the graph keeps it reserved)
    int    fork = 0;   // fork region (pdn_branch_begin) and branch the block was allocated in while recording; 0 = none
    int    branch = 0;
  };
  std::multimap<size_t, void*>     free_blocks;
  std::unordered_map<void*, Block> live;
  std::multimap<size_t, void*>     capture_free;  // blocks freed DURING the active capture: reusable inside it only
  // recorded fork region (pdn_branch_begin .. pdn_branch_end): the branches run concurrently when the graph replays, so a block
  // freed in a branch may only be handed out again in THAT branch, and only if it was allocated there (nobody else can hold it);
  // everything else waits for the join
  std::multimap<size_t, void*> branch_free[4];
  std::vector<std::pair<size_t, void*>> parked;
  cudaStream_t                 branch[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t>     edge_events;  // fork / join / ordering edges of recorded branches
  size_t                       edge_next = 0;
  uint64_t in_use = 0, cached = 0, n_malloc = 0;
};
static DevState   g_dev[16];
static std::mutex g_mu;
static bool       g_capturing = false;
static int        g_capture_id = 0;   // id of the graph being captured
static int        g_next_graph_id = 1;
static std::unordered_map<void*, int> g_exec_graph;  // cudaGraphExec_t -> graph id
static std::unordered_map<void*, uint64_t> g_exec_kernels;  // kernels recorded in the graph (for the launch counter)
static uint64_t g_capture_launch0 = 0;
static int      g_nbranch = 1, g_branch = 0, g_fork = 0, g_next_fork = 1;  // recorded fork region: see pdn_branch_begin

static DevState* cur() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 16) return nullptr;
  return &g_dev[d];
}

int ensure_init() {
  DevState* s = cur();
  if (!s) {
    set_error("no CUDA device available (cudaGetDevice failed) — libpdn_b200 has no CPU fallback");
    return PDN_ERR_CUDA;
  }
  if (s->inited) return 0;
  int d = 0;
  PDN_CUDA(cudaGetDevice(&d));
  cudaDeviceProp prop;
  PDN_CUDA(cudaGetDeviceProperties(&prop, d));
  s->sms = prop.multiProcessorCount;
  PDN_CUDA(cudaStreamCreateWithFlags(&s->compute, cudaStreamNonBlocking));
  PDN_CUDA(cudaStreamCreateWithFlags(&s->comm, cudaStreamNonBlocking));
  PDN_CUDA(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
  PDN_CUDA(cudaEventCreateWithFlags(&s->copy_fence, cudaEventDisableTiming));
  s->inited = true;
  return 0;
}

cudaStream_t stream() {
  DevState* s = cur();
  if (!s || !s->inited) return nullptr;
  return g_branch > 0 ? s->branch[g_branch] : s->compute;
}
cudaStream_t comm_stream() {
  DevState* s = cur();
  return s && s->inited ? s->comm : nullptr;
}
int sm_count() {
  DevState* s = cur();
  return s && s->sms > 0 ? s->sms : 148;
}

static size_t round_size(size_t b) {
  if (b == 0) b = 1;
  if (b <= (1u << 20)) return (b + 511) & ~(size_t)511;            // 512 B granules up to 1 MiB
  return (b + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);  // 2 MiB granules above (one TLB page)
}

static void release_cached(DevState* s) {
  for (auto& kv : s->free_blocks) cudaFree(kv.second);
  s->free_blocks.clear();
  s->cached = 0;
}

int dev_alloc(void** p, size_t bytes) {
  PDN_TRY(ensure_init());
  DevState* s = cur();
  size_t    want = round_size(bytes);
  std::lock_guard<std::mutex> lk(g_mu);
  const int gid = g_capturing ? g_capture_id : 0;
  if (g_capturing) {  // memory released earlier in this capture is ordered before us inside the graph: reuse it first
    auto& pool = g_fork ? s->branch_free[g_branch] : s->capture_free;
    auto  it = pool.lower_bound(want);
    if (it != pool.end() && it->first <= want + want / 4) {
      *p = it->second;
      pool.erase(it);
      auto& blk = s->live[*p];
      blk.user_live = true;
      blk.fork = g_fork, blk.branch = g_branch;
      return 0;
    }
  }
  auto it = s->free_blocks.lower_bound(want);
  // accept a cached block only if it wastes < 25 % (or is an exact granule match)
  if (it != s->free_blocks.end() && (it->first == want || it->first <= want + want / 4)) {
    *p = it->second;
    size_t got = it->first;
    s->free_blocks.erase(it);
    s->cached -= got;
    s->live[*p] = DevState::Block{got, gid, true, g_fork, g_branch};
    s->in_use += got;
    return 0;
  }
  // (cudaMalloc is legal during a capture started in relaxed mode; it is not a stream operation)
  cudaError_t e = cudaMalloc(p, want);
  if (e == cudaErrorMemoryAllocation && !g_capturing) {
    cudaGetLastError();
    cudaStreamSynchronize(s->compute);
    release_cached(s);
    e = cudaMalloc(p, want);
  }
  if (e != cudaSuccess) {
    *p = nullptr;
    return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
  }
  s->n_malloc++;
  s->live[*p] = DevState::Block{want, gid, true, g_fork, g_branch};
  s->in_use += want;
  return 0;
}

static void free_in(DevState& d, std::unordered_map<void*, DevState::Block>::iterator it) {
  void* p = it->first;
  if (it->second.graph != 0) {
    // referenced by a CUDA graph: never hand it to unrelated work while the graph exists
    it->second.user_live = false;
    if (g_capturing && it->second.graph == g_capture_id) {
      if (!g_fork) d.capture_free.emplace(it->second.size, p);
      else if (it->second.fork == g_fork && it->second.branch == g_branch) d.branch_free[g_branch].emplace(it->second.size, p);
      else d.parked.emplace_back(it->second.size, p);  // may still be read by another branch: reusable after the join
    }
    return;
  }
  // single compute stream => a freed block can be handed out again immediately (stream order)
  d.free_blocks.emplace(it->second.size, p);
  d.cached += it->second.size;
  d.in_use -= it->second.size;
  d.live.erase(it);
}

size_t dev_block_size(void* p) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& d : g_dev) {
    auto it = d.live.find(p);
    if (it != d.live.end()) return it->second.size;
  }
  return 0;
}

bool is_capturing() { return g_capturing; }

void dev_free(void* p) {
  if (!p) return;
  DevState* s = cur();
  if (!s) return;
  plane_cache_drop_range(p, dev_block_size(p));  // cached GEMM operand planes derived from this block die with it
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = s->live.find(p);
  if (it != s->live.end()) {
    free_in(*s, it);
    return;
  }
  for (auto& d : g_dev) {  // allocated on another device
    auto jt = d.live.find(p);
    if (jt != d.live.end()) {
      free_in(d, jt);
      return;
    }
  }
}

// called when a graph is destroyed: blocks it kept reserved go back to the cache (or to their still-living owner)
static void release_graph_blocks(int gid) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& d : g_dev) {
    for (auto it = d.live.begin(); it != d.live.end();) {
      if (it->second.graph == gid) {
        it->second.graph = 0;
        if (!it->second.user_live) {
          d.free_blocks.emplace(it->second.size, it->first);
          d.cached += it->second.size;
          d.in_use -= it->second.size;
          it = d.live.erase(it);
          continue;
        }
      }
      ++it;
    }
  }
}

}  // namespace pdn

using namespace pdn;

extern "C" {

const char* pdn_last_error(void) { return g_err; }

int pdn_device_count(int* n) {
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *n = 0;
  }
  return 0;
}

int pdn_init(int device) {
  PDN_CUDA(cudaSetDevice(device));
  return ensure_init();
}
int pdn_set_device(int device) {
  PDN_CUDA(cudaSetDevice(device));
  return ensure_init();
}
int pdn_get_device(int* device) {
  PDN_CUDA(cudaGetDevice(device));
  return 0;
}
int pdn_device_name(char* buf, int buflen) {
  int d = 0;
  PDN_CUDA(cudaGetDevice(&d));
  cudaDeviceProp prop;
  PDN_CUDA(cudaGetDeviceProperties(&prop, d));
  snprintf(buf, buflen, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  return 0;
}
int pdn_sm_count(int* n) {
  PDN_TRY(ensure_init());
  *n = sm_count();
  return 0;
}

int pdn_malloc(void** p, size_t bytes) { return dev_alloc(p, bytes); }
int pdn_free(void* p) {
  dev_free(p);
  return 0;
}
int pdn_malloc_host(void** p, size_t bytes) {
  PDN_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
  return 0;
}
int pdn_free_host(void* p) {
  PDN_CUDA(cudaFreeHost(p));
  return 0;
}
int pdn_mem_stats(uint64_t* bytes_in_use, uint64_t* bytes_cached, uint64_t* n_cuda_malloc) {
  PDN_TRY(ensure_init());
  DevState* s = cur();
  *bytes_in_use = s->in_use;
  *bytes_cached = s->cached;
  *n_cuda_malloc = s->n_malloc;
  return 0;
}
int pdn_empty_cache(void) {
  PDN_TRY(ensure_init());
  DevState* s = cur();
  PDN_CUDA(cudaStreamSynchronize(s->compute));
  std::lock_guard<std::mutex> lk(g_mu);
  release_cached(s);
  return 0;
}

int pdn_memcpy_h2d(void* dst, const void* src, size_t bytes) {
  PDN_TRY(ensure_init());
  if (!bytes) return 0;
  // pageable source: cudaMemcpyAsync stages and returns once the source is consumed
  PDN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
  PDN_CUDA(cudaStreamSynchronize(stream()));
  return 0;
}
int pdn_memcpy_d2h(void* dst, const void* src, size_t bytes) {
  PDN_TRY(ensure_init());
  if (bytes) PDN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
  PDN_CUDA(cudaStreamSynchronize(stream()));
  return 0;
}
int pdn_memcpy_d2d(void* dst, const void* src, size_t bytes) {
  PDN_TRY(ensure_init());
  if (bytes) PDN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream()));
  return 0;
}
int pdn_memcpy_h2d_async(void* dst, const void* src, size_t bytes) {
  PDN_TRY(ensure_init());
  if (bytes) PDN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream()));
  return 0;
}
// Input pipeline: copy a batch from PINNED host memory into `dst` on a dedicated copy stream, overlapping the compute stream's
// kernels. `dst` comes from the stream-ordered allocator (a block may have been handed back by compute work that is still queued),
// so the copy first waits for everything queued on the compute stream so far; `done` (an event from pdn_event_create) is recorded
// after the copy: the consumer calls pdn_stream_wait_event(done) before its first kernel that reads dst, and
// pdn_event_synchronize(done) before the pinned staging buffer is overwritten.
int pdn_prefetch_h2d(void* dst, const void* pinned_src, size_t bytes, void* done) {
  PDN_TRY(ensure_init());
  DevState* s = cur();
  PDN_CHECK(done != nullptr, "prefetch_h2d: null event");
  PDN_CUDA(cudaEventRecord(s->copy_fence, s->compute));
  PDN_CUDA(cudaStreamWaitEvent(s->copy, s->copy_fence, 0));
  if (bytes) PDN_CUDA(cudaMemcpyAsync(dst, pinned_src, bytes, cudaMemcpyHostToDevice, s->copy));
  PDN_CUDA(cudaEventRecord((cudaEvent_t)done, s->copy));
  return 0;
}
int pdn_stream_wait_event(void* ev) {
  PDN_TRY(ensure_init());
  PDN_CUDA(cudaStreamWaitEvent(stream(), (cudaEvent_t)ev, 0));
  return 0;
}
int pdn_event_synchronize(void* ev) {
  PDN_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return 0;
}
int pdn_memcpy_d2h_async(void* dst, const void* src, size_t bytes) {
  PDN_TRY(ensure_init());
  if (bytes) PDN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream()));
  return 0;
}
int pdn_memset(void* dst, int byte, size_t bytes) {
  PDN_TRY(ensure_init());
  if (bytes) PDN_CUDA(cudaMemsetAsync(dst, byte, bytes, stream()));
  return 0;
}
int pdn_sync(void) {
  PDN_TRY(ensure_init());
  PDN_CUDA(cudaStreamSynchronize(stream()));
  PDN_CUDA(cudaStreamSynchronize(comm_stream()));
  return 0;
}

uint64_t pdn_kernel_launch_count(void) { return g_launches; }
void     pdn_reset_launch_count(void) { g_launches = 0; }
void     pdn_watch_launches(const char* name) {
  g_watched = 0;
  snprintf(g_watch, sizeof g_watch, "%s", name ? name : "");
}
uint64_t pdn_watched_launch_count(void) { return g_watched; }

int pdn_event_create(void** ev) {
  PDN_TRY(ensure_init());
  cudaEvent_t e;
  PDN_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return 0;
}
int pdn_event_destroy(void* ev) {
  PDN_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
int pdn_event_record(void* ev) {
  PDN_TRY(ensure_init());
  // while a CUDA graph is being recorded the event becomes an event-record NODE of the graph (cudaEventRecordExternal): every
  // replay stamps it, so a kernel inside a replayed graph can be timed with an ordinary event pair
  PDN_CUDA(cudaEventRecordWithFlags((cudaEvent_t)ev, stream(), g_capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  return 0;
}
int pdn_event_elapsed_ms(void* start, void* stop, float* ms) {
  PDN_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  PDN_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}

// ---- recorded fork regions ---------------------------------------------------------------------------------------------------
// Inside a graph recording, pdn_branch_begin(n) forks the compute stream into n branches (branch 0 is the compute stream itself),
// pdn_branch_select(i) makes branch i the stream every following launch / allocation of this thread goes to, pdn_branch_mark /
// pdn_branch_wait add an edge from a marked point of one branch to what another does next, pdn_branch_end joins. When the graph replays
// the branches run concurrently: a latency-bound launch chain of one batch slice fills the gaps of another's. Outside a recording
// the calls are no-ops (everything stays on the one compute stream, which is what the caching allocator's reuse rule assumes).
static int next_edge_event(DevState* s, cudaEvent_t* ev) {
  if (s->edge_next == s->edge_events.size()) {
    cudaEvent_t e;
    PDN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s->edge_events.push_back(e);
  }
  *ev = s->edge_events[s->edge_next++];
  return 0;
}
int pdn_branch_begin(int n) {
  PDN_TRY(ensure_init());
  PDN_CHECK(n >= 1 && n <= 4, "pdn_branch_begin: 1..4 branches");
  if (!g_capturing || n == 1) return 0;
  PDN_CHECK(g_fork == 0, "pdn_branch_begin: fork regions do not nest");
  DevState*   s = cur();
  cudaEvent_t ev;
  PDN_TRY(next_edge_event(s, &ev));
  PDN_CUDA(cudaEventRecord(ev, s->compute));
  for (int i = 1; i < n; ++i) {
    if (!s->branch[i]) PDN_CUDA(cudaStreamCreateWithFlags(&s->branch[i], cudaStreamNonBlocking));
    PDN_CUDA(cudaStreamWaitEvent(s->branch[i], ev, 0));
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_nbranch = n, g_branch = 0, g_fork = g_next_fork++;
  return 0;
}
int pdn_branch_select(int i) {
  if (!g_fork) return 0;
  PDN_CHECK(i >= 0 && i < g_nbranch, "pdn_branch_select: branch %d of %d", i, g_nbranch);
  std::lock_guard<std::mutex> lk(g_mu);
  g_branch = i;
  return 0;
}
int pdn_branch_mark(int* token) {  // an edge source at the current tail of the selected branch
  *token = -1;
  if (!g_fork) return 0;
  DevState*   s = cur();
  cudaEvent_t ev;
  PDN_TRY(next_edge_event(s, &ev));
  PDN_CUDA(cudaEventRecord(ev, stream()));
  *token = (int)s->edge_next - 1;
  return 0;
}
int pdn_branch_wait(int token) {  // what the selected branch does next waits for the marked point
  if (!g_fork || token < 0) return 0;
  DevState* s = cur();
  PDN_CHECK((size_t)token < s->edge_next, "pdn_branch_wait: unknown token %d", token);
  PDN_CUDA(cudaStreamWaitEvent(stream(), s->edge_events[token], 0));
  return 0;
}
int pdn_branch_end(void) {
  if (!g_fork) return 0;
  DevState* s = cur();
  for (int i = 1; i < g_nbranch; ++i) {
    cudaEvent_t ev;
    PDN_TRY(next_edge_event(s, &ev));
    PDN_CUDA(cudaEventRecord(ev, s->branch[i]));
    PDN_CUDA(cudaStreamWaitEvent(s->compute, ev, 0));
  }
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < 4; ++i) {  // after the join everything the branches released is ordered before what follows
    for (auto& kv : s->branch_free[i]) s->capture_free.emplace(kv.first, kv.second);
    s->branch_free[i].clear();
  }
  for (auto& kv : s->parked) s->capture_free.emplace(kv.first, kv.second);
  s->parked.clear();
  g_nbranch = 1, g_branch = 0, g_fork = 0;
  return 0;
}

int pdn_graph_begin(void) {
  PDN_TRY(ensure_init());
  PDN_CHECK(!g_capturing, "graph capture already active");
  PDN_CUDA(cudaStreamBeginCapture(stream(), cudaStreamCaptureModeRelaxed));
  std::lock_guard<std::mutex> lk(g_mu);
  g_capturing = true;
  g_capture_id = g_next_graph_id++;
  g_capture_launch0 = g_launches;
  if (DevState* s = cur()) s->edge_next = 0;  // edge events are only graph edges: reusable by every recording
  return 0;
}
/* debugging aid: 0 = no capture, 1 = capture active, 2 = capture invalidated by an operation that cannot be recorded */
int pdn_graph_status(int* status) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaError_t e = cudaStreamIsCapturing(stream(), &st);
  if (e != cudaSuccess) cudaGetLastError();
  *status = (e != cudaSuccess || st == cudaStreamCaptureStatusInvalidated) ? 2 : (st == cudaStreamCaptureStatusActive ? 1 : 0);
  return 0;
}
int pdn_graph_end(void** graph_exec) {
  PDN_CHECK(g_capturing, "no graph capture active");
  int gid;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_capturing = false;
    gid = g_capture_id;
    g_capture_id = 0;
    for (auto& d : g_dev) d.capture_free.clear();  // stay reserved for the graph (user_live == false)
  }
  PDN_CHECK(g_fork == 0, "graph capture ended inside a fork region (pdn_branch_end missing)");
  cudaGraph_t g;
  cudaError_t e = cudaStreamEndCapture(stream(), &g);
  if (e != cudaSuccess) {
    release_graph_blocks(gid);
    return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
  }
  cudaGraphExec_t ge;
  e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    release_graph_blocks(gid);
    return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_exec_graph[(void*)ge] = gid;
    g_exec_kernels[(void*)ge] = g_launches - g_capture_launch0;
    g_launches = g_capture_launch0;  // recording is not launching
  }
  *graph_exec = (void*)ge;
  return 0;
}
int pdn_graph_launch(void* graph_exec) {
  PDN_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, stream()));
  auto it = g_exec_kernels.find(graph_exec);
  g_launches += it != g_exec_kernels.end() ? it->second : 1;  // every kernel node of the replayed graph is one of ours
  return 0;
}
int pdn_graph_destroy(void* graph_exec) {
  PDN_CUDA(cudaStreamSynchronize(stream()));
  PDN_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  int gid = 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_exec_graph.find(graph_exec);
    if (it != g_exec_graph.end()) { gid = it->second; g_exec_graph.erase(it); }
  }
  if (gid) release_graph_blocks(gid);
  return 0;
}

}  // extern "C"

// ---- shared host helper: dimension collapsing -------------------------------------------------
namespace pdn {
int make_desc(int ndim, const int64_t* shape, int nops, const int64_t* const* strides, StridedDesc* out) {
  PDN_CHECK(ndim >= 0 && ndim <= PDN_MAXD, "ndim %d out of range (max %d)", ndim, PDN_MAXD);
  PDN_CHECK(nops >= 1 && nops <= 4, "bad operand count");
  int64_t sh[PDN_MAXD], st[4][PDN_MAXD];
  int     nd = 0;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    PDN_CHECK(shape[i] >= 0, "negative dimension");
    n *= shape[i];
    if (shape[i] == 1) continue;  // size-1 dims carry no information
    sh[nd] = shape[i];
    for (int o = 0; o < nops; ++o) st[o][nd] = strides[o][i];
    ++nd;
  }
  // merge dim k+1 into k when for every operand stride[k] == shape[k+1]*stride[k+1]
  int w = 0;
  for (int k = 0; k < nd; ++k) {
    if (w > 0) {
      bool merge = true;
      for (int o = 0; o < nops; ++o)
        if (st[o][w - 1] != sh[k] * st[o][k]) merge = false;
      if (merge) {
        sh[w - 1] *= sh[k];
        for (int o = 0; o < nops; ++o) st[o][w - 1] = st[o][k];
        continue;
      }
    }
    sh[w] = sh[k];
    for (int o = 0; o < nops; ++o) st[o][w] = st[o][k];
    ++w;
  }
  out->ndim = w;
  out->n = n;
  for (int k = 0; k < PDN_MAXD; ++k) {
    out->shape[k] = k < w ? sh[k] : 1;
    for (int o = 0; o < 4; ++o) out->s[o][k] = (k < w && o < nops) ? st[o][k] : 0;
  }
  return 0;
}
}  // namespace pdn

// ---- NVTX ranges (profiling aid): phases of a step as named ranges for ncu --nvtx / nsys -------------------------
extern "C" int pdn_nvtx_push(const char* name) {
  nvtxRangePushA(name ? name : "pdn");
  return 0;
}
extern "C" int pdn_nvtx_pop(void) {
  nvtxRangePop();
  return 0;
}
