// Comm implementations (comm.h):
//   ThreadComm  ranks are host threads of one process whose buffers live on the same
//               device (CPU port; also used to exercise the strip kernels on one GPU)
//   NcclComm    one process per GPU, NCCL over NVLink/NVSwitch (product build only);
//               libnccl is opened lazily so that single-GPU users do not need it.
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "comm.h"

namespace gb200 {

// ---------------------------------------------------------------------------
struct ThreadGroup {
  int world;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<void*> bufs;
  explicit ThreadGroup(int w) : world(w), bufs(w, nullptr) {}
  bool aborted = false;
  // a rank that failed calls this so that the others do not wait for it forever
  void abort() {
    std::lock_guard<std::mutex> lock(mu);
    aborted = true;
    cv.notify_all();
  }
  void barrier() {
    std::unique_lock<std::mutex> lock(mu);
    if (aborted) throw std::runtime_error("strip mode: another rank failed");
    const long gen = generation;
    if (++arrived == world) {
      arrived = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lock, [&] { return generation != gen || aborted; });
      if (generation == gen) throw std::runtime_error("strip mode: another rank failed");
    }
  }
};

class ThreadComm : public Comm {
 public:
  ThreadComm(ThreadGroup* g, int rank) : g_(g), rank_(rank) {}
  int rank() const override { return rank_; }
  int world() const override { return g_->world; }
  void allgather_inplace(void* dev_buf, size_t elem_bytes, const std::vector<size_t>& offset,
                         const std::vector<size_t>& count, Stream s) override {
    stream_sync(s);  // my segment is complete
    g_->bufs[rank_] = dev_buf;
    g_->barrier();
    for (int r = 0; r < g_->world; ++r) {
      if (r == rank_ || count[r] == 0) continue;
      const size_t off = offset[r] * elem_bytes;
      d2d(static_cast<char*>(dev_buf) + off, static_cast<const char*>(g_->bufs[r]) + off, count[r] * elem_bytes, s);
    }
    stream_sync(s);
    g_->barrier();  // nobody overwrites a buffer that is still being read
  }

 private:
  ThreadGroup* g_;
  int rank_;
};

ThreadGroup* thread_group_create(int world) { return new ThreadGroup(world); }
void thread_group_abort(ThreadGroup* g) { g->abort(); }
void thread_group_destroy(ThreadGroup* g) { delete g; }
Comm* thread_comm_create(ThreadGroup* g, int rank) { return new ThreadComm(g, rank); }

#if !defined(GB200_HOSTSIM)
// ---------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId {
  char internal[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) return;
    api.GetUniqueId = reinterpret_cast<int (*)(ncclUniqueId*)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank =
        reinterpret_cast<int (*)(ncclComm_t*, int, ncclUniqueId, int)>(dlsym(api.lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<int (*)(ncclComm_t)>(dlsym(api.lib, "ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupEnd"));
    api.Broadcast = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, ncclComm_t, void*)>(
        dlsym(api.lib, "ncclBroadcast"));
    api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
  });
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.Broadcast)
    throw std::runtime_error("guetzli_b200: libnccl.so.2 not available");
  return api;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    const char* msg = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
    throw std::runtime_error(std::string("NCCL error in ") + what + ": " + msg);
  }
}

class NcclComm : public Comm {
 public:
  NcclComm(const uint8_t id[128], int rank, int world) : rank_(rank), world_(world) {
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    nccl_check(nccl().CommInitRank(&comm_, world, uid, rank), "ncclCommInitRank");
    // NCCL sets up its channels lazily at the first collective (seconds): pay that here,
    // not inside the first image
    std::vector<size_t> off(world), cnt(world, 1);
    for (int r = 0; r < world; ++r) off[r] = r;
    void* tmp = dev_alloc(static_cast<size_t>(world) * 4);
    dev_zero(tmp, static_cast<size_t>(world) * 4, nullptr);
    allgather_inplace(tmp, 4, off, cnt, nullptr);
    stream_sync(nullptr);
    dev_free(tmp);
  }
  ~NcclComm() override {
    if (comm_) nccl().CommDestroy(comm_);
  }
  int rank() const override { return rank_; }
  int world() const override { return world_; }
  void allgather_inplace(void* dev_buf, size_t elem_bytes, const std::vector<size_t>& offset,
                         const std::vector<size_t>& count, Stream s) override {
    // uneven segments: one in-place broadcast per owner, fused into one NCCL group
    nccl_check(nccl().GroupStart(), "ncclGroupStart");
    for (int r = 0; r < world_; ++r) {
      if (count[r] == 0) continue;
      char* p = static_cast<char*>(dev_buf) + offset[r] * elem_bytes;
      nccl_check(nccl().Broadcast(p, p, count[r] * elem_bytes, /*ncclInt8*/ 0, r, comm_, s), "ncclBroadcast");
    }
    nccl_check(nccl().GroupEnd(), "ncclGroupEnd");
  }

 private:
  ncclComm_t comm_ = nullptr;
  int rank_, world_;
};
}  // namespace

void nccl_unique_id(uint8_t out[128]) {
  ncclUniqueId uid;
  nccl_check(nccl().GetUniqueId(&uid), "ncclGetUniqueId");
  memcpy(out, uid.internal, 128);
}
Comm* nccl_comm_create(const uint8_t id[128], int rank, int world) { return new NcclComm(id, rank, world); }
#endif

}  // namespace gb200
