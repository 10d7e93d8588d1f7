// fake_nccl.cpp -- TEST INFRASTRUCTURE: the ten NCCL entry points the engine resolves with dlsym, implemented
// over POSIX shared memory between PROCESSES OF ONE HOST, so that the multi-rank logic of the product (slab
// ownership, halo exchange, migration, the distributed rebuild criterion, the collective downloads) can run on
// the CPU together with the kernel emulator (tests/cusim/cusim.h). Everything is synchronous; the stream is
// ignored. A collective that not every rank enters -- the class of bug that hangs a real multi-GPU job -- is
// reported after FAKE_NCCL_TIMEOUT seconds (default 60) with the rank and the operation, and the process aborts.
//
// Built by tests/cusim/build.py into tests/_build/libnccl_fake.so; the engine loads it only when the
// environment variable EMDEE_NCCL_LIB points at it.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "include/nccl.h"

typedef void* cudaStream_t;

namespace {

constexpr size_t SLOT_BYTES = 8u << 20;   // per-rank staging area for collectives
constexpr size_t MAIL_BYTES = 1u << 20;   // per ordered rank pair, for send/recv
constexpr int MAX_RANKS = 16;

struct Mail {
  std::atomic<int> full;
  size_t bytes;
};

struct Header {
  std::atomic<int> attached;
  std::atomic<unsigned> bar_count;
  std::atomic<unsigned> bar_gen;
  int nranks;
};

struct Comm {
  int rank = 0, nranks = 1;
  Header* hdr = nullptr;
  unsigned char* base = nullptr;
  size_t total = 0;
  char name[128];
  unsigned char* slot(int r) const { return base + 4096 + (size_t)r * SLOT_BYTES; }
  Mail* mail(int src, int dst) const {
    return reinterpret_cast<Mail*>(base + 4096 + (size_t)nranks * SLOT_BYTES + ((size_t)src * nranks + dst) * (MAIL_BYTES + 64));
  }
  unsigned char* mail_data(int src, int dst) const { return reinterpret_cast<unsigned char*>(mail(src, dst)) + 64; }
};

double timeout_seconds() {
  const char* e = std::getenv("FAKE_NCCL_TIMEOUT");
  return e ? std::atof(e) : 60.0;
}

struct Deadline {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  const Comm* c;
  const char* what;
  Deadline(const Comm* c_, const char* w) : c(c_), what(w) {}
  void spin() {
    sched_yield();
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_seconds()) {
      std::fprintf(stderr, "fake NCCL: rank %d of %d has waited %.0f s in %s -- the ranks took different paths\n", c->rank,
                   c->nranks, timeout_seconds(), what);
      std::abort();
    }
  }
};

void barrier(Comm* c, const char* what) {
  Header* h = c->hdr;
  const unsigned gen = h->bar_gen.load();
  if (h->bar_count.fetch_add(1) + 1 == (unsigned)c->nranks) {
    h->bar_count.store(0);
    h->bar_gen.fetch_add(1);
  } else {
    Deadline d(c, what);
    while (h->bar_gen.load() == gen) d.spin();
  }
}

size_t type_size(ncclDataType_t t) {
  switch ((int)t) {
    case 0: case 1: return 1;
    case 2: case 3: case 7: return 4;
    case 4: case 5: case 8: return 8;
    case 6: return 2;
    default: return 0;
  }
}

template <class T>
void reduce_into(T* out, const T* in, size_t n, ncclRedOp_t op, bool first) {
  for (size_t i = 0; i < n; ++i) {
    if (first) out[i] = in[i];
    else if (op == ncclSum) out[i] = out[i] + in[i];
    else if (op == ncclProd) out[i] = out[i] * in[i];
    else if (op == ncclMax) out[i] = in[i] > out[i] ? in[i] : out[i];
    else out[i] = in[i] < out[i] ? in[i] : out[i];
  }
}

struct P2P {
  bool send;
  unsigned char* buf;
  size_t bytes, done;
  int peer;
};
int group_depth = 0;
std::vector<P2P> pending;
Comm* pending_comm = nullptr;

void run_group() {
  Comm* c = pending_comm;
  if (c == nullptr || pending.empty()) {
    pending.clear();
    return;
  }
  Deadline d(c, "a send/recv group");
  for (;;) {
    bool all_done = true, progressed = false;
    // per peer and direction, only the OLDEST unfinished operation may move (NCCL matches in order)
    bool send_busy[MAX_RANKS] = {false}, recv_busy[MAX_RANKS] = {false};
    for (P2P& op : pending) {
      if (op.done == op.bytes) continue;
      all_done = false;
      bool* busy = op.send ? send_busy : recv_busy;
      if (busy[op.peer]) continue;
      busy[op.peer] = true;
      Mail* m = op.send ? c->mail(c->rank, op.peer) : c->mail(op.peer, c->rank);
      unsigned char* data = op.send ? c->mail_data(c->rank, op.peer) : c->mail_data(op.peer, c->rank);
      if (op.send && m->full.load() == 0) {
        const size_t n = std::min(MAIL_BYTES, op.bytes - op.done);
        std::memcpy(data, op.buf + op.done, n);
        m->bytes = n;
        m->full.store(1);
        op.done += n;
        progressed = true;
      } else if (!op.send && m->full.load() == 1) {
        const size_t n = m->bytes;
        if (n > op.bytes - op.done) {
          std::fprintf(stderr, "fake NCCL: rank %d receives %zu bytes from %d but expects at most %zu\n", c->rank, n, op.peer,
                       op.bytes - op.done);
          std::abort();
        }
        std::memcpy(op.buf + op.done, data, n);
        m->full.store(0);
        op.done += n;
        progressed = true;
      }
    }
    if (all_done) break;
    if (!progressed) d.spin();
  }
  pending.clear();
  pending_comm = nullptr;
}

}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
  std::memset(id, 0, sizeof(*id));
  std::snprintf(id->internal, sizeof(id->internal), "/emdee_fakenccl_%d_%lld", (int)getpid(),
                (long long)std::chrono::steady_clock::now().time_since_epoch().count());
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* out, int nranks, ncclUniqueId id, int rank) {
  if (nranks > MAX_RANKS) return (ncclResult_t)1;
  Comm* c = new Comm();
  c->rank = rank;
  c->nranks = nranks;
  std::snprintf(c->name, sizeof(c->name), "%s", id.internal);
  c->total = 4096 + (size_t)nranks * SLOT_BYTES + (size_t)nranks * nranks * (MAIL_BYTES + 64);
  int fd = -1;
  if (rank == 0) {
    fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)c->total) != 0) return (ncclResult_t)1;
  } else {
    Deadline d(c, "ncclCommInitRank (waiting for rank 0)");
    struct stat sb;
    for (;;) {
      fd = shm_open(c->name, O_RDWR, 0600);
      if (fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size == c->total) break;
      if (fd >= 0) close(fd);
      d.spin();
    }
  }
  c->base = static_cast<unsigned char*>(mmap(nullptr, c->total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
  close(fd);
  if (c->base == MAP_FAILED) return (ncclResult_t)1;
  c->hdr = reinterpret_cast<Header*>(c->base);   // a fresh shm object is zero-filled: counters start at 0
  c->hdr->nranks = nranks;
  c->hdr->attached.fetch_add(1);
  {
    Deadline d(c, "ncclCommInitRank");
    while (c->hdr->attached.load() < nranks) d.spin();
  }
  barrier(c, "ncclCommInitRank");
  if (rank == 0) shm_unlink(c->name);   // everybody has it mapped: the name can go
  *out = reinterpret_cast<ncclComm_t>(c);
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (c != nullptr) {
    munmap(c->base, c->total);
    delete c;
  }
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t comm,
                           cudaStream_t) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  const size_t ts = type_size(type);
  const size_t per = SLOT_BYTES / ts;
  for (size_t off = 0; off < count || off == 0; off += per) {
    const size_t n = std::min(per, count - off);
    std::memcpy(c->slot(c->rank), static_cast<const unsigned char*>(send) + off * ts, n * ts);
    barrier(c, "ncclAllReduce");
    unsigned char* dst = static_cast<unsigned char*>(recv) + off * ts;
    for (int r = 0; r < c->nranks; ++r) {   // fixed rank order: every rank computes the same bits
      const unsigned char* src = c->slot(r);
      if ((int)type == 8) reduce_into(reinterpret_cast<double*>(dst), reinterpret_cast<const double*>(src), n, op, r == 0);
      else if ((int)type == 7) reduce_into(reinterpret_cast<float*>(dst), reinterpret_cast<const float*>(src), n, op, r == 0);
      else if ((int)type == 2) reduce_into(reinterpret_cast<int*>(dst), reinterpret_cast<const int*>(src), n, op, r == 0);
      else if ((int)type == 3) reduce_into(reinterpret_cast<unsigned*>(dst), reinterpret_cast<const unsigned*>(src), n, op, r == 0);
      else if ((int)type == 4) reduce_into(reinterpret_cast<long long*>(dst), reinterpret_cast<const long long*>(src), n, op, r == 0);
      else if ((int)type == 5) reduce_into(reinterpret_cast<unsigned long long*>(dst), reinterpret_cast<const unsigned long long*>(src), n, op, r == 0);
      else reduce_into(reinterpret_cast<signed char*>(dst), reinterpret_cast<const signed char*>(src), n, op, r == 0);
    }
    barrier(c, "ncclAllReduce");
    if (count == 0) break;
  }
  return ncclSuccess;
}

ncclResult_t ncclAllGather(const void* send, void* recv, size_t sendcount, ncclDataType_t type, ncclComm_t comm, cudaStream_t) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  const size_t bytes = sendcount * type_size(type);
  if (bytes > SLOT_BYTES) {
    std::fprintf(stderr, "fake NCCL: all-gather of %zu bytes per rank exceeds the staging slot\n", bytes);
    std::abort();
  }
  std::memcpy(c->slot(c->rank), send, bytes);   // `send` may alias a part of `recv`: staged first
  barrier(c, "ncclAllGather");
  for (int r = 0; r < c->nranks; ++r) std::memcpy(static_cast<unsigned char*>(recv) + (size_t)r * bytes, c->slot(r), bytes);
  barrier(c, "ncclAllGather");
  return ncclSuccess;
}

ncclResult_t ncclGroupStart() {
  ++group_depth;
  return ncclSuccess;
}

ncclResult_t ncclGroupEnd() {
  if (--group_depth == 0) run_group();
  return ncclSuccess;
}

ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
  pending_comm = reinterpret_cast<Comm*>(comm);
  pending.push_back(P2P{true, const_cast<unsigned char*>(static_cast<const unsigned char*>(buf)), count * type_size(type), 0, peer});
  if (group_depth == 0) run_group();
  return ncclSuccess;
}

ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t) {
  pending_comm = reinterpret_cast<Comm*>(comm);
  pending.push_back(P2P{false, static_cast<unsigned char*>(buf), count * type_size(type), 0, peer});
  if (group_depth == 0) run_group();
  return ncclSuccess;
}

const char* ncclGetErrorString(ncclResult_t) { return "fake NCCL error"; }

}  // extern "C"
